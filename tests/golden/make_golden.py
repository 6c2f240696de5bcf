#!/usr/bin/env python3
"""Generates the committed golden fixtures FROM THE ORACLE (oracle/_ref, i.e. the reference's
own sources over the Embree shim).  Run in the development container, where /root/reference
exists:  python tests/golden/make_golden.py

paths_<scene>.npz : xys (n x 3 int32: x, y(sampler), s) and records (n x 32 float32)
sampler_kat.npz   : q (n x 4 int32: x, y, s, dim) and values (float32), incl. the survey's probe
film_<scene>.npz  : a low-resolution, low-spp film (color, normal, albedo)
"""
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from common import ANALYTIC_SCENES, pixel_samples  # noqa: E402
from quetzalcoatlus_b200.harness import Harness  # noqa: E402

OUT = Path(__file__).resolve().parent


def obj_viewer_1m(orc) -> None:
    """One-off (the reference's regex OBJ loader needs ~25 minutes for the 1M-triangle text): replayed paths of the
    obj_viewer configuration on the FULL synthetic mesh, plus the checksum of the OBJ text they belong to."""
    import hashlib

    sys.path.insert(0, str(ROOT / "tools"))
    import gen_mesh

    path = gen_mesh.ensure_obj("/tmp", 1_000_000)
    digest = hashlib.sha256(open(path, "rb").read()).hexdigest()
    with orc.build_scene("obj_viewer", obj_path=path, obj_material="alluminum", obj_light="point") as sc:
        xys = pixel_samples(sc, 4096, seed=1_000_003)
        rec = sc.trace_paths(xys)
        r = sc.render(1)
    np.savez_compressed(OUT / "paths_obj_viewer_1m.npz", xys=xys, records=rec, obj_sha256=np.array(digest),
                        cpu_seconds_1spp=np.array(r.seconds), cpu_threads=np.array(r.n_threads), cpu_rays_1spp=np.array(r.rays))
    print("paths obj_viewer_1m", digest, r.seconds)


def main() -> None:
    if "--obj1m" in sys.argv:
        obj_viewer_1m(Harness(ROOT / "oracle" / "_ref" / "liboracle_ref.so", "orc_"))
        return
    orc = Harness(ROOT / "oracle" / "_ref" / "liboracle_ref.so", "orc_")
    rng = np.random.default_rng(2026)
    q = np.stack([rng.integers(0, 1920, 4096), rng.integers(0, 1080, 4096), rng.integers(0, 256, 4096), rng.integers(0, 1000, 4096)], 1)
    q[:6] = [[10, 20, 1, d] for d in range(6)]  # SURVEY.md section 4 probe
    q = q.astype(np.int32)
    np.savez_compressed(OUT / "sampler_kat.npz", q=q, values=orc.sampler_eval(256, 1920, 1080, q), res=np.array([256, 1920, 1080]))
    q2 = q.copy()
    q2[:, 0] %= 800
    q2[:, 1] %= 800
    q2[:, 2] %= 4
    np.savez_compressed(OUT / "sampler_kat_800.npz", q=q2, values=orc.sampler_eval(4, 800, 800, q2), res=np.array([4, 800, 800]))
    for name in ANALYTIC_SCENES:
        with orc.build_scene(name) as sc:
            xys = pixel_samples(sc, 512, seed=zlib.crc32(name.encode()) % 1000 + 7)
            np.savez_compressed(OUT / f"paths_{name}.npz", xys=xys, records=sc.trace_paths(xys))
        print("paths", name)
    for name, (w, h, spp) in {"cornell_box": (48, 48, 4), "textures": (40, 40, 3), "kitchen_sink": (40, 30, 4)}.items():
        with orc.build_scene(name, w, h) as sc:
            r = sc.render(spp)
            np.savez_compressed(OUT / f"film_{name}.npz", color=r.color, normal=r.normal, albedo=r.albedo, spp=spp)
        print("film", name)


if __name__ == "__main__":
    main()
