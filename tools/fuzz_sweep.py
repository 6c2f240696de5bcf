"""Offline sweep of the seeded random parity scenes (scenes/scenes.hpp: build_fuzz) and random meshes through the
device-code emulation (tests/emu) against the oracle, on seed ranges the test suite does not hold.

TEST INFRASTRUCTURE, CPU only: loads oracle/_ref and tests/emu/_build, never the product library.

    python tools/fuzz_sweep.py --exact 1000:3000 --fast 5000:5600 --mesh 1000:1400 --jobs 8 --out profiles/r02_fuzz_sweep.json

exact: every field of every replayed path bit for bit (tests/test_emu_parity.py::test_random_scenes_bit_exact)
fast : the radiometric build: discrete decisions identical, finite radiance within the reported relative deviation
film : whole films (pixel loop, ordered accumulation, division) bit for bit, three planes
obj  : OBJ text loader on torture files against the reference's own regex loader, record by record
sampler: the Halton sampler at random resolutions / sample counts, bit for bit
memo : the sample memo's host side at random resolutions / row lengths / pass sizes / shards: film with == film without
rays : closest-hit probe with axis-aligned, vertex-aimed, edge-grazing, tiny and huge rays on random meshes
mesh : OBJ loader -> LBVH -> wide-BVH traversal: closest hits and paths bit for bit
"""
from __future__ import annotations

import argparse
import json
import sys
import tempfile
import time
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

N_PATHS = 1200


def _libs():
    from quetzalcoatlus_b200.harness import Harness

    build = ROOT / "tests" / "emu" / "_build"
    return (Harness(ROOT / "oracle" / "_ref" / "liboracle_ref.so", "orc_"), Harness(build / "libqz_emu_harness.so", "qzh_"),
            Harness(build / "libqz_emu_fast_harness.so", "qzh_"))


def run_exact(seed: int) -> dict:
    from common import bits_equal, pixel_samples

    oracle, emu, _ = _libs()
    name = f"fuzz:{seed}"
    with emu.build_scene(name) as se, oracle.build_scene(name) as so:
        cam = bool(bits_equal(se.camera_fields(), so.camera_fields()).all())
        xys = pixel_samples(so, N_PATHS, seed=seed)
        got, want = se.trace_paths(xys), so.trace_paths(xys)
    bad = ~bits_equal(got, want).all(1)
    return {"kind": "exact", "seed": seed, "paths": len(bad), "differing": int(bad.sum()) + (0 if cam else len(bad))}


def run_fast(seed: int) -> dict:
    from common import NORMAL, RADIANCE, RAYS, bits_equal, pixel_samples

    oracle, _, fast = _libs()
    name = f"fuzz:{seed}"
    with fast.build_scene(name) as sf, oracle.build_scene(name) as so:
        xys = pixel_samples(so, N_PATHS, seed=seed)
        got, want = sf.trace_paths(xys), so.trace_paths(xys)
    discrete = (got[:, RAYS] == want[:, RAYS]) & bits_equal(got[:, :4], want[:, :4]).all(1) & bits_equal(got[:, NORMAL], want[:, NORMAL]).all(1)
    g, w = got[:, RADIANCE].astype(np.float64), want[:, RADIANCE].astype(np.float64)
    nonfinite_same = bool((np.isfinite(g) == np.isfinite(w)).all() and (g[~np.isfinite(w)] == w[~np.isfinite(w)]).all())
    finite = np.isfinite(w).all(1) & np.isfinite(g).all(1)
    worst, over5, over4 = 0.0, 0, 0
    if finite.any():
        scale = max(np.abs(w[finite]).max(), 1e-6)
        rel = (np.abs(g[finite] - w[finite]) / np.maximum(np.abs(w[finite]), 1e-2 * scale)).max(1)
        worst, over5, over4 = float(rel.max()), int((rel > 1e-5).sum()), int((rel > 1e-4).sum())
    return {"kind": "fast", "seed": seed, "paths": len(discrete), "discrete_differing": int((~discrete).sum()),
            "nonfinite_same": nonfinite_same, "worst_rel": worst, "over_1e-5": over5, "over_1e-4": over4}


def run_mesh(seed: int) -> dict:
    from common import bits_equal, pixel_samples
    from test_emu_parity import _random_mesh_obj

    oracle, emu, _ = _libs()
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "mesh.obj"
        rng = _random_mesh_obj(path, seed)
        kw = dict(obj_path=str(path), obj_material=["diffuse", "glass", "alluminum", "copper"][seed % 4],
                  obj_light=["point", "area", "ambient"][seed % 3])
        with emu.build_scene("obj_viewer", 48, 36, **kw) as se, oracle.build_scene("obj_viewer", 48, 36, **kw) as so:
            o = np.tile(so.camera_fields()[0], (3000, 1)) + rng.normal(0, 0.4, (3000, 3))
            d = rng.normal(0, 1, (3000, 3))
            d[:, 2] -= 1.5
            rays = np.concatenate([o, d], 1).astype(np.float32)
            hits = int((~bits_equal(se.intersect(rays), so.intersect(rays)).all(1)).sum())
            xys = pixel_samples(so, 1000, seed=seed)
            paths = int((~bits_equal(se.trace_paths(xys), so.trace_paths(xys)).all(1)).sum())
    return {"kind": "mesh", "seed": seed, "rays": 3000, "paths": 1000, "hits_differing": hits, "paths_differing": paths}


def run_film(seed: int) -> dict:
    """Pixel loop, ordered accumulation and division (render.cpp:260-294): three planes.  (The emulation has no wavefront;
    the wavefront film is compared with the replayed paths by the -m gpu tests.)"""
    from common import bits_equal

    oracle, emu, _ = _libs()
    w, h, spp = 20 + seed % 13, 12 + seed % 11, 2 + seed % 3
    with emu.build_scene(f"fuzz:{seed}", w, h) as se, oracle.build_scene(f"fuzz:{seed}", w, h) as so:
        a, b = se.render(spp), so.render(spp)
    bad = sum(int((~bits_equal(getattr(a, p), getattr(b, p))).sum()) for p in ("color", "normal", "albedo"))
    return {"kind": "film", "seed": seed, "values": 9 * w * h, "differing": bad}


def run_obj(seed: int) -> dict:
    """OBJ text loader on a torture file (tests/test_abi_and_host.py: _torture_obj_lines) against the reference's own regex
    loader compiled into the oracle: ObjData record by record."""
    from test_abi_and_host import _torture_obj_lines

    oracle, emu, _ = _libs()
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "torture.obj"
        path.write_bytes("\n".join(_torture_obj_lines(np.random.default_rng(seed))).encode())
        got, want = emu.obj_load(path), oracle.obj_load(path)
    bad = 0
    for key in ("vertices", "normals"):
        if got[key].shape != want[key].shape:
            bad += 1
        else:
            bad += int((got[key].view(np.uint32) != want[key].view(np.uint32)).sum())
    gf, wf = got["faces"], want["faces"]
    if gf.shape != wf.shape:
        bad += 1
    else:
        defined = (gf[:, 8:12] == 0).all(1) | (gf[:, 4:8] != 0).any(1)
        bad += int((gf[:, :4] != wf[:, :4]).sum() + (gf[:, 8:] != wf[:, 8:]).sum() + (gf[defined, 4:8] != wf[defined, 4:8]).sum())
    return {"kind": "obj", "seed": seed, "records": int(len(want["vertices"]) + len(want["normals"]) + len(wf)), "differing": bad}


def run_sampler(seed: int) -> dict:
    """Owen-scrambled Halton sampler (sampler.cpp:161-454) at a random resolution and sample count: (x, y, s, dimension) ->
    float, bit for bit; dimensions across the whole 1000-prime table."""
    from common import bits_equal

    oracle, emu, _ = _libs()
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(1, 8193)), int(rng.integers(1, 8193))
    spp = int(rng.choice([1, 2, 3, 7, 24, 36, 128, 1000, 1024, 4096, int(rng.integers(1, 69000))]))   # (beyond 69042 x 31104 the reference's int index overflows: refused by qz_render)
    n = 2000
    q = np.stack([rng.integers(0, w, n), rng.integers(0, h, n), rng.integers(0, spp, n), rng.integers(0, 1000, n)], 1)
    bad = int((~bits_equal(emu.sampler_eval(spp, w, h, q), oracle.sampler_eval(spp, w, h, q))).sum())
    return {"kind": "sampler", "seed": seed, "values": n, "differing": bad}


def run_memo(seed: int) -> dict:
    """Sample memo on the host (csrc/memo_plan.h, sampler.cuh: owned pixel classes, row layout, fill and read side shared
    with the kernels) at a random resolution, sample count, row length, pass size and multi-GPU shard: the owned rows
    of the film with the table forced on must equal the film without it, bit for bit."""
    from common import bits_equal
    from quetzalcoatlus_b200.harness import QZ_FLAG_FORCE_MEMO

    _, emu, _ = _libs()
    rng = np.random.default_rng(seed)
    w, h, spp = int(rng.integers(1, 260)), int(rng.integers(1, 260)), int(rng.integers(1, 4))
    name = str(rng.choice(["cornell_box", "textures", "kitchen_sink", f"fuzz:{seed}"]))
    region = None
    if rng.random() < 0.6:
        n = int(rng.integers(1, 9))
        region = (int(rng.integers(1, 17)), n, int(rng.integers(0, n)))
    reserved = int(rng.choice([0, 1, 2, 5, 8]))
    per_pass = int(rng.choice([0, 1, 2]))
    with emu.build_scene(name, w, h) as sc:
        plain, _ = sc.render_flags(spp, region=region)
        memo, st = sc.render_flags(spp, flags=QZ_FLAG_FORCE_MEMO, region=region, reserved=reserved, samples_per_pass=per_pass)
    bad = sum(int((~bits_equal(getattr(plain, p), getattr(memo, p))).sum()) for p in ("color", "normal", "albedo"))
    return {"kind": "memo", "seed": seed, "values": 9 * w * h, "differing": bad, "classes": int(st["iterations"])}


def run_rays(seed: int) -> dict:
    """Closest-hit probe on a random mesh with the rays a BVH finds hardest: one or two direction components exactly zero
    (or -0), rays aimed exactly at a vertex along an axis, rays along the line through two vertices (grazing edges),
    directions scaled by 1e-20 and 1e18.  (t, u, v, Ng, ids) bit for bit against the oracle."""
    from common import bits_equal
    from test_emu_parity import _random_mesh_obj

    oracle, emu, _ = _libs()
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "mesh.obj"
        rng = _random_mesh_obj(path, seed)
        verts = np.array([[float(x) for x in line.split()[1:4]] for line in open(path) if line.startswith("v ")], dtype=np.float32)
        kw = dict(obj_path=str(path), obj_material="diffuse", obj_light="point")
        with emu.build_scene("obj_viewer", 48, 36, **kw) as se, oracle.build_scene("obj_viewer", 48, 36, **kw) as so:
            n = 2000
            o = rng.normal(0, 2.0, (n, 3)).astype(np.float32)
            d = rng.normal(0, 1, (n, 3)).astype(np.float32)
            kind = rng.integers(0, 8, n)
            for i in range(n):
                if kind[i] == 0:
                    d[i, rng.integers(0, 3)] = 0
                elif kind[i] == 1:
                    d[i, rng.choice(3, 2, replace=False)] = 0
                elif kind[i] == 2:
                    v, ax = verts[rng.integers(0, len(verts))], rng.integers(0, 3)
                    o[i] = v
                    o[i, ax] += np.float32(rng.choice([-3, 3]))
                    d[i] = 0
                    d[i, ax] = -np.sign(o[i, ax] - v[ax])
                elif kind[i] == 3:
                    d[i] *= np.float32(1e-20)
                elif kind[i] == 4:
                    d[i] *= np.float32(1e18)
                elif kind[i] == 5:
                    d[i, rng.integers(0, 3)] = -0.0
                elif kind[i] == 6:
                    a, b = verts[rng.integers(0, len(verts))], verts[rng.integers(0, len(verts))]
                    o[i] = a - (b - a) * np.float32(0.5)
                    d[i] = b - a
            rays = np.concatenate([o, d], 1).astype(np.float32)
            bad = int((~bits_equal(se.intersect(rays), so.intersect(rays)).all(1)).sum())
    return {"kind": "rays", "seed": seed, "rays": n, "differing": bad}


def _guard(fn, seed):
    try:
        return fn(seed)
    except Exception as e:  # a failing seed must be reported, not lost
        return {"kind": fn.__name__[4:], "seed": seed, "error": f"{type(e).__name__}: {e}"}


def _span(s: str) -> range:
    a, b = (int(v) for v in s.split(":"))
    return range(a, b)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--exact", default="")
    ap.add_argument("--fast", default="")
    ap.add_argument("--mesh", default="")
    ap.add_argument("--film", default="")
    ap.add_argument("--obj", default="")
    ap.add_argument("--sampler", default="")
    ap.add_argument("--memo", default="")
    ap.add_argument("--rays", default="")
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    jobs = [(run_exact, s) for s in (_span(a.exact) if a.exact else [])]
    jobs += [(run_fast, s) for s in (_span(a.fast) if a.fast else [])]
    jobs += [(run_mesh, s) for s in (_span(a.mesh) if a.mesh else [])]
    jobs += [(run_film, s) for s in (_span(a.film) if a.film else [])]
    jobs += [(run_obj, s) for s in (_span(a.obj) if a.obj else [])]
    jobs += [(run_sampler, s) for s in (_span(a.sampler) if a.sampler else [])]
    jobs += [(run_memo, s) for s in (_span(a.memo) if a.memo else [])]
    jobs += [(run_rays, s) for s in (_span(a.rays) if a.rays else [])]
    t0 = time.time()
    with ProcessPoolExecutor(a.jobs) as pool:
        rows = list(pool.map(_guard, *zip(*jobs), chunksize=4))
    summary = {"command": " ".join(sys.argv), "seconds": round(time.time() - t0, 1)}
    ex = [r for r in rows if r["kind"] == "exact" and "error" not in r]
    fa = [r for r in rows if r["kind"] == "fast" and "error" not in r]
    me = [r for r in rows if r["kind"] == "mesh" and "error" not in r]
    if ex:
        summary["exact"] = {"seeds": a.exact, "scenes": len(ex), "paths": sum(r["paths"] for r in ex),
                            "differing_paths": sum(r["differing"] for r in ex), "seeds_with_differences": [r["seed"] for r in ex if r["differing"]]}
    if fa:
        worst = max(fa, key=lambda r: r["worst_rel"])
        summary["fast"] = {"seeds": a.fast, "scenes": len(fa), "paths": sum(r["paths"] for r in fa),
                           "paths_with_a_different_discrete_decision": sum(r["discrete_differing"] for r in fa),
                           "seeds_with_a_different_discrete_decision": [r["seed"] for r in fa if r["discrete_differing"]],
                           "seeds_with_different_nonfinite_paths": [r["seed"] for r in fa if not r["nonfinite_same"]],
                           "worst_relative_radiance_deviation": worst["worst_rel"], "worst_seed": worst["seed"],
                           "paths_beyond_1e-5": sum(r["over_1e-5"] for r in fa), "paths_beyond_1e-4": sum(r["over_1e-4"] for r in fa),
                           "worst_without_the_worst_seed": max([r["worst_rel"] for r in fa if r is not worst], default=0.0)}
    if me:
        summary["mesh"] = {"seeds": a.mesh, "meshes": len(me), "rays": sum(r["rays"] for r in me), "paths": sum(r["paths"] for r in me),
                           "differing_hits": sum(r["hits_differing"] for r in me), "differing_paths": sum(r["paths_differing"] for r in me)}
    fi = [r for r in rows if r["kind"] == "film" and "error" not in r]
    if fi:
        summary["film"] = {"seeds": a.film, "films": len(fi), "values": sum(r["values"] for r in fi),
                           "differing_values": sum(r["differing"] for r in fi), "seeds_with_differences": [r["seed"] for r in fi if r["differing"]]}
    ob = [r for r in rows if r["kind"] == "obj" and "error" not in r]
    if ob:
        summary["obj"] = {"seeds": a.obj, "files": len(ob), "records": sum(r["records"] for r in ob),
                          "differing_values": sum(r["differing"] for r in ob), "seeds_with_differences": [r["seed"] for r in ob if r["differing"]]}
    sa = [r for r in rows if r["kind"] == "sampler" and "error" not in r]
    if sa:
        summary["sampler"] = {"seeds": a.sampler, "configurations": len(sa), "values": sum(r["values"] for r in sa),
                              "differing_values": sum(r["differing"] for r in sa), "seeds_with_differences": [r["seed"] for r in sa if r["differing"]]}
    mm = [r for r in rows if r["kind"] == "memo" and "error" not in r]
    if mm:
        summary["memo"] = {"seeds": a.memo, "configurations": len(mm), "film_values": sum(r["values"] for r in mm),
                           "differing_values": sum(r["differing"] for r in mm), "seeds_with_differences": [r["seed"] for r in mm if r["differing"]],
                           "configurations_that_built_a_table": sum(1 for r in mm if r["classes"] > 0)}
    ra = [r for r in rows if r["kind"] == "rays" and "error" not in r]
    if ra:
        summary["rays"] = {"seeds": a.rays, "meshes": len(ra), "rays": sum(r["rays"] for r in ra),
                           "differing_hits": sum(r["differing"] for r in ra), "seeds_with_differences": [r["seed"] for r in ra if r["differing"]]}
    summary["errors"] = [r for r in rows if "error" in r]
    text = json.dumps(summary, indent=1)
    print(text)
    if a.out:
        Path(a.out).write_text(text + "\n")


if __name__ == "__main__":
    main()
