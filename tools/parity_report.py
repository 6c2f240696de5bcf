"""Per-scene parity statistics of the CUDA path against the oracle (replayed pixel-samples):
prints a markdown table; run on the GPU box (needs oracle/_ref/liboracle_ref.so)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tools"))
from common import ANALYTIC_SCENES, RADIANCE, bits_equal, path_agreement, pixel_samples
from quetzalcoatlus_b200 import load_harness
from quetzalcoatlus_b200.harness import Harness
import gen_mesh

qz = load_harness()
orc = Harness(ROOT / "oracle" / "_ref" / "liboracle_ref.so", "orc_")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
mesh = "/tmp/qz_parity_mesh.obj"
pos, nrm, tris = gen_mesh.knot_mesh(4000)
gen_mesh.write_obj(mesh, pos, nrm, tris)
cases = [(s, {}) for s in ANALYTIC_SCENES] + [("obj_viewer", dict(obj_path=mesh, obj_material=m, obj_light=l))
                                               for m, l in (("alluminum", "point"), ("glass", "area"), ("diffuse", "ambient"))]
print("| scene | paths replayed | within 1e-4 rel. | bit-identical radiance | same ray count | max rel. error of agreeing paths |")
print("|---|---|---|---|---|---|")
for name, kw in cases:
    with qz.build_scene(name, **kw) as sg, orc.build_scene(name, **kw) as so:
        xys = pixel_samples(so, n, seed=11)
        got, want = sg.trace_paths(xys), so.trace_paths(xys)
    ok = path_agreement(want, got, 1e-4)
    exact = bits_equal(got[:, RADIANCE], want[:, RADIANCE]).all(1)
    rays = got[:, 15] == want[:, 15]
    scale = max(float(np.abs(want[:, RADIANCE]).max()), 1e-6)
    rel = np.abs(got[:, RADIANCE] - want[:, RADIANCE]) / np.maximum(np.abs(want[:, RADIANCE]), 1e-2 * scale)
    label = name + ("" if not kw else f" ({kw['obj_material']}, {kw['obj_light']})")
    print(f"| {label} | {len(xys)} | {ok.mean():.4%} | {exact.mean():.4%} | {rays.mean():.4%} | {rel[ok].max():.2e} |")
