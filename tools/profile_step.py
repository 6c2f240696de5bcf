"""One render of a bench workload and nothing else: the command profiled under ncu (bench.py runs
warm-ups, end-to-end steps and instrumented steps, which under ncu's per-kernel replay takes tens of minutes).

    python tools/profile_step.py --workload cornell_box --spp 16 [--flags N] [--pool N]
"""
import argparse, ctypes, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from quetzalcoatlus_b200 import load_harness
from quetzalcoatlus_b200.harness import QzRenderOptions, QzStats

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cornell_box")
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--pool", type=int, default=0)
ap.add_argument("--mesh-triangles", type=int, default=1_000_000)
a = ap.parse_args()
qz = load_harness()
_, w, h, _, mb = bench.WORKLOADS[a.workload]
sc = bench.build_scene(qz, a.workload, w, h, a.mesh_triangles)
handle, cam = ctypes.c_void_p(sc.c_scene_handle()), sc.c_camera()
film = torch.zeros((3, h, w, 3), dtype=torch.float32, device="cuda")
st, opts = QzStats(), QzRenderOptions(a.flags, a.pool, 0, 0)
rc = qz.lib.qz_render_device(handle, ctypes.byref(cam), a.spp, mb, None, ctypes.byref(opts), ctypes.c_void_p(film[0].data_ptr()),
                             ctypes.c_void_p(film[1].data_ptr()), ctypes.c_void_p(film[2].data_ptr()), None, ctypes.byref(st))
torch.cuda.synchronize()
d = st.as_dict()
print({k: d[k] for k in ("paths", "rays_closest", "rays_shadow", "shade_calls", "iterations", "kernel_launches", "ms_total")}, "rc", rc)
