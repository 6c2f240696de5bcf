"""What one rank of an N-way sharded render does, measured on ONE GPU: the workload's image with qz_region{strip_rows, N, shard 0}
and N times the samples (bench.py's weak scaling), against the unsharded render at the per-GPU sample count.  The ratio is the
compute side of the weak-scaling efficiency (the NCCL reduce of a 23 MB film is the rest).

    python tools/shard_probe.py [--workload cornell_box] [--shards 2,4,8]
"""
import argparse, ctypes, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from quetzalcoatlus_b200 import load_harness
from quetzalcoatlus_b200.harness import QzRegion, QzRenderOptions, QzStats

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cornell_box")
ap.add_argument("--shards", default="2,4,8")
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
qz = load_harness()
_, w, h, spp0, mb = bench.WORKLOADS[a.workload]
sc = bench.build_scene(qz, a.workload, w, h, 1_000_000)
handle, cam = ctypes.c_void_p(sc.c_scene_handle()), sc.c_camera()
film = torch.zeros((3, h, w, 3), dtype=torch.float32, device="cuda")
base = None
for n in [1] + [int(x) for x in a.shards.split(",")]:
    region = QzRegion(bench.strip_rows_for(h, n), n, 0)
    ms = []
    for i in range(a.steps + 2):
        st, opts = QzStats(), QzRenderOptions(0, 0, 0, 0)
        rc = qz.lib.qz_render_device(handle, ctypes.byref(cam), spp0 * n, mb, ctypes.byref(region), ctypes.byref(opts), ctypes.c_void_p(film[0].data_ptr()),
                                     ctypes.c_void_p(film[1].data_ptr()), ctypes.c_void_p(film[2].data_ptr()), None, ctypes.byref(st))
        assert rc == 0
        if i >= 2:
            ms.append(st.ms_total)
    t = sum(ms) / len(ms)
    base = base or t
    print(json.dumps({"workload": a.workload, "shards": n, "strip_rows": region.strip_rows, "spp": spp0 * n, "paths_of_this_shard": st.paths,
                      "device_ms": t, "compute_efficiency_vs_unsharded": base / t}))
