"""Multi-GPU BEHIND render(): one process, the scene replicated on N devices (qz_set_device_count), the host library's
render() timed end to end (RenderResult on the host).  Weak scaling like bench.py: 128 spp per GPU.

    python tools/inproc_multi_gpu.py [--workload cornell_box] [--devices 1,2,4,8] [--steps 3]
"""
import argparse, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from quetzalcoatlus_b200 import load_harness

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cornell_box")
ap.add_argument("--devices", default="1,2,4,8")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--spp", type=int, default=0)
a = ap.parse_args()
qz = load_harness()
_, w, h, spp0, mb = bench.WORKLOADS[a.workload]
for n in [int(x) for x in a.devices.split(",")]:
    if n > torch.cuda.device_count():
        continue
    qz.lib.qz_set_device_count(n)
    sc = bench.build_scene(qz, a.workload, w, h, 1_000_000)
    spp = (a.spp or spp0) * n
    with bench.quiet_stdout():
        for _ in range(2):
            sc.render_only(spp, mb)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            sc.render_only(spp, mb)
        dt = (time.perf_counter() - t0) / a.steps
    st = sc.last_stats()
    print(json.dumps({"in_process_devices": n, "workload": a.workload, "width": w, "height": h, "spp": spp, "ms_per_render": dt * 1e3,
                      "mpaths_per_s_e2e": w * h * spp / dt / 1e6, "device_ms_max": st["ms_total"], "paths": st["paths"],
                      "call": "render(camera, scene, spp, bounces) -> RenderResult, scene replicated by qz_set_device_count"}))
    sc.close()
    qz.lib.qz_set_device_count(0)
