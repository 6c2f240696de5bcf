"""Times the host-buffer entry (qz_render) against the device-buffer entry (qz_render_device), call by call."""
import ctypes, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from quetzalcoatlus_b200 import load_harness
from quetzalcoatlus_b200.harness import QzRenderOptions, QzStats

qz = load_harness()
lib = qz.lib
sc = qz.build_scene("cornell_box", 800, 800)
handle, cam = ctypes.c_void_p(sc.c_scene_handle()), sc.c_camera()
film = torch.zeros((3, 800, 800, 3), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream()
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for rep in range(4):
    st = QzStats(); opts = QzRenderOptions(0, 0, 0, 0)
    t0 = time.perf_counter()
    lib.qz_render_device(handle, ctypes.byref(cam), spp, 64, None, ctypes.byref(opts), ctypes.c_void_p(film[0].data_ptr()),
                         ctypes.c_void_p(film[1].data_ptr()), ctypes.c_void_p(film[2].data_ptr()), ctypes.c_void_p(stream.cuda_stream), ctypes.byref(st))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"device entry: wall {1e3*(t1-t0):8.2f} ms   device {st.ms_total:8.2f} ms  iterations {st.iterations}")
for rep in range(4):
    color = np.zeros((800, 800, 3), np.float32); normal = np.zeros_like(color); albedo = np.zeros_like(color)
    st = QzStats(); opts = QzRenderOptions(0, 0, 0, 0)
    t0 = time.perf_counter()
    lib.qz_render(handle, ctypes.byref(cam), spp, 64, None, ctypes.byref(opts), color.ctypes.data_as(ctypes.c_void_p),
                  normal.ctypes.data_as(ctypes.c_void_p), albedo.ctypes.data_as(ctypes.c_void_p), ctypes.byref(st))
    t1 = time.perf_counter()
    print(f"host entry:   wall {1e3*(t1-t0):8.2f} ms   device {st.ms_total:8.2f} ms  iterations {st.iterations}")
