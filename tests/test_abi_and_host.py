"""CPU-only checks of the boundary: the C-ABI library loads without a GPU and exports every
symbol include/qz_b200.h declares; without a device it fails LOUDLY (no CPU fallback); the
reference's own example programs and colour test compile UNCHANGED against the host headers."""
import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "quetzalcoatlus_b200"
HEADER = ROOT / "include" / "qz_b200.h"
CUDA_LIB = PKG / "_lib" / "libqz_b200.so"
HARNESS_LIB = PKG / "_lib" / "libqz_harness.so"
REFERENCE = Path("/root/reference")


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(qz_[a-z_0-9]+)\s*\(", text)))


def has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ["qz_init", "qz_scene_create", "qz_scene_commit", "qz_scene_destroy", "qz_render", "qz_render_device",
                 "qz_trace_paths", "qz_sampler_eval", "qz_intersect", "qz_last_error"]:
        assert must in syms


def test_cuda_library_exports_every_declared_symbol():
    if not CUDA_LIB.exists():
        pytest.fail(f"{CUDA_LIB} missing: run __graft_entry__.build()")
    lib = ctypes.CDLL(str(CUDA_LIB))
    for name in declared_symbols():
        assert hasattr(lib, name), f"libqz_b200.so does not export {name}"
    assert lib.qz_abi_version() == 1


def test_pod_struct_sizes_match_the_python_bindings():
    from quetzalcoatlus_b200.harness import QzCamera, QzRegion, QzRenderOptions, QzStats

    src = '#include <stdio.h>\n#include "qz_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          "sizeof(qz_spectrum),sizeof(qz_texture),sizeof(qz_material),sizeof(qz_light),sizeof(qz_geometry),sizeof(qz_prim)," \
          "sizeof(qz_camera),sizeof(qz_region),sizeof(qz_render_options),sizeof(qz_stats));return 0;}"
    exe = Path("/tmp/qz_sizes")
    subprocess.run(["gcc", "-x", "c", "-", f"-I{ROOT / 'include'}", "-o", str(exe)], input=src, text=True, check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes[:6] == [32, 32, 32, 80, 32, 64]
    assert sizes[6:] == [ctypes.sizeof(QzCamera), ctypes.sizeof(QzRegion), ctypes.sizeof(QzRenderOptions), ctypes.sizeof(QzStats)]


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_a_cpu_fallback():
    lib = ctypes.CDLL(str(CUDA_LIB))
    lib.qz_last_error.restype = ctypes.c_char_p
    assert lib.qz_init(0) == 1  # QZ_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.qz_last_error()
    scene = ctypes.c_void_p()
    assert lib.qz_scene_create(ctypes.byref(scene)) == 1
    out = np.zeros(4, np.float32)
    q = np.zeros(4, np.int32)
    assert lib.qz_sampler_eval(1, 8, 8, 1, q.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)) == 1
    # the host library reports it the way the reference reports a failed device and leaves the scene not ready
    from quetzalcoatlus_b200 import load_harness

    qz = load_harness()
    with pytest.raises(RuntimeError):
        qz.build_scene("cornell_box", 8, 8)


def test_product_libraries_do_not_link_the_oracle_or_the_emulation():
    for lib in (CUDA_LIB, HARNESS_LIB):
        out = subprocess.run(["ldd", str(lib)], capture_output=True, text=True).stdout
        assert "oracle" not in out and "emu" not in out, out
    for py in list(PKG.glob("*.py")) + [ROOT / "quetzalcoatlus_b200" / "Makefile"]:
        text = py.read_text()
        assert "liboracle_ref" not in text and "libqz_emu" not in text, py


@pytest.mark.skipif(not REFERENCE.exists(), reason="/root/reference not present")
@pytest.mark.parametrize("example", ["cornell_box", "glass_spheres", "textures", "opposing_planes", "mandelbrot"])
def test_reference_examples_compile_unchanged_against_the_host_headers(example, tmp_path):
    """Drop-in at the source level: the reference's example mains, untouched, against our headers."""
    src = REFERENCE / "examples" / f"{example}.cpp"
    cmd = ["g++", "-std=c++20", "-O0", "-w", "-fsyntax-only", f"-I{PKG / 'host'}", f"-I{ROOT / 'include'}", str(src)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-3000:]


@pytest.mark.skipif(not REFERENCE.exists(), reason="/root/reference not present")
def test_reference_color_test_against_the_host_library(tmp_path):
    """The reference's colour smoke test, compiled unchanged against the host colour classes, prints
    the values the reference prints."""
    exe = tmp_path / "color_test_host"
    # compile a byte-identical copy placed outside the reference tree, so that its quoted
    # #include "color.hpp" resolves to OUR headers instead of the file's own directory
    src = tmp_path / "color_test.cpp"
    src.write_bytes((REFERENCE / "src" / "color" / "color_test.cpp").read_bytes())
    cmd = ["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-w", f"-I{PKG / 'host'}", f"-I{PKG / 'host' / 'color'}",
           f"-I{ROOT / 'include'}", str(src), str(PKG / "host" / "src" / "color.cpp"),
           "-ldl", "-o", str(exe)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-3000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True, env={"QZ_DATA_DIR": str(PKG / "data")}).stdout
    for want in ["whitepoint: 0.3367 0.357918", "0.333314 0.333288", "0.771653 0.470475 0.980432", "0.769878 0.544533 0.671843",
                 "0.446457 0.594196 0.175637", "0.758855 0.589091 1.49932", "0.656894 0.511026 0.176181", "0.609913 0.599273 1.37064"]:
        assert want in out, (want, out)


def test_obj_loader_semantics(emu, oracle, tmp_path):
    """Face forms, triangle-as-degenerate-quad, comments, extra corners, bad lines: same paths as
    the reference's regex loader produces."""
    import sys

    sys.path.insert(0, str(ROOT / "tests"))
    from common import bits_equal, pixel_samples

    verts = "v -1 0.2 -0.5\nv 1 0.2 -0.5\nv 1 2.2 -0.8 1.0\nv -1 2.2 -0.8\nv 0 3.0 -1\n"
    plain = tmp_path / "plain.obj"
    plain.write_text("# comment\n" + verts + "vt 0 0\n\ng grp\nf 1 2 3\nf 1/1 3/1 4/1\nf 4 3 5 5 1\nf 1 2\nx junk\n")
    smooth = tmp_path / "smooth.obj"
    smooth.write_text(verts + "vn 0 0 1\nvn 0.1 0 1\nvn 0 0.1 1\n"
                      "f 1//1 2//2 3//3\nf 1/1/1 3/1/3 4/1/2\nf 4//1 3//2 5//3 5//3\n")
    for path in (plain, smooth):
        kw = dict(obj_path=str(path), obj_material="diffuse", obj_light="point")
        with emu.build_scene("obj_viewer", 32, 24, **kw) as se, oracle.build_scene("obj_viewer", 32, 24, **kw) as so:
            xys = pixel_samples(so, 600, seed=2)
            assert bits_equal(se.trace_paths(xys), so.trace_paths(xys)).all(), path.name


def test_strip_rows_are_balanced_and_keep_a_shard_on_its_own_pixel_classes():
    """bench.py / qz_render's multi-GPU strips: every rank owns the same number of rows, and where a power-of-two height
    does that, strip * ranks divides 128 -- a rank then owns 1/ranks of the (y mod 128) pixel classes, which is what its
    sample memo tabulates (csrc/sampler.cuh).  5-row strips over 8 ranks were balanced too, but touched all 128."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for height in (800, 600, 1080, 2160, 96, 54):
        for world in (1, 2, 4, 8):
            rows = bench.strip_rows_for(height, world)
            owned = [sum(1 for r in range(height) if (r // rows) % world == k) for k in range(world)]
            if height % world == 0:
                assert max(owned) == min(owned), (height, world, rows, owned)
            if height % (world * rows) == 0 and rows in (1, 2, 4, 8):
                classes = {(height - 1 - r) % 128 for r in range(height) if (r // rows) % world == 0}
                assert len(classes) <= 128 // world, (height, world, rows, len(classes))
    assert bench.strip_rows_for(800, 8) == 4 and bench.strip_rows_for(2160, 8) == 2 and bench.strip_rows_for(800, 4) == 8


def test_reference_arm_of_the_bench_runs_on_the_cpu(tmp_path):
    """`bench.py --impl reference` (the reference integrator over the oracle, host cores only) prints ONE JSON line with the
    keys the driver reads, on a box without a GPU."""
    import json
    import sys

    lib = ROOT / "oracle" / "_ref" / "liboracle_ref.so"
    if not lib.exists():
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--width", "48",
                          "--height", "40"], capture_output=True, text=True, timeout=300, cwd=tmp_path)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mpaths/s" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
