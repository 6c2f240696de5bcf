"""CPU-only checks of the boundary: the C-ABI library loads without a GPU and exports every
symbol include/qz_b200.h declares; without a device it fails LOUDLY (no CPU fallback); the
reference's own example programs and colour test compile UNCHANGED against the host headers."""
import ctypes
import os
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
def test_same_program_against_both_libraries_behaves_the_same(emu, oracle, tmp_path):
    """tests/api_behaviour.cpp -- written against the reference's public scene API only -- is compiled twice: against
    the reference's headers and linked with the reference's own objects (oracle/_ref), and against the host library's
    headers linked with the host library (over the emulation of the device code, so that it runs here).  What a user
    observes must be the same: ready() before / after commit, add_obj on a missing file, render() on a scene that was
    not committed (message + zero film of the right size), the films of two renders bit for bit, a render without
    samples (the reference's 0 / 0 film), a second camera on the same scene, the geometry records add_* hands back, a
    background light, an empty committed scene, and degenerate add_grid / add_obj inputs (a grid without cells still
    takes a geometry ID; a grid beyond 65535 per side, an empty OBJ and an OBJ without faces yield nullptr), both sensor
    presets with explicit imaging ratios, four fields of view under a rotated camera, a scaled camera transform -- and
    the state render() leaves std::cout in (fixed notation, three decimals: render.cpp:394)."""
    if not REFERENCE.exists():
        pytest.skip("/root/reference is absent")
    src = ROOT / "tests" / "api_behaviour.cpp"
    common = ["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-w", str(src)]
    ref_dir, emu_dir = ROOT / "oracle" / "_ref", ROOT / "tests" / "emu" / "_build"
    subprocess.run(common + ["-DNO_BOOST", f"-I{REFERENCE / 'src'}", f"-I{ROOT / 'oracle' / 'embree_shim'}", f"-L{ref_dir}", "-loracle_ref",
                             f"-Wl,-rpath,{ref_dir}", "-o", str(tmp_path / "api_ref")], check=True, capture_output=True, timeout=300)
    subprocess.run(common + [f"-I{ROOT / 'include'}", f"-I{ROOT / 'quetzalcoatlus_b200' / 'host'}", f"-L{emu_dir}", "-lqz_emu_harness",
                             f"-Wl,-rpath,{emu_dir}", "-o", str(tmp_path / "api_emu")], check=True, capture_output=True, timeout=300)
    env = dict(os.environ, QZ_DATA_DIR=str(ROOT / "quetzalcoatlus_b200" / "data"), API_TMP=str(tmp_path))

    def observed(exe):
        out = subprocess.run([str(tmp_path / exe)], capture_output=True, text=True, check=True, timeout=120, env=env,
                             cwd=ROOT / "quetzalcoatlus_b200" / "data").stdout
        skip = ("Rendering with", "Render time", "[")   # thread count, timing and the progress bar
        return [line for line in out.splitlines() if line.strip() and not line.startswith(skip)]

    ref, ours = observed("api_ref"), observed("api_emu")
    assert ref == ours
    assert "Scene must be committed before rendering." in ours and any(line.startswith("film 8x6 color") for line in ours)


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


def _torture_obj_lines(rng):
    """Lines the reference's loader survives (its std::stof throws on a token without a number): number formats of every
    kind, blank runs of every kind, CR line ends, indented records, foreign line types, every face form, faces with mixed,
    truncated, surplus and malformed corners."""
    def num():
        k = rng.integers(0, 9)
        sign = str(rng.choice(["", "-", "+"])) if rng.random() < 0.5 else ""
        if k == 0: return sign + str(rng.integers(0, 10 ** int(rng.integers(1, 12))))
        if k == 1: return sign + f"{abs(rng.uniform(-1, 1)) * 10.0 ** int(rng.integers(-6, 6)):.{rng.integers(0, 13)}f}"
        if k == 2: return sign + f"{rng.uniform(0, 1000):.{rng.integers(1, 8)}e}"
        if k == 3: return sign + "0" * int(rng.integers(1, 5)) + f"{rng.uniform(0, 10):.{rng.integers(0, 9)}f}"
        if k == 4: return sign + f".{rng.integers(0, 10 ** 6)}"
        if k == 5: return sign + f"{rng.integers(0, 1000)}."
        if k == 6: return sign + f"{rng.integers(0, 2 ** 24 + 50)}"   # around the largest integer a float holds exactly
        if k == 7: return sign + f"{rng.integers(16777000, 16777300) / 10.0 ** int(rng.integers(0, 11)):.{rng.integers(0, 11)}f}"
        return sign + f"{rng.uniform(0, 3):.7g}" + str(rng.choice(["", "abc", "f", "e", "#"]))

    lines = ["# torture", "", "   ", "\t", "\r", "vt 0.5 0.5", "g name", "o obj", "s off", "usemtl m", "vp 1 2 3", "vnn 1 2 3", "ff 1 2 3"]
    for _ in range(250):
        n, sep = rng.choice([3, 3, 3, 4, 5]), str(rng.choice([" ", "  ", "\t", " \t ", "\v", "\f "]))
        lines.append("v" + sep + sep.join(num() for _ in range(n)) + str(rng.choice(["", " ", "\r", " \r"])))
        if rng.random() < 0.5:
            lines.append("vn" + sep + sep.join(num() for _ in range(rng.choice([3, 3, 4]))))
    lines += [" v 1 2 3", "\tv 4 5 6", " vn 0 0 1", " f 1 2 3", "v 1 2", "vn 1 2", "v", "vn", "f", "f 1 2", "#v 1 2 3", "v 1 2 3#x",
              "vv 1 2 3 v 1 2 3", "g v 1 2 3", "x f 1 2 3", "f 1 2 3 f 4 5 6", "f 1/2 f 3 4 5", "f 1/ 2 3 f 7/8 9/10 11/12", "f1 2 3", "v1 2 3"]
    # NUL bytes inside the text: non-blank characters like any other to the reference's patterns (`\S`), the end of the
    # number to std::stof (a token that STARTS with one makes the reference throw, like every token without a number:
    # not in this file, the oracle would take the test process down with it)
    lines += ["v 1 2 3\x00 4", "v 1 2\x00 3 4", "vn 0 0\x001 1 5", "f 1 2 3\x00 4", "f 1 2 3 \x004", "\x00v 1 2 3", "v\x00 1 2 3",
              "v 5 6 7 8\x00", "f 1/2/3 4/5/6 7/8/9\x00 1/1/1"]
    idx = lambda: str(rng.integers(1, 250))
    for _ in range(200):
        form, n = rng.integers(0, 6), rng.choice([3, 4, 5, 2])
        if form == 0: c = [idx() for _ in range(n)]
        elif form == 1: c = [idx() + "/" + idx() for _ in range(n)]
        elif form == 2: c = [idx() + "//" + idx() for _ in range(n)]
        elif form == 3: c = [idx() + "/" + idx() + "/" + idx() for _ in range(n)]
        elif form == 4: c = [idx() + str(rng.choice(["", "/" + idx(), "//" + idx(), "/" + idx() + "/" + idx(), "/", "//"])) for _ in range(n)]
        else:
            c = [idx() for _ in range(n)]
            c[rng.integers(0, n)] = str(rng.choice(["-1", "a", "1.5", "0", "+3", "3x"]))
        lines.append("f " + str(rng.choice([" ", "  ", "\t"])).join(c) + str(rng.choice(["", " ", "\r"])))
    return lines


def test_obj_loader_records_equal_the_reference_loaders(emu, oracle, tmp_path):
    """ObjData of a torture file, record by record, against the reference's own regex loader (obj/obj.cpp:9-175) compiled
    into the oracle: vertices, normals, face indices and corner counts bit for bit -- well-formed and malformed lines
    alike.  Only `textures` of `a//n` faces is left out: the reference leaves it indeterminate there."""
    rng = np.random.default_rng(7)
    path = tmp_path / "torture.obj"
    path.write_bytes("\n".join(_torture_obj_lines(rng)).encode())   # (no newline at the end of the file)
    got, want = emu.obj_load(path), oracle.obj_load(path)
    assert got["vertices"].shape == want["vertices"].shape and len(got["vertices"]) > 240
    assert (got["vertices"].view(np.uint32) == want["vertices"].view(np.uint32)).all()
    assert got["normals"].shape == want["normals"].shape and (got["normals"].view(np.uint32) == want["normals"].view(np.uint32)).all()
    gf, wf = got["faces"], want["faces"]
    assert gf.shape == wf.shape and len(gf) > 80
    assert (gf[:, :4] == wf[:, :4]).all() and (gf[:, 8:] == wf[:, 8:]).all()
    textures_defined = (gf[:, 8:12] == 0).all(1) | (gf[:, 4:8] != 0).any(1)
    assert textures_defined.sum() > 40 and (gf[textures_defined, 4:8] == wf[textures_defined, 4:8]).all()
    assert (gf[~textures_defined, 4:8] == 0).all()


def test_obj_loader_ranges_and_number_parsing_at_scale(emu, tmp_path):
    """A file large enough to be cut into several ranges (one per host thread), with malformed and foreign lines spread
    through every range: record order, counts and values against libc's strtof -- what std::stof calls -- on every token
    (the loader's exact-decimal fast path must agree with it bit for bit)."""
    import ctypes

    libc = ctypes.CDLL(None)
    libc.strtof.restype, libc.strtof.argtypes = ctypes.c_float, [ctypes.c_char_p, ctypes.c_void_p]
    rng = np.random.default_rng(3)
    n = 60000
    coords = rng.uniform(-50, 50, (n, 3))
    digits = rng.integers(0, 10, (n, 3))
    lines, want_v, want_n, want_f = [], [], [], []
    for i in range(n):
        tok = [f"{coords[i, k]:.{digits[i, k]}f}" for k in range(3)]
        lines.append("v " + " ".join(tok))
        want_v.append(tok)
        if i % 3 == 0:
            lines.append("vn " + " ".join(tok[::-1]))
            want_n.append(tok[::-1])
        a, b, c = (int(x) for x in rng.integers(1, n, 3))
        kind = i % 7
        if kind == 0: lines.append(f"f {a} {b} {c}"); want_f.append([a, b, c, c, 0, 0, 0, 0, 0, 0, 0, 0, 3])
        elif kind == 1: lines.append(f"f {a}//{c} {b}//{a} {c}//{b}"); want_f.append([a, b, c, c, 0, 0, 0, 0, c, a, b, b, 3])
        elif kind == 2: lines.append(f"f {a}/{b} {b}/{c} {c}/{a} {a}/{a}"); want_f.append([a, b, c, a, b, c, a, a, 0, 0, 0, 0, 4])
        elif kind == 3: lines.append(f"f {a}/{b}/{c} {b}/{c}/{a} {c}/{a}/{b}"); want_f.append([a, b, c, c, b, c, a, 0, c, a, b, b, 3])
        elif kind == 4: lines += ["v 1 2", f"f {a} {b}", f" v {a} {b} {c}", "# v 1 2 3", "vt 0 0", ""]   # nothing of this is a record
        elif kind == 5: lines.append(f"f {a} {b} {c}/{a}"); want_f.append([a, b, c, c, 0, 0, 0, 0, 0, 0, 0, 0, 3])
        else: lines.append(f"f {a} {b} {c} {a} {b}"); want_f.append([a, b, c, a, 0, 0, 0, 0, 0, 0, 0, 0, 4])
    path = tmp_path / "large.obj"
    path.write_bytes(("\n".join(lines) + "\n").encode())
    assert path.stat().st_size > 3 << 20
    got = emu.obj_load(path)
    value = lambda t: libc.strtof(t.encode(), None)
    want_vertices = np.array([[value(t) for t in tok] + [1.0] for tok in want_v], np.float32)
    want_normals = np.array([[value(t) for t in tok] for tok in want_n], np.float32)
    assert got["vertices"].shape == want_vertices.shape and (got["vertices"].view(np.uint32) == want_vertices.view(np.uint32)).all()
    assert got["normals"].shape == want_normals.shape and (got["normals"].view(np.uint32) == want_normals.view(np.uint32)).all()
    assert got["faces"].shape == (len(want_f), 13) and (got["faces"] == np.array(want_f, np.int32)).all()


def test_obj_loader_file_that_ends_exactly_at_a_page_boundary(emu, oracle, tmp_path):
    """The loader parses the mapped file in place and relies on a NUL behind the last byte; a file that fills its last
    page exactly and has no final newline takes the copying path instead.  Also: empty file, missing file."""
    import mmap

    tail = b"v 1.25 2.5 3.75\nvn 0 0 1\nf 1 1 1\nv 9 8 7.5"
    for size in (mmap.PAGESIZE, 2 * mmap.PAGESIZE, 2 * mmap.PAGESIZE + 1):
        path = tmp_path / f"exact_{size}.obj"
        path.write_bytes(b"#" * (size - len(tail) - 1) + b"\n" + tail)
        assert path.stat().st_size == size
        got, want = emu.obj_load(path), oracle.obj_load(path)
        assert got["vertices"].tolist() == want["vertices"].tolist() == [[1.25, 2.5, 3.75, 1.0], [9.0, 8.0, 7.5, 1.0]]
        assert got["normals"].tolist() == [[0.0, 0.0, 1.0]] and got["faces"][:, :4].tolist() == [[1, 1, 1, 1]]
    empty = tmp_path / "empty.obj"
    empty.write_bytes(b"")
    assert all(len(a) == 0 for a in emu.obj_load(empty).values())
    assert emu.obj_load(tmp_path / "missing.obj") is None and oracle.obj_load(tmp_path / "missing.obj") is None


def _tone_torture_rgb():
    rng = np.random.default_rng(21)
    special = np.array([0.0, -0.0, 0.5 / 255, 1.5 / 255, 2.5 / 255, 127.5 / 255, 254.5 / 255, 1.0, 255.4 / 255, 255.5 / 255, 256.0 / 255, 2.0,
                        -0.002, -1.0, np.nan, np.inf, -np.inf, 8.4e6, 8.5e6, 1e10, 1e30, -1e10, 3.4e38], np.float32)
    rgb = np.concatenate([rng.uniform(0.0, 1.1, 30000), rng.uniform(-0.1, 40.0, 3000), rng.integers(0, 512, 3000) / 510.0,
                          np.tile(special, 40), special[:3 * 7]]).astype(np.float32)
    rgb = rgb[: len(rgb) // 3 * 3]
    rng.shuffle(rgb)   # specials at every position of OpenCV's vector loops and their scalar tails
    return rgb.reshape(1, -1, 3)


def test_tone_path_against_opencv_itself(emu, oracle, tmp_path):
    """The 8-bit values of Image::save (image.cpp:7-19) are whatever cv::imwrite makes of the CV_32FC3 matrix.  The C++
    OpenCV the reference links is absent, but the image has OpenCV's Python build: the reference's matrix (255 * powf,
    B and R swapped, from the oracle's libm expression) goes through cv2.imwrite -> cv2.imread, and both the oracle's
    stand-in and the device-code restatement (csrc/math.cuh: tone_u8) must store the same bytes -- halves to even,
    saturation, and NaN / infinities / values beyond int32 all 0."""
    cv2 = pytest.importorskip("cv2")
    rgb = _tone_torture_rgb()
    for gamma in (1.0, 0.45454547):
        bgr255, want_u8 = oracle.tone(rgb.reshape(-1, 3), gamma)
        with np.errstate(all="ignore"):
            assert cv2.imwrite(str(tmp_path / "ref.png"), bgr255.reshape(rgb.shape))
        stored = cv2.imread(str(tmp_path / "ref.png"), cv2.IMREAD_UNCHANGED)
        assert stored.dtype == np.uint8 and stored.shape == rgb.shape
        assert (stored.reshape(-1, 3) == want_u8).all()
        got_f, got_u8 = emu.tone(rgb.reshape(-1, 3), gamma)
        if gamma == 1.0:
            assert (got_u8 == stored.reshape(-1, 3)).all()
        else:   # pow in double rounded once vs glibc powf: a value may sit on the other side of a half
            assert (np.abs(got_u8.astype(int) - stored.reshape(-1, 3).astype(int)) <= 1).all() and (got_u8 != stored.reshape(-1, 3)).mean() < 1e-3


def test_image_save_writes_the_png_opencv_would(emu, oracle, tmp_path):
    """Image::save("*.png") of the host library (its own encoder, host/src/png.cpp) against the reference's recipe run
    through OpenCV: same size, same pixels when decoded by OpenCV and by PIL; chunk CRCs and the zlib stream valid."""
    import struct
    import zlib

    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for h, w in [(1, 1), (3, 5), (97, 131), (240, 320)]:   # (the last one spans several stored deflate blocks)
        film = rng.uniform(-0.05, 1.3, (h, w, 3)).astype(np.float32)
        film[0, 0] = [np.nan, np.inf, 2.0]
        ours, ref = tmp_path / f"ours_{h}x{w}.PNG", tmp_path / f"ref_{h}x{w}.png"
        emu.image_save(film, ours, 1.0)
        bgr255, _ = oracle.tone(film.reshape(-1, 3), 1.0)
        with np.errstate(all="ignore"):
            assert cv2.imwrite(str(ref), bgr255.reshape(h, w, 3))
        a, b = cv2.imread(str(ours), cv2.IMREAD_UNCHANGED), cv2.imread(str(ref), cv2.IMREAD_UNCHANGED)
        assert a is not None and a.shape == b.shape == (h, w, 3) and (a == b).all()
        # structure: signature, IHDR / IDAT / IEND with correct CRCs, stored zlib stream with a correct Adler-32
        blob = ours.read_bytes()
        assert blob[:8] == b"\x89PNG\r\n\x1a\n"
        pos, chunks = 8, []
        while pos < len(blob):
            n, kind = struct.unpack(">I4s", blob[pos:pos + 8])
            data = blob[pos + 8:pos + 8 + n]
            assert struct.unpack(">I", blob[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(kind + data)
            chunks.append((kind, data))
            pos += 12 + n
        assert [k for k, _ in chunks] == [b"IHDR", b"IDAT", b"IEND"]
        assert chunks[0][1] == struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)
        raw = zlib.decompress(chunks[1][1])
        rows = np.frombuffer(raw, np.uint8).reshape(h, 1 + 3 * w)
        assert (rows[:, 0] == 0).all() and (rows[:, 1:].reshape(h, w, 3)[..., ::-1] == b).all()
    try:
        from PIL import Image as PILImage
    except ImportError:
        return
    assert (np.asarray(PILImage.open(ours).convert("RGB"))[..., ::-1] == b).all()


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
    # the library's own rule (in-library multi-GPU render) is the same function
    lib = ctypes.CDLL(str(ROOT / "quetzalcoatlus_b200" / "_lib" / "libqz_b200.so"))
    lib.qz_strip_rows.restype, lib.qz_strip_rows.argtypes = ctypes.c_uint32, [ctypes.c_uint32, ctypes.c_uint32]
    for height in (1, 7, 54, 96, 600, 800, 1080, 2160, 4320):
        for world in (1, 2, 3, 4, 8, 16):
            assert lib.qz_strip_rows(height, world) == bench.strip_rows_for(height, world), (height, world)


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
