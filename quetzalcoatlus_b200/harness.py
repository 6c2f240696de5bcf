"""ctypes binding of the scene harness entry points.

The product harness (``libqz_harness.so``, prefix ``qzh_``) drives the C++ host library --
the name-for-name mirror of the reference's scene API -- which in turn calls the CUDA
library through the C ABI of ``include/qz_b200.h``.  The oracle build exports the same
entry points with the prefix ``orc_`` (``oracle/ref_harness.cpp``); the tests bind it with
this same class, so both sides are driven by identical Python code.

Nothing here computes anything: PyTorch/NumPy only hold buffers.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
DATA_DIR = PKG_DIR / "data"

TRACE_RECORD_FLOATS = 32

_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_int_p = ctypes.POINTER(ctypes.c_int)


class QzStats(ctypes.Structure):
    """``qz_stats`` of include/qz_b200.h."""

    _fields_ = [
        ("paths", ctypes.c_uint64),
        ("rays_closest", ctypes.c_uint64),
        ("rays_shadow", ctypes.c_uint64),
        ("shade_calls", ctypes.c_uint64),
        ("iterations", ctypes.c_uint64),
        ("kernel_launches", ctypes.c_uint64),
        ("node_visits", ctypes.c_uint64),
        ("prim_tests", ctypes.c_uint64),
        ("ms_total", ctypes.c_float),
        ("ms_closest", ctypes.c_float),
        ("ms_shadow", ctypes.c_float),
        ("ms_shade", ctypes.c_float),
        ("ms_other", ctypes.c_float),
        ("bvh_nodes", ctypes.c_uint32),
        ("bvh_bytes", ctypes.c_uint32),
        ("ms_sample", ctypes.c_float),
        ("stack_overflows", ctypes.c_uint32),
    ]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


class QzCamera(ctypes.Structure):
    """``qz_camera`` of include/qz_b200.h."""

    _fields_ = [
        ("image_width", ctypes.c_uint32),
        ("image_height", ctypes.c_uint32),
        ("pos", ctypes.c_float * 3),
        ("viewport_bottom_left", ctypes.c_float * 3),
        ("pixel_delta_u", ctypes.c_float * 3),
        ("pixel_delta_v", ctypes.c_float * 3),
        ("sensor_rgb", ctypes.c_void_p),
        ("imaging_ratio", ctypes.c_float),
    ]


class QzRegion(ctypes.Structure):
    _fields_ = [("strip_rows", ctypes.c_uint32), ("n_shards", ctypes.c_uint32), ("shard", ctypes.c_uint32)]


class QzRenderOptions(ctypes.Structure):
    _fields_ = [
        ("flags", ctypes.c_uint32),
        ("pool_paths", ctypes.c_uint32),
        ("samples_per_pass", ctypes.c_uint32),
        ("reserved", ctypes.c_uint32),
    ]


QZ_FLAG_UNSORTED_SHADING = 1
QZ_FLAG_COUNT_TRAVERSAL = 2
QZ_FLAG_STAGE_TIMING = 4
QZ_FLAG_FORCE_BVH = 8
QZ_FLAG_LANE_TRAVERSAL = 16   # round-1 evidence arms: rejected with QZ_ERR_INVALID
QZ_FLAG_OCTET_TRAVERSAL = 32
QZ_FLAG_EXACT_ARITHMETIC = 64
QZ_FLAG_FORCE_MEMO = 128
QZ_FLAG_NO_MEMO = 256


@dataclass
class RenderOutput:
    color: np.ndarray
    normal: np.ndarray
    albedo: np.ndarray
    seconds: float
    rays: int
    n_threads: int


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


class Scene:
    """A built scene (scene + camera + materials) living behind a harness handle."""

    def __init__(self, harness: "Harness", handle: int, name: str):
        self._h = harness
        self._handle = ctypes.c_void_p(handle)
        self.name = name
        w, h, spp, mb = (ctypes.c_int() for _ in range(4))
        harness._fn("scene_info")(self._handle, ctypes.byref(w), ctypes.byref(h), ctypes.byref(spp), ctypes.byref(mb))
        self.width, self.height, self.default_spp, self.default_max_bounces = w.value, h.value, spp.value, mb.value

    def close(self) -> None:
        if self._handle:
            self._h._fn("scene_free")(self._handle)
            self._handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def render(self, spp: int | None = None, max_bounces: int | None = None) -> RenderOutput:
        spp = self.default_spp if spp is None else spp
        max_bounces = self.default_max_bounces if max_bounces is None else max_bounces
        shape = (self.height, self.width, 3)
        color, normal, albedo = (np.zeros(shape, np.float32) for _ in range(3))
        sec, rays, nthr = ctypes.c_double(), ctypes.c_ulonglong(), ctypes.c_int()
        rc = self._h._fn("render")(
            self._handle, spp, max_bounces, color.ctypes.data_as(_c_float_p), normal.ctypes.data_as(_c_float_p),
            albedo.ctypes.data_as(_c_float_p), ctypes.byref(sec), ctypes.byref(rays), ctypes.byref(nthr))
        if rc != 0:
            raise RuntimeError(f"{self._h.prefix}render failed with code {rc}")
        return RenderOutput(color, normal, albedo, sec.value, rays.value, nthr.value)

    def render_only(self, spp: int, max_bounces: int) -> float:
        """render() with the RenderResult kept on the C++ side (product harness): returns one value of the colour plane."""
        sec, v = ctypes.c_double(), ctypes.c_float()
        rc = self._h._fn("render_only")(self._handle, spp, max_bounces, ctypes.byref(sec), ctypes.byref(v))
        if rc != 0:
            raise RuntimeError(f"{self._h.prefix}render_only failed with code {rc}")
        return v.value

    def trace_paths(self, xys, spp: int | None = None, max_bounces: int | None = None) -> np.ndarray:
        """Replay pixel-samples (x, y, s) -- y is the sampler/camera y (= H-1-row) -- and return
        one 32-float record per path (layout in include/qz_b200.h: qz_trace_paths)."""
        spp = self.default_spp if spp is None else spp
        max_bounces = self.default_max_bounces if max_bounces is None else max_bounces
        xys = _i32(xys).reshape(-1, 3)
        rec = np.zeros((len(xys), TRACE_RECORD_FLOATS), np.float32)
        rc = self._h._fn("trace_paths")(self._handle, spp, max_bounces, len(xys), xys.ctypes.data_as(_c_int_p),
                                        rec.ctypes.data_as(_c_float_p))
        if rc != 0:
            raise RuntimeError(f"{self._h.prefix}trace_paths failed with code {rc}")
        return rec

    def camera_fields(self) -> np.ndarray:
        out = np.zeros(21, np.float32)
        self._h._fn("camera_fields")(self._handle, out.ctypes.data_as(_c_float_p))
        return out.reshape(7, 3)

    def sensor_eval(self, u_and_l) -> np.ndarray:
        a = _f32(u_and_l).reshape(-1, 5)
        out = np.zeros((len(a), 3), np.float32)
        rc = self._h._fn("sensor_eval")(self._handle, len(a), a.ctypes.data_as(_c_float_p), out.ctypes.data_as(_c_float_p))
        if rc != 0:
            raise RuntimeError(f"{self._h.prefix}sensor_eval failed with code {rc}")
        return out

    def intersect(self, rays) -> np.ndarray:
        """rays: n x (o, d) -> n x (t, u, v, Ng.xyz, geomID, primID); t = -1 on a miss."""
        r = _f32(rays).reshape(-1, 6)
        out = np.zeros((len(r), 8), np.float32)
        rc = self._h._fn("intersect")(self._handle, len(r), r.ctypes.data_as(_c_float_p), out.ctypes.data_as(_c_float_p))
        if rc != 0:
            raise RuntimeError(f"{self._h.prefix}intersect failed with code {rc}")
        return out

    # ---- product side only: raw C-ABI access for the bench and the multi-GPU driver
    def c_scene_handle(self) -> int:
        fn = self._h._fn("scene_handle")
        fn.restype = ctypes.c_void_p
        return fn(self._handle)

    def c_camera(self) -> QzCamera:
        cam = QzCamera()
        self._h._fn("camera")(self._handle, ctypes.byref(cam))
        return cam

    def render_flags(self, spp: int, max_bounces: int | None = None, flags: int = 0, pool: int = 0, samples_per_pass: int = 0,
                     region: tuple[int, int, int] | None = None, reserved: int = 0):
        """qz_render through the raw C ABI with explicit render options; returns (RenderOutput, stats).
        region = (strip_rows, n_shards, shard): rows of other shards keep their zeros."""
        max_bounces = self.default_max_bounces if max_bounces is None else max_bounces
        shape = (self.height, self.width, 3)
        color, normal, albedo = (np.zeros(shape, np.float32) for _ in range(3))
        cam, st, opts = self.c_camera(), QzStats(), QzRenderOptions(flags, pool, samples_per_pass, reserved)
        reg = ctypes.byref(QzRegion(*region)) if region else None
        lib = self._h.lib
        lib.qz_render.restype = ctypes.c_int
        rc = lib.qz_render(ctypes.c_void_p(self.c_scene_handle()), ctypes.byref(cam), spp, max_bounces, reg, ctypes.byref(opts),
                           color.ctypes.data_as(ctypes.c_void_p), normal.ctypes.data_as(ctypes.c_void_p),
                           albedo.ctypes.data_as(ctypes.c_void_p), ctypes.byref(st))
        if rc != 0:
            lib.qz_last_error.restype = ctypes.c_char_p
            raise RuntimeError(f"qz_render failed: {lib.qz_last_error().decode()}")
        return RenderOutput(color, normal, albedo, st.ms_total * 1e-3, st.rays_closest + st.rays_shadow, 0), st.as_dict()

    def last_stats(self) -> dict:
        st = QzStats()
        self._h._fn("last_stats")(ctypes.byref(st))
        return st.as_dict()


class Harness:
    def __init__(self, lib_path: os.PathLike | str, prefix: str, data_dir: os.PathLike | str = DATA_DIR,
                 mode: int = ctypes.DEFAULT_MODE):
        self.path = str(lib_path)
        self.prefix = prefix
        self.lib = ctypes.CDLL(self.path, mode=mode)
        self._fn("scene_build").restype = ctypes.c_void_p
        self._fn("impl").restype = ctypes.c_char_p
        rc = self._fn("init")(str(data_dir).encode())
        if rc != 0:
            raise RuntimeError(f"{prefix}init({data_dir}) failed with code {rc}")

    def _fn(self, name: str):
        return getattr(self.lib, self.prefix + name)

    def math_probe(self, op: int, x) -> np.ndarray:
        """qz_math_probe of the C ABI: op 0 -> (sinf, cosf) per argument."""
        a = _f32(x).ravel()
        out = np.zeros((len(a), 2), np.float32)
        rc = self.lib.qz_math_probe(op, len(a), a.ctypes.data_as(_c_float_p), out.ctypes.data_as(_c_float_p))
        if rc != 0:
            raise RuntimeError(f"qz_math_probe failed with code {rc}")
        return out

    def set_default_flags(self, flags: int) -> int:
        """QZ_FLAG_* bits OR-ed into every later render / trace_paths call of this thread (product library only;
        the oracle and the host emulation have one arithmetic).  Returns the previous value."""
        fn = getattr(self.lib, "qz_set_default_flags", None)
        if fn is None:
            return 0
        fn.restype = ctypes.c_uint32
        return int(fn(ctypes.c_uint32(flags)))

    def arithmetic(self, exact: bool):
        """Context manager: run the enclosed calls in the reference's arithmetic (exact=True) or in the default
        radiometric mode (include/qz_b200.h: QZ_FLAG_EXACT_ARITHMETIC)."""
        import contextlib

        @contextlib.contextmanager
        def scope():
            old = self.set_default_flags(QZ_FLAG_EXACT_ARITHMETIC if exact else 0)
            try:
                yield self
            finally:
                self.set_default_flags(old)

        return scope()

    def build_times(self) -> dict:
        """Scene-build times (ms) of the last build_scene on this thread: OBJ text parse, the whole Scene::commit()
        (flatten + upload + BVH build) and the device part of it (product harness only)."""
        fn = getattr(self.lib, self.prefix + "build_times", None)
        if fn is None:
            return {}
        a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        fn(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return {"obj_parse_ms": a.value, "commit_ms": b.value, "bvh_build_ms": c.value}

    def impl(self) -> str:
        return self._fn("impl")().decode()

    def build_scene(self, name: str, width: int = 0, height: int = 0, obj_path: str | None = None,
                    obj_material: str = "", obj_light: str = "") -> Scene:
        handle = self._fn("scene_build")(name.encode(), width, height, obj_path.encode() if obj_path else None,
                                         obj_material.encode(), obj_light.encode())
        if not handle:
            raise RuntimeError(f"{self.prefix}scene_build({name!r}) failed")
        return Scene(self, handle, name)

    def sampler_eval(self, spp: int, width: int, height: int, q) -> np.ndarray:
        q = _i32(q).reshape(-1, 4)
        out = np.zeros(len(q), np.float32)
        rc = self._fn("sampler_eval")(spp, width, height, len(q), q.ctypes.data_as(_c_int_p), out.ctypes.data_as(_c_float_p))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}sampler_eval failed with code {rc}")
        return out

    def tone(self, rgb, gamma: float = 1.0):
        """Image::save's tone path (image.cpp:7-19): (n, 3) RGB floats -> (255 * powf(c, gamma) as BGR floats, the same as
        8-bit BGR)."""
        a = _f32(rgb).reshape(-1, 3)
        f = np.zeros_like(a)
        u8 = np.zeros(a.shape, np.uint8)
        rc = self._fn("tone")(a.ctypes.data_as(_c_float_p), len(a), ctypes.c_float(gamma), f.ctypes.data_as(_c_float_p),
                              u8.ctypes.data_as(ctypes.c_void_p))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}tone failed with code {rc}")
        return f, u8

    def image_save(self, rgb, path, gamma: float = 1.0) -> None:
        """Image::save (image.cpp:7-19) of an (H, W, 3) float plane through the host library."""
        a = _f32(rgb)
        h, w, _ = a.shape
        self._fn("image_save")(a.ctypes.data_as(_c_float_p), h, w, str(path).encode(), ctypes.c_float(gamma))

    def obj_load(self, path) -> dict:
        """ObjData of an OBJ file as this library's loader reads it (obj/obj.hpp): vertices (n, 4), normals (n, 3),
        faces (n, 13) = vertices[4], textures[4], normals[4], n_vertices.  None when the loader gives up."""
        fn = self._fn("obj_load")
        counts = (ctypes.c_long * 3)()
        if fn(str(path).encode(), counts, None, None, None) != 0:
            return None
        v, n, f = np.zeros((counts[0], 4), np.float32), np.zeros((counts[1], 3), np.float32), np.zeros((counts[2], 13), np.int32)
        fn(str(path).encode(), counts, v.ctypes.data_as(_c_float_p), n.ctypes.data_as(_c_float_p), f.ctypes.data_as(ctypes.c_void_p))
        return {"vertices": v, "normals": n, "faces": f}

    def eval_spectrum(self, name: str, lambdas) -> np.ndarray:
        lam = _f32(lambdas).ravel()
        out = np.zeros(len(lam), np.float32)
        rc = self._fn("eval_spectrum")(name.encode(), len(lam), lam.ctypes.data_as(_c_float_p), out.ctypes.data_as(_c_float_p))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}eval_spectrum({name!r}) failed with code {rc}")
        return out
