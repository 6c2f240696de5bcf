"""GPU parity tests proper: the product (host library -> C ABI -> sm_100a kernels) against the
oracle on the same seeded inputs, and against the committed golden fixtures.

Integer / table work (sampler, spectra, sensor, intersection) is BIT-EXACT.  Per-path
radiance follows the north-star tolerance: within 1e-4 relative on replayed sampler
sequences.  The library has two ARITHMETIC MODES (csrc/common.cuh; include/qz_b200.h:
QZ_FLAG_EXACT_ARITHMETIC) and every path test runs in both:
  * exact -- every float operation as the reference's x86-64 build performs it.  sinf / cosf are
    glibc's own algorithm, restated (csrc/math.cuh), so paths are bit-identical to the oracle's
    except through atan2f / acosf (uv of textured spheres: CUDA double, rounded once);
  * fast (the default, what render() runs) -- geometry and discrete decisions as in exact mode,
    radiometric values with fused multiply-adds and hardware reciprocals: radiance within ~1e-6.
An ulp can flip a discrete decision (Russian roulette, a texel), so the tests bound the FRACTION
of paths outside tolerance instead of demanding zero.
"""
import os
from pathlib import Path

import numpy as np
import pytest

from common import ANALYTIC_SCENES, EDGE_SCENES, RADIANCE, RAYS, bits_equal, path_agreement, pixel_samples

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
REL_TOL = 1e-4            # north_star: per-path radiance within 1e-4 relative
MAX_DIVERGENT = 5e-3      # fraction of paths allowed outside it (knife-edge decisions)
MODES = ["exact", "fast"]
EXACT_SCENES = ["cornell_box", "glass_spheres", "opposing_planes", "cornell_mixed", "mandelbrot"]   # no textured spheres


@pytest.fixture(params=MODES)
def mode(request, qz):
    """Runs the test once per arithmetic mode; the library default (fast) is restored afterwards."""
    with qz.arithmetic(exact=request.param == "exact"):
        yield request.param


def test_sincos_is_glibc(qz):
    """csrc/math.cuh restates glibc's sinf / cosf (double-precision polynomial, rounded once): bit-identical to the
    host libm on the arguments the warps produce and well beyond."""
    import ctypes

    libm = ctypes.CDLL("libm.so.6")
    libm.sinf.restype = libm.cosf.restype = ctypes.c_float
    libm.sinf.argtypes = libm.cosf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(12)
    x = np.concatenate([rng.uniform(-1.0, 7.0, 60000), rng.uniform(-100.0, 100.0, 20000), 10.0 ** rng.uniform(-8, 0, 10000),
                        np.array([0.0, 0.75, 0.7499999, 0.785398, 1.5707964, 3.1415927, 6.2831855, 2.0 ** -12, 119.9])]).astype(np.float32)
    got = qz.math_probe(0, x)
    want = np.array([[libm.sinf(float(v)), libm.cosf(float(v))] for v in x], np.float32)
    assert bits_equal(got, want).all(), int((~bits_equal(got, want)).sum())



@pytest.mark.parametrize("fixture", ["sampler_kat.npz", "sampler_kat_800.npz"])
def test_sampler_bit_exact(qz, fixture):
    g = np.load(GOLDEN / fixture)
    spp, w, h = (int(v) for v in g["res"])
    assert bits_equal(qz.sampler_eval(spp, w, h, g["q"]), g["values"]).all()


def test_sampler_bit_exact_large(qz, oracle):
    rng = np.random.default_rng(9)
    n = 200_000
    q = np.stack([rng.integers(0, 3840, n), rng.integers(0, 2160, n), rng.integers(0, 1024, n), rng.integers(0, 1000, n)], 1)
    assert bits_equal(qz.sampler_eval(1024, 3840, 2160, q), oracle.sampler_eval(1024, 3840, 2160, q)).all()


@pytest.mark.parametrize("name", ["X", "D65", "CANON_R", "CANON_B", "AL_IOR", "CU_ABSORPTION", "GLASS_SF11_IOR", "rgb:0.8,0.4,0.1",
                                  "rgb:0.5,0.5,0.5", "rgbu:1.1,1.8,3.0", "rgbi:8,2,4", "const:1.5"])
def test_spectra_bit_exact(qz, oracle, name):
    lam = np.concatenate([np.random.default_rng(1).uniform(355, 835, 3000), np.arange(355, 836, dtype=np.float64)]).astype(np.float32)
    assert bits_equal(qz.eval_spectrum(name, lam), oracle.eval_spectrum(name, lam)).all()


@pytest.mark.parametrize("name", ANALYTIC_SCENES)
def test_sensor_and_intersection_bit_exact(qz, oracle, name):
    rng = np.random.default_rng(11)
    with qz.build_scene(name) as sg, oracle.build_scene(name) as so:
        assert bits_equal(sg.camera_fields(), so.camera_fields()).all()
        ul = rng.uniform(0, 1, (2000, 5)).astype(np.float32)
        ul[:, 1:] *= 60.0
        assert bits_equal(sg.sensor_eval(ul), so.sensor_eval(ul)).all()
        o = np.tile(so.camera_fields()[0], (20000, 1)) + rng.normal(0, 0.5, (20000, 3))
        d = rng.normal(0, 1, (20000, 3))
        d[:, 2] -= 1.5
        rays = np.concatenate([o, d], 1).astype(np.float32)
        assert bits_equal(sg.intersect(rays), so.intersect(rays)).all()


@pytest.mark.parametrize("name", ANALYTIC_SCENES)
def test_paths_against_golden(qz, mode, name):
    g = np.load(GOLDEN / f"paths_{name}.npz")
    with qz.build_scene(name) as sc:
        got = sc.trace_paths(g["xys"])
    ok = path_agreement(g["records"], got, REL_TOL)
    assert 1.0 - ok.mean() <= 4 * MAX_DIVERGENT, f"{(~ok).sum()} of {len(ok)} golden paths outside tolerance"
    assert bits_equal(got[:, :4], g["records"][:, :4]).all()  # wavelengths come from the sampler: exact


@pytest.mark.parametrize("name", ANALYTIC_SCENES)
def test_paths_against_oracle(qz, oracle, mode, name):
    with qz.build_scene(name) as sg, oracle.build_scene(name) as so:
        xys = pixel_samples(so, 20000, seed=21)
        got, want = sg.trace_paths(xys), so.trace_paths(xys)
    ok = path_agreement(want, got, REL_TOL)
    frac = 1.0 - ok.mean()
    exact = bits_equal(got, want).all(1).mean()
    print(f"{name} [{mode}]: {exact:.4%} of paths bit-identical, {frac:.4%} outside rel {REL_TOL}")
    assert frac <= MAX_DIVERGENT
    # geometry and discrete decisions are the reference's in BOTH modes: same ray count, same first-hit normal
    assert (got[:, RAYS] == want[:, RAYS]).mean() >= 1.0 - MAX_DIVERGENT
    assert bits_equal(got[:, 12:15], want[:, 12:15]).all(1).mean() >= 1.0 - MAX_DIVERGENT
    if mode == "exact" and name in EXACT_SCENES:
        # no atan2f / acosf on these scenes' paths: everything the path touches is restated bit for bit
        assert exact >= 0.9995, f"{name}: only {exact:.4%} of the paths are bit-identical to the oracle's"


def test_glass_spheres_is_bit_exact(qz, oracle):
    """All-specular scene: in exact mode the GPU must reproduce the oracle to the bit through up to 64 bounces; in
    the default mode the geometry (ray counts, wavelengths, normals) still does and the radiance stays within 1e-5."""
    with qz.build_scene("glass_spheres") as sg, oracle.build_scene("glass_spheres") as so:
        xys = pixel_samples(so, 20000, seed=5)
        want = so.trace_paths(xys)
        with qz.arithmetic(exact=True):
            assert bits_equal(sg.trace_paths(xys), want).all()
        got = sg.trace_paths(xys)
    assert (got[:, RAYS] == want[:, RAYS]).all() and bits_equal(got[:, :4], want[:, :4]).all() and bits_equal(got[:, 12:15], want[:, 12:15]).all()
    assert path_agreement(want, got, 1e-5).all()


@pytest.mark.parametrize("material,light", [("alluminum", "point"), ("glass", "area"), ("diffuse", "ambient")])
def test_mesh_scene(qz, oracle, mode, small_mesh, material, light):
    kw = dict(obj_path=small_mesh, obj_material=material, obj_light=light)
    with qz.build_scene("obj_viewer", **kw) as sg, oracle.build_scene("obj_viewer", **kw) as so:
        xys = pixel_samples(so, 5000, seed=4)
        got, want = sg.trace_paths(xys), so.trace_paths(xys)
    assert 1.0 - path_agreement(want, got, REL_TOL).mean() <= MAX_DIVERGENT


def test_wavefront_film_equals_replayed_paths(qz, mode):
    """The wavefront pipeline (queues, regeneration, ordered film sum) against the same paths
    replayed one thread per path on the same GPU: bit-identical film, in either arithmetic mode (a fused
    multiply-add exists only where the source says so, csrc/common.cuh)."""
    for name, (w, h, spp) in {"cornell_box": (40, 36, 6), "kitchen_sink": (32, 24, 5), "textures": (36, 36, 3)}.items():
        with qz.build_scene(name, w, h) as sc:
            film = sc.render(spp)
            ys, xs, ss = np.meshgrid(np.arange(h), np.arange(w), np.arange(spp), indexing="ij")
            xys = np.stack([xs.ravel(), (h - 1 - ys).ravel(), ss.ravel()], 1).astype(np.int32)  # row -> sampler y
            rec = sc.trace_paths(xys, spp=spp).reshape(h, w, spp, 32)
        color = np.zeros((h, w, 3), np.float32)
        normal = np.zeros((h, w, 3), np.float32)
        albedo = np.zeros((h, w, 3), np.float32)
        for s in range(spp):  # ascending s, float32 adds: the reference's order
            color += rec[:, :, s, 20:23]
            albedo += rec[:, :, s, 23:26]
            normal += rec[:, :, s, 12:15]
        n = np.float32(spp)
        assert bits_equal(film.color, color / n).all(), name
        assert bits_equal(film.albedo, albedo / n).all(), name
        assert bits_equal(film.normal, normal / n).all(), name


def _replayed_film(sc, w, h, spp):
    ys, xs, ss = np.meshgrid(np.arange(h), np.arange(w), np.arange(spp), indexing="ij")
    xys = np.stack([xs.ravel(), (h - 1 - ys).ravel(), ss.ravel()], 1).astype(np.int32)
    rec = sc.trace_paths(xys, spp=spp).reshape(h, w, spp, 32)
    planes = [np.zeros((h, w, 3), np.float32) for _ in range(3)]
    for s in range(spp):
        planes[0] += rec[:, :, s, 20:23]
        planes[1] += rec[:, :, s, 12:15]
        planes[2] += rec[:, :, s, 23:26]
    return [p / np.float32(spp) for p in planes]


@pytest.mark.parametrize("material,light", [("alluminum", "point"), ("glass", "area"), ("diffuse", "ambient")])
def test_bvh_traversal_kernels_equal_replayed_paths(qz, mode, small_mesh, material, light):
    """The wavefront BVH kernel on a mesh scene (phase-scheduled traversal, shared-memory short stack) against the
    scalar traversal of the per-path replay: bit-identical films, and counters that agree on the rays."""
    from quetzalcoatlus_b200.harness import QZ_FLAG_COUNT_TRAVERSAL

    w, h, spp = 96, 72, 4
    with qz.build_scene("obj_viewer", w, h, obj_path=small_mesh, obj_material=material, obj_light=light) as sc:
        want = _replayed_film(sc, w, h, spp)
        rays = None
        for flags in (0, QZ_FLAG_COUNT_TRAVERSAL):
            for pool in (0, 1000):
                film, st = sc.render_flags(spp, flags=flags, pool=pool)
                assert st["stack_overflows"] == 0
                for got, ref, plane in zip((film.color, film.normal, film.albedo), want, ("color", "normal", "albedo")):
                    assert bits_equal(got, ref).all(), (flags, pool, plane, float(1.0 - bits_equal(got, ref).mean()))
                rays = rays or (st["rays_closest"], st["rays_shadow"])
                assert (st["rays_closest"], st["rays_shadow"]) == rays
                if flags & QZ_FLAG_COUNT_TRAVERSAL:
                    assert st["node_visits"] >= st["rays_closest"] and st["prim_tests"] > 0


def test_removed_evidence_arms_are_rejected(qz):
    from quetzalcoatlus_b200.harness import QZ_FLAG_LANE_TRAVERSAL

    with qz.build_scene("cornell_box", 16, 16) as sc:
        with pytest.raises(RuntimeError, match="no longer built"):
            sc.render_flags(1, flags=QZ_FLAG_LANE_TRAVERSAL)


def test_forced_bvh_on_analytic_scenes(qz, mode):
    """Scenes small enough for the flat kernel, pushed through the BVH kernels instead (QZ_FLAG_FORCE_BVH):
    spheres, quads, grid cells and hoisted huge primitives in one tree.  BVH answer == brute force, to the bit --
    including rays that leave the 2000-unit planes of opposing_planes towards its 0.8-unit spheres."""
    from quetzalcoatlus_b200.harness import QZ_FLAG_FORCE_BVH

    for name, (w, h, spp) in {"cornell_box": (40, 36, 4), "kitchen_sink": (64, 48, 6), "opposing_planes": (96, 54, 8),
                              "mandelbrot": (32, 32, 3), "glass_spheres": (32, 32, 4)}.items():
        with qz.build_scene(name, w, h) as sc:
            base, _ = sc.render_flags(spp)
            film, st = sc.render_flags(spp, flags=QZ_FLAG_FORCE_BVH)
            assert st["stack_overflows"] == 0
            assert bits_equal(film.color, base.color).all() and bits_equal(film.normal, base.normal).all() \
                and bits_equal(film.albedo, base.albedo).all(), name


def test_film_against_golden(qz, mode):
    for name in ["cornell_box", "textures", "kitchen_sink"]:
        g = np.load(GOLDEN / f"film_{name}.npz")
        h, w, _ = g["color"].shape
        with qz.build_scene(name, w, h) as sc:
            r = sc.render(int(g["spp"]))
        for plane in ("color", "normal", "albedo"):
            err = np.abs(getattr(r, plane) - g[plane])
            tol = 1e-4 * np.maximum(np.abs(g[plane]), 1e-2)
            assert (err <= tol).mean() >= 0.98, (name, plane, float((err <= tol).mean()))


def test_film_is_independent_of_pool_pass_and_sharding(qz, mode):
    """Bit-stable film: pool size, pass size and row sharding must not change a single bit."""
    import ctypes

    from quetzalcoatlus_b200.harness import QzRegion, QzRenderOptions, QzStats

    lib = qz.lib
    lib.qz_render.restype = ctypes.c_int
    with qz.build_scene("cornell_box", 48, 40) as sc:
        base = sc.render(8).color
        handle, cam = ctypes.c_void_p(sc.c_scene_handle()), sc.c_camera()

        def render(opts=None, regions=(None,)):
            color = np.zeros((40, 48, 3), np.float32)
            for reg in regions:
                st = QzStats()
                rc = lib.qz_render(handle, ctypes.byref(cam), 8, 64, ctypes.byref(reg) if reg else None,
                                   ctypes.byref(opts) if opts else None, color.ctypes.data_as(ctypes.c_void_p), None, None, ctypes.byref(st))
                assert rc == 0
            return color

        assert bits_equal(render(QzRenderOptions(0, 257, 0, 0)), base).all()      # tiny odd pool: heavy regeneration
        assert bits_equal(render(QzRenderOptions(0, 0, 3, 0)), base).all()        # 3 samples per pass: 3 passes
        assert bits_equal(render(QzRenderOptions(1, 0, 0, 0)), base).all()        # unsorted (uber-kernel) shading
        assert bits_equal(render(None, [QzRegion(4, 3, k) for k in range(3)]), base).all()  # 3 shards of 4-row strips


def test_mandelbrot_grid_at_full_scale(qz, oracle, mode):
    """examples/mandelbrot.cpp with its shipped 1200 x 1200 height grid (1.44 M grid cells =
    2.87 M triangles, rough copper): replayed paths against the oracle, and the wavefront BVH kernels
    against the replay on a crop-sized film."""
    with qz.build_scene("mandelbrot_full", 96, 96) as sg, oracle.build_scene("mandelbrot_full", 96, 96) as so:
        xys = pixel_samples(so, 3000, seed=9)
        got, want = sg.trace_paths(xys), so.trace_paths(xys)
        assert 1.0 - path_agreement(want, got, REL_TOL).mean() <= MAX_DIVERGENT
        film, st = sg.render_flags(3)
        ref = _replayed_film(sg, 96, 96, 3)
    assert st["stack_overflows"] == 0
    for got_plane, want_plane in zip((film.color, film.normal, film.albedo), ref):
        assert bits_equal(got_plane, want_plane).all()


@pytest.mark.parametrize("name", EDGE_SCENES)
def test_edge_scenes(qz, oracle, mode, name):
    """No lights / no geometry: replayed paths against the oracle, wavefront film against the replay."""
    with qz.build_scene(name) as sg, oracle.build_scene(name) as so:
        xys = pixel_samples(so, 5000, seed=6)
        got, want = sg.trace_paths(xys), so.trace_paths(xys)
        assert 1.0 - path_agreement(want, got, REL_TOL).mean() <= MAX_DIVERGENT
        w, h = sg.width, sg.height
        film, st = sg.render_flags(3)
        ref = _replayed_film(sg, w, h, 3)
    for got_plane, want_plane in zip((film.color, film.normal, film.albedo), ref):
        assert bits_equal(got_plane, want_plane).all()
    assert st["paths"] == w * h * 3


@pytest.mark.parametrize("max_bounces", [0, 1, 2])
def test_bounce_limits(qz, oracle, mode, max_bounces):
    """max_bounces 0, 1, 2: the loop's early break (render.cpp:137), incl. the separate conductor albedo stage."""
    for name in ("cornell_box", "kitchen_sink"):
        with qz.build_scene(name, 64, 48) as sg, oracle.build_scene(name, 64, 48) as so:
            xys = pixel_samples(so, 3000, seed=8, spp=4)
            got = sg.trace_paths(xys, spp=4, max_bounces=max_bounces)
            want = so.trace_paths(xys, spp=4, max_bounces=max_bounces)
            assert 1.0 - path_agreement(want, got, REL_TOL).mean() <= MAX_DIVERGENT, name
            film, _ = sg.render_flags(4, max_bounces=max_bounces)
            ys, xs, ss = np.meshgrid(np.arange(48), np.arange(64), np.arange(4), indexing="ij")
            all_xys = np.stack([xs.ravel(), (47 - ys).ravel(), ss.ravel()], 1).astype(np.int32)
            rec = sg.trace_paths(all_xys, spp=4, max_bounces=max_bounces).reshape(48, 64, 4, 32)
            color = np.zeros((48, 64, 3), np.float32); albedo = np.zeros_like(color)
            for s in range(4):
                color += rec[:, :, s, 20:23]; albedo += rec[:, :, s, 23:26]
            assert bits_equal(film.color, color / np.float32(4)).all(), name
            assert bits_equal(film.albedo, albedo / np.float32(4)).all(), name


def test_paths_deeper_than_255_bounces(qz, oracle):
    """max_bounces beyond the 8 bits the path depth had in round 1 (now a 16-bit field next to the sampler dimension):
    kitchen_sink at max_bounces = 3000 keeps a twentieth of its paths alive for thousands of rays (lossless dielectrics).
    The sampler dimension wraps the 1000-prime table many times on the way (sampler.cpp:404-454), all of it through the
    late queue / k_sample.  Replay against the oracle within tolerance (ray counts included), and the wavefront film and its ray
    statistics against the replay bit for bit."""
    w, h, spp, deep_limit = 32, 24, 2, 3000
    ys, xs, ss = np.meshgrid(np.arange(h), np.arange(w), np.arange(spp), indexing="ij")
    all_xys = np.stack([xs.ravel(), (h - 1 - ys).ravel(), ss.ravel()], 1).astype(np.int32)
    with qz.build_scene("kitchen_sink", w, h) as sg, oracle.build_scene("kitchen_sink", w, h) as so:
        got = sg.trace_paths(all_xys, spp=spp, max_bounces=deep_limit)
        want = so.trace_paths(all_xys, spp=spp, max_bounces=deep_limit)
        assert (want[:, RAYS] > 1000).sum() >= 10, "the scene no longer has deep paths: pick another"
        agree = path_agreement(want, got, REL_TOL)
        print(f"deep paths: {(want[:, RAYS] > 1000).sum()} of {len(want)} beyond 1000 rays, longest {int(want[:, RAYS].max())}; "
              f"{1.0 - agree.mean():.4%} outside rel {REL_TOL}")
        assert 1.0 - agree.mean() <= MAX_DIVERGENT
        film, st = sg.render_flags(spp, max_bounces=deep_limit)
        rec = got.reshape(h, w, spp, 32)
        color = np.zeros((h, w, 3), np.float32)
        for s in range(spp):
            color += rec[:, :, s, 20:23]
        assert bits_equal(film.color, color / np.float32(spp)).all()
        assert st["rays_closest"] + st["rays_shadow"] == int(got[:, RAYS].astype(np.int64).sum())
    with pytest.raises(RuntimeError, match="65535"):
        with qz.build_scene("cornell_box", 8, 8) as sc:
            sc.render_flags(1, max_bounces=65536)


@pytest.mark.skipif(not os.environ.get("QZ_GPU_FUZZ"), reason="opt-in (QZ_GPU_FUZZ=1): written after this round's GPU budget was spent, "
                    "not yet run on a GPU; the same scenes are bit-exact in the CPU emulation (tests/test_emu_parity.py)")
@pytest.mark.parametrize("seed", range(8))
def test_random_scenes_against_oracle(qz, oracle, mode, seed):
    """The seeded random scenes of the CPU suite through the GPU library: per-path agreement within the north-star
    tolerance on the paths whose reference radiance is finite, and the same set of non-finite paths."""
    name = f"fuzz:{seed}"
    with qz.build_scene(name) as sg, oracle.build_scene(name) as so:
        xys = pixel_samples(so, 3000, seed=seed)
        got, want = sg.trace_paths(xys), so.trace_paths(xys)
    finite = np.isfinite(want[:, RADIANCE]).all(1)
    assert (np.isfinite(got[:, RADIANCE]).all(1) == finite).mean() >= 1.0 - MAX_DIVERGENT
    if finite.any():
        assert 1.0 - path_agreement(want[finite], got[finite], REL_TOL).mean() <= MAX_DIVERGENT


def test_pipelines_do_not_change_the_film(qz, mode, small_mesh):
    """Renders large enough for the pool to be split into concurrent pipelines (two streams drawing
    paths from one cursor) against the single-pipeline run of the same call (stage timing forces one
    pipeline): bit-identical films and identical ray counts."""
    from quetzalcoatlus_b200.harness import QZ_FLAG_STAGE_TIMING

    cases = [("cornell_box", dict(), 320, 240, 6), ("kitchen_sink", dict(), 256, 256, 4),
             ("obj_viewer", dict(obj_path=small_mesh, obj_material="alluminum", obj_light="point"), 320, 240, 4)]
    for name, kw, w, h, spp in cases:
        with qz.build_scene(name, w, h, **kw) as sc:
            one, st1 = sc.render_flags(spp, flags=QZ_FLAG_STAGE_TIMING)
            two, st2 = sc.render_flags(spp)
            odd, st3 = sc.render_flags(spp, pool=70001, samples_per_pass=3)
        for other, st in ((two, st2), (odd, st3)):
            assert bits_equal(one.color, other.color).all(), name
            assert bits_equal(one.normal, other.normal).all(), name
            assert bits_equal(one.albedo, other.albedo).all(), name
            assert (st["rays_closest"], st["rays_shadow"], st["shade_calls"]) == (st1["rays_closest"], st1["rays_shadow"], st1["shade_calls"]), name


@pytest.mark.parametrize("name,w,h,spp", [("cornell_box", 40, 40, 8), ("opposing_planes", 48, 27, 8), ("textures", 40, 40, 6)])
def test_equal_spp_rmse_is_indistinguishable_from_cpu(qz, oracle, name, w, h, spp):
    """north_star / SURVEY 8.c: at equal spp the GPU film's RMSE against a high-spp reference must be statistically
    indistinguishable from the CPU film's.  Eight DISJOINT sample-index ranges are rendered on both sides (replayed
    pixel-samples s in [k*spp, (k+1)*spp), summed in the film's order), each film's RMSE is taken against a
    1024-spp oracle render, and the eight paired RMSE differences go through a paired t-test (7 degrees of freedom,
    two-sided 1 %: |t| < 3.499) plus a bound on the mean difference relative to the CPU's spread."""
    ranges = 8
    with oracle.build_scene(name, w, h) as so, qz.build_scene(name, w, h) as sg:
        ref = so.render(1024).color
        ys, xs, ss = np.meshgrid(np.arange(h), np.arange(w), np.arange(ranges * spp), indexing="ij")
        xys = np.stack([xs.ravel(), (h - 1 - ys).ravel(), ss.ravel()], 1).astype(np.int32)
        cpu = so.trace_paths(xys, spp=1024).reshape(h, w, ranges, spp, 32)[..., 20:23]
        gpu = sg.trace_paths(xys, spp=1024).reshape(h, w, ranges, spp, 32)[..., 20:23]   # default (fast) mode: what render() runs

    def rmse(rec, k):
        film = np.zeros((h, w, 3), np.float32)
        for s in range(spp):
            film += rec[:, :, k, s]
        return float(np.sqrt(np.mean((film / np.float32(spp) - ref) ** 2)))

    r_cpu = np.array([rmse(cpu, k) for k in range(ranges)])
    r_gpu = np.array([rmse(gpu, k) for k in range(ranges)])
    diff = r_gpu - r_cpu
    sd = diff.std(ddof=1)
    t = 0.0 if sd == 0.0 else float(diff.mean() / (sd / np.sqrt(ranges)))
    print(f"{name}: rmse cpu {r_cpu.mean():.6f} +- {r_cpu.std(ddof=1):.6f}, gpu {r_gpu.mean():.6f}; paired diff {diff.mean():.3e} +- {sd:.3e}, t = {t:.2f}")
    assert abs(t) < 3.499 or abs(diff.mean()) < 1e-3 * r_cpu.mean(), (name, t)
    assert abs(diff.mean()) <= 0.25 * r_cpu.std(ddof=1) + 1e-3 * r_cpu.mean()


def test_obj_viewer_1m_against_golden(qz):
    """BASELINE config 4 at FULL size: the obj_viewer scene on the synthetic ~1M-triangle mesh (tools/gen_mesh.py
    dragon_mesh), 4096 replayed paths against the fixture the oracle produced once with the reference's own OBJ
    loader (tests/golden/make_golden.py --obj1m).  The fixture carries the SHA-256 of the OBJ text it belongs to."""
    import hashlib
    import os
    import sys

    fixture = GOLDEN / "paths_obj_viewer_1m.npz"
    if not fixture.exists():
        pytest.skip("tests/golden/paths_obj_viewer_1m.npz has not been generated")
    g = np.load(fixture)
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import gen_mesh

    path = gen_mesh.ensure_obj(os.environ.get("QZ_MESH_DIR", "/tmp"), 1_000_000)
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == str(g["obj_sha256"]), "the generated OBJ text differs from the fixture's"
    with qz.build_scene("obj_viewer", obj_path=path, obj_material="alluminum", obj_light="point") as sc:
        for exact in (True, False):
            with qz.arithmetic(exact=exact):
                got = sc.trace_paths(g["xys"])
            ok = path_agreement(g["records"], got, REL_TOL)
            print(f"obj_viewer 1M [{'exact' if exact else 'fast'}]: {bits_equal(got, g['records']).all(1).mean():.4%} bit-identical, "
                  f"{1.0 - ok.mean():.4%} outside rel {REL_TOL}")
            assert 1.0 - ok.mean() <= MAX_DIVERGENT
            assert bits_equal(got[:, :4], g["records"][:, :4]).all()


@pytest.mark.parametrize("gamma", [1.0, 0.45454547, 2.2])
def test_tone_path_against_oracle(qz, oracle, gamma):
    """Image::save's tone path (image.cpp:10-15: x255, powf gamma, BGR) as a kernel, through the C ABI
    (qz_tone), against the oracle's libm expression: bit-exact at the default gamma of 1, within one ulp otherwise
    (device double pow rounded once vs glibc powf), 8-bit output equal except at rounding knife edges."""
    rng = np.random.default_rng(11)
    rgb = np.concatenate([rng.uniform(0.0, 1.2, (200000, 3)), rng.uniform(0.0, 40.0, (2000, 3)),
                          [[0.0, 1.0, 0.5], [1.0 / 255, 2.0 / 255, 0.5 / 255], [-0.25, 0.0, 1e-30]]]).astype(np.float32)
    got_f, got_u8 = qz.tone(rgb, gamma)
    want_f, want_u8 = oracle.tone(rgb, gamma)
    if gamma == 1.0:
        assert bits_equal(got_f, want_f).all() and (got_u8 == want_u8).all()
    else:
        # one ulp of powf, then the rounding of the x255 product: two ulps of the result
        ok = np.isnan(want_f) & np.isnan(got_f) | (np.abs(got_f - want_f) <= 2 * np.spacing(np.abs(want_f)))
        print(f"tone gamma {gamma}: {(got_f != want_f).mean():.2e} of values differ from glibc powf (by one ulp of the power)")
        assert ok.all()
        assert (np.abs(got_u8.astype(int) - want_u8.astype(int)) <= 1).all() and (got_u8 != want_u8).mean() < 1e-3
    # NaN, infinities and values beyond int32 after the x255: OpenCV's conversion stores 0 for all of them (the oracle's
    # stand-in is pinned against cv2.imwrite on the CPU, tests/test_abi_and_host.py)
    special = np.array([[np.nan, np.inf, -np.inf], [1e10, 8.5e6, 8.4e6], [3.4e38, -1e10, 2.0]], np.float32)
    assert (qz.tone(special, gamma)[1] == oracle.tone(special, gamma)[1]).all()
    if gamma == 1.0:
        assert qz.tone(special, gamma)[1].tolist() == [[0, 0, 0], [255, 0, 0], [255, 0, 0]]


def test_image_save_png_through_the_product_library(qz, oracle, tmp_path):
    """Image::save("*.png") of the host library as shipped -- tone kernel on the GPU, the library's own PNG encoder --
    decodes to the bytes the reference's recipe stores (oracle's tone arithmetic, pinned against OpenCV on the CPU)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(9)
    film = rng.uniform(-0.05, 1.3, (60, 84, 3)).astype(np.float32)
    film[0, 0] = [np.nan, np.inf, 1e10]
    qz.image_save(film, tmp_path / "film.png", 1.0)
    stored = cv2.imread(str(tmp_path / "film.png"), cv2.IMREAD_UNCHANGED)
    _, want_bgr8 = oracle.tone(film.reshape(-1, 3), 1.0)
    assert stored is not None and stored.shape == film.shape and (stored.reshape(-1, 3) == want_bgr8).all()


def test_film_stays_on_the_device_for_a_denoiser(qz):
    """AOV hand-off (RenderResult::denoise, image.cpp:47-95): after render() the three planes are still on the device
    (qz_film_device) and equal what render() returned; the tone kernel runs on them in place (qz_tone_device)."""
    import ctypes

    import torch
    from cuda.bindings import runtime as cudart

    lib = qz.lib
    with qz.build_scene("cornell_box", 48, 40) as sc:
        out = sc.render(spp=2, max_bounces=4)
        ptrs = [ctypes.c_void_p() for _ in range(3)]
        w, h = ctypes.c_uint32(), ctypes.c_uint32()
        rc = lib.qz_film_device(ctypes.c_void_p(sc.c_scene_handle()), ctypes.byref(ptrs[0]), ctypes.byref(ptrs[1]), ctypes.byref(ptrs[2]),
                                ctypes.byref(w), ctypes.byref(h))
        assert rc == 0 and (w.value, h.value) == (48, 40) and all(p.value for p in ptrs)
        n = 48 * 40
        planes = []
        for p in ptrs:
            host = np.zeros((40, 48, 3), np.float32)
            (err,) = cudart.cudaMemcpy(host.ctypes.data, p.value, n * 12, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
            assert int(err) == 0
            planes.append(host)
        assert bits_equal(planes[0], out.color).all() and bits_equal(planes[1], out.normal).all() and bits_equal(planes[2], out.albedo).all()
        bgr8 = torch.zeros(n * 3, dtype=torch.uint8, device="cuda")
        assert lib.qz_tone_device(ptrs[0], n, ctypes.c_float(1.0), None, ctypes.c_void_p(bgr8.data_ptr()), None) == 0
        torch.cuda.synchronize()
        want = np.clip(np.rint(255.0 * out.color[..., ::-1].astype(np.float32)), 0, 255).astype(np.uint8)
        assert (bgr8.cpu().numpy().reshape(40, 48, 3) == want).all()


@pytest.mark.parametrize("name,w,h,spp,bounces,spp_pass", [("cornell_box", 200, 136, 6, 12, 0), ("cornell_box", 96, 80, 7, 40, 3),
                                                            ("textures", 72, 64, 4, 8, 0), ("opposing_planes", 96, 54, 4, 24, 0),
                                                            ("mandelbrot", 64, 48, 4, 8, 2)])
def test_sample_memo_does_not_change_the_film(qz, mode, name, w, h, spp, bounces, spp_pass):
    """The per-pass sample memo (csrc/sampler.cuh: sampler values and hot spectra tabulated per Halton index, read by the
    shading kernels instead of running the digit loops / spectrum lookups per bounce) is a pure memoisation: forced on
    (QZ_FLAG_FORCE_MEMO; real renders enable it from 65536 pixels) and forced off, the film must be the same bit for bit --
    single and multiple passes, bounces within and beyond the memo's eight, flat and BVH scenes."""
    from quetzalcoatlus_b200.harness import QZ_FLAG_FORCE_MEMO, QZ_FLAG_NO_MEMO

    with qz.build_scene(name, w, h) as sc:
        on, st_on = sc.render_flags(spp, bounces, flags=QZ_FLAG_FORCE_MEMO, samples_per_pass=spp_pass)
        off, st_off = sc.render_flags(spp, bounces, flags=QZ_FLAG_NO_MEMO, samples_per_pass=spp_pass)
    assert bits_equal(on.color, off.color).all() and bits_equal(on.normal, off.normal).all() and bits_equal(on.albedo, off.albedo).all()
    assert (st_on["rays_closest"], st_on["rays_shadow"], st_on["shade_calls"]) == (st_off["rays_closest"], st_off["rays_shadow"], st_off["shade_calls"])


def test_in_library_multi_gpu_render_is_bit_identical(qz):
    """Multi-GPU behind render() (qz_set_device_count / QZ_DEVICES): the scene is replicated, every device renders its
    interleaved strips and writes them into device 0's film over peer memory; the film must equal the one-GPU film bit
    for bit, for a flat scene and for a BVH scene.  Needs two visible GPUs (gpurun --gpus 2)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = min(torch.cuda.device_count(), 4)
    for name, w, h in [("cornell_box", 96, 80), ("mandelbrot", 64, 48)]:
        with qz.build_scene(name, w, h) as sc:
            want = sc.render(spp=4, max_bounces=8)
        for flags in (0, 128):   # 128 = QZ_FLAG_FORCE_MEMO: every device tabulates the pixel classes of its own strips
            qz.lib.qz_set_device_count(n)
            old = qz.set_default_flags(flags)
            try:
                with qz.build_scene(name, w, h) as sc:
                    got = sc.render(spp=4, max_bounces=8)
                    stats = sc.last_stats()
            finally:
                qz.lib.qz_set_device_count(0)
                qz.set_default_flags(old)
            assert stats["paths"] == w * h * 4
            assert bits_equal(got.color, want.color).all() and bits_equal(got.normal, want.normal).all() and bits_equal(got.albedo, want.albedo).all()


def test_error_conventions(qz):
    import ctypes

    from quetzalcoatlus_b200.harness import QzCamera

    lib = qz.lib
    lib.qz_last_error.restype = ctypes.c_char_p
    scene = ctypes.c_void_p()
    assert lib.qz_scene_create(ctypes.byref(scene)) == 0
    cam = QzCamera()
    out = np.zeros(3, np.float32)
    # render on an uncommitted scene: the reference prints and returns an empty image (render.cpp:328-331)
    assert lib.qz_render(scene, ctypes.byref(cam), 1, 1, None, None, out.ctypes.data_as(ctypes.c_void_p), None, None, None) == 3
    assert b"committed" in lib.qz_last_error()
    assert lib.qz_scene_destroy(scene) == 0
    # more samples than the sampler's index arithmetic defines: the reference computes `sample_index * sample_stride` in
    # int (sampler.cpp:419), which overflows past 2^31 -- stride 31104 from 128 x 128 pixels on, i.e. beyond 69042 samples
    with qz.build_scene("cornell_box", 128, 128) as sc:
        with pytest.raises(RuntimeError, match="Halton"):
            sc.render_flags(69043)
