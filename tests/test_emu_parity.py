"""Device-code restatement (csrc/*.cuh compiled for the host, tests/emu) against the oracle
and the golden fixtures: BIT-EXACT, every field of every record.  In this build the
transcendental functions are the host libm's, so nothing is left to tolerance; what the GPU
adds on top (CUDA's double-rounded sin/cos/atan2/acos, the wavefront scheduling) is checked
by the -m gpu tests."""
from pathlib import Path

import numpy as np
import pytest

from common import ANALYTIC_SCENES, EDGE_SCENES, bits_equal, pixel_samples

GOLDEN = Path(__file__).resolve().parent / "golden"

SPECTRA = ["X", "Y", "Z", "D65", "CANON_R", "CANON_G", "CANON_B", "AL_IOR", "AL_ABSORPTION", "CU_IOR", "CU_ABSORPTION",
           "GLASS_BK7_IOR", "GLASS_SF11_IOR", "rgb:0.8,0.4,0.1", "rgb:0.5,0.5,0.5", "rgb:0,0,0", "rgb:1,1,1", "rgbu:1.1,1.8,3.0",
           "rgbi:8,2,4", "rgbi:2,1,2", "const:1.5"]


@pytest.mark.parametrize("fixture", ["sampler_kat.npz", "sampler_kat_800.npz"])
def test_sampler_bit_exact(emu, fixture):
    g = np.load(GOLDEN / fixture)
    spp, w, h = (int(v) for v in g["res"])
    assert bits_equal(emu.sampler_eval(spp, w, h, g["q"]), g["values"]).all()


def test_sampler_small_images_and_dimension_wrap(emu, oracle):
    rng = np.random.default_rng(3)
    for w, h in [(1, 1), (3, 2), (100, 81), (127, 129), (4096, 2160)]:
        q = np.stack([rng.integers(0, w, 500), rng.integers(0, h, 500), rng.integers(0, 64, 500), rng.integers(0, 1000, 500)], 1)
        assert bits_equal(emu.sampler_eval(64, w, h, q), oracle.sampler_eval(64, w, h, q)).all(), (w, h)


def test_sampler_up_to_the_largest_defined_sample_number(emu, oracle):
    """The reference computes `sample_index * sample_stride` in int (sampler.cpp:419): defined up to 69042 samples at the
    stride of 31104 that images of 128 x 128 pixels and more have.  Bit-exact right up to that, refused beyond it (by
    the probes here, by qz_render for the sample count: tests/test_gpu_parity.py::test_error_conventions)."""
    rng = np.random.default_rng(5)
    for w, h in [(128, 128), (800, 800), (3840, 2160)]:
        q = np.stack([rng.integers(0, w, 400), rng.integers(0, h, 400), rng.integers(68000, 69042, 400), rng.integers(0, 1000, 400)], 1)
        q[0, 2] = 69041
        assert bits_equal(emu.sampler_eval(69042, w, h, q), oracle.sampler_eval(69042, w, h, q)).all(), (w, h)
        q[0, 2] = 69042
        with pytest.raises(RuntimeError):
            emu.sampler_eval(69043, w, h, q)
    # small images have a smaller stride (8 x 9 = 72 for 8 x 8 pixels) and a correspondingly larger range
    q = np.stack([rng.integers(0, 8, 400), rng.integers(0, 8, 400), rng.integers(0, 2 ** 31 // 72 - 1, 400), rng.integers(0, 1000, 400)], 1)
    assert bits_equal(emu.sampler_eval(1, 8, 8, q), oracle.sampler_eval(1, 8, 8, q)).all()


@pytest.mark.parametrize("name", SPECTRA)
def test_spectra_bit_exact(emu, oracle, name):
    lam = np.concatenate([np.random.default_rng(1).uniform(355, 835, 3000), np.arange(355, 836, dtype=np.float64)]).astype(np.float32)
    assert bits_equal(emu.eval_spectrum(name, lam), oracle.eval_spectrum(name, lam)).all()


@pytest.mark.parametrize("name", ANALYTIC_SCENES)
def test_paths_match_golden_bit_exact(emu, name):
    g = np.load(GOLDEN / f"paths_{name}.npz")
    with emu.build_scene(name) as sc:
        got = sc.trace_paths(g["xys"])
    bad = ~bits_equal(got, g["records"]).all(1)
    assert not bad.any(), f"{bad.sum()} of {len(bad)} paths differ, first at {g['xys'][np.argmax(bad)]}"


@pytest.mark.parametrize("name", ANALYTIC_SCENES)
def test_flatten_camera_sensor_intersection(emu, oracle, name):
    rng = np.random.default_rng(11)
    with emu.build_scene(name) as se, oracle.build_scene(name) as so:
        assert bits_equal(se.camera_fields(), so.camera_fields()).all()
        ul = rng.uniform(0, 1, (1000, 5)).astype(np.float32)
        ul[:, 1:] *= 60.0  # some above the saturation clamp
        assert bits_equal(se.sensor_eval(ul), so.sensor_eval(ul)).all()
        o = np.tile(so.camera_fields()[0], (5000, 1)) + rng.normal(0, 0.5, (5000, 3))
        d = rng.normal(0, 1, (5000, 3))
        d[:, 2] -= 1.5
        rays = np.concatenate([o, d], 1).astype(np.float32)
        assert bits_equal(se.intersect(rays), so.intersect(rays)).all()


@pytest.mark.parametrize("material,light", [("alluminum", "point"), ("glass", "area"), ("diffuse", "ambient")])
def test_mesh_scene_bit_exact(emu, oracle, small_mesh, material, light):
    """obj_viewer on the synthetic mesh: OBJ loader, smooth normals, LBVH, hoisted planes."""
    kw = dict(obj_path=small_mesh, obj_material=material, obj_light=light)
    with emu.build_scene("obj_viewer", **kw) as se, oracle.build_scene("obj_viewer", **kw) as so:
        xys = pixel_samples(so, 1500, seed=4)
        assert bits_equal(se.trace_paths(xys), so.trace_paths(xys)).all()


def test_film_matches_oracle_bit_exact(emu):
    """Pixel loop + ordered accumulation + division (render.cpp:260-294) on a small film."""
    for name in ["cornell_box", "textures", "kitchen_sink"]:
        g = np.load(GOLDEN / f"film_{name}.npz")
        h, w, _ = g["color"].shape
        with emu.build_scene(name, w, h) as sc:
            r = sc.render(int(g["spp"]))
        for plane in ("color", "normal", "albedo"):
            assert bits_equal(getattr(r, plane), g[plane]).all(), (name, plane)


def test_ragged_and_edge_inputs(emu, oracle):
    with emu.build_scene("cornell_box", 5, 3) as se, oracle.build_scene("cornell_box", 5, 3) as so:
        assert se.trace_paths(np.zeros((0, 3), np.int32)).shape == (0, 32)
        xys = np.array([[0, 0, 0], [4, 2, 127], [2, 1, 5]], np.int32)
        assert bits_equal(se.trace_paths(xys), so.trace_paths(xys)).all()
        # max_bounces = 0: only directly visible emission
        assert bits_equal(se.trace_paths(xys, max_bounces=0), so.trace_paths(xys, max_bounces=0)).all()
        assert bits_equal(se.render(2).color, so.render(2).color).all()


@pytest.mark.parametrize("name", EDGE_SCENES)
def test_edge_scenes_bit_exact(emu, oracle, name):
    """A scene without lights (sample_lights' unused draws, render.cpp:62-66) and a scene without geometry."""
    with emu.build_scene(name) as se, oracle.build_scene(name) as so:
        xys = pixel_samples(so, 1500, seed=6)
        assert bits_equal(se.trace_paths(xys), so.trace_paths(xys)).all()
        a, b = se.render(2), so.render(2)
        assert bits_equal(a.color, b.color).all() and bits_equal(a.normal, b.normal).all() and bits_equal(a.albedo, b.albedo).all()


@pytest.mark.parametrize("max_bounces", [0, 1, 2])
def test_bounce_limits_bit_exact(emu, oracle, max_bounces):
    """depth == max_bounces breaks the path loop before the BSDF is built (render.cpp:137): limits 0, 1, 2."""
    for name in ("cornell_box", "kitchen_sink"):
        with emu.build_scene(name) as se, oracle.build_scene(name) as so:
            xys = pixel_samples(so, 800, seed=8)
            assert bits_equal(se.trace_paths(xys, max_bounces=max_bounces), so.trace_paths(xys, max_bounces=max_bounces)).all(), name


FAST_EMU_LIB = Path(__file__).resolve().parent / "emu" / "_build" / "libqz_emu_fast_harness.so"


@pytest.mark.parametrize("name", ANALYTIC_SCENES)
def test_fast_arithmetic_build_stays_within_tolerance(emu, oracle, name):
    """The radiometric ("fast") build of the device headers (csrc/common.cuh, ARITHMETIC MODES), compiled for the host:
    geometry and discrete decisions identical to the oracle's (ray counts, wavelengths, first-hit normals bit for
    bit), radiance within 1e-5 relative -- ten times tighter than the north-star tolerance.  (The host stands in for
    the MUFU approximations with exact operations; the GPU tests measure those.)"""
    from common import NORMAL, RAYS, path_agreement
    from quetzalcoatlus_b200.harness import Harness

    fast = Harness(FAST_EMU_LIB, "qzh_")
    with fast.build_scene(name) as sf, oracle.build_scene(name) as so:
        xys = pixel_samples(so, 3000, seed=33)
        got, want = sf.trace_paths(xys), so.trace_paths(xys)
    assert (got[:, RAYS] == want[:, RAYS]).all()
    assert bits_equal(got[:, :4], want[:, :4]).all() and bits_equal(got[:, NORMAL], want[:, NORMAL]).all()
    assert path_agreement(want, got, 1e-5).all()


def _tone_inputs():
    rng = np.random.default_rng(11)
    rgb = np.concatenate([rng.uniform(0.0, 1.2, (20000, 3)), rng.uniform(0.0, 40.0, (2000, 3)),
                          [[0.0, 1.0, 0.5], [1.0 / 255, 2.0 / 255, 0.5 / 255], [-0.25, 0.0, 1e-30]]]).astype(np.float32)
    return rgb


@pytest.mark.parametrize("gamma", [1.0, 0.45454547, 2.2])
def test_tone_path_restatement(emu, oracle, gamma):
    """Image::save's arithmetic (image.cpp:10-15) as restated in csrc/math.cuh (tone_value / tone_u8) against the oracle's
    libm expression: bit-exact for the default gamma of 1, within one ulp otherwise (double pow rounded once vs glibc powf)."""
    rgb = _tone_inputs()
    got_f, got_u8 = emu.tone(rgb, gamma)
    want_f, want_u8 = oracle.tone(rgb, gamma)
    if gamma == 1.0:
        assert bits_equal(got_f, want_f).all() and (got_u8 == want_u8).all()
    else:
        ok = np.isnan(want_f) & np.isnan(got_f) | (np.abs(got_f - want_f) <= 2 * np.spacing(np.abs(want_f)))
        assert ok.all()
        assert (np.abs(got_u8.astype(int) - want_u8.astype(int)) <= 1).all() and (got_u8 != want_u8).mean() < 1e-3


def _planes_equal(a, b, rows=slice(None)):
    return all(bits_equal(getattr(a, p)[rows], getattr(b, p)[rows]).all() for p in ("color", "normal", "albedo"))


@pytest.mark.parametrize("name,size,spp", [("cornell_box", (150, 140), 3), ("kitchen_sink", (131, 70), 2), ("textures", (140, 133), 2)])
def test_sample_memo_on_the_host_does_not_change_the_film(emu, name, size, spp):
    """SAMPLE MEMO (csrc/sampler.cuh, memo_plan.h) with the wavefront taken away: the emulation builds the table from the
    product's own pieces -- the owned-class tables, the row layout, memo_entry_bits (what k_memo_fill stores),
    memo_fill_hot (what k_memo_spectra runs) -- and the per-path loop reads draws and hot spectra through memo_row /
    bounce_in_memo exactly as k_shade does.  Films must be bit-identical with the table, without it, with rows that
    hold only two bounces, and with a table that covers only the first sample number (a later pass's samples)."""
    from quetzalcoatlus_b200.harness import QZ_FLAG_FORCE_MEMO

    w, h = size
    with emu.build_scene(name, w, h) as sc:
        plain, _ = sc.render_flags(spp)
        memo, st = sc.render_flags(spp, flags=QZ_FLAG_FORCE_MEMO)
        n_cls = min(w, 128) * min(h, 128)
        assert st["iterations"] == n_cls and st["shade_calls"] == n_cls * spp * (3 + 8 * 8)   # the table really was built
        assert _planes_equal(plain, memo)
        short, _ = sc.render_flags(spp, flags=QZ_FLAG_FORCE_MEMO, reserved=2)
        assert _planes_equal(plain, short)
        partial, st1 = sc.render_flags(spp, flags=QZ_FLAG_FORCE_MEMO, samples_per_pass=1, reserved=1)
        assert st1["shade_calls"] == n_cls * (3 + 8)
        assert _planes_equal(plain, partial)


def test_render_hands_back_large_films_whole(emu):
    """The host library's render() lets the C ABI write into a scratch area while a helper thread allocates the
    RenderResult, then copies the planes over in several ranges (host/src/scene.cpp): a film large enough for several
    ranges must equal the one the C ABI writes straight into caller buffers -- twice, the second call reusing the scratch."""
    with emu.build_scene("textures", 331, 257) as sc:
        direct, _ = sc.render_flags(1)
        for _ in range(2):
            assert _planes_equal(sc.render(1), direct)
    with emu.build_scene("cornell_box", 64, 48) as sc:   # a smaller film after a larger one: the scratch is not resized down
        direct, _ = sc.render_flags(2)
        assert _planes_equal(sc.render(2), direct)


def test_sample_memo_of_a_shard_holds_only_its_own_classes(emu):
    """A multi-GPU shard (interleaved strips, qz_region) tabulates only the (x mod 128, y mod 128) classes of its own rows:
    1/N of them when strip x N divides 128, and its rows of the film are those of the unsharded render, bit for bit."""
    from quetzalcoatlus_b200.harness import QZ_FLAG_FORCE_MEMO

    w, h, spp = 130, 256, 1
    with emu.build_scene("cornell_box", w, h) as sc:
        whole, _ = sc.render_flags(spp)
        for strip, n in [(4, 8), (8, 2), (5, 3)]:
            shard = n - 1
            part, st = sc.render_flags(spp, flags=QZ_FLAG_FORCE_MEMO, region=(strip, n, shard))
            own = np.array([(r // strip) % n == shard for r in range(h)])
            if 128 % (strip * n) == 0:
                assert st["iterations"] == 128 * 128 // n
            else:
                assert 128 * 128 // n < st["iterations"] <= 128 * 128
            assert _planes_equal(whole, part, own)
            assert not part.color[~own].any()


@pytest.mark.parametrize("what", ["dims", "hot"])
def test_sample_memo_negative_control(emu, monkeypatch, what):
    """The paths really read the table: entries off by 5e-4 relative (the draws, or the hot spectra) change the film."""
    from quetzalcoatlus_b200.harness import QZ_FLAG_FORCE_MEMO

    with emu.build_scene("cornell_box", 40, 30) as sc:
        plain, _ = sc.render_flags(2)
        monkeypatch.setenv("QZ_EMU_MEMO_CORRUPT", what)
        off, _ = sc.render_flags(2, flags=QZ_FLAG_FORCE_MEMO)
        monkeypatch.delenv("QZ_EMU_MEMO_CORRUPT")
        again, _ = sc.render_flags(2, flags=QZ_FLAG_FORCE_MEMO)
    assert _planes_equal(plain, again)
    lit = plain.color != 0   # (few pixels find the small light with two samples)
    differing = (~bits_equal(plain.color, off.color))[lit].mean()
    assert differing > 0.5, differing
    if what == "dims":
        assert (~bits_equal(plain.albedo, off.albedo)).mean() > 0.5
    else:
        assert _planes_equal(plain, off) is False and bits_equal(plain.normal, off.normal).all()   # geometry does not depend on the spectra


def test_paths_thousands_of_bounces_deep_bit_exact(emu, oracle):
    """kitchen_sink at max_bounces = 3000: lossless dielectrics keep a twentieth of the paths alive for thousands of rays,
    and the sampler's dimension counter wraps its 1000-prime table many times on the way (sampler.cpp:404-454)."""
    from common import RAYS

    with emu.build_scene("kitchen_sink", 64, 48) as se, oracle.build_scene("kitchen_sink", 64, 48) as so:
        xys = pixel_samples(so, 1500, seed=12)
        got, want = se.trace_paths(xys, max_bounces=3000), so.trace_paths(xys, max_bounces=3000)
    assert (want[:, RAYS] > 1000).sum() >= 10
    assert bits_equal(got, want).all()


@pytest.mark.parametrize("seed", range(24))
def test_random_scenes_bit_exact(emu, oracle, seed):
    """Seeded random scenes (scenes/scenes.hpp: build_fuzz -- the same generator compiled against the reference and against
    the host library): every material / texture / light / shape type in combinations the shipped examples do not have
    (anisotropic conductors, three-way mixed materials, blackbody and RGB emitters, two-sided area lights, no light at
    all).  Every field of every replayed path bit for bit -- including the paths whose radiance the reference itself
    drives to infinity (blackbody emitters).  600 seeds were run once this way; 24 stay in the suite."""
    name = f"fuzz:{seed}"
    with emu.build_scene(name) as se, oracle.build_scene(name) as so:
        assert bits_equal(se.camera_fields(), so.camera_fields()).all()
        xys = pixel_samples(so, 1200, seed=seed)
        got, want = se.trace_paths(xys), so.trace_paths(xys)
    bad = ~bits_equal(got, want).all(1)
    assert not bad.any(), f"{name}: {bad.sum()} of {len(bad)} paths differ, first at {xys[np.argmax(bad)]}"


@pytest.mark.parametrize("seed", range(100, 112))
def test_random_scenes_fast_arithmetic(oracle, seed):
    """The radiometric ("fast") build on the random scenes: discrete decisions identical (ray counts, wavelengths, normals,
    and which paths the reference drives to infinity), finite radiance within 1e-5 relative."""
    from common import NORMAL, RADIANCE, RAYS
    from quetzalcoatlus_b200.harness import Harness

    fast = Harness(FAST_EMU_LIB, "qzh_")
    name = f"fuzz:{seed}"
    with fast.build_scene(name) as sf, oracle.build_scene(name) as so:
        xys = pixel_samples(so, 1200, seed=seed)
        got, want = sf.trace_paths(xys), so.trace_paths(xys)
    assert (got[:, RAYS] == want[:, RAYS]).all()
    assert bits_equal(got[:, :4], want[:, :4]).all() and bits_equal(got[:, NORMAL], want[:, NORMAL]).all()
    g, w = got[:, RADIANCE].astype(np.float64), want[:, RADIANCE].astype(np.float64)
    assert (np.isfinite(g) == np.isfinite(w)).all() and (g[~np.isfinite(w)] == w[~np.isfinite(w)]).all()
    finite = np.isfinite(w).all(1)
    if finite.any():
        scale = max(np.abs(w[finite]).max(), 1e-6)
        assert (np.abs(g[finite] - w[finite]) <= 1e-5 * np.maximum(np.abs(w[finite]), 1e-2 * scale)).all()


@pytest.mark.parametrize("seed", range(200, 206))
def test_random_scene_films_bit_exact(emu, oracle, seed):
    """Pixel loop, ordered accumulation and division (render.cpp:260-294) on the random scenes: all three planes."""
    with emu.build_scene(f"fuzz:{seed}", 28, 20) as se, oracle.build_scene(f"fuzz:{seed}", 28, 20) as so:
        a, b = se.render(3), so.render(3)
    for plane in ("color", "normal", "albedo"):
        assert bits_equal(getattr(a, plane), getattr(b, plane)).all(), plane


def _random_mesh_obj(path, seed):
    """Triangle soups, coplanar overlapping triangles with exact duplicates (ties in t: the tie-break by ids decides),
    slivers next to huge triangles, and quad sheets -- what an LBVH over Morton codes finds hardest to keep in order."""
    rng = np.random.default_rng(seed)
    n, kind = int(rng.integers(1, 400)), seed % 4
    with open(path, "w") as f:
        if kind == 3:
            g = int(np.sqrt(n)) + 2
            xs, ys = np.meshgrid(np.linspace(-1.5, 1.5, g), np.linspace(-1.5, 1.5, g))
            for p in np.stack([xs, ys, 0.3 * np.sin(3 * xs) * np.cos(2 * ys)], -1).reshape(-1, 3):
                f.write("v %.6f %.6f %.6f\n" % tuple(p))
            for y in range(g - 1):
                for x in range(g - 1):
                    a = y * g + x + 1
                    f.write("f %d %d %d %d\n" % (a, a + 1, a + g + 1, a + g))
            return rng
        if kind == 0:
            tri = rng.normal(0, 1.0, (n, 1, 3)) + rng.normal(0, 0.3, (n, 3, 3))
        elif kind == 1:
            tri = np.concatenate([rng.uniform(-1.5, 1.5, (n, 3, 2)), np.zeros((n, 3, 1))], 2)
            tri[n // 2:] = tri[: n - n // 2]
        else:
            tri = rng.normal(0, 1.0, (n, 1, 3)) + rng.normal(0, 1.0, (n, 3, 3)) * rng.choice([1e-3, 1.0, 30.0], (n, 1, 1))
        for t in tri.reshape(-1, 3):
            f.write("v %.6f %.6f %.6f\n" % tuple(t))
        for i in range(len(tri)):
            f.write("f %d %d %d\n" % (3 * i + 1, 3 * i + 2, 3 * i + 3))
    return rng


@pytest.mark.parametrize("seed", range(12))
def test_random_meshes_bit_exact(emu, oracle, tmp_path, seed):
    """OBJ loader -> LBVH build -> wide-BVH traversal on random meshes against the oracle's brute-force-checked shim:
    closest hits (t, u, v, Ng, ids) and whole paths bit for bit.  300 seeds were run once this way; 12 stay in the suite."""
    path = tmp_path / "mesh.obj"
    rng = _random_mesh_obj(path, seed)
    kw = dict(obj_path=str(path), obj_material=["diffuse", "glass", "alluminum", "copper"][seed % 4], obj_light=["point", "area", "ambient"][seed % 3])
    with emu.build_scene("obj_viewer", 48, 36, **kw) as se, oracle.build_scene("obj_viewer", 48, 36, **kw) as so:
        o = np.tile(so.camera_fields()[0], (3000, 1)) + rng.normal(0, 0.4, (3000, 3))
        d = rng.normal(0, 1, (3000, 3))
        d[:, 2] -= 1.5
        rays = np.concatenate([o, d], 1).astype(np.float32)
        assert bits_equal(se.intersect(rays), so.intersect(rays)).all()
        xys = pixel_samples(so, 1000, seed=seed)
        assert bits_equal(se.trace_paths(xys), so.trace_paths(xys)).all()
