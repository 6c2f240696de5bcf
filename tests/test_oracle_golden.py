"""The oracle against the only known answers the reference itself holds, and against the
committed fixtures generated from it (tests/golden/make_golden.py)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from common import ANALYTIC_SCENES, bits_equal

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = Path(__file__).resolve().parent / "golden"

# values printed by the reference's color_test built from unmodified sources
COLOR_TEST_EXPECTED = """whitepoint: 0.3367 0.357918
0.333314 0.333288
0.771653 0.470475 0.980432
0.769878 0.544533 0.671843
0.446457 0.594196 0.175637
0.758855 0.589091 1.49932
0.656894 0.511026 0.176181
0.609913 0.599273 1.37064""".splitlines()


def test_sampler_known_answers_from_survey(oracle):
    # Sampler(4, 800, 800, 0), start_pixel_sample(10, 20, 1): sample_pixel() then dims 2..5
    q = [[10, 20, 1, d] for d in range(6)]
    got = oracle.sampler_eval(4, 800, 800, q)
    want = np.array([0.400390625, 0.971193612, 0.237410799, 0.619343162, 0.398415118, 0.742496133], np.float32)
    assert bits_equal(got, want).all()


def test_reference_color_test_output(oracle):
    exe = ROOT / "oracle" / "_ref" / "color_test"
    if not exe.exists():
        pytest.skip("oracle/_ref/color_test not built")
    out = subprocess.run([str(exe)], cwd=ROOT / "quetzalcoatlus_b200" / "data", capture_output=True, text=True, check=True).stdout
    lines = [l.strip() for l in out.splitlines() if l.strip() and not l.startswith("Should") and not l.startswith("Sensor")]
    for want in COLOR_TEST_EXPECTED:
        assert want.strip() in lines, f"color_test no longer prints {want!r}"


def test_shipped_rgb_to_spectrum_table_is_what_the_reference_optimiser_writes(tmp_path):
    """data/coeffs_SRGB_32.dat must be byte for byte what rgb_to_spectrum_opt.cpp:788-898 produces.
    The reference's color_test, built from unmodified sources, runs the optimiser when its working directory holds
    no cached table (get_coeffs, :873-898) and writes the file there."""
    exe = ROOT / "oracle" / "_ref" / "color_test"
    if not exe.exists():
        pytest.skip("oracle/_ref/color_test not built")
    subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, text=True, check=True, timeout=600)
    fresh = (tmp_path / "coeffs_SRGB_32.dat").read_bytes()
    shipped = (ROOT / "quetzalcoatlus_b200" / "data" / "coeffs_SRGB_32.dat").read_bytes()
    assert len(fresh) == (32 + 3 * 32 ** 3 * 3) * 4
    assert fresh == shipped


@pytest.mark.parametrize("fixture", ["sampler_kat.npz", "sampler_kat_800.npz"])
def test_oracle_reproduces_sampler_fixture(oracle, fixture):
    g = np.load(GOLDEN / fixture)
    spp, w, h = (int(v) for v in g["res"])
    assert bits_equal(oracle.sampler_eval(spp, w, h, g["q"]), g["values"]).all()


@pytest.mark.parametrize("name", ANALYTIC_SCENES)
def test_oracle_reproduces_path_fixture(oracle, name):
    g = np.load(GOLDEN / f"paths_{name}.npz")
    with oracle.build_scene(name) as sc:
        assert bits_equal(sc.trace_paths(g["xys"]), g["records"]).all()


def test_shim_bvh_equals_its_brute_force(oracle):
    """The shim's BVH must give the brute-force answer it defines (t, u, v, Ng, ids bit-identical)."""
    rng = np.random.default_rng(5)
    for name in ["glass_spheres", "mandelbrot"]:
        with oracle.build_scene(name) as bvh:
            o = np.tile(bvh.camera_fields()[0], (4000, 1)) + rng.normal(0, 0.5, (4000, 3))
            d = rng.normal(0, 1, (4000, 3))
            d[:, 2] -= 1.5
            d[:100] = [0.0, 0.0, -1.0]  # axis-parallel rays too
            rays = np.concatenate([o, d], 1).astype(np.float32)
            with_bvh = bvh.intersect(rays)
        oracle._fn("force_brute_force")(1)
        try:
            with oracle.build_scene(name) as brute:
                without = brute.intersect(rays)
        finally:
            oracle._fn("force_brute_force")(0)
        assert bits_equal(with_bvh, without).all()
        assert (with_bvh[:, 0] >= 0).mean() > 0.05
