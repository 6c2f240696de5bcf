"""Helpers shared by the parity tests."""
import numpy as np

ANALYTIC_SCENES = ["cornell_box", "glass_spheres", "textures", "opposing_planes", "cornell_mixed", "mandelbrot", "kitchen_sink"]

# scenes for the branches no shipped example reaches (no golden fixtures: compared with the oracle directly)
EDGE_SCENES = ["no_lights", "empty_sky"]

# record layout of trace_paths (include/qz_b200.h)
LAMBDA, PDF, RADIANCE, NORMAL, RAYS, ALBEDO, RGB, ARGB = (slice(0, 4), slice(4, 8), slice(8, 12), slice(12, 15), 15,
                                                          slice(16, 20), slice(20, 23), slice(23, 26))


def pixel_samples(scene, n, seed, spp=None):
    """n random (x, y, s) with y the sampler/camera y."""
    rng = np.random.default_rng(seed)
    spp = scene.default_spp if spp is None else spp
    return np.stack([rng.integers(0, scene.width, n), rng.integers(0, scene.height, n), rng.integers(0, spp, n)], 1).astype(np.int32)


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    return a == b


def path_agreement(ref, got, rel=1e-4):
    """Per-path agreement under the north-star tolerance: radiance within `rel` relative (with an
    absolute floor of rel * the scene's radiance scale), same ray count.  Returns the boolean mask."""
    scale = max(float(np.abs(ref[:, RADIANCE]).max()), 1e-6)
    tol = rel * np.maximum(np.abs(ref[:, RADIANCE]), 1e-2 * scale)
    ok = (np.abs(ref[:, RADIANCE] - got[:, RADIANCE]) <= tol).all(1)
    ok &= ref[:, RAYS] == got[:, RAYS]
    return ok
