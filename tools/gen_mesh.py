#!/usr/bin/env python3
"""Deterministic synthetic OBJ mesh for the obj_viewer config.

No mesh ships with the reference (its dragon/bust OBJ files are not in the tree), so the
benchmark uses a closed, knotted tube with a displaced surface: a (p, q) torus knot swept by
a circle whose radius is modulated by a few fixed sine waves.  The surface is a
(nu x nv) periodic grid -> nu*nv vertices, 2*nu*nv triangles, written with `vn` normals and
`f a//na b//nb c//nc` faces so that the smooth-normal path (scene.cpp:85-98) is exercised.

    python tools/gen_mesh.py out.obj --triangles 1000000
"""
import argparse

import numpy as np


def knot_mesh(n_tri: int):
    nv = 256 if n_tri >= 200_000 else (64 if n_tri >= 10_000 else 16)
    nu = max(8, n_tri // (2 * nv))
    u = np.linspace(0.0, 2.0 * np.pi, nu, endpoint=False)
    v = np.linspace(0.0, 2.0 * np.pi, nv, endpoint=False)
    p, q = 2, 3
    # centre line of the knot, scaled to fit the obj_viewer camera (about 3 units across, resting near y = 1.6)
    r = 1.0 + 0.45 * np.cos(q * u)
    c = np.stack([r * np.cos(p * u), 0.45 * np.sin(q * u) * 1.4 + 1.6, r * np.sin(p * u)], 1)
    t = np.roll(c, -1, 0) - np.roll(c, 1, 0)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    up = np.array([0.0, 1.0, 0.0])
    n1 = np.cross(t, up)
    n1 /= np.linalg.norm(n1, axis=1, keepdims=True)
    n2 = np.cross(t, n1)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    rad = 0.28 * (1.0 + 0.18 * np.sin(7 * uu + 3 * vv) + 0.08 * np.sin(23 * uu - 5 * vv) + 0.04 * np.sin(61 * uu + 11 * vv))
    pos = c[:, None, :] + rad[..., None] * (np.cos(vv)[..., None] * n1[:, None, :] + np.sin(vv)[..., None] * n2[:, None, :])
    # vertex normals from central differences of the periodic grid
    du = np.roll(pos, -1, 0) - np.roll(pos, 1, 0)
    dv = np.roll(pos, -1, 1) - np.roll(pos, 1, 1)
    nrm = np.cross(dv, du)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=2, keepdims=True), 1e-12)
    idx = (np.arange(nu)[:, None] * nv + np.arange(nv)[None, :])
    a = idx
    b = np.roll(idx, -1, 0)
    cc = np.roll(np.roll(idx, -1, 0), -1, 1)
    d = np.roll(idx, -1, 1)
    tris = np.concatenate([np.stack([a, b, cc], -1).reshape(-1, 3), np.stack([a, cc, d], -1).reshape(-1, 3)], 0)
    return pos.reshape(-1, 3), nrm.reshape(-1, 3), tris


def write_obj(path: str, pos, nrm, tris) -> None:
    with open(path, "w") as f:
        f.write("# synthetic torus-knot mesh (tools/gen_mesh.py)\n")
        f.write("".join("v %.6f %.6f %.6f\n" % tuple(p) for p in pos))
        f.write("".join("vn %.6f %.6f %.6f\n" % tuple(n) for n in nrm))
        t1 = tris + 1
        f.write("".join("f %d//%d %d//%d %d//%d\n" % (a, a, b, b, c, c) for a, b, c in t1))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--triangles", type=int, default=1_000_000)
    args = ap.parse_args()
    pos, nrm, tris = knot_mesh(args.triangles)
    write_obj(args.out, pos, nrm, tris)
    print(f"{args.out}: {len(pos)} vertices, {len(tris)} triangles")


if __name__ == "__main__":
    main()
