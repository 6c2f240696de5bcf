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


def dragon_mesh(n_tri: int):
    """The obj_viewer stand-in: a bulky closed body that fills the obj_viewer frame the way the reference's dragon
    render does (examples/xyz_dragon_obj.png: a quarter to a third of the pixels).  A latitude-longitude grid over a
    displaced ellipsoid -- overlapping scale rows, a dorsal ridge of spikes, fine noise -- closed by two pole fans:
    nu rings x nv segments -> nu*nv + 2 vertices, 2*nv*(nu - 1) + 2*nv triangles."""
    nv = max(8, int(round(np.sqrt(n_tri / 2.0))))
    nu = max(4, n_tri // (2 * nv))
    eps = np.pi / (2.0 * nu)
    th = np.linspace(eps, np.pi - eps, nu)            # polar angle from +x (the body's long axis)
    ph = np.linspace(0.0, 2.0 * np.pi, nv, endpoint=False)
    tt, pp = np.meshgrid(th, ph, indexing="ij")

    def surface(t, p):
        scales = 0.045 * np.abs(np.sin(19.0 * t + 0.5 * np.sin(14.0 * p))) * np.abs(np.sin(14.0 * p))
        ridge = 0.32 * np.exp(-((p - 0.5 * np.pi) / 0.22) ** 2) * (0.55 + 0.45 * np.sin(23.0 * t)) * np.sin(t) ** 2
        bulge = 0.10 * np.sin(3.0 * t) * np.cos(2.0 * p) + 0.06 * np.sin(5.0 * t + 1.3) * np.sin(3.0 * p + 0.4)
        fine = 0.012 * np.sin(61.0 * t + 11.0 * p) + 0.008 * np.sin(97.0 * p - 29.0 * t)
        r = 1.0 + scales + ridge + bulge + fine
        d = np.stack([np.cos(t), np.sin(t) * np.sin(p), np.sin(t) * np.cos(p)], -1)
        return d * r[..., None] * np.array([2.85, 1.9, 1.95]) + np.array([0.0, 2.2, 0.0])

    pos = surface(tt, pp)
    h = 1e-4
    dt = surface(tt + h, pp) - surface(tt - h, pp)
    dp = surface(tt, pp + h) - surface(tt, pp - h)
    nrm = np.cross(dp, dt)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=2, keepdims=True), 1e-12)
    poles = np.stack([surface(np.array(0.0), np.array(0.0)), surface(np.array(np.pi), np.array(0.0))])
    pole_n = np.array([[1.0, 0.0, 0.0], [-1.0, 0.0, 0.0]])
    idx = np.arange(nu)[:, None] * nv + np.arange(nv)[None, :]
    a, b = idx[:-1], idx[1:]
    c, d = np.roll(b, -1, 1), np.roll(a, -1, 1)
    tris = [np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)]
    p0, p1 = nu * nv, nu * nv + 1
    tris.append(np.stack([np.full(nv, p0), idx[0], np.roll(idx[0], -1)], -1))
    tris.append(np.stack([np.full(nv, p1), np.roll(idx[-1], -1), idx[-1]], -1))
    return np.concatenate([pos.reshape(-1, 3), poles]), np.concatenate([nrm.reshape(-1, 3), pole_n]), np.concatenate(tris, 0)


def write_obj(path: str, pos, nrm, tris) -> None:
    with open(path, "w") as f:
        f.write("# synthetic torus-knot mesh (tools/gen_mesh.py)\n")
        f.write("".join("v %.6f %.6f %.6f\n" % tuple(p) for p in pos))
        f.write("".join("vn %.6f %.6f %.6f\n" % tuple(n) for n in nrm))
        t1 = tris + 1
        f.write("".join("f %d//%d %d//%d %d//%d\n" % (a, a, b, b, c, c) for a, b, c in t1))


def ensure_obj(directory: str, n_tri: int = 1_000_000, shape: str = "dragon") -> str:
    """Path of the cached OBJ text of `shape` with ~n_tri triangles (written on first use)."""
    import os

    path = os.path.join(directory, f"qz_{shape}_{n_tri}.obj")
    if not os.path.exists(path):
        pos, nrm, tris = (knot_mesh if shape == "knot" else dragon_mesh)(n_tri)
        write_obj(path + ".tmp", pos, nrm, tris)
        os.replace(path + ".tmp", path)
    return path


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--triangles", type=int, default=1_000_000)
    ap.add_argument("--shape", choices=["dragon", "knot"], default="dragon")
    args = ap.parse_args()
    pos, nrm, tris = (knot_mesh if args.shape == "knot" else dragon_mesh)(args.triangles)
    write_obj(args.out, pos, nrm, tris)
    print(f"{args.out}: {len(pos)} vertices, {len(tris)} triangles")


if __name__ == "__main__":
    main()
