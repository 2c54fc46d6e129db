"""Seeded synthetic point-cloud pairs shaped like the reference's datasets.

There is no network for 3DMatch / ModelNet40 / MCD, so the benchmark and the parity tests run on
synthetic clouds of the same size, density and extent.  All clouds are in generic position
(continuous random coordinates, no duplicated points or lattices) because the reference's order
for exactly-tied neighbour distances is unspecified (unstable sort).
"""
from __future__ import annotations

import numpy as np


def random_pose(rng: np.random.Generator, max_deg: float = 45.0, max_t: float = 0.5) -> np.ndarray:
    """[3,4] rigid transform: rotation by <= max_deg about a random axis, |t|_inf <= max_t."""
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(0, max_deg))
    kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    rot = np.eye(3) + np.sin(ang) * kx + (1 - np.cos(ang)) * kx @ kx
    return np.concatenate([rot, rng.uniform(-max_t, max_t, size=(3, 1))], 1).astype(np.float32)


def voxel_barycentres(pts: np.ndarray, dl: float) -> np.ndarray:
    """Order-free voxel-grid barycentres (input pre-processing only, like the datasets' 2.5 cm grid)."""
    ijk = np.floor(pts / dl).astype(np.int64)
    ijk -= ijk.min(0)
    dims = ijk.max(0) + 1
    key = ijk[:, 0] + dims[0] * (ijk[:, 1] + dims[1] * ijk[:, 2])
    _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    out = np.zeros((cnt.shape[0], 3), np.float64)
    np.add.at(out, inv, pts.astype(np.float64))
    return (out / cnt[:, None]).astype(np.float32)


def _apply(pose: np.ndarray, pts: np.ndarray) -> np.ndarray:
    return (pts.astype(np.float64) @ pose[:, :3].T.astype(np.float64) + pose[:, 3].astype(np.float64)).astype(np.float32)


def _room_surface(rng, n, patches, noise):
    k = patches.shape[0]
    which = rng.integers(0, k, size=n)
    uv = rng.uniform(-0.5, 0.5, size=(n, 2))
    p = patches[which, 0] + uv[:, :1] * patches[which, 1] + uv[:, 1:] * patches[which, 2]
    return p + rng.normal(scale=noise, size=p.shape)


def threedmatch_pair(seed: int, n_raw: int = 36000, voxel: float = 0.025, thickness: float = 0.004):
    """~20 k points per cloud after a 2.5 cm voxel grid: six 1.6 m planar patches in a 2.5 m room.

    Returns (src [Ns,3] f32, tgt [Nt,3] f32, pose [3,4] f32 with tgt ~= R src + t).
    """
    rng = np.random.default_rng(seed)
    patches = []
    for _ in range(6):
        c = rng.uniform(-0.45, 0.45, size=3)
        a = rng.normal(size=3)
        a /= np.linalg.norm(a)
        b = np.cross(a, rng.normal(size=3))
        b /= np.linalg.norm(b)
        patches.append(np.stack([c, 1.6 * a, 1.6 * b]))
    patches = np.stack(patches)
    pose = random_pose(rng, 45.0, 0.5)
    src = voxel_barycentres(_room_surface(rng, n_raw, patches, thickness), voxel)
    tgt = voxel_barycentres(_apply(pose, _room_surface(rng, n_raw, patches, thickness)), voxel)
    return src, tgt, pose


def modelnet_pair(seed: int, n: int = 717):
    """2 x 717 points on a random smooth closed surface inside [-1,1]^3 (+ clipped N(0,0.01) jitter)."""
    rng = np.random.default_rng(seed)
    coef = rng.normal(scale=0.12, size=(4, 3))

    def shape(m):
        d = rng.normal(size=(m, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        rad = 0.7 + d @ coef[0] + (d ** 2) @ coef[1] + np.sin(3 * d) @ coef[2] + np.cos(2 * d) @ coef[3]
        p = d * np.clip(rad, 0.3, 1.0)[:, None] * np.array([1.0, 0.8, 0.6])
        return p + np.clip(rng.normal(scale=0.01, size=p.shape), -0.05, 0.05)

    pose = random_pose(rng, 45.0, 0.5)
    return shape(n).astype(np.float32), _apply(pose, shape(n)), pose


def mcd_pair(seed: int, n: int = 120000):
    """2 x 120 k LiDAR-like points: ground disc (r < 40 m, sqrt-uniform) + 10 wall strips, 1 cm noise."""
    rng = np.random.default_rng(seed)
    walls = [(rng.uniform(-30, 30, size=2), rng.uniform(0, np.pi), rng.uniform(5, 20)) for _ in range(10)]

    def scan(m):
        n_ground = int(0.6 * m)
        rad = 40.0 * np.sqrt(rng.uniform(size=n_ground))
        ang = rng.uniform(0, 2 * np.pi, size=n_ground)
        parts = [np.stack([rad * np.cos(ang), rad * np.sin(ang), np.zeros(n_ground)], 1)]
        per = (m - n_ground) // len(walls)
        for i, (c, th, length) in enumerate(walls):
            k = per if i < len(walls) - 1 else m - n_ground - per * (len(walls) - 1)
            u = rng.uniform(-0.5, 0.5, size=k) * length
            parts.append(np.stack([c[0] + u * np.cos(th), c[1] + u * np.sin(th), rng.uniform(0, 4, size=k)], 1))
        p = np.concatenate(parts, 0)
        return p + rng.normal(scale=0.01, size=p.shape)

    pose = random_pose(rng, 10.0, 1.0)
    return scan(n).astype(np.float32), _apply(pose, scan(n)), pose


def kabsch_inputs(seed: int, n_sets: int = 6, n_pts: int = 1200, noise: float = 0.01):
    """Correspondence sets shaped like RegTR's decoder output (SURVEY.md §8d 'Kabsch inputs').

    Returns (a [n_sets,n_pts,3], b = R a + t + N(0,noise), w = sigmoid(N(0,2)), pose [3,4]).
    """
    rng = np.random.default_rng(seed)
    pose = random_pose(rng, 45.0, 0.5)
    a = rng.uniform(-1.25, 1.25, size=(n_sets, n_pts, 3)).astype(np.float32)
    b = (a.astype(np.float64) @ pose[:, :3].T + pose[:, 3] + rng.normal(scale=noise, size=a.shape)).astype(np.float32)
    w = (1.0 / (1.0 + np.exp(-rng.normal(scale=2.0, size=(n_sets, n_pts))))).astype(np.float32)
    return a, b, w, pose
