"""Kernel-point dispositions for KPConv.

Same contract as the reference's ``load_kernels(radius, num_kpoints, dimension, fixed)``
(models/backbone_kpconv/kernels/kernel_points.py:387-469): take a unit-sphere disposition (read from
``kernels/dispositions/k_{K:03d}_{fixed}_{D}D.ply`` relative to the working directory when that file
exists, as the reference does), add N(0, 0.01) noise, scale by ``radius`` and rotate about z by a
random angle drawn from numpy's global RNG.  When no file is present the disposition is generated
here by a small repulsion solve (deterministic), instead of the reference's 100-candidate optimiser.

Kernel points end up in ``KPConv.kernel_points`` (a non-trainable Parameter stored in checkpoints),
so loading a reference checkpoint reproduces the reference's kernel points exactly.
"""
from __future__ import annotations

import os
from functools import lru_cache

import numpy as np


def _read_ply_xyz(path: str) -> np.ndarray:
    """Minimal reader for the binary little-endian vertex-only PLY files KPConv stores."""
    with open(path, "rb") as fh:
        raw = fh.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode("ascii", errors="replace").splitlines()
    if not any(line.startswith("format binary_little_endian") for line in header):
        raise ValueError(f"{path}: only binary_little_endian PLY is supported")
    n_vertex, props = 0, []
    for line in header:
        tok = line.split()
        if tok[:2] == ["element", "vertex"]:
            n_vertex = int(tok[2])
        elif tok and tok[0] == "property":
            props.append((tok[2], {"float64": "<f8", "double": "<f8", "float32": "<f4", "float": "<f4"}[tok[1]]))
    data = np.frombuffer(raw[end:], dtype=np.dtype(props), count=n_vertex)
    return np.stack([data["x"], data["y"], data["z"]], 1).astype(np.float64)


@lru_cache(maxsize=None)
def _generated_disposition(num_kpoints: int, dimension: int, fixed: str) -> np.ndarray:
    """Points in the unit ball that repel each other (1/d potential) and are pulled to the centre;
    with fixed='center' point 0 stays at the origin.  Rescaled so the mean radius of the free points
    is 0.66 (the KPConv convention, kernel_points.py:380-381)."""
    rng = np.random.default_rng(20240 + 131 * num_kpoints + dimension)
    pts = rng.normal(size=(num_kpoints, dimension))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    pts *= rng.uniform(0.3, 0.9, size=(num_kpoints, 1))
    if fixed == "center":
        pts[0] = 0.0
    step = 0.05
    for it in range(4000):
        diff = pts[:, None, :] - pts[None, :, :]
        dist = np.sqrt((diff ** 2).sum(-1)) + np.eye(num_kpoints)
        grad = -(diff / dist[..., None] ** 3).sum(1) + 2.0 * pts  # repulsion + quadratic well
        if fixed == "center":
            grad[0] = 0.0
        norm = np.linalg.norm(grad, axis=1, keepdims=True)
        pts -= step * grad / np.maximum(norm, 1e-9) * np.minimum(norm, 1.0)
        if it % 500 == 499:
            step *= 0.6
    radii = np.linalg.norm(pts, axis=1)
    free = radii[1:] if fixed == "center" else radii
    return pts * (0.66 / free.mean())


def load_kernels(radius, num_kpoints, dimension, fixed, lloyd=False):
    """[num_kpoints, dimension] float32 kernel points scaled to ``radius`` (randomly z-rotated)."""
    kernel_file = os.path.join("kernels/dispositions", "k_{:03d}_{:s}_{:d}D.ply".format(num_kpoints, fixed, dimension))
    if os.path.exists(kernel_file):
        kernel_points = _read_ply_xyz(kernel_file)[:, :dimension]
    else:
        kernel_points = _generated_disposition(int(num_kpoints), int(dimension), str(fixed)).copy()

    rot = np.eye(dimension)
    theta = np.random.rand() * 2 * np.pi
    c, s = np.cos(theta), np.sin(theta)
    if dimension == 2 and fixed != "vertical":
        rot = np.array([[c, -s], [s, c]], dtype=np.float32)
    elif dimension == 3:
        if fixed != "vertical":
            rot = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)
        else:
            # random axis (theta, phi) and angle alpha: Rodrigues' formula
            phi = (np.random.rand() - 0.5) * np.pi
            axis = np.array([np.cos(theta) * np.cos(phi), np.sin(theta) * np.cos(phi), np.sin(phi)])
            alpha = np.random.rand() * 2 * np.pi
            kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
            rot = (np.eye(3) + np.sin(alpha) * kx + (1 - np.cos(alpha)) * kx @ kx).astype(np.float32)
    kernel_points = kernel_points + np.random.normal(scale=0.01, size=kernel_points.shape)
    kernel_points = radius * kernel_points
    return np.matmul(kernel_points, rot).astype(np.float32)
