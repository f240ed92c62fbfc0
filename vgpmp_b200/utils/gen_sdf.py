"""Mesh -> `.sdf` grid (reference: gpflow_vgpmp/utils/gen_sdf.py:16-43 shells out to an external SDFGen binary with
(mesh, delta, padding); the produced grids are missing from the reference snapshot).

`mesh_to_sdf` parses a Wavefront OBJ whose `o` groups are convex pieces (the reference's scene meshes:
data/scenes/bookshelves/bookshelves_center.obj, data/scenes/industrial/industrial-acd.obj), and evaluates the signed
distance on the GPU (`vgpmp_mesh_to_sdf`, csrc/mesh_sdf.cu).  Grid geometry follows SDFGen: the box is the mesh bounding
box grown by `padding` cells, nodes at origin + (i,j,k)*delta.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np
import torch

from .. import _cabi
from .sdf_utils import SignedDistanceField

__all__ = ["PADDING", "load_obj_convex_pieces", "mesh_to_sdf", "scene_mesh_path"]

PADDING = 20  # gen_sdf.py:9


def scene_mesh_path(name: str) -> Path:
    """Scene meshes shipped with the package (copied as data by tools/lift_reference_data.py)."""
    files = {"bookshelves": "bookshelves_center.obj", "industrial": "industrial-acd.obj"}
    return Path(__file__).resolve().parents[1] / "data" / "scenes" / files[name]


def load_obj_convex_pieces(path):
    """-> tri [T,3,3], plane [T,4] (outward n, d), piece_end [pieces].  Polygons are fanned into triangles; each face
    normal is oriented away from its piece's centroid (the pieces are convex, so this is the outward side)."""
    verts, pieces, cur = [], [], None
    with open(path) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                verts.append([float(t) for t in tok[1:4]])
            elif tok[0] in ("o", "g"):
                cur = []
                pieces.append(cur)
            elif tok[0] == "f":
                if cur is None:
                    cur = []
                    pieces.append(cur)
                idx = [int(t.split("/")[0]) for t in tok[1:]]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    cur.append((idx[0], idx[k], idx[k + 1]))
    V = np.asarray(verts, dtype=np.float64)
    tris, planes, ends = [], [], []
    for faces in pieces:
        if not faces:
            continue
        F = np.asarray(faces, dtype=np.int64)
        T = V[F]                                              # [f,3,3]
        centroid = V[np.unique(F)].mean(axis=0)
        n = np.cross(T[:, 1] - T[:, 0], T[:, 2] - T[:, 0])
        area2 = np.linalg.norm(n, axis=1)
        keep = area2 > 1e-14                                  # drop degenerate slivers
        T, n = T[keep], n[keep] / area2[keep, None]
        flip = np.einsum("fi,fi->f", n, T[:, 0] - centroid) < 0
        n[flip] *= -1.0
        d = np.einsum("fi,fi->f", n, T[:, 0])
        tris.append(T)
        planes.append(np.concatenate([n, d[:, None]], axis=1))
        ends.append((ends[-1] if ends else 0) + len(T))
    return np.concatenate(tris), np.concatenate(planes), np.asarray(ends, dtype=np.int32)


def grid_geometry(tri, delta, padding):
    lo, hi = tri.reshape(-1, 3).min(axis=0), tri.reshape(-1, 3).max(axis=0)
    origin = lo - padding * delta
    shape = np.floor((hi + padding * delta - origin) / delta).astype(np.int64) + 1
    return origin, tuple(int(v) for v in shape)


def mesh_to_sdf(obj_path, delta: float, padding: int = PADDING, device=None, origin=None, shape=None) -> SignedDistanceField:
    """`origin` + `shape` (both or neither) replace the SDFGen geometry (mesh bounding box grown by `padding` cells) with an
    explicit grid, e.g. the 512^3 grid around the robot's reach volume of BASELINE config 5."""
    if not torch.cuda.is_available():
        raise _cabi.VgpmpError("mesh_to_sdf runs on the GPU (vgpmp_mesh_to_sdf); there is no CPU fallback")
    tri, plane, piece_end = load_obj_convex_pieces(obj_path)
    if (origin is None) != (shape is None):
        raise ValueError("give both origin and shape, or neither")
    if origin is None:
        origin, shape = grid_geometry(tri, delta, padding)
    else:
        origin, shape = np.asarray(origin, dtype=np.float64).reshape(3), tuple(int(v) for v in shape)
    out = np.empty(shape, dtype=np.float64)
    tri_c, plane_c = np.ascontiguousarray(tri.reshape(-1, 9)), np.ascontiguousarray(plane)
    org = np.ascontiguousarray(origin, dtype=np.float64)
    lib = _cabi.load()
    dev = torch.cuda.current_device() if device is None else int(device)
    rc = lib.vgpmp_mesh_to_sdf(dev, tri_c.ctypes.data_as(_cabi.c_double_p), plane_c.ctypes.data_as(_cabi.c_double_p),
                               piece_end.ctypes.data_as(_cabi.c_int32_p), len(tri), len(piece_end), *shape,
                               org.ctypes.data_as(_cabi.c_double_p), float(delta), out.ctypes.data_as(_cabi.c_double_p))
    if rc != 0:
        raise _cabi.VgpmpError(f"vgpmp_mesh_to_sdf failed (status {rc})")
    return SignedDistanceField(out, origin, delta)
