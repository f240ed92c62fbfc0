"""`SignedDistanceField` (reference: gpflow_vgpmp/utils/sdf_utils.py:25-44,62-136,195-210).

Container + text loader on the host; nearest-voxel lookup and the clipped 7-point central-difference
gradient (exact zeros -> 0.1) run in csrc/kinematics.cu.  mayavi visualisation / CLI are out of scope.
"""
from __future__ import annotations

import pickle

import numpy as np

from ..engine import Engine, RobotConstants

__all__ = ["SignedDistanceField", "synthetic_shelf_sdf"]


class SignedDistanceField:
    """data[x, y, z], float64, C-order (z fastest)."""

    def __init__(self, data: np.ndarray, origin: np.ndarray, delta: float):
        self.data = np.ascontiguousarray(np.asarray(data, dtype=np.float64))
        if self.data.ndim != 3:
            raise ValueError("SDF data must be a 3-D array")
        self.nx, self.ny, self.nz = self.data.shape
        self.origin = np.asarray(origin, dtype=np.float64).reshape(3)
        self.delta = float(delta)
        self.min_coords = self.origin
        self.max_coords = self.origin + self.delta * np.array(self.data.shape)
        self._engine = None

    def _eng(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(RobotConstants.dummy(1), self.data, self.origin, self.delta)
        return self._engine

    def get_distance_tf(self, rel_pos):
        """rel_pos [...,3] (already relative to the grid) -> grid value at clip(trunc((x-origin)/delta)) (sdf_utils.py:73-76)."""
        p = self._eng().dev(rel_pos)
        dist, _ = self._eng().sdf_lookup(p.reshape(-1, 3), with_grad=False)
        return dist.reshape(p.shape[:-1])

    def get_distance_grad_tf(self, rel_pos):
        """6 clipped-neighbour central differences /(2 delta); exact zeros become 0.1 (sdf_utils.py:100-136)."""
        p = self._eng().dev(rel_pos)
        _, grad = self._eng().sdf_lookup(p.reshape(-1, 3), with_grad=True)
        return grad.reshape(p.shape)

    get_distance = get_distance_tf
    get_distance_grad = get_distance_grad_tf

    @classmethod
    def from_sdf(cls, sdf_file):
        """'nx ny nz' / 'x0 y0 z0' / 'delta' / one value per line, x fastest (sdf_utils.py:195-210)."""
        with open(sdf_file, "r") as fh:
            nx, ny, nz = (int(v) for v in fh.readline().split())
            origin = np.array([float(v) for v in fh.readline().split()])
            delta = float(fh.readline().strip())
            vals = np.loadtxt(fh, dtype=np.float64, ndmin=1)
        if vals.size < nx * ny * nz:
            raise ValueError(f"{sdf_file}: expected {nx * ny * nz} values, found {vals.size}")
        data = vals[:nx * ny * nz].reshape(nz, ny, nx).transpose(2, 1, 0)
        return cls(data, origin, delta)

    def dump(self, pkl_file):
        with open(pkl_file, "wb") as fh:
            pickle.dump({"data": self.data, "origin": self.origin, "delta": self.delta}, fh, protocol=2)

    @classmethod
    def from_pkl(cls, pkl_file):
        with open(pkl_file, "rb") as fh:
            d = pickle.load(fh)
        return cls(d["data"], d["origin"], d["delta"])

    def to_sdf(self, path):
        with open(path, "w") as fh:
            fh.write(f"{self.nx} {self.ny} {self.nz}\n{float(self.origin[0])!r} {float(self.origin[1])!r} {float(self.origin[2])!r}\n{self.delta!r}\n")
            np.savetxt(fh, self.data.transpose(2, 1, 0).reshape(-1), fmt="%.17g")


def _box_sdf(pts, lo, hi):
    c, hsz = 0.5 * (lo + hi), 0.5 * (hi - lo)
    q = np.abs(pts - c) - hsz
    return np.linalg.norm(np.maximum(q, 0.0), axis=-1) + np.minimum(q.max(axis=-1), 0.0)


def synthetic_shelf_sdf(shape=(96, 96, 96), delta=0.02, origin=(-0.96, -0.96, -0.96), seed=0, n_boxes=12,
                        dtype=np.float64) -> "SignedDistanceField":
    """Analytic stand-in for the reference's missing .sdf grids (.MISSING_LARGE_BLOBS): signed distance to a union
    of axis-aligned boxes -- two uprights, a back panel and shelves in front of the robot, plus seeded clutter.
    The grid frame is the *scene* frame: the likelihood subtracts the scene offset before the lookup."""
    rng = np.random.default_rng(seed)
    origin = np.asarray(origin, dtype=np.float64)
    ext = delta * np.asarray(shape)
    boxes = [(np.array([-0.10, -0.45, -0.80]), np.array([0.20, -0.42, 0.60])),    # left upright
             (np.array([-0.10, 0.42, -0.80]), np.array([0.20, 0.45, 0.60])),      # right upright
             (np.array([0.18, -0.45, -0.80]), np.array([0.20, 0.45, 0.60]))]      # back panel
    for z in (-0.80, -0.35, 0.10, 0.55):
        boxes.append((np.array([-0.10, -0.45, z]), np.array([0.20, 0.45, z + 0.03])))
    for _ in range(max(0, n_boxes - len(boxes))):
        c = origin + ext * rng.uniform(0.15, 0.85, size=3)
        h = rng.uniform(0.03, 0.12, size=3)
        boxes.append((c - h, c + h))
    ax = [origin[k] + delta * np.arange(shape[k]) for k in range(3)]
    out = np.full(shape, np.inf)
    for i0 in range(0, shape[0], 16):          # slabs keep the temporary small
        X, Y, Zc = np.meshgrid(ax[0][i0:i0 + 16], ax[1], ax[2], indexing="ij")
        pts = np.stack([X, Y, Zc], axis=-1)
        slab = np.full(pts.shape[:-1], np.inf)
        for lo, hi in boxes:
            slab = np.minimum(slab, _box_sdf(pts, lo, hi))
        out[i0:i0 + 16] = slab
    return SignedDistanceField(out.astype(dtype).astype(np.float64), origin, delta)
