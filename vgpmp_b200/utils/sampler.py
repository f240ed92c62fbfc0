"""`Sampler`: forward kinematics of the collision spheres (reference: gpflow_vgpmp/utils/sampler.py).

Same constructor and method names as the reference; the arithmetic runs in the CUDA kernels of
csrc/kinematics.cu through the C-ABI.  The per-robot hand-tuned sphere-offset remaps of
Sampler.get_mat (sampler.py:68-101) are kept as a data table.
"""
from __future__ import annotations

import numpy as np

from ..engine import Engine, RobotConstants
from .robot import Robot, get_base

__all__ = ["Sampler", "set_base"]


def set_base(translation) -> np.ndarray:
    assert len(translation) == 3
    return get_base(np.eye(3), translation)


# (first index, last index inclusive, source axis per output axis, sign per output axis, additive shift)
_REMAP = {
    "wam": [(0, 7, (0, 1, 2), (1, -1, 1), (-0.045, 0.0, 0.0)),
            (8, 8, (0, 1, 2), (0, 0, 0), (0.0, 0.0, 0.0)),
            (9, 12, (0, 1, 2), (1, -1, 1), (0.045, -0.05, 0.0)),
            (13, 14, (0, 1, 2), (1, -1, 1), (0.0, 0.0, 0.0)),
            (15, 10 ** 6, (0, 1, 2), (1, 1, 1), (0.0, 0.0, 0.0))],
    "ur10": [(0, 0, (2, 0, 1), (1, 1, 1), (0.0, 0.0, 0.0)),
             (1, 6, (2, 0, 1), (1, 1, 1), (0.0, 0.0, 0.163941 + 0.05)),
             (7, 10 ** 6, (2, 0, 1), (1, 1, 1), (0.0, 0.0, 0.0))],
    "kuka": [(0, 1, (0, 1, 2), (1, 1, 1), (0.0, 0.0, 0.0)),
             (2, 4, (0, 2, 1), (1, -1, 1), (0.0, 0.18, 0.0)),
             (5, 7, (0, 2, 1), (1, 1, 1), (0.0, 0.0, 0.0)),
             (8, 10, (0, 2, 1), (1, 1, -1), (0.0, -0.18, 0.0)),
             (11, 14, (0, 2, 1), (1, -1, 1), (0.0, 0.0, 0.0)),
             (15, 16, (0, 2, 1), (1, 1, 1), (0.0, 0.1, -0.06)),
             (17, 19, (0, 2, 1), (1, 1, 1), (0.0, -0.07, 0.0)),
             (20, 10 ** 6, (0, 1, 2), (1, 1, 1), (0.0, 0.0, 0.0))],
}


class Sampler:
    def __init__(self, config=None, robot: Robot = None):
        if robot is None:
            raise AssertionError("Sampler needs a robot")
        assert robot.sphere_offsets is not None and len(robot.sphere_offsets) > 0
        assert robot.base_pose is not None
        self.robot = robot
        self.name, self.dof = robot.name, robot.dof
        self.DH, self.twist = robot.DH.copy(), robot.twist.copy()
        self.d, self.a, self.alpha = (self.DH[:, k].reshape(self.dof, 1) for k in range(3))
        self.fk_slice = list(robot.fk_slice)
        self.craig_notation = robot.craig_notation
        self.base_pose = robot.base_pose.copy()
        self.num_spheres_per_link = list(robot.num_spheres_per_link)
        self.sphere_offsets = np.stack([self.get_mat(self.name, i, o) for i, o in enumerate(robot.sphere_offsets)])
        self._engine = None

    def get_mat(self, robot_name, index, offset) -> np.ndarray:
        """4x4 translation of sphere `index` in its link frame (sampler.py:68-101)."""
        off = np.asarray(offset, dtype=np.float64)
        for lo, hi, src, sign, shift in _REMAP.get(robot_name, ()):
            if lo <= index <= hi:
                off = np.array([sign[k] * off[src[k]] + shift[k] for k in range(3)])
                break
        return set_base(off)

    # ---- constants handed to the C-ABI -------------------------------------------------------------
    def constants(self) -> RobotConstants:
        r = self.robot
        frame = np.repeat(np.asarray(self.fk_slice, dtype=np.int32), self.num_spheres_per_link).astype(np.int32)
        return RobotConstants(dof=self.dof, craig=self.craig_notation, dh=self.DH, twist=self.twist.reshape(-1),
                              base_pose=self.base_pose, sphere_frame=frame,
                              sphere_offsets=self.sphere_offsets[:, :3, 3].copy(),
                              sphere_radii=np.asarray(r.sphere_radii, dtype=np.float64),
                              limits_lo=r.limits_lo, limits_hi=r.limits_hi)

    def _eng(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self.constants(), np.zeros((1, 1, 1)), (0.0, 0.0, 0.0), 1.0)
        return self._engine

    # ---- reference API ---------------------------------------------------------------------------
    def forward_kinematics(self, thetas):
        """thetas [D,1] (or [D]) -> [D+1,4,4]; a batch [n,D] gives [n,D+1,4,4] (sampler.py:103-120)."""
        t = self._eng().dev(thetas)
        single = t.numel() == self.dof
        out = self._eng().fk_frames(t.reshape(-1, self.dof))
        return out[0] if single else out

    def forward_kinematics_cost(self, joint_config):
        """joint_config [D,1] -> sphere centres [P,3]; a batch [n,D] gives [n,P,3] (sampler.py:216-235)."""
        t = self._eng().dev(joint_config)
        single = t.numel() == self.dof
        out = self._eng().fk_spheres(t.reshape(-1, self.dof))
        return out[0] if single else out

    def _forward_kinematics_joints_to_spheres(self, joint_config):
        """Frames of every sphere [P,4,4] (sampler.py:237-244)."""
        frames = self.forward_kinematics(joint_config)
        idx = np.repeat(np.asarray(self.fk_slice), self.num_spheres_per_link)
        return frames[..., idx, :, :]
