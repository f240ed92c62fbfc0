"""`VariationalMonteCarloLikelihood` (reference: gpflow_vgpmp/likelihoods/likelihood.py:18-176).

Joint configurations -> sphere centres -> signed distance - radius -> hinge -> Gaussian-like log-probability, with
the SDF custom gradient.  One fused CUDA kernel (csrc/kinematics.cu: loglik_kernel) does the forward pass and the
reverse pass to the input.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from ..engine import Engine
from ..utils.robot import Robot
from ..utils.sampler import Sampler
from ..utils.sdf_utils import SignedDistanceField

__all__ = ["VariationalMonteCarloLikelihood", "JointSigmoid"]


class JointSigmoid:
    """tfb.Sigmoid(low=joint_limits[:,1], high=joint_limits[:,0]) (likelihood.py:49-52)."""

    def __init__(self, low, high):
        self.low, self.high = np.asarray(low, dtype=np.float64), np.asarray(high, dtype=np.float64)

    def _lh(self, like):
        if isinstance(like, torch.Tensor):
            return (torch.as_tensor(self.low, device=like.device), torch.as_tensor(self.high, device=like.device))
        return self.low, self.high

    def forward(self, x):
        lo, hi = self._lh(x)
        if isinstance(x, torch.Tensor):
            return lo + (hi - lo) * torch.sigmoid(x)
        x = np.asarray(x, dtype=np.float64)
        return lo + (hi - lo) / (1.0 + np.exp(-x))

    __call__ = forward

    def inverse(self, y):
        lo, hi = self._lh(y)
        if isinstance(y, torch.Tensor):
            u = (y - lo) / (hi - lo)
            return torch.log(u) - torch.log1p(-u)
        u = (np.asarray(y, dtype=np.float64) - lo) / (hi - lo)
        return np.log(u) - np.log1p(-u)


class VariationalMonteCarloLikelihood:
    def __init__(self, sigma_obs: float, robot: Robot, sampler: Sampler, sdf: SignedDistanceField, offset: List[float],
                 epsilon: float = 0.05, DEFAULT_VARIANCE_LOWER_BOUND=1e-5, **kwargs):
        if not float(sigma_obs) > DEFAULT_VARIANCE_LOWER_BOUND:
            raise ValueError("sigma_obs must exceed the positive() lower bound of likelihood.variance")
        self.sdf, self.sampler, self.robot = sdf, sampler, robot
        self.sigma_obs = float(sigma_obs)
        self.variance = np.full((1, robot.num_spheres), self.sigma_obs)     # holds sigma_obs itself (likelihood.py:37-41)
        self.offset = np.asarray(offset, dtype=np.float64).reshape(1, 3)
        self.sphere_radii = np.asarray(robot.sphere_radii, dtype=np.float64).reshape(1, -1)
        self.joint_constraints = np.asarray(robot.joint_limits, dtype=np.float64).reshape(-1, 2)
        self.velocity_constraints = np.asarray(robot.velocity_limits, dtype=np.float64).reshape(-1, 2)
        self.joint_sigmoid = JointSigmoid(low=self.joint_constraints[:, 1], high=self.joint_constraints[:, 0])
        self.epsilon = float(epsilon)
        self.p = robot.num_spheres
        self._engines = {}          # alpha -> Engine; all of them share ONE copy of the SDF records on the device
        self._share_from = kwargs.get("share_engine")   # an Engine of another likelihood over the same robot + SDF

    def _engine(self, alpha: float = 1.0) -> Engine:
        """The handle that carries this likelihood's constants and `alpha` (VGPMP.alpha is baked into the handle).  A
        likelihood used with several alphas, or several likelihoods over one environment (`share_engine`), get handles
        made with `vgpmp_create_shared`: the SDF is uploaded once."""
        alpha = float(alpha)
        if alpha not in self._engines:
            # the SignedDistanceField object owns the device copy of its records (sdf._eng()); every handle made here shares it
            base = next(iter(self._engines.values()), None) or self._share_from or self.sdf._eng()
            self._engines[alpha] = Engine(self.sampler.constants(), None, (0, 0, 0), 1.0, share_from=base,
                                          sigma_obs=self.sigma_obs, epsilon=self.epsilon, alpha=alpha,
                                          scene_offset=self.offset.reshape(3))
        return self._engines[alpha]

    # ---- reference API ---------------------------------------------------------------------------
    def log_prob(self, F):
        """F [S,N,D] joint configurations -> log p(e|f) [S,N] (likelihood.py:57-99)."""
        logp, _ = self._engine().loglik(F, squash=False, need_grad=False)
        return logp

    _log_prob = log_prob

    def log_prob_and_grad(self, F, squash=False, upstream=1.0):
        return self._engine().loglik(F, squash=squash, upstream=upstream, need_grad=True)

    def _sample_config_cost(self, f):
        """[S,N,D] -> sphere centres [S,N,P,3] (likelihood.py:101-125)."""
        eng = self._engine()
        t = eng.dev(f)
        return eng.fk_spheres(t.reshape(-1, eng.D)).reshape(*t.shape[:-1], self.p, 3)

    def _compute_forward_kinematics_cost(self, joint_config):
        return self.sampler.forward_kinematics_cost(joint_config)

    def _signed_distance_grad(self, data):
        """[...,P,3] world positions -> (dist [...,P], dist_grad [...,P,3]) (likelihood.py:146-176)."""
        eng = self._engine()
        d = eng.dev(data)
        rel = d - eng.dev(self.offset.reshape(3))
        dist, grad = eng.sdf_lookup(rel.reshape(-1, 3), with_grad=True)
        dist = dist.reshape(d.shape[:-1]) - eng.dev(self.sphere_radii.reshape(-1))
        return dist, grad.reshape(d.shape)

    def _hinge_loss(self, data):
        dist, _ = self._signed_distance_grad(data)
        return torch.clamp(self.epsilon - dist, min=0.0)

    def _scalar_log_prob(self, f):
        cost = self._hinge_loss(f)
        var = self._engine().dev(self.variance.reshape(-1))
        return -0.5 * torch.sum(cost / var * cost, dim=-1)
