"""Thin host wrapper around one `vgpmp_handle` (include/vgpmp_b200.h).

PyTorch is used only to own device buffers and streams; every computation is a call into
libvgpmp_b200.so.  A missing library or a missing GPU is a hard error (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _cabi

F64 = torch.float64


@dataclass
class RobotConstants:
    """What Sampler.__init__ / Robot.initialise leave behind (utils/sampler.py:28-56, utils/robot.py:482-550)."""
    dof: int
    craig: bool
    dh: np.ndarray              # [D,3]
    twist: np.ndarray           # [D]
    base_pose: np.ndarray       # [4,4]
    sphere_frame: np.ndarray    # [P] int32
    sphere_offsets: np.ndarray  # [P,3]
    sphere_radii: np.ndarray    # [P]
    limits_lo: np.ndarray       # [D]
    limits_hi: np.ndarray       # [D]

    @classmethod
    def dummy(cls, dof: int) -> "RobotConstants":
        return cls(dof=dof, craig=False, dh=np.zeros((dof, 3)), twist=np.zeros(dof), base_pose=np.eye(4),
                   sphere_frame=np.zeros(1, np.int32), sphere_offsets=np.zeros((1, 3)), sphere_radii=np.zeros(1),
                   limits_lo=-np.ones(dof), limits_hi=np.ones(dof))


def _c64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_cabi.c_double_p)


class Engine:
    """One handle = robot constants + SDF grid (resident in HBM) + likelihood constants on one GPU."""

    def __init__(self, robot: RobotConstants, sdf_data: np.ndarray, sdf_origin: Sequence[float], sdf_delta: float,
                 sigma_obs: float = 1.0, epsilon: float = 0.0, alpha: float = 1.0,
                 scene_offset: Sequence[float] = (0.0, 0.0, 0.0), jitter: float = 1e-6, device: Optional[int] = None,
                 share_from: Optional["Engine"] = None):
        """`share_from`: another Engine on the same device whose SDF records (and robot constants) this handle re-uses
        (`vgpmp_create_shared`): one copy of the grid per device however many handles / streams drive it."""
        self.lib = _cabi.load()
        if not torch.cuda.is_available():
            raise _cabi.VgpmpError("vgpmp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.robot = robot
        self.D, self.P = int(robot.dof), int(len(robot.sphere_radii))
        self.sigma_obs, self.epsilon, self.alpha, self.jitter = float(sigma_obs), float(epsilon), float(alpha), float(jitter)
        self.scene_offset = tuple(float(v) for v in scene_offset)
        self._ws = None
        keep = dict(dh=_c64(robot.dh), twist=_c64(robot.twist), base=_c64(robot.base_pose),
                    frame=np.ascontiguousarray(robot.sphere_frame, dtype=np.int32), off=_c64(robot.sphere_offsets),
                    rad=_c64(robot.sphere_radii), lo=_c64(robot.limits_lo), hi=_c64(robot.limits_hi))
        rd = _cabi.RobotDesc(self.D, int(bool(robot.craig)), self.P, _dp(keep["dh"]), _dp(keep["twist"]),
                             _dp(keep["base"]), keep["frame"].ctypes.data_as(_cabi.c_int32_p), _dp(keep["off"]),
                             _dp(keep["rad"]), _dp(keep["lo"]), _dp(keep["hi"]))
        ld = _cabi.LikDesc(self.sigma_obs, self.epsilon, self.alpha, (C.c_double * 3)(*self.scene_offset), self.jitter)
        h = C.c_void_p()
        if share_from is not None:
            if share_from.device_index != self.device_index:
                raise _cabi.VgpmpError("share_from must live on the same device")
            rc = self.lib.vgpmp_create_shared(C.byref(h), share_from.h, C.byref(rd), C.byref(ld))
            if rc != 0:
                raise _cabi.VgpmpError(f"vgpmp_create_shared failed ({rc}): {self.lib.vgpmp_last_error(None).decode()}")
            self.h = h
            return
        keep["grid"] = _c64(sdf_data)
        nx, ny, nz = keep["grid"].shape
        sd = _cabi.SdfDesc(nx, ny, nz, _dp(keep["grid"]), (C.c_double * 3)(*map(float, sdf_origin)), float(sdf_delta))
        rc = self.lib.vgpmp_create(C.byref(h), self.device_index, C.byref(rd), C.byref(sd), C.byref(ld))
        if rc != 0:
            raise _cabi.VgpmpError(f"vgpmp_create failed ({rc}): {self.lib.vgpmp_last_error(None).decode()}")
        self.h = h

    def shared(self, alpha: Optional[float] = None) -> "Engine":
        """A second handle on the same SDF records (e.g. for a sub-batch driven on its own stream)."""
        return Engine(self.robot, None, (0, 0, 0), 1.0, sigma_obs=self.sigma_obs, epsilon=self.epsilon,
                      alpha=self.alpha if alpha is None else alpha, scene_offset=self.scene_offset, jitter=self.jitter,
                      device=self.device_index, share_from=self)

    @property
    def sdf_records_id(self) -> int:
        return int(self.lib.vgpmp_sdf_records_id(self.h))

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                self.lib.vgpmp_destroy(h)
            except Exception:
                pass

    # ---- helpers -------------------------------------------------------------------------------
    _gp_cache = {}

    @classmethod
    def for_gp(cls, dof: int) -> "Engine":
        """Handle for the GP-only entry points (Kuu, Kuf, prior_kl): dof = number of latent GPs, no robot, no SDF."""
        key = (dof, torch.cuda.current_device() if torch.cuda.is_available() else -1)
        if key not in cls._gp_cache:
            cls._gp_cache[key] = cls(RobotConstants.dummy(dof), np.zeros((1, 1, 1)), (0, 0, 0), 1.0)
        return cls._gp_cache[key]

    def dev(self, x, shape=None) -> torch.Tensor:
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64))
        t = t.to(device=self.device, dtype=F64).contiguous()
        if shape is not None:
            t = t.reshape(shape)
        return t

    def empty(self, *shape) -> torch.Tensor:
        return torch.empty(*shape, dtype=F64, device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, rc, what):
        _cabi.check(self.h, rc, what)

    def dims(self, Bp, M, N, S, B, total_samples=0, kl_shards=0) -> _cabi.Dims:
        return _cabi.Dims(int(Bp), int(M), int(N), int(S), int(B), int(total_samples), int(kl_shards))

    def workspace(self, dims: _cabi.Dims) -> torch.Tensor:
        need = int(self.lib.vgpmp_workspace_bytes(self.h, C.byref(dims)))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def set_option(self, name: str, value: int):
        self._chk(self.lib.vgpmp_set_option(self.h, name.encode(), int(value)), "set_option")

    @property
    def launch_count(self) -> int:
        return int(self.lib.vgpmp_launch_count(self.h))

    # ---- stage kernels -------------------------------------------------------------------------
    def fk_frames(self, joints) -> torch.Tensor:
        j = self.dev(joints).reshape(-1, self.D)
        out = self.empty(j.shape[0], self.D + 1, 4, 4)
        self._chk(self.lib.vgpmp_fk_frames(self.h, j.data_ptr(), out.data_ptr(), j.shape[0], self._stream()), "fk_frames")
        return out

    def fk_spheres(self, joints) -> torch.Tensor:
        j = self.dev(joints).reshape(-1, self.D)
        out = self.empty(j.shape[0], self.P, 3)
        self._chk(self.lib.vgpmp_fk_spheres(self.h, j.data_ptr(), out.data_ptr(), j.shape[0], self._stream()), "fk_spheres")
        return out

    def sdf_lookup(self, pts, with_grad=True):
        p = self.dev(pts).reshape(-1, 3)
        dist = self.empty(p.shape[0])
        grad = self.empty(p.shape[0], 3) if with_grad else None
        self._chk(self.lib.vgpmp_sdf_lookup(self.h, p.data_ptr(), dist.data_ptr(), grad.data_ptr() if with_grad else None,
                                            p.shape[0], self._stream()), "sdf_lookup")
        return dist, grad

    def loglik(self, x, squash: bool, upstream: float = 1.0, need_grad: bool = True):
        """x [...,D] -> logp [...], d_x [...,D] (= upstream * dlogp/dx) or None."""
        xin = self.dev(x)
        flat = xin.reshape(-1, self.D)
        logp = self.empty(flat.shape[0])
        dx = torch.empty_like(flat) if need_grad else None
        self._chk(self.lib.vgpmp_loglik_fwd_bwd(self.h, flat.data_ptr(), int(bool(squash)), float(upstream), logp.data_ptr(),
                                                dx.data_ptr() if need_grad else None, flat.shape[0], self._stream()),
                  "loglik_fwd_bwd")
        return logp.reshape(xin.shape[:-1]), (dx.reshape(xin.shape) if need_grad else None)

    def clearance(self, joints) -> torch.Tensor:
        j = self.dev(joints)
        flat = j.reshape(-1, self.D)
        out = self.empty(flat.shape[0])
        self._chk(self.lib.vgpmp_clearance(self.h, flat.data_ptr(), out.data_ptr(), flat.shape[0], self._stream()), "clearance")
        return out.reshape(j.shape[:-1])

    def predict_f_mean(self, dims, params: _cabi.Params, Xq) -> torch.Tensor:
        Xq = self.dev(Xq).reshape(-1, self.D)
        mean = self.empty(dims.num_problems, Xq.shape[0], self.D)
        ws = self.workspace(dims)
        self._chk(self.lib.vgpmp_predict_f_mean(self.h, C.byref(dims), C.byref(params), Xq.data_ptr(), Xq.shape[0],
                                                mean.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()),
                  "predict_f_mean")
        return mean

    def kuu(self, Z, lengthscales, variances, jitter=0.0) -> torch.Tensor:
        Zd = self.dev(Z).reshape(-1, self.D)
        ls, var = self.dev(lengthscales).reshape(-1, self.D), self.dev(variances).reshape(-1, self.D)
        Bp, M = ls.shape[0], Zd.shape[0]
        K = self.empty(Bp, self.D, M + 2, M + 2)
        self._chk(self.lib.vgpmp_kuu(self.h, Zd.data_ptr(), ls.data_ptr(), var.data_ptr(), float(jitter), K.data_ptr(),
                                     Bp, M, self._stream()), "kuu")
        return K

    def kuf(self, Z, X, lengthscales, variances) -> torch.Tensor:
        Zd, Xd = self.dev(Z).reshape(-1, self.D), self.dev(X).reshape(-1, self.D)
        ls, var = self.dev(lengthscales).reshape(-1, self.D), self.dev(variances).reshape(-1, self.D)
        Bp, M, N = ls.shape[0], Zd.shape[0], Xd.shape[0]
        K = self.empty(Bp, self.D, M + 2, N)
        self._chk(self.lib.vgpmp_kuf(self.h, Zd.data_ptr(), Xd.data_ptr(), ls.data_ptr(), var.data_ptr(), K.data_ptr(),
                                     Bp, M, N, self._stream()), "kuf")
        return K

    # ---- GP / fused iteration ------------------------------------------------------------------
    @staticmethod
    def params_struct(q_mu, q_sqrt, ls, var, query_latent, Z, X) -> _cabi.Params:
        ps = _cabi.Params(q_mu.data_ptr(), q_sqrt.data_ptr(), ls.data_ptr(), var.data_ptr(), query_latent.data_ptr(),
                          Z.data_ptr(), X.data_ptr() if X is not None else None)
        # the struct only holds raw addresses: keep the tensors alive as long as it lives, otherwise a temporary (e.g.
        # `dev(X)`) is returned to the caching allocator and re-used by the very call that still reads it
        ps._keep = (q_mu, q_sqrt, ls, var, query_latent, Z, X)
        return ps

    @staticmethod
    def _ptr(t):
        return None if t is None else t.data_ptr()

    @staticmethod
    def draws_struct(d: dict) -> _cabi.Draws:
        ds = _cabi.Draws(*(Engine._ptr(d[k]) for k in ("omega", "tau", "w", "eps_u", "eps_j")))
        ds._keep = tuple(d[k] for k in ("omega", "tau", "w", "eps_u", "eps_j"))
        return ds

    def alloc_draws(self, dims: _cabi.Dims, lazy_only: bool = False) -> dict:
        """`lazy_only`: omega / tau / w stay None (a set for `rng_fill_lazy` that is never materialised: at BASELINE
        config 5, w alone would be 8192 x 7 x 256 x 1024 doubles = 112 GiB)."""
        Bp, D, B, S, Mp = dims.num_problems, self.D, dims.num_bases, dims.num_samples, dims.num_inducing + 2
        if lazy_only:
            return dict(omega=None, tau=None, w=None, eps_u=self.empty(Bp, D, S, Mp), eps_j=self.empty(Bp, D, S, Mp))
        return dict(omega=self.empty(Bp, D, B, D), tau=self.empty(Bp, D, B), w=self.empty(Bp, D, S, B),
                    eps_u=self.empty(Bp, D, S, Mp), eps_j=self.empty(Bp, D, S, Mp))

    def rng_fill(self, dims: _cabi.Dims, seed: int, iteration: int, draws: Optional[dict] = None, problem_offset: int = 0,
                 sample_offset: int = 0) -> dict:
        d = self.alloc_draws(dims) if draws is None else draws
        self._chk(self.lib.vgpmp_rng_fill(self.h, C.byref(dims), int(seed), int(iteration), int(problem_offset),
                                          int(sample_offset), d["omega"].data_ptr(), d["tau"].data_ptr(),
                                          d["w"].data_ptr(), d["eps_u"].data_ptr(), d["eps_j"].data_ptr(),
                                          self._stream()), "rng_fill")
        return d

    def rng_fill_lazy(self, dims: _cabi.Dims, seed: int, iteration: int, draws: dict, problem_offset: int = 0,
                      sample_offset: int = 0) -> dict:
        """`vgpmp_rng_fill_lazy`: eps_u / eps_j are written, omega / tau / w are produced by whoever consumes `draws` next
        (inside the sampler kernel for equispaced inputs).  Their tensors hold unspecified values afterwards."""
        d = draws
        self._chk(self.lib.vgpmp_rng_fill_lazy(self.h, C.byref(dims), int(seed), int(iteration), int(problem_offset),
                                               int(sample_offset), self._ptr(d["omega"]), self._ptr(d["tau"]),
                                               self._ptr(d["w"]), d["eps_u"].data_ptr(), d["eps_j"].data_ptr(),
                                               self._stream()), "rng_fill_lazy")
        return d

    def rng_fill_async(self, dims: _cabi.Dims, seed: int, iteration: int, draws: dict, slot: int, problem_offset: int = 0,
                       sample_offset: int = 0):
        """Generate `draws` on the handle's side stream (overlaps with work queued on the current stream)."""
        self._chk(self.lib.vgpmp_rng_fill_async(self.h, C.byref(dims), int(seed), int(iteration), int(problem_offset),
                                                int(sample_offset), draws["omega"].data_ptr(), draws["tau"].data_ptr(),
                                                draws["w"].data_ptr(), draws["eps_u"].data_ptr(), draws["eps_j"].data_ptr(),
                                                int(slot)), "rng_fill_async")

    def rng_join(self, slot: int):
        self._chk(self.lib.vgpmp_rng_join(self.h, int(slot), self._stream()), "rng_join")

    def rng_release(self, slot: int):
        self._chk(self.lib.vgpmp_rng_release(self.h, int(slot), self._stream()), "rng_release")

    def gp_prepare(self, dims, params: _cabi.Params):
        Bp, Mp = dims.num_problems, dims.num_inducing + 2
        Lc, Sf, kl = self.empty(Bp, self.D, Mp, Mp), self.empty(Bp, self.D, Mp, Mp), self.empty(Bp)
        ws = self.workspace(dims)
        self._chk(self.lib.vgpmp_gp_prepare(self.h, C.byref(dims), C.byref(params), Lc.data_ptr(), Sf.data_ptr(),
                                            kl.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()), "gp_prepare")
        return Lc, Sf, kl

    def pathwise_sample(self, dims, params: _cabi.Params, draws: dict, Xq) -> torch.Tensor:
        Xq = self.dev(Xq).reshape(-1, self.D)
        Nq = Xq.shape[0]
        dq = self.dims(dims.num_problems, dims.num_inducing, Nq, dims.num_samples, dims.num_bases, dims.total_samples,
                       dims.kl_shards)
        f = self.empty(dims.num_problems, dims.num_samples, Nq, self.D)
        ws = self.workspace(dq)
        ds = self.draws_struct(draws)
        self._chk(self.lib.vgpmp_pathwise_sample(self.h, C.byref(dims), C.byref(params), C.byref(ds), Xq.data_ptr(), Nq,
                                                 f.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()),
                  "pathwise_sample")
        return f

    def elbo_fwd_bwd(self, dims, params: _cabi.Params, draws: dict, need_grad=True, want_aux=False, out=None):
        """`out` may hold preallocated 'elbo' / 'd_*' tensors (e.g. views of one packed all-reduce buffer)."""
        Bp, M, N, S = dims.num_problems, dims.num_inducing, dims.num_timesteps, dims.num_samples
        pre = out or {}
        elbo = pre.get("elbo", None)
        if elbo is None:
            elbo = self.empty(Bp)
        out = dict(elbo=elbo)
        gs = None
        if need_grad:
            for key, shp in (("d_q_mu", (Bp, M, self.D)), ("d_q_sqrt", (Bp, self.D, M, M)),
                             ("d_lengthscales", (Bp, self.D)), ("d_variances", (Bp, self.D))):
                out[key] = pre[key] if key in pre else self.empty(*shp)
            gs = _cabi.Grads(out["d_q_mu"].data_ptr(), out["d_q_sqrt"].data_ptr(), out["d_lengthscales"].data_ptr(),
                             out["d_variances"].data_ptr())
        aux = None
        if want_aux:
            out.update(f=self.empty(Bp, S, N, self.D), logp=self.empty(Bp, S, N), kl=self.empty(Bp))
            aux = _cabi.Aux(out["f"].data_ptr(), out["logp"].data_ptr(), out["kl"].data_ptr())
        ws = self.workspace(dims)
        ds = self.draws_struct(draws)
        self._chk(self.lib.vgpmp_elbo_fwd_bwd(self.h, C.byref(dims), C.byref(params), C.byref(ds), elbo.data_ptr(),
                                              C.byref(gs) if gs is not None else None,
                                              C.byref(aux) if aux is not None else None, ws.data_ptr(), ws.numel(),
                                              self._stream()), "elbo_fwd_bwd")
        return out
