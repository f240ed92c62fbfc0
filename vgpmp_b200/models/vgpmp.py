"""`VGPMP` planner model (reference: gpflow_vgpmp/models/vgpmp.py:59-339).

Same constructor / `initialize` / property / method names as the reference.  All state lives in HBM as float64
buffers; `elbo`, its gradients and the Adam update are calls into libvgpmp_b200.so.  A model holds a *batch* of
Bp independent planning problems (query_states [Bp,2,D]); Bp = 1 reproduces the reference object.
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import List, Optional

import numpy as np
import torch

from .. import _cabi
from ..inducing_variables import ConditionedVariableInducingPoints, SharedIndependentInducingVariables
from ..kernels import Matern52, SeparateIndependent, SharedIndependent, VanillaConditioningSeparateIndependent
from ..likelihoods import VariationalMonteCarloLikelihood

__all__ = ["VGPMP", "initialize_Z", "AdamConfig"]

DEFAULT_JITTER = 1e-6


def initialize_Z(num_latent_gps, num_inducing):
    """linspace(.1,.9,M) replicated over the latent GPs (models/vgpmp.py:37-42; sigmoid-bounded to (.09,.91) there)."""
    return np.array([np.full(num_latent_gps, t) for t in np.linspace(0.1, 0.9, num_inducing)], dtype=np.float64)


def _softplus_inv(y):
    y = np.asarray(y, dtype=np.float64)
    return y + np.log(-np.expm1(-y))


class AdamConfig:
    """tf.optimizers.Adam(learning_rate, beta_1=0.8, beta_2=0.95) (models/vgpmp.py:77); epsilon = Keras default 1e-7."""

    def __init__(self, learning_rate, beta_1=0.8, beta_2=0.95, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = float(learning_rate), beta_1, beta_2, epsilon


class VGPMP:
    def __init__(self, kernel, likelihood: VariationalMonteCarloLikelihood, inducing_variable, *, num_latent_gps: int,
                 alpha: float, query_states, num_samples: int, num_bases: int, num_inducing: int, learning_rate: float,
                 q_mu=None, num_data=None, whiten: bool = False, prior=None, variance_lower: float = 0.0, seed: int = 0,
                 **kwargs):
        if whiten:
            raise NotImplementedError("the reference builds the model with whiten=False (models/vgpmp.py:198)")
        if not isinstance(kernel, SeparateIndependent):
            raise AssertionError("Kernels must be a list of kernels or a SharedIndependent kernel.")
        self.kernel, self.likelihood, self.inducing_variable = kernel, likelihood, inducing_variable
        self.num_latent_gps = D = int(num_latent_gps)
        self.num_samples, self.num_bases, self.num_inducing = int(num_samples), int(num_bases), int(num_inducing)
        self.lazy_draws = True      # train_step: draws generated inside the sampler kernel (vgpmp_rng_fill_lazy)
        self.problem_offset = 0     # global index of this model's first problem (Philox keys use global problem indices)
        self.num_data, self.prior, self.seed = num_data, prior, int(seed)
        self.optimizer = AdamConfig(learning_rate)
        self.alpha = float(alpha)
        self.trainable = dict(q_mu=True, q_sqrt=True, lengthscales=True, kernel_variance=True)
        self._eng = eng = likelihood._engine(alpha=self.alpha)   # the handle carries alpha: one (shared-SDF) handle per value
        if eng.D != D:
            raise AssertionError("num_latent_gps must equal the robot's degrees of freedom")

        qs = np.asarray(query_states, dtype=np.float64)
        qs = qs.reshape(1, 2, D) if qs.ndim == 2 else qs.reshape(-1, 2, D)
        self.num_problems = Bp = qs.shape[0]
        self._query_joint = qs
        self._query_states = eng.dev(likelihood.joint_sigmoid.inverse(qs))            # [Bp,2,D] latent (vgpmp.py:75-76)
        from ..covariances import unwrap_inducing
        Zh = unwrap_inducing(inducing_variable)        # also rejects conditioned timesteps other than t = 0, 1 (the kernels hard-code Zy = [0; 1; Z])
        if Zh.shape[0] != self.num_inducing:
            raise AssertionError("inducing variable size and num_inducing disagree")
        self._Z = eng.dev(Zh)                                                         # [M,D]
        self._init_variational_parameters(self.num_inducing, q_mu, None, False)
        ls, var = kernel.hyper_arrays()
        if np.any(var <= variance_lower):
            raise ValueError("kernel variance must exceed the lower bound of its positive() transform "
                             "(SURVEY.md appendix C #14: UR10/industrial variance=0.1 is un-representable)")
        self._variance_lower = float(variance_lower)
        self._lengthscales = eng.dev(np.broadcast_to(ls, (Bp, D)).copy())
        self._variances = eng.dev(np.broadcast_to(var, (Bp, D)).copy())
        self._raw_lengthscales = eng.dev(_softplus_inv(np.broadcast_to(ls, (Bp, D))))
        self._raw_variances = eng.dev(_softplus_inv(np.broadcast_to(var, (Bp, D)) - variance_lower))
        for i, k in enumerate(kernel.kernels):
            k._bind(self, i)
        per = self.num_inducing * D + D * self.num_inducing ** 2 + 2 * D
        self._adam_m = torch.zeros(Bp * per, dtype=torch.float64, device=eng.device)
        self._adam_v = torch.zeros_like(self._adam_m)
        self._step = 0
        self._grads = None
        self._draw_buf = None
        self.last_aux = None
        self._shard = None          # set by enable_sample_sharding (single-problem large-sample mode)
        self._pipe = None           # double-buffered device draws of train_step (next step drawn on a side stream)

    # ---- construction ------------------------------------------------------------------------------
    @classmethod
    def initialize(cls, sdf, robot, sampler, lengthscales: List[float], query_states, sigma_obs: float = 0.05,
                   alpha: float = 1.0, variance: float = 0.1, learning_rate: float = 0.1, num_inducing: int = 14,
                   num_samples: int = 51, num_bases: int = 1024, scene_offset: List[float] = None, num_data=None,
                   num_output_dims=None, kernel=None, num_latent_gps: int = None, epsilon: float = 0.05, q_mu=None,
                   interpolation_method: Optional[str] = 'linear', **kwargs):
        """models/vgpmp.py:84-198.  Extra planner_params keys (num_steps, time_spacing_*) are swallowed like there."""
        qs = np.asarray(query_states, dtype=np.float64)
        if num_output_dims is None:
            num_output_dims = qs.shape[-1]
        if num_latent_gps is None:
            num_latent_gps = num_output_dims
        if num_data is None:
            num_data = 2
        assert lengthscales is not None, "Lengthscales have not been set."
        assert qs.size > 0, "Must pass a motion plan to initialize the model."
        if scene_offset is None:
            scene_offset = [0, 0, 0]
            warnings.warn("Offset has not been set. Defaulting to [0, 0, 0].")
        assert len(lengthscales) == num_latent_gps
        assert num_output_dims == num_latent_gps
        if kernel is not None:
            assert isinstance(kernel, (SeparateIndependent, SharedIndependent)), \
                "Kernels must be a list of kernels or a SharedIndependent kernel."
        else:
            kernel = VanillaConditioningSeparateIndependent(
                [Matern52(lengthscales=lengthscales[i], variance=variance) for i in range(num_latent_gps)])
        conditioned_timesteps = np.stack([np.zeros(num_output_dims), np.ones(num_output_dims)])
        Z = initialize_Z(num_latent_gps, num_inducing)
        _Z = ConditionedVariableInducingPoints(Z=Z, conditioned_timesteps=conditioned_timesteps)
        likelihood = VariationalMonteCarloLikelihood(sigma_obs=sigma_obs, robot=robot, sdf=sdf, sampler=sampler,
                                                     offset=scene_offset, epsilon=epsilon,
                                                     share_engine=kwargs.get("share_engine"))
        batched = qs.ndim == 3
        q3 = qs.reshape(-1, 2, num_output_dims)
        if q_mu is None:
            if interpolation_method is None:
                q_mu = likelihood.joint_sigmoid(np.zeros((q3.shape[0], num_inducing, num_latent_gps)))
            elif interpolation_method == 'linear':
                steps = (np.arange(num_inducing, dtype=np.float64) / num_inducing)[None, :, None]   # i / M, never reaches the goal
                q_mu = q3[:, :1, :] + (q3[:, 1:2, :] - q3[:, :1, :]) * steps
            elif interpolation_method == 'waypoint':
                raise NotImplementedError("'waypoint' builds a [3,D] q_mu in the reference, which its own shape check "
                                          "rejects unless num_inducing == 3 (models/vgpmp.py:172-175)")
            else:
                raise NotImplementedError
        else:
            assert isinstance(q_mu, np.ndarray)
            q_mu = np.broadcast_to(q_mu.reshape(-1, num_inducing, num_latent_gps), (q3.shape[0], num_inducing, num_latent_gps))
        return cls(kernel=kernel, likelihood=likelihood, inducing_variable=SharedIndependentInducingVariables(_Z),
                   num_latent_gps=num_latent_gps, num_samples=num_samples, num_bases=num_bases, num_data=num_data,
                   query_states=qs if batched else q3[0], num_inducing=num_inducing, learning_rate=learning_rate,
                   alpha=alpha, q_mu=q_mu if batched else q_mu[0], whiten=False,
                   **{k: v for k, v in kwargs.items() if k in ("variance_lower", "seed")})

    def _init_variational_parameters(self, num_inducing, q_mu, q_sqrt, q_diag):
        """models/vgpmp.py:255-263: _q_mu = joint_sigmoid.inverse(q_mu); _q_sqrt = I per latent."""
        eng, D, Bp = self._eng, self.num_latent_gps, self.num_problems
        q = np.asarray(q_mu, dtype=np.float64).reshape(-1, num_inducing, D)
        q = np.broadcast_to(q, (Bp, num_inducing, D))
        self._q_mu = eng.dev(self.likelihood.joint_sigmoid.inverse(q))                              # [Bp,M,D]
        self._q_sqrt = eng.dev(np.broadcast_to(np.eye(num_inducing), (Bp, D, num_inducing, num_inducing)).copy())

    # ---- properties -------------------------------------------------------------------------------
    def _squeeze(self, t):
        return t[0] if self.num_problems == 1 else t

    @property
    def query_states(self):
        return self._squeeze(self._query_states)

    @property
    def q_mu(self):
        """concat([query_states, _q_mu]) -> [Mp,D] (models/vgpmp.py:200-202)."""
        return self._squeeze(torch.cat([self._query_states, self._q_mu], dim=1))

    @property
    def q_sqrt(self):
        """Lc @ pad(_q_sqrt) + jitter * diag(1,1,0,..) -> [D,Mp,Mp] (models/vgpmp.py:208-218)."""
        _, Sfull, _ = self._eng.gp_prepare(self._dims(1), self._params(None))
        return self._squeeze(Sfull)

    @property
    def trainable_variables(self):
        out = []
        if self.trainable["q_mu"]:
            out.append(self._q_mu)
        if self.trainable["q_sqrt"]:
            out.append(self._q_sqrt)
        if self.trainable["lengthscales"]:
            out.append(self._raw_lengthscales)
        if self.trainable["kernel_variance"]:
            out.append(self._raw_variances)
        return out

    # ---- plumbing ----------------------------------------------------------------------------------
    def enable_sample_sharding(self, rank: int, world: int, group=None):
        """Large-sample mode (BASELINE config 4): this process keeps samples [lo, hi) of the model's num_samples; ELBO and
        gradients are summed over the ranks with one all-reduce per step (vgpmp_b200/utils/sharding.py)."""
        from ..utils.sharding import packed_layout, shard_range
        total = self.num_samples if self._shard is None else self._shard["total"]
        lo, hi = shard_range(total, rank, world)
        if hi <= lo:
            raise ValueError("more ranks than samples")
        self.num_samples = hi - lo
        n = packed_layout(self.num_problems, self.num_inducing, self.num_latent_gps)["_total"][0]
        self._shard = dict(rank=rank, world=world, total=total, offset=lo, group=group,
                           flat=torch.zeros(n, dtype=torch.float64, device=self._eng.device))
        self._draw_buf = None
        return self

    def _dims(self, N, S=None, Bp=None):
        Bp = self.num_problems if Bp is None else Bp
        if self._shard is not None and S is None:
            return self._eng.dims(Bp, self.num_inducing, N, self.num_samples, self.num_bases,
                                  total_samples=self._shard["total"], kl_shards=self._shard["world"])
        return self._eng.dims(Bp, self.num_inducing, N, self.num_samples if S is None else S, self.num_bases)

    def _params(self, X):
        return self._eng.params_struct(self._q_mu, self._q_sqrt, self._lengthscales, self._variances, self._query_states,
                                       self._Z, X)

    def _make_draws(self, dims, draws):
        """Explicit draws (parity mode: dict of arrays shaped like vgpmp_draws) or device Philox draws (seed, step)."""
        eng = self._eng
        if draws is not None:
            Bp, D, B, S, Mp = dims.num_problems, eng.D, dims.num_bases, dims.num_samples, dims.num_inducing + 2
            shapes = dict(omega=(Bp, D, B, D), tau=(Bp, D, B), w=(Bp, D, S, B), eps_u=(Bp, D, S, Mp), eps_j=(Bp, D, S, Mp))
            return {k: eng.dev(draws[k]).reshape(shapes[k]) for k in shapes}
        key = (dims.num_problems, dims.num_samples, dims.num_bases)
        if self._draw_buf is None or self._draw_buf[0] != key:
            self._draw_buf = (key, eng.alloc_draws(dims))
        off = self._shard["offset"] if (self._shard is not None and dims.total_samples > 0) else 0
        return eng.rng_fill(dims, self.seed, self._step, self._draw_buf[1], problem_offset=self.problem_offset, sample_offset=off)

    # ---- reference API ---------------------------------------------------------------------------
    def elbo(self, data, draws=None):
        """alpha * sum_n mean_s log p - KL (models/vgpmp.py:265-289).  Scalar tensor (or [Bp])."""
        X = self._eng.dev(data).reshape(-1, self.num_latent_gps)
        dims = self._dims(X.shape[0])
        out = self._eng.elbo_fwd_bwd(dims, self._params(X), self._make_draws(dims, draws), need_grad=False)
        return self._squeeze(out["elbo"])

    def elbo_and_grads(self, data, draws=None, want_aux=False):
        X = self._eng.dev(data).reshape(-1, self.num_latent_gps)
        dims = self._dims(X.shape[0])
        return self._eng.elbo_fwd_bwd(dims, self._params(X), self._make_draws(dims, draws), need_grad=True,
                                      want_aux=want_aux)

    def maximum_log_likelihood_objective(self, data):
        return self.elbo(data)

    def training_loss(self, data):
        return -self.elbo(data)

    def training_loss_closure(self, data, **_):
        X = self._eng.dev(data).reshape(-1, self.num_latent_gps)

        def closure():
            return -self.elbo(X)
        closure.model, closure.data = self, X
        return closure

    def _adam_struct(self, lo: int = 0, hi: Optional[int] = None) -> _cabi.Adam:
        """Adam state of problems [lo, hi) (every buffer is problem-major, so a slice is contiguous)."""
        o, t = self.optimizer, self.trainable
        hi = self.num_problems if hi is None else hi
        per = self._adam_m.numel() // self.num_problems
        st = _cabi.Adam(self._q_mu[lo:hi].data_ptr(), self._q_sqrt[lo:hi].data_ptr(),
                        self._raw_lengthscales[lo:hi].data_ptr(), self._raw_variances[lo:hi].data_ptr(),
                        self._lengthscales[lo:hi].data_ptr(), self._variances[lo:hi].data_ptr(),
                        self._adam_m[lo * per:hi * per].data_ptr(), self._adam_v[lo * per:hi * per].data_ptr(),
                        self._variance_lower, o.learning_rate, o.beta_1, o.beta_2, o.epsilon, self._step, int(t["q_mu"]),
                        int(t["q_sqrt"]), int(t["lengthscales"]), int(t["kernel_variance"]))
        return st

    # ---- batching plan ------------------------------------------------------------------------------
    max_workspace_bytes = 32 << 30   # per launch: workspace + draws; larger batches run as several problem chunks

    def _inputs_are_reference_grids(self, X) -> bool:
        """X and Z equispaced and identical across columns (what init_trainset / initialize_Z always produce)?  Same
        criterion as the device-side probe (8 ulp).  One device->host read per distinct X buffer."""
        key = (X.data_ptr(), tuple(X.shape), self._Z.data_ptr())
        if getattr(self, "_grid_key", None) != key:
            def ok(T):
                n = T.shape[0]
                if n < 2:
                    return False
                t0, t1 = T[0, 0], T[-1, 0]
                ref = t0 + (t1 - t0) / (n - 1) * torch.arange(n, dtype=T.dtype, device=T.device)
                tol = 8.0 * 2.220446049250313e-16 * torch.clamp(torch.maximum(t0.abs(), t1.abs()), min=1e-300)
                return bool(((T - ref[:, None]).abs() <= tol).all())
            self._grid_key, self._grid_ok = key, ok(X) and ok(self._Z)
        return self._grid_ok

    def _plan(self, X, explicit_draws: bool):
        """-> (chunks [(lo, hi)], lazy_only): problem chunks that keep workspace + draws under `max_workspace_bytes`, and
        whether omega / tau / w can stay unallocated (lazy draws, inputs are the reference's grids, and the sampler the
        library picks for this shape generates them in-kernel)."""
        eng, N, Bp = self._eng, X.shape[0], self.num_problems
        one = self._dims(N, Bp=1)
        lazy_only = (not explicit_draws and self.lazy_draws and self._inputs_are_reference_grids(X)
                     and bool(eng.lib.vgpmp_sampler_generates_draws(eng.h, C.byref(one))))
        per = int(eng.lib.vgpmp_workspace_bytes(eng.h, C.byref(one)))
        per += int(eng.lib.vgpmp_draws_bytes_lazy(C.byref(one), eng.D) if lazy_only
                   else (1 if self.lazy_draws else 2) * eng.lib.vgpmp_draws_bytes(C.byref(one), eng.D))
        nchunks = max(1, -(-Bp * per // int(self.max_workspace_bytes)))
        size = -(-Bp // nchunks)
        return [(lo, min(Bp, lo + size)) for lo in range(0, Bp, size)], lazy_only

    def train_step(self, X, draws=None):
        """One optimization_step (utils/miscellaneous.py:68-84): ELBO forward + reverse + Adam; returns loss = -ELBO.
        Batches whose workspace would exceed `max_workspace_bytes` run as consecutive problem chunks (independent
        problems; the Philox keys use global problem indices, so chunking does not change a single bit)."""
        eng, D, Bp, M = self._eng, self.num_latent_gps, self.num_problems, self.num_inducing
        X = eng.dev(X).reshape(-1, D)
        chunks, lazy_only = self._plan(X, draws is not None)
        if self._shard is not None and len(chunks) > 1:
            raise NotImplementedError("sample-sharded mode holds one problem per model; it does not chunk")
        if self._grads is None or self._grads["elbo"].shape[0] != Bp:
            self._grads = dict(elbo=eng.empty(Bp), d_q_mu=eng.empty(Bp, M, D), d_q_sqrt=eng.empty(Bp, D, M, M),
                               d_lengthscales=eng.empty(Bp, D), d_variances=eng.empty(Bp, D))
        for lo, hi in chunks:
            self._train_chunk(X, lo, hi, draws, lazy_only, single=len(chunks) == 1)
        self._step += 1
        return self._squeeze(-self._grads["elbo"])

    def _train_chunk(self, X, lo, hi, draws, lazy_only, single):
        eng = self._eng
        dims = self._dims(X.shape[0], Bp=hi - lo)
        poff = self.problem_offset + lo
        soff = self._shard["offset"] if self._shard is not None else 0
        slot = None
        if draws is not None:
            use = self._make_draws(dims, {k: v[lo:hi] for k, v in draws.items()} if not single else draws)
        else:
            key = (dims.num_problems, dims.num_samples, dims.num_bases, dims.total_samples, lazy_only)
            if self._pipe is None or self._pipe["key"] != key:
                # lazy draws need one set (of which only eps_u / eps_j are ever written); the prefetch pipeline needs two
                self._pipe = dict(key=key, sets=[eng.alloc_draws(dims, lazy_only=lazy_only)], ready=None)
            if self.lazy_draws:
                # omega / tau / w are generated inside the sampler kernel from the same Philox keys: nothing to prefetch
                use = eng.rng_fill_lazy(dims, self.seed, self._step, self._pipe["sets"][0], problem_offset=poff,
                                        sample_offset=soff)
            elif not single:
                use = eng.rng_fill(dims, self.seed, self._step, self._pipe["sets"][0], problem_offset=poff, sample_offset=soff)
            else:
                # materialised draws, double-buffered: this step's set was generated on the side stream during the previous step
                if len(self._pipe["sets"]) < 2:
                    self._pipe["sets"].append(eng.alloc_draws(dims))
                slot = self._step & 1
                if self._pipe["ready"] == self._step:
                    eng.rng_join(slot)
                else:
                    eng.rng_fill(dims, self.seed, self._step, self._pipe["sets"][slot], problem_offset=poff, sample_offset=soff)
                use = self._pipe["sets"][slot]
        params = eng.params_struct(self._q_mu[lo:hi], self._q_sqrt[lo:hi], self._lengthscales[lo:hi], self._variances[lo:hi],
                                   self._query_states[lo:hi], self._Z, X)
        if self._shard is not None:
            from ..utils.sharding import allreduce_packed, packed_views
            views = packed_views(self._shard["flat"], self.num_problems, self.num_inducing, self.num_latent_gps)
            out = eng.elbo_fwd_bwd(dims, params, use, need_grad=True, out=views)
            allreduce_packed(self._shard["flat"], self._shard["group"])      # one collective: gradients || ELBO
            self._grads = out
        else:
            out = eng.elbo_fwd_bwd(dims, params, use, need_grad=True, out={k: v[lo:hi] for k, v in self._grads.items()})
        if slot is not None:
            eng.rng_release(slot)
            eng.rng_fill_async(dims, self.seed, self._step + 1, self._pipe["sets"][slot ^ 1], slot ^ 1,
                               problem_offset=poff, sample_offset=soff)
            self._pipe["ready"] = self._step + 1
        gs = _cabi.Grads(out["d_q_mu"].data_ptr(), out["d_q_sqrt"].data_ptr(), out["d_lengthscales"].data_ptr(),
                         out["d_variances"].data_ptr())
        st = self._adam_struct(lo, hi)
        eng._chk(eng.lib.vgpmp_adam_step(eng.h, C.byref(dims), C.byref(st), C.byref(gs), eng._stream()), "adam_step")

    def train_step_host(self, X_host: torch.Tensor, wait: bool = True, stream: int = None, loss_out: torch.Tensor = None):
        """The same optimisation step driven from HOST buffers through `vgpmp_train_step_host`: X [N,D] is read from
        (pinned) host memory, copied to the device, the step's randomness is drawn on the device, and loss = -ELBO [Bp]
        is copied back to pinned host memory before the call returns (like `loss = tf_optimization_step(...)` feeding
        the tqdm readout, utils/miscellaneous.py:101-103).  Returns a CPU tensor (a copy of the pinned loss buffer).
        `stream` (a raw CUDA stream handle; default: torch's current stream) and `loss_out` (a pinned float64 [Bp] buffer the
        loss is written to instead of the model's own) are for callers that drive several models, see StreamedVGPMP.

        The steady state is ONE ctypes call: the argument block is cached per (host buffer, stream, optimiser settings)."""
        hs = getattr(self, "_host_state", None)
        if self._shard is not None or (hs is not None and hs.get("generic_numel") == X_host.numel()):
            return self._train_step_host_generic(X_host, wait, stream, loss_out)
        o, t = self.optimizer, self.trainable
        sid = stream if stream is not None else torch.cuda.current_stream(self._eng.device).cuda_stream
        key = (X_host.data_ptr(), X_host.numel(), sid, self.seed, self.problem_offset, o.learning_rate, o.beta_1, o.beta_2,
               o.epsilon, t["q_mu"], t["q_sqrt"], t["lengthscales"], t["kernel_variance"],
               None if loss_out is None else loss_out.data_ptr())
        call = hs["call"] if hs is not None else None
        if call is None or call["key"] != key:
            call = self._host_call(X_host, sid, loss_out, key)
            if call is None:
                return self._train_step_host_generic(X_host, wait, stream, loss_out)
        if hs is None:
            hs = self._host_state
        if hs["pending"]:
            raise RuntimeError("train_step_host(wait=False) was called twice without train_step_host_wait(): the pinned "
                               "loss buffer of the step in flight would be overwritten")
        st = call["st"]
        st.step = self._step
        rc = call["begin"](*call["args"])
        if rc:
            self._eng._chk(rc, "train_step_host_begin")
        self._step = st.step
        hs["pending"] = True
        if not wait:
            return None
        return self.train_step_host_wait()

    def _host_call(self, X_host, sid, loss_out, key):
        """Slow path of `train_step_host`: checks the host tensor, (re)allocates the per-shape device state and builds the
        argument block of the C call."""
        eng, D = self._eng, self.num_latent_gps
        if X_host.device.type != "cpu" or X_host.dtype != torch.float64 or not X_host.is_contiguous():
            raise TypeError("train_step_host expects a contiguous float64 CPU tensor (pinned for async copies)")
        N = X_host.numel() // D
        dims = self._dims(N)
        if len(self._plan(eng.dev(X_host.reshape(N, D)), False)[0]) != 1:
            # more than one workspace chunk: the single C call (and its CUDA graph) covers one chunk, so this batch steps
            # through train_step with the host copies around it
            hs = getattr(self, "_host_state", None) or dict(N=-1, pending=False, call=None)
            hs["generic_numel"] = X_host.numel()
            self._host_state = hs
            return None
        hs = getattr(self, "_host_state", None)
        if hs is None or hs["N"] != N:
            Xc = X_host.reshape(N, D)
            probe = eng.dev(Xc)
            compact = (self.lazy_draws and self._inputs_are_reference_grids(probe)
                       and bool(eng.lib.vgpmp_sampler_generates_draws(eng.h, C.byref(dims))))
            # compact: eps_u / eps_j only (lazy draws never materialised); else two full sets (next step drawn on the side stream)
            nbytes = int(eng.lib.vgpmp_draws_bytes_lazy(C.byref(dims), D)) if compact \
                else 2 * int(eng.lib.vgpmp_draws_bytes(C.byref(dims), D))
            hs = dict(N=N, X_dev=eng.empty(N, D), draws=torch.empty(nbytes, dtype=torch.uint8, device=eng.device),
                      elbo=eng.empty(self.num_problems),
                      loss=torch.empty(self.num_problems, dtype=torch.float64).pin_memory(), pending=False, call=None,
                      g=dict(d_q_mu=eng.empty(self.num_problems, self.num_inducing, D),
                             d_q_sqrt=eng.empty(self.num_problems, D, self.num_inducing, self.num_inducing),
                             d_lengthscales=eng.empty(self.num_problems, D), d_variances=eng.empty(self.num_problems, D)))
            self._host_state = hs
        if loss_out is not None and (loss_out.device.type != "cpu" or loss_out.dtype != torch.float64
                                     or loss_out.numel() != self.num_problems or not loss_out.is_contiguous()):
            raise TypeError("loss_out must be a contiguous float64 CPU tensor with one entry per problem")
        loss = hs["loss"] if loss_out is None else loss_out
        g = hs["g"]
        gs = _cabi.Grads(g["d_q_mu"].data_ptr(), g["d_q_sqrt"].data_ptr(), g["d_lengthscales"].data_ptr(),
                         g["d_variances"].data_ptr())
        st = self._adam_struct()
        ws = eng.workspace(dims)
        stream = C.c_void_p(sid)
        call = dict(key=key, st=st, gs=gs, dims=dims, stream=stream, keep=(ws, X_host, loss), loss=loss,
                    begin=eng.lib.vgpmp_train_step_host_begin, end=eng.lib.vgpmp_train_step_host_end,
                    args=(eng.h, C.byref(dims), C.byref(st), self._query_states.data_ptr(), self._Z.data_ptr(),
                          X_host.data_ptr(), hs["X_dev"].data_ptr(), self.seed, int(self.problem_offset),
                          hs["draws"].data_ptr(), hs["draws"].numel(), C.byref(gs), hs["elbo"].data_ptr(),
                          loss.data_ptr(), ws.data_ptr(), ws.numel(), stream),
                    end_args=(eng.h, C.byref(dims), loss.data_ptr(), stream))
        hs["call"] = call
        return call

    def _train_step_host_generic(self, X_host, wait, stream, loss_out):
        """Host-buffer step for models the one-call path does not cover (sample-sharded: the all-reduce sits between the
        reverse pass and Adam; batches of several workspace chunks): pinned X -> device, `train_step`, loss -> pinned host."""
        eng, D = self._eng, self.num_latent_gps
        if X_host.device.type != "cpu" or X_host.dtype != torch.float64 or not X_host.is_contiguous():
            raise TypeError("train_step_host expects a contiguous float64 CPU tensor (pinned for async copies)")
        if stream is not None and stream != torch.cuda.current_stream(eng.device).cuda_stream:
            raise NotImplementedError("the generic host step runs on torch's current stream")
        N = X_host.numel() // D
        hs = getattr(self, "_host_state", None) or dict(N=-1, pending=False, call=None)
        self._host_state = hs
        gs = hs.get("generic")
        if gs is None or gs["N"] != N:
            gs = dict(N=N, X_dev=eng.empty(N, D), loss=torch.empty(self.num_problems, dtype=torch.float64).pin_memory())
            hs["generic"] = gs
        if hs["pending"]:
            raise RuntimeError("train_step_host(wait=False) was called twice without train_step_host_wait(): the pinned "
                               "loss buffer of the step in flight would be overwritten")
        gs["X_dev"].copy_(X_host.reshape(N, D), non_blocking=True)
        loss = self.train_step(gs["X_dev"])
        out = gs["loss"] if loss_out is None else loss_out
        out.copy_(loss.reshape(-1), non_blocking=True)
        hs["pending"], hs["generic_out"] = True, out
        if not wait:
            return None
        return self.train_step_host_wait()

    def train_step_host_wait(self, copy: bool = True):
        """Second half of `train_step_host(..., wait=False)`: blocks until the loss of the step in flight is in host
        memory and returns a copy of it (the pinned buffer is re-used by the next step; `copy=False` returns the buffer)."""
        hs = self._host_state
        if hs.get("generic_out") is not None:
            torch.cuda.current_stream(self._eng.device).synchronize()
            hs["pending"] = False
            out, hs["generic_out"] = hs["generic_out"], None
            hs["generic_last"] = out
            return out.clone() if copy else out
        if hs["call"] is None and hs.get("generic_last") is not None:      # a second wait after a generic step: idempotent
            return hs["generic_last"].clone() if copy else hs["generic_last"]
        call = hs["call"]
        rc = call["end"](*call["end_args"])
        if rc:
            self._eng._chk(rc, "train_step_host_end")
        hs["pending"] = False
        return call["loss"].clone() if copy else call["loss"]

    def predict_f_samples(self, X, num_samples=None, draws=None):
        """temporary_paths + predict_f_samples (models/vgpmp.py:281-282): [S,N,D] latent samples."""
        S = self.num_samples if num_samples is None else int(num_samples)
        dims = self._dims(1, S)
        f = self._eng.pathwise_sample(dims, self._params(None), self._make_draws(dims, draws), X)
        return self._squeeze(f)

    def predict_f_mean(self, X):
        """`self.posterior().predict_f(X)[0]`: SVGP posterior mean in latent space, [N,D] (or [Bp,N,D])."""
        return self._squeeze(self._eng.predict_f_mean(self._dims(1, 1), self._params(None), X))

    def sample_from_posterior(self, X, robot=None, compute_uncertainty=False, num_samples=150):
        """models/vgpmp.py:312-331: (mean trajectory, best of 150 posterior samples, the first 7 samples,
        2*sqrt(uncertainty)) in joint space.  `compute_uncertainty` needs the pybullet robot in the reference and is
        not part of the hot path; the constant 1.0 of its False branch is returned."""
        if compute_uncertainty:
            raise NotImplementedError("end-effector uncertainty uses the pybullet robot (models/vgpmp.py:322-327)")
        sig = self.likelihood.joint_sigmoid
        mu = sig(self.predict_f_mean(X))
        samples = sig(self.predict_f_samples(X, num_samples=num_samples))          # [S,N,D] or [Bp,S,N,D]
        best = self.get_best_sample(samples)
        if self.num_problems == 1:
            best_sample = samples[best]
            first = samples[:7]
        else:
            idx = best.view(-1, 1, 1, 1).expand(-1, 1, samples.shape[2], samples.shape[3])
            best_sample = torch.gather(samples, 1, idx)[:, 0]
            first = samples[:, :7]
        return mu, best_sample, first, 2.0 * torch.ones((), dtype=torch.float64, device=samples.device)

    def collision_free(self, trajectory):
        """Verdict of SURVEY.md 8f-3 on joint-space trajectories [N,D] (or [Bp,N,D]): min over timesteps and spheres of
        (sdf - radius) > 0.  Returns (verdict, min clearance)."""
        clr = self._eng.clearance(trajectory).amin(dim=-1)
        return clr > 0, clr

    def debug_likelihood(self, data):
        lp = self.likelihood.log_prob(data)
        return torch.sum(torch.mean(lp, dim=0))

    def get_best_sample(self, samples):
        """argmax_s sum_n log_prob (models/vgpmp.py:336-339)."""
        cost = torch.sum(self.likelihood.log_prob(samples), dim=-1)
        return torch.argmax(cost, dim=-1)

    def initialize_optimizer(self, learning_rate):
        return AdamConfig(learning_rate)
