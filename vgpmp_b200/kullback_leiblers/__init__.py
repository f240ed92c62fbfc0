"""`prior_kl` (reference: gpflow_vgpmp/kullback_leiblers/prior_kl.py:13-35): KL(q(u) || p(u | start, goal)) with the
prior mean obtained by conditioning on the two query states, whitened with chol(Kuu).  Runs in gp_prepare_kernel."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ..covariances import kernel_hypers, unwrap_inducing
from ..engine import Engine

__all__ = ["prior_kl"]


def prior_kl(inducing_variable, kernel, q_mu, q_sqrt, query_states):
    """q_mu [M,L], q_sqrt [L,M,M] (lower), query_states [2,L] in latent space -> scalar tensor."""
    Z = unwrap_inducing(inducing_variable)
    ls, var = kernel_hypers(kernel)
    M, D = Z.shape
    eng = Engine.for_gp(D)
    dims = eng.dims(1, M, 1, 1, 1)
    bufs = [eng.dev(q_mu).reshape(1, M, D), eng.dev(q_sqrt).reshape(1, D, M, M), eng.dev(ls).reshape(1, D),
            eng.dev(var).reshape(1, D), eng.dev(query_states).reshape(1, 2, D), eng.dev(Z)]
    params = eng.params_struct(*bufs, None)
    _, _, kl = eng.gp_prepare(dims, params)
    return kl[0]
