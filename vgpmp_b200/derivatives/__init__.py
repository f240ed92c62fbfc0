"""Fused reverse pass of the ELBO hot path.

In the reference `gpflow_vgpmp/derivatives/` holds kernel-derivative covariances for velocity conditioning, a
branch `VGPMP.initialize` never builds; the gradient of the ELBO itself comes from TF
autodiff (utils/miscellaneous.py:68-84).  Here the name is kept for the hand-written reverse pass:

    d ELBO / d f              loglik_bwd_kernel<D>     (SDF custom gradient -> sphere wrenches -> joint axes -> sigmoid)
    d ELBO / d {_q_mu, _q_sqrt, lengthscales, variances}
                              gp_backward_kernel, gp_backward_samples_kernel   (pathwise update, Cholesky reverse, KL)
"""
from __future__ import annotations


def elbo_gradients(model, X, draws=None):
    """{'elbo', 'd_q_mu', 'd_q_sqrt', 'd_lengthscales', 'd_variances'} for a VGPMP model (constrained space)."""
    return model.elbo_and_grads(X, draws=draws)


def loglik_gradient(likelihood, F, squash=False, upstream=1.0):
    """(logp, upstream * d logp / d F) through FK, sphere placement, the SDF custom gradient and the hinge."""
    return likelihood._engine().loglik(F, squash=squash, upstream=upstream, need_grad=True)
