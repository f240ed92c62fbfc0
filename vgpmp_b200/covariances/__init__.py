"""`Kuu`, `Kuf`, `Kfu` for the conditioned inducing variables (reference: gpflow_vgpmp/covariances/multioutput/
Kuus.py:42-53, Kufs.py:26-34, covariances/Kfus.py:36-42,67-75).  The reference registers these into GPflow's /
GPflowSampling's multiple-dispatch tables; here they are plain functions over the same argument types, backed by
`vgpmp_kuu` / `vgpmp_kuf` (csrc/gp.cu)."""
from __future__ import annotations

import numpy as np
import torch

from ..engine import Engine
from ..inducing_variables import (ConditionedVariableInducingPoints, SeparateIndependentInducingVariables,
                                  SharedIndependentInducingVariables)
from ..kernels import FirstOrderKernelDerivativeSeparateIndependent, Matern52, SeparateIndependent

__all__ = ["Kuu", "Kuf", "Kfu", "unwrap_inducing", "kernel_hypers"]


def unwrap_inducing(iv) -> np.ndarray:
    """-> trainable inducing inputs _Z [M,D] (the 2 conditioned timesteps are implied: Zy = [0; 1; _Z])."""
    if isinstance(iv, SharedIndependentInducingVariables):
        iv = iv.inducing_variable
    if isinstance(iv, SeparateIndependentInducingVariables):
        cols = [v._Z[:, 0] if v._Z.shape[1] == 1 else v._Z[:, i] for i, v in enumerate(iv.inducing_variable_list)]
        ny = iv.inducing_variable_list[0].ny
        return _check_ny(np.stack(cols, axis=1), ny)
    if isinstance(iv, ConditionedVariableInducingPoints):
        return _check_ny(iv._Z, iv.ny)
    raise NotImplementedError(f"no Kuu/Kuf registration for {type(iv).__name__}")


def _check_ny(Z, ny):
    if ny.shape[0] != 2 or not (np.all(ny[0] == 0.0) and np.all(ny[1] == 1.0)):
        raise NotImplementedError("the CUDA path conditions on exactly the timesteps t=0 and t=1 (models/vgpmp.py:144-146)")
    return Z


def kernel_hypers(kernel):
    if isinstance(kernel, FirstOrderKernelDerivativeSeparateIndependent):
        raise NotImplementedError("derivative-conditioned kernels are a dead branch of the reference (SURVEY.md 2 #7)")
    if isinstance(kernel, Matern52):
        kernel = SeparateIndependent([kernel])
    if not isinstance(kernel, SeparateIndependent):
        raise NotImplementedError(f"no registration for kernel type {type(kernel).__name__}")
    ls = [k.lengthscales for k in kernel.kernels]
    var = [k.variance for k in kernel.kernels]
    if isinstance(ls[0], torch.Tensor):
        return torch.stack(ls, dim=-1), torch.stack(var, dim=-1)
    return np.array(ls, dtype=np.float64), np.array(var, dtype=np.float64)


def Kuu(inducing_variable, kernel, *, jitter: float = 0.0):
    """[L,Mp,Mp] = K(Zy,Zy) + jitter I  (a leading batch axis appears when the kernel is bound to a batched model)."""
    Z = unwrap_inducing(inducing_variable)
    ls, var = kernel_hypers(kernel)
    eng = Engine.for_gp(Z.shape[1])
    K = eng.kuu(Z, ls, var, jitter)
    return K[0] if K.shape[0] == 1 else K


def Kuf(inducing_variable, kernel, Xnew):
    """[L,Mp,N] = K(Zy, Xnew), per-latent 1-D kernels on matching columns (cond_kernel.py:17-25)."""
    Z = unwrap_inducing(inducing_variable)
    ls, var = kernel_hypers(kernel)
    eng = Engine.for_gp(Z.shape[1])
    K = eng.kuf(Z, Xnew, ls, var)
    return K[0] if K.shape[0] == 1 else K


def Kfu(inducing_variable, kernel, X, **kwargs):
    """[L,N,Mp]: transpose of Kuf (covariances/Kfus.py:36-42)."""
    return Kuf(inducing_variable, kernel, X).transpose(-1, -2)
