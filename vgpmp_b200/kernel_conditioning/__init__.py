"""`K_conditioned` (reference: gpflow_vgpmp/kernel_conditioning/dispatch.py:3, cond_kernel.py:7-22,
multioutput/cond_kernel.py:17-48): per-latent 1-D kernel Gram matrices stacked to [L,|Z|,|X|]."""
from __future__ import annotations

import numpy as np

from ..covariances import kernel_hypers
from ..engine import Engine
from ..inducing_variables import ConditionedVariableInducingPoints

__all__ = ["K_conditioned"]


def K_conditioned(Z, X, kernel):
    """Z: ConditionedVariableInducingPoints or an array [A,D] that starts with the rows t=0, t=1; X: [N,D]."""
    Zy = Z.Zy if isinstance(Z, ConditionedVariableInducingPoints) else np.asarray(Z, dtype=np.float64)
    Xn = X.Zy if isinstance(X, ConditionedVariableInducingPoints) else X
    if Zy.shape[0] < 3 or not (np.all(Zy[0] == 0.0) and np.all(Zy[1] == 1.0)):
        raise NotImplementedError("K_conditioned expects Zy = [0; 1; Z] (inducing_variables.py:56-62)")
    ls, var = kernel_hypers(kernel)
    eng = Engine.for_gp(Zy.shape[1])
    K = eng.kuf(Zy[2:], Xn, ls, var)
    return K[0] if K.shape[0] == 1 else K
