"""Kernel objects (reference: gpflow_vgpmp/kernels/kernels.py:4-16 + the GPflow classes they subclass).

Host-side parameter holders / dispatch tags only: the Matern-5/2 arithmetic runs in csrc/gp.cu.  Values are
constrained-space floats until a VGPMP model adopts the kernel, after which `.lengthscales` / `.variance`
read the model's device-resident state.
"""
from __future__ import annotations

import numpy as np

__all__ = ["Kernel", "Matern52", "SeparateIndependent", "SharedIndependent", "VanillaConditioningSeparateIndependent",
           "VanillaConditioningSharedIndependent", "FirstOrderKernelDerivativeSeparateIndependent"]


class Kernel:
    pass


class Matern52(Kernel):
    """k(r) = variance (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r), r = |x - x'| / lengthscales (GPflow Matern52)."""

    def __init__(self, variance=1.0, lengthscales=1.0, name=None):
        if not float(variance) > 0 or not float(lengthscales) > 0:
            raise ValueError("Matern52 needs variance > 0 and lengthscales > 0")
        self._variance, self._lengthscales = float(variance), float(lengthscales)
        self.trainable = {"variance": True, "lengthscales": True}
        self._model, self._index = None, None
        self.name = name or "matern52"

    def _bind(self, model, index):
        self._model, self._index = model, index

    @property
    def variance(self):
        if self._model is not None:
            return self._model._variances[..., self._index]
        return self._variance

    @property
    def lengthscales(self):
        if self._model is not None:
            return self._model._lengthscales[..., self._index]
        return self._lengthscales


class SeparateIndependent(Kernel):
    def __init__(self, kernels, name=None):
        self.kernels = list(kernels)
        self.name = name
        if not self.kernels or not all(isinstance(k, Matern52) for k in self.kernels):
            raise NotImplementedError("the CUDA path implements Matern52 latent kernels only")

    @property
    def num_latent_gps(self):
        return len(self.kernels)

    def hyper_arrays(self):
        ls = np.array([float(np.asarray(k._lengthscales)) for k in self.kernels])
        var = np.array([float(np.asarray(k._variance)) for k in self.kernels])
        return ls, var


class SharedIndependent(Kernel):
    def __init__(self, kernel, output_dim, name=None):
        self.kernel, self.output_dim, self.name = kernel, int(output_dim), name


class VanillaConditioningSeparateIndependent(SeparateIndependent):
    """The live dispatch key of the reference (models/vgpmp.py:136-142)."""


class VanillaConditioningSharedIndependent(SharedIndependent):
    pass


class FirstOrderKernelDerivativeSeparateIndependent(SeparateIndependent):
    """Velocity-conditioning variant; never built by VGPMP.initialize (dead branch of the reference)."""
