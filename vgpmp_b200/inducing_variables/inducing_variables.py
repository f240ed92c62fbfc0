"""Inducing points with the two conditioned timesteps prepended
(reference: gpflow_vgpmp/inducing_variables/inducing_variables.py:26-82)."""
from __future__ import annotations

import numpy as np

__all__ = ["ConditionedVariableInducingPoints", "SharedIndependentInducingVariables",
           "SeparateIndependentInducingVariables", "InducingPointsInterface"]


class InducingPointsInterface:
    def __init__(self, Z, conditioned_timesteps, name=None):
        self._Z = np.asarray(Z, dtype=np.float64)
        self.conditioned_timesteps = np.asarray(conditioned_timesteps, dtype=np.float64)
        assert self._Z.ndim == 2 and self.conditioned_timesteps.ndim == 2
        assert self._Z.shape[1] == self.conditioned_timesteps.shape[1], \
            "The number of degrees of freedom of the trainable inducing points and the conditioned timesteps must " \
            "be the same. Right now it is {} and {}, respectively.".format(self._Z.shape[1],
                                                                           self.conditioned_timesteps.shape[1])
        self.len_ny = self.conditioned_timesteps.shape[0]
        self.name = name
        self.trainable = False

    @property
    def num_inducing(self) -> int:
        return len(self)

    def __len__(self) -> int:
        return int(self._Z.shape[0])

    @property
    def ny(self):
        return self.conditioned_timesteps

    @property
    def Z(self):
        return np.concatenate([self.ny, self._Z], axis=0)


class ConditionedVariableInducingPoints(InducingPointsInterface):
    @property
    def Zy(self):
        return np.concatenate([self.ny, self._Z], axis=0)


class SharedIndependentInducingVariables:
    def __init__(self, inducing_variable):
        self.inducing_variable = inducing_variable

    @property
    def num_inducing(self):
        return self.inducing_variable.num_inducing


class SeparateIndependentInducingVariables:
    def __init__(self, inducing_variable_list):
        self.inducing_variable_list = list(inducing_variable_list)

    @property
    def num_inducing(self):
        return self.inducing_variable_list[0].num_inducing
