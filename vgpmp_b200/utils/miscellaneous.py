"""Planning driver for the hot path (reference: gpflow_vgpmp/utils/miscellaneous.py:68-127,141-343).

`optimization_step` / `training_loop` / `init_trainset` / `disable_param_opt` keep the reference names and argument
meaning.  pybullet drawing / trajectory execution are out of scope; `solve_planning_problem` here batches problems.
"""
from __future__ import annotations

import itertools
import json
from pathlib import Path
from typing import Optional

import numpy as np
import torch

DATA_DIR = Path(__file__).resolve().parents[1] / "data"

__all__ = ["optimization_step", "training_loop", "init_trainset", "disable_param_opt", "set_trainable",
           "load_problemset", "default_trainable_params", "solve_planning_problems"]


def optimization_step(model, closure, optimizer=None):
    """One Adam step on loss = closure() = -ELBO (miscellaneous.py:68-84).  The tape + apply_gradients of the reference
    are one fused call: forward, reverse pass and the update all run on the GPU."""
    data = getattr(closure, "data", None)
    if data is None:
        raise TypeError("closure must come from model.training_loss_closure(data)")
    if optimizer is not None:
        model.optimizer = optimizer
    return model.train_step(data)


def training_loop(model, data, num_steps, print_summary=False, randomize=False, callback=None):
    """miscellaneous.py:87-112.  Returns the list of losses (the reference only shows them in tqdm)."""
    if randomize:
        raise NotImplementedError("randomize=True resamples timesteps with tf.random; not on the benchmark path")
    closure = model.training_loss_closure(data)
    losses = []
    for it in range(num_steps):
        loss = optimization_step(model, closure, model.optimizer)
        losses.append(loss)
        if callback is not None:
            callback(it, loss)
    if print_summary:
        for kern in model.kernel.kernels:
            print(f"model lengthscale: {kern.lengthscales} \nmodel variance: {kern.variance}")
        print(f"model q_mu: {model.q_mu}")
    return losses


def init_trainset(grid_spacing_X, grid_spacing_Xnew, input_dimension, degree_of_freedom, start_joints, end_joints,
                  scale=100, end_time=1):
    """X [N,D], y [2,D], Xnew [Nnew,D]; every column of X is linspace(0, end_time*scale, N) (miscellaneous.py:115-127)."""
    X = np.repeat(np.linspace(0, end_time * scale, grid_spacing_X)[:, None], input_dimension, axis=1)
    Xnew = np.repeat(np.linspace(0, end_time * scale, grid_spacing_Xnew)[:, None], input_dimension, axis=1)
    y = np.concatenate([np.asarray(start_joints, dtype=np.float64).reshape(1, degree_of_freedom),
                        np.asarray(end_joints, dtype=np.float64).reshape(1, degree_of_freedom)], axis=0)
    return X.astype(np.float64), y, Xnew.astype(np.float64)


def set_trainable(obj, flag: bool):
    if hasattr(obj, "trainable") and not isinstance(obj.trainable, dict):
        obj.trainable = bool(flag)
    else:
        raise TypeError("set_trainable: use model.trainable[...] for model-owned parameters")


def default_trainable_params() -> dict:
    """parameters.yaml:35-42 of the reference."""
    return json.loads((DATA_DIR / "trainable_params.json").read_text())


def disable_param_opt(planner, trainable_params):
    """Trainable flags (miscellaneous.py:324-343).  The two priors the reference attaches there sit on parameters that are
    non-trainable under the default flags, so they add a constant to the loss and nothing to any gradient."""
    for key in ("sigma_obs", "inducing_variable", "alpha"):
        if trainable_params.get(key, False):
            raise NotImplementedError(f"training {key} is outside the hot path (parameters.yaml:35-42 keeps it fixed)")
    planner.trainable.update(q_mu=bool(trainable_params["q_mu"]), q_sqrt=bool(trainable_params["q_sqrt"]),
                             lengthscales=bool(trainable_params["lengthscales"]),
                             kernel_variance=bool(trainable_params["kernel_variance"]))


def load_problemset(robot_name: str, environment: str) -> dict:
    """Tables of data/problemsets/<robot>.py: states, queries = combinations(states, 2) (parameter_loader.py:138),
    robot pos/orn, scene offset, planner_params."""
    entry = json.loads((DATA_DIR / "problemsets.json").read_text())[robot_name][environment]
    states = [np.asarray(s, dtype=np.float64) for s in entry["states"]]
    queries = list(itertools.combinations(states, 2))
    return dict(states=states, queries=queries, robot_pos_and_orn=entry["pos_and_orn"],
                scene_offset=entry["object_positions"][0] if entry["object_positions"] else [0.0, 0.0, 0.0],
                planner_params=dict(entry["planner_params"]))


def solve_planning_problems(sdf, robot, sampler, queries, scene_offset, planner_params, trainable_params=None,
                            seed: int = 0, num_steps: Optional[int] = None):
    """Batched counterpart of solve_planning_problem (miscellaneous.py:141-321): build the model for all start/goal pairs,
    run the training loop, return the model and the per-step losses [steps, Bp]."""
    from ..models.vgpmp import VGPMP
    pp = dict(planner_params)
    steps = pp["num_steps"] if num_steps is None else num_steps
    dof = robot.dof
    q = np.stack([np.stack([np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)]) for a, b in queries])
    X, _, Xnew = init_trainset(pp["time_spacing_X"], pp["time_spacing_Xnew"], dof, dof, q[0, 0], q[0, 1], scale=1)
    planner = VGPMP.initialize(sdf=sdf, robot=robot, sampler=sampler, query_states=q, scene_offset=scene_offset,
                               q_mu=None, interpolation_method="linear", seed=seed, **pp)
    disable_param_opt(planner, trainable_params or default_trainable_params())
    losses = training_loop(planner, X, steps)
    return planner, torch.stack([l.reshape(-1) for l in losses]) if losses else None, Xnew
