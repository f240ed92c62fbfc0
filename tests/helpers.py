"""Builds matched (oracle problem, CUDA model) pairs on identical inputs.  Test infrastructure only."""
import json
from pathlib import Path

import numpy as np

from oracle import vgpmp_oracle as O

ROOT = Path(__file__).resolve().parents[1]
DATA = ROOT / "vgpmp_b200" / "data"


def tables():
    return json.loads((DATA / "robots.json").read_text()), json.loads((DATA / "problemsets.json").read_text())


def oracle_robot(name, env="bookshelves"):
    robots, probs = tables()
    return O.OracleRobot.from_tables(robots[name], probs[name][env]["pos_and_orn"])


def small_sdf(seed=0, shape=(48, 48, 48), delta=0.04, origin=(-0.96, -0.96, -0.96)):
    from vgpmp_b200.utils.sdf_utils import synthetic_shelf_sdf
    return synthetic_shelf_sdf(shape=shape, delta=delta, origin=origin, seed=seed)


def planner_params(name, env):
    _, probs = tables()
    return dict(probs[name][env]["planner_params"])


def make_case(name="franka", env="bookshelves", num_problems=2, seed=0, S=None, N=None, M=None, B=64, sdf=None,
              perturb=True):
    """Returns dict with: oracle problems (list), product-side objects, params, explicit draws."""
    from vgpmp_b200.utils.miscellaneous import load_problemset
    rng = np.random.default_rng(seed)
    ps = load_problemset(name, env)
    pp = ps["planner_params"]
    S = pp["num_samples"] if S is None else S
    N = pp["time_spacing_X"] if N is None else N
    M = pp["num_inducing"] if M is None else M
    sdf = small_sdf() if sdf is None else sdf
    orob = oracle_robot(name, env)
    D = orob.dof
    X = np.repeat(np.linspace(0, 1, N)[:, None], D, axis=1)
    Z = np.array([np.full(D, t) for t in np.linspace(0.1, 0.9, M)])
    queries = [ps["queries"][i % len(ps["queries"])] for i in range(num_problems)]
    osdf = O.OracleSDF(sdf.data, sdf.origin, sdf.delta)
    oprobs = [O.OracleProblem(robot=orob, sdf=osdf, scene_offset=np.asarray(ps["scene_offset"], dtype=np.float64),
                              sigma_obs=pp["sigma_obs"], epsilon=pp["epsilon"], alpha=pp["alpha"], X=X, Z=Z,
                              query_states=np.stack(q)) for q in queries]
    q_mu = np.stack([O.init_q_mu_linear(p, M) for p in oprobs])                       # [Bp,M,D]
    q_sqrt = np.broadcast_to(np.eye(M), (num_problems, D, M, M)).copy()
    ls = np.broadcast_to(np.asarray(pp["lengthscales"], dtype=np.float64), (num_problems, D)).copy()
    var = np.full((num_problems, D), float(pp["variance"]))
    if perturb:  # move away from the symmetric initial point so every gradient path is exercised
        q_mu = q_mu + 0.3 * rng.standard_normal(q_mu.shape)
        q_sqrt = np.tril(q_sqrt + 0.2 * rng.standard_normal(q_sqrt.shape))
        idx = np.arange(M)
        q_sqrt[..., idx, idx] = np.abs(q_sqrt[..., idx, idx]) + 0.3
        ls = ls * np.exp(0.2 * rng.standard_normal(ls.shape))
        var = var * np.exp(0.2 * rng.standard_normal(var.shape))
    draws = [O.make_draws(rng, D, S, B, M + 2) for _ in range(num_problems)]
    stacked = {k: np.stack([d[k] for d in draws]) for k in draws[0]}
    return dict(name=name, env=env, ps=ps, pp=pp, S=S, N=N, M=M, B=B, D=D, sdf=sdf, X=X, Z=Z, queries=queries,
                oracle=oprobs, q_mu=q_mu, q_sqrt=q_sqrt, ls=ls, var=var, draws=draws, draws_stacked=stacked)


def make_model(case, seed=0):
    """CUDA-side VGPMP for the case, with its state overwritten by the case's (perturbed) parameters."""
    from vgpmp_b200.models import VGPMP
    from vgpmp_b200.utils.robot import Robot
    from vgpmp_b200.utils.sampler import Sampler
    pp = dict(case["pp"])
    pp.update(num_samples=case["S"], num_inducing=case["M"])
    robot = Robot.from_tables(case["name"], case["env"])
    sampler = Sampler(None, robot)
    q = np.stack([np.stack(qq) for qq in case["queries"]])
    model = VGPMP.initialize(sdf=case["sdf"], robot=robot, sampler=sampler, query_states=q,
                             scene_offset=case["ps"]["scene_offset"], num_bases=case["B"], seed=seed, **pp)
    eng = model._eng
    model._q_mu.copy_(eng.dev(case["q_mu"]))
    model._q_sqrt.copy_(eng.dev(case["q_sqrt"]))
    model._lengthscales.copy_(eng.dev(case["ls"]))
    model._variances.copy_(eng.dev(case["var"]))
    model._raw_lengthscales.copy_(eng.dev(O.softplus_inv(case["ls"])))
    model._raw_variances.copy_(eng.dev(O.softplus_inv(case["var"] - model._variance_lower)))
    return model


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
