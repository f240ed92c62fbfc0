#!/usr/bin/env python
"""Per-stage CUDA-event times of one training step at an arbitrary shape (experiments; the numbers that matter are
bench.py's).  Example (BASELINE config-5 shape on a slice of the problems):

    python tools/stage_probe.py --robot franka --env bookshelves --problems 64 --samples 256 --timesteps 64
"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robot", default="franka")
    ap.add_argument("--env", default="bookshelves")
    ap.add_argument("--problems", type=int, default=64)
    ap.add_argument("--samples", type=int, default=256)
    ap.add_argument("--timesteps", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--option", action="append", default=[])
    a = ap.parse_args()
    from vgpmp_b200 import _cabi
    from vgpmp_b200.models import VGPMP
    from vgpmp_b200.utils.gen_sdf import PADDING, mesh_to_sdf, scene_mesh_path
    from vgpmp_b200.utils.miscellaneous import default_trainable_params, disable_param_opt, init_trainset, load_problemset
    from vgpmp_b200.utils.robot import Robot
    from vgpmp_b200.utils.sampler import Sampler
    ps = load_problemset(a.robot, a.env)
    pp = dict(ps["planner_params"], num_samples=a.samples, time_spacing_X=a.timesteps)
    robot = Robot.from_tables(a.robot, a.env)
    sdf = mesh_to_sdf(scene_mesh_path(a.env), 0.01, PADDING)
    queries = [ps["queries"][i % len(ps["queries"])] for i in range(a.problems)]
    q = np.stack([np.stack(p) for p in queries])
    X, _, _ = init_trainset(pp["time_spacing_X"], pp["time_spacing_Xnew"], robot.dof, robot.dof, q[0, 0], q[0, 1], scale=1)
    model = VGPMP.initialize(sdf=sdf, robot=robot, sampler=Sampler(None, robot), query_states=q,
                             scene_offset=ps["scene_offset"], seed=7, **pp)
    disable_param_opt(model, default_trainable_params())
    eng = model._eng
    for kv in a.option:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    Xd = eng.dev(X)
    for _ in range(2):
        model.train_step(Xd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        model.train_step(Xd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    stage_ms = (C.c_double * _cabi.NUM_STAGES)()
    stage_n = (C.c_int64 * _cabi.NUM_STAGES)()
    eng.lib.vgpmp_profile_enable(eng.h, 1)
    for _ in range(a.steps):
        model.train_step(Xd)
    eng.lib.vgpmp_profile_collect(eng.h, stage_ms, stage_n)
    eng.lib.vgpmp_profile_enable(eng.h, 0)
    stages = {eng.lib.vgpmp_stage_name(i).decode(): stage_ms[i] / max(1, stage_n[i]) for i in range(_cabi.NUM_STAGES)}
    print(json.dumps({"robot": a.robot, "problems": a.problems, "S": a.samples, "N": X.shape[0], "M": model.num_inducing,
                      "ms_per_step": ms, "stage_ms": stages, "options": a.option,
                      "problem_it_per_s": a.problems / (ms * 1e-3),
                      "sdf_evals_per_s": a.problems * a.samples * X.shape[0] * robot.num_spheres / (ms * 1e-3)}))


if __name__ == "__main__":
    main()
