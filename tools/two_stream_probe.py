#!/usr/bin/env python
"""Experiment: the 275-problem batch as two half-batches on two CUDA streams (latency-bound kernels of one half can fill
the gaps of the FP64-bound sampler of the other).  GPU only.   python tools/two_stream_probe.py"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from vgpmp_b200.models import VGPMP
from vgpmp_b200.utils.miscellaneous import disable_param_opt, init_trainset, load_problemset, default_trainable_params
from vgpmp_b200.utils.robot import Robot
from vgpmp_b200.utils.sampler import Sampler
from vgpmp_b200.utils.sdf_utils import synthetic_shelf_sdf


def build(queries, ps, sdf, robot, sampler, seed):
    pp = ps["planner_params"]
    q = np.stack([np.stack(pair) for pair in queries])
    m = VGPMP.initialize(sdf=sdf, robot=robot, sampler=sampler, query_states=q, scene_offset=ps["scene_offset"], seed=seed, **pp)
    disable_param_opt(m, default_trainable_params())
    return m


def main():
    ps = load_problemset("franka", "bookshelves")
    pp = ps["planner_params"]
    robot = Robot.from_tables("franka", "bookshelves")
    sampler = Sampler(None, robot)
    sdf = synthetic_shelf_sdf(shape=(128, 128, 128), delta=0.02, origin=(-1.6, -1.0, -1.8), seed=0, n_boxes=16)
    queries = [ps["queries"][i % len(ps["queries"])] for i in range(275)]
    X, _, _ = init_trainset(pp["time_spacing_X"], pp["time_spacing_Xnew"], robot.dof, robot.dof, queries[0][0], queries[0][1], scale=1)
    for nsplit in (1, 2, 3):
        parts = np.array_split(np.arange(275), nsplit)
        models = [build([queries[i] for i in idx], ps, sdf, robot, sampler, 1 + k) for k, idx in enumerate(parts)]
        streams = [torch.cuda.Stream() for _ in models]
        Xd = [m._eng.dev(X) for m in models]
        def step():
            for m, s, x in zip(models, streams, Xd):
                with torch.cuda.stream(s):
                    m.train_step(x)
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 200
        print(f"{nsplit} stream(s): {dt * 1e3:.4f} ms per 275-problem step = {275 / dt:,.0f} problem-it/s")


if __name__ == "__main__":
    main()
