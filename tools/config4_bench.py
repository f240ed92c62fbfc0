#!/usr/bin/env python
"""BASELINE config 4: UR10 / bookshelves, ONE problem, S pathwise samples sharded over the ranks, one NCCL all-reduce of
the packed gradient per step.  Run under torchrun (or plain python for one GPU); prints one JSON line on rank 0.

  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config4_bench.py --samples 65536
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from vgpmp_b200.models import VGPMP
    from vgpmp_b200.utils.gen_sdf import PADDING, mesh_to_sdf, scene_mesh_path
    from vgpmp_b200.utils.miscellaneous import default_trainable_params, disable_param_opt, init_trainset, load_problemset
    from vgpmp_b200.utils.robot import Robot
    from vgpmp_b200.utils.sampler import Sampler
    ps = load_problemset("ur10", "bookshelves")
    pp = dict(ps["planner_params"], num_samples=a.samples)
    robot = Robot.from_tables("ur10", "bookshelves")
    sdf = mesh_to_sdf(scene_mesh_path("bookshelves"), 0.01, PADDING)
    start, goal = ps["queries"][0]
    X, y, _ = init_trainset(pp["time_spacing_X"], pp["time_spacing_Xnew"], robot.dof, robot.dof, start, goal, scale=1)
    model = VGPMP.initialize(sdf=sdf, robot=robot, sampler=Sampler(None, robot), query_states=y,
                             scene_offset=ps["scene_offset"], seed=1234 + 4, **pp)
    disable_param_opt(model, default_trainable_params())
    model.enable_sample_sharding(rank, world)
    Xd = model._eng.dev(X)
    for _ in range(a.warmup):
        model.train_step(Xd)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = model.train_step(Xd)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        q = [torch.zeros_like(model._q_mu) for _ in range(world)]
        dist.all_gather(q, model._q_mu)
        assert all(torch.equal(t, q[0]) for t in q), "ranks diverged"
    if rank == 0:
        P, N = robot.num_spheres, X.shape[0]
        print(json.dumps({"config": "ur10/bookshelves, 1 problem, sample-sharded", "n_gpus": world, "samples_total": a.samples,
                          "samples_per_gpu": model.num_samples, "ms_per_step": float(ms),
                          "elbo_iterations_per_s": 1000.0 / float(ms),
                          "sdf_evals_per_s": a.samples * N * P / (float(ms) * 1e-3), "loss": float(loss),
                          "allreduce_scalars": int(model._shard["flat"].numel())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
