#!/usr/bin/env python
"""SDF-gather stage at scale (BASELINE config 5 class): the fused likelihood kernel (FK + one {value,gradient} record per
sphere + hinge + reverse pass) on a large grid whose records exceed L2 by far, with configurations spread over the whole
reach volume so that record loads miss L2.  Prints one JSON line: sphere-SDF evals/s, algorithmic GB/s (32 B per eval)
and its fraction of the measured HBM bandwidth.

    python tools/sdf_stage_bench.py [--configs 16777216] [--delta 0.004] [--padding 90] [--iters 10]
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", type=int, default=1 << 24)      # 8192 problems x 64 timesteps x 256 samples = 2^27
    ap.add_argument("--delta", type=float, default=0.004)
    ap.add_argument("--padding", type=int, default=90)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--spread", type=float, default=1.5, help="std of the latent joint values (sigmoid-squashed)")
    a = ap.parse_args()
    from vgpmp_b200.engine import Engine
    from vgpmp_b200.utils.gen_sdf import mesh_to_sdf, scene_mesh_path
    from vgpmp_b200.utils.miscellaneous import load_problemset
    from vgpmp_b200.utils.robot import Robot
    from vgpmp_b200.utils.sampler import Sampler

    ps = load_problemset("franka", "bookshelves")
    pp = ps["planner_params"]
    sdf = mesh_to_sdf(scene_mesh_path("bookshelves"), a.delta, a.padding)
    robot = Robot.from_tables("franka", "bookshelves")
    eng = Engine(Sampler(None, robot).constants(), sdf.data, sdf.origin, sdf.delta, sigma_obs=pp["sigma_obs"],
                 epsilon=pp["epsilon"], alpha=pp["alpha"], scene_offset=ps["scene_offset"])
    n, D, P = a.configs, robot.dof, robot.num_spheres
    g = torch.Generator(device="cuda").manual_seed(0)
    f = a.spread * torch.randn(n, D, dtype=torch.float64, device="cuda", generator=g)
    for _ in range(3):
        logp, df = eng.loglik(f, squash=True, upstream=1.0, need_grad=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        logp, df = eng.loglik(f, squash=True, upstream=1.0, need_grad=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    evals = n * P
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    gbs = evals * 32 / (ms * 1e-3) / 1e9
    io_gbs = (evals * 32 + n * (2 * D + 1) * 8) / (ms * 1e-3) / 1e9          # + f in, df and logp out
    in_hinge = float((logp < 0).double().mean())
    print(json.dumps({"stage": "loglik_fwd_bwd at scale", "configs": n, "spheres": P, "sdf_grid": list(sdf.data.shape),
                      "records_GiB": sdf.data.nbytes * 4 / 2**30, "ms_per_launch": ms, "sdf_evals_per_s": evals / (ms * 1e-3),
                      "algorithmic_GBps_records_only": gbs, "algorithmic_GBps_with_io": io_gbs,
                      "frac_of_measured_hbm": gbs / peaks["hbm_gbs"], "frac_with_io": io_gbs / peaks["hbm_gbs"],
                      "hbm_peak_GBps": peaks["hbm_gbs"], "configs_touching_hinge": in_hinge}))


if __name__ == "__main__":
    main()
