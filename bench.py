#!/usr/bin/env python
"""bench.py -- ELBO iterations/s (and sphere-SDF evals/s) of the vgpmp hot path on N B200s.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps K ...  # the reference arithmetic on the host cores (float64 port)
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, problems sharded, no collective

A "step" is one pass of the hot path over one batch of synthetic problems: draw the step's randomness, ELBO forward,
reverse pass to (_q_mu, _q_sqrt, lengthscales, variances), Adam update.  Workload at any N (weak scaling): BASELINE.json
configs[1] per GPU -- Franka Panda, bookshelves planner_params (S=7, N=70, M=24, B=1024), all 55 start/goal pairs x
total_runs=5 (benchmarking.py:70) = 275 independent problems in one batch; the SDF is regenerated on the GPU from the
reference's bookshelves mesh (its .sdf grids are missing blobs).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "elbo_problem_iterations_per_s"
UNIT = "problem-iterations/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--runs", type=int, default=5, help="total_runs of benchmarking.py:70 (copies of the 55 pairs)")
    ap.add_argument("--sdf", default="bookshelves_mesh", choices=["bookshelves_mesh", "synthetic"],
                    help="bookshelves_mesh: signed distance to the reference's bookshelves_center.obj, produced on the GPU "
                         "(delta 1 cm, padding 20 as utils/gen_sdf.py:9); synthetic: analytic union of boxes")
    ap.add_argument("--sdf-dim", type=int, default=256)
    ap.add_argument("--cpu-problems", type=int, default=16, help="problems per step of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=INT",
                    help="experiments only: vgpmp_set_option(NAME, INT) before timing (e.g. mma_sampler=0)")
    ap.add_argument("--streams", type=int, default=2,
                    help="sub-batches of the problem batch, each on its own CUDA stream (StreamedVGPMP); 1 = one VGPMP model")
    ap.add_argument("--bases", type=int, default=0, help="experiments only: number of random Fourier bases (default 1024)")
    ap.add_argument("--problems", type=int, default=0, help="experiments only: truncate / cycle the batch to this many problems")
    return ap.parse_args()


def workload(runs):
    from vgpmp_b200.utils.miscellaneous import load_problemset
    ps = load_problemset("franka", "bookshelves")
    queries = [q for _ in range(runs) for q in ps["queries"]]
    return ps, queries


def bench_sdf(kind, dim):
    if kind == "bookshelves_mesh":
        from vgpmp_b200.utils.gen_sdf import PADDING, mesh_to_sdf, scene_mesh_path
        return mesh_to_sdf(scene_mesh_path("bookshelves"), 0.01, PADDING), "bookshelves_center.obj -> GPU SDF, delta=0.01, padding=20"
    from vgpmp_b200.utils.sdf_utils import synthetic_shelf_sdf
    # scene frame: robot base sits at -scene_offset = (-0.62, 0.15, -0.834); 2.56 m cube around the reach box
    return (synthetic_shelf_sdf(shape=(dim, dim, dim), delta=2.56 / dim, origin=(-1.6, -1.0, -1.8), seed=0, n_boxes=16),
            f"synthetic shelf {dim}^3")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        smax = float(rows[0][2]) if rows and rows[0][2].replace(".", "").isdigit() else None
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(rows)}


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_port_rate(problems, steps, warmup=1):
    """The float64 oracle (reference arithmetic restated; TF/GPflow are not installable) run like the reference runs:
    problems one after another, forward + autograd reverse + Keras-Adam, all host threads.  -> problem-iterations/s."""
    import torch
    from oracle import vgpmp_oracle as O
    from tests import helpers as H
    torch.set_num_threads(os.cpu_count() or 1)
    case = H.make_case("franka", "bookshelves", num_problems=problems, B=1024, seed=1234 + 2, perturb=False,
                       sdf=H.small_sdf(seed=0, shape=(128, 128, 128), delta=0.02, origin=(-1.6, -1.0, -1.8)))
    rng = np.random.default_rng(0)
    lr = case["pp"]["learning_rate"]
    state = [dict(p=dict(q_mu=case["q_mu"][b].copy(), q_sqrt=case["q_sqrt"][b].copy(),
                         raw_ls=O.softplus_inv(case["ls"][b]), raw_var=O.softplus_inv(case["var"][b])),
                  adam=O.AdamState()) for b in range(problems)]

    def one_step():
        for b, prob in enumerate(case["oracle"]):
            pr = state[b]["p"]
            draws = O.make_draws(rng, case["D"], case["S"], case["B"], case["M"] + 2)   # fresh draws every iteration
            ref = O.elbo_and_grads(prob, pr["q_mu"], pr["q_sqrt"], O.softplus(pr["raw_ls"]), O.softplus(pr["raw_var"]), draws)
            sig_ls, sig_var = 1 / (1 + np.exp(-pr["raw_ls"])), 1 / (1 + np.exp(-pr["raw_var"]))
            O.adam_step(pr, dict(q_mu=-ref["d_q_mu"], q_sqrt=-ref["d_q_sqrt"], raw_ls=-ref["d_lengthscales"] * sig_ls,
                                 raw_var=-ref["d_variances"] * sig_var), state[b]["adam"], lr)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(warmup):
            one_step()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_step()
        dt = time.perf_counter() - t0
    return problems * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        import gpflow, gpflow_sampling, tensorflow  # noqa: F401,E401
        note = "tensorflow/gpflow importable but the reference driver also needs pybullet; using the float64 port"
    except Exception:
        note = "tensorflow/gpflow/gpflow_sampling/pybullet not installable here: float64 port of the reference arithmetic"
    steps = max(1, min(args.steps, 60))
    rate, dt = cpu_port_rate(args.cpu_problems, steps, warmup=min(args.warmup, 2))
    cores = os.cpu_count() or 1
    sample = f"{args.cpu_problems} Franka/bookshelves problems per step x {steps} steps ({dt:.1f} s), solved one after another"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 2), "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "franka/bookshelves planner_params S=7 N=70 M=24 B=1024, bounded sample of the 275-problem batch",
                   "sdf": "synthetic shelf 128^3 float64", "note": note},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_b200(args):
    import torch
    import torch.distributed as dist
    from vgpmp_b200 import _cabi
    from vgpmp_b200.models import StreamedVGPMP, VGPMP
    from vgpmp_b200.utils.miscellaneous import default_trainable_params, disable_param_opt, init_trainset
    from vgpmp_b200.utils.robot import Robot
    from vgpmp_b200.utils.sampler import Sampler
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: vgpmp_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    ps, queries = workload(args.runs)
    if args.problems:
        queries = [queries[i % len(queries)] for i in range(args.problems)]
    pp = dict(ps["planner_params"])
    sdf, sdf_desc = bench_sdf(args.sdf, args.sdf_dim)
    robot = Robot.from_tables("franka", "bookshelves")
    sampler = Sampler(None, robot)
    q = np.stack([np.stack(pair) for pair in queries])
    X, _, _ = init_trainset(pp["time_spacing_X"], pp["time_spacing_Xnew"], robot.dof, robot.dof, q[0, 0], q[0, 1], scale=1)
    extra = {"num_bases": args.bases} if args.bases else {}
    init_kw = dict(sdf=sdf, robot=robot, sampler=sampler, scene_offset=ps["scene_offset"], seed=1234 + 2 + 1000 * rank, **pp,
                   **extra)

    def configure(m):
        disable_param_opt(m, default_trainable_params())
        for kv in args.option:
            k, v = kv.split("=")
            m._eng.set_option(k, int(v))
        return m

    # `model`: the whole batch in one VGPMP (used for the per-kernel stage profile: every launch covers all problems).
    # `runner`: what is timed - the same batch as `--streams` sub-batches on their own CUDA streams (StreamedVGPMP), whose
    # latency-bound kernels fill the gaps of each other's FP64-bound sampler; identical results (global Philox keys).
    model = configure(VGPMP.initialize(query_states=q, **init_kw))
    eng = model._eng
    if args.streams > 1:
        runner = StreamedVGPMP.initialize(query_states=q, num_streams=args.streams, **init_kw)
        for m in runner.models:
            configure(m)
        count_launches = lambda: runner.launch_count
    else:
        runner = model
        count_launches = lambda: eng.launch_count
    Bp, S, N, M, B, D, P = model.num_problems, model.num_samples, X.shape[0], model.num_inducing, model.num_bases, robot.dof, robot.num_spheres
    Xd = eng.dev(X)

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        runner.train_step(Xd)
    torch.cuda.synchronize()

    # ------------------------------------------------------------------ timed region (inputs resident in HBM)
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = count_launches()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        runner.train_step(Xd)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = count_launches() - launches0

    # ------------------------------------------------------------------ stage profile: untimed extra pass, one launch per
    # kernel over the whole batch, CUDA events recorded by the library around every stage (they cost ~5 %: not in `value`)
    stage_ms = (C.c_double * _cabi.NUM_STAGES)()
    stage_n = (C.c_int64 * _cabi.NUM_STAGES)()
    for _ in range(3):
        model.train_step(Xd)
    torch.cuda.synchronize()
    eng.lib.vgpmp_profile_collect(eng.h, stage_ms, stage_n)
    eng.lib.vgpmp_profile_enable(eng.h, 1)
    for _ in range(min(args.steps, 50)):
        model.train_step(Xd)
    eng.lib.vgpmp_profile_collect(eng.h, stage_ms, stage_n)
    eng.lib.vgpmp_profile_enable(eng.h, 0)

    # ------------------------------------------------------------------ e2e: host buffers through the public step
    Xh = torch.from_numpy(X.copy()).pin_memory()
    for _ in range(3):
        runner.train_step_host(Xh)
    barrier()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = None
    for _ in range(args.steps):
        last = runner.train_step_host(Xh)      # H2D X, device RNG, fwd+bwd+Adam, D2H loss, stream sync(s)
    f1.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    assert np.all(np.isfinite(last.numpy()))
    clk = clocks.stop(t0, t2)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        per_step = ms / args.steps
        value = world * Bp * args.steps / (ms / 1000.0)
        evals_per_step = Bp * S * N * P
        peak, peak_src = measured_peaks()
        stages = {}
        tot = sum(stage_ms) or 1.0
        for i in range(_cabi.NUM_STAGES):
            name = eng.lib.vgpmp_stage_name(i).decode()
            if stage_n[i]:
                stages[name] = {"ms_per_launch": stage_ms[i] / stage_n[i], "share": stage_ms[i] / tot, "launches": int(stage_n[i])}
        dominant = max(stages, key=lambda k: stages[k]["share"])
        sdf_ms = stages["loglik_fwd_bwd"]["ms_per_launch"]
        sdf_bytes = evals_per_step * 32                         # one 32-byte {value, gradient} record per sphere-SDF eval
        achieved = sdf_bytes / (sdf_ms * 1e-3) / 1e9
        roofline = {"kernel": "loglik_kernel<7,true,4> (FK + one 256-bit {value,gradient} record load per sphere + hinge + reverse pass)", "bound": "hbm",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": sdf_bytes,
                    "ms_per_launch": sdf_ms, "share_of_step": stages["loglik_fwd_bwd"]["share"],
                    "dominant_stage": dominant}
        tr = ROOT / "profiles" / "roofline_traffic.json"
        if tr.exists():
            t = json.loads(tr.read_text()).get("loglik_fwd_bwd")
            if t:
                roofline["traffic"] = t["dram_bytes_per_launch"]
                roofline["traffic_source"] = t["source"]
                if "gather_probe" in t:     # measured ceiling of random 32-byte gathers (context for `frac`, not the peak)
                    roofline["gather_probe"] = t["gather_probe"]
                    roofline["frac_of_gather_probe"] = achieved / t["gather_probe"]["l2_resident_grid_GBps"]
        # dominant stage by time = the sampler: FP64-pipe bound (B200 runs DMMA on the FP64 pipe at the DFMA rate, see
        # profiles/r1_v9_dmma_probe.txt).  Algorithmic flops = the two contractions (f0 and d f0/d lengthscale: S*A*B FMAs
        # each, 2 flops per FMA) plus 6 FP64 operations per generated feature pair; peak = FP64 FMA rate measured now
        # by vgpmp_probe_fp64_tflops.  (Rounds up to r1_v8 multiplied the contraction by a stray factor 2; their
        # recorded "0.57" fractions are 0.33 on this count.)  The stage time includes gp_prepare (~0.1 ms).
        A = N + M + 2
        pw_flops = Bp * D * (2 * 2 * S * A * B + 6 * A * B)
        pw = stages.get("pathwise_sample")
        out_dom = None
        if pw:
            peak64 = float(eng.lib.vgpmp_probe_fp64_tflops(local))
            ach = pw_flops / (pw["ms_per_launch"] * 1e-3) / 1e12
            out_dom = {"kernel": "pathwise_rr_kernel<9> + gp_prepare_update_kernel (register-resident rotation chains feeding DMMA m8n8k4, then GP preparation + pathwise update)",
                       "bound": "fp64", "achieved": ach, "peak": peak64, "unit": "TFLOP/s", "frac": ach / peak64 if peak64 > 0 else None,
                       "peak_source": "measured now: vgpmp_probe_fp64_tflops (DFMA chains, CUDA events)", "traffic": None,
                       "algorithmic_flops_per_launch": pw_flops, "ms_per_launch": pw["ms_per_launch"], "share_of_step": pw["share"]}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"franka/bookshelves: 55 start-goal pairs x total_runs={args.runs} = {Bp} problems per GPU per step",
                       "S": S, "N": N, "M": M, "B": B, "dof": D, "spheres": P,
                       "sdf": f"{sdf_desc}; grid {sdf.data.shape} float64, {sdf.data.nbytes * 4 / 2**20:.0f} MiB of "
                              "{value,gradient} records in HBM (the reference's own .sdf grids are missing blobs)",
                       "rng": "device Philox4x32-10, fresh draws every step (lazy: generated inside the sampler kernel)",
                       "streams": args.streams,
                       "l2": "no explicit flush: every step streams its whole working set, which exceeds the 126 MB L2 (86 MiB of SDF records + ~110 MB of GP factors, prior draws, samples, gradients and Adam state; omega/tau/w are generated in-kernel and never stored)"},
            "sdf_evals_per_s": world * evals_per_step * args.steps / (ms / 1000.0),
            "e2e": {"value": world * Bp * args.steps / (ms_e2e / 1000.0), "unit": UNIT,
                    "h2d_bytes_per_step": int(X.nbytes) * max(args.streams, 1), "d2h_bytes_per_step": int(Bp * 8),
                    "ms_per_step": ms_e2e / args.steps, "api": ("StreamedVGPMP.train_step_host -> vgpmp_train_step_host_begin/_end per sub-batch" if args.streams > 1 else "VGPMP.train_step_host -> vgpmp_train_step_host")},
            "gpu_launches": int(launches), "stages": stages, "roofline": roofline, "roofline_dominant_stage": out_dom,
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            steps = 40
            rate, dt = cpu_port_rate(args.cpu_problems, steps)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                   "sample": f"{args.cpu_problems} of the {Bp} problems x {steps} steps ({dt:.1f} s), float64 "
                                             "torch-CPU restatement of the reference (TF/GPflow not installable)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
