#!/usr/bin/env python
"""bench.py -- ELBO iterations/s (and sphere-SDF evals/s) of the vgpmp hot path on N B200s.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps K ...  # the reference arithmetic on the host cores (float64 port)
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU

A "step" is one pass of the hot path over one batch of synthetic problems: draw the step's randomness, ELBO forward,
reverse pass to (_q_mu, _q_sqrt, lengthscales, variances), Adam update.

Headline workload (`--config 2`, the default; BASELINE.json configs[1]): Franka Panda, bookshelves planner_params (S=7, N=70,
M=24, B=1024), all 55 start/goal pairs x total_runs=5 (benchmarking.py:70) = 275 independent problems per GPU (weak
scaling, no collective); the SDF is regenerated on the GPU from the reference's bookshelves mesh (its .sdf grids are
missing blobs).  The same JSON line carries a `configs` block with BASELINE configs 3, 4 and 5 at this GPU count (strong
scaling: the work is divided over the ranks), each with its own stage times and SDF-stage roofline:
  3  Kuka iiwa7 / industrial, 1024 synthetic start/goal problems sharded over the ranks (no collective)
  4  UR10 / bookshelves, ONE problem, 65 536 pathwise samples sharded over the ranks + one NCCL all-reduce per step
  5  Franka, 512^3 SDF grid (4 GiB of records), 8192 problems x 64 timesteps x 256 samples sharded over the ranks
`--config K` makes K the headline instead.  One JSON line on stdout (rank 0).
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
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json config used as the headline")
    ap.add_argument("--extra-configs", default=None,
                    help="comma list of further configs measured into the `configs` block (default: 3,4,5 with --config 2)")
    ap.add_argument("--extra-steps", type=int, default=10, help="timed steps of each extra config")
    ap.add_argument("--runs", type=int, default=5, help="total_runs of benchmarking.py:70 (copies of the 55 pairs)")
    ap.add_argument("--cpu-problems", type=int, default=16, help="problems per step of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=INT",
                    help="experiments only: vgpmp_set_option(NAME, INT) before timing (e.g. tc_sampler=0)")
    ap.add_argument("--streams", type=int, default=2,
                    help="config 2: sub-batches of the problem batch, each on its own CUDA stream (StreamedVGPMP)")
    ap.add_argument("--bases", type=int, default=0, help="experiments only: number of random Fourier bases (default 1024)")
    ap.add_argument("--problems", type=int, default=0, help="experiments only: override the number of problems")
    ap.add_argument("--samples", type=int, default=0, help="experiments only: override the number of samples")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------- workloads
CONFIGS = {
    2: dict(robot="franka", env="bookshelves", sdf="bookshelves", scaling="weak",
            what="franka/bookshelves: 55 start-goal pairs x total_runs=5 = 275 problems per GPU per step"),
    3: dict(robot="kuka", env="industrial", sdf="industrial", scaling="strong", problems=1024,
            what="kuka iiwa7/industrial planner_params, 1024 synthetic start/goal problems sharded over the GPUs"),
    4: dict(robot="ur10", env="bookshelves", sdf="bookshelves", scaling="strong", problems=1, samples=65536,
            what="ur10/bookshelves, ONE problem, 65536 pathwise samples sharded over the GPUs, one NCCL all-reduce of the "
                 "packed gradient per step"),
    5: dict(robot="franka", env="bookshelves", sdf="bookshelves512", scaling="strong", problems=8192, samples=256, timesteps=64,
            what="franka, 512^3 SDF grid, 8192 synthetic problems x 64 timesteps x 256 samples sharded over the GPUs"),
}

_SDF_CACHE = {}


def build_sdf(kind):
    """GPU producer (vgpmp_mesh_to_sdf) on the reference's scene meshes; cached per process."""
    if kind in _SDF_CACHE:
        return _SDF_CACHE[kind]
    from vgpmp_b200.utils.gen_sdf import PADDING, mesh_to_sdf, scene_mesh_path
    if kind == "bookshelves512":
        # 512^3 nodes at 5 mm around the Franka's reach volume (robot base at -scene_offset = (-0.62, 0.15, -0.834))
        sdf = mesh_to_sdf(scene_mesh_path("bookshelves"), 0.005, origin=(-1.9, -1.13, -1.73), shape=(512, 512, 512))
        desc = "bookshelves_center.obj -> GPU SDF, 512^3 nodes, delta=0.005"
    else:
        sdf = mesh_to_sdf(scene_mesh_path(kind), 0.01, PADDING)
        desc = f"{scene_mesh_path(kind).name} -> GPU SDF, delta=0.01, padding={PADDING}"
    _SDF_CACHE[kind] = (sdf, desc)
    return sdf, desc


def problem_queries(cfg_id, cfg, ps, robot, args):
    """[Bp,2,D] start/goal joint states of the WHOLE job (before sharding)."""
    real = [np.stack(p) for p in ps["queries"]]
    if cfg_id == 2:
        q = [real[i % len(real)] for i in range(args.runs * len(real))]
        if args.problems:
            q = [q[i % len(q)] for i in range(args.problems)]
        return np.stack(q)
    n = args.problems or cfg["problems"]
    if n == 1:
        return np.stack(real[:1])
    # synthetic problems: the table's pairs first, then uniform draws inside the middle 80 % of the joint range
    rng = np.random.default_rng(1234 + cfg_id)
    lo, hi = robot.limits_lo, robot.limits_hi
    mid, half = 0.5 * (lo + hi), 0.4 * (hi - lo)
    q = list(real[:n])
    while len(q) < n:
        q.append(mid + half * rng.uniform(-1.0, 1.0, size=(2, robot.dof)))
    return np.stack(q)


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML is polled in-process every
    2 ms (the default timed region is ~15 ms: `nvidia-smi -lms 50` would see it once); without the NVML binding the
    nvidia-smi loop of the recipe is used."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx, self.nvml, self.smax, self.stop_flag = [], None, gpu_index, None, None, False

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.idx < len(ids) and ids[self.idx].isdigit():
                return int(ids[self.idx])
        return self.idx

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self._physical_index()), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.dev, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.dev))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev))
                self.rows.append((time.perf_counter(), mhz, mask))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.nvml is not None:
            self.stop_flag = True
            rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
            reasons = sorted(k for k, bit in self.BITS.items() if any(r[2] & bit for r in rows))
            return {"sm_mhz": statistics.median([r[1] for r in rows]) if rows else None, "sm_max_mhz": self.smax,
                    "reasons": reasons, "samples": len(rows), "source": "nvml, 2 ms period"}
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
                "samples": len(rows), "source": "nvidia-smi -lms 50"}


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------- CPU arm
def cpu_bookshelves_sdf():
    """The SAME grid as the GPU arm's (bookshelves_center.obj, delta 1 cm, padding 20), produced on the host by the oracle's
    vectorised float64 mesh->SDF (no repo .so is loaded); cached in the temp dir between the two legs that use it."""
    import tempfile
    from oracle import vgpmp_oracle as O
    from vgpmp_b200.utils.gen_sdf import PADDING, grid_geometry, load_obj_convex_pieces, scene_mesh_path
    from vgpmp_b200.utils.sdf_utils import SignedDistanceField
    tri, plane, piece_end = load_obj_convex_pieces(scene_mesh_path("bookshelves"))
    origin, shape = grid_geometry(tri, 0.01, PADDING)
    cache = Path(tempfile.gettempdir()) / f"vgpmp_b200_cpu_sdf_bookshelves_{shape[0]}x{shape[1]}x{shape[2]}.npy"
    if cache.exists():
        data = np.load(cache)
    else:
        data = O.mesh_sdf_vectorised(tri, plane, piece_end, origin, 0.01, shape)
        try:
            np.save(cache, data)
        except OSError:
            pass
    return SignedDistanceField(data, origin, 0.01)


def cpu_port_rate(problems, steps, warmup=3):
    """The float64 oracle (reference arithmetic restated; TF/GPflow are not installable) run like the reference runs:
    problems one after another, forward + autograd reverse + Keras-Adam, all host threads, on the first `problems` problems
    of the config-2 batch and the same bookshelves grid as the GPU arm.  -> problem-iterations/s."""
    import torch
    from oracle import vgpmp_oracle as O
    from tests import helpers as H
    torch.set_num_threads(os.cpu_count() or 1)
    case = H.make_case("franka", "bookshelves", num_problems=problems, B=1024, seed=1234 + 2, perturb=False,
                       sdf=cpu_bookshelves_sdf())
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
    warm = max(3, min(args.warmup, 5))
    rate, dt = cpu_port_rate(args.cpu_problems, steps, warmup=warm)
    cores = os.cpu_count() or 1
    sample = (f"the first {args.cpu_problems} of the 275 Franka/bookshelves problems per step x {steps} steps ({dt:.1f} s), "
              "solved one after another")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the B200 arm's config 2 (same workload string, same shape keys); what was actually timed is in cpu_baseline.sample
        "config": {"workload": CONFIGS[2]["what"], "baseline_config": 2, "problems_total": 275 * max(args.gpus, 1),
                   "problems_this_gpu": 275, "S": 7, "N": 70, "M": 24, "B": 1024, "dof": 7, "spheres": 37,
                   "sdf": "bookshelves_center.obj -> float64 SDF on the host (oracle mesh producer), delta=0.01, padding=20: "
                          "the grid of the GPU arm", "sample": sample, "note": note},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ----------------------------------------------------------------------------------------------------------- GPU arm
class Job:
    """One BASELINE config on this rank: model(s), inputs, and how much of the whole job this rank holds."""

    def __init__(self, cfg_id, args, rank, world, headline):
        import torch
        from vgpmp_b200.models import StreamedVGPMP, VGPMP
        from vgpmp_b200.utils.miscellaneous import (default_trainable_params, disable_param_opt, init_trainset,
                                                    load_problemset)
        from vgpmp_b200.utils.robot import Robot
        from vgpmp_b200.utils.sampler import Sampler
        from vgpmp_b200.utils.sharding import shard_range
        self.id, self.cfg, self.rank, self.world = cfg_id, CONFIGS[cfg_id], rank, world
        cfg = self.cfg
        ps = load_problemset(cfg["robot"], cfg["env"])
        pp = dict(ps["planner_params"])
        if cfg.get("samples") or args.samples:
            pp["num_samples"] = args.samples or cfg["samples"]
        if cfg.get("timesteps"):
            pp["time_spacing_X"] = cfg["timesteps"]
        if args.bases:
            pp["num_bases"] = args.bases
        self.sdf, self.sdf_desc = build_sdf(cfg["sdf"])
        robot = Robot.from_tables(cfg["robot"], cfg["env"])
        self.robot = robot
        q_all = problem_queries(cfg_id, cfg, ps, robot, args)
        self.total_problems = len(q_all) * (world if cfg["scaling"] == "weak" else 1)
        self.sample_sharded = cfg_id == 4
        if cfg["scaling"] == "strong" and not self.sample_sharded:
            lo, hi = shard_range(len(q_all), rank, world)
            q, self.problem_offset = q_all[lo:hi], lo
        else:
            q, self.problem_offset = q_all, (rank * len(q_all) if cfg["scaling"] == "weak" else 0)
        X, _, _ = init_trainset(pp["time_spacing_X"], pp["time_spacing_Xnew"], robot.dof, robot.dof, q[0, 0], q[0, 1], scale=1)
        self.X = X
        seed = 1234 + cfg_id
        kw = dict(sdf=self.sdf, robot=robot, sampler=Sampler(None, robot), scene_offset=ps["scene_offset"], seed=seed, **pp)

        def configure(m):
            disable_param_opt(m, default_trainable_params())
            for kv in args.option:
                k, v = kv.split("=")
                m._eng.set_option(k, int(v))
            return m

        # `model`: the rank's whole batch in one VGPMP (stage profile: every launch covers all its problems).
        # `runner`: what is timed.  Config 2 drives the batch as `--streams` sub-batches on their own CUDA streams
        # (StreamedVGPMP: their latency-bound kernels fill the gaps of each other's FP64-bound sampler; identical results,
        # the Philox keys use global problem indices).  All handles share one copy of the SDF records.
        self.model = configure(VGPMP.initialize(query_states=q if len(q) > 1 else q[0], **kw))
        self.model.problem_offset = self.problem_offset
        if self.sample_sharded:
            self.model.enable_sample_sharding(rank, world)
        self.streams = args.streams if (cfg_id == 2 and args.streams > 1 and len(q) > 1) else 1
        if self.streams > 1:
            self.runner = StreamedVGPMP.initialize(query_states=q, num_streams=self.streams, share_engine=self.model._eng, **kw)
            for m in self.runner.models:
                configure(m)
                m.problem_offset += self.problem_offset
            self.count_launches = lambda: self.runner.launch_count
        else:
            self.runner = self.model
            self.count_launches = lambda: self.model._eng.launch_count
        m = self.model
        self.Bp, self.S, self.N, self.M, self.B = m.num_problems, m.num_samples, X.shape[0], m.num_inducing, m.num_bases
        self.S_total = m._shard["total"] if m._shard is not None else m.num_samples
        self.D, self.P = robot.dof, robot.num_spheres
        self.Xd = m._eng.dev(X)
        self.torch = torch

    # whole-job units per step (all ranks)
    def problem_iterations(self):
        return self.total_problems

    def sdf_evals(self):
        return self.total_problems * self.S_total * self.N * self.P

    def step(self):
        return self.runner.train_step(self.Xd)

    def timed(self, steps, warmup, barrier):
        torch = self.torch
        for _ in range(max(warmup, 3)):
            self.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = self.count_launches()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            self.step()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        return e0.elapsed_time(e1), self.count_launches() - l0, t0, t1

    def stage_profile(self, steps):
        """Untimed extra pass of the one-model batch with the library's per-stage CUDA events on (they cost ~5 %)."""
        import ctypes as C
        from vgpmp_b200 import _cabi
        eng = self.model._eng
        stage_ms = (C.c_double * _cabi.NUM_STAGES)()
        stage_n = (C.c_int64 * _cabi.NUM_STAGES)()
        for _ in range(2):
            self.model.train_step(self.Xd)
        self.torch.cuda.synchronize()
        eng.lib.vgpmp_profile_collect(eng.h, stage_ms, stage_n)
        eng.lib.vgpmp_profile_enable(eng.h, 1)
        for _ in range(steps):
            self.model.train_step(self.Xd)
        eng.lib.vgpmp_profile_collect(eng.h, stage_ms, stage_n)
        eng.lib.vgpmp_profile_enable(eng.h, 0)
        stages, tot = {}, sum(stage_ms) or 1.0
        for i in range(_cabi.NUM_STAGES):
            if stage_n[i]:
                stages[eng.lib.vgpmp_stage_name(i).decode()] = {
                    "ms_per_launch": stage_ms[i] / stage_n[i], "share": stage_ms[i] / tot, "launches": int(stage_n[i]),
                    "ms_per_step": stage_ms[i] / steps}
        return stages

    def sdf_roofline(self, stages, traffic_key=None):
        """The SDF / likelihood kernel (the stage north_star names): algorithmic bytes = one 32-byte {value, gradient}
        record per sphere-SDF evaluation of THIS rank, over the kernel's CUDA-event time."""
        peak, peak_src = measured_peaks()
        st = stages["loglik_fwd_bwd"]
        evals = self.Bp * self.S * self.N * self.P
        nbytes = evals * 32
        ach = nbytes / (st["ms_per_step"] * 1e-3) / 1e9
        rec_mib = self.sdf.data.nbytes * 4 / 2**20
        out = {"kernel": f"loglik_bwd_kernel<{self.D},3,4> (FK + one 256-bit {{value,gradient}} record load per sphere + hinge + reverse pass)",
               "bound": "hbm" if rec_mib > 126 else "l2", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
               "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": nbytes,
               "ms_per_launch": st["ms_per_step"], "share_of_step": st["share"], "records_MiB": rec_mib,
               "sphere_evals_per_s_this_gpu": evals / (st["ms_per_step"] * 1e-3)}
        tr = ROOT / "profiles" / "roofline_traffic.json"
        if tr.exists() and traffic_key:
            t = json.loads(tr.read_text()).get(traffic_key)
            if t:       # DRAM bytes of the ncu capture, for the problems this stage time covers
                out["traffic"] = t["dram_bytes_per_launch"] * self.Bp / t["problems_in_launch"] if "dram_bytes_per_launch" in t \
                    else t["dram_bytes_per_problem"] * self.Bp
                out["traffic_source"] = t["source"]
                if "gather_probe" in t:     # measured ceiling of 32-byte gathers from an L2-resident grid (context for `frac`)
                    out["gather_probe"] = t["gather_probe"]
                    # the probe gathers at RANDOM; neighbouring timesteps of a trajectory share sectors, so an HBM-resident
                    # grid can exceed the random-gather figure
                    key = "hbm_resident_grid_GBps" if rec_mib > 126 else "l2_resident_grid_GBps"
                    out["frac_of_gather_probe"] = ach / t["gather_probe"][key]
                    out["gather_probe_used"] = key
        return out


def measure_allreduce(job, world, iters=50):
    """Config 4's collective alone: one all-reduce of the packed (gradient || ELBO) buffer, CUDA events."""
    import torch
    import torch.distributed as dist
    flat = job.model._shard["flat"].clone()
    if world > 1:
        for _ in range(5):
            dist.all_reduce(flat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if world > 1:
        for _ in range(iters):
            dist.all_reduce(flat)
    e1.record()
    torch.cuda.synchronize()
    return {"payload_bytes": int(flat.numel() * 8), "us_per_allreduce": (1000.0 * e0.elapsed_time(e1) / iters) if world > 1 else 0.0}


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: vgpmp_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()

    def reduce_max(vals):
        if world == 1:
            return vals
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    head = Job(args.config, args, rank, world, True)
    eng = head.model._eng

    # ------------------------------------------------------------------ timed region (inputs resident in HBM)
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.15)
    ms, launches, t0, t1 = head.timed(args.steps, args.warmup, barrier)
    clk = clocks.stop(t0, t1)      # the poller stops here: its wake-ups would otherwise jitter the host-driven e2e loop below

    # ------------------------------------------------------------------ stage profile (untimed extra pass)
    stages = head.stage_profile(min(args.steps, 50))

    # ------------------------------------------------------------------ e2e: host buffers through the public step
    ms_e2e, e2e = None, None
    if True:      # every configuration has a host-buffer step (one C call + CUDA graph when the batch is one unsharded chunk)
        Xh = torch.from_numpy(head.X.copy()).pin_memory()
        for _ in range(3):
            head.runner.train_step_host(Xh)
        barrier()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        last = None
        for _ in range(args.steps):
            last = head.runner.train_step_host(Xh)      # H2D X, device RNG, fwd+bwd+Adam, D2H loss, stream sync(s)
        f1.record()
        torch.cuda.synchronize()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
        assert np.all(np.isfinite(last.numpy()))
    t2 = time.perf_counter()
    ms, ms_e2e_r = reduce_max([ms, ms_e2e if ms_e2e is not None else 0.0])

    # ------------------------------------------------------------------ the other BASELINE configs at this GPU count
    extra_ids = ([3, 4, 5] if args.config == 2 else []) if args.extra_configs is None else \
        [int(v) for v in args.extra_configs.split(",") if v]
    extras = {}

    def summarize(job, ms_total, steps, stg):
        per = ms_total / steps
        out = {"workload": job.cfg["what"], "scaling": job.cfg["scaling"], "n_gpus": world, "ms_per_step": per,
               "problem_iterations_per_s": job.problem_iterations() / (per * 1e-3),
               "sdf_evals_per_s": job.sdf_evals() / (per * 1e-3),
               "problems_total": job.total_problems, "problems_this_gpu": job.Bp, "samples_total": job.S_total,
               "samples_this_gpu": job.S, "N": job.N, "M": job.M, "B": job.B, "dof": job.D, "spheres": job.P,
               "sdf": f"{job.sdf_desc}; grid {job.sdf.data.shape}, {job.sdf.data.nbytes * 4 / 2**20:.0f} MiB of records",
               "chunks_per_step": len(job.model._plan(job.Xd, False)[0]),
               "stages_ms_per_step": {k: v["ms_per_step"] for k, v in stg.items()},
               "sdf_stage_roofline": job.sdf_roofline(stg, f"config{job.id}_loglik")}
        return out

    for cid in extra_ids:
        if cid == args.config:
            continue
        job = None
        try:
            job = Job(cid, args, rank, world, False)
            ems, _, _, _ = job.timed(args.extra_steps, 3, barrier)
            (ems,) = reduce_max([ems])
            stg = job.stage_profile(min(args.extra_steps, 5))
            extras[str(cid)] = summarize(job, ems, args.extra_steps, stg)
            if cid == 4:
                ar = measure_allreduce(job, world)
                ar["share_of_step"] = ar["us_per_allreduce"] * 1e-3 / extras[str(cid)]["ms_per_step"]
                extras[str(cid)]["allreduce"] = ar
        except Exception as exc:     # an extra config must not take the headline down with it
            extras[str(cid)] = {"error": f"{type(exc).__name__}: {exc}"}
        del job
        torch.cuda.empty_cache()

    if rank == 0:
        per_step = ms / args.steps
        value = head.problem_iterations() * args.steps / (ms / 1000.0)
        sdf_roof = head.sdf_roofline(stages, "loglik_fwd_bwd" if args.config == 2 else f"config{args.config}_loglik")
        dominant = max(stages, key=lambda k: stages[k]["share"])
        # `roofline` = the kernel that dominates the step.  At config 2 that is the register-resident FP64 sampler: B200 runs
        # DMMA on the FP64 pipe at the DFMA rate (profiles/r1_v9_dmma_probe.txt), so its roofline is the FP64 FMA rate,
        # measured now by vgpmp_probe_fp64_tflops (MEASURED_PEAKS.json has no FP64 entry).  Algorithmic flops = the two
        # contractions (f0 and d f0 / d lengthscale: S*A*B FMAs each) plus 6 FP64 operations per generated feature pair.
        # With the tensor-core sampler (S >= 64) the denominator is the measured bf16 peak / 2 (TF32 dense rate) and the
        # flops are the 3 passes the split issues.
        A = head.N + head.M + 2
        pw = stages.get("pathwise_sample")
        roofline = sdf_roof
        if dominant == "pathwise_sample" and pw:
            tc = head.S >= 64 and A <= 112
            if tc:
                peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
                peak = 0.5 * float(peaks.get("bf16_tflops", 1590.0))
                flops = head.Bp * head.D * 3 * 2 * 2 * head.S * A * head.B
                kern = "pathwise_tc_kernel (tcgen05.mma kind::tf32, 3-pass split, TMEM partial sums) + gp_prepare_kernel + pathwise_update_mma_kernel"
                bound, src = "tensor", "0.5 x measured bf16 dense peak (MEASURED_PEAKS.json) = TF32 dense rate"
            else:
                peak = float(eng.lib.vgpmp_probe_fp64_tflops(local))
                flops = head.Bp * head.D * (2 * 2 * head.S * A * head.B + 6 * A * head.B)
                kern = "pathwise_rr_kernel + gp_prepare_update_kernel (register-resident rotation chains feeding DMMA m8n8k4, then GP preparation + pathwise update)"
                bound, src = "fp64", "measured now: vgpmp_probe_fp64_tflops (DFMA chains, CUDA events)"
            ach = flops / (pw["ms_per_step"] * 1e-3) / 1e12
            roofline = {"kernel": kern, "bound": bound, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                        "frac": ach / peak if peak > 0 else None, "peak_source": src, "traffic": None,
                        "algorithmic_flops_per_launch": flops, "ms_per_launch": pw["ms_per_step"], "share_of_step": pw["share"]}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": per_step, "higher_is_better": True, "scaling": head.cfg["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": head.cfg["what"], "baseline_config": args.config,
                       "problems_total": head.total_problems, "problems_this_gpu": head.Bp,
                       "S": head.S_total, "N": head.N, "M": head.M, "B": head.B, "dof": head.D, "spheres": head.P,
                       "sdf": f"{head.sdf_desc}; grid {head.sdf.data.shape} float64, {head.sdf.data.nbytes * 4 / 2**20:.0f} MiB of "
                              "{value,gradient} records in HBM, one copy per device (the reference's own .sdf grids are missing blobs)",
                       "rng": "device Philox4x32-10, fresh draws every step (lazy: omega/tau/w generated inside the sampler kernel, never allocated)",
                       "streams": head.streams,
                       "l2": "no explicit flush: every step streams its whole working set, which exceeds the 126 MB L2 (86 MiB of SDF records + ~110 MB of GP factors, prior draws, samples, gradients and Adam state)"},
            "sdf_evals_per_s": head.sdf_evals() * args.steps / (ms / 1000.0),
            "gpu_launches": int(launches), "stages": stages, "roofline": roofline, "roofline_sdf_stage": sdf_roof,
            "dominant_stage": dominant, "configs": extras, "clocks": clk,
        }
        if ms_e2e is not None:
            out["e2e"] = {"value": head.problem_iterations() * args.steps / (ms_e2e_r / 1000.0), "unit": UNIT,
                          "h2d_bytes_per_step": int(head.X.nbytes) * max(head.streams, 1), "d2h_bytes_per_step": int(head.Bp * 8),
                          "ms_per_step": ms_e2e_r / args.steps,
                          "api": ("StreamedVGPMP.train_step_host -> vgpmp_train_step_host_begin/_end per sub-batch"
                                  if head.streams > 1 else
                                  "VGPMP.train_step_host -> pinned X to device, train_step (chunks / all-reduce), loss to pinned host"
                                  if (head.sample_sharded or len(head.model._plan(head.Xd, False)[0]) > 1) else
                                  "VGPMP.train_step_host -> vgpmp_train_step_host")}
        else:
            out["e2e"] = None
        if world == 1 and not args.no_cpu_baseline:
            steps = 40
            rate, dt = cpu_port_rate(args.cpu_problems, steps)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                   "sample": f"the first {args.cpu_problems} of the 275 config-2 problems x {steps} steps ({dt:.1f} s) "
                                             "on the same bookshelves grid, float64 torch-CPU restatement of the reference "
                                             "(TF/GPflow not installable)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
