"""A batch of planning problems driven as several sub-batches on their own CUDA streams.

The per-iteration path alternates an FP64-pipe-bound sampler with latency-bound small-matrix kernels (GP preparation,
reverse pass).  The problems of a batch are independent (`benchmarking.py:70-100` solves them one after another), so two
sub-batches whose steps are issued on two streams fill each other's gaps: +10 % on one B200 for the 275-problem batch.
Draws are keyed by the GLOBAL problem index (`VGPMP.problem_offset`), so the split batch follows the unsplit batch's
optimisation trajectory bit for bit.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from .vgpmp import VGPMP


class StreamedVGPMP:
    """`num_streams` VGPMP models over contiguous slices of the problem batch, one CUDA stream each.

    Same calls as `VGPMP` for the training path: `train_step`, `train_step_host`, plus gathered views of the variational
    state.  Everything else (prediction, sampling, verdicts) is available per sub-model through `.models`."""

    def __init__(self, models: List[VGPMP], device=None):
        self.models = models
        dev = models[0]._eng.device
        self.streams = [torch.cuda.Stream(device=dev) for _ in models]
        self.num_problems = sum(m.num_problems for m in models)
        self._host_loss = None

    @classmethod
    def initialize(cls, query_states, num_streams: int = 2, **kw) -> "StreamedVGPMP":
        """`VGPMP.initialize(**kw)` per slice of `query_states` [Bp,2,D]; all slices share the seed and differ only in
        `problem_offset`."""
        qs = np.asarray(query_states, dtype=np.float64)
        if qs.ndim != 3:
            raise ValueError("StreamedVGPMP needs a batch of problems: query_states [Bp,2,D]")
        num_streams = max(1, min(int(num_streams), qs.shape[0]))
        models, start = [], 0
        for idx in np.array_split(np.arange(qs.shape[0]), num_streams):
            if models and "share_engine" not in kw:
                kw = dict(kw, share_engine=models[0]._eng)      # one copy of the SDF records, one handle per stream
            m = VGPMP.initialize(query_states=qs[idx], **kw)
            m.problem_offset = start
            start += len(idx)
            models.append(m)
        return cls(models)

    # ------------------------------------------------------------------ training path
    def _fork(self):
        cur = torch.cuda.current_stream()
        for s in self.streams:
            s.wait_stream(cur)

    def _join(self):
        cur = torch.cuda.current_stream()
        for s in self.streams:
            cur.wait_stream(s)

    def train_step(self, X) -> torch.Tensor:
        """One optimisation step of every sub-batch; returns loss = -ELBO [Bp] (device tensor)."""
        self._fork()
        out = []
        for m, s in zip(self.models, self.streams):
            with torch.cuda.stream(s):
                out.append(m.train_step(X).reshape(-1))
        self._join()
        return torch.cat(out)

    def train_step_host(self, X_host: torch.Tensor) -> torch.Tensor:
        """Host-buffer step (`vgpmp_train_step_host_begin/_end` per sub-batch): all sub-batches are enqueued before the
        first one is waited for; their losses land in slices of ONE pinned buffer.  Returns a CPU tensor [Bp] (a copy)."""
        if self._host_loss is None:
            self._host_loss = torch.empty(self.num_problems, dtype=torch.float64).pin_memory()
            parts, lo = [], 0
            for m in self.models:
                parts.append(self._host_loss[lo:lo + m.num_problems])
                lo += m.num_problems
            self._host_parts = parts
            for m, s, part in zip(self.models, self.streams, parts):      # first step: allocations under the right stream
                with torch.cuda.stream(s):
                    m.train_step_host(X_host, wait=False, stream=s.cuda_stream, loss_out=part)
        else:
            for m, s, part in zip(self.models, self.streams, self._host_parts):
                m.train_step_host(X_host, wait=False, stream=s.cuda_stream, loss_out=part)
        for m in self.models:
            m.train_step_host_wait(copy=False)
        return self._host_loss.clone()

    # ------------------------------------------------------------------ gathered state
    def _cat(self, name):
        return torch.cat([getattr(m, name) for m in self.models])

    @property
    def q_mu(self):
        return self._cat("_q_mu")

    @property
    def q_sqrt(self):
        return self._cat("_q_sqrt")

    @property
    def lengthscales(self):
        return self._cat("_lengthscales")

    @property
    def variances(self):
        return self._cat("_variances")

    @property
    def launch_count(self) -> int:
        return sum(m._eng.launch_count for m in self.models)
