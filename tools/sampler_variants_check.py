#!/usr/bin/env python
"""Agreement of the equispaced sampler variants with the general per-point sincos kernel (GPU only).

    python tools/sampler_variants_check.py
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
from tests import helpers as H  # noqa: E402


def main():
    case = H.make_case("franka", "bookshelves", num_problems=8, B=1024, seed=3)
    model = H.make_model(case)
    eng = model._eng
    out = {}
    for name, opts in {"general": dict(grid_fast_path=0), "cta_fma": dict(grid_fast_path=1, dmma_sampler=0),
                       "dmma": dict(grid_fast_path=1, dmma_sampler=1, rr_sampler=0),
                       "rr": dict(grid_fast_path=1, dmma_sampler=1, rr_sampler=1)}.items():
        for k, v in opts.items():
            eng.set_option(k, v)
        out[name] = model.predict_f_samples(case["X"], draws=case["draws_stacked"]).cpu().numpy()
    ref = out["general"]
    for k in ("cta_fma", "dmma", "rr"):
        print(f"{k:8s} vs general: max|df| {np.abs(out[k] - ref).max():.3e}   rel {H.rel_err(out[k], ref):.3e}")


if __name__ == "__main__":
    main()
