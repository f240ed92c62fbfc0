#!/usr/bin/env python
"""SHA-256 of the device draws for a few fixed shapes / offsets (GPU only): pins the Philox key layout across refactors.
    python tools/rng_checksum.py            # prints one line per case; tests/golden/rng_sha256.json holds the pinned values
"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tests import helpers as H  # noqa: E402

CASES = [dict(Bp=3, S=9, M=5, B=37, seed=7, it=3, po=0, so=0), dict(Bp=2, S=7, M=24, B=1024, seed=1234, it=0, po=5, so=0),
         dict(Bp=1, S=33, M=12, B=64, seed=99, it=17, po=2, so=40)]


def checksums():
    case = H.make_case(num_problems=1, S=3, N=5, M=3, B=8, seed=1)
    eng = H.make_model(case)._eng
    out = {}
    for c in CASES:
        dims = eng.dims(c["Bp"], c["M"], 10, c["S"], c["B"])
        d = eng.rng_fill(dims, c["seed"], c["it"], problem_offset=c["po"], sample_offset=c["so"])
        key = "Bp{Bp}_S{S}_M{M}_B{B}_seed{seed}_it{it}_po{po}_so{so}".format(**c)
        out[key] = {k: hashlib.sha256(v.cpu().numpy().tobytes()).hexdigest() for k, v in sorted(d.items())}
    return out


if __name__ == "__main__":
    print(json.dumps(checksums(), indent=1, sort_keys=True))
