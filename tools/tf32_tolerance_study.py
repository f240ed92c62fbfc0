#!/usr/bin/env python
"""Tolerance study for moving the random-Fourier contraction  f0 = Phi w  (the path's only dense contraction) to tensor cores.

north_star allows 3xTF32 "only if it holds tolerance".  A 3xTF32 (or bf16x9) split reproduces float32-class products with
float32 accumulation, so its error is emulated here by evaluating the contraction in float32 inside the float64 oracle and
leaving everything else untouched.  Reported per problem: max |df| in the latent paths, number of (sample, timestep)
cells whose log-likelihood changed (nearest-voxel flips), relative ELBO error, relative gradient errors.

    python tools/tf32_tolerance_study.py [--problems 6]
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import vgpmp_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402


def _tf32(x):
    """round-to-nearest float32 -> tf32 (10 explicit mantissa bits), returned as float32"""
    u = x.contiguous().view(torch.int32)
    return ((u + 0x1000) & ~0x1FFF).view(torch.float32)


def _split(x64):
    hi = _tf32(x64.float())
    lo = _tf32((x64 - hi.double()).float())
    return hi, lo


MODE = "f32"
FLUSH = 0          # 3xtf32: bases per float32 accumulation run before the partial sum is added into a float64 total


def contract(phi, w):
    """phi [l,a,b] (float64), w [l,s,b] (float64) -> [l,s,a] float64, computed the way MODE says"""
    if MODE == "f32":
        return torch.einsum("lab,lsb->lsa", phi.float(), w.float()).double()
    exact = torch.einsum("lab,lsb->lsa", phi, w)          # carries the gradient (the kernel's d/dlengthscale contraction
    ph, pl = _split(phi.detach())                         # would run in the same reduced precision: same error class)
    wh, wl = _split(w.detach())
    B = phi.shape[-1]
    step = FLUSH or B
    out = torch.zeros(phi.shape[0], w.shape[1], phi.shape[1], dtype=torch.float64)
    for b0 in range(0, B, step):
        sl = slice(b0, b0 + step)
        part = (torch.einsum("lab,lsb->lsa", pl[..., sl], wh[..., sl]) + torch.einsum("lab,lsb->lsa", ph[..., sl], wl[..., sl])
                + torch.einsum("lab,lsb->lsa", ph[..., sl], wh[..., sl]))
        out += part.double()
    return exact + (out - exact).detach()


class F32Problem(O.OracleProblem):
    """OracleProblem whose Fourier contraction runs in float32 / emulated 3xTF32 (everything else float64)."""

    def sample_paths(self, Xq, q_mu, q_sqrt, lengthscales, variances, draws):
        omega, tau, w = O._t(draws["omega"]), O._t(draws["tau"]), O._t(draws["w"])
        eps_u, eps_j = O._t(draws["eps_u"]), O._t(draws["eps_j"])
        D, B = tau.shape
        Zy, Xq = O._t(self.Zy), O._t(Xq)

        def prior(pts):
            proj = torch.einsum("ad,lbd->lab", pts, omega) / lengthscales[:, None, None]
            phi = torch.sqrt(2.0 * variances / B)[:, None, None] * torch.cos(proj + tau[:, None, :])
            return contract(phi, w)     # <- reduced-precision contraction

        mu = self.q_mu_full(q_mu).T
        Sfull = self.q_sqrt_full(q_sqrt, lengthscales, variances)
        u = mu[:, None, :] + torch.einsum("lmk,lsk->lsm", Sfull, eps_u)
        K = O.kuu(Zy, lengthscales, variances, 0.0) + O.JITTER * torch.eye(Zy.shape[0], dtype=torch.float64)[None]
        Lc = torch.linalg.cholesky(K)
        err = u - prior(Zy) - np.sqrt(O.JITTER) * eps_j
        v = torch.cholesky_solve(err.transpose(1, 2), Lc).transpose(1, 2)
        Kfu = O.k_conditioned(Zy, Xq, lengthscales, variances).transpose(1, 2)
        return (prior(Xq) + torch.einsum("lnm,lsm->lsn", Kfu, v)).permute(1, 2, 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=6)
    ap.add_argument("--mode", choices=["f32", "3xtf32"], default="f32")
    ap.add_argument("--flush", type=int, default=0, help="3xtf32: float32 accumulation run length in bases (0 = all)")
    a = ap.parse_args()
    global MODE, FLUSH
    MODE, FLUSH = a.mode, a.flush
    case = H.make_case("franka", "bookshelves", num_problems=a.problems, B=1024, seed=11, perturb=True)   # perturbed: trajectories touch the obstacles
    worst = dict(f=0.0, elbo=0.0, grad=0.0, flips=0)
    print("problem  max|df|     flipped cells   rel ELBO err   rel grad err (q_mu, q_sqrt, ls, var)")
    for b, p in enumerate(case["oracle"]):
        args = (case["q_mu"][b], case["q_sqrt"][b], case["ls"][b], case["var"][b], case["draws"][b])
        ref = O.elbo_and_grads(p, *args)
        p32 = F32Problem(**p.__dict__)
        got = O.elbo_and_grads(p32, *args)
        df = np.abs(got["f"] - ref["f"]).max()
        flips = int((np.abs(got["logp"] - ref["logp"]) > 1e-9 * max(1.0, np.abs(ref["logp"]).max())).sum())
        e = abs(got["elbo"] - ref["elbo"]) / abs(ref["elbo"])
        g = [H.rel_err(got["d_" + k], ref["d_" + k]) for k in ("q_mu", "q_sqrt", "lengthscales", "variances")]
        worst = dict(f=max(worst["f"], df), elbo=max(worst["elbo"], e), grad=max(worst["grad"], max(g)), flips=worst["flips"] + flips)
        active = int((ref["logp"] != 0).sum())
        print(f"{b:7d}  {df:9.2e}  {flips:6d}/{ref['logp'].size:<6d} (active {active:4d})  {e:12.2e}   " + " ".join(f"{x:8.1e}" for x in g))
    print(f"worst: max|df| {worst['f']:.2e}, ELBO {worst['elbo']:.2e} (tolerance 1e-4), gradients {worst['grad']:.2e} (tolerance 1e-3), "
          f"{worst['flips']} flipped cells")


if __name__ == "__main__":
    main()
