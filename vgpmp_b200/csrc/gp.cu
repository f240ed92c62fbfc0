// Gaussian-process side of the vgpmp ELBO hot path (sm_100a, float64):
//   Kuu / Kuf builds, Cholesky, q_sqrt un-whitening, endpoint-conditioned KL      (gp_prepare_kernel)
//   random-Fourier prior + pathwise update -> latent trajectory samples             (pathwise_kernel)
//   reverse pass to _q_mu, _q_sqrt, lengthscales, variances                         (gp_backward_kernel)
//   ELBO assembly, Keras-Adam update, Philox draw generator.
//
// Reference semantics restated (files under /root/reference; GPflow / GPflowSampling are un-vendored):
//   kernel_conditioning/multioutput/cond_kernel.py:17-25, cond_kernel.py:19-22   per-latent 1-D Matern-5/2
//   covariances/multioutput/Kuus.py:42-53, Kufs.py:26-34, covariances/Kfus.py:36-42
//   models/vgpmp.py:200-218 (q_mu / q_sqrt properties), :265-289 (elbo)
//   kullback_leiblers/prior_kl.py:16-35 + gpflow.kullback_leiblers.gauss_kl (white)
//   GPflowSampling random_fourier / PathwiseSVGP.generate_paths / exact_update
//
// One CTA per (problem, latent GP).  The inducing covariance is at most 32x32, so all dense algebra on it lives in
// shared memory; triangular solves are warp-per-right-hand-side with the running vector in registers and the pivot
// broadcast by shuffle.  Matrices in shared memory use a row stride of 33 doubles so that both row and column walks
// are bank-conflict free.
#include <cstdlib>
#include <algorithm>

#include "common.cuh"
#include "device_utils.cuh"

namespace {

constexpr int LDM = 33;          // shared-memory leading dimension of the Mp x Mp matrices
constexpr int kST = 8;           // samples per register tile in the Fourier contraction
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double zy_at(const double* __restrict__ Z, int D, int l, int m) {
  return m == 0 ? 0.0 : (m == 1 ? 1.0 : Z[(size_t)(m - 2) * D + l]);  // Zy = [0; 1; Z]  inducing_variables.py:56-62
}

__device__ __forceinline__ double block_sum(double v, double* red) {
  // deterministic block reduction (fixed tree), result valid in every thread
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// In-place Cholesky of the leading Mp x Mp block of A (lower), whole CTA cooperates (warps walk rows, lanes walk
// columns: no integer division in the index math).  Upper triangle zeroed.
__device__ void chol_inplace(double* A, int Mp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int k = 0; k < Mp; ++k) {
    __syncthreads();
    // pivot: 1/sqrt by the hardware seed + Newton, sqrt from it with one correction step (residual in FMA): the sqrt /
    // divide pair of the first version was half of this routine's instructions, and every thread runs it
    const double akk = A[k * LDM + k];
    double inv = rsqrt(akk);
    double piv = akk * inv;
    piv = fma(0.5 * inv, fma(-piv, piv, akk), piv);        // piv += (a - piv^2) / (2 piv)
    inv = fma(inv, fma(-piv, inv, 1.0), inv);              // inv += inv (1 - piv inv)
    if (tid > k && tid < Mp) A[tid * LDM + k] *= inv;      // own element only: no hazard with the a_kk reads
    __syncthreads();
    if (tid == 0) A[k * LDM + k] = piv;                    // nobody reads a_kk any more; the update touches j > k
    const int j = k + 1 + lane;
    for (int i = k + 1 + warp; i < Mp; i += nw)
      if (j <= i) A[i * LDM + j] -= A[i * LDM + k] * A[j * LDM + k];
  }
  __syncthreads();
  for (int i = warp; i < Mp; i += nw)
    if (lane > i && lane < Mp) A[i * LDM + lane] = 0.0;
  __syncthreads();
}

// Warp-level triangular solves; lane i owns component i (i < Mp <= 32).  L lower-triangular in shared memory; every lane
// keeps the reciprocal of its own diagonal entry so the pivot step is a multiply (the reciprocal travels with the shuffle).
__device__ __forceinline__ double warp_fwd_subst(const double* L, int Mp, double d) {  // returns (L^-1 d)_lane
  const int lane = threadIdx.x & 31;
  const double inv = lane < Mp ? 1.0 / L[lane * LDM + lane] : 0.0;
  double out = 0.0;
  for (int k = 0; k < Mp; ++k) {
    const double bk = __shfl_sync(kFull, d * inv, k);
    if (lane == k) out = bk;
    if (lane > k && lane < Mp) d -= L[lane * LDM + k] * bk;
  }
  return out;
}
__device__ __forceinline__ double warp_bwd_subst(const double* L, int Mp, double d) {  // returns (L^-T d)_lane
  const int lane = threadIdx.x & 31;
  const double inv = lane < Mp ? 1.0 / L[lane * LDM + lane] : 0.0;
  double out = 0.0;
  for (int k = Mp - 1; k >= 0; --k) {
    const double bk = __shfl_sync(kFull, d * inv, k);
    if (lane == k) out = bk;
    if (lane < k) d -= L[k * LDM + lane] * bk;
  }
  return out;
}

// Products with the explicit inverse factor Li = L^-1 (lower, shared memory): lane i owns component i, the input vector
// sits in shared memory (all lanes read the same element -> broadcast), no shuffles and no cross-lane dependency chain.
__device__ __forceinline__ double warp_lower_mv(const double* Li, int Mp, const double* x) {   // (Li x)_lane
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  if (lane < Mp)
    for (int k = 0; k <= lane; ++k) acc += Li[lane * LDM + k] * x[k];
  return acc;
}
__device__ __forceinline__ double warp_lowerT_mv(const double* Li, int Mp, const double* x) {  // (Li^T x)_lane
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  if (lane < Mp)
    for (int k = lane; k < Mp; ++k) acc += Li[k * LDM + lane] * x[k];
  return acc;
}

// single-exp evaluation of the Matern-5/2 profile and its derivative
__device__ __forceinline__ void matern52_both(double r, double& k, double& dk) {
  const double a = VG_SQRT5 * r, e = exp(-a);
  k = (1.0 + a + (5.0 / 3.0) * r * r) * e;
  dk = -(5.0 / 3.0) * r * (1.0 + a) * e;
}

__global__ void kuu_kernel(int D, int M, double jitter, const double* __restrict__ Z, const double* __restrict__ ls,
                           const double* __restrict__ var, double* __restrict__ K) {
  const int pl = blockIdx.x, l = pl % D, Mp = M + 2;
  const double ell = ls[pl], s2 = var[pl];
  for (int idx = threadIdx.x; idx < Mp * Mp; idx += blockDim.x) {
    const int i = idx / Mp, j = idx % Mp;
    const double r = fabs(zy_at(Z, D, l, i) - zy_at(Z, D, l, j)) / ell;
    K[(size_t)pl * Mp * Mp + idx] = s2 * vg_matern52(r) + (i == j ? jitter : 0.0);
  }
}

__global__ void kuf_kernel(int D, int M, int N, const double* __restrict__ Z, const double* __restrict__ X,
                           const double* __restrict__ ls, const double* __restrict__ var, double* __restrict__ Kuf) {
  const int pl = blockIdx.x, l = pl % D, Mp = M + 2;
  const double ell = ls[pl], s2 = var[pl];
  for (int idx = threadIdx.x; idx < Mp * N; idx += blockDim.x) {
    const int m = idx / N, n = idx % N;
    const double r = fabs(zy_at(Z, D, l, m) - X[(size_t)n * D + l]) / ell;
    Kuf[(size_t)pl * Mp * N + idx] = s2 * vg_matern52(r);
  }
}

// ---------------------------------------------------------------------------------------------
// Kuu + chol + q_sqrt un-whitening + KL, one CTA (128 threads) per (problem, latent)
// ---------------------------------------------------------------------------------------------
constexpr int kPrepSmem = 3 * 32 * LDM + 2 * 32 + 64;   // doubles of shared memory gp_prepare_body needs: K (later q, then q_sqrt_full) | L | L^-1 | zy | mu | scratch

// Body of the GP preparation for one (problem, latent), working on a caller-provided shared-memory block.
// `smem` must hold kPrepSmem doubles; needs blockDim.x >= 32 and a multiple of 32.
__device__ void gp_prepare_body(int D, int M, double jitter, const vgpmp_params& P, int pl, double* __restrict__ Lc_out,
                                double* __restrict__ S_out, double* __restrict__ kl_l, double* __restrict__ kvec,
                                double* __restrict__ Linv_out, double* smem, double* Ssm = nullptr) {
  // Ssm (optional, [32][LDM] in shared memory, may alias smem = the K block): a copy of q_sqrt_full for a caller that
  // continues with the pathwise update in the same CTA; the explicit inverse factor then stays in smem + 3*32*LDM.
  double* Ksm = smem;              // K until the KL term is done, then pad(q), then q_sqrt_full
  double* Lsm = Ksm + 32 * LDM;
  double* Li = Lsm + 32 * LDM;
  double* qsm = Ksm;
  double* zy = Li + 32 * LDM;
  double* mu = zy + 32;
  double* red = mu + 32;           // [8] block_sum scratch, then [32] reciprocal diagonal of L
  double* cvec = red + 8 + 32;
  const int p = pl / D, l = pl % D, Mp = M + 2, tid = threadIdx.x;
  const double ell = P.lengthscales[pl], s2 = P.variances[pl];
  if (tid < Mp) {
    zy[tid] = zy_at(P.Z, D, l, tid);
    mu[tid] = tid < 2 ? P.query_latent[((size_t)p * 2 + tid) * D + l] : P.q_mu[((size_t)p * M + tid - 2) * D + l];
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const double* q = P.q_sqrt + (size_t)pl * M * M;
  for (int idx = tid; idx < 32 * LDM; idx += blockDim.x) Lsm[idx] = 0.0;
  __syncthreads();
  for (int i = warp; i < Mp; i += nw) {
    if (lane < Mp) {
      const double k = s2 * vg_matern52(fabs(zy[i] - zy[lane]) / ell) + (i == lane ? jitter : 0.0);
      Ksm[i * LDM + lane] = k;
      Lsm[i * LDM + lane] = k;
    }
  }
  chol_inplace(Lsm, Mp);
  if (Lc_out != nullptr)
    for (int i = warp; i < Mp; i += nw)
      if (lane < Mp) Lc_out[(size_t)pl * Mp * Mp + i * Mp + lane] = Lsm[i * LDM + lane];
  // explicit inverse factor: thread j holds column j of L^-1 in registers,
  //   x_i = (delta_ij - sum_{k<i} L[i][k] x_k) / L[i][i]   (x_k = 0 for k < j; L is zero-padded beyond Mp)
  if (Linv_out != nullptr || Ssm != nullptr) {
    if (tid >= 32 && tid < 64) {                          // reciprocal diagonal, off the critical path of warp 0
      const int i = tid - 32;
      red[8 + i] = i < Mp ? 1.0 / Lsm[i * LDM + i] : 0.0;  // (red has 8 + 32 doubles of scratch behind it: cvec starts later)
    }
    __syncthreads();
    if (tid < 32) {
      const int j = tid;
      double x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) acc -= Lsm[i * LDM + k] * x[k];
        x[i] = (i < Mp && i >= j && j < Mp) ? acc * red[8 + i] : 0.0;
        Li[i * LDM + j] = x[i];
      }
    }
    __syncthreads();
    if (Linv_out != nullptr)
      for (int i = warp; i < Mp; i += nw)
        if (lane < Mp) Linv_out[(size_t)pl * Mp * Mp + i * Mp + lane] = Li[i * LDM + lane];
  }

  // prior_kl: p_mu = K[:, :2] chol_solve(L[:2,:2], q~);  a = (L^-1 (mu - p_mu))[2:]   prior_kl.py:24-34
  if (tid == 0) {
    const double l00 = Lsm[0], l10 = Lsm[LDM], l11 = Lsm[LDM + 1];
    const double y0 = mu[0] / l00, y1 = (mu[1] - l10 * y0) / l11;
    const double c1 = y1 / l11, c0 = (y0 - l10 * c1) / l00;
    cvec[0] = c0; cvec[1] = c1;
  }
  __syncthreads();
  double part = 0.0;
  if (tid < 32) {
    double d = 0.0;
    if (tid < Mp) d = mu[tid] - (Ksm[tid * LDM] * cvec[0] + Ksm[tid * LDM + 1] * cvec[1]);
    const double b = warp_fwd_subst(Lsm, Mp, d);
    if (kvec != nullptr) {
      if (tid < Mp) kvec[(size_t)pl * (Mp + 4) + tid] = b;
      if (tid < 2) kvec[(size_t)pl * (Mp + 4) + Mp + tid] = cvec[tid];
    }
    if (tid >= 2 && tid < Mp) part = b * b;  // mahalanobis of the whitened difference
  }
  // gauss_kl(white): 0.5 * (maha - M - sum log diag(q)^2 + sum q^2)
  for (int a_ = warp; a_ < M; a_ += nw) {
    if (lane <= a_) {
      const double v = q[a_ * M + lane];
      part += v * v;
    }
  }
  if (tid < M) {                         // log of the diagonal: one pass of one warp (inside the row loop every row paid for it)
    const double v = q[tid * M + tid];
    part -= log(v * v);
  }
  const double tot = block_sum(part, red);   // (its barriers also retire the last readers of K)
  if (tid == 0 && kl_l != nullptr) kl_l[pl] = 0.5 * (tot - (double)M);
  // q_sqrt property: Lc @ pad(_q_sqrt) + jitter * diag(1,1,0,...)   models/vgpmp.py:208-218
  // K is dead: its block takes q, and then (through registers) q_sqrt_full itself when the caller passes Ssm == smem
  if (S_out != nullptr || Ssm != nullptr) {
    for (int i = warp; i < M; i += nw)
      if (lane < M) qsm[i * LDM + lane] = lane <= i ? q[i * M + lane] : 0.0;
    __syncthreads();
    double sreg[8];                    // rows warp, warp + nw, ... (nw >= 4 -> at most 8 rows of 32)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int i = warp + r * nw, j = lane;
      double acc = 0.0;
      if (i < Mp && j < Mp) {
        if (i >= 2 && j >= 2 && j <= i)
          for (int k = j; k <= i; ++k) acc += Lsm[i * LDM + k] * qsm[(k - 2) * LDM + (j - 2)];
        if (i == j && i < 2) acc += jitter;
      }
      sreg[r] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int i = warp + r * nw, j = lane;
      if (i < Mp && j < Mp) {
        if (S_out != nullptr) S_out[(size_t)pl * Mp * Mp + i * Mp + j] = sreg[r];
        if (Ssm != nullptr) Ssm[i * LDM + j] = sreg[r];
      }
    }
  }

}

__global__ void __launch_bounds__(128) gp_prepare_kernel(int D, int M, double jitter, vgpmp_params P,
                                                        double* __restrict__ Lc_out, double* __restrict__ S_out,
                                                        double* __restrict__ kl_l, double* __restrict__ kvec,
                                                        double* __restrict__ Linv_out) {
  __shared__ double smem[kPrepSmem];
  gp_prepare_body(D, M, jitter, P, blockIdx.x, Lc_out, S_out, kl_l, kvec, Linv_out, smem);
}

// ---------------------------------------------------------------------------------------------
// Random-Fourier prior draw + pathwise (Matheron) update, one CTA per (problem, latent).
//   phase 1: f0[s,x] = sum_b phi_b(x) w[s,b],  h0 = d f0 / d lengthscale, x over the Nq query points then the Mp
//            inducing points.  A warp owns 32 points and one slice of the bases; the cos/sin feature lives in a
//            register and is consumed immediately by the kST sample accumulators (w is a warp-uniform load).
//   phase 2: u = mu + S eps_u;  v = Khat^-1 (u - f0(Zy) - sqrt(jitter) eps_j);  f = f0(X) + Kfu v.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(768) pathwise_kernel(PathwiseArgs a, const double* __restrict__ meta) {
  if (meta != nullptr && meta[0] != 0.0) return;  // equispaced rank-1 inputs: the equispaced samplers do the work
  extern __shared__ double sm[];
  // work items = (problem, latent, sample chunk); the grid may be smaller than their number (the launcher sends only a
  // couple of CTAs per SM when the equispaced path is expected to take the work: an empty launch of 1925 x 768 threads
  // cost 6 us per step)
  for (int item = blockIdx.x; item < a.items; item += gridDim.x) {
  const int D = a.D, M = a.M, Mp = M + 2, Nq = a.Nq, S = a.S, B = a.B, A = Nq + Mp;
  const int pl = item / a.nchunk, p = pl / D, l = pl % D;
  const int s_begin = (item % a.nchunk) * a.chunk, s_end = min(a.S, s_begin + a.chunk);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x;
  const int XP = a.XG * 32;
  // shared-memory carve-up
  double* red = sm;                               // [KS][2][kST][XP]
  double* Lsm = red + (size_t)a.KS * 2 * kST * XP;  // [Mp][LDM]
  double* Ssm = Lsm + 32 * LDM;                   // [Mp][LDM]
  double* Kfu = Ssm + 32 * LDM;                   // [Nq][Mp]
  double* vs = Kfu + (size_t)Nq * Mp;             // [kST][32]
  double* mu = vs + kST * 32;                     // [32]
  double* zy = mu + 32;                           // [32]

  const double ell = a.ls[pl], s2 = a.var[pl];
  const double amp = sqrt(2.0 * s2 / (double)B);
  const double sqrtj = sqrt(a.jitter);
  if (tid < Mp) {
    zy[tid] = zy_at(a.Z, D, l, tid);
    mu[tid] = tid < 2 ? a.query_latent[((size_t)p * 2 + tid) * D + l] : a.q_mu[((size_t)p * M + tid - 2) * D + l];
  }
  for (int idx = tid; idx < Mp * Mp; idx += nt) {
    const int i = idx / Mp, j = idx % Mp;
    Lsm[i * LDM + j] = a.Lc[(size_t)pl * Mp * Mp + idx];
    Ssm[i * LDM + j] = a.Sfull[(size_t)pl * Mp * Mp + idx];
  }
  __syncthreads();
  for (int idx = tid; idx < Nq * Mp; idx += nt) {
    const int n = idx / Mp, m = idx % Mp;
    Kfu[idx] = s2 * vg_matern52(fabs(a.Xq[(size_t)n * D + l] - zy[m]) / ell);
  }

  // this thread's point, scaled by 1/lengthscale (GPflow kernel.scale)
  const int xg = warp % a.XG, ks = warp / a.XG;
  const int x = xg * 32 + lane;
  double pt[VGPMP_MAX_DOF];
#pragma unroll
  for (int d = 0; d < VGPMP_MAX_DOF; ++d) {
    double v = 0.0;
    if (d < D && x < A) {
      if (x < Nq) v = a.Xq[(size_t)x * D + d];
      else v = zy_at(a.Z, D, d, x - Nq);
    }
    pt[d] = v / ell;
  }
  const int Bs = (B + a.KS - 1) / a.KS;
  const int b0 = ks * Bs, b1 = min(B, b0 + Bs);
  const double* om = a.omega + (size_t)pl * B * D;
  const double* ta = a.tau + (size_t)pl * B;
  const double* wp = a.w + (size_t)pl * S * B;

  for (int s0 = s_begin; s0 < s_end; s0 += kST) {
    double acc0[kST], acc1[kST];
#pragma unroll
    for (int i = 0; i < kST; ++i) { acc0[i] = 0.0; acc1[i] = 0.0; }
    const int ns = min(kST, s_end - s0);
    if (ks < a.KS) {
      for (int b = b0; b < b1; ++b) {
        double pr = 0.0;
#pragma unroll
        for (int d = 0; d < VGPMP_MAX_DOF; ++d)
          if (d < D) pr += pt[d] * __ldg(om + (size_t)b * D + d);
        double sn, cs;
        sincos(pr + __ldg(ta + b), &sn, &cs);
        const double c = amp * cs, g = amp * sn * pr / ell;
#pragma unroll
        for (int i = 0; i < kST; ++i) {
          if (i < ns) {
            const double wv = __ldg(wp + (size_t)(s0 + i) * B + b);
            acc0[i] += c * wv;
            acc1[i] += g * wv;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < kST; ++i) {
        red[((size_t)(ks * 2 + 0) * kST + i) * XP + x] = acc0[i];
        red[((size_t)(ks * 2 + 1) * kST + i) * XP + x] = acc1[i];
      }
    }
    __syncthreads();
    // reduce the base slices (fixed order) into slice 0 and publish f0 / h0
    for (int idx = tid; idx < 2 * kST * XP; idx += nt) {
      double t = red[idx];
      for (int k = 1; k < a.KS; ++k) t += red[(size_t)k * 2 * kST * XP + idx];
      red[idx] = t;
      const int which = idx / (kST * XP), i = (idx / XP) % kST, xx = idx % XP;
      if (i < ns && xx < A) {
        double* dst = which == 0 ? a.f0 : a.h0;
        if (dst != nullptr) dst[((size_t)pl * S + s0 + i) * A + xx] = t;
      }
    }
    __syncthreads();
    // pathwise update, one warp per sample of the tile
    if (warp < ns) {
      const int s = s0 + warp;
      const double* eu = a.eps_u + ((size_t)pl * S + s) * Mp;
      double r = 0.0;
      if (lane < Mp) {
        double u = mu[lane];
        for (int k = 0; k <= lane; ++k) u += Ssm[lane * LDM + k] * eu[k];
        r = u - red[(size_t)warp * XP + Nq + lane] - sqrtj * a.eps_j[((size_t)pl * S + s) * Mp + lane];
      }
      const double y = warp_fwd_subst(Lsm, Mp, r);
      const double vv = warp_bwd_subst(Lsm, Mp, y);
      vs[warp * 32 + lane] = lane < Mp ? vv : 0.0;
      if (lane < Mp && a.v != nullptr) a.v[((size_t)pl * S + s) * Mp + lane] = vv;
    }
    __syncthreads();
    for (int idx = tid; idx < ns * Nq; idx += nt) {
      const int i = idx / Nq, n = idx % Nq;
      double fv = red[(size_t)i * XP + n];
      for (int m = 0; m < Mp; ++m) fv += Kfu[n * Mp + m] * vs[i * 32 + m];
      a.f[f_index(a.f_planar, p, s0 + i, n, l, S, Nq, D)] = fv;
    }
    __syncthreads();
  }
  __syncthreads();
  }
}


// ---------------------------------------------------------------------------------------------
// Input-structure probe.  The reference always evaluates on X = linspace replicated over the D columns
// (utils/miscellaneous.py:115-127) with inducing inputs linspace(.1,.9,M) replicated likewise (models/vgpmp.py:37-42).
// When both are rank-1 and equispaced the Fourier phase of basis b at point n is (t0 + n dt) c_b + tau_b, so the
// cos/sin features follow from one rotation per step instead of one sincos per point.  meta = {flag, t0, dt, z0, dz}.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) analyze_grid_kernel(int D, int M, int Nq, const double* __restrict__ Xq,
                                                          const double* __restrict__ Z, double* __restrict__ meta) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  const double t0 = Xq[0], dt = Nq > 1 ? (Xq[(size_t)(Nq - 1) * D] - Xq[0]) / (double)(Nq - 1) : 0.0;
  const double z0 = Z[0], dz = M > 1 ? (Z[(size_t)(M - 1) * D] - Z[0]) / (double)(M - 1) : 0.0;
  const double tolx = 8.0 * 2.220446049250313e-16 * fmax(fmax(fabs(t0), fabs(t0 + dt * (Nq - 1))), 1e-300);
  const double tolz = 8.0 * 2.220446049250313e-16 * fmax(fmax(fabs(z0), fabs(z0 + dz * (M - 1))), 1e-300);
  int mybad = 0;
  for (int i = threadIdx.x; i < Nq * D; i += blockDim.x) {
    const int n = i / D;
    if (!(fabs(Xq[i] - (t0 + dt * n)) <= tolx)) mybad = 1;
  }
  for (int i = threadIdx.x; i < M * D; i += blockDim.x) {
    const int m = i / D;
    if (!(fabs(Z[i] - (z0 + dz * m)) <= tolz)) mybad = 1;
  }
  if (mybad) bad = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    meta[0] = bad ? 0.0 : 1.0;
    meta[1] = t0; meta[2] = dt; meta[3] = z0; meta[4] = dz;
  }
}

// ---------------------------------------------------------------------------------------------
// Equispaced sampler with the contraction on the FP64 tensor path (DMMA m8n8k4), features through shared memory: the
// fallback of the register-resident sampler below for point sets that need more than its 12 row tiles (N + M + 2 <= 192,
// e.g. the 150-point prediction grid).  Per tile of 32 bases:
//   phase 1  thread (basis, chunk of points): one complex rotation per point from tabulated start / step phasors; cos
//            feature and lengthscale-derivative feature go to shared memory [point][basis];
//   phase 2  C[8 points x 8 samples] += A[8 x 4 bases] B[4 x 8]  with A fragments read straight from the feature tile
//            (one 64-bit shared load per 256 FMAs) and C held in 2*PT registers per thread for the whole base loop.
// B200 executes DMMA at the DFMA rate (profiles/r1_v9_dmma_probe.txt), so the gain over an FMA contraction is issue
// slots and operand traffic, not flops.  The warps that are light in phase 2 also precompute, for the NEXT tile,
// everything phase 1 would otherwise recompute per chunk (step phasors of both grids, the two conditioned endpoints).
// The kernel stops at the prior draw f0 / h0; gp_prepare_update_kernel finishes the sample paths.
// ---------------------------------------------------------------------------------------------
constexpr int kDB = 32;    // bases per tile
constexpr int kDBP = 36;   // padded basis stride of a feature row: the 4 rows x 4 lanes of an A fragment hit 16 distinct banks
constexpr int kWS = 12;    // padded sample stride of a staged weight row (same argument for the B fragment)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// one chunk of an equispaced run written point-major (feature of point n at fc[n * kDBP]); start phasor (amplitude
// folded in) and step phasor come from the tables the previous tile's contraction phase prepared
__device__ __forceinline__ void rotate_run_pm(double* fc, double* fd, int col0, int first, int count, double g0, double dg,
                                              double c, double inv_ell, double cs, double sn, double cd, double sd) {
  if (count <= 0) return;
  const double tn = g0 + dg * first;
  double q = tn * c * inv_ell;
  const double dq = dg * c * inv_ell;
  fc += (size_t)(col0 + first) * kDBP;
  fd += (size_t)(col0 + first) * kDBP;
  for (int k = 0; k < count; ++k) {
    fc[k * kDBP] = cs;
    fd[k * kDBP] = sn * q;
    const double c2 = cs * cd - sn * sd;
    sn = sn * cd + cs * sd;
    cs = c2;
    q += dq;
  }
}

template <int PT>
__global__ void __launch_bounds__(256, 3) pathwise_dmma_kernel(PathwiseArgs a, const double* __restrict__ meta) {
  extern __shared__ __align__(16) double sm[];
  constexpr int ROWS = 32 * PT;                 // padded number of points: 8 warps x PT units = 4*PT point tiles x 2 features
  constexpr int PLANE = ROWS * kDBP;            // one feature plane [point][basis]
  const int D = a.D, M = a.M, Mp = M + 2, Nq = a.Nq, S = a.S, B = a.B, A = Nq + Mp;
  const int pl = blockIdx.x / a.nchunk, p = pl / D, l = pl % D;
  const int s_begin = (blockIdx.x % a.nchunk) * a.chunk, s_end = min(a.S, s_begin + a.chunk);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x;
  const int T = (B + kDB - 1) / kDB;
  // main-loop view
  double* feat = sm;                            // [2 features][ROWS][kDBP]
  double* Wt = feat + 2 * PLANE;                // [3][kDB][kWS]
  double* osum = Wt + 3 * kDB * kWS;            // [3][4][kDB]  partial sums over the input dims of omega_b
  double* tb = osum + 3 * 4 * kDB;              // [3][kDB]     tau_b
  // tables for the tile about to be generated, written during the previous tile's contraction phase
  double* stt = tb + 3 * kDB;                   // [8 chunks][kDB][2]  chunk-start phasors, amplitude folded in
  double* stp = stt + 8 * kDB * 2;              // [2][kDB][2]  step phasors (cos, sin) of the query grid | inducing grid
  double* ept = stp + 2 * kDB * 2;              // [2][kDB][2]  unit phasors of the conditioned endpoints Zy = 0 | 1
  // after the basis loop the feature tile is dead: the accumulators are transposed through it
  double* red = sm;                             // [2][kST][ROWS]

  if (meta[0] == 0.0) return;  // not an equispaced rank-1 grid: the general kernel does the sampling
  const double ell = a.ls[pl], s2 = a.var[pl];
  const double amp = sqrt(2.0 * s2 / (double)B), inv_ell = 1.0 / ell;
  const double t0 = meta[1], dt = meta[2], z0 = meta[3], dz = meta[4];
  const double* om = a.omega + (size_t)pl * B * D;
  const double* ta = a.tau + (size_t)pl * B;
  const double* wp = a.w + (size_t)pl * S * B;

  // phase-1 roles: lane = basis of the tile; warps 0-5 walk chunks of the query grid, warps 6-7 the two halves of Z
  const int per = (Nq + 5) / 6, mhalf = (M + 1) / 2;
  // phase-2 roles: warp owns PT (feature, point tile) units; MMA fragment coordinates g = lane / 4, t4 = lane % 4
  const int g = lane >> 2, t4 = lane & 3;

  // tables of tile `tile` (operands staged in buffer ob): every warp evaluates its chunk's start phasor, warps 0-3 one
  // more sincos each (the two step phasors, the two endpoint phasors).  Runs inside the previous tile's contraction
  // phase, where its latency hides behind the DMMAs of the same warp.
  auto fill_tables = [&](int tile, int ob) {
    const double* os = osum + (size_t)ob * 4 * kDB + lane;
    const double c = (os[0] + os[kDB] + os[2 * kDB] + os[3 * kDB]) * inv_ell, tau = tb[ob * kDB + lane];
    const double ab = (tile * kDB + lane < B) ? amp : 0.0;
    const double tn = warp < 6 ? t0 + dt * (warp * per) : z0 + dz * (warp == 6 ? 0 : mhalf);
    double sn, cs;
    sincos(tn * c + tau, &sn, &cs);
    double* d0 = stt + ((size_t)warp * kDB + lane) * 2;
    d0[0] = ab * cs; d0[1] = ab * sn;
    if (warp < 4) {
      const double arg = warp == 0 ? dt * c : (warp == 1 ? dz * c : (warp == 2 ? tau : c + tau));
      sincos(arg, &sn, &cs);
      double* d1 = (warp < 2 ? stp : ept) + ((size_t)(warp & 1) * kDB + lane) * 2;
      d1[0] = cs; d1[1] = sn;
    }
  };

  for (int s0 = s_begin; s0 < s_end; s0 += kST) {
    const int ns = min(kST, s_end - s0);
    double acc[PT][2];
#pragma unroll
    for (int j = 0; j < PT; ++j) acc[j][0] = acc[j][1] = 0.0;

    // staging registers: this thread's share of a tile's operands (warps 0-3: omega[base = lane][warp] and [warp + 4]; tid < 32: tau; all: one weight), one tile ahead of the
    // shared-memory copy, which itself is one tile ahead of its use
    double st_o[2] = {0.0, 0.0}, st_t = 0.0, st_w = 0.0;
    auto prefetch = [&](int b0) {
      const int nb = min(kDB, B - b0);
      if (warp < 4) {
        st_o[0] = (lane < nb && warp < D) ? om[(size_t)(b0 + lane) * D + warp] : 0.0;
        st_o[1] = (lane < nb && warp + 4 < D) ? om[(size_t)(b0 + lane) * D + warp + 4] : 0.0;
      }
      st_t = (tid < kDB && tid < nb) ? ta[b0 + tid] : 0.0;
      const int i = tid / kDB, b = tid % kDB;         // kDB * kST = 256 = blockDim.x
      st_w = (i < ns && b < nb) ? wp[(size_t)(s0 + i) * B + b0 + b] : 0.0;
    };
    auto stage = [&](int ob) {
      if (warp < 4) osum[((size_t)ob * 4 + warp) * kDB + lane] = st_o[0] + st_o[1];
      if (tid < kDB) tb[ob * kDB + tid] = st_t;
      Wt[((size_t)ob * kDB + tid % kDB) * kWS + tid / kDB] = st_w;
    };
    __syncthreads();          // previous sample tile's tail is done with the shared memory
    prefetch(0);
    stage(0);
    if (kDB < B) prefetch(kDB);
    __syncthreads();
    fill_tables(0, 0);

    for (int t = 0; t < T; ++t) {
      const int cur = t % 3, nxt = (t + 1) % 3;
      if (t + 1 < T) {
        stage(nxt);
        if ((t + 2) * kDB < B) prefetch((t + 2) * kDB);
      }
      __syncthreads();  // staging and tables visible; previous tile's contraction finished -> feature tile is free
      {  // phase 1: features by rotation along the grid (no transcendental on this path)
        const double* os = osum + (size_t)cur * 4 * kDB + lane;
        const double c = (os[0] + os[kDB] + os[2 * kDB] + os[3 * kDB]) * inv_ell;
        double* fc = feat + lane;
        double* fd = fc + PLANE;
        const double2 st = *reinterpret_cast<const double2*>(stt + ((size_t)warp * kDB + lane) * 2);
        const double2 sp = *reinterpret_cast<const double2*>(stp + ((size_t)(warp < 6 ? 0 : 1) * kDB + lane) * 2);
        if (warp < 6) {
          const int n0 = warp * per;
          rotate_run_pm(fc, fd, 0, n0, min(Nq, n0 + per) - n0, t0, dt, c, inv_ell, st.x, st.y, sp.x, sp.y);
        } else {
          const double2 e = *reinterpret_cast<const double2*>(ept + ((size_t)(warp - 6) * kDB + lane) * 2);
          const double ab = (t * kDB + lane < B) ? amp : 0.0;
          const int col = Nq + (warp - 6);
          fc[(size_t)col * kDBP] = ab * e.x;                                   // Zy[0] = 0, Zy[1] = 1
          fd[(size_t)col * kDBP] = warp == 6 ? 0.0 : ab * e.y * c * inv_ell;
          if (warp == 6) rotate_run_pm(fc, fd, Nq + 2, 0, mhalf, z0, dz, c, inv_ell, st.x, st.y, sp.x, sp.y);
          else rotate_run_pm(fc, fd, Nq + 2, mhalf, M - mhalf, z0, dz, c, inv_ell, st.x, st.y, sp.x, sp.y);
        }
      }
      __syncthreads();
      {  // phase 2: DMMA contraction, and the next tile's tables in its shadow
        if (t + 1 < T) fill_tables(t + 1, nxt);
        const double* wsrc = Wt + (size_t)cur * kDB * kWS + (size_t)t4 * kWS + g;
        const double* asrc[PT];
#pragma unroll
        for (int j = 0; j < PT; ++j) {
          const int u = warp * PT + j, f = u / (4 * PT), tile = u % (4 * PT);
          asrc[j] = feat + (size_t)f * PLANE + (size_t)(tile * 8 + g) * kDBP + t4;
        }
#pragma unroll
        for (int k = 0; k < kDB / 4; ++k) {
          const double wk = wsrc[(size_t)4 * k * kWS];
#pragma unroll
          for (int j = 0; j < PT; ++j) dmma884(acc[j][0], acc[j][1], asrc[j][4 * k], wk);
        }
      }
    }
    __syncthreads();
    // publish red[feature][sample][point] over the dead feature tile
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      const int u = warp * PT + j, f = u / (4 * PT), tile = u % (4 * PT);
      red[((size_t)f * kST + 2 * t4) * ROWS + tile * 8 + g] = acc[j][0];
      red[((size_t)f * kST + 2 * t4 + 1) * ROWS + tile * 8 + g] = acc[j][1];
    }
    __syncthreads();
    for (int idx = tid; idx < 2 * kST * ROWS; idx += nt) {
      const int wh = idx / (kST * ROWS), i = (idx / ROWS) % kST, xx = idx % ROWS;
      if (i < ns && xx < A) {
        double* dst = wh == 0 ? a.f0 : a.h0;
        if (dst != nullptr) dst[((size_t)pl * S + s0 + i) * A + xx] = red[idx];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Equispaced sampler, register-resident features (N + 2 and M each padded to tiles of 8 rows, <= 12 tiles in total).
// The DMMA A fragment of thread (g = lane/4, t = lane%4) is the feature of row g at basis t of a 4-basis step - exactly
// the state of one rotation chain.  So a consumer warp keeps the chain of (row g, basis t) in registers, walks it from
// point tile to point tile with the 8-step phasor E^8, and feeds the cos / sin registers straight into two DMMAs per
// tile: the feature matrix never exists in memory, the main loop has no CTA-wide barrier and no shared-memory feature
// traffic, and one warp alone keeps its sub-partition's FP64 pipe busy (32 pipe cycles of DMMA hide the 2-level
// rotation dependency).  d f0 / d lengthscale uses the separable form  h0[s,x] = t_x sum_b sin(theta_xb) (w_sb c_b / l):
// same chain registers, weights scaled once per step, rows scaled by their coordinate at the end.
//   warps 0-3  consumers: 4-basis steps round-robin; accumulators C[12 tiles][cos|sin][2] for the whole basis loop;
//   warps 4-7  producers: lane = basis of a 32-basis slot; 6 branch-free sincos (start, step of both grids, the two
//              conditioned endpoints), then the 8 row starts S E^g and E^8 by complex products; 42 doubles per basis
//              into a 4-slot ring handed over with mbarriers (one arrival per warp).
// ---------------------------------------------------------------------------------------------
constexpr int kRB = 32;        // bases per table slot
constexpr int kRE = 52;        // doubles per basis in a slot: Sx[8] | E8x | Sz[8] | E8z | e0 | e1 | c/l | pad | w[8 samples] | pad
                               // (kRE % 16 == 4: the 4 bases x 8 rows a consumer warp reads at once hit distinct banks)
constexpr int kRS = 4;         // ring slots
constexpr int kRT = 12;        // point tiles (rows / 8) a consumer carries

// JX = tiles of 8 rows holding the query grid and the two conditioned endpoints (inducing rows follow); GEN = lazy draws
template <int JX, bool GEN>
__global__ void __launch_bounds__(256, 2) pathwise_rr_kernel(PathwiseArgs a, const double* __restrict__ meta) {
  constexpr int kRC = 4;               // consumer warps (4 + 4 in both modes: the fold order, hence every bit, is shared)
  constexpr int kRP = 8 - kRC;         // producer warps
  extern __shared__ __align__(16) double sm[];
  const int D = a.D, M = a.M, Mp = M + 2, Nq = a.Nq, S = a.S, B = a.B, A = Nq + Mp;
  const int pl = blockIdx.x / a.nchunk;
  const int s_begin = (blockIdx.x % a.nchunk) * a.chunk, s_end = min(a.S, s_begin + a.chunk);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x;
  const int T = (B + kRB - 1) / kRB;            // table slots to produce per sample tile
  constexpr int ROWS = kRT * 8;
  constexpr size_t kFold = (size_t)kRC * 2 * kST * ROWS, kRing = (size_t)kRS * kRB * kRE;
  double* tab = sm;                             // [kRS][kRB][kRE]
  double* red = sm;                             // after the basis loop: [kRC][2][kST][ROWS]
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (kFold > kRing ? kFold : kRing));   // [kRS] slot filled
  uint64_t* empty = full + kRS;                                                          // [kRS] slot drained

  if (meta[0] == 0.0) return;  // not an equispaced rank-1 grid: the general kernel does the sampling
  if (GEN && a.iter_dev != nullptr) a.iteration += *a.iter_dev;
  if (tid == 0) {
    for (int i = 0; i < kRS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, kRC); }   // one arrival per warp
  }
  const double ell = a.ls[pl], s2 = a.var[pl];
  const double amp = sqrt(2.0 * s2 / (double)B), inv_ell = 1.0 / ell;
  const double t0 = meta[1], dt = meta[2], z0 = meta[3], dz = meta[4];
  const double* om = a.omega + (size_t)pl * B * D;
  const double* ta = a.tau + (size_t)pl * B;
  const double* wp = a.w + (size_t)pl * S * B;
  const int g = lane >> 2, t4 = lane & 3;

  int it = 0;                                   // sample-tile iteration: slot sequence numbers keep counting across tiles
  for (int s0 = s_begin; s0 < s_end; s0 += kST, ++it) {
    const int ns = min(kST, s_end - s0);
    __syncthreads();   // barriers initialised / previous sample tile's fold is done with the shared memory
    if (warp >= kRC) {
      // ---------------- producer: slots n = pw, pw + kRP, ... ----------------
      const int pw = warp - kRC;
      // operands of slot n are loaded one slot ahead of their use; `ld.volatile`-style asm pins the loads where they
      // are written (the compiler otherwise sinks them next to their first use and the warp eats the DRAM latency)
      double ov[VGPMP_MAX_DOF], wv[kST], tau = 0.0;
      auto fetch = [&](int n) {
        const int b = n * kRB + lane;
        const bool ok = n < T && b < B;
#pragma unroll
        for (int d = 0; d < VGPMP_MAX_DOF; ++d) {
          ov[d] = 0.0;
          if (ok && d < D) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(ov[d]) : "l"(om + (size_t)b * D + d));
        }
        tau = 0.0;
        if (ok) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(tau) : "l"(ta + b));
        // this basis' weights of the sample tile: 32 consecutive bases per row = one coalesced 256-byte request per sample
        // (the consumers used to gather them 8 sectors at a time, one step ahead, and sat on the DRAM latency)
#pragma unroll
        for (int i = 0; i < kST; ++i) {
          wv[i] = 0.0;
          if (ok && i < ns) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(wv[i]) : "l"(wp + (size_t)(s0 + i) * B + b));
        }
      };
      // lazy draws: the same Philox keys and arithmetic as rng_fill_kernel, evaluated here instead of read from memory
      constexpr bool gen = GEN;
      const uint64_t pairkey = ((uint64_t)(pl / D) + (uint64_t)a.problem_offset) * (uint64_t)D + (uint64_t)(pl % D);
      const uint32_t B4 = ((uint32_t)B + 3) / 4;
      if (!gen) fetch(pw);
      for (int n = pw; n < T; n += kRP) {
        const int N = it * T + n, slot = N % kRS, use = N / kRS;
        const bool live = n * kRB + lane < B;
        double c = 0.0, taub = 0.0;
        double wcur[kST];       // load mode: w[sample][this basis]; lazy mode: [0..3] | [4..7] = two Philox blocks (see below)
        if (gen) {
          if (live) {
            const uint64_t key = pairkey * (uint64_t)B + (uint64_t)(n * kRB + lane);
            double z[16];
            const int ncall = (5 + D + 3) / 4;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < ncall) normal4(a.rk, a.iteration, 1u, key * 4 + k, z + 4 * k);
            const double gam = (z[0] * z[0] + z[1] * z[1] + z[2] * z[2] + z[3] * z[3] + z[4] * z[4]) / 5.0;
            const double rs = rsqrt(gam);
#pragma unroll
            for (int d = 0; d < VGPMP_MAX_DOF; ++d)
              if (d < D) c = __dadd_rn(c, __dmul_rn(z[5 + d], rs));   // as stored then summed by the load path: no FMA
            uint32_t cc[4] = {(uint32_t)key, (uint32_t)(key >> 32), (uint32_t)a.iteration, 3u};
            philox4x32(cc, a.rk);
            taub = 6.283185307179586476925 * u01(cc[0], cc[1]);
          }
          // weights: lane (quad, r) draws the two blocks (quad of 4 bases, samples r and r + 4) and later stores each
          // component into the row of the basis it belongs to - no shuffles, 64 blocks per slot in total
          const uint32_t b4 = (uint32_t)n * (kRB / 4) + (uint32_t)(lane >> 2);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int i = (lane & 3) + 4 * hh;
            double z4[4] = {0.0, 0.0, 0.0, 0.0};
            if (i < ns && 4 * b4 < (uint32_t)B) {
              const uint64_t sg = (uint64_t)(s0 + i) + (uint64_t)a.sample_offset;
              normal4(a.rk, a.iteration, 4u, (pairkey * (uint64_t)(1u << 24) + sg) * B4 + b4, z4);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) wcur[4 * hh + k] = z4[k];
          }
        } else {
#pragma unroll
          for (int d = 0; d < VGPMP_MAX_DOF; ++d) c += ov[d];
          taub = tau;
#pragma unroll
          for (int i = 0; i < kST; ++i) wcur[i] = wv[i];
          fetch(n + kRP);
        }
        const double cb = c * inv_ell;
        const double ab = live ? amp : 0.0;
        const double ax0 = t0 * cb + taub, ax1 = dt * cb, az0 = z0 * cb + taub, az1 = dz * cb, ae1 = cb + taub;
        const double arg[6] = {ax0, ax1, az0, az1, taub, ae1};
        double sv[6], cv[6];
        const double big = fmax(fmax(fabs(ax0), fabs(ax1)), fmax(fmax(fabs(az0), fabs(az1)), fmax(fabs(taub), fabs(ae1))));
        if (big < 1048576.0) {
          sincos_bf6(arg, sv, cv);
        } else {
#pragma unroll
          for (int k = 0; k < 6; ++k) sincos(arg[k], &sv[k], &cv[k]);
        }
        double sx = sv[0], cx = cv[0], sdx = sv[1], cdx = cv[1], sz = sv[2], cz = cv[2], sdz = sv[3], cdz = cv[3];
        const double se0 = sv[4], ce0 = cv[4], se1 = sv[5], ce1 = cv[5];
        mbar_wait(empty + slot, (use & 1) ^ 1);            // consumers are done with the slot's previous contents
        double* e = tab + ((size_t)slot * kRB + lane) * kRE;
        double c1 = ab * cx, s1 = ab * sx, c2 = ab * cz, s2z = ab * sz;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          *reinterpret_cast<double2*>(e + 2 * r) = make_double2(c1, s1);
          *reinterpret_cast<double2*>(e + 18 + 2 * r) = make_double2(c2, s2z);
          const double n1 = c1 * cdx - s1 * sdx, n2 = c2 * cdz - s2z * sdz;
          s1 = s1 * cdx + c1 * sdx; s2z = s2z * cdz + c2 * sdz;
          c1 = n1; c2 = n2;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {   // E^8 by three squarings
          const double n1 = cdx * cdx - sdx * sdx, n2 = cdz * cdz - sdz * sdz;
          sdx = 2.0 * cdx * sdx; sdz = 2.0 * cdz * sdz;
          cdx = n1; cdz = n2;
        }
        *reinterpret_cast<double2*>(e + 16) = make_double2(cdx, sdx);
        *reinterpret_cast<double2*>(e + 34) = make_double2(cdz, sdz);
        *reinterpret_cast<double2*>(e + 36) = make_double2(ab * ce0, ab * se0);
        *reinterpret_cast<double2*>(e + 38) = make_double2(ab * ce1, ab * se1);
        *reinterpret_cast<double2*>(e + 40) = make_double2(cb * inv_ell, 0.0);
        *reinterpret_cast<double2*>(e + 50) = make_double2(0.0, 0.0);     // the consumers' zero pair (padding rows)
        if (gen) {
          double* eq = tab + ((size_t)slot * kRB + (lane & ~3)) * kRE + 42 + (lane & 3);   // row of basis 4*quad, column of sample r
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const bool okb = n * kRB + (lane & ~3) + k < B;
              eq[(size_t)k * kRE + 4 * hh] = okb ? wcur[4 * hh + k] : 0.0;
            }
        } else {
#pragma unroll
          for (int i = 0; i < kST; i += 2) *reinterpret_cast<double2*>(e + 42 + i) = make_double2(wcur[i], wcur[i + 1]);
        }
        __syncwarp();                                      // all 32 bases written ...
        if (lane == 0) mbar_arrive(full + slot);           // ... one release for the warp (32 lanes arriving on one
                                                           // mbarrier serialise: that was most of the kernel's time)
      }
    } else {
      // ---------------- consumer: 4-basis steps q = warp, warp + kRC, ... ----------------
      double acc[kRT][2][2];
#pragma unroll
      for (int j = 0; j < kRT; ++j) acc[j][0][0] = acc[j][0][1] = acc[j][1][0] = acc[j][1][1] = 0.0;
      // rows Nq and Nq+1 (the conditioned endpoints) live in tile JX-1, or JX-2 and JX-1.  This lane's role there is loop
      // invariant: -1 = its chain value, else the offset of the (cos, sin) pair to use instead - endpoint 0 / 1 at 36 / 38,
      // padding rows at 50 (a pair the producers keep at zero).  (Deciding this inside the basis loop cost ~35 of the ~165
      // instructions of a step; the slot / quad index arithmetic below another ~20.)
      const int ea = 8 * (JX - 2) + g - Nq, eb = 8 * (JX - 1) + g - Nq;
      const int offa = (JX >= 2 && ea >= 0) ? (ea < 2 ? 36 + 2 * ea : 50) : -1;
      const int offb = eb >= 0 ? (eb < 2 ? 36 + 2 * eb : 50) : -1;
      unsigned roles = (offa >= 0 ? (unsigned)offa : 63u) | ((offb >= 0 ? (unsigned)offb : 63u) << 6);   // 63 = chain value
      asm volatile("" : "+r"(roles));      // one opaque register: otherwise the compiler re-derives both in every step
      const double* ebase = tab + (size_t)(warp * 4 + t4) * kRE;            // this warp's first quad of a slot, this lane's basis
      for (int n = 0; n < T; ++n) {
        const int N = it * T + n, slot = N % kRS;
        mbar_wait(full + slot, (N / kRS) & 1);                             // this slot is filled (acquire)
        const double* e = ebase + (size_t)slot * kRB * kRE;
#pragma unroll 1
        for (int hq = 0; hq < (kRB / 4) / kRC; ++hq, e += (size_t)kRC * 4 * kRE) {   // quads warp, warp + kRC of the slot
          const double wk = e[42 + g];                                        // w[sample g][basis], staged by the producer
          // Chain state: (cos, sin) of row g (amplitude in), advanced from tile to tile by the three-term recurrence
          //   x_{j+1} = 2 cos(theta8) x_j - x_{j-1}      (theta8 = phase step of 8 rows)
          // - one DFMA per value instead of the DMUL + DFMA of a complex product, and a one-deep dependency in front of the
          // next tile's DMMA.  x_{-1} comes from one inverse rotation per grid.  Rounding errors grow like sin(m theta)/sin(theta)
          // <= m <= 8 and the rounding of 2 cos(theta8) shifts the phase by <= 8 eps / theta8: far below the 1e-7 the samples
          // are checked to, for any step a Student-t frequency produces.
          double2 ph = *reinterpret_cast<const double2*>(e + 2 * g);
          double2 st = *reinterpret_cast<const double2*>(e + 16);             // E^8 of the query grid
          double2 pm = make_double2(ph.x * st.x + ph.y * st.y, ph.y * st.x - ph.x * st.y);   // x_{-1} = x_0 E^-8
          double k2 = 2.0 * st.x;
          const double wq = wk * e[40];                                       // w c / l for the d/dlengthscale contraction
#pragma unroll
          for (int j = 0; j < kRT; ++j) {
            if (j == JX) {                                                    // inducing grid starts here
              ph = *reinterpret_cast<const double2*>(e + 18 + 2 * g);
              st = *reinterpret_cast<const double2*>(e + 34);
              pm = make_double2(ph.x * st.x + ph.y * st.y, ph.y * st.x - ph.x * st.y);
              k2 = 2.0 * st.x;
            }
            double ac = ph.x, as = ph.y;
            if (j == JX - 2 && JX >= 2) {
              const unsigned o = roles & 63u;
              if (o != 63u) { const double2 ep = *reinterpret_cast<const double2*>(e + o); ac = ep.x; as = ep.y; }
            }
            if (j == JX - 1) {
              const unsigned o = roles >> 6;
              if (o != 63u) { const double2 ep = *reinterpret_cast<const double2*>(e + o); ac = ep.x; as = ep.y; }
            }
            dmma884(acc[j][0][0], acc[j][0][1], ac, wk);
            dmma884(acc[j][1][0], acc[j][1][1], as, wq);
            const double nc = fma(k2, ph.x, -pm.x), ns_ = fma(k2, ph.y, -pm.y);
            pm = ph;
            ph = make_double2(nc, ns_);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);                          // done with the slot
      }
      // the ring is dead only after ALL consumers left the loop: consumer-only barrier, then the partial sums of this
      // warp go to red[warp][feature][sample][row] over it
      asm volatile("bar.sync 1, %0;" ::"r"(kRC * 32) : "memory");
#pragma unroll
      for (int j = 0; j < kRT; ++j)
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2)
            red[(((size_t)warp * 2 + f) * kST + 2 * t4 + e2) * ROWS + j * 8 + g] = acc[j][f][e2];
    }
    __syncthreads();
    // fold the consumer slices in fixed order; rows -> points: query rows x < Nq, endpoints Nq, Nq+1, inducing rows after
    for (int idx = tid; idx < 2 * kST * ROWS; idx += nt) {
      const int wh = idx / (kST * ROWS), i = (idx / ROWS) % kST, row = idx % ROWS;
      double tsum = red[idx];
      for (int k = 1; k < kRC; ++k) tsum += red[(size_t)k * 2 * kST * ROWS + idx];
      int point = -1;
      double coord = 0.0;
      if (row < JX * 8) {
        if (row < Nq) { point = row; coord = t0 + dt * row; }
        else if (row < Nq + 2) { point = row; coord = (double)(row - Nq); }
      } else {
        const int m = row - JX * 8;
        if (m < M) { point = Nq + 2 + m; coord = z0 + dz * m; }
      }
      if (i < ns && point >= 0) {
        if (wh == 0) { if (a.f0 != nullptr) a.f0[((size_t)pl * S + s0 + i) * A + point] = tsum; }
        else if (a.h0 != nullptr) a.h0[((size_t)pl * S + s0 + i) * A + point] = tsum * coord;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Register-resident sampler for 9 .. 63 samples (e.g. the reference's S = 20 planner_params of Kuka / WAM / Franka-industrial):
// NT tiles of 8 samples per CTA pass.  pathwise_rr_kernel above repeats the whole producer pass (Philox, 6 sincos per basis)
// for every 8-sample tile and issues 2 DMMAs per rotation; here a consumer warp owns a QUARTER OF THE POINT TILES (3 of 12)
// for ALL NT sample tiles and walks every 4-basis step: one rotation feeds 2 NT DMMAs, accumulators are 12 NT registers,
// the basis tables are produced once per 8 NT samples, and no cross-warp fold is needed (a warp's tiles are its own).
// A warp that does not start at row 0 of a grid multiplies the row starts by the group phasor E8^(tile offset) the
// producers add to the slot.  Slot entry per basis: Sx[8] | E8x | Sz[8] | E8z | e0 | e1 | c/l^2 | pad | P1 P2 P3 | w[8 NT].
// ---------------------------------------------------------------------------------------------
constexpr int kRMW = 48;       // first weight of a slot entry

template <int NT, bool GEN>
__global__ void __launch_bounds__(256, 2) pathwise_rrm_kernel(PathwiseArgs a, const double* __restrict__ meta) {
  constexpr int kRC = 4, kRP = 4;                // consumer / producer warps
  constexpr int kTPW = kRT / kRC;                // point tiles per consumer warp
  constexpr int kEM = (kRMW + 8 * NT + 11) / 16 * 16 + 4;   // doubles per basis in a slot, padded to 4 mod 16: the 4 bases x 8
                                                           // rows a consumer warp reads at once then hit distinct banks
  constexpr int kSamples = 8 * NT;
  extern __shared__ __align__(16) double sm[];
  const int D = a.D, M = a.M, Nq = a.Nq, S = a.S, B = a.B, A = Nq + M + 2;
  const int JX = (Nq + 2 + 7) / 8;               // tiles holding the query grid and the two conditioned endpoints
  const int tpw = (JX + (M + 7) / 8 + kRC - 1) / kRC;   // tiles per consumer warp actually in use (<= kTPW): the useful tiles, evenly
  const int pl = blockIdx.x / a.nchunk;
  const int s0 = (blockIdx.x % a.nchunk) * a.chunk, ns = min(kSamples, S - s0);   // a.chunk == 8 NT
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt_ = blockDim.x;
  const int T = (B + kRB - 1) / kRB;
  constexpr int ROWS = kRT * 8;
  constexpr size_t kFold = (size_t)2 * kSamples * ROWS, kRing = (size_t)kRS * kRB * kEM;
  double* tab = sm;                              // [kRS][kRB][kEM]
  double* red = sm;                              // after the basis loop: [2][8 NT][ROWS]
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (kFold > kRing ? kFold : kRing));
  uint64_t* empty = full + kRS;

  if (meta[0] == 0.0) return;  // not an equispaced rank-1 grid: the general kernel does the sampling
  if (GEN && a.iter_dev != nullptr) a.iteration += *a.iter_dev;
  if (tid == 0) {
    for (int i = 0; i < kRS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, kRC); }
  }
  const double ell = a.ls[pl], s2 = a.var[pl];
  const double amp = sqrt(2.0 * s2 / (double)B), inv_ell = 1.0 / ell;
  const double t0 = meta[1], dt = meta[2], z0 = meta[3], dz = meta[4];
  const int g = lane >> 2, t4 = lane & 3;
  __syncthreads();
  if (warp >= kRC) {
    // ---------------- producer: slots n = pw, pw + kRP, ... ----------------
    const int pw = warp - kRC;
    const double* om = GEN ? nullptr : a.omega + (size_t)pl * B * D;
    const double* ta = GEN ? nullptr : a.tau + (size_t)pl * B;
    const double* wp = GEN ? nullptr : a.w + (size_t)pl * S * B;
    const uint64_t pairkey = ((uint64_t)(pl / D) + (uint64_t)a.problem_offset) * (uint64_t)D + (uint64_t)(pl % D);
    const uint32_t B4 = ((uint32_t)B + 3) / 4;
    for (int n = pw; n < T; n += kRP) {
      const int slot = n % kRS, use = n / kRS;
      const int b = n * kRB + lane;
      const bool live = b < B;
      double c = 0.0, taub = 0.0;
      if (GEN) {
        if (live) {
          const uint64_t key = pairkey * (uint64_t)B + (uint64_t)b;
          double z[16];
          const int ncall = (5 + D + 3) / 4;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < ncall) normal4(a.rk, a.iteration, 1u, key * 4 + k, z + 4 * k);
          const double gam = (z[0] * z[0] + z[1] * z[1] + z[2] * z[2] + z[3] * z[3] + z[4] * z[4]) / 5.0;
          const double rs = rsqrt(gam);
#pragma unroll
          for (int d = 0; d < VGPMP_MAX_DOF; ++d)
            if (d < D) c = __dadd_rn(c, __dmul_rn(z[5 + d], rs));   // as stored then summed by the load path: no FMA
          uint32_t cc[4] = {(uint32_t)key, (uint32_t)(key >> 32), (uint32_t)a.iteration, 3u};
          philox4x32(cc, a.rk);
          taub = 6.283185307179586476925 * u01(cc[0], cc[1]);
        }
      } else if (live) {
        for (int d = 0; d < D; ++d) c += __ldg(om + (size_t)b * D + d);
        taub = __ldg(ta + b);
      }
      const double cb = c * inv_ell;
      const double ab = live ? amp : 0.0;
      const double ax0 = t0 * cb + taub, ax1 = dt * cb, az0 = z0 * cb + taub, az1 = dz * cb, ae1 = cb + taub;
      const double arg[6] = {ax0, ax1, az0, az1, taub, ae1};
      double sv[6], cv[6];
      const double big = fmax(fmax(fabs(ax0), fabs(ax1)), fmax(fmax(fabs(az0), fabs(az1)), fmax(fabs(taub), fabs(ae1))));
      if (big < 1048576.0) {
        sincos_bf6(arg, sv, cv);
      } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) sincos(arg[k], &sv[k], &cv[k]);
      }
      double sdx = sv[1], cdx = cv[1], sdz = sv[3], cdz = cv[3];
      mbar_wait(empty + slot, (use & 1) ^ 1);            // consumers are done with the slot's previous contents
      double* e = tab + ((size_t)slot * kRB + lane) * kEM;
      double c1 = ab * cv[0], s1 = ab * sv[0], c2 = ab * cv[2], s2z = ab * sv[2];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        *reinterpret_cast<double2*>(e + 2 * r) = make_double2(c1, s1);
        *reinterpret_cast<double2*>(e + 18 + 2 * r) = make_double2(c2, s2z);
        const double n1 = c1 * cdx - s1 * sdx, n2 = c2 * cdz - s2z * sdz;
        s1 = s1 * cdx + c1 * sdx; s2z = s2z * cdz + c2 * sdz;
        c1 = n1; c2 = n2;
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {   // E^8 by three squarings
        const double n1 = cdx * cdx - sdx * sdx, n2 = cdz * cdz - sdz * sdz;
        sdx = 2.0 * cdx * sdx; sdz = 2.0 * cdz * sdz;
        cdx = n1; cdz = n2;
      }
      *reinterpret_cast<double2*>(e + 16) = make_double2(cdx, sdx);
      *reinterpret_cast<double2*>(e + 34) = make_double2(cdz, sdz);
      *reinterpret_cast<double2*>(e + 36) = make_double2(ab * cv[4], ab * sv[4]);
      *reinterpret_cast<double2*>(e + 38) = make_double2(ab * cv[5], ab * sv[5]);
      *reinterpret_cast<double2*>(e + 40) = make_double2(cb * inv_ell, 0.0);
      *reinterpret_cast<double2*>(e + kRMW + 8 * NT) = make_double2(0.0, 0.0);     // the consumers' zero pair
      // group phasors: consumer warp w starts at tile 3 w, i.e. E8x^(3w) on the query grid or E8z^(3w - JX) on the inducing grid
      {
        double px = 1.0, qx = 0.0, pz = 1.0, qz = 0.0;
        int jx = 0, jz = 0;                              // powers reached so far
#pragma unroll
        for (int w = 1; w < kRC; ++w) {
          const int j0 = w * tpw;
          double oc, os;
          if (j0 < JX) {
            while (jx < j0) { const double nn = px * cdx - qx * sdx; qx = qx * cdx + px * sdx; px = nn; ++jx; }
            oc = px; os = qx;
          } else {
            while (jz < j0 - JX) { const double nn = pz * cdz - qz * sdz; qz = qz * cdz + pz * sdz; pz = nn; ++jz; }
            oc = pz; os = qz;
          }
          *reinterpret_cast<double2*>(e + 40 + 2 * w) = make_double2(oc, os);
        }
      }
      // weights of the 8 NT samples
      if (GEN) {
        const uint32_t b4 = (uint32_t)n * (kRB / 4) + (uint32_t)(lane >> 2);
        double* eq = tab + ((size_t)slot * kRB + (lane & ~3)) * kEM + kRMW + (lane & 3);   // row of basis 4*quad, column of sample r
#pragma unroll
        for (int hh = 0; hh < 2 * NT; ++hh) {
          const int i = (lane & 3) + 4 * hh;
          double z4[4] = {0.0, 0.0, 0.0, 0.0};
          if (i < ns && 4 * b4 < (uint32_t)B) {
            const uint64_t sg = (uint64_t)(s0 + i) + (uint64_t)a.sample_offset;
            normal4(a.rk, a.iteration, 4u, (pairkey * (uint64_t)(1u << 24) + sg) * B4 + b4, z4);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool okb = n * kRB + (lane & ~3) + k < B;
            eq[(size_t)k * kEM + 4 * hh] = okb ? z4[k] : 0.0;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < kSamples; ++i)
          e[kRMW + i] = (live && i < ns) ? __ldg(wp + (size_t)(s0 + i) * B + b) : 0.0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(full + slot);
    }
  } else {
    // ---------------- consumer: point tiles [3 warp, 3 warp + 3), every 4-basis step ----------------
    double acc[kTPW][2][NT][2];
#pragma unroll
    for (int j = 0; j < kTPW; ++j)
#pragma unroll
      for (int i = 0; i < NT; ++i) acc[j][0][i][0] = acc[j][0][i][1] = acc[j][1][i][0] = acc[j][1][i][1] = 0.0;
    const int j0 = warp * tpw;
    // per-tile roles of this lane's row (loop invariant): 0 chain value, 1 / 2 conditioned endpoint 0 / 1, 3 padding (zero);
    // sw: the inducing grid starts at this tile (reload the row start, no group offset)
    // (packed into one word: 2 bits of role + 1 bit of sw per tile - kept as arrays the compiler re-derived them in every
    // step of the basis loop, 12 % of the kernel's instructions)
    constexpr int kZero = kRMW + 8 * NT;          // a pair the producers keep at zero (padding rows read it)
    static_assert(kZero + 2 <= kEM && kZero < 127, "slot entry has no room for the zero pair");
    unsigned roles = 0;                           // per tile: 7 bits = offset of the pair used instead of the chain value (127: none),
#pragma unroll                                    //           bit 7 = the inducing grid starts at this tile
    for (int jj = 0; jj < kTPW; ++jj) {
      const int j = j0 + jj, ex = 8 * j + g - Nq;
      const unsigned off = (j < JX && ex >= 0) ? (ex < 2 ? 36u + 2u * ex : (unsigned)kZero) : 127u;
      roles |= (off | ((jj > 0 && j == JX) ? 128u : 0u)) << (8 * jj);
    }
    asm volatile("" : "+r"(roles));                // opaque from here on: one register, extracted by shifts
    const bool startx = j0 < JX;
    const int Q = T * (kRB / 4);
    // (fetching a step's operands one step ahead was tried: 3.40 -> 3.88 ms at 1024 Kuka problems, the extra live registers
    // cost more than the exposed shared-memory round trip)
    int cur = -1;
    const double* e = tab;
    for (int q = 0; q < Q; ++q, e += 4 * kEM) {
      if ((q & (kRB / 4 - 1)) == 0) {                 // first step of slot n = q / 8
        const int n = q / (kRB / 4);
        if (cur >= 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(empty + cur % kRS);
        }
        mbar_wait(full + n % kRS, (n / kRS) & 1);
        cur = n;
        e = tab + ((size_t)(n % kRS) * kRB + t4) * kEM;
      }
      double wk[NT], wq[NT];
      const double cl = e[40];
#pragma unroll
      for (int i = 0; i < NT; ++i) { wk[i] = e[kRMW + 8 * i + g]; wq[i] = wk[i] * cl; }
      double2 ph = *reinterpret_cast<const double2*>(e + (startx ? 0 : 18) + 2 * g);
      double2 st = *reinterpret_cast<const double2*>(e + (startx ? 16 : 34));
      if (warp > 0) {
        const double2 po = *reinterpret_cast<const double2*>(e + 40 + 2 * warp);
        const double c2 = ph.x * po.x - ph.y * po.y;
        ph.y = ph.y * po.x + ph.x * po.y;
        ph.x = c2;
      }
#pragma unroll
      for (int jj = 0; jj < kTPW; ++jj) {
        if (jj >= tpw) break;
        const unsigned role = (roles >> (8 * jj)) & 255u;
        if (role & 128u) {
          ph = *reinterpret_cast<const double2*>(e + 18 + 2 * g);
          st = *reinterpret_cast<const double2*>(e + 34);
        }
        double ac = ph.x, as = ph.y;
        const unsigned o = role & 127u;
        if (o != 127u) { const double2 ep = *reinterpret_cast<const double2*>(e + o); ac = ep.x; as = ep.y; }
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          dmma884(acc[jj][0][i][0], acc[jj][0][i][1], ac, wk[i]);
          dmma884(acc[jj][1][i][0], acc[jj][1][i][1], as, wq[i]);
        }
        const double c2 = ph.x * st.x - ph.y * st.y;
        ph.y = ph.y * st.x + ph.x * st.y;
        ph.x = c2;
      }
    }
    __syncwarp();
    if (cur >= 0 && lane == 0) mbar_arrive(empty + cur % kRS);
    // the ring is dead only after ALL consumers left the loop
    asm volatile("bar.sync 1, %0;" ::"r"(kRC * 32) : "memory");
#pragma unroll
    for (int jj = 0; jj < kTPW; ++jj)
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2)
            if (jj < tpw) red[((size_t)f * kSamples + 8 * i + 2 * t4 + e2) * ROWS + (j0 + jj) * 8 + g] = acc[jj][f][i][e2];
  }
  __syncthreads();
  // rows -> points: query rows x < Nq, endpoints Nq, Nq+1, inducing rows after
  for (int idx = tid; idx < 2 * kSamples * ROWS; idx += nt_) {
    const int wh = idx / (kSamples * ROWS), i = (idx / ROWS) % kSamples, row = idx % ROWS;
    int point = -1;
    double coord = 0.0;
    if (row < JX * 8) {
      if (row < Nq) { point = row; coord = t0 + dt * row; }
      else if (row < Nq + 2) { point = row; coord = (double)(row - Nq); }
    } else {
      const int m = row - JX * 8;
      if (m < M) { point = Nq + 2 + m; coord = z0 + dz * m; }
    }
    if (i < ns && point >= 0) {
      if (wh == 0) { if (a.f0 != nullptr) a.f0[((size_t)pl * S + s0 + i) * A + point] = red[idx]; }
      else if (a.h0 != nullptr) a.h0[((size_t)pl * S + s0 + i) * A + point] = red[idx] * coord;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GP preparation fused with the pathwise update (used behind the DMMA sampler, which only needs the draws and the
// lengthscale): one CTA of 128 threads per (problem, latent[, sample chunk]).  gp_prepare_body leaves L, L^-1, q_sqrt_full,
// mu and Zy in shared memory; every warp then finishes whole samples on its own, no CTA barrier between them:
//   u = mu + S eps_u;  r = u - f0(Zy) - sqrt(jitter) eps_j;  v = L^-T L^-1 r;  f = f0(X) + Kfu v.
// Small CTAs, latency hidden by occupancy - the job the 256-thread sampler CTAs did badly in their tails.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 7) gp_prepare_update_kernel(PathwiseArgs a, vgpmp_params P, double* __restrict__ Lc_out,
                                                                  double* __restrict__ S_out, double* __restrict__ kl_l,
                                                                  double* __restrict__ kvec, double* __restrict__ Linv_out,
                                                                  const double* __restrict__ meta) {
  __shared__ __align__(16) double prep[kPrepSmem];    // gp_prepare_body's block: K -> q -> q_sqrt_full | L | L^-1 | zy | mu | ...
  __shared__ double vsm[kST * 32];                    // v of one tile of samples
  __shared__ double ez[2 * 32];                       // exp(+c zy[m]) | exp(-c zy[m]), c = sqrt(5) / lengthscale
  const int D = a.D, M = a.M, Mp = M + 2, Nq = a.Nq, S = a.S, A = Nq + Mp;
  const int pl = blockIdx.x / a.nchunk, chunk = blockIdx.x % a.nchunk, p = pl / D, l = pl % D;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
  const bool first = chunk == 0;                      // chunk 0 publishes the factors for the reverse pass
  gp_prepare_body(D, M, a.jitter, P, pl, first ? Lc_out : nullptr, first ? S_out : nullptr, first ? kl_l : nullptr,
                  first ? kvec : nullptr, first ? Linv_out : nullptr, prep, prep);
  if (meta[0] == 0.0) return;                         // general sampler runs after this kernel and does its own update
  const double* Ssm = prep;
  const double* Lism = prep + 2 * 32 * LDM;
  const double* zy = prep + 3 * 32 * LDM;
  const double* mu = zy + 32;
  const double ell = a.ls[pl], s2 = a.var[pl], sqrtj = sqrt(a.jitter);
  const int s_begin = chunk * a.chunk, s_end = min(S, s_begin + a.chunk);
  // Matern-5/2 in the update below: exp(-c |x - z|) = exp(-c x) exp(c z) for x >= z (and mirrored), so a thread needs two
  // exponentials for its query point instead of one per inducing point (the exp was 15 % of this kernel's instructions).
  // |c x|, |c z| < 300 keeps both factors far from overflow; otherwise the direct form is used.
  const double cexp = VG_SQRT5 / ell;
  if (tid < 32) {
    const double z = tid < Mp ? zy[tid] : 0.0;
    ez[tid] = exp(cexp * z);
    ez[32 + tid] = fabs(cexp * z) < 300.0 ? exp(-cexp * z) : -1.0;     // negative marks "use the direct form"
  }
  for (int s0 = s_begin; s0 < s_end; s0 += kST) {
    const int ns = min(kST, s_end - s0);
    __syncthreads();   // q_sqrt_full complete / previous tile's vsm consumed
    // warp per sample: u = mu + S eps_u;  r = u - f0(Zy) - sqrt(jitter) eps_j;  v = L^-T (L^-1 r)
    for (int i = warp; i < ns; i += nw) {
      const int s = s0 + i;
      double* row = vsm + i * 32;
      row[lane] = lane < Mp ? a.eps_u[((size_t)pl * S + s) * Mp + lane] : 0.0;
      __syncwarp();
      double r = 0.0;
      if (lane < Mp) {
        double u = mu[lane];
        for (int k = 0; k <= lane; ++k) u += Ssm[lane * LDM + k] * row[k];
        r = u - a.f0[((size_t)pl * S + s) * A + Nq + lane] - sqrtj * a.eps_j[((size_t)pl * S + s) * Mp + lane];
      }
      __syncwarp();
      row[lane] = r;
      __syncwarp();
      const double y = warp_lower_mv(Lism, Mp, row);
      __syncwarp();
      row[lane] = y;
      __syncwarp();
      const double vv = warp_lowerT_mv(Lism, Mp, row);
      __syncwarp();
      row[lane] = lane < Mp ? vv : 0.0;
      if (lane < Mp && a.v != nullptr) a.v[((size_t)pl * S + s) * Mp + lane] = vv;
    }
    __syncthreads();
    // thread per query point: f[s][n] = f0[s][n] + sum_m k(x_n, z_m) v[s][m]; each kernel value is formed once and
    // used by every sample of the tile
    for (int n = tid; n < Nq; n += nt) {
      const double xn = a.Xq[(size_t)n * D + l];
      double acc[kST];
#pragma unroll
      for (int i = 0; i < kST; ++i) acc[i] = i < ns ? a.f0[((size_t)pl * S + s0 + i) * A + n] : 0.0;
      const bool sep = fabs(cexp * xn) < 300.0;
      const double exp_p = sep ? exp(cexp * xn) : 0.0, exp_m = sep ? exp(-cexp * xn) : 0.0;
      for (int m = 0; m < Mp; ++m) {
        const double dx = xn - zy[m], av = cexp * fabs(dx);
        double ev;
        if (sep && ez[32 + m] > 0.0) ev = dx >= 0.0 ? exp_m * ez[m] : exp_p * ez[32 + m];
        else ev = exp(-av);
        const double kv = s2 * ((1.0 + av + (1.0 / 3.0) * av * av) * ev);     // (5/3) r^2 = a^2 / 3 with a = sqrt(5) r
#pragma unroll
        for (int i = 0; i < kST; ++i) acc[i] += kv * vsm[i * 32 + m];
      }
#pragma unroll
      for (int i = 0; i < kST; ++i)
        if (i < ns) a.f[f_index(a.f_planar, p, s0 + i, n, l, S, Nq, D)] = acc[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Pathwise update for many samples per (problem, latent) (num_samples >= 64; the tensor-core sampler's companion).
// gp_prepare_kernel has published L^-1 and q_sqrt_full; one CTA of 256 threads per (pair, chunk of samples) stages them and
// Kfu^T in shared memory once, then every warp finishes whole samples on its own - no CTA barrier in the sample loop:
//   u = mu + S eps_u;  r = u - f0(Zy) - sqrt(jitter) eps_j;  v = L^-T L^-1 r;  f = f0(X) + Kfu v.
// (gp_prepare_update_kernel repeats the Cholesky per sample chunk and re-evaluates the Matern kernel per 8-sample tile:
// right for S = 7, 40 ms per step at 8192 problems x 256 samples.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pathwise_update_kernel(PathwiseArgs a, const double* __restrict__ meta, int chunk) {
  extern __shared__ __align__(16) double sm[];
  if (meta[0] == 0.0) return;                         // general sampler does its own update
  const int D = a.D, M = a.M, Mp = M + 2, Nq = a.Nq, S = a.S, A = Nq + Mp;
  const int nchunk = (S + chunk - 1) / chunk;
  const int pl = blockIdx.x / nchunk, p = pl / D, l = pl % D;
  const int s_begin = (blockIdx.x % nchunk) * chunk, s_end = min(S, s_begin + chunk);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
  const int NP = Nq | 1;                              // odd row length of Kfu^T: conflict-free for lanes over n
  double* Li = sm;                                    // [32][LDM]
  double* Ss = Li + 32 * LDM;                         // [32][LDM]
  double* KT = Ss + 32 * LDM;                         // [Mp][NP]   Kfu transposed
  double* mu = KT + (size_t)Mp * NP;                  // [32]
  double* zy = mu + 32;                               // [32]
  double* rows = zy + 32;                             // [nw][32]
  const double ell = a.ls[pl], s2 = a.var[pl], sqrtj = sqrt(a.jitter), inv_ell = 1.0 / ell;
  for (int idx = tid; idx < 32 * LDM; idx += nt) { Li[idx] = 0.0; Ss[idx] = 0.0; }
  if (tid < 32) {
    zy[tid] = tid < Mp ? zy_at(a.Z, D, l, tid) : 0.0;
    mu[tid] = tid >= Mp ? 0.0 : (tid < 2 ? a.query_latent[((size_t)p * 2 + tid) * D + l] : a.q_mu[((size_t)p * M + tid - 2) * D + l]);
  }
  __syncthreads();
  for (int i = warp; i < Mp; i += nw)
    if (lane < Mp) {
      Li[i * LDM + lane] = a.Linv[(size_t)pl * Mp * Mp + i * Mp + lane];
      Ss[i * LDM + lane] = a.Sfull[(size_t)pl * Mp * Mp + i * Mp + lane];
    }
  for (int idx = tid; idx < Mp * Nq; idx += nt) {
    const int m = idx / Nq, n = idx - m * Nq;
    KT[m * NP + n] = s2 * vg_matern52(fabs(a.Xq[(size_t)n * D + l] - zy[m]) * inv_ell);
  }
  __syncthreads();
  double* row = rows + warp * 32;
  for (int s = s_begin + warp; s < s_end; s += nw) {
    const size_t ps = (size_t)pl * S + s;
    row[lane] = lane < Mp ? a.eps_u[ps * Mp + lane] : 0.0;
    const double f0z = lane < Mp ? a.f0[ps * A + Nq + lane] : 0.0;
    const double ej = lane < Mp ? a.eps_j[ps * Mp + lane] : 0.0;
    __syncwarp();
    double r = 0.0;
    if (lane < Mp) {
      double u = mu[lane];
      for (int k = 0; k <= lane; ++k) u += Ss[lane * LDM + k] * row[k];
      r = u - f0z - sqrtj * ej;
    }
    __syncwarp();
    row[lane] = r;
    __syncwarp();
    const double y = warp_lower_mv(Li, Mp, row);
    __syncwarp();
    row[lane] = y;
    __syncwarp();
    const double vv = warp_lowerT_mv(Li, Mp, row);
    __syncwarp();
    row[lane] = lane < Mp ? vv : 0.0;
    if (lane < Mp && a.v != nullptr) a.v[ps * Mp + lane] = vv;
    __syncwarp();
    for (int n = lane; n < Nq; n += 32) {
      double acc = a.f0[ps * A + n];
      for (int m = 0; m < Mp; ++m) acc += KT[m * NP + n] * row[m];
      a.f[f_index(a.f_planar, p, s, n, l, S, Nq, D)] = acc;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Reverse pass of the GP side, one CTA (256 threads) per (problem, latent).
// ---------------------------------------------------------------------------------------------
struct BackwardArgs {
  int D, M, N, S, B;
  int nchunk, chunk;   // sample-loop split (MODE 1/2 of gp_backward_kernel)
  double* partial;     // [Bp*D*nchunk][kPartial]
  double jitter, klw;
  const double *Z, *X, *ls, *var, *q_sqrt;
  const double *eps_u;
  const double *Lc, *Linv, *kvec, *v, *f0, *h0, *df;
  int df_planar;       // layout of df, as PathwiseArgs::f_planar
  double *d_q_mu, *d_q_sqrt, *d_ls, *d_var;
};

constexpr int kBT = 8;  // samples per tile (= warps) in the reverse pass

constexpr int kPartial = 2 * 32 * 32 + 32 + 4;   // doubles per (problem, latent, sample chunk) partial: G, GS, gmu, acc_var, acc_ls

// MODE 0: one CTA per (problem, latent) does everything.  For a single problem with very many samples the sample loop
// is split over CTAs: MODE 1 = sample phase of one chunk -> partial sums in global memory; MODE 2 = fold the partials in
// fixed order (deterministic) and run the sample-independent phase.
template <int MODE>
__global__ void __launch_bounds__(256, 4) gp_backward_kernel(BackwardArgs a) {
  extern __shared__ double sm[];
  const int D = a.D, M = a.M, Mp = M + 2, N = a.N, S = a.S, A = N + Mp;
  const int pl = MODE == 1 ? blockIdx.x / a.nchunk : blockIdx.x, p = pl / D, l = pl % D;
  const int s_begin = MODE == 1 ? (blockIdx.x % a.nchunk) * a.chunk : 0;
  const int s_end = MODE == 1 ? min(S, s_begin + a.chunk) : (MODE == 2 ? 0 : S);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
  // Kfu is only live inside the sample loop, L and d ELBO / d Lc only after it: they share one region (one matrix
  // less resident -> 4 CTAs per SM instead of 3, i.e. 4 waves of the 13 CTAs per SM instead of 5).
  const size_t kreg = (size_t)N * Mp > (size_t)2 * 32 * LDM ? (size_t)N * Mp : (size_t)2 * 32 * LDM;
  double* Lism = sm;                   // [32][LDM] explicit inverse L^-1
  double* G = Lism + 32 * LDM;         // [32][LDM] d ELBO / d Khat (general, not symmetrised)
  double* GS = G + 32 * LDM;           // [32][LDM] d ELBO / d q_sqrt_full
  double* Lsm = GS + 32 * LDM;         // [32][LDM] chol factor L            (after the sample loop)
  double* GL = Lsm + 32 * LDM;         // [32][LDM] d ELBO / d Lc            (after the sample loop)
  double* Kfu = Lsm;                   // [N][Mp]                            (inside the sample loop)
  double* gv = Lsm + kreg;             // [kBT][32]
  double* gr = gv + kBT * 32;          // [kBT][32]
  double* vsm = gr + kBT * 32;         // [kBT][32]
  double* epsm = vsm + kBT * 32;       // [kBT][32] this tile's eps_u
  double* gmu = epsm + kBT * 32;       // [32]
  double* zy = gmu + 32;               // [32]
  double* bvec = zy + 32;              // [32]
  double* gd = bvec + 32;              // [32]
  double* red = gd + 32;               // [8]
  double* dft = red + 8;               // [kBT][N]  this tile's slice of d ELBO / d f

  if (MODE == 0) {
    // Everything this CTA will read is requested from DRAM now (the phases below meet their inputs one dependent round trip
    // at a time otherwise: a third of the stall samples of this latency-bound kernel were first-touch loads).
    auto l2_prefetch = [&](const double* base, size_t count) { l2_prefetch_span(base, count, tid, nt); };
    l2_prefetch(a.Linv + (size_t)pl * Mp * Mp, (size_t)Mp * Mp);
    l2_prefetch(a.Lc + (size_t)pl * Mp * Mp, (size_t)Mp * Mp);
    l2_prefetch(a.q_sqrt + (size_t)pl * M * M, (size_t)M * M);
    l2_prefetch(a.v + (size_t)pl * S * Mp, (size_t)S * Mp);
    l2_prefetch(a.eps_u + (size_t)pl * S * Mp, (size_t)S * Mp);
    l2_prefetch(a.f0 + (size_t)pl * S * A, (size_t)S * A);
    l2_prefetch(a.h0 + (size_t)pl * S * A, (size_t)S * A);
    if (a.df_planar) l2_prefetch(a.df + ((size_t)p * D + l) * S * N, (size_t)S * N);
  }
  const double ell = a.ls[pl], s2 = a.var[pl], inv_ell = 1.0 / ell;
  if (tid < 32) {
    zy[tid] = tid < Mp ? zy_at(a.Z, D, l, tid) : 0.0;
    bvec[tid] = tid < Mp ? a.kvec[(size_t)pl * (Mp + 4) + tid] : 0.0;
    gmu[tid] = 0.0;
  }
  for (int idx = tid; idx < 32 * LDM; idx += nt) { G[idx] = 0.0; GS[idx] = 0.0; Lism[idx] = 0.0; }
  __syncthreads();
  for (int i = warp; i < Mp; i += nw)
    if (lane < Mp) Lism[i * LDM + lane] = a.Linv[(size_t)pl * Mp * Mp + i * Mp + lane];
  if (s_begin < s_end) {
    // Kfu through the separable form of exp(-c |x - z|) (see gp_prepare_update_kernel): 2 N + 2 Mp exponentials per CTA
    // instead of N Mp; the tables borrow the df tile, which is only filled inside the sample loop
    const double cexp = VG_SQRT5 * inv_ell;
    double* exn = dft;                                   // [2][N]  exp(+c x_n) | exp(-c x_n)
    for (int n = tid; n < N; n += nt) {
      const double cx = cexp * a.X[(size_t)n * D + l];
      exn[n] = exp(cx);
      exn[N + n] = fabs(cx) < 300.0 ? exp(-cx) : -1.0;   // negative marks "use the direct form"
    }
    const double cz = cexp * zy[lane & 31];
    const double ezp = exp(cz), ezm = exp(-cz);
    const bool zok = fabs(cz) < 300.0;
    __syncthreads();
    for (int n = warp; n < N; n += nw)
      if (lane < Mp) {
        const double dx = a.X[(size_t)n * D + l] - zy[lane], av = cexp * fabs(dx);
        const double ev = (zok && exn[N + n] > 0.0) ? (dx >= 0.0 ? exn[N + n] * ezp : exn[n] * ezm) : exp(-av);
        Kfu[n * Mp + lane] = s2 * ((1.0 + av + (1.0 / 3.0) * av * av) * ev);
      }
  }
  __syncthreads();

  double acc_ls = 0.0, acc_var = 0.0;  // per-thread partial hyper-parameter gradients
  // df[s,n] at dfp[(s*N+n)*dfs]
  const double* dfp = a.df + (size_t)p * S * N * D + (a.df_planar ? (size_t)l * S * N : (size_t)l);
  const int dfs = a.df_planar ? 1 : D;

  for (int s0 = s_begin; s0 < s_end; s0 += kBT) {
    const int ns = min(kBT, s_end - s0);
    // stage this tile's df (strided in global: [s,n,D]), v and eps_u in shared memory
    for (int idx = tid; idx < kBT * N; idx += nt) {
      const int i = idx / N, n = idx - i * N;
      dft[idx] = i < ns ? dfp[((size_t)(s0 + i) * N + n) * dfs] : 0.0;
    }
    for (int idx = tid; idx < kBT * 32; idx += nt) {
      const int i = idx >> 5, m = idx & 31;
      const bool ok = i < ns && m < Mp;
      vsm[idx] = ok ? a.v[((size_t)pl * S + s0 + i) * Mp + m] : 0.0;
      epsm[idx] = ok ? a.eps_u[((size_t)pl * S + s0 + i) * Mp + m] : 0.0;
    }
    __syncthreads();
    // (1) gv[s,m] = sum_n Kfu[n,m] df[s,n]
    for (int idx = tid; idx < kBT * 32; idx += nt) {
      const int i = idx >> 5, m = idx & 31;
      double t = 0.0;
      if (i < ns && m < Mp) {
        const double* dfs = dft + (size_t)i * N;
#pragma unroll 4
        for (int n = 0; n < N; ++n) t += Kfu[n * Mp + m] * dfs[n];
      }
      gv[idx] = t;
    }
    __syncthreads();
    // (3) gr = Khat^-1 gv = Li^T (Li gv), one warp per sample, dense products with the explicit inverse factor
    for (int i = warp; i < kBT; i += nw) {
      const double y = warp_lower_mv(Lism, Mp, gv + i * 32);
      gr[i * 32 + lane] = y;                      // scratch
      __syncwarp();
      const double z = warp_lowerT_mv(Lism, Mp, gr + i * 32);
      __syncwarp();
      gr[i * 32 + lane] = (i < ns && lane < Mp) ? z : 0.0;
    }
    __syncthreads();
    // (2) Kfu path to the hyper-parameters: sum_s df[s,n] v[s,m] dKfu[n,m]/dtheta
    for (int n = warp; n < N; n += nw) {
      if (lane < Mp) {
        double t = 0.0;
#pragma unroll 4
        for (int i = 0; i < ns; ++i) t += dft[(size_t)i * N + n] * vsm[i * 32 + lane];
        // k = s2 (1 + a + a^2/3) e^-a is in Kfu; dk/dr = -(5/3) r (1 + a) e^-a follows from it by a division (no second exp)
        const double r = fabs(a.X[(size_t)n * D + l] - zy[lane]) * inv_ell, av = VG_SQRT5 * r;
        const double ks = Kfu[n * Mp + lane];
        acc_var += t * (ks / s2);
        acc_ls += t * ks * ((5.0 / 3.0) * r * r * (1.0 + av) / (1.0 + av + (1.0 / 3.0) * av * av)) * inv_ell;
      }
    }
    // (4) G -= gr v^T ; (6) GS += gr eps_u^T ; gmu += gr
    for (int i = warp; i < Mp; i += nw) {
      if (lane < Mp) {
        double t = 0.0, u = 0.0;
#pragma unroll 4
        for (int k = 0; k < ns; ++k) {
          const double g = gr[k * 32 + i];
          t += g * vsm[k * 32 + lane];
          u += g * epsm[k * 32 + lane];
        }
        G[i * LDM + lane] -= t;
        GS[i * LDM + lane] += u;
      }
    }
    if (tid < Mp) {
      double t = 0.0;
      for (int k = 0; k < ns; ++k) t += gr[k * 32 + tid];
      gmu[tid] += t;
    }
    // (5) prior path: d f0(X) = df, d f0(Zy) = -gr
    for (int idx = tid; idx < ns * A; idx += nt) {
      const int i = idx / A, xx = idx - i * A;
      const double g = xx < N ? dft[(size_t)i * N + xx] : -gr[i * 32 + xx - N];
      const size_t o = ((size_t)pl * S + s0 + i) * A + xx;
      acc_var += g * a.f0[o] / (2.0 * s2);
      acc_ls += g * a.h0[o];
    }
    __syncthreads();
  }

  if (MODE == 1) {  // publish this chunk's partial sums
    double* part = a.partial + (size_t)blockIdx.x * kPartial;
    for (int idx = tid; idx < 32 * 32; idx += nt) {
      part[idx] = G[(idx >> 5) * LDM + (idx & 31)];
      part[1024 + idx] = GS[(idx >> 5) * LDM + (idx & 31)];
    }
    if (tid < 32) part[2048 + tid] = gmu[tid];
    const double tv = block_sum(acc_var, red);
    const double tl = block_sum(acc_ls, red);
    if (tid == 0) { part[2080] = tv; part[2081] = tl; }
    return;
  }
  if (MODE == 2) {  // fold the partial sums of all chunks, chunk 0 first
    for (int idx = tid; idx < 32 * 32; idx += nt) {
      double g = 0.0, gs = 0.0;
#pragma unroll 8
      for (int c = 0; c < a.nchunk; ++c) {        // read-only loads, several chunks in flight; the sums stay in chunk order
        const double* part = a.partial + ((size_t)pl * a.nchunk + c) * kPartial;
        g += __ldg(part + idx);
        gs += __ldg(part + 1024 + idx);
      }
      G[(idx >> 5) * LDM + (idx & 31)] = g;
      GS[(idx >> 5) * LDM + (idx & 31)] = gs;
    }
    if (tid < 32) {
      double t = 0.0;
#pragma unroll 8
      for (int c = 0; c < a.nchunk; ++c) t += __ldg(a.partial + ((size_t)pl * a.nchunk + c) * kPartial + 2048 + tid);
      gmu[tid] = t;
    }
    if (tid == 0)
      for (int c = 0; c < a.nchunk; ++c) {
        acc_var += a.partial[((size_t)pl * a.nchunk + c) * kPartial + 2080];
        acc_ls += a.partial[((size_t)pl * a.nchunk + c) * kPartial + 2081];
      }
    __syncthreads();
  }

  // the sample loop is over: its Kfu region becomes L | d ELBO / d Lc
  __syncthreads();
  for (int idx = tid; idx < 2 * 32 * LDM; idx += nt) Lsm[idx] = 0.0;
  __syncthreads();
  for (int i = warp; i < Mp; i += nw)
    if (lane < Mp) Lsm[i * LDM + lane] = a.Lc[(size_t)pl * Mp * Mp + i * Mp + lane];
  __syncthreads();
  // (7) q_sqrt_full = Lc pad(q) + jitter diag  ->  d q = tril((Lc^T GS)[2:,2:]),  GL += GS pad(q)^T
  const double* q = a.q_sqrt + (size_t)pl * M * M;
  double* dq = a.d_q_sqrt + (size_t)pl * M * M;
  for (int r_ = warp; r_ < M; r_ += nw) {
    if (lane < M) {
      const int c_ = lane;
      double t = 0.0;
      if (c_ <= r_) {
#pragma unroll 4
        for (int i = r_ + 2; i < Mp; ++i) t += Lsm[i * LDM + r_ + 2] * GS[i * LDM + c_ + 2];
        const double qv = q[r_ * M + c_];
        t -= a.klw * (qv - (r_ == c_ ? 1.0 / qv : 0.0));  // - d KL / d q_sqrt
      }
      dq[r_ * M + c_] = t;
    }
  }
  for (int i = warp; i < Mp; i += nw) {
    const int k = lane;
    if (k >= 2 && k <= i) {
      double t = 0.0;
#pragma unroll 4
      for (int j = 2; j <= k; ++j) t += GS[i * LDM + j] * q[(k - 2) * M + (j - 2)];
      GL[i * LDM + k] += t;
    }
  }
  // (8) KL reverse: b = L^-1 d, KL = 0.5 sum_{i>=2} b_i^2 + ...;  ELBO carries -KL
  if (warp == 0) gd[lane] = (lane >= 2 && lane < Mp) ? -a.klw * bvec[lane] : 0.0;   // d ELBO / d b
  __syncthreads();
  if (warp == 0) {
    const double g = warp_lowerT_mv(Lism, Mp, gd);  // d ELBO / d d = L^-T gb
    __syncwarp();
    gd[lane] = lane < Mp ? g : 0.0;
  }
  __syncthreads();
  const double c0 = a.kvec[(size_t)pl * (Mp + 4) + Mp], c1 = a.kvec[(size_t)pl * (Mp + 4) + Mp + 1];
  for (int i = warp; i < Mp; i += nw)
    if (lane <= i) GL[i * LDM + lane] -= gd[i] * bvec[lane];
  if (tid < Mp) {
    gmu[tid] += gd[tid];
    // p_mu = Khat[:, :2] c  ->  d Khat[i, 0:2] += (-gd_i) c
    G[tid * LDM + 0] -= gd[tid] * c0;
    G[tid * LDM + 1] -= gd[tid] * c1;
  }
  __syncthreads();
  if (warp == 0) {
    // gc = Khat[:, :2]^T (-gd);  c = K22^-1 q~  ->  d K22 -= (K22^-1 gc) c^T     (lane i owns row i of Khat[:, :2])
    double gc0 = 0.0, gc1 = 0.0;
    if (lane < Mp) {
      const double k0 = s2 * vg_matern52(fabs(zy[lane] - zy[0]) * inv_ell) + (lane == 0 ? a.jitter : 0.0);
      const double k1 = s2 * vg_matern52(fabs(zy[lane] - zy[1]) * inv_ell) + (lane == 1 ? a.jitter : 0.0);
      gc0 = -k0 * gd[lane];
      gc1 = -k1 * gd[lane];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gc0 += __shfl_xor_sync(kFull, gc0, o);
      gc1 += __shfl_xor_sync(kFull, gc1, o);
    }
    if (lane == 0) {
      const double l00 = Lsm[0], l10 = Lsm[LDM], l11 = Lsm[LDM + 1];
      const double y0 = gc0 / l00, y1 = (gc1 - l10 * y0) / l11;
      const double e1 = y1 / l11, e0 = (y0 - l10 * e1) / l00;
      G[0] -= e0 * c0;       G[1] -= e0 * c1;
      G[LDM] -= e1 * c0;     G[LDM + 1] -= e1 * c1;
    }
  }
  __syncthreads();
  // (9) Cholesky reverse (Murray 2016): Khat_bar = L^-T Phi(L^T L_bar) L^-1, Phi = tril with halved diagonal.
  //     With the explicit inverse this is three triangular products, no substitutions.
  double* Pm = GS;  // reuse: P = Phi(L^T GL)
  for (int i = warp; i < Mp; i += nw) {
    const int j = lane;
    double t = 0.0;
    if (j <= i) {
#pragma unroll 4
      for (int k = i; k < Mp; ++k) t += Lsm[k * LDM + i] * GL[k * LDM + j];
      if (i == j) t *= 0.5;
    }
    Pm[i * LDM + j] = t;
  }
  __syncthreads();
  double* T1 = GL;  // reuse: T1 = P Li  (lower x lower)
  for (int i = warp; i < Mp; i += nw) {
    const int j = lane;
    double t = 0.0;
    if (j <= i)
#pragma unroll 4
      for (int k = j; k <= i; ++k) t += Pm[i * LDM + k] * Lism[k * LDM + j];
    T1[i * LDM + j] = t;
  }
  __syncthreads();
  for (int i = warp; i < Mp; i += nw) {   // G += Li^T T1
    const int j = lane;
    if (j < Mp) {
      double t = 0.0;
#pragma unroll 4
      for (int k = max(i, j); k < Mp; ++k) t += Lism[k * LDM + i] * T1[k * LDM + j];
      G[i * LDM + j] += t;
    }
  }
  __syncthreads();
  // (10) contract with d Khat / d theta
  for (int i = warp; i < Mp; i += nw) {
    if (lane < Mp) {
      const double r = fabs(zy[i] - zy[lane]) * inv_ell;
      double k, dk;
      matern52_both(r, k, dk);
      const double g = G[i * LDM + lane];
      acc_var += g * k;
      acc_ls += g * s2 * dk * (-r * inv_ell);
    }
  }
  const double tv = block_sum(acc_var, red);
  const double tl = block_sum(acc_ls, red);
  if (tid == 0) {
    a.d_var[pl] = tv;
    a.d_ls[pl] = tl;
  }
  if (tid >= 2 && tid < Mp) a.d_q_mu[((size_t)p * M + tid - 2) * D + l] = gmu[tid];
}

// ---------------------------------------------------------------------------------------------
// Sample phase of the reverse pass for many samples per (problem, latent) (num_samples >= 64): the same sums as the sample
// loop of gp_backward_kernel, organised as small GEMMs over tiles of 32 samples on the FP64 tensor path (DMMA m8n8k4):
//   GV = DF Kfu            [32 x N][N x Mp]        gv[s,m] = sum_n Kfu[n,m] df[s,n]
//   GR = (GV Li^T) Li      [32 x Mp][Mp x Mp] x2   gr = Khat^-1 gv
//   T += DF^T V            [N x 32][32 x Mp]       hyper-parameter path through Kfu (contracted with dKfu/dtheta at the end)
//   G -= GR^T V,  GS += GR^T EPS                   d ELBO / d Khat, d ELBO / d q_sqrt_full
//   gmu += sum_s GR,  prior path: <[DF | -GR], f0>, <[DF | -GR], h0>
// One CTA of 256 threads per (pair, chunk of samples); T, G, GS live in DMMA accumulator registers for the whole chunk.  The
// result goes out in gp_backward_kernel's partial-sum format; gp_backward_kernel<2> folds the chunks in fixed order and
// runs the sample-independent phase.  (The 8-sample tiles, six CTA barriers per tile and per-tile Matern evaluations of
// gp_backward_kernel cost 57 ms per step at 8192 problems x 256 samples.)
// ---------------------------------------------------------------------------------------------
constexpr int kBS = 32;    // samples per tile of the batched reverse pass
constexpr int kLD = 36;    // leading dimension of 32-wide DMMA operands: the 8 rows x 4 lanes of a fragment hit 16 distinct double banks

__host__ __device__ inline int dmma_ld(int n) {   // smallest ld >= n with ld % 16 == 4 (same bank argument for wider operands)
  const int n4 = (n + 3) & ~3;
  return n4 + ((4 - (n4 & 15)) + 16) % 16;
}

// C[8x8] += A[8 x 4 KS] B[4 KS x 8];  A row-major (this warp's 8 rows), Bn "n-major": element (k, n) at Bn[n * ldb + k]
template <int KS>
__device__ __forceinline__ void wmma_rn(double& c0, double& c1, const double* __restrict__ A, int lda,
                                        const double* __restrict__ Bn, int ldb, int g, int t) {
#pragma unroll
  for (int k = 0; k < KS; ++k) dmma884(c0, c1, A[g * lda + 4 * k + t], Bn[g * ldb + 4 * k + t]);
}
// same with A given transposed: element (row, k) at At[k * lda + row]
template <int KS>
__device__ __forceinline__ void wmma_tn(double& c0, double& c1, const double* __restrict__ At, int lda,
                                        const double* __restrict__ Bn, int ldb, int g, int t) {
#pragma unroll
  for (int k = 0; k < KS; ++k) dmma884(c0, c1, At[(4 * k + t) * lda + g], Bn[g * ldb + 4 * k + t]);
}

template <int NTILES>   // point tiles of 8 (N <= 8 NTILES)
__global__ void __launch_bounds__(256, 2) gp_backward_samples_kernel(BackwardArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int D = a.D, M = a.M, Mp = M + 2, N = a.N, S = a.S, A = N + Mp;
  const int pl = blockIdx.x / a.nchunk, p = pl / D, l = pl % D;
  const int s_begin = (blockIdx.x % a.nchunk) * a.chunk, s_end = min(S, s_begin + a.chunk);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
  const int g = lane >> 2, t = lane & 3;
  constexpr int NP = NTILES * 8;
  const int ldN = dmma_ld(NP);
  double* Li = sm;                       // [32][kLD]   L^-1 (lower)
  double* LiT = Li + 32 * kLD;           // [32][kLD]   its transpose
  double* KT = LiT + 32 * kLD;           // [32][ldN]   Kfu transposed: KT[m][n]
  double* DF = KT + 32 * ldN;            // [32][ldN]   this tile's d ELBO / d f
  double* VT = DF + 32 * ldN;            // [32][kLD]   v transposed: VT[m][s]
  double* ET = VT + 32 * kLD;            // [32][kLD]   eps_u transposed
  double* B1 = ET + 32 * kLD;            // [32][kLD]   GV, then GR
  double* B2 = B1 + 32 * kLD;            // [32][kLD]   Y = GV Li^T
  double* gmu = B2 + 32 * kLD;           // [32]
  double* zy = gmu + 32;                 // [32]
  double* red = zy + 32;                 // [8]
  {                                                   // first tile's inputs: DRAM -> L2 while the operands are set up
    const int n1 = min(kBS, s_end - s_begin);
    const size_t ps1 = (size_t)pl * S + s_begin;
    if (n1 > 0) {
      if (a.df_planar) l2_prefetch_span(a.df + ((size_t)p * D + l) * S * N + (size_t)s_begin * N, (size_t)n1 * N, tid, nt);
      l2_prefetch_span(a.f0 + ps1 * A, (size_t)n1 * A, tid, nt);
      l2_prefetch_span(a.h0 + ps1 * A, (size_t)n1 * A, tid, nt);
      l2_prefetch_span(a.v + ps1 * Mp, (size_t)n1 * Mp, tid, nt);
      l2_prefetch_span(a.eps_u + ps1 * Mp, (size_t)n1 * Mp, tid, nt);
    }
  }
  const double ell = a.ls[pl], s2 = a.var[pl], inv_ell = 1.0 / ell;
  for (int idx = tid; idx < 2 * 32 * kLD + 32 * ldN; idx += nt) sm[idx] = 0.0;     // Li, LiT, KT (padding must be zero)
  if (tid < 32) { zy[tid] = tid < Mp ? zy_at(a.Z, D, l, tid) : 0.0; gmu[tid] = 0.0; }
  __syncthreads();
  for (int i = warp; i < Mp; i += nw)
    if (lane < Mp) {
      const double v = a.Linv[(size_t)pl * Mp * Mp + i * Mp + lane];
      Li[i * kLD + lane] = v;
      LiT[lane * kLD + i] = v;
    }
  for (int idx = tid; idx < Mp * N; idx += nt) {
    const int m = idx / N, n = idx - m * N;
    KT[m * ldN + n] = s2 * vg_matern52(fabs(a.X[(size_t)n * D + l] - zy[m]) * inv_ell);
  }
  // persistent accumulators: T tiles (NTILES x 4), G and GS tiles (4 x 4 each) dealt round-robin to the 8 warps
  constexpr int kTT = (NTILES * 4 + 7) / 8;
  double accT[kTT][2], accG[2][2], accS[2][2];
#pragma unroll
  for (int i = 0; i < kTT; ++i) accT[i][0] = accT[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < 2; ++i) accG[i][0] = accG[i][1] = accS[i][0] = accS[i][1] = 0.0;
  double acc_ls = 0.0, acc_var = 0.0, acc_f0 = 0.0;   // acc_f0: sum g f0, scaled by 1 / (2 s2) once at the end
  const double* dfp = a.df + (size_t)p * S * N * D + (a.df_planar ? (size_t)l * S * N : (size_t)l);
  const int dfs = a.df_planar ? 1 : D;

  for (int s0 = s_begin; s0 < s_end; s0 += kBS) {
    const int ns = min(kBS, s_end - s0);
    __syncthreads();                                    // previous tile fully consumed
    if (s0 + kBS < s_end) {                             // next tile's inputs: DRAM -> L2 while this tile computes
      const int s1 = s0 + kBS, n1 = min(kBS, s_end - s1);
      const size_t ps1 = (size_t)pl * S + s1;
      if (a.df_planar) l2_prefetch_span(dfp + (size_t)s1 * N, (size_t)n1 * N, tid, nt);
      l2_prefetch_span(a.f0 + ps1 * A, (size_t)n1 * A, tid, nt);
      l2_prefetch_span(a.h0 + ps1 * A, (size_t)n1 * A, tid, nt);
      l2_prefetch_span(a.v + ps1 * Mp, (size_t)n1 * Mp, tid, nt);
      l2_prefetch_span(a.eps_u + ps1 * Mp, (size_t)n1 * Mp, tid, nt);
    }
    // (staging DF one tile ahead with cp.async changed nothing here: 2758 vs 2792 us at 1024 problems x 256 samples, and its
    // second buffer costs the second CTA per SM once N > 64)
#pragma unroll 2
    for (int idx = tid; idx < kBS * NP; idx += nt) {    // DF[s][n], zero padded; its share of the prior path rides along
      const int i = idx / NP, n = idx - i * NP;           // (d f0(X) = df: the f0 / h0 loads overlap with the df loads)
      double dv = 0.0;
      if (i < ns && n < N) {
        dv = __ldg(dfp + ((size_t)(s0 + i) * N + n) * dfs);
        const size_t o = ((size_t)pl * S + s0 + i) * A + n;
        acc_f0 += dv * __ldg(a.f0 + o);
        acc_ls += dv * __ldg(a.h0 + o);
      }
      DF[i * ldN + n] = dv;
    }
#pragma unroll 2
    for (int idx = tid; idx < kBS * 32; idx += nt) {    // VT[m][s], ET[m][s]
      const int i = idx >> 5, m = idx & 31;
      const bool ok = i < ns && m < Mp;
      VT[m * kLD + i] = ok ? __ldg(a.v + ((size_t)pl * S + s0 + i) * Mp + m) : 0.0;
      ET[m * kLD + i] = ok ? __ldg(a.eps_u + ((size_t)pl * S + s0 + i) * Mp + m) : 0.0;
    }
    __syncthreads();
    // GV = DF KT^T: 4 x 4 output tiles, 2 per warp
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int tile = warp * 2 + u, mi = tile >> 2, ni = tile & 3;
      double c0 = 0.0, c1 = 0.0;
      wmma_rn<NTILES * 2>(c0, c1, DF + mi * 8 * ldN, ldN, KT + ni * 8 * ldN, ldN, g, t);
      *reinterpret_cast<double2*>(B1 + (mi * 8 + g) * kLD + ni * 8 + 2 * t) = make_double2(c0, c1);
    }
    __syncthreads();
    // Y = GV Li^T  (element (k, n) of the right operand = Li[n][k]: Li itself is n-major)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int tile = warp * 2 + u, mi = tile >> 2, ni = tile & 3;
      double c0 = 0.0, c1 = 0.0;
      wmma_rn<8>(c0, c1, B1 + mi * 8 * kLD, kLD, Li + ni * 8 * kLD, kLD, g, t);
      *reinterpret_cast<double2*>(B2 + (mi * 8 + g) * kLD + ni * 8 + 2 * t) = make_double2(c0, c1);
    }
    __syncthreads();
    // GR = Y Li  (right operand element (k, n) = Li[k][n] = LiT[n][k])
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int tile = warp * 2 + u, mi = tile >> 2, ni = tile & 3;
      double c0 = 0.0, c1 = 0.0;
      wmma_rn<8>(c0, c1, B2 + mi * 8 * kLD, kLD, LiT + ni * 8 * kLD, kLD, g, t);
      *reinterpret_cast<double2*>(B1 + (mi * 8 + g) * kLD + ni * 8 + 2 * t) = make_double2(c0, c1);
    }
    __syncthreads();
    // T += DF^T V ; G -= GR^T V ; GS += GR^T EPS   (K = the 32 samples of the tile)
#pragma unroll
    for (int u = 0; u < kTT; ++u) {
      const int tile = warp + 8 * u;
      if (tile < NTILES * 4) {
        const int ni_ = tile >> 2, mi_ = tile & 3;      // point tile, inducing tile
        wmma_tn<8>(accT[u][0], accT[u][1], DF + ni_ * 8, ldN, VT + mi_ * 8 * kLD, kLD, g, t);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int tile = warp * 2 + u, ii = tile >> 2, jj = tile & 3;
      wmma_tn<8>(accG[u][0], accG[u][1], B1 + ii * 8, kLD, VT + jj * 8 * kLD, kLD, g, t);
      wmma_tn<8>(accS[u][0], accS[u][1], B1 + ii * 8, kLD, ET + jj * 8 * kLD, kLD, g, t);
    }
    if (tid < 32) {
      double tsum = 0.0;
      for (int i = 0; i < ns; ++i) tsum += B1[i * kLD + tid];
      gmu[tid] += tsum;
    }
    // prior path at the inducing points: d f0(Zy) = -gr   (sample = warp-strided, inducing point = lane: no index division)
    for (int i = warp; i < ns; i += nw)
      if (lane < Mp) {
        const double gg = -B1[i * kLD + lane];
        const size_t o = ((size_t)pl * S + s0 + i) * A + N + lane;
        acc_f0 += gg * __ldg(a.f0 + o);
        acc_ls += gg * __ldg(a.h0 + o);
      }
  }
  acc_var += acc_f0 / (2.0 * s2);
  // hyper-parameter path through Kfu: sum_{n,m} T[n,m] dKfu[n,m]/dtheta
#pragma unroll
  for (int u = 0; u < kTT; ++u) {
    const int tile = warp + 8 * u;
    if (tile < NTILES * 4) {
      const int n = (tile >> 2) * 8 + g;
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        const int m = (tile & 3) * 8 + 2 * t + e2;
        if (n < N && m < Mp) {
          const double r = fabs(a.X[(size_t)n * D + l] - zy[m]) * inv_ell;
          double k, dk;
          matern52_both(r, k, dk);
          acc_var += accT[u][e2] * k;
          acc_ls += accT[u][e2] * s2 * dk * (-r * inv_ell);
        }
      }
    }
  }
  // publish the chunk's partial sums in gp_backward_kernel's format: G | GS (32 x 32 each) | gmu | acc_var | acc_ls
  double* part = a.partial + (size_t)blockIdx.x * kPartial;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int tile = warp * 2 + u, i = (tile >> 2) * 8 + g, j = (tile & 3) * 8 + 2 * t;
    part[i * 32 + j] = -accG[u][0];          part[i * 32 + j + 1] = -accG[u][1];
    part[1024 + i * 32 + j] = accS[u][0];    part[1024 + i * 32 + j + 1] = accS[u][1];
  }
  __syncthreads();
  if (tid < 32) part[2048 + tid] = gmu[tid];
  const double tv = block_sum(acc_var, red);
  const double tl = block_sum(acc_ls, red);
  if (tid == 0) { part[2080] = tv; part[2081] = tl; }
}

// ---------------------------------------------------------------------------------------------
// Pathwise update for many samples per (problem, latent), batched as small GEMMs on the FP64 tensor path (the S >= 64
// companion of the tensor-core sampler; pathwise_update_kernel above is its warp-per-sample form, ~9x the instructions):
//   U = mu + EPS S^T;  R = U - F0(Zy) - sqrt(jitter) EPSJ;  Y = R Li^T;  V = Y Li;  F = F0(X) + V Kfu^T
// One CTA of 256 threads per (pair, chunk of samples); a warp owns a tile of 8 samples end to end (its operands chain
// through a private shared-memory tile, so the sample loop has no CTA barrier).
// ---------------------------------------------------------------------------------------------
template <int NTILES>   // point tiles of 8 (Nq <= 8 NTILES)
__global__ void __launch_bounds__(256, 2) pathwise_update_mma_kernel(PathwiseArgs a, const double* __restrict__ meta, int chunk) {
  extern __shared__ __align__(16) double sm[];
  if (meta[0] == 0.0) return;                         // general sampler does its own update
  const int D = a.D, M = a.M, Mp = M + 2, Nq = a.Nq, S = a.S, A = Nq + Mp;
  const int nchunk = (S + chunk - 1) / chunk;
  const int pl = blockIdx.x / nchunk, p = pl / D, l = pl % D;
  const int s_begin = (blockIdx.x % nchunk) * chunk, s_end = min(S, s_begin + chunk);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
  const int g = lane >> 2, t = lane & 3;
  {                                                   // the chunk's first tiles: DRAM -> L2 while the operands are set up
    const int n1 = min(nw * 8, s_end - s_begin);
    const size_t ps1 = (size_t)pl * S + s_begin;
    if (n1 > 0) {
      l2_prefetch_span(a.eps_u + ps1 * Mp, (size_t)n1 * Mp, tid, nt);
      l2_prefetch_span(a.eps_j + ps1 * Mp, (size_t)n1 * Mp, tid, nt);
      l2_prefetch_span(a.f0 + ps1 * A, (size_t)n1 * A, tid, nt);
    }
  }
  double* Ss = sm;                                    // [32][kLD]  q_sqrt_full, n-major for U = EPS S^T
  double* Li = Ss + 32 * kLD;                         // [32][kLD]  L^-1
  double* LiT = Li + 32 * kLD;                        // [32][kLD]
  double* KF = LiT + 32 * kLD;                        // [8 NTILES][kLD]  Kfu[n][m] (n-major for F = V Kfu^T)
  double* mu = KF + (size_t)NTILES * 8 * kLD;         // [32]
  double* zy = mu + 32;                               // [32]
  double* tiles = zy + 32;                            // [nw][2][8][kLD] per-warp operand tiles
  const double ell = a.ls[pl], s2 = a.var[pl], sqrtj = sqrt(a.jitter), inv_ell = 1.0 / ell;
  for (int idx = tid; idx < (3 * 32 + NTILES * 8) * kLD; idx += nt) sm[idx] = 0.0;
  if (tid < 32) {
    zy[tid] = tid < Mp ? zy_at(a.Z, D, l, tid) : 0.0;
    mu[tid] = tid >= Mp ? 0.0 : (tid < 2 ? a.query_latent[((size_t)p * 2 + tid) * D + l] : a.q_mu[((size_t)p * M + tid - 2) * D + l]);
  }
  __syncthreads();
  for (int i = warp; i < Mp; i += nw)
    if (lane < Mp) {
      const double v = a.Linv[(size_t)pl * Mp * Mp + i * Mp + lane];
      Li[i * kLD + lane] = v;
      LiT[lane * kLD + i] = v;
      Ss[i * kLD + lane] = a.Sfull[(size_t)pl * Mp * Mp + i * Mp + lane];
    }
  for (int idx = tid; idx < Nq * Mp; idx += nt) {
    const int n = idx / Mp, m = idx - n * Mp;
    KF[n * kLD + m] = s2 * vg_matern52(fabs(a.Xq[(size_t)n * D + l] - zy[m]) * inv_ell);
  }
  __syncthreads();
  double* T0 = tiles + (size_t)warp * 2 * 8 * kLD;    // [8][kLD]
  double* T1 = T0 + 8 * kLD;
  for (int s0 = s_begin + warp * 8; s0 < s_end; s0 += nw * 8) {
    const int ns = min(8, s_end - s0);
    const size_t ps0 = (size_t)pl * S + s0;
    if (s0 + nw * 8 < s_end) {                          // this warp's next tile: DRAM -> L2 while this one computes
      const int s1 = s0 + nw * 8, n1 = min(8, s_end - s1);
      const size_t ps1 = (size_t)pl * S + s1;
      l2_prefetch_span(a.eps_u + ps1 * Mp, (size_t)n1 * Mp, lane, 32);
      l2_prefetch_span(a.eps_j + ps1 * Mp, (size_t)n1 * Mp, lane, 32);
      l2_prefetch_span(a.f0 + ps1 * A, (size_t)n1 * A, lane, 32);
    }
    // This tile's inputs are requested up front through the read-only path (nothing in this kernel writes them): one round
    // trip for EPS, F0(Zy) and EPSJ instead of one per 8-column block behind the shared-memory stores.
    double e8[8], fz[4][2], ej[4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) e8[i] = (i < ns && lane < Mp) ? __ldg(a.eps_u + (ps0 + i) * Mp + lane) : 0.0;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        const int m = ni * 8 + 2 * t + e2;
        const bool ok = g < ns && m < Mp;
        fz[ni][e2] = ok ? __ldg(a.f0 + (ps0 + g) * A + Nq + m) : 0.0;
        ej[ni][e2] = ok ? __ldg(a.eps_j + (ps0 + g) * Mp + m) : 0.0;
      }
    // T0 <- EPS tile [8 samples][32]
#pragma unroll
    for (int i = 0; i < 8; ++i) T0[i * kLD + lane] = e8[i];
    __syncwarp();
    // R = mu + EPS S^T - F0(Zy) - sqrt(jitter) EPSJ  -> T1      (right operand element (k, n) = S[n][k]: Ss is n-major)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      double c0 = 0.0, c1 = 0.0;
      wmma_rn<8>(c0, c1, T0, kLD, Ss + ni * 8 * kLD, kLD, g, t);
      const int m0 = ni * 8 + 2 * t;
      double r0 = 0.0, r1 = 0.0;
      if (g < ns) {
        if (m0 < Mp) r0 = mu[m0] + c0 - fz[ni][0] - sqrtj * ej[ni][0];
        if (m0 + 1 < Mp) r1 = mu[m0 + 1] + c1 - fz[ni][1] - sqrtj * ej[ni][1];
      }
      *reinterpret_cast<double2*>(T1 + g * kLD + m0) = make_double2(r0, r1);
    }
    __syncwarp();
    // F0(X) of this tile: requested now, needed after the two triangular products
    double fx[NTILES][2];
#pragma unroll
    for (int ni = 0; ni < NTILES; ++ni)
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        const int n = ni * 8 + 2 * t + e2;
        fx[ni][e2] = (g < ns && n < Nq) ? __ldg(a.f0 + (ps0 + g) * A + n) : 0.0;
      }
    // Y = R Li^T -> T0
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      double c0 = 0.0, c1 = 0.0;
      wmma_rn<8>(c0, c1, T1, kLD, Li + ni * 8 * kLD, kLD, g, t);
      *reinterpret_cast<double2*>(T0 + g * kLD + ni * 8 + 2 * t) = make_double2(c0, c1);
    }
    __syncwarp();
    // V = Y Li -> T1 (and to memory for the reverse pass)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      double c0 = 0.0, c1 = 0.0;
      wmma_rn<8>(c0, c1, T0, kLD, LiT + ni * 8 * kLD, kLD, g, t);
      const int m0 = ni * 8 + 2 * t;
      *reinterpret_cast<double2*>(T1 + g * kLD + m0) = make_double2(c0, c1);
      if (a.v != nullptr && g < ns) {
        if (m0 < Mp) a.v[(ps0 + g) * Mp + m0] = c0;
        if (m0 + 1 < Mp) a.v[(ps0 + g) * Mp + m0 + 1] = c1;
      }
    }
    __syncwarp();
    // F = F0(X) + V Kfu^T        (right operand element (k = m, n) = Kfu[n][m]: KF is n-major)
#pragma unroll
    for (int ni = 0; ni < NTILES; ++ni) {
      double c0 = 0.0, c1 = 0.0;
      wmma_rn<8>(c0, c1, T1, kLD, KF + ni * 8 * kLD, kLD, g, t);
      const int n0 = ni * 8 + 2 * t;
      if (g < ns) {
        if (n0 < Nq) a.f[f_index(a.f_planar, p, s0 + g, n0, l, S, Nq, D)] = c0 + fx[ni][0];
        if (n0 + 1 < Nq) a.f[f_index(a.f_planar, p, s0 + g, n0 + 1, l, S, Nq, D)] = c1 + fx[ni][1];
      }
    }
    __syncwarp();
  }
}

// SVGP posterior mean at Xq (GPflow posterior().predict_f with whiten=False, models/vgpmp.py:315):
//   mean[n,l] = Kfu[n,:] Khat^-1 q_mu_full[:,l]        one CTA per (problem, latent)
__global__ void __launch_bounds__(128) predict_mean_kernel(int D, int M, int Nq, vgpmp_params P,
                                                          const double* __restrict__ Xq, const double* __restrict__ Lc,
                                                          double* __restrict__ mean) {
  __shared__ double Lsm[32 * LDM], wv[32], zy[32];
  const int pl = blockIdx.x, p = pl / D, l = pl % D, Mp = M + 2, tid = threadIdx.x;
  const double ell = P.lengthscales[pl], s2 = P.variances[pl];
  for (int idx = tid; idx < Mp * Mp; idx += blockDim.x) Lsm[(idx / Mp) * LDM + idx % Mp] = Lc[(size_t)pl * Mp * Mp + idx];
  if (tid < 32) zy[tid] = tid < Mp ? zy_at(P.Z, D, l, tid) : 0.0;
  __syncthreads();
  if (tid < 32) {
    double mu = 0.0;
    if (tid < Mp)
      mu = tid < 2 ? P.query_latent[((size_t)p * 2 + tid) * D + l] : P.q_mu[((size_t)p * M + tid - 2) * D + l];
    const double y = warp_fwd_subst(Lsm, Mp, mu);
    wv[tid] = warp_bwd_subst(Lsm, Mp, y);
  }
  __syncthreads();
  for (int n = tid; n < Nq; n += blockDim.x) {
    const double x = Xq[(size_t)n * D + l];
    double acc = 0.0;
    for (int m = 0; m < Mp; ++m) acc += s2 * vg_matern52(fabs(x - zy[m]) / ell) * wv[m];
    mean[((size_t)p * Nq + n) * D + l] = acc;
  }
}

// ELBO[p] = alpha/S * sum_{s,n} logp - sum_l KL_l      models/vgpmp.py:287-289
// nseg > 1 (few problems with very many samples, e.g. the single-problem 65 536-sample mode): the sum over (s, n) is cut
// into nseg fixed segments, one CTA each, and elbo_finish_kernel adds the partial sums in segment order (deterministic).
__global__ void __launch_bounds__(256) elbo_reduce_kernel(int D, int SN, double scale, double klw,
                                                         const double* __restrict__ logp,
                                                         const double* __restrict__ kl_l, double* __restrict__ elbo,
                                                         double* __restrict__ kl_out, double* __restrict__ loss_out,
                                                         int nseg, double* __restrict__ partial) {
  __shared__ double red[8];
  const int p = blockIdx.x / nseg, seg = blockIdx.x - p * nseg;
  const int per = (SN + nseg - 1) / nseg, lo = seg * per, hi = min(SN, lo + per);
  double t = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) t += logp[(size_t)p * SN + i];
  const double lik = block_sum(t, red);
  if (threadIdx.x != 0) return;
  if (nseg > 1) { partial[blockIdx.x] = lik; return; }
  double kl = 0.0;
  for (int l = 0; l < D; ++l) kl += kl_l[(size_t)p * D + l];
  const double e = scale * lik - klw * kl;
  elbo[p] = e;
  if (kl_out != nullptr) kl_out[p] = kl;
  if (loss_out != nullptr) loss_out[p] = -e;      // training_loss = -ELBO (utils/miscellaneous.py:77-79)
}

__global__ void elbo_finish_kernel(int D, int Bp, double scale, double klw, const double* __restrict__ partial, int nseg,
                                   const double* __restrict__ kl_l, double* __restrict__ elbo, double* __restrict__ kl_out,
                                   double* __restrict__ loss_out) {
  // one warp per problem: lane-strided sums of the segment partials, then a fixed shuffle tree (deterministic)
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (p >= Bp) return;
  double lik = 0.0, kl = 0.0;
  for (int i = lane; i < nseg; i += 32) lik += partial[(size_t)p * nseg + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lik += __shfl_xor_sync(kFull, lik, o);
  if (lane != 0) return;
  for (int l = 0; l < D; ++l) kl += kl_l[(size_t)p * D + l];
  const double e = scale * lik - klw * kl;
  elbo[p] = e;
  if (kl_out != nullptr) kl_out[p] = kl;
  if (loss_out != nullptr) loss_out[p] = -e;
}

// Keras Adam on the unconstrained variables, loss = -ELBO.
__device__ __forceinline__ double softplus_d(double x) { return x > 30.0 ? x : log1p(exp(x)); }

__global__ void adam_kernel(int D, int M, int Bp, vgpmp_adam st, vgpmp_grads g, const unsigned long long* __restrict__ step_dev) {
  // bias-corrected rate of this step, lr * sqrt(1 - beta2^t) / (1 - beta1^t): evaluated on the device (once per CTA) in both
  // the plain and the CUDA-graph path, so that the two are bit-identical; in graph replay the step counter is device-resident
  __shared__ double lr_sh;
  if (threadIdx.x == 0) {
    const double t = (double)((step_dev != nullptr ? *step_dev : (unsigned long long)st.step) + 1ull);
    lr_sh = st.learning_rate * sqrt(1.0 - pow(st.beta2, t)) / (1.0 - pow(st.beta1, t));
  }
  __syncthreads();
  const double lr_t = lr_sh;
  const int per = M * D + D * M * M + 2 * D;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)Bp * per) return;
  const int p = (int)(gid / per), e = (int)(gid % per);
  double grad, *var_ptr;
  int kind;  // 0 q_mu, 1 q_sqrt, 2 ls, 3 var
  int local;
  if (e < M * D) { kind = 0; local = e; }
  else if (e < M * D + D * M * M) { kind = 1; local = e - M * D; }
  else if (e < M * D + D * M * M + D) { kind = 2; local = e - M * D - D * M * M; }
  else { kind = 3; local = e - M * D - D * M * M - D; }
  bool train;
  if (kind == 0) { train = st.train_q_mu; var_ptr = st.q_mu + (size_t)p * M * D + local; grad = g.d_q_mu[(size_t)p * M * D + local]; }
  else if (kind == 1) {
    const int rc = local % (M * M);
    train = st.train_q_sqrt && (rc % M) <= (rc / M);
    var_ptr = st.q_sqrt + (size_t)p * D * M * M + local; grad = g.d_q_sqrt[(size_t)p * D * M * M + local];
  } else if (kind == 2) {
    train = st.train_lengthscales; var_ptr = st.raw_lengthscales + (size_t)p * D + local;
    grad = g.d_lengthscales[(size_t)p * D + local];
  } else {
    train = st.train_variances; var_ptr = st.raw_variances + (size_t)p * D + local;
    grad = g.d_variances[(size_t)p * D + local];
  }
  if (!train) return;
  double x = *var_ptr;
  if (kind >= 2) grad *= 1.0 / (1.0 + exp(-x));  // d softplus / d raw
  grad = -grad;                                   // minimise -ELBO
  double m = st.m[gid], v = st.v[gid];
  m += (grad - m) * (1.0 - st.beta1);
  v += (grad * grad - v) * (1.0 - st.beta2);
  st.m[gid] = m;
  st.v[gid] = v;
  x -= lr_t * m / (sqrt(v) + st.eps);
  *var_ptr = x;
  if (kind == 2) st.lengthscales[(size_t)p * D + local] = softplus_d(x);
  if (kind == 3) st.variances[(size_t)p * D + local] = st.variance_lower + softplus_d(x);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 draw generator: element-wise counters, so any launch shape / any GPU produces the same draw for the
// same (seed, iteration, stream, problem, latent, row, column).
// ---------------------------------------------------------------------------------------------
struct RngArgs {
  int D, B, S, Mp, Bp;
  int64_t problem_offset, sample_offset;
  uint64_t seed, iteration;
  PhiloxKeys rk;                         // round keys of `seed`
  const unsigned long long* iter_dev;   // CUDA-graph replay: added to `iteration`
  double *omega, *tau, *w, *eps_u, *eps_j;
};

// Blocks are grouped per (problem, latent) pair: `bpp` blocks of 256 threads walk the pair's B basis rows, then its
// S*ceil(B/4) weight quads, then its S*ceil(Mp/2) eps pairs.  All index arithmetic is 32-bit (the flat 64-bit div/mod
// of the first version was half of the kernel's instructions); the Philox keys are unchanged.
__global__ void __launch_bounds__(256) rng_fill_kernel(RngArgs a, uint32_t bpp, uint32_t nblocks,
                                                       const double* __restrict__ skip_if_grid) {
  // lazy draws: this launch only matters when the equispaced sampler did NOT run (it generated its own omega / tau / w)
  if (skip_if_grid != nullptr && skip_if_grid[0] != 0.0) return;
  if (a.iter_dev != nullptr) a.iteration += *a.iter_dev;
  const uint32_t B = a.B, S = a.S, D = a.D, Mp = a.Mp, B4 = (B + 3) / 4, M2 = (Mp + 1) / 2;
  const uint32_t nA = a.omega != nullptr ? B : 0, nW = a.w != nullptr ? S * B4 : 0, nE = a.eps_u != nullptr ? S * M2 : 0;
  for (uint32_t vb = blockIdx.x; vb < nblocks; vb += gridDim.x) {   // (the gated launch uses a small grid: it usually exits above)
  const uint32_t pair = vb / bpp;                                // local (problem, latent)
  uint32_t t = (vb - pair * bpp) * 256u + threadIdx.x;            // index inside the pair
  const uint32_t pl_ = pair / D, l = pair - pl_ * D;
  const uint64_t p = (uint64_t)pl_ + (uint64_t)a.problem_offset;
  if (t < nA) {
    const uint32_t b = t;
    const uint64_t key = (p * D + l) * B + b;
    const size_t gid = (size_t)pair * B + b;
    // Gamma(5/2, rate 5/2) = chi^2_5 / 5  -> omega = z / sqrt(gamma): Matern-5/2 spectral density (Student-t_5)
    double z[16];
    const int ncall = (5 + a.D + 3) / 4;                    // <= 4 for D <= 8
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < ncall) normal4(a.rk, a.iteration, 1u, key * 4 + k, z + 4 * k);
    const double gam = (z[0] * z[0] + z[1] * z[1] + z[2] * z[2] + z[3] * z[3] + z[4] * z[4]) / 5.0;
    const double rs = rsqrt(gam);
#pragma unroll
    for (int d = 0; d < VGPMP_MAX_DOF; ++d)
      if (d < a.D) a.omega[gid * a.D + d] = z[5 + d] * rs;
    uint32_t c[4] = {(uint32_t)key, (uint32_t)(key >> 32), (uint32_t)a.iteration, 3u};
    philox4x32(c, a.rk);
    a.tau[gid] = 6.283185307179586476925 * u01(c[0], c[1]);
    continue;
  }
  t -= nA;
  if (t < nW) {
    const uint32_t sl = t / B4, b4 = t - sl * B4;
    const uint64_t s = (uint64_t)sl + (uint64_t)a.sample_offset;
    const uint64_t key = ((p * D + l) * (uint64_t)(1u << 24) + s) * B4 + b4;  // sample index < 2^24
    double z[4];
    normal4(a.rk, a.iteration, 4u, key, z);
    double* dst = a.w + ((size_t)pair * S + sl) * B + (size_t)b4 * 4;
    if (b4 * 4 + 4 <= B && (B & 1) == 0) {
      reinterpret_cast<double2*>(dst)[0] = make_double2(z[0], z[1]);
      reinterpret_cast<double2*>(dst)[1] = make_double2(z[2], z[3]);
    } else {
      for (uint32_t k = 0; k < 4; ++k)
        if (b4 * 4 + k < B) dst[k] = z[k];
    }
    continue;
  }
  t -= nW;
  if (t < nE) {
    const uint32_t sl = t / M2, m2 = t - sl * M2;
    const uint64_t s = (uint64_t)sl + (uint64_t)a.sample_offset;
    const uint64_t key = ((p * D + l) * (uint64_t)(1u << 24) + s) * 16 + m2;
    double z[4];
    normal4(a.rk, a.iteration, 5u, key, z);
    const size_t o = ((size_t)pair * S + sl) * Mp + (size_t)m2 * 2;
    a.eps_u[o] = z[0];
    a.eps_j[o] = z[2];
    if (m2 * 2 + 1 < Mp) { a.eps_u[o + 1] = z[1]; a.eps_j[o + 1] = z[3]; }
  }
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_kuu(vgpmp_handle* h, const double* Z, const double* ls, const double* var, double jitter, double* K,
                       int Bp, int M, cudaStream_t s) {
  kuu_kernel<<<Bp * h->robot.dof, 128, 0, s>>>(h->robot.dof, M, jitter, Z, ls, var, K);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_kuf(vgpmp_handle* h, const double* Z, const double* X, const double* ls, const double* var,
                       double* Kuf, int Bp, int M, int N, cudaStream_t s) {
  kuf_kernel<<<Bp * h->robot.dof, 128, 0, s>>>(h->robot.dof, M, N, Z, X, ls, var, Kuf);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_gp_prepare(vgpmp_handle* h, const vgpmp_dims& d, const vgpmp_params& p, double* Lc, double* Sfull,
                              double* kl_l, double* kvec, double* Linv, cudaStream_t s) {
  gp_prepare_kernel<<<d.num_problems * h->robot.dof, 128, 0, s>>>(h->robot.dof, d.num_inducing, h->lik.jitter, p, Lc,
                                                                  Sfull, kl_l, kvec, Linv);
  h->launches++;
  return cudaGetLastError();
}

// gated on the device-side input probe: the caller handed over NO omega / tau / w buffers (lazy draws that are never
// materialised) but X / Z turned out not to be an equispaced rank-1 grid, so no sampler could run.  Fail loudly: the
// sample paths (hence the ELBO and every gradient) become NaN.
__global__ void poison_paths_kernel(double* __restrict__ f, size_t n, const double* __restrict__ meta) {
  if (meta[0] != 0.0) return;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) f[i] = nan;
}

// Which sampler takes equispaced rank-1 inputs of this shape: tensor-core (tcgen05, 3xTF32) for large sample counts,
// register-resident DMMA when the points fit its 12 row tiles, shared-memory DMMA up to 192 points, else the general kernel.
enum SamplerKind { SAMPLER_GENERAL = 0, SAMPLER_DMMA, SAMPLER_RR, SAMPLER_TC };

static SamplerKind pick_sampler(const vgpmp_handle* h, const PathwiseArgs& a) {
  const int A = a.Nq + a.M + 2;
  if (!h->allow_grid_path || a.Nq < 2 || a.M < 2) return SAMPLER_GENERAL;
  if (h->allow_tc_path && a.S >= h->tc_min_samples && pathwise_tc_supported(a)) return SAMPLER_TC;
  if (h->allow_rr_path && (a.Nq + 2 + 7) / 8 + (a.M + 7) / 8 <= kRT) return SAMPLER_RR;
  if (h->allow_dmma_path && A <= 192) return SAMPLER_DMMA;
  return SAMPLER_GENERAL;
}

int sampler_generates_draws(const vgpmp_handle* h, const vgpmp_dims& d) {
  PathwiseArgs a{};
  a.D = h->robot.dof; a.M = d.num_inducing; a.Nq = d.num_timesteps; a.S = d.num_samples; a.B = d.num_bases;
  const SamplerKind k = pick_sampler(h, a);
  return (k == SAMPLER_TC || k == SAMPLER_RR) ? 1 : 0;
}

cudaError_t launch_pathwise(vgpmp_handle* h, const vgpmp_dims& d, const vgpmp_params& p, const vgpmp_draws& r,
                            const double* Xq, int Nq, double* Lc, double* Sfull, double* Linv, double* kl_l, double* kvec,
                            double* f, double* v, double* f0, double* h0, double* meta, int f_planar, cudaStream_t s) {
  // Schedule: [draw materialisation if needed] -> input probe -> equispaced sampler (stops at the prior draw f0 / h0; it
  // needs no GP factor) -> gp_prepare_update_kernel (Kuu, Cholesky, L^-1, q_sqrt_full, KL, then the pathwise update) ->
  // general sampler, which exits at once unless the probe found X / Z not to be an equispaced rank-1 grid.
  // (Fusing the preparation into the sampler CTAs was tried: neutral-to-slower.)
  PathwiseArgs a;
  a.D = h->robot.dof; a.M = d.num_inducing; a.Nq = Nq; a.S = d.num_samples; a.B = d.num_bases;
  const int Mp = a.M + 2, A = Nq + Mp;
  a.XG = (A + 31) / 32;
  if (a.XG > 24) return cudaErrorInvalidValue;
  a.KS = 24 / a.XG;
  if (a.KS > a.B) a.KS = a.B;
  a.jitter = h->lik.jitter;
  {
    // enough CTAs to fill the machine a few times over even for a single problem with many samples
    const int pairs = d.num_problems * a.D, tiles = (a.S + kST - 1) / kST;
    int want = (4 * h->num_sms + pairs - 1) / pairs;
    a.nchunk = std::max(1, std::min(tiles, want));
    a.chunk = ((tiles + a.nchunk - 1) / a.nchunk) * kST;
    a.nchunk = (a.S + a.chunk - 1) / a.chunk;
  }
  a.Z = p.Z; a.Xq = Xq; a.ls = p.lengthscales; a.var = p.variances; a.q_mu = p.q_mu; a.query_latent = p.query_latent;
  a.omega = r.omega; a.tau = r.tau; a.w = r.w; a.eps_u = r.eps_u; a.eps_j = r.eps_j;
  a.Lc = Lc; a.Sfull = Sfull; a.Linv = Linv; a.f = f; a.v = v; a.f0 = f0; a.h0 = h0; a.f_planar = f_planar;
  cudaError_t e;
  const int pairs = d.num_problems * a.D;
  const SamplerKind kind = f0 != nullptr ? pick_sampler(h, a) : SAMPLER_GENERAL;
  const bool tc_path = kind == SAMPLER_TC, rr_path = kind == SAMPLER_RR;
  const bool fast = kind != SAMPLER_GENERAL;
  // lazy draws (vgpmp_rng_fill_lazy): omega / tau / w of this set were never written.  The tensor-core and the
  // register-resident samplers generate them in their producer warps; anything else needs them in memory first.
  // The state is one-shot: whatever happens below, this call consumes it.
  const bool lazy = h->lazy.valid && h->lazy.omega == r.omega && h->lazy.tau == r.tau && h->lazy.w == r.w &&
                    h->lazy.num_problems == d.num_problems && h->lazy.num_samples == d.num_samples &&
                    h->lazy.num_bases == d.num_bases;
  const vgpmp_handle::LazyDraws lz = h->lazy;
  h->lazy.valid = false;
  const bool have_buffers = r.omega != nullptr && r.tau != nullptr && r.w != nullptr;
  if (!lazy && !have_buffers) return cudaErrorInvalidValue;   // NULL draw buffers are only legal right after vgpmp_rng_fill_lazy
  a.gen_draws = 0; a.seed = 0; a.iteration = 0; a.problem_offset = 0; a.sample_offset = 0;
  a.iter_dev = h->capture_iter_dev;
  if (lazy && (tc_path || rr_path)) {
    a.gen_draws = 1;
    a.seed = lz.seed; a.iteration = lz.iteration; a.problem_offset = lz.problem_offset; a.sample_offset = lz.sample_offset;
    philox_round_keys(a.seed, a.rk);
  } else if (lazy) {
    if (!have_buffers) return cudaErrorInvalidValue;          // no in-kernel generator for this shape: the caller must provide buffers
    if ((e = launch_rng_fill(h, d, lz.seed, lz.iteration, lz.problem_offset, lz.sample_offset, const_cast<double*>(r.omega),
                             const_cast<double*>(r.tau), const_cast<double*>(r.w), nullptr, nullptr, s, nullptr)) != cudaSuccess)
      return e;
  }
  if (!fast) {
    if ((e = launch_gp_prepare(h, d, p, Lc, Sfull, kl_l, kvec, Linv, s)) != cudaSuccess) return e;
  } else {
    analyze_grid_kernel<<<1, 256, 0, s>>>(a.D, a.M, Nq, Xq, p.Z, meta);
    h->launches++;
    if (tc_path) {
      if ((e = launch_pathwise_tc(h, a, pairs, meta, s)) != cudaSuccess) return e;
    } else if (rr_path) {
      // 9 .. 63 samples: NT tiles of 8 samples per CTA pass (pathwise_rrm_kernel), as long as that leaves enough CTAs
      const int tiles = (a.S + kST - 1) / kST;
      int NT = std::min(4, tiles);
      const size_t min_ctas = h->rrm_min_ctas >= 0 ? (size_t)h->rrm_min_ctas : (size_t)2 * h->num_sms;
      while (NT > 1 && (size_t)pairs * ((tiles + NT - 1) / NT) < min_ctas) --NT;
      if (NT > 1) {
        NT = (tiles + ((tiles + NT - 1) / NT) - 1) / ((tiles + NT - 1) / NT);     // balance the passes
        a.chunk = NT * kST;
        a.nchunk = (a.S + a.chunk - 1) / a.chunk;
        const size_t ring = (size_t)kRS * kRB * ((kRMW + 8 * NT + 11) / 16 * 16 + 4), fold = (size_t)2 * 8 * NT * kRT * 8;
        const size_t smem_r = sizeof(double) * (std::max(ring, fold) + 2 * kRS);
        void (*kern)(PathwiseArgs, const double*) = nullptr;
        switch (NT) {
          case 2: kern = a.gen_draws ? pathwise_rrm_kernel<2, true> : pathwise_rrm_kernel<2, false>; break;
          case 3: kern = a.gen_draws ? pathwise_rrm_kernel<3, true> : pathwise_rrm_kernel<3, false>; break;
          default: kern = a.gen_draws ? pathwise_rrm_kernel<4, true> : pathwise_rrm_kernel<4, false>; break;
        }
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r)) != cudaSuccess) return e;
        kern<<<pairs * a.nchunk, 256, smem_r, s>>>(a, meta);
        h->launches++;
      } else {
      const size_t ring = (size_t)kRS * kRB * kRE, fold = (size_t)4 * 2 * kST * kRT * 8;
      const size_t smem_r = sizeof(double) * (std::max(ring, fold) + 2 * kRS);
      void (*kern)(PathwiseArgs, const double*) = nullptr;
      switch ((Nq + 2 + 7) / 8) {
        case 1: kern = a.gen_draws ? pathwise_rr_kernel<1, true> : pathwise_rr_kernel<1, false>; break;
        case 2: kern = a.gen_draws ? pathwise_rr_kernel<2, true> : pathwise_rr_kernel<2, false>; break;
        case 3: kern = a.gen_draws ? pathwise_rr_kernel<3, true> : pathwise_rr_kernel<3, false>; break;
        case 4: kern = a.gen_draws ? pathwise_rr_kernel<4, true> : pathwise_rr_kernel<4, false>; break;
        case 5: kern = a.gen_draws ? pathwise_rr_kernel<5, true> : pathwise_rr_kernel<5, false>; break;
        case 6: kern = a.gen_draws ? pathwise_rr_kernel<6, true> : pathwise_rr_kernel<6, false>; break;
        case 7: kern = a.gen_draws ? pathwise_rr_kernel<7, true> : pathwise_rr_kernel<7, false>; break;
        case 8: kern = a.gen_draws ? pathwise_rr_kernel<8, true> : pathwise_rr_kernel<8, false>; break;
        case 9: kern = a.gen_draws ? pathwise_rr_kernel<9, true> : pathwise_rr_kernel<9, false>; break;
        case 10: kern = a.gen_draws ? pathwise_rr_kernel<10, true> : pathwise_rr_kernel<10, false>; break;
        default: kern = a.gen_draws ? pathwise_rr_kernel<11, true> : pathwise_rr_kernel<11, false>; break;
      }
      if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r)) != cudaSuccess) return e;
      kern<<<pairs * a.nchunk, 256, smem_r, s>>>(a, meta);
      h->launches++;
      }
    } else {
      const int PT = A <= 96 ? 3 : 6, ROWS = 32 * PT;
      const size_t main_view = (size_t)2 * ROWS * kDBP + 3 * kDB * kWS + 3 * 4 * kDB + 3 * kDB + 8 * kDB * 2 + 2 * (2 * kDB * 2);
      const size_t smem_m = sizeof(double) * std::max(main_view, (size_t)2 * kST * ROWS);
      if (smem_m > 227 * 1024) return cudaErrorInvalidValue;
      void (*kern)(PathwiseArgs, const double*) = PT == 3 ? pathwise_dmma_kernel<3> : pathwise_dmma_kernel<6>;
      if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m)) != cudaSuccess)
        return e;
      kern<<<pairs * a.nchunk, 256, smem_m, s>>>(a, meta);
      h->launches++;
    }
    if (a.S >= 64) {
      // many samples per pair: the Cholesky once per pair, then the update over (pair, sample chunk) CTAs
      if ((e = launch_gp_prepare(h, d, p, Lc, Sfull, kl_l, kvec, Linv, s)) != cudaSuccess) return e;
      const int chunk = (size_t)pairs * ((a.S + 127) / 128) >= (size_t)8 * h->num_sms ? 128 : 64;
      const int ntiles = (Nq + 7) / 8;
      if (ntiles <= 12) {
        const size_t smem_u = sizeof(double) * ((size_t)(3 * 32 + ntiles * 8) * kLD + 64 + (size_t)8 * 2 * 8 * kLD);
        void (*kern)(PathwiseArgs, const double*, int) = nullptr;
        switch (ntiles) {
          case 1: kern = pathwise_update_mma_kernel<1>; break;
          case 2: kern = pathwise_update_mma_kernel<2>; break;
          case 3: kern = pathwise_update_mma_kernel<3>; break;
          case 4: kern = pathwise_update_mma_kernel<4>; break;
          case 5: kern = pathwise_update_mma_kernel<5>; break;
          case 6: kern = pathwise_update_mma_kernel<6>; break;
          case 7: kern = pathwise_update_mma_kernel<7>; break;
          case 8: kern = pathwise_update_mma_kernel<8>; break;
          case 9: kern = pathwise_update_mma_kernel<9>; break;
          case 10: kern = pathwise_update_mma_kernel<10>; break;
          case 11: kern = pathwise_update_mma_kernel<11>; break;
          default: kern = pathwise_update_mma_kernel<12>; break;
        }
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u)) != cudaSuccess) return e;
        kern<<<pairs * ((a.S + chunk - 1) / chunk), 256, smem_u, s>>>(a, meta, chunk);
      } else {
      const size_t smem_u = sizeof(double) * (2 * 32 * LDM + (size_t)Mp * (Nq | 1) + 64 + 8 * 32);
      if ((e = cudaFuncSetAttribute(pathwise_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u)) != cudaSuccess)
        return e;
      pathwise_update_kernel<<<pairs * ((a.S + chunk - 1) / chunk), 256, smem_u, s>>>(a, meta, chunk);
      }
    } else {
      gp_prepare_update_kernel<<<pairs * a.nchunk, 128, 0, s>>>(a, p, Lc, Sfull, kl_l, kvec, Linv, meta);
    }
    h->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (a.gen_draws) {
      // should the probe find the inputs not to be an equispaced grid, the general sampler below reads the draws from
      // memory: write them now (gated on the probe), or - without buffers - poison the result
      if (have_buffers) {
        if ((e = launch_rng_fill(h, d, a.seed, a.iteration, a.problem_offset, a.sample_offset, const_cast<double*>(r.omega),
                                 const_cast<double*>(r.tau), const_cast<double*>(r.w), nullptr, nullptr, s, meta)) != cudaSuccess)
          return e;
      } else {
        poison_paths_kernel<<<2 * h->num_sms, 256, 0, s>>>(f, (size_t)d.num_problems * a.S * Nq * a.D, meta);
        h->launches++;
        return cudaGetLastError();
      }
    }
  }
  const int threads_gen = 32 * a.XG * a.KS;
  const size_t smem = sizeof(double) * ((size_t)a.KS * 2 * kST * a.XG * 32 + 2 * 32 * LDM + (size_t)Nq * Mp + kST * 32 + 64);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  e = cudaFuncSetAttribute(pathwise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  a.items = pairs * a.nchunk;
  const int grid_gen = fast ? std::min(a.items, 2 * h->num_sms) : a.items;
  pathwise_kernel<<<grid_gen, threads_gen, smem, s>>>(a, fast ? meta : nullptr);
  h->launches++;
  return cudaGetLastError();
}

// split of the reverse pass's sample loop: only when there are too few (problem, latent) pairs to fill the GPU
// How the reverse pass splits its sample loop.  S >= 64 (and N <= 96): the DMMA-batched sample kernel over (pair, chunk)
// CTAs, chunks a multiple of its 32-sample tile, then the fold kernel.  Otherwise the fused kernel, split into
// partial-sum CTAs only when there are too few (problem, latent) pairs to fill the GPU.
bool backward_plan(int num_sms, int pairs, int S, int N, int* nchunk, int* chunk) {
  const bool batched = S >= 64 && N <= 96;
  if (batched) {
    const int want = std::max(1, (4 * num_sms + pairs - 1) / std::max(1, pairs));
    const int n = std::max(1, std::min(want, (S + 63) / 64));
    const int c = (((S + n - 1) / n) + kBS - 1) / kBS * kBS;
    *chunk = c;
    *nchunk = (S + c - 1) / c;
    return true;
  }
  const int tiles = (S + kBT - 1) / kBT;
  int want = (2 * num_sms) / std::max(1, pairs);
  int n = std::max(1, std::min(tiles / 4, want));      // at least 4 tiles per chunk, else the fused kernel wins
  const int c = ((tiles + n - 1) / n) * kBT;
  *chunk = c;
  *nchunk = (S + c - 1) / c;
  return false;
}

size_t backward_partial_doubles(int num_sms, int pairs, int S, int N) {
  int n, c;
  const bool batched = backward_plan(num_sms, pairs, S, N, &n, &c);
  return (batched || n > 1) ? (size_t)pairs * n * kPartial : 0;
}

cudaError_t launch_gp_backward(vgpmp_handle* h, const vgpmp_dims& d, const vgpmp_params& p, const vgpmp_draws& r,
                               const GpScratch& ws, const vgpmp_grads& g, cudaStream_t s) {
  BackwardArgs a;
  a.D = h->robot.dof; a.M = d.num_inducing; a.N = d.num_timesteps; a.S = d.num_samples; a.B = d.num_bases;
  a.jitter = h->lik.jitter;
  a.klw = d.kl_shards > 1 ? 1.0 / (double)d.kl_shards : 1.0;
  a.Z = p.Z; a.X = p.X; a.ls = p.lengthscales; a.var = p.variances; a.q_sqrt = p.q_sqrt;
  a.eps_u = r.eps_u;
  a.Lc = ws.Lc; a.Linv = ws.Linv; a.kvec = ws.kvec; a.v = ws.v; a.f0 = ws.f0; a.h0 = ws.h0; a.df = ws.df; a.df_planar = ws.planar;
  a.d_q_mu = g.d_q_mu; a.d_q_sqrt = g.d_q_sqrt; a.d_ls = g.d_lengthscales; a.d_var = g.d_variances;
  const int Mp = a.M + 2;
  const size_t kreg = std::max((size_t)a.N * Mp, (size_t)2 * 32 * LDM);   // Kfu inside the sample loop, L | GL after it
  const size_t smem = sizeof(double) * (3 * 32 * LDM + kreg + 4 * kBT * 32 + 4 * 32 + 16 + (size_t)kBT * a.N);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  const int pairs = d.num_problems * a.D;
  const bool batched = backward_plan(h->num_sms, pairs, a.S, a.N, &a.nchunk, &a.chunk) && ws.partial != nullptr;
  a.partial = ws.partial;
  cudaError_t e;
  if (batched) {
    const int ntiles = (a.N + 7) / 8;
    const size_t smem_b = sizeof(double) * (6 * 32 * kLD + 2 * 32 * (size_t)dmma_ld(ntiles * 8) + 72);
    void (*kern)(BackwardArgs) = nullptr;
    switch (ntiles) {
      case 1: kern = gp_backward_samples_kernel<1>; break;
      case 2: kern = gp_backward_samples_kernel<2>; break;
      case 3: kern = gp_backward_samples_kernel<3>; break;
      case 4: kern = gp_backward_samples_kernel<4>; break;
      case 5: kern = gp_backward_samples_kernel<5>; break;
      case 6: kern = gp_backward_samples_kernel<6>; break;
      case 7: kern = gp_backward_samples_kernel<7>; break;
      case 8: kern = gp_backward_samples_kernel<8>; break;
      case 9: kern = gp_backward_samples_kernel<9>; break;
      case 10: kern = gp_backward_samples_kernel<10>; break;
      case 11: kern = gp_backward_samples_kernel<11>; break;
      default: kern = gp_backward_samples_kernel<12>; break;
    }
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(gp_backward_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return e;
    kern<<<pairs * a.nchunk, 256, smem_b, s>>>(a);
    gp_backward_kernel<2><<<pairs, 256, smem, s>>>(a);
    h->launches += 2;
  } else if (a.nchunk <= 1 || ws.partial == nullptr) {
    if (ws.partial == nullptr) { a.nchunk = 1; a.chunk = a.S; }
    if ((e = cudaFuncSetAttribute(gp_backward_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return e;
    gp_backward_kernel<0><<<pairs, 256, smem, s>>>(a);
    h->launches++;
  } else {
    if ((e = cudaFuncSetAttribute(gp_backward_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(gp_backward_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return e;
    gp_backward_kernel<1><<<pairs * a.nchunk, 256, smem, s>>>(a);
    gp_backward_kernel<2><<<pairs, 256, smem, s>>>(a);
    h->launches += 2;
  }
  return cudaGetLastError();
}

cudaError_t launch_predict_mean(vgpmp_handle* h, const vgpmp_dims& d, const vgpmp_params& p, const double* Xq, int Nq,
                                const double* Lc, double* mean, cudaStream_t s) {
  predict_mean_kernel<<<d.num_problems * h->robot.dof, 128, 0, s>>>(h->robot.dof, d.num_inducing, Nq, p, Xq, Lc, mean);
  h->launches++;
  return cudaGetLastError();
}

int elbo_reduce_segments(int num_sms, int Bp, int SN) {
  if (Bp >= 2 * num_sms) return 1;
  return std::max(1, std::min((SN + 4095) / 4096, (4 * num_sms + Bp - 1) / Bp));
}

cudaError_t launch_elbo_reduce(vgpmp_handle* h, const vgpmp_dims& d, const double* logp, const double* kl_l,
                               double* elbo, double* kl_out, double* loss_out, double* partial, cudaStream_t s) {
  const int stot = d.total_samples > 0 ? d.total_samples : d.num_samples;
  const double klw = d.kl_shards > 1 ? 1.0 / (double)d.kl_shards : 1.0;
  const int SN = d.num_samples * d.num_timesteps, D = h->robot.dof;
  const double scale = h->lik.alpha / (double)stot;
  const int nseg = partial != nullptr ? elbo_reduce_segments(h->num_sms, d.num_problems, SN) : 1;
  elbo_reduce_kernel<<<d.num_problems * nseg, 256, 0, s>>>(D, SN, scale, klw, logp, kl_l, elbo, kl_out, loss_out, nseg, partial);
  h->launches++;
  if (nseg > 1) {
    elbo_finish_kernel<<<(d.num_problems + 3) / 4, 128, 0, s>>>(D, d.num_problems, scale, klw, partial, nseg, kl_l, elbo,
                                                                    kl_out, loss_out);
    h->launches++;
  }
  return cudaGetLastError();
}

// CUDA-graph replay of the training step: the step counter lives in device memory; the samplers and the draw generator add
// it to their Philox iteration, Adam reads it, and this epilogue advances it.
__global__ void step_epilogue_kernel(unsigned long long* __restrict__ step) { *step += 1ull; }

cudaError_t launch_step_epilogue(vgpmp_handle* h, cudaStream_t s) {
  step_epilogue_kernel<<<1, 1, 0, s>>>(const_cast<unsigned long long*>(h->capture_iter_dev));
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_adam(vgpmp_handle* h, const vgpmp_dims& d, const vgpmp_adam& st, const vgpmp_grads& g, cudaStream_t s) {
  const int D = h->robot.dof, M = d.num_inducing;
  const size_t total = (size_t)d.num_problems * (M * D + D * M * M + 2 * D);
  adam_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(D, M, d.num_problems, st, g, h->capture_iter_dev);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_rng_fill(vgpmp_handle* h, const vgpmp_dims& d, uint64_t seed, uint64_t iteration,
                            int64_t problem_offset, int64_t sample_offset, double* omega, double* tau, double* w,
                            double* eps_u, double* eps_j, cudaStream_t s, const double* skip_if_grid) {
  RngArgs a;
  a.D = h->robot.dof; a.B = d.num_bases; a.S = d.num_samples; a.Mp = d.num_inducing + 2; a.Bp = d.num_problems;
  a.problem_offset = problem_offset; a.sample_offset = sample_offset; a.seed = seed; a.iteration = iteration;
  philox_round_keys(seed, a.rk);
  a.iter_dev = h->capture_iter_dev;
  a.omega = omega; a.tau = tau; a.w = w; a.eps_u = eps_u; a.eps_j = eps_j;
  const size_t per_pair = (omega ? (size_t)a.B : 0) + (w ? (size_t)a.S * ((a.B + 3) / 4) : 0) +
                          (eps_u ? (size_t)a.S * ((a.Mp + 1) / 2) : 0);
  if (per_pair == 0) return cudaSuccess;
  const size_t bpp = (per_pair + 255) / 256, blocks = bpp * (size_t)a.Bp * a.D;
  if (per_pair >= (1ull << 32) || blocks >= (1ull << 31)) return cudaErrorInvalidValue;
  const size_t grid = skip_if_grid != nullptr ? std::min(blocks, (size_t)8 * h->num_sms) : blocks;
  rng_fill_kernel<<<(unsigned)grid, 256, 0, s>>>(a, (uint32_t)bpp, (uint32_t)blocks, skip_if_grid);
  h->launches++;
  return cudaGetLastError();
}
