// Device-side helpers shared by the sampler translation units (gp.cu, sampler_tc.cu): kernel argument block of the
// pathwise samplers, Philox4x32-10 draws, mbarrier hand-over, lock-step sincos.
#pragma once
#include "common.cuh"

// Philox4x32-10 round keys (seed_lo + r W0, seed_hi + r W1): computed on the host, read as constant-bank operands by the
// in-kernel generators (bumping the key in registers cost 18 of the ~100 instructions of a block of four normals)
struct PhiloxKeys {
  uint32_t k[20];
};

struct PathwiseArgs {
  int D, M, Nq, S, B, XG, KS;
  int gen_draws;       // 1: omega / tau / w are not in memory, the register-resident sampler generates them from the key below
  uint64_t seed, iteration;
  PhiloxKeys rk;       // round keys of `seed` (filled with gen_draws)
  const unsigned long long* iter_dev;   // CUDA-graph replay: the iteration lives in device memory and is ADDED to `iteration`
  int64_t problem_offset, sample_offset;
  int items;           // work items of the general sampler (pairs x nchunk)
  int nchunk, chunk;   // the S samples are split into nchunk CTAs per (problem, latent), `chunk` samples each (multiple of kST)
  double jitter;
  const double *Z, *Xq, *ls, *var, *q_mu, *query_latent;
  const double *omega, *tau, *w, *eps_u, *eps_j;
  const double *Lc, *Sfull, *Linv;
  double *f, *v, *f0, *h0;
  int f_planar;        // 0: f [Bp,S,Nq,D] (the reference's layout, public entry points); 1: f [Bp,D,S,Nq] (inside the fused step)
};

// element (problem p, sample s, point n, latent l) of the latent samples / their cotangents in either layout
__device__ __forceinline__ size_t f_index(int planar, int p, int s, int n, int l, int S, int N, int D) {
  return planar ? (((size_t)p * D + l) * S + s) * N + n : (((size_t)p * S + s) * N + n) * D + l;
}

// sampler_tc.cu: equispaced sampler on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3-pass hi/lo split,
// accumulators in TMEM).  `supported` = shape and sample count it is built for; the launcher stops at f0 / h0.
bool pathwise_tc_supported(const PathwiseArgs& a);
cudaError_t launch_pathwise_tc(vgpmp_handle* h, const PathwiseArgs& a, int pairs, const double* meta, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// Counter-based draws (Philox4x32-10).  Shared by rng_fill_kernel and by the register-resident sampler, whose producer
// warps can generate omega / tau / w in place ("lazy" draws: same keys, same arithmetic, bit-identical values).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ void philox4x32(uint32_t c[4], const PhiloxKeys& rk) {   // same block function, precomputed round keys
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ rk.k[2 * r], n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ rk.k[2 * r + 1];
    c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
  }
}
inline void philox_round_keys(uint64_t seed, PhiloxKeys& rk) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) { rk.k[2 * r] = k0; rk.k[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {  // (0,1), 53 bits
  const uint64_t x = ((uint64_t)hi << 32 | lo) >> 11;
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}
// Four independent N(0,1) from one Philox block.  The draws are random inputs, not arithmetic of the reference: the
// Box-Muller transform runs in float32 on the SFU (32-bit uniforms, |z| < 6.7) and is widened to float64.  Parity
// tests feed the *materialised* draws to the oracle, so this choice cannot leak into a parity result.
__device__ __forceinline__ void box_muller4f(const uint32_t c[4], float z[4]);
__device__ __forceinline__ void normal4f(uint64_t seed, uint64_t iter, uint32_t stream, uint64_t idx, float z[4]) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)iter, stream ^ ((uint32_t)(iter >> 32) << 8)};
  philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  box_muller4f(c, z);
}
__device__ __forceinline__ void normal4f(const PhiloxKeys& rk, uint64_t iter, uint32_t stream, uint64_t idx, float z[4]) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)iter, stream ^ ((uint32_t)(iter >> 32) << 8)};
  philox4x32(c, rk);
  box_muller4f(c, z);
}
__device__ __forceinline__ void box_muller4f(const uint32_t c[4], float z[4]) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float u1 = ((float)c[2 * k] + 0.5f) * 2.3283064365386963e-10f;       // (0,1]
    const float u2 = ((float)c[2 * k + 1] + 0.5f) * 2.3283064365386963e-10f;
    // t = -2 ln u1 > 1e-7 and r = sqrt(t) = t rsqrt(t), each ONE SFU instruction: the .ftz forms drop the denormal pre-scaling
    // of __logf / rsqrtf (4 instructions each; u1 >= 1.2e-10 and t >= 1.2e-7 are never denormal, so every bit is unchanged)
    float l2, rs;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(fminf(u1, 0.99999994f)));
    const float t = l2 * -1.3862943649291992f;                  // -2 * float(ln 2), the constant __logf multiplies by
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(t));
    const float r = t * rs;
    float sn, cs;
    __sincosf(6.283185307179586f * u2, &sn, &cs);
    z[2 * k] = r * cs;
    z[2 * k + 1] = r * sn;
  }
}
__device__ __forceinline__ void normal4(const PhiloxKeys& rk, uint64_t iter, uint32_t stream, uint64_t idx, double z[4]) {
  float zf[4];
  normal4f(rk, iter, stream, idx, zf);
#pragma unroll
  for (int k = 0; k < 4; ++k) z[k] = (double)zf[k];
}
__device__ __forceinline__ void normal4(uint64_t seed, uint64_t iter, uint32_t stream, uint64_t idx, double z[4]) {
  float zf[4];
  normal4f(seed, iter, stream, idx, zf);
#pragma unroll
  for (int k = 0; k < 4; ++k) z[k] = (double)zf[k];
}



// L2 prefetch of `count` doubles starting at `base`, one 128-byte line per participating thread and round
__device__ __forceinline__ void l2_prefetch_span(const double* base, size_t count, int tid, int nthreads) {
  const char* b0 = reinterpret_cast<const char*>(base);
  for (size_t off = (size_t)tid * 128; off < count * sizeof(double); off += (size_t)nthreads * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(b0 + off));
}

// shared-memory mbarrier helpers (producer/consumer hand-over without coupling the consumers to each other)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_test(uint64_t* b, uint32_t parity) {   // one non-blocking probe (acquire on success)
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok;
}
#ifndef VGPMP_MBAR_HINT_NS
#define VGPMP_MBAR_HINT_NS 1000000
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  // try_wait parks the thread until the phase completes or the suspend-time hint runs out; with the default (short) limit
  // a waiting warp re-issued the probe ~100 times per stage of the tensor-core sampler: 22 % of that kernel's issue slots
  uint32_t ok;
  do {
#if VGPMP_MBAR_HINT_NS > 0
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity), "r"((uint32_t)VGPMP_MBAR_HINT_NS) : "memory");
#else
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
#endif
  } while (!ok);
}

// Six branch-free double sincos evaluated in lock-step (|x| < 2^20, the caller checks and falls back): three-term
// Cody-Waite reduction by pi/2, Taylor kernels on [-pi/4, pi/4] truncated below 1e-17.  Written as loops over the six
// arguments so every Horner step issues six independent FMAs: a lone producer warp otherwise crawls through six
// back-to-back dependency chains at one instruction per FP64 latency (measured: the producers, not the DMMAs, set the
// kernel time before this).
__device__ __forceinline__ void sincos_bf6(const double (&x)[6], double (&sn)[6], double (&cs)[6]) {
  double r[6], z[6], ps[6], pc[6];
  int q[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double n = rint(x[k] * 0.63661977236758134308);
    q[k] = __double2int_rn(n);
    r[k] = fma(-n, 6.123233995736766e-17, fma(-n, 1.5707963267948966, x[k]));
    z[k] = r[k] * r[k];
    ps[k] = 1.0 / 1307674368000.0;
    pc[k] = 1.0 / 20922789888000.0;
  }
  const double S[6] = {1.0 / 6227020800.0, 1.0 / 39916800.0, 1.0 / 362880.0, 1.0 / 5040.0, 1.0 / 120.0, 1.0 / 6.0};
  const double C[7] = {1.0 / 87178291200.0, 1.0 / 479001600.0, 1.0 / 3628800.0, 1.0 / 40320.0, 1.0 / 720.0, 1.0 / 24.0, 0.5};
#pragma unroll
  for (int t = 0; t < 6; ++t)
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      ps[k] = fma(ps[k], -z[k], S[t]);
      pc[k] = fma(pc[k], -z[k], C[t]);
    }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    pc[k] = fma(pc[k], -z[k], C[6]);
    const double s = fma(-z[k] * r[k], ps[k], r[k]);
    const double c = fma(-z[k], pc[k], 1.0);
    const double a = (q[k] & 1) ? c : s, b = (q[k] & 1) ? s : c;
    sn[k] = (q[k] & 2) ? -a : a;
    cs[k] = ((q[k] + 1) & 2) ? -b : b;
  }
}

