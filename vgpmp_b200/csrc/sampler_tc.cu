// Random-Fourier prior draw on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32 with the accumulators
// in TMEM, used for large sample counts (num_samples >= 64, N + M + 2 <= 112; BASELINE configs 4 and 5).
//
// Reference arithmetic restated (GPflowSampling random_fourier + MultioutputDenseSampler.__call__, called from
// gpflow_vgpmp/models/vgpmp.py:281-282): for every (problem, latent GP)
//     f0[s, x] = sum_b phi_b(x) w[s, b],   phi_b(x) = sqrt(2 sigma^2 / B) cos(x c_b + tau_b),  c_b = sum_d omega_bd / l,
// and h0 = d f0 / d lengthscale = t_x sum_b (amp sin(theta_xb) c_b / l) w[s, b], at x in X (N equispaced points), the two
// conditioned timesteps 0 and 1, and the M equispaced inducing inputs.  This is the only dense contraction of the path:
// per pair a [S x B] . [B x 2(N + M + 2)] product (S = 256, B = 1024, 90 points at config 5: 94 MFLOP, 2.7 TFLOP per step).
//
// Mapping.  One CTA per (pair, block of 128 samples):  D[128 samples x 2*AP] += W[128 x 8 bases] . F[2*AP x 8 bases]^T,
// AP = points padded to 16, the cos rows in columns [0, AP), the sin rows (pre-scaled by c_b / l) in [AP, 2 AP).
// float32-class accuracy from TF32 tensor cores by the 3-pass split  w f ~ w_lo f_hi + w_hi f_lo + w_hi f_hi  (hi = value
// rounded to 11 significant bits, lo = the exact float32 remainder; only lo x lo, ~2^-22, is dropped).  The tensor core adds
// into its FP32 accumulator with truncation (profiles/r2_umma_probe.txt: a 1024-deep chain drifts by ~1e-5 relative), so a
// TMEM accumulator only ever holds the partial sum of 64 bases: two TMEM regions alternate, and the worker warps fold
// every finished partial into FP32 registers with round-to-nearest adds (16 folds per output) while the next 64 bases run.
//
// Warp roles (640 threads = five warps per scheduler, so that ptxas may use 96 registers; 1 CTA per SM, persistent):
//   warps 0-15 workers.  Per stage of 32 bases: features by complex rotation along the two grids in float64 (warp =
//              segment of rows, lane = basis; start / step phasors come from the table warps), split into hi / lo and
//              stored in the UMMA canonical K-major no-swizzle layout [basis / 4][row][4]; the stage's 128 x 32 weights
//              from Philox4x32-10 (same keys as rng_fill_kernel: the trajectory is the one materialised draws give) or
//              from memory; every second stage the fold described above (warp = TMEM lane quarter x column quarter).
//              The producer side is issue / latency bound (Philox and Box-Muller are serial chains), hence 16 warps.
//   warps 16-18 table producers, one table slot each (stage t belongs to warp t mod 3): lane = basis; omega / tau draws,
//              6 lock-step sincos, segment start phasors by powers (a serial chain of ~1 200 instructions per table).
//   warp 19    lane 0 issues the MMAs (12 per stage) and commits them to the mbarriers that free the stage / publish the
//              group; the warp also owns the TMEM allocation.
#include <algorithm>
#include <cstdlib>

#include "device_utils.cuh"

namespace {

constexpr int kTM = 128;          // samples per CTA = MMA M
constexpr int kTB = 32;           // bases per stage
constexpr int kGroup = 2;         // stages per TMEM partial sum (64 bases)
#ifndef VGPMP_TC_WORKERS
#define VGPMP_TC_WORKERS 16
#endif
constexpr int kWorkers = VGPMP_TC_WORKERS;   // worker warps (8 or 16)
#ifndef VGPMP_TC_MMA_SLEEP_NS
#define VGPMP_TC_MMA_SLEEP_NS 200   // the issuing thread polls its barriers this often: a stage lasts ~2 us and two are in flight
#endif
#ifndef VGPMP_TC_TABWARPS
#define VGPMP_TC_TABWARPS 3
#endif
constexpr int kTabWarps = VGPMP_TC_TABWARPS;   // table producer warps, one table slot each (stage t is produced by warp t mod kTabWarps).
                                                // 16 + 3 + 1 = 20 warps: five per scheduler, so ptxas may use 96 registers (21 warps: 80, with spills)
constexpr int kThreads = (kWorkers + kTabWarps + 1) * 32;
constexpr int kTE = 2 * kWorkers + 10;   // doubles per basis in a table slot: segment starts [16][2] | Ex | Ez | e0 | e1 | cl | pad
constexpr int kEx = 2 * kWorkers, kEz = kEx + 2, kE0 = kEx + 4, kE1 = kEx + 6, kCl = kEx + 8;
constexpr int kWChunk = kTM * 4;  // floats per 4-basis chunk of the weight operand (LBO = 2048 B)

struct TcShape {
  int AP;          // points padded to a multiple of 16
  int nx;          // worker warps walking the query grid (the other 16 - nx walk the inducing grid)
  int perx, perz;  // rows per segment
  int blocks;      // sample blocks per pair
  int items;       // pairs * blocks
  int ablate;      // experiments only (VGPMP_TC_ABLATE): 1 no MMAs, 2 no weight draws, 4 no features, 8 no table math, 16 no fold,
                   // 32 MMA thread spins without nanosleep, 64 no proxy fence, 128 no fold hand-shake, 256 no write-out
};

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
  // K-major, no swizzle: core matrix = 8 rows x 16 B, contiguous along rows (SBO = 128 B), LBO between the two K halves
  return (uint64_t)((saddr >> 4) & 0x3fff) | (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16 | (uint64_t)(128 >> 4) << 32 |
         (uint64_t)1 << 46;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}

// value -> (hi, lo): hi keeps 11 significant bits (round to nearest), lo is the exact float32 remainder
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}

__device__ __forceinline__ void cmul(double& c, double& s, double c2, double s2) {
  const double n = c * c2 - s * s2;
  s = s * c2 + c * s2;
  c = n;
}

// NC16 = AP / 16: columns per worker thread in the fold (compile time: the running sums live in registers)
template <int NC16, bool GEN>
__global__ void __launch_bounds__(kThreads, 1) pathwise_tc_kernel(PathwiseArgs a, TcShape sh, const double* __restrict__ meta) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int AP = NC16 * 16;
  constexpr int NROW = 2 * AP;                 // MMA N: cos rows then sin rows
  constexpr int FChunk = (NROW + 1) * 4;       // floats per 4-basis chunk of the feature operand (odd row count: the 8
                                               // chunks a warp writes at once land in 8 different bank groups)
  constexpr int kWOp = (kTB / 4) * kWChunk, kFOp = (kTB / 4) * FChunk;   // floats per operand copy
  constexpr int kStage = 2 * kWOp + 2 * kFOp;                            // W hi | W lo | F hi | F lo
  float* stage0 = reinterpret_cast<float*>(smem_raw);
  double* tab = reinterpret_cast<double*>(stage0 + 2 * kStage);          // [kTabWarps][kTE][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab + kTabWarps * kTE * 32);
  uint64_t* full = bars;            // [2] stage written (16 worker warps)
  uint64_t* empty = bars + 2;       // [2] stage consumed (MMA commit)
  uint64_t* grp_full = bars + 4;    // [2] TMEM partial complete (MMA commit)
  uint64_t* grp_empty = bars + 6;   // [2] TMEM partial folded (worker warps)
  uint64_t* tab_full = bars + 8;    // [kTabWarps] table written (its table warp)
  uint64_t* tab_empty = bars + 12;  // [kTabWarps] table read (worker warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  if (meta[0] == 0.0) return;       // not an equispaced rank-1 grid: the general kernel does the sampling
  if (GEN && a.iter_dev != nullptr) a.iteration += *a.iter_dev;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = a.D, M = a.M, Nq = a.Nq, S = a.S, B = a.B, A = Nq + M + 2;
  const int T = (B + kTB - 1) / kTB;             // stages per item
  const int G = (T + kGroup - 1) / kGroup;       // TMEM partials per item
  const double t0 = meta[1], dt = meta[2], z0 = meta[3], dz = meta[4];

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(full + i, kWorkers); mbar_init(empty + i, 1);
      mbar_init(grp_full + i, 1); mbar_init(grp_empty + i, kWorkers);
    }
    for (int i = 0; i < kTabWarps; ++i) { mbar_init(tab_full + i, 1); mbar_init(tab_empty + i, kWorkers); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWorkers + kTabWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // the padding rows of the feature operands are never written again: clear both stages once
  for (int i = tid; i < 2 * kStage; i += kThreads) stage0[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  int tg = 0;    // stages issued so far by this role (all roles count alike): slot = tg & 1, use = tg >> 1
  int gg = 0;    // groups so far: region = gg & 1, use = gg >> 1

  if (warp < kWorkers) {
    // =========================================== workers ===========================================
    constexpr int kParts = kWorkers / 4;         // column parts of the fold (2: cos | sin; 4: two each)
    constexpr int CW = NROW / kParts;            // columns of the fold per thread
    constexpr int NCH = CW / 8;                  // 8-column TMEM loads per fold
    const int q = warp & 3, cp = warp >> 2;      // fold role: TMEM lane quarter, column part
    // feature role: segment `warp` of the rows, basis = lane
    int rowbase, cnt;
    if (warp < sh.nx) { const int r0 = warp * sh.perx; rowbase = r0; cnt = min(Nq, r0 + sh.perx) - r0; }
    else { const int r0 = (warp - sh.nx) * sh.perz; rowbase = Nq + 2 + r0; cnt = min(M, r0 + sh.perz) - r0; }
    const bool onx = warp < sh.nx;
    const int endpoint = kWorkers - 1 - warp;    // the last two warps (short inducing segments) also place Zy = 0 and Zy = 1
    const int fcol = (lane >> 2) * FChunk + (lane & 3);                  // this basis' column inside a feature operand
    constexpr int kQStep = kWorkers / 4;                                  // quads advance by the number of 128-thread groups
    constexpr int kQPer = 8 / kQStep;                                     // quads (Philox blocks) per thread and stage
    const int sl = tid & (kTM - 1), qb = tid >> 7;                        // weight role: sample row, first quad
    for (int item = blockIdx.x; item < sh.items; item += gridDim.x) {
      const int pl = item / sh.blocks, s0 = (item % sh.blocks) * kTM;
      const uint64_t pairkey = ((uint64_t)(pl / D) + (uint64_t)a.problem_offset) * (uint64_t)D + (uint64_t)(pl % D);
      const uint32_t B4 = ((uint32_t)B + 3) / 4;
      const double* wp = GEN ? nullptr : a.w + (size_t)pl * S * B;
      const int srow = s0 + sl;
      const uint64_t wkey = (pairkey * (uint64_t)(1u << 24) + (uint64_t)srow + (uint64_t)a.sample_offset) * B4;
      float acc[CW];
#pragma unroll
      for (int i = 0; i < CW; ++i) acc[i] = 0.f;

      int folded = 0;                             // groups of this item folded so far
      auto fold_next = [&]() {
        const int gi = gg + folded;               // global group counter of the group to fold
        mbar_wait(grp_full + (gi & 1), (gi >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((gi & 1) * 256 + cp * CW);
#pragma unroll
        for (int c = 0; c < NCH; c += 2) {
          if (sh.ablate & 16) break;
          float v0[8], v1[8];
          tmem_ld8(taddr + 8 * c, v0);
          if (c + 1 < NCH) tmem_ld8(taddr + 8 * (c + 1), v1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[8 * c + k] += v0[k];
          if (c + 1 < NCH) {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[8 * (c + 1) + k] += v1[k];
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(grp_empty + (gi & 1));
        ++folded;
      };

      for (int t = 0; t < T; ++t, ++tg) {
        const int slot = tg & 1, use = tg >> 1;
        float* Whi = stage0 + (size_t)slot * kStage;
        float* Wlo = Whi + kWOp;
        float* Fhi = Wlo + kWOp;
        float* Flo = Fhi + kFOp;
        // probe the two barriers this stage needs now, consume the answers after the weight draws: the try_wait round
        // trips (~hundreds of cycles each) then overlap with the Philox / Box-Muller chains instead of preceding them
        const int tslot = tg % kTabWarps, tuse = tg / kTabWarps;
        const uint32_t ok_empty = mbar_test(empty + slot, (use & 1) ^ 1);
        const uint32_t ok_tab = mbar_test(tab_full + tslot, tuse & 1);
        // ---- weights first (they need nothing but the key): 128 samples x 8 quads of 4 bases ----
        float4 whi[kQPer], wlo[kQPer];
#pragma unroll
        for (int j = 0; j < kQPer; ++j) {
          const int b4 = qb + kQStep * j;                     // quad inside the stage
          const uint32_t b4g = (uint32_t)t * (kTB / 4) + (uint32_t)b4;
          float z[4] = {0.f, 0.f, 0.f, 0.f};
          if (srow < S && 4 * b4g < (uint32_t)B && !(sh.ablate & 2)) {
            if (GEN) {
              normal4f(a.rk, a.iteration, 4u, wkey + b4g, z);
            } else {
              const double* src = wp + (size_t)srow * B + 4 * b4g;
#pragma unroll
              for (int k = 0; k < 4; ++k) z[k] = (float)__ldg(src + min(k, B - 1 - (int)(4 * b4g)));
            }
            if (4 * b4g + 4 > (uint32_t)B) {                  // ragged last quad (B not a multiple of 4)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (4 * b4g + k >= (uint32_t)B) z[k] = 0.f;
            }
          }
          split_tf32(z[0], whi[j].x, wlo[j].x); split_tf32(z[1], whi[j].y, wlo[j].y);
          split_tf32(z[2], whi[j].z, wlo[j].z); split_tf32(z[3], whi[j].w, wlo[j].w);
        }
        if (!ok_empty) mbar_wait(empty + slot, (use & 1) ^ 1);   // the MMAs that read this slot two stages ago are done
#pragma unroll
        for (int j = 0; j < kQPer; ++j) {
          const int b4 = qb + kQStep * j;
          *reinterpret_cast<float4*>(Whi + b4 * kWChunk + sl * 4) = whi[j];
          *reinterpret_cast<float4*>(Wlo + b4 * kWChunk + sl * 4) = wlo[j];
        }
        if (!ok_tab) mbar_wait(tab_full + tslot, tuse & 1);  // this stage's table is there
        {  // ---- features: rotation chain along this warp's rows ----
          const double* e = tab + (size_t)tslot * kTE * 32 + lane;
          double cs = e[(2 * warp) * 32], sn = e[(2 * warp + 1) * 32];
          const double cd = e[(onx ? kEx : kEz) * 32], sd = e[(onx ? kEx + 1 : kEz + 1) * 32];
          const double cl = e[kCl * 32];
          float* ph = Fhi + fcol + rowbase * 4;
          float* pl_ = Flo + fcol + rowbase * 4;
          for (int i = 0; i < ((sh.ablate & 4) ? 0 : cnt); ++i) {
            float hi, lo;
            split_tf32((float)cs, hi, lo);
            ph[i * 4] = hi; pl_[i * 4] = lo;
            split_tf32((float)(sn * cl), hi, lo);
            ph[(AP + i) * 4] = hi; pl_[(AP + i) * 4] = lo;
            cmul(cs, sn, cd, sd);
          }
          if (endpoint < 2) {                                 // the two conditioned timesteps: rows Nq, Nq + 1
            const double ec = e[(kE0 + 2 * endpoint) * 32], es = e[(kE0 + 1 + 2 * endpoint) * 32];
            float hi, lo;
            split_tf32((float)ec, hi, lo);
            Fhi[fcol + (Nq + endpoint) * 4] = hi; Flo[fcol + (Nq + endpoint) * 4] = lo;
            split_tf32((float)(es * cl), hi, lo);
            Fhi[fcol + (AP + Nq + endpoint) * 4] = hi; Flo[fcol + (AP + Nq + endpoint) * 4] = lo;
          }
        }
        if (!(sh.ablate & 64)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        __syncwarp();
        if (lane == 0) { mbar_arrive(tab_empty + tslot); mbar_arrive(full + slot); }
        // ---- fold the partial sum that finished one group ago (its MMAs were issued >= 2 stages back) ----
        const int gl = t / kGroup;
        if (!(sh.ablate & 128) && (t % kGroup == kGroup - 1 || t == T - 1) && gl >= 1 && folded < gl) fold_next();
      }
      while (!(sh.ablate & 128) && folded < G) fold_next();
      gg += G;
      // ---- write-out through shared memory: every MMA of this item has been folded, so the stages are idle.  The
      // transposes live in the weight areas only (rewritten in full by every stage; the zero padding rows of the feature
      // areas stay intact): the first half of the warps in stage 0's, the second half in stage 1's.
      constexpr int kHalfW = kWorkers / 2;
      float* tr = stage0 + (size_t)(warp / kHalfW) * kStage + (size_t)(warp % kHalfW) * 32 * (CW + 1);   // [32 samples][CW + 1]
#pragma unroll
      for (int i = 0; i < CW; ++i) tr[lane * (CW + 1) + i] = acc[i];
      __syncwarp();
      const int hf = cp * CW >= AP ? 1 : 0;                     // cos columns -> f0, sin columns -> h0
      double* dst = hf == 0 ? a.f0 : a.h0;
      if (dst != nullptr && !(sh.ablate & 256)) {
        for (int xl = lane; xl < CW; xl += 32) {
          const int x = cp * CW + xl - hf * AP;                 // point index
          if (x >= A) continue;
          double scale = 1.0;
          if (hf == 1) scale = x < Nq ? t0 + dt * x : (x < Nq + 2 ? (double)(x - Nq) : z0 + dz * (x - Nq - 2));
          const int sbase = s0 + q * 32, rmax = min(32, S - sbase);
          double* outp = dst + ((size_t)pl * S + sbase) * A + x;
          const float* trp = tr + xl;
#pragma unroll 4
          for (int r = 0; r < rmax; ++r) outp[(size_t)r * A] = (double)trp[r * (CW + 1)] * scale;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(kWorkers * 32) : "memory");   // transposes read: the stages may be refilled
    }
  } else if (warp < kWorkers + kTabWarps) {
    // =========================================== table producers ===========================================
    // warp j owns table slot j and produces the stages with (global stage counter) mod kTabWarps == j: one warp alone needs
    // ~8 000 cycles per table (a serial chain of ~1 200 mostly float64 instructions), four in flight keep ahead of the workers
    const int tw = warp - kWorkers;
    for (int item = blockIdx.x; item < sh.items; item += gridDim.x) {
      const int pl = item / sh.blocks;
      const double ell = a.ls[pl], s2 = a.var[pl];
      const double amp = sqrt(2.0 * s2 / (double)B), inv_ell = 1.0 / ell;
      const uint64_t pairkey = ((uint64_t)(pl / D) + (uint64_t)a.problem_offset) * (uint64_t)D + (uint64_t)(pl % D);
      const double* om = GEN ? nullptr : a.omega + (size_t)pl * B * D;
      const double* ta = GEN ? nullptr : a.tau + (size_t)pl * B;
      for (int t = 0; t < T; ++t, ++tg) {
        if (tg % kTabWarps != tw) continue;
        const int slot = tw, use = tg / kTabWarps;
        const int b = t * kTB + lane;
        const bool live = b < B;
        double c = 0.0, taub = 0.0;
        if (live && !(sh.ablate & 8)) {
          if (GEN) {
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
              if (d < D) c = __dadd_rn(c, __dmul_rn(z[5 + d], rs));
            uint32_t cc[4] = {(uint32_t)key, (uint32_t)(key >> 32), (uint32_t)a.iteration, 3u};
            philox4x32(cc, a.rk);
            taub = 6.283185307179586476925 * u01(cc[0], cc[1]);
          } else {
            for (int d = 0; d < D; ++d) c += om[(size_t)b * D + d];
            taub = ta[b];
          }
        }
        const double cb = c * inv_ell;
        const double ab = live ? amp : 0.0;
        const double arg[6] = {t0 * cb + taub, dt * cb, z0 * cb + taub, dz * cb, taub, cb + taub};
        double sv[6], cv[6];
        double big = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) big = fmax(big, fabs(arg[k]));
        if (sh.ablate & 8) {
#pragma unroll
          for (int k = 0; k < 6; ++k) { sv[k] = 0.0; cv[k] = 1.0; }
        } else if (big < 1048576.0) {
          sincos_bf6(arg, sv, cv);
        } else {
#pragma unroll
          for (int k = 0; k < 6; ++k) sincos(arg[k], &sv[k], &cv[k]);
        }
        // segment starts: S E^(j per) by repeated products with E^per (binary powers)
        auto power = [](double c1, double s1, int n, double& co, double& so) {
          co = 1.0; so = 0.0;
          while (n > 0) {
            if (n & 1) cmul(co, so, c1, s1);
            cmul(c1, s1, c1, s1);
            n >>= 1;
          }
        };
        double pxc, pxs, pzc, pzs;
        power(cv[1], sv[1], sh.perx, pxc, pxs);
        power(cv[3], sv[3], sh.perz, pzc, pzs);
        mbar_wait(tab_empty + slot, (use & 1) ^ 1);          // workers are done with this slot's previous table
        double* e = tab + (size_t)slot * kTE * 32 + lane;
        double cx = ab * cv[0], sx = ab * sv[0], cz = ab * cv[2], sz = ab * sv[2];
        for (int j = 0; j < kWorkers; ++j) {
          if (j < sh.nx) { e[(2 * j) * 32] = cx; e[(2 * j + 1) * 32] = sx; cmul(cx, sx, pxc, pxs); }
          else { e[(2 * j) * 32] = cz; e[(2 * j + 1) * 32] = sz; cmul(cz, sz, pzc, pzs); }
        }
        e[kEx * 32] = cv[1]; e[(kEx + 1) * 32] = sv[1];
        e[kEz * 32] = cv[3]; e[(kEz + 1) * 32] = sv[3];
        e[kE0 * 32] = ab * cv[4]; e[(kE0 + 1) * 32] = ab * sv[4];
        e[kE1 * 32] = ab * cv[5]; e[(kE1 + 1) * 32] = ab * sv[5];
        e[kCl * 32] = cb * inv_ell;
        __syncwarp();
        if (lane == 0) mbar_arrive(tab_full + slot);
      }
    }
  } else {
    // =========================================== MMA issuer ===========================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NROW >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
      constexpr uint32_t lbo_w = kWChunk * 4, lbo_f = FChunk * 4;
      const int ab = sh.ablate;
      auto wait_sleepy = [ab](uint64_t* b, uint32_t parity) {     // one thread polling: leave the issue slots to the workers
        uint32_t ok;
        for (;;) {
          asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                       : "=r"(ok) : "r"(smem_u32(b)), "r"(parity), "r"((uint32_t)VGPMP_MBAR_HINT_NS) : "memory");
          if (ok) break;
          if (!(ab & 32)) __nanosleep(VGPMP_TC_MMA_SLEEP_NS);
        }
      };
      for (int item = blockIdx.x; item < sh.items; item += gridDim.x) {
        for (int t = 0; t < T; ++t, ++tg) {
          const int slot = tg & 1, use = tg >> 1;
          const int gl = t / kGroup;
          const int gi = gg + gl;
          const bool first = t % kGroup == 0;
          if (first && !(ab & 128)) wait_sleepy(grp_empty + (gi & 1), ((gi >> 1) & 1) ^ 1);   // the partial two groups back has been folded
          wait_sleepy(full + slot, use & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = smem_u32(stage0 + (size_t)slot * kStage);
          const uint32_t whi = base, wlo = base + kWOp * 4, fhi = base + 2 * kWOp * 4, flo = fhi + kFOp * 4;
          const uint32_t dcol = tmem + (uint32_t)((gi & 1) * 256);
#pragma unroll
          for (int ks = 0; ks < kTB / 8; ++ks) {
            if (sh.ablate & 1) break;
            const uint32_t wo = ks * 2 * lbo_w, fo = ks * 2 * lbo_f;
            const uint64_t dwh = umma_desc(whi + wo, lbo_w), dwl = umma_desc(wlo + wo, lbo_w);
            const uint64_t dfh = umma_desc(fhi + fo, lbo_f), dfl = umma_desc(flo + fo, lbo_f);
            umma_tf32(dcol, dwl, dfh, idesc, (first && ks == 0) ? 0u : 1u);   // small terms first
            umma_tf32(dcol, dwh, dfl, idesc, 1u);
            umma_tf32(dcol, dwh, dfh, idesc, 1u);
          }
          umma_commit(empty + slot);
          if (!(ab & 128) && (t % kGroup == kGroup - 1 || t == T - 1)) umma_commit(grp_full + (gi & 1));
        }
        gg += G;
      }
    }
  }
  // teardown: every MMA has been folded by the workers before they leave their loop
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kWorkers + kTabWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

size_t tc_smem_bytes(int AP) {
  const size_t fchunk = (size_t)(2 * AP + 1) * 4;
  const size_t stage = 2 * (size_t)(kTB / 4) * kWChunk + 2 * (size_t)(kTB / 4) * fchunk;
  return 2 * stage * sizeof(float) + (size_t)kTabWarps * kTE * 32 * sizeof(double) + 24 * sizeof(uint64_t);
}

template <int NC16>
cudaError_t launch_nc(vgpmp_handle* h, const PathwiseArgs& a, const TcShape& sh, const double* meta, cudaStream_t s) {
  const size_t smem = tc_smem_bytes(NC16 * 16);
  void (*kern)(PathwiseArgs, TcShape, const double*) = a.gen_draws ? pathwise_tc_kernel<NC16, true> : pathwise_tc_kernel<NC16, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int grid = std::min(sh.items, h->num_sms);
  kern<<<grid, kThreads, smem, s>>>(a, sh, meta);
  h->launches++;
  return cudaGetLastError();
}

}  // namespace

bool pathwise_tc_supported(const PathwiseArgs& a) {
  const int A = a.Nq + a.M + 2;
  return a.S >= 1 && A <= 112 && a.Nq >= 2 && a.M >= 2 && a.B >= 64;   // the sample-count threshold is the caller's (tc_min_samples)
}

cudaError_t launch_pathwise_tc(vgpmp_handle* h, const PathwiseArgs& a, int pairs, const double* meta, cudaStream_t s) {
  TcShape sh;
  const int A = a.Nq + a.M + 2;
  sh.AP = (A + 15) / 16 * 16;
  sh.nx = std::max(1, std::min(kWorkers - 2, (kWorkers * a.Nq + (a.Nq + a.M) / 2) / (a.Nq + a.M)));   // >= 2 inducing segments: they also place the endpoints
  sh.perx = (a.Nq + sh.nx - 1) / sh.nx;
  sh.perz = (a.M + (kWorkers - sh.nx) - 1) / (kWorkers - sh.nx);
  sh.blocks = (a.S + kTM - 1) / kTM;
  sh.items = pairs * sh.blocks;
  { const char* ab = getenv("VGPMP_TC_ABLATE"); sh.ablate = ab ? atoi(ab) : 0; }
  switch (sh.AP / 16) {
    case 1: return launch_nc<1>(h, a, sh, meta, s);
    case 2: return launch_nc<2>(h, a, sh, meta, s);
    case 3: return launch_nc<3>(h, a, sh, meta, s);
    case 4: return launch_nc<4>(h, a, sh, meta, s);
    case 5: return launch_nc<5>(h, a, sh, meta, s);
    case 6: return launch_nc<6>(h, a, sh, meta, s);
    case 7: return launch_nc<7>(h, a, sh, meta, s);
    default: return cudaErrorInvalidValue;
  }
}
