// tcgen05 / TMEM probe for the large-S sampler (standalone; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_probe tools/umma_probe.cu).
//
// Answers, on a B200, the questions the 3xTF32 sampler design rests on:
//   1. which shared-memory descriptor encoding (LBO / SBO roles) the no-swizzle K-major canonical layout takes
//      ([k/4][row][4 floats]: 8 rows x 16 B core matrices, contiguous along the row dimension);
//   2. the error of a 1024-deep FP32 TMEM accumulation against float64, for plain TF32 and for the 3-pass hi/lo split;
//   3. the sustained tcgen05.mma kind::tf32 rate for M=128, N=192, K=8 issued back to back.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
  return d;                 // base offset 0, lbo mode 0, layout type 0 = no swizzle
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                 // D format F32
  d |= 2u << 7;                 // A format TF32
  d |= 2u << 10;                // B format TF32
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;                     // K-major A and B, no negate, dense
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* b, uint32_t parity, long long limit) {
  uint32_t ok = 0;
  for (long long i = 0; i < limit && !ok; ++i)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}

// element (row, k) of an operand tile with `rows` rows: [k/4][row][k%4]
__device__ __forceinline__ int op_index(int rows, int row, int k) { return ((k >> 2) * rows + row) * 4 + (k & 3); }

constexpr int kM = 128;
constexpr int kKB = 32;     // bases per staged block

// mode 0: correctness / accuracy.  D[M,N] = A[M,K] B[N,K]^T, K a multiple of 32, `passes` = 1 (plain TF32) or 3 (hi/lo split).
// swap = 1 exchanges the LBO / SBO roles in the descriptor.
// mode 1: throughput. `reps` x (4 MMAs of K=8) on the same staged block, no waits in between.
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                           float* __restrict__ D, int N, int K, int passes, int swap,
                                                           int mode, int reps, long long* cycles, int* status) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* Ahi = reinterpret_cast<float*>(smem_raw);          // [kKB/4][128][4]
  float* Alo = Ahi + kM * kKB;
  float* Bhi = Alo + kM * kKB;                               // [kKB/4][N][4]
  float* Blo = Bhi + N * kKB;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc_tf32(kM, N);
  const uint32_t a_lbo = kM * 16, b_lbo = N * 16, sbo = 128;
  uint32_t phase = 0;
  bool ok = true;
  const int nblk = mode == 0 ? K / kKB : 1;
  for (int blk = 0; blk < nblk && ok; ++blk) {
    // stage one block of 32 bases in the canonical layout
    for (int i = tid; i < kM * kKB; i += blockDim.x) {
      const int row = i / kKB, k = i % kKB;
      const float x = A[(size_t)row * K + blk * kKB + k];
      const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
      Ahi[op_index(kM, row, k)] = hi;
      Alo[op_index(kM, row, k)] = x - hi;
    }
    for (int i = tid; i < N * kKB; i += blockDim.x) {
      const int row = i / kKB, k = i % kKB;
      const float x = B[(size_t)row * K + blk * kKB + k];
      const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
      Bhi[op_index(N, row, k)] = hi;
      Blo[op_index(N, row, k)] = x - hi;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long t0 = clock64();
      const int R = mode == 0 ? 1 : reps;
      for (int r = 0; r < R; ++r) {
        for (int ks = 0; ks < kKB / 8; ++ks) {
          // one MMA covers 8 bases = 2 chunks of 4; chunk stride (K direction) = rows * 16 B
          const uint32_t aoff = ks * 2 * a_lbo, boff = ks * 2 * b_lbo;
          const uint64_t dah = swap ? make_desc(smem_u32(Ahi) + aoff, sbo, a_lbo) : make_desc(smem_u32(Ahi) + aoff, a_lbo, sbo);
          const uint64_t dal = swap ? make_desc(smem_u32(Alo) + aoff, sbo, a_lbo) : make_desc(smem_u32(Alo) + aoff, a_lbo, sbo);
          const uint64_t dbh = swap ? make_desc(smem_u32(Bhi) + boff, sbo, b_lbo) : make_desc(smem_u32(Bhi) + boff, b_lbo, sbo);
          const uint64_t dbl = swap ? make_desc(smem_u32(Blo) + boff, sbo, b_lbo) : make_desc(smem_u32(Blo) + boff, b_lbo, sbo);
          const uint32_t acc = (blk | ks | r) != 0;
          if (passes == 3) {          // small terms first
            umma_tf32(tmem, dal, dbh, idesc, acc);
            umma_tf32(tmem, dah, dbl, idesc, 1);
            umma_tf32(tmem, dah, dbh, idesc, 1);
          } else {
            umma_tf32(tmem, dah, dbh, idesc, acc);
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      ok = mbar_wait_bounded(&bar, phase, 200000000LL);
      const long long t1 = clock64();
      if (cycles != nullptr && blockIdx.x == 0) *cycles = t1 - t0;
      if (!ok) *status = 1;
    }
    phase ^= 1;
    __syncthreads();   // (thread 0 waited for the MMAs: the staged block may be overwritten)
    ok = *((volatile int*)status) == 0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (mode == 0 && ok) {
    // epilogue: warp w owns TMEM lanes 32w .. 32w+31 (= rows of D); 16 columns per load
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int row = warp * 32 + lane;
      for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

static double urand(uint64_t& s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return ((s >> 11) + 0.5) / 9007199254740992.0; }
static double nrand(uint64_t& s) { const double u = urand(s), v = urand(s); return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v); }

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d, %d SMs, %d kHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.clockRate);
  const int N = 192;
  int* status; long long* cycles;
  CK(cudaMalloc(&status, 4)); CK(cudaMalloc(&cycles, 8));
  const size_t smem = sizeof(float) * (2 * kM * kKB + 2 * N * kKB);
  CK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  for (int test = 0; test < 3; ++test) {
    // test 0: small integers, K=32 (layout check, exact in TF32); test 1: features x weights like the sampler, K=1024
    const int K = test == 0 ? 32 : 1024;
    std::vector<float> A((size_t)kM * K), B((size_t)N * K);
    uint64_t seed = 12345 + test;
    if (test == 0) {
      for (auto& x : A) x = (float)((int)(urand(seed) * 9.0) - 4);
      for (auto& x : B) x = (float)((int)(urand(seed) * 9.0) - 4);
    } else if (test == 1) {
      const double amp = sqrt(2.0 * 0.5 / K);
      for (auto& x : A) x = (float)nrand(seed);                                            // weights w[s][b]
      for (auto& x : B) x = (float)(amp * cos(6.283185307179586 * urand(seed)));           // features phi[x][b]
    } else {                                                                               // same-sign worst case for a truncating accumulator
      for (auto& x : A) x = (float)fabs(nrand(seed));
      for (auto& x : B) x = (float)(0.03 * urand(seed));
    }
    std::vector<double> ref((size_t)kM * N, 0.0);
    for (int i = 0; i < kM; ++i)
      for (int j = 0; j < N; ++j) {
        double t = 0.0;
        for (int k = 0; k < K; ++k) t += (double)A[(size_t)i * K + k] * (double)B[(size_t)j * K + k];
        ref[(size_t)i * N + j] = t;
      }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, (size_t)kM * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    for (int swap = 0; swap < 1; ++swap)   // (swap = 1, LBO and SBO exchanged, faults: the encoding below is the right one)
      for (int passes = 1; passes <= 3; passes += 2) {
        CK(cudaMemset(status, 0, 4));
        CK(cudaMemset(dD, 0xff, (size_t)kM * N * 4));
        umma_probe_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, passes, swap, 0, 1, cycles, status);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("test %d swap %d passes %d: CUDA error %s\n", test, swap, passes, cudaGetErrorString(e)); return 1; }
        int st = 0;
        CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
        std::vector<float> Dh((size_t)kM * N);
        CK(cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0.0, maxref = 0.0, bias = 0.0;
        for (size_t i = 0; i < Dh.size(); ++i) {
          maxerr = fmax(maxerr, fabs((double)Dh[i] - ref[i]));
          maxref = fmax(maxref, fabs(ref[i]));
          bias += ((double)Dh[i] - ref[i]) * (ref[i] >= 0 ? 1.0 : -1.0);
        }
        printf("test %d K=%d swap(LBO<->SBO)=%d passes=%d: status %d  max|err| %.3e  max|ref| %.3e  rel %.3e  mean signed err (toward +|ref|) %.3e\n",
               test, K, swap, passes, st, maxerr, maxref, maxerr / maxref, bias / Dh.size());
      }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  // throughput: every SM issues reps x 4 MMAs (M=128, N=192, K=8) back to back
  {
    std::vector<float> A((size_t)kM * 32, 1.0f), B((size_t)N * 32, 1.0f);
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, (size_t)kM * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    for (int passes = 1; passes <= 3; passes += 2) {
      const int reps = 2000;
      CK(cudaMemset(status, 0, 4));
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      umma_probe_kernel<<<prop.multiProcessorCount, 128, smem>>>(dA, dB, dD, N, 32, passes, 0, 1, 10, cycles, status);
      cudaEventRecord(e0);
      umma_probe_kernel<<<prop.multiProcessorCount, 128, smem>>>(dA, dB, dD, N, 32, passes, 0, 1, reps, cycles, status);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      long long cyc = 0;
      CK(cudaMemcpy(&cyc, cycles, 8, cudaMemcpyDeviceToHost));
      const double mmas = (double)reps * 4 * passes;
      const double flops = mmas * 2.0 * kM * N * 8 * prop.multiProcessorCount;
      printf("throughput passes=%d: %.3f ms, %.1f cycles per MMA (M=128,N=%d,K=8), %.1f TFLOP/s tf32 over %d SMs\n", passes, ms,
             (double)cyc / mmas, N, flops / (ms * 1e-3) / 1e12, prop.multiProcessorCount);
    }
  }
  return 0;
}
