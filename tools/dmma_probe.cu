// Micro-benchmark: FP64 FMA pipe vs the FP64 tensor path (mma.sync m8n8k4 f64 and the larger sm_90+ f64 shapes) on
// sm_100a, alone and interleaved.  Answers one design question for the pathwise sampler's contraction: does DMMA run
// beside DFMA (separate pipe) or on it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/dmma_probe tools/dmma_probe.cu && gpurun_out/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
               "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// MODE 0: DFMA only (NF independent chains)   1: m8n8k4 only (NM independent accumulators)   2: both interleaved
// MODE 3: m16n8k8    4: m16n8k16   5: m16n8k8 + DFMA   6: DMMA and DFMA alternating instruction by instruction
template <int MODE, int NF, int NM>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double seed) {
  double f[NF > 0 ? NF : 1];
  double c[NM > 0 ? NM : 1][4];
  const double a = seed + threadIdx.x * 1e-9, b = 1.0 - seed * 1e-3;
  double av[8], bv[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) av[i] = a + i * 1e-7;
#pragma unroll
  for (int i = 0; i < 4; ++i) bv[i] = b + i * 1e-7;
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = i * 0.5;
#pragma unroll
  for (int i = 0; i < NM; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 0.5 * i; c[i][3] = 0.25 * i; }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2 || MODE == 5) {
#pragma unroll
      for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < NM; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    if (MODE == 3 || MODE == 5) {
#pragma unroll
      for (int i = 0; i < NM; ++i) dmma1688(c[i], av, bv);
    }
    if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < NM; ++i) dmma16816(c[i], av, bv);
    }
    if (MODE == 6) {   // fine-grained alternation: NF/NM vector ops after every DMMA (the register-resident sampler's pattern)
#pragma unroll
      for (int i = 0; i < NM; ++i) {
        dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int k = 0; k < NF / NM; ++k) f[i * (NF / NM) + k] = fma(f[i * (NF / NM) + k], a, b);
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
#pragma unroll
  for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NF, int NM>
static void run(const char* name, double fma_per_thread_iter, double mma_fma_per_warp_iter, int ctas_per_sm) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * ctas_per_sm, threads = 256, iters = 20000;
  double* out;
  cudaMalloc(&out, sizeof(double) * grid * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE, NF, NM><<<grid, threads>>>(out, 200, 0.999);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    probe<MODE, NF, NM><<<grid, threads>>>(out, iters, 0.999);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double nthreads = (double)grid * threads, nwarps = nthreads / 32;
  const double vec = 2.0 * fma_per_thread_iter * nthreads * iters / (best * 1e-3) * 1e-12;
  const double ten = 2.0 * mma_fma_per_warp_iter * nwarps * iters / (best * 1e-3) * 1e-12;
  printf("%-34s ctas/sm %d  %8.3f ms   vector %6.2f TFLOP/s   tensor %6.2f TFLOP/s   sum %6.2f  err=%s\n", name, ctas_per_sm,
         best, vec, ten, vec + ten, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  // dependent-issue latency: ONE warp per SM sub-partition (128 threads, 1 CTA/SM), 1..4 independent chains per thread
  printf("-- latency: 1 CTA/SM x 256 threads = 2 warps per sub-partition; TFLOP/s scales with chains until the pipe saturates\n");
  run<0, 1, 0>("dfma x1", 1, 0, 1);
  run<0, 2, 0>("dfma x2", 2, 0, 1);
  run<0, 4, 0>("dfma x4", 4, 0, 1);
  run<1, 0, 1>("dmma m8n8k4 x1", 0, 1 * 256.0, 1);
  run<1, 0, 2>("dmma m8n8k4 x2", 0, 2 * 256.0, 1);
  run<1, 0, 3>("dmma m8n8k4 x3", 0, 3 * 256.0, 1);
  run<1, 0, 4>("dmma m8n8k4 x4", 0, 4 * 256.0, 1);
  printf("-- alternation granularity, 2 CTAs/SM: same instruction counts, grouped vs interleaved\n");
  run<2, 16, 8>("grouped: dfma x16, then dmma x8", 16, 8 * 256.0, 2);
  run<6, 16, 8>("interleaved: (dmma, dfma, dfma) x8", 16, 8 * 256.0, 2);
  run<2, 8, 8>("grouped: dfma x8, then dmma x8", 8, 8 * 256.0, 2);
  run<6, 8, 8>("interleaved: (dmma, dfma) x8", 8, 8 * 256.0, 2);
  for (int c = 1; c <= 4; c *= 2) {
    run<0, 8, 0>("dfma x8", 8, 0, c);
    run<1, 0, 8>("dmma m8n8k4 x8", 0, 8 * 256.0, c);
    run<3, 0, 4>("dmma m16n8k8 x4", 0, 4 * 1024.0, c);
    run<4, 0, 4>("dmma m16n8k16 x4", 0, 4 * 2048.0, c);
    run<2, 8, 8>("dfma x8 + m8n8k4 x8", 8, 8 * 256.0, c);
    run<2, 8, 1>("dfma x8 + m8n8k4 x1", 8, 1 * 256.0, c);
    run<2, 8, 2>("dfma x8 + m8n8k4 x2", 8, 2 * 256.0, c);
    run<2, 4, 4>("dfma x4 + m8n8k4 x4", 4, 4 * 256.0, c);
    run<5, 8, 2>("dfma x8 + m16n8k8 x2", 8, 2 * 1024.0, c);
  }
  return 0;
}
