// Micro-benchmark for the 3xTF32 sampler contraction: throughput of mma.sync.m16n8k8 tf32 on sm_100a, of the
// double->float conversion, of double sincos, and how each co-issues with a DFMA stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tf32_probe tools/tf32_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// MODE 0 dfma x8 | 1 mma x8 | 2 dfma x8 + mma x8 | 3 f2f x8 | 4 f2f x8 + dfma x8 | 5 sincos x2 | 6 dfma x8 + mma x2
template <int MODE>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double seed) {
  double f[8];
  float c[8][4];
  uint32_t a[4], b[2];
  float g[8];
  double x = seed + threadIdx.x * 1e-3, sacc = 0.0;
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + i + threadIdx.x * 1e-3f) & 0xffffe000u;
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + i) & 0xffffe000u;
  const double ka = seed, kb = 1.0 - seed * 1e-3;
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = i * 0.5 + threadIdx.x; g[i] = 0.f; c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2 || MODE == 4 || MODE == 6) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fma(f[i], ka, kb);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mma_tf32(c[i], a, b);
    }
    if (MODE == 6) {
#pragma unroll
      for (int i = 0; i < 2; ++i) mma_tf32(c[i], a, b);
    }
    if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { g[i] += __double2float_rn(f[i]); f[i] = f[i] + 1.0; }   // 1 F2F + 1 FADD + 1 DADD
    }
    if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] += __double2float_rn(f[i]);                        // 1 F2F + 1 FADD beside the DFMA
    }
    if (MODE == 5) {
      double s, co;
      sincos(x, &s, &co);
      sacc += s * co;
      x += 0.37;
      sincos(x * 1.5, &s, &co);
      sacc += s - co;
    }
  }
  double s = sacc;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i] + g[i] + c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, double unit_per_thread_iter, const char* unit, int ctas_per_sm) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * ctas_per_sm, threads = 256, iters = 20000;
  double* out;
  cudaMalloc(&out, sizeof(double) * grid * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<grid, threads>>>(out, 200, 0.999);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    probe<MODE><<<grid, threads>>>(out, iters, 0.999);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double nthreads = (double)grid * threads;
  const double clk = 1.9e9;   // nominal; only used for the per-SM-per-clock column
  const double per_s = unit_per_thread_iter * nthreads * iters / (best * 1e-3);
  printf("%-28s ctas/sm %d  %8.3f ms   %10.3e %s/s   %7.1f per SM per clk @1.9GHz   err=%s\n", name, ctas_per_sm, best, per_s, unit,
         per_s / sms / clk, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int c = 2; c <= 4; c *= 2) {
    run<0>("dfma x8", 8, "dfma", c);
    run<1>("mma.tf32 m16n8k8 x8", 8 * 1024.0 / 32, "mac", c);
    run<2>("dfma x8 + mma x8 (dfma)", 8, "dfma", c);
    run<2>("dfma x8 + mma x8 (mac)", 8 * 1024.0 / 32, "mac", c);
    run<6>("dfma x8 + mma x2 (dfma)", 8, "dfma", c);
    run<3>("f2f+dadd x8", 8, "f2f", c);
    run<4>("f2f x8 + dfma x8 (pairs)", 8, "pair", c);
    run<5>("sincos x2", 2, "sincos", c);
  }
  return 0;
}
