// What can a B200 deliver for the SDF stage's access pattern?  Independent 32-byte record gathers (one aligned
// ld.global.nc.v4.f64 = one sector per lookup) at random positions of a table, no other work.  Run for a table that fits
// L2 (the 86 MiB bookshelves record grid of bench.py) and one that does not (2.4 GiB, tools/sdf_stage_bench.py).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/gather_probe tools/gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int UNROLL>
__global__ void __launch_bounds__(256) gather(const double4* __restrict__ rec, uint64_t nrec, int iters, double* out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  for (int it = 0; it < iters; ++it) {
    double4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint64_t h = ((uint64_t)hash32(tid * 2654435761u + it * UNROLL + u) << 20) ^ hash32(tid + 977u * (it * UNROLL + u));
      const double4* p = rec + h % nrec;
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[u].x), "=d"(v[u].y), "=d"(v[u].z), "=d"(v[u].w) : "l"(p));
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].w;
  }
  out[tid] = acc;
}

template <int UNROLL>
static void run(const char* name, size_t bytes, int ctas_per_sm) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const uint64_t nrec = bytes / sizeof(double4);
  double4* rec; double* out;
  cudaMalloc(&rec, nrec * sizeof(double4));
  cudaMemset(rec, 0, nrec * sizeof(double4));
  const int grid = sms * ctas_per_sm, iters = 64;
  cudaMalloc(&out, sizeof(double) * grid * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather<UNROLL><<<grid, 256>>>(rec, nrec, 4, out);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    gather<UNROLL><<<grid, 256>>>(rec, nrec, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double lookups = (double)grid * 256 * iters * UNROLL;
  printf("%-22s table %7.1f MiB  %d CTAs/SM x 256 thr, %d loads in flight/thread: %7.1f G lookups/s = %7.1f GB/s  (%s)\n", name,
         bytes / 1048576.0, ctas_per_sm, UNROLL, lookups / (best * 1e-3) * 1e-9, lookups * 32.0 / (best * 1e-3) * 1e-9,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(rec); cudaFree(out);
}

int main() {
  const size_t small = (size_t)101 * 153 * 182 * 32, big = (size_t)2400 << 20;
  run<4>("L2-resident grid", small, 8);
  run<8>("L2-resident grid", small, 8);
  run<4>("L2-resident grid", small, 3);     // the likelihood kernel's occupancy: 3 CTAs x 128 threads would be half of this
  run<4>("HBM-resident grid", big, 8);
  run<8>("HBM-resident grid", big, 8);
  run<4>("HBM-resident grid", big, 3);
  return 0;
}
