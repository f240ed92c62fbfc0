// Internal declarations shared by the vgpmp_b200 translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <memory>
#include <string>
#include <vector>

#include "../../include/vgpmp_b200.h"

#define VG_SQRT5 2.23606797749978969641

// Robot constants passed BY VALUE as a kernel parameter: they land in the constant bank, and every
// thread of a warp walks joints/spheres in lock-step, so each access is a uniform broadcast load.
struct RobotDev {
  int dof, craig, num_spheres, pad_;
  double dh[VGPMP_MAX_DOF][3];      // d, a, alpha
  double cos_alpha[VGPMP_MAX_DOF], sin_alpha[VGPMP_MAX_DOF];
  double twist[VGPMP_MAX_DOF];
  double lo[VGPMP_MAX_DOF], hi[VGPMP_MAX_DOF];
  double base[12];                  // rows 0..2 of the 4x4 base pose
  int sphere_frame[VGPMP_MAX_SPHERES];
  int frame_end[VGPMP_MAX_DOF + 2];  // spheres of frame k are [frame_end[k-1], frame_end[k])
  double sphere_off[VGPMP_MAX_SPHERES][3];
  double sphere_rad[VGPMP_MAX_SPHERES];
};

// The SDF lives in HBM as one 32-byte record per voxel: {value, d/dx, d/dy, d/dz} with the reference's clipped central
// differences and its "exact zero -> 0.1" rule already applied (they depend only on the grid and the voxel index, so
// hoisting them to vgpmp_create is bit-identical to evaluating the 7-point stencil at lookup time).  One sphere-SDF
// evaluation = one aligned 256-bit load = one DRAM sector, instead of seven scattered 8-byte gathers.
struct SdfDev {
  const double4* rec;  // [nx,ny,nz] records, z fastest
  int nx, ny, nz, pad_;
  double origin[3];
  double delta;
  double inv_delta;   // 1/delta (index fast path, see voxel_index)
};

struct LikDev {
  double sigma_obs, epsilon, alpha, jitter;
  double offset[3];
};

struct vgpmp_handle {
  int device = 0;
  int num_sms = 0;
  RobotDev robot{};
  SdfDev sdf{};
  LikDev lik{};
  std::shared_ptr<double4> rec_owner;   // the SDF records in HBM; shared (ref-counted) by handles made with vgpmp_create_shared
  double4* rec_dev = nullptr;
  uint64_t launches = 0;
  std::string err;
  // draw prefetch: the generator for step t+1 runs on a side stream while step t computes (two buffer slots)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_filled[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  bool consumed_valid[2] = {false, false};
  // lazy draws (vgpmp_rng_fill_lazy): omega / tau / w of the set below were not written; this is how to produce them
  struct LazyDraws {
    bool valid = false;
    const double *omega = nullptr, *tau = nullptr, *w = nullptr;
    uint64_t seed = 0, iteration = 0;
    int64_t problem_offset = 0, sample_offset = 0;
    int num_problems = 0, num_samples = 0, num_bases = 0;
  } lazy;
  bool allow_lazy_draws = true;   // vgpmp_set_option("lazy_draws"): 0 makes vgpmp_rng_fill_lazy behave like vgpmp_rng_fill
  int64_t prefetched_step = -1;   // vgpmp_train_step_host: which step's draws are in flight / ready
  uint64_t prefetched_seed = 0;
  // stage profiling (bench.py): event pairs recorded on the launching stream
  bool allow_tc_path = true;    // tcgen05 / TMEM 3xTF32 sampler for large sample counts (sampler_tc.cu)
  int tc_min_samples = 64;      // ... from this many samples per problem on (below: the float64 DMMA samplers)
  bool allow_dmma_path = true;  // shared-memory DMMA sampler (N + M + 2 <= 192), else the general kernel
  bool allow_rr_path = true;    // register-resident warp-specialised DMMA sampler (<= 12 point tiles), else the shared-memory one
  // CUDA-graph replay of vgpmp_train_step_host (one captured graph per argument signature)
  const unsigned long long* capture_iter_dev = nullptr;   // non-null only while capturing: kernels read the iteration from it
  unsigned long long* step_dev = nullptr;                 // [1] step counter
  cudaGraphExec_t step_graph = nullptr;
  uint64_t step_graph_sig = 0;
  int64_t step_graph_next = -1;                           // the value the device counter holds
  int step_graph_launches = 0;
  bool allow_step_graph = true;
  int rrm_min_ctas = -1;        // multi-tile register-resident sampler only if it still launches this many CTAs (-1: 2 per SM)
  bool allow_grid_path = true;  // equispaced rank-1 fast path of the pathwise sampler (vgpmp_set_option)
  bool profiling = false;
  struct Span { int stage; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
};

enum VgStage { ST_RNG = 0, ST_PREPARE, ST_PATHWISE, ST_LOGLIK, ST_REDUCE, ST_BACKWARD, ST_ADAM };

// ---- launchers implemented in kinematics.cu -------------------------------------------------
cudaError_t launch_fk_frames(vgpmp_handle* h, const double* joints, double* frames, int64_t n, cudaStream_t s);
cudaError_t launch_fk_spheres(vgpmp_handle* h, const double* joints, double* centres, int64_t n, cudaStream_t s);
cudaError_t launch_sdf_build(vgpmp_handle* h, const double* raw_dev, cudaStream_t s);
cudaError_t launch_sdf_lookup(vgpmp_handle* h, const double* pts, double* dist, double* grad, int64_t n, cudaStream_t s);
cudaError_t launch_clearance(vgpmp_handle* h, const double* joints, double* clearance, int64_t n, cudaStream_t s);
// planar_sn = 0: in / d_in are [n,D]; else in / d_in are [n / planar_sn][D][planar_sn] (latent-major per problem)
cudaError_t launch_loglik(vgpmp_handle* h, const double* in, int squash, double upstream, double* logp, double* d_in,
                          int64_t n, int64_t planar_sn, cudaStream_t s);

// ---- launchers implemented in gp.cu ---------------------------------------------------------
struct GpScratch {  // carved from the caller's workspace by cabi.cu
  double* Lc;      // [Bp,D,Mp,Mp]
  double* Sfull;   // [Bp,D,Mp,Mp]
  double* Linv;    // [Bp,D,Mp,Mp]  explicit inverse of the Cholesky factor
  double* kl_l;    // [Bp,D]
  double* kvec;    // [Bp,D,Mp+4]  saved b = Lc^-1 (mu - p_mu) and c = K22^-1 q
  double* v;       // [Bp,D,S,Mp]
  double* f0;      // [Bp,D,S,A]   prior draws at X then Zy
  double* h0;      // [Bp,D,S,A]   d f0 / d lengthscale
  double* f;       // [Bp,D,S,N] inside the fused step (planar != 0), [Bp,S,N,D] when the caller supplied the buffer
  double* df;      // same layout as f
  int planar;
  double* logp;    // [Bp,S,N]
  double* meta;    // [8] input-structure probe {grid flag, t0, dt, z0, dz}
  double* loss;    // [Bp] -ELBO, what vgpmp_train_step_host copies back
  double* lik_part; // [Bp * segments] partial sums of the ELBO reduction (few problems, many samples)
  double* partial; // reverse-pass partial sums when the sample loop is split over CTAs (else nullptr)
};
size_t backward_partial_doubles(int num_sms, int pairs, int S, int N);

cudaError_t launch_kuu(vgpmp_handle* h, const double* Z, const double* ls, const double* var, double jitter, double* K,
                       int Bp, int M, cudaStream_t s);
cudaError_t launch_kuf(vgpmp_handle* h, const double* Z, const double* X, const double* ls, const double* var,
                       double* Kuf, int Bp, int M, int N, cudaStream_t s);
cudaError_t launch_gp_prepare(vgpmp_handle* h, const vgpmp_dims& d, const vgpmp_params& p, double* Lc, double* Sfull,
                              double* kl_l, double* kvec, double* Linv, cudaStream_t s);
cudaError_t launch_pathwise(vgpmp_handle* h, const vgpmp_dims& d, const vgpmp_params& p, const vgpmp_draws& r,
                            const double* Xq, int Nq, double* Lc, double* Sfull, double* Linv, double* kl_l, double* kvec,
                            double* f, double* v, double* f0, double* h0, double* meta, int f_planar, cudaStream_t s);
int sampler_generates_draws(const vgpmp_handle* h, const vgpmp_dims& d);
cudaError_t launch_gp_backward(vgpmp_handle* h, const vgpmp_dims& d, const vgpmp_params& p, const vgpmp_draws& r,
                               const GpScratch& ws, const vgpmp_grads& g, cudaStream_t s);
cudaError_t launch_predict_mean(vgpmp_handle* h, const vgpmp_dims& d, const vgpmp_params& p, const double* Xq, int Nq,
                                const double* Lc, double* mean, cudaStream_t s);
int elbo_reduce_segments(int num_sms, int Bp, int SN);
cudaError_t launch_elbo_reduce(vgpmp_handle* h, const vgpmp_dims& d, const double* logp, const double* kl_l,
                               double* elbo, double* kl_out, double* loss_out, double* partial, cudaStream_t s);
cudaError_t launch_adam(vgpmp_handle* h, const vgpmp_dims& d, const vgpmp_adam& st, const vgpmp_grads& g, cudaStream_t s);
cudaError_t launch_step_epilogue(vgpmp_handle* h, cudaStream_t s);
cudaError_t launch_rng_fill(vgpmp_handle* h, const vgpmp_dims& d, uint64_t seed, uint64_t iteration,
                            int64_t problem_offset, int64_t sample_offset, double* omega, double* tau, double* w,
                            double* eps_u, double* eps_j, cudaStream_t s, const double* skip_if_grid = nullptr);

// Matern-5/2 radial profile and its derivative wrt r (GPflow Matern52.K_r).
__host__ __device__ inline double vg_matern52(double r) {
  const double a = VG_SQRT5 * r;
  return (1.0 + a + (5.0 / 3.0) * r * r) * exp(-a);
}
__host__ __device__ inline double vg_matern52_dr(double r) {  // d/dr
  const double a = VG_SQRT5 * r;
  return -(5.0 / 3.0) * r * (1.0 + a) * exp(-a);
}
