// C-ABI of libvgpmp_b200.so: handle lifetime, workspace carving, argument checks, entry points.
// Every entry point documents the reference interface it replaces in include/vgpmp_b200.h.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>

#include "common.cuh"

namespace {

std::string g_create_error;

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

int fail(vgpmp_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}

int check_cuda(vgpmp_handle* h, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return VGPMP_OK;
  return fail(h, VGPMP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

cudaEvent_t take_event(vgpmp_handle* h) {
  if (!h->event_pool.empty()) {
    cudaEvent_t e = h->event_pool.back();
    h->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

// RAII bracket of one stage launch; a no-op unless profiling is on
struct StageSpan {
  vgpmp_handle* h;
  cudaStream_t s;
  cudaEvent_t b = nullptr;
  StageSpan(vgpmp_handle* h_, int stage, cudaStream_t s_) : h(h_), s(s_) {
    if (!h->profiling) return;
    cudaEvent_t a = take_event(h);
    b = take_event(h);
    cudaEventRecord(a, s);
    h->spans.push_back({stage, a, b});
  }
  ~StageSpan() {
    if (b) cudaEventRecord(b, s);
  }
};

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  double* take(size_t n_doubles) {
    double* r = reinterpret_cast<double*>(base + off);
    off = align_up(off + n_doubles * sizeof(double));
    return r;
  }
};

GpScratch carve(void* ws, int D, const vgpmp_dims& d, size_t* total, int num_sms = 148) {
  const size_t Bp = d.num_problems, M = d.num_inducing, Mp = M + 2, N = d.num_timesteps, S = d.num_samples;
  const size_t A = N + Mp;
  Carver c(ws);
  GpScratch g;
  g.Lc = c.take(Bp * D * Mp * Mp);
  g.Sfull = c.take(Bp * D * Mp * Mp);
  g.Linv = c.take(Bp * D * Mp * Mp);
  g.kl_l = c.take(Bp * D);
  g.kvec = c.take(Bp * D * (Mp + 4));
  g.v = c.take(Bp * D * S * Mp);
  g.f0 = c.take(Bp * D * S * A);
  g.h0 = c.take(Bp * D * S * A);
  g.f = c.take(Bp * S * N * D);
  g.df = c.take(Bp * S * N * D);
  g.logp = c.take(Bp * S * N);
  g.meta = c.take(8);
  g.loss = c.take(Bp);
  g.lik_part = c.take(Bp * (size_t)elbo_reduce_segments(num_sms, (int)Bp, (int)(S * N)));
  const size_t np = backward_partial_doubles(num_sms, (int)(Bp * D), (int)S, (int)N);
  g.partial = np ? c.take(np) : nullptr;
  *total = c.off;
  return g;
}

// Sampler.__init__ constants -> the by-value kernel parameter block; returns an error text or nullptr
const char* fill_robot(RobotDev& r, const vgpmp_robot_desc* robot) {
  if (robot->dof < 1 || robot->dof > VGPMP_MAX_DOF) return "dof must be in 1..8";
  if (robot->num_spheres < 1 || robot->num_spheres > VGPMP_MAX_SPHERES) return "num_spheres must be in 1..64";
  std::memset(&r, 0, sizeof(r));
  r.dof = robot->dof; r.craig = robot->craig ? 1 : 0; r.num_spheres = robot->num_spheres;
  for (int j = 0; j < r.dof; ++j) {
    for (int k = 0; k < 3; ++k) r.dh[j][k] = robot->dh[3 * j + k];
    r.cos_alpha[j] = std::cos(r.dh[j][2]);
    r.sin_alpha[j] = std::sin(r.dh[j][2]);
    r.twist[j] = robot->twist[j];
    r.lo[j] = robot->limits_lo[j];
    r.hi[j] = robot->limits_hi[j];
  }
  for (int i = 0; i < 12; ++i) r.base[i] = robot->base_pose[i];
  int prev = 0;
  for (int p = 0; p < r.num_spheres; ++p) {
    const int fr = robot->sphere_frame[p];
    if (fr < prev || fr > r.dof) return "sphere_frame must be non-decreasing and within 0..dof";
    prev = fr;
    r.sphere_frame[p] = fr;
    for (int k = 0; k < 3; ++k) r.sphere_off[p][k] = robot->sphere_offsets[3 * p + k];
    r.sphere_rad[p] = robot->sphere_radii[p];
  }
  for (int k = 0; k <= r.dof; ++k) {
    int end = 0;
    while (end < r.num_spheres && r.sphere_frame[end] <= k) ++end;
    r.frame_end[k] = end;
  }
  return nullptr;
}

int check_dims(vgpmp_handle* h, const vgpmp_dims* d) {
  if (!h || !d) return fail(h, VGPMP_ERR_INVALID, "null handle or dims");
  if (d->num_problems < 1 || d->num_inducing < 1 || d->num_timesteps < 1 || d->num_samples < 1 || d->num_bases < 1)
    return fail(h, VGPMP_ERR_INVALID, "all dims must be >= 1");
  if (d->num_inducing + 2 > VGPMP_MAX_MP) return fail(h, VGPMP_ERR_INVALID, "num_inducing + 2 must be <= 32");
  if (d->num_timesteps + d->num_inducing + 2 > 768) return fail(h, VGPMP_ERR_INVALID, "num_timesteps + Mp must be <= 768");
  if (d->num_samples >= (1 << 24) || d->total_samples >= (1 << 24))
    return fail(h, VGPMP_ERR_INVALID, "num_samples must be < 2^24");
  if (d->total_samples != 0 && d->total_samples < d->num_samples)
    return fail(h, VGPMP_ERR_INVALID, "total_samples must be 0 or >= num_samples");
  if (d->kl_shards < 0) return fail(h, VGPMP_ERR_INVALID, "kl_shards must be >= 0");
  return VGPMP_OK;
}

}  // namespace

extern "C" {

const char* vgpmp_version(void) { return "vgpmp_b200 0.1 (sm_100a, float64)"; }

const char* vgpmp_last_error(const vgpmp_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

uint64_t vgpmp_launch_count(const vgpmp_handle* h) { return h ? h->launches : 0; }

int vgpmp_create(vgpmp_handle** out, int device, const vgpmp_robot_desc* robot, const vgpmp_sdf_desc* sdf,
                 const vgpmp_lik_desc* lik) {
  if (!out || !robot || !sdf || !lik) return fail(nullptr, VGPMP_ERR_INVALID, "null argument");
  *out = nullptr;
  if (robot->dof < 1 || robot->dof > VGPMP_MAX_DOF) return fail(nullptr, VGPMP_ERR_INVALID, "dof must be in 1..8");
  if (robot->num_spheres < 1 || robot->num_spheres > VGPMP_MAX_SPHERES)
    return fail(nullptr, VGPMP_ERR_INVALID, "num_spheres must be in 1..64");
  if (sdf->nx < 1 || sdf->ny < 1 || sdf->nz < 1 || !(sdf->delta > 0.0) || !sdf->data)
    return fail(nullptr, VGPMP_ERR_INVALID, "bad SDF description");
  if (!(lik->sigma_obs > 0.0)) return fail(nullptr, VGPMP_ERR_INVALID, "sigma_obs must be > 0");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, VGPMP_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(nullptr, VGPMP_ERR_INVALID, "device index out of range");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return fail(nullptr, VGPMP_ERR_CUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, VGPMP_ERR_NO_DEVICE, "this library holds sm_100a code only; device is sm_" +
                                                  std::to_string(prop.major) + std::to_string(prop.minor));
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, VGPMP_ERR_CUDA, cudaGetErrorString(e));

  vgpmp_handle* h = new (std::nothrow) vgpmp_handle();
  if (!h) return fail(nullptr, VGPMP_ERR_INVALID, "out of host memory");
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  if (const char* msg = fill_robot(h->robot, robot)) {
    delete h;
    return fail(nullptr, VGPMP_ERR_INVALID, msg);
  }
  h->lik.sigma_obs = lik->sigma_obs; h->lik.epsilon = lik->epsilon; h->lik.alpha = lik->alpha;
  h->lik.jitter = lik->jitter;
  for (int k = 0; k < 3; ++k) h->lik.offset[k] = lik->scene_offset[k];

  const size_t cells = (size_t)sdf->nx * sdf->ny * sdf->nz;
  if (cells >= ((size_t)1 << 32)) { delete h; return fail(nullptr, VGPMP_ERR_INVALID, "SDF grids are limited to 2^32 - 1 cells (128 GiB of records)"); }
  h->sdf.nx = sdf->nx; h->sdf.ny = sdf->ny; h->sdf.nz = sdf->nz;
  for (int k = 0; k < 3; ++k) h->sdf.origin[k] = sdf->origin[k];
  h->sdf.delta = sdf->delta;
  h->sdf.inv_delta = 1.0 / sdf->delta;
  // upload the raw grid, expand it into {value, gradient} records (see SdfDev), release the raw copy
  double* raw = nullptr;
  if ((e = cudaMalloc(&raw, cells * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(&h->rec_dev, cells * sizeof(double4))) != cudaSuccess ||
      (e = cudaMemcpy(raw, sdf->data, cells * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = launch_sdf_build(h, raw, nullptr)) != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) {
    std::string msg = std::string("SDF upload: ") + cudaGetErrorString(e);
    if (raw) cudaFree(raw);
    if (h->rec_dev) cudaFree(h->rec_dev);
    delete h;
    return fail(nullptr, VGPMP_ERR_CUDA, msg);
  }
  cudaFree(raw);
  const int dev_of_records = device;
  h->rec_owner = std::shared_ptr<double4>(h->rec_dev, [dev_of_records](double4* q) {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(dev_of_records);
    cudaFree(q);
    cudaSetDevice(cur);
  });
  h->sdf.rec = h->rec_dev;
  *out = h;
  return VGPMP_OK;
}

int vgpmp_create_shared(vgpmp_handle** out, const vgpmp_handle* src, const vgpmp_robot_desc* robot,
                        const vgpmp_lik_desc* lik) {
  if (!out || !src) return fail(nullptr, VGPMP_ERR_INVALID, "create_shared: null argument");
  *out = nullptr;
  if (lik && !(lik->sigma_obs > 0.0)) return fail(nullptr, VGPMP_ERR_INVALID, "sigma_obs must be > 0");
  vgpmp_handle* h = new (std::nothrow) vgpmp_handle();
  if (!h) return fail(nullptr, VGPMP_ERR_INVALID, "out of host memory");
  h->device = src->device;
  h->num_sms = src->num_sms;
  h->robot = src->robot;
  if (robot) {
    if (const char* msg = fill_robot(h->robot, robot)) {
      delete h;
      return fail(nullptr, VGPMP_ERR_INVALID, msg);
    }
  }
  h->sdf = src->sdf;
  h->lik = src->lik;
  h->rec_owner = src->rec_owner;     // one copy of the records per device, released with the last handle
  h->rec_dev = src->rec_dev;
  if (lik) {
    h->lik.sigma_obs = lik->sigma_obs; h->lik.epsilon = lik->epsilon; h->lik.alpha = lik->alpha;
    h->lik.jitter = lik->jitter;
    for (int k = 0; k < 3; ++k) h->lik.offset[k] = lik->scene_offset[k];
  }
  *out = h;
  return VGPMP_OK;
}

uint64_t vgpmp_sdf_records_id(const vgpmp_handle* h) { return h ? (uint64_t)(uintptr_t)h->rec_dev : 0; }

int vgpmp_set_option(vgpmp_handle* h, const char* name, int value) {
  if (!h || !name) return fail(h, VGPMP_ERR_INVALID, "set_option: bad argument");
  if (std::strcmp(name, "grid_fast_path") == 0) { h->allow_grid_path = value != 0; return VGPMP_OK; }
  if (std::strcmp(name, "dmma_sampler") == 0) { h->allow_dmma_path = value != 0; return VGPMP_OK; }
  if (std::strcmp(name, "rr_sampler") == 0) { h->allow_rr_path = value != 0; return VGPMP_OK; }
  if (std::strcmp(name, "lazy_draws") == 0) { h->allow_lazy_draws = value != 0; return VGPMP_OK; }
  if (std::strcmp(name, "step_graph") == 0) { h->allow_step_graph = value != 0; return VGPMP_OK; }
  if (std::strcmp(name, "rrm_min_ctas") == 0) { h->rrm_min_ctas = value; return VGPMP_OK; }
  if (std::strcmp(name, "tc_sampler") == 0) { h->allow_tc_path = value != 0; return VGPMP_OK; }
  if (std::strcmp(name, "tc_min_samples") == 0) { h->tc_min_samples = value < 1 ? 1 : value; return VGPMP_OK; }
  return fail(h, VGPMP_ERR_INVALID, std::string("set_option: unknown option ") + name);
}

int vgpmp_profile_enable(vgpmp_handle* h, int on) {
  if (!h) return VGPMP_ERR_INVALID;
  h->profiling = on != 0;
  return VGPMP_OK;
}

int vgpmp_profile_collect(vgpmp_handle* h, double* stage_ms, int64_t* stage_launches) {
  if (!h || !stage_ms || !stage_launches) return fail(h, VGPMP_ERR_INVALID, "profile_collect: bad argument");
  for (int i = 0; i < VGPMP_NUM_STAGES; ++i) { stage_ms[i] = 0.0; stage_launches[i] = 0; }
  int rc = check_cuda(h, cudaDeviceSynchronize(), "profile_collect sync");
  for (auto& sp : h->spans) {
    float ms = 0.f;
    if (!rc && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
      stage_ms[sp.stage] += ms;
      stage_launches[sp.stage] += 1;
    }
    h->event_pool.push_back(sp.a);
    h->event_pool.push_back(sp.b);
  }
  h->spans.clear();
  return rc;
}

const char* vgpmp_stage_name(int stage) {
  static const char* names[VGPMP_NUM_STAGES] = {"rng_fill", "gp_prepare", "pathwise_sample", "loglik_fwd_bwd",
                                                "elbo_reduce", "gp_backward", "adam"};
  return (stage >= 0 && stage < VGPMP_NUM_STAGES) ? names[stage] : "?";
}

int vgpmp_destroy(vgpmp_handle* h) {
  if (!h) return VGPMP_OK;
  cudaSetDevice(h->device);
  for (auto& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (auto e : h->event_pool) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_filled[i]) cudaEventDestroy(h->ev_filled[i]);
    if (h->ev_consumed[i]) cudaEventDestroy(h->ev_consumed[i]);
  }
  if (h->side) cudaStreamDestroy(h->side);
  if (h->step_graph) cudaGraphExecDestroy(h->step_graph);
  if (h->step_dev) cudaFree(h->step_dev);
  h->rec_owner.reset();
  delete h;
  return VGPMP_OK;
}

size_t vgpmp_workspace_bytes(const vgpmp_handle* h, const vgpmp_dims* dims) {
  if (!h || !dims) return 0;
  size_t total = 0;
  carve(nullptr, h->robot.dof, *dims, &total, h->num_sms);
  return total;
}

int vgpmp_sampler_generates_draws(const vgpmp_handle* h, const vgpmp_dims* dims) {
  if (!h || !dims) return 0;
  return sampler_generates_draws(h, *dims);
}

size_t vgpmp_draws_bytes_lazy(const vgpmp_dims* d, int dof) {
  if (!d) return 0;
  const size_t Bp = d->num_problems, D = dof, S = d->num_samples, Mp = d->num_inducing + 2;
  return 2 * align_up(Bp * D * S * Mp * 8);
}

size_t vgpmp_draws_bytes(const vgpmp_dims* d, int dof) {
  if (!d) return 0;
  const size_t Bp = d->num_problems, D = dof, B = d->num_bases, S = d->num_samples, Mp = d->num_inducing + 2;
  return align_up(Bp * D * B * D * 8) + align_up(Bp * D * B * 8) + align_up(Bp * D * S * B * 8) +
         2 * align_up(Bp * D * S * Mp * 8);
}

int vgpmp_fk_frames(vgpmp_handle* h, const double* joints, double* frames, int64_t n, void* stream) {
  if (!h || n < 0 || (n > 0 && (!joints || !frames))) return fail(h, VGPMP_ERR_INVALID, "fk_frames: bad argument");
  return check_cuda(h, launch_fk_frames(h, joints, frames, n, (cudaStream_t)stream), "fk_frames");
}

int vgpmp_fk_spheres(vgpmp_handle* h, const double* joints, double* centres, int64_t n, void* stream) {
  if (!h || n < 0 || (n > 0 && (!joints || !centres))) return fail(h, VGPMP_ERR_INVALID, "fk_spheres: bad argument");
  return check_cuda(h, launch_fk_spheres(h, joints, centres, n, (cudaStream_t)stream), "fk_spheres");
}

int vgpmp_sdf_lookup(vgpmp_handle* h, const double* pts, double* dist, double* grad, int64_t n, void* stream) {
  if (!h || n < 0 || (n > 0 && (!pts || !dist))) return fail(h, VGPMP_ERR_INVALID, "sdf_lookup: bad argument");
  return check_cuda(h, launch_sdf_lookup(h, pts, dist, grad, n, (cudaStream_t)stream), "sdf_lookup");
}

int vgpmp_loglik_fwd_bwd(vgpmp_handle* h, const double* in, int squash, double upstream, double* logp, double* d_in,
                         int64_t n, void* stream) {
  if (!h || n < 0 || (n > 0 && (!in || !logp))) return fail(h, VGPMP_ERR_INVALID, "loglik_fwd_bwd: bad argument");
  return check_cuda(h, launch_loglik(h, in, squash, upstream, logp, d_in, n, 0, (cudaStream_t)stream), "loglik_fwd_bwd");
}

int vgpmp_clearance(vgpmp_handle* h, const double* joints, double* clearance, int64_t n, void* stream) {
  if (!h || n < 0 || (n > 0 && (!joints || !clearance))) return fail(h, VGPMP_ERR_INVALID, "clearance: bad argument");
  return check_cuda(h, launch_clearance(h, joints, clearance, n, (cudaStream_t)stream), "clearance");
}

int vgpmp_predict_f_mean(vgpmp_handle* h, const vgpmp_dims* dims, const vgpmp_params* p, const double* Xq, int num_query,
                         double* mean, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_dims(h, dims);
  if (rc) return rc;
  if (!p || !Xq || !mean || !ws || num_query < 1) return fail(h, VGPMP_ERR_INVALID, "predict_f_mean: bad argument");
  size_t need = 0;
  GpScratch g = carve(ws, h->robot.dof, *dims, &need, h->num_sms);
  if (ws_bytes < need) return fail(h, VGPMP_ERR_WORKSPACE, "predict_f_mean: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = check_cuda(h, launch_gp_prepare(h, *dims, *p, g.Lc, g.Sfull, g.kl_l, g.kvec, g.Linv, s), "gp_prepare"))) return rc;
  return check_cuda(h, launch_predict_mean(h, *dims, *p, Xq, num_query, g.Lc, mean, s), "predict_f_mean");
}

int vgpmp_kuu(vgpmp_handle* h, const double* Z, const double* lengthscales, const double* variances, double jitter,
              double* K, int num_problems, int num_inducing, void* stream) {
  if (!h || !Z || !lengthscales || !variances || !K || num_problems < 1 || num_inducing < 1 ||
      num_inducing + 2 > VGPMP_MAX_MP)
    return fail(h, VGPMP_ERR_INVALID, "kuu: bad argument");
  return check_cuda(h, launch_kuu(h, Z, lengthscales, variances, jitter, K, num_problems, num_inducing,
                                  (cudaStream_t)stream), "kuu");
}

int vgpmp_kuf(vgpmp_handle* h, const double* Z, const double* X, const double* lengthscales, const double* variances,
              double* Kuf, int num_problems, int num_inducing, int num_points, void* stream) {
  if (!h || !Z || !X || !lengthscales || !variances || !Kuf || num_problems < 1 || num_inducing < 1 || num_points < 1 ||
      num_inducing + 2 > VGPMP_MAX_MP)
    return fail(h, VGPMP_ERR_INVALID, "kuf: bad argument");
  return check_cuda(h, launch_kuf(h, Z, X, lengthscales, variances, Kuf, num_problems, num_inducing, num_points,
                                  (cudaStream_t)stream), "kuf");
}

int vgpmp_gp_prepare(vgpmp_handle* h, const vgpmp_dims* dims, const vgpmp_params* p, double* Lc, double* q_sqrt_full,
                     double* kl, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_dims(h, dims);
  if (rc) return rc;
  if (!p || !ws) return fail(h, VGPMP_ERR_INVALID, "gp_prepare: null params or workspace");
  size_t need = 0;
  GpScratch g = carve(ws, h->robot.dof, *dims, &need, h->num_sms);
  if (ws_bytes < need) return fail(h, VGPMP_ERR_WORKSPACE, "gp_prepare: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  rc = check_cuda(h, launch_gp_prepare(h, *dims, *p, Lc ? Lc : g.Lc, q_sqrt_full ? q_sqrt_full : g.Sfull, g.kl_l,
                                       g.kvec, g.Linv, s), "gp_prepare");
  if (rc || !kl) return rc;
  // kl[p] = sum_l kl_l: reuse the ELBO reducer with an empty likelihood term
  vgpmp_dims d0 = *dims;
  d0.num_timesteps = 0;
  return check_cuda(h, launch_elbo_reduce(h, d0, g.logp, g.kl_l, g.f0 /*scratch for -kl*/, kl, nullptr, nullptr, s), "kl_reduce");
}

int vgpmp_pathwise_sample(vgpmp_handle* h, const vgpmp_dims* dims, const vgpmp_params* p, const vgpmp_draws* r,
                          const double* Xq, int num_query, double* f, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_dims(h, dims);
  if (rc) return rc;
  if (!p || !r || !Xq || !f || !ws || num_query < 1) return fail(h, VGPMP_ERR_INVALID, "pathwise_sample: bad argument");
  vgpmp_dims dq = *dims;
  dq.num_timesteps = num_query;
  if ((rc = check_dims(h, &dq))) return rc;
  size_t need = 0;
  GpScratch g = carve(ws, h->robot.dof, dq, &need, h->num_sms);
  if (ws_bytes < need) return fail(h, VGPMP_ERR_WORKSPACE, "pathwise_sample: workspace too small (size it with num_timesteps = num_query)");
  cudaStream_t s = (cudaStream_t)stream;
  return check_cuda(h, launch_pathwise(h, dq, *p, *r, Xq, num_query, g.Lc, g.Sfull, g.Linv, g.kl_l, g.kvec, f, nullptr, g.f0, nullptr, g.meta, 0, s),
                    "pathwise_sample");
}

int vgpmp_elbo_fwd_bwd(vgpmp_handle* h, const vgpmp_dims* dims, const vgpmp_params* p, const vgpmp_draws* r,
                       double* elbo, const vgpmp_grads* gr, const vgpmp_aux* aux, void* ws, size_t ws_bytes,
                       void* stream) {
  int rc = check_dims(h, dims);
  if (rc) return rc;
  if (!p || !r || !elbo || !ws) return fail(h, VGPMP_ERR_INVALID, "elbo_fwd_bwd: bad argument");
  size_t need = 0;
  GpScratch g = carve(ws, h->robot.dof, *dims, &need, h->num_sms);
  if (ws_bytes < need) return fail(h, VGPMP_ERR_WORKSPACE, "elbo_fwd_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int D = h->robot.dof;
  const int64_t ncfg = (int64_t)dims->num_problems * dims->num_samples * dims->num_timesteps;
  // Inside the step the samples and their cotangents are latent-major, [Bp,D,S,N]: the GP kernels (one CTA per problem and
  // latent) and the likelihood (lanes = consecutive timesteps) then read and write contiguous runs.  A caller who asks for
  // the samples (aux->f) gets the reference's [Bp,S,N,D].
  const int planar = (aux && aux->f) ? 0 : 1;
  double* f = (aux && aux->f) ? aux->f : g.f;
  double* logp = (aux && aux->logp) ? aux->logp : g.logp;
  const bool bwd = gr != nullptr;
  {
    // GP preparation kernel + pathwise sampling kernels (one stage for the profile)
    StageSpan sp(h, ST_PATHWISE, s);
    if ((rc = check_cuda(h, launch_pathwise(h, *dims, *p, *r, p->X, dims->num_timesteps, g.Lc, g.Sfull, g.Linv, g.kl_l,
                                            g.kvec, f, bwd ? g.v : nullptr, g.f0, bwd ? g.h0 : nullptr,
                                            g.meta, planar, s), "pathwise")))
      return rc;
  }
  {
    StageSpan sp(h, ST_LOGLIK, s);
    if ((rc = check_cuda(h, launch_loglik(h, f, 1, h->lik.alpha / (double)(dims->total_samples > 0 ? dims->total_samples : dims->num_samples), logp,
                                          bwd ? g.df : nullptr, ncfg,
                                          planar ? (int64_t)dims->num_samples * dims->num_timesteps : 0, s), "loglik")))
      return rc;
  }
  {
    StageSpan sp(h, ST_REDUCE, s);
    if ((rc = check_cuda(h, launch_elbo_reduce(h, *dims, logp, g.kl_l, elbo, aux ? aux->kl : nullptr, g.loss, g.lik_part, s), "elbo_reduce")))
      return rc;
  }
  if (bwd) {
    GpScratch gs = g;
    gs.f = f;
    gs.planar = planar;
    StageSpan sp(h, ST_BACKWARD, s);
    if ((rc = check_cuda(h, launch_gp_backward(h, *dims, *p, *r, gs, *gr, s), "gp_backward"))) return rc;
  }
  (void)D;
  return VGPMP_OK;
}

int vgpmp_adam_step(vgpmp_handle* h, const vgpmp_dims* dims, vgpmp_adam* st, const vgpmp_grads* g, void* stream) {
  int rc = check_dims(h, dims);
  if (rc) return rc;
  if (!st || !g) return fail(h, VGPMP_ERR_INVALID, "adam_step: bad argument");
  {
    StageSpan sp(h, ST_ADAM, (cudaStream_t)stream);
    rc = check_cuda(h, launch_adam(h, *dims, *st, *g, (cudaStream_t)stream), "adam_step");
  }
  if (!rc) st->step += 1;
  return rc;
}

int vgpmp_rng_fill(vgpmp_handle* h, const vgpmp_dims* dims, uint64_t seed, uint64_t iteration, int64_t problem_offset,
                   int64_t sample_offset, double* omega, double* tau, double* w, double* eps_u, double* eps_j,
                   void* stream) {
  int rc = check_dims(h, dims);
  if (rc) return rc;
  if ((omega == nullptr) != (tau == nullptr) || (eps_u == nullptr) != (eps_j == nullptr))
    return fail(h, VGPMP_ERR_INVALID, "rng_fill: omega/tau and eps_u/eps_j come in pairs");
  if (h->lazy.valid && (omega == h->lazy.omega || w == h->lazy.w)) h->lazy.valid = false;   // these buffers are real again
  StageSpan sp(h, ST_RNG, (cudaStream_t)stream);
  return check_cuda(h, launch_rng_fill(h, *dims, seed, iteration, problem_offset, sample_offset, omega, tau, w, eps_u,
                                       eps_j, (cudaStream_t)stream), "rng_fill");
}

int vgpmp_rng_fill_lazy(vgpmp_handle* h, const vgpmp_dims* dims, uint64_t seed, uint64_t iteration, int64_t problem_offset,
                        int64_t sample_offset, double* omega, double* tau, double* w, double* eps_u, double* eps_j,
                        void* stream) {
  if (!h) return VGPMP_ERR_INVALID;
  const bool no_buffers = !omega && !tau && !w;     // never materialised: only an in-kernel generating sampler can consume this set
  if (!no_buffers && (!h->allow_lazy_draws || !omega || !tau || !w))
    return vgpmp_rng_fill(h, dims, seed, iteration, problem_offset, sample_offset, omega, tau, w, eps_u, eps_j, stream);
  int rc = check_dims(h, dims);
  if (rc) return rc;
  if ((eps_u == nullptr) != (eps_j == nullptr)) return fail(h, VGPMP_ERR_INVALID, "rng_fill_lazy: eps_u/eps_j come in pairs");
  {
    StageSpan sp(h, ST_RNG, (cudaStream_t)stream);
    rc = check_cuda(h, launch_rng_fill(h, *dims, seed, iteration, problem_offset, sample_offset, nullptr, nullptr, nullptr,
                                       eps_u, eps_j, (cudaStream_t)stream), "rng_fill_lazy");
  }
  if (rc) return rc;
  h->lazy.valid = true;
  h->lazy.omega = omega; h->lazy.tau = tau; h->lazy.w = w;
  h->lazy.seed = seed; h->lazy.iteration = iteration;
  h->lazy.problem_offset = problem_offset; h->lazy.sample_offset = sample_offset;
  h->lazy.num_problems = dims->num_problems; h->lazy.num_samples = dims->num_samples; h->lazy.num_bases = dims->num_bases;
  return VGPMP_OK;
}

static int ensure_side(vgpmp_handle* h) {
  if (h->side) return VGPMP_OK;
  cudaError_t e = cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&h->ev_filled[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_consumed[i], cudaEventDisableTiming);
  }
  return check_cuda(h, e, "side stream");
}

int vgpmp_rng_fill_async(vgpmp_handle* h, const vgpmp_dims* dims, uint64_t seed, uint64_t iteration,
                         int64_t problem_offset, int64_t sample_offset, double* omega, double* tau, double* w,
                         double* eps_u, double* eps_j, int slot) {
  if (h && h->lazy.valid && (omega == h->lazy.omega || w == h->lazy.w)) h->lazy.valid = false;
  int rc = check_dims(h, dims);
  if (rc) return rc;
  if (slot < 0 || slot > 1 || !omega || !tau || !w || !eps_u || !eps_j)
    return fail(h, VGPMP_ERR_INVALID, "rng_fill_async: bad argument");
  if ((rc = ensure_side(h))) return rc;
  if (h->consumed_valid[slot] &&
      (rc = check_cuda(h, cudaStreamWaitEvent(h->side, h->ev_consumed[slot], 0), "rng_fill_async wait")))
    return rc;
  {
    StageSpan sp(h, ST_RNG, h->side);
    if ((rc = check_cuda(h, launch_rng_fill(h, *dims, seed, iteration, problem_offset, sample_offset, omega, tau, w,
                                            eps_u, eps_j, h->side), "rng_fill_async")))
      return rc;
  }
  return check_cuda(h, cudaEventRecord(h->ev_filled[slot], h->side), "rng_fill_async record");
}

int vgpmp_rng_join(vgpmp_handle* h, int slot, void* stream) {
  if (!h || slot < 0 || slot > 1 || !h->side) return fail(h, VGPMP_ERR_INVALID, "rng_join: nothing in flight");
  return check_cuda(h, cudaStreamWaitEvent((cudaStream_t)stream, h->ev_filled[slot], 0), "rng_join");
}

int vgpmp_rng_release(vgpmp_handle* h, int slot, void* stream) {
  if (!h || slot < 0 || slot > 1) return fail(h, VGPMP_ERR_INVALID, "rng_release: bad slot");
  int rc = ensure_side(h);
  if (rc) return rc;
  h->consumed_valid[slot] = true;
  return check_cuda(h, cudaEventRecord(h->ev_consumed[slot], (cudaStream_t)stream), "rng_release");
}

// The body of one host-buffer step, enqueued on `s`.  With graph_mode the iteration number and Adam's bias-corrected rate are
// read from device memory (h->capture_iter_dev / capture_lr_dev), so that the enqueued work is the same for every step and
// can be captured once and replayed.
static int enqueue_train_step(vgpmp_handle* h, const vgpmp_dims* dims, vgpmp_adam* st, const double* query_latent, const double* Z,
                              const double* X_host, double* X_dev, uint64_t seed, int64_t problem_offset, double* draws_ws,
                              size_t draws_bytes, const vgpmp_grads* g, double* elbo_dev, double* loss_host, void* ws,
                              size_t ws_bytes, cudaStream_t s, bool graph_mode) {
  const int D = h->robot.dof;
  const bool compact = draws_bytes < vgpmp_draws_bytes(dims, D);
  const size_t Bp = dims->num_problems, B = dims->num_bases, S = dims->num_samples, Mp = dims->num_inducing + 2;
  int rc;
  void* stream = (void*)s;
  if ((rc = check_cuda(h, cudaMemcpyAsync(X_dev, X_host, sizeof(double) * dims->num_timesteps * D, cudaMemcpyHostToDevice, s),
                       "H2D X")))
    return rc;
  const size_t one_set = vgpmp_draws_bytes(dims, D);
  const bool pipelined = !graph_mode && draws_bytes >= 2 * one_set;   // two draw sets: step t+1 is drawn on the side stream during step t
  const int slot = pipelined ? (st->step & 1) : 0;
  const uint64_t iteration = graph_mode ? 0 : (uint64_t)st->step;     // graph mode: the kernels add the device-resident counter
  double *omega, *tau, *w, *eps_u, *eps_j;
  auto carve_set = [&](int which) {
    Carver c(static_cast<char*>(static_cast<void*>(draws_ws)) + (size_t)which * one_set);
    omega = compact ? nullptr : c.take(Bp * D * B * D);
    tau = compact ? nullptr : c.take(Bp * D * B);
    w = compact ? nullptr : c.take(Bp * D * S * B);
    eps_u = c.take(Bp * D * S * Mp);
    eps_j = c.take(Bp * D * S * Mp);
  };
  carve_set(slot);
  const bool lazy_ok = h->allow_lazy_draws;   // draws generated inside the sampler: no prefetch pipeline needed
  if (lazy_ok) {
    if ((rc = vgpmp_rng_fill_lazy(h, dims, seed, iteration, problem_offset, 0, omega, tau, w, eps_u, eps_j, stream))) return rc;
  } else if (pipelined && h->prefetched_step == (int64_t)st->step && h->prefetched_seed == seed) {
    if ((rc = vgpmp_rng_join(h, slot, stream))) return rc;
  } else {
    if ((rc = vgpmp_rng_fill(h, dims, seed, iteration, problem_offset, 0, omega, tau, w, eps_u, eps_j, stream))) return rc;
  }
  vgpmp_params p{st->q_mu, st->q_sqrt, st->lengthscales, st->variances, query_latent, Z, X_dev};
  vgpmp_draws r{omega, tau, w, eps_u, eps_j};
  if ((rc = vgpmp_elbo_fwd_bwd(h, dims, &p, &r, elbo_dev, g, nullptr, ws, ws_bytes, stream))) return rc;
  if ((rc = vgpmp_adam_step(h, dims, st, g, stream))) return rc;   // st->step is now the NEXT step
  if (graph_mode && (rc = check_cuda(h, launch_step_epilogue(h, s), "step epilogue"))) return rc;
  if (pipelined && !lazy_ok) {
    if ((rc = vgpmp_rng_release(h, slot, stream))) return rc;
    carve_set(slot ^ 1);
    if ((rc = vgpmp_rng_fill_async(h, dims, seed, (uint64_t)st->step, problem_offset, 0, omega, tau, w, eps_u, eps_j, slot ^ 1))) return rc;
    h->prefetched_step = st->step;
    h->prefetched_seed = seed;
  }
  {
    size_t need = 0;
    GpScratch gsc = carve(ws, D, *dims, &need, h->num_sms);   // loss = -ELBO was written by the ELBO reduction
    if ((rc = check_cuda(h, cudaMemcpyAsync(loss_host, gsc.loss, sizeof(double) * Bp, cudaMemcpyDeviceToHost, s), "D2H loss")))
      return rc;
  }
  return VGPMP_OK;
}

static uint64_t mix64(uint64_t hsh, uint64_t v) {
  hsh ^= v + 0x9E3779B97F4A7C15ull + (hsh << 6) + (hsh >> 2);
  return hsh;
}

int vgpmp_train_step_host_begin(vgpmp_handle* h, const vgpmp_dims* dims, vgpmp_adam* st, const double* query_latent,
                                const double* Z, const double* X_host, double* X_dev, uint64_t seed, int64_t problem_offset,
                                double* draws_ws, size_t draws_bytes, const vgpmp_grads* g, double* elbo_dev,
                                double* loss_host, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_dims(h, dims);
  if (rc) return rc;
  if (!st || !query_latent || !Z || !X_host || !X_dev || !draws_ws || !g || !elbo_dev || !loss_host || !ws)
    return fail(h, VGPMP_ERR_INVALID, "train_step_host: bad argument");
  if (dims->total_samples > 0 || dims->kl_shards > 1)
    return fail(h, VGPMP_ERR_INVALID, "train_step_host: a sample-sharded step needs the all-reduce between the reverse pass and Adam; "
                                      "drive it with vgpmp_elbo_fwd_bwd + vgpmp_adam_step");
  const int D = h->robot.dof;
  // a buffer that only holds eps_u / eps_j (vgpmp_draws_bytes_lazy) selects lazy draws that are never materialised
  const bool compact = draws_bytes < vgpmp_draws_bytes(dims, D);
  if (compact && (draws_bytes < vgpmp_draws_bytes_lazy(dims, D) || !h->allow_lazy_draws))
    return fail(h, VGPMP_ERR_WORKSPACE, "train_step_host: draws buffer too small");
  cudaStream_t s = (cudaStream_t)stream;
  // CUDA-graph replay: with lazy draws every step enqueues the same launches on the same buffers; only the iteration number
  // changes, and that lives in device memory.  Captured once per argument signature, replayed with one cudaGraphLaunch.
  const bool want_graph = h->allow_step_graph && h->allow_lazy_draws && !h->profiling && s != nullptr;
  if (!want_graph)
    return enqueue_train_step(h, dims, st, query_latent, Z, X_host, X_dev, seed, problem_offset, draws_ws, draws_bytes, g,
                              elbo_dev, loss_host, ws, ws_bytes, s, false);
  uint64_t sig = 0x1234567ull;
  const uint64_t words[] = {(uint64_t)dims->num_problems, (uint64_t)dims->num_inducing, (uint64_t)dims->num_timesteps,
                            (uint64_t)dims->num_samples, (uint64_t)dims->num_bases, (uint64_t)(uintptr_t)st->q_mu,
                            (uint64_t)(uintptr_t)st->q_sqrt, (uint64_t)(uintptr_t)st->raw_lengthscales,
                            (uint64_t)(uintptr_t)st->raw_variances, (uint64_t)(uintptr_t)st->lengthscales,
                            (uint64_t)(uintptr_t)st->variances, (uint64_t)(uintptr_t)st->m, (uint64_t)(uintptr_t)st->v,
                            (uint64_t)(uintptr_t)query_latent, (uint64_t)(uintptr_t)Z, (uint64_t)(uintptr_t)X_host,
                            (uint64_t)(uintptr_t)X_dev, seed, (uint64_t)problem_offset, (uint64_t)(uintptr_t)draws_ws,
                            (uint64_t)draws_bytes, (uint64_t)(uintptr_t)g->d_q_mu, (uint64_t)(uintptr_t)g->d_q_sqrt,
                            (uint64_t)(uintptr_t)g->d_lengthscales, (uint64_t)(uintptr_t)g->d_variances,
                            (uint64_t)(uintptr_t)elbo_dev, (uint64_t)(uintptr_t)loss_host, (uint64_t)(uintptr_t)ws,
                            (uint64_t)ws_bytes, (uint64_t)(uintptr_t)s, (uint64_t)st->train_q_mu, (uint64_t)st->train_q_sqrt,
                            (uint64_t)st->train_lengthscales, (uint64_t)st->train_variances,
                            (uint64_t)(h->allow_tc_path | (h->allow_rr_path << 1) | (h->allow_dmma_path << 2) | (h->allow_grid_path << 3)) |
                                ((uint64_t)h->tc_min_samples << 8),
                            (uint64_t)(int64_t)h->rrm_min_ctas};
  for (uint64_t wv : words) sig = mix64(sig, wv);
  double dw[6] = {st->learning_rate, st->beta1, st->beta2, st->eps, st->variance_lower, 0.0};
  for (double dv : dw) { uint64_t bits; std::memcpy(&bits, &dv, 8); sig = mix64(sig, bits); }
  if (h->step_dev == nullptr) {
    if ((rc = check_cuda(h, cudaMalloc(&h->step_dev, 32), "step counter"))) return rc;
  }
  if (h->step_graph == nullptr || h->step_graph_sig != sig) {
    if (h->step_graph) { cudaGraphExecDestroy(h->step_graph); h->step_graph = nullptr; }
    h->capture_iter_dev = h->step_dev;
    const uint64_t launches0 = h->launches;
    const int step0 = st->step;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    int rc2 = VGPMP_OK;
    if (e == cudaSuccess) {
      rc2 = enqueue_train_step(h, dims, st, query_latent, Z, X_host, X_dev, seed, problem_offset, draws_ws, draws_bytes, g,
                               elbo_dev, loss_host, ws, ws_bytes, s, true);
      e = cudaStreamEndCapture(s, &graph);
    }
    h->capture_iter_dev = nullptr;
    st->step = step0;                                   // nothing ran yet
    h->step_graph_launches = (int)(h->launches - launches0);
    h->launches = launches0;
    if (e == cudaSuccess && rc2 == VGPMP_OK && graph != nullptr) e = cudaGraphInstantiate(&h->step_graph, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess || rc2 != VGPMP_OK || h->step_graph == nullptr) {
      // capture is an optimisation: fall back to plain launches for good on this handle
      (void)cudaGetLastError();
      h->step_graph = nullptr;
      h->allow_step_graph = false;
      return enqueue_train_step(h, dims, st, query_latent, Z, X_host, X_dev, seed, problem_offset, draws_ws, draws_bytes, g,
                                elbo_dev, loss_host, ws, ws_bytes, s, false);
    }
    h->step_graph_sig = sig;
    h->step_graph_next = -1;
  }
  if (h->step_graph_next != (int64_t)st->step) {        // (re)synchronise the device-resident counter with the caller's step
    const unsigned long long v = (unsigned long long)st->step;
    if ((rc = check_cuda(h, cudaMemcpyAsync(h->step_dev, &v, sizeof(v), cudaMemcpyHostToDevice, s), "step counter upload"))) return rc;
    if ((rc = check_cuda(h, cudaStreamSynchronize(s), "step counter upload"))) return rc;   // `v` is a stack variable
  }
  if ((rc = check_cuda(h, cudaGraphLaunch(h->step_graph, s), "graph launch"))) return rc;
  h->launches += (uint64_t)h->step_graph_launches;
  st->step += 1;
  h->step_graph_next = st->step;
  h->lazy.valid = false;
  return VGPMP_OK;   // the loss copy is in flight: vgpmp_train_step_host_end waits for it
}

int vgpmp_train_step_host_end(vgpmp_handle* h, const vgpmp_dims* dims, double* loss_host, void* stream) {
  int rc = check_dims(h, dims);
  if (rc) return rc;
  if (!loss_host) return fail(h, VGPMP_ERR_INVALID, "train_step_host_end: bad argument");
  return check_cuda(h, cudaStreamSynchronize((cudaStream_t)stream), "sync");   // idempotent: the loss was negated on the device
}

int vgpmp_train_step_host(vgpmp_handle* h, const vgpmp_dims* dims, vgpmp_adam* st, const double* query_latent,
                          const double* Z, const double* X_host, double* X_dev, uint64_t seed, double* draws_ws,
                          size_t draws_bytes, const vgpmp_grads* g, double* elbo_dev, double* loss_host, void* ws,
                          size_t ws_bytes, void* stream) {
  int rc = vgpmp_train_step_host_begin(h, dims, st, query_latent, Z, X_host, X_dev, seed, 0, draws_ws, draws_bytes, g,
                                       elbo_dev, loss_host, ws, ws_bytes, stream);
  if (rc) return rc;
  return vgpmp_train_step_host_end(h, dims, loss_host, stream);
}

}  // extern "C"
