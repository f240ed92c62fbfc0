/*
 * vgpmp_b200 -- C-ABI of the B200-native vgpmp ELBO hot path (sm_100a).
 *
 * The reference (luke-ck/vgpmp) has no FFI: its extension points are GPflow-style
 * dispatchers and subclassing.  Every entry point below names the
 * reference Python interface it replaces (file:line under /root/reference).  A
 * maintainer binds these with ctypes; the Python package
 * vgpmp_b200/ does exactly that.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in signatures (a stream is
 *     passed as void* holding a cudaStream_t / CUstream; NULL = legacy default stream);
 *   - unless a parameter is documented "host", pointers are DEVICE pointers on the
 *     handle's device; the caller owns every buffer; the library allocates device memory
 *     only in vgpmp_create (constants + SDF grid) and works inside the caller's workspace;
 *   - every call is asynchronous on the given stream and returns 0 or a negative
 *     vgpmp_status; vgpmp_last_error() gives the text.  No exceptions cross the ABI;
 *   - a handle is bound to one device and is not thread-safe; handles are independent;
 *   - all arrays are row-major float64 with the innermost dimension as written.
 *
 * Symbols:  D = dof = latent GPs (<= 8), M inducing points, Mp = M+2 (<= 32), N timesteps,
 *           S samples, B Fourier bases, P spheres (<= 64), Bp problems in the batch.
 */
#ifndef VGPMP_B200_H
#define VGPMP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define VGPMP_MAX_DOF 8
#define VGPMP_MAX_SPHERES 64
#define VGPMP_MAX_MP 32

typedef enum {
  VGPMP_OK = 0,
  VGPMP_ERR_INVALID = -1,   /* bad argument / unsupported size */
  VGPMP_ERR_CUDA = -2,      /* CUDA runtime error (text in vgpmp_last_error) */
  VGPMP_ERR_WORKSPACE = -3, /* workspace too small */
  VGPMP_ERR_NO_DEVICE = -4  /* no usable sm_100 device */
} vgpmp_status;

typedef struct vgpmp_handle vgpmp_handle;

/* Constants left behind by Sampler.__init__ / Robot.initialise
 * (gpflow_vgpmp/utils/sampler.py:28-56, utils/robot.py:196-203,482-499,534-550).  HOST pointers. */
typedef struct {
  int32_t dof;                  /* D */
  int32_t craig;                /* craig_dh_convention (sampler.py:190-214) else Spong (:142-168) */
  int32_t num_spheres;          /* P */
  const double* dh;             /* [D,3] (d, a, alpha) */
  const double* twist;          /* [D] */
  const double* base_pose;      /* [4,4] */
  const int32_t* sphere_frame;  /* [P] index into the D+1 prefix frames = repeat(fk_slice, spheres_per_link) (sampler.py:237-244); non-decreasing */
  const double* sphere_offsets; /* [P,3] translations after Sampler.get_mat (sampler.py:68-101) */
  const double* sphere_radii;   /* [P] */
  const double* limits_lo;      /* [D] joint_limits[:,1] (likelihood.py:45-52) */
  const double* limits_hi;      /* [D] joint_limits[:,0] */
} vgpmp_robot_desc;

/* SignedDistanceField(data, origin, delta) (utils/sdf_utils.py:31-44); data[x,y,z] C-order, z fastest.  HOST pointer. */
typedef struct {
  int32_t nx, ny, nz;
  const double* data;
  double origin[3];
  double delta;
} vgpmp_sdf_desc;

/* VariationalMonteCarloLikelihood(sigma_obs, ..., offset, epsilon) (likelihoods/likelihood.py:23-55) + VGPMP.alpha */
typedef struct {
  double sigma_obs;       /* likelihood.variance holds sigma_obs itself (likelihood.py:37-41,99) */
  double epsilon;
  double alpha;
  double scene_offset[3]; /* likelihood.offset */
  double jitter;          /* gpflow.default_jitter() = 1e-6 */
} vgpmp_lik_desc;

typedef struct {
  int32_t num_problems; /* Bp */
  int32_t num_inducing; /* M  */
  int32_t num_timesteps;/* N  */
  int32_t num_samples;  /* S  */
  int32_t num_bases;    /* B  */
  /* Single-problem large-sample mode (samples sharded over GPUs): this call holds num_samples of
   * total_samples Monte-Carlo samples and 1/kl_shards of the (replicated) KL term, so that SUMMING elbo and gradients
   * over the shards (one NCCL all-reduce) gives exactly the unsharded result.  0 means "not sharded". */
  int32_t total_samples;
  int32_t kl_shards;
} vgpmp_dims;

/* Model state of a batch of planning problems (models/vgpmp.py:59-82,200-218,255-263). */
typedef struct {
  const double* q_mu;         /* [Bp,M,D]   VGPMP._q_mu (latent space) */
  const double* q_sqrt;       /* [Bp,D,M,M] VGPMP._q_sqrt, lower triangle used */
  const double* lengthscales; /* [Bp,D]     Matern52.lengthscales (constrained) */
  const double* variances;    /* [Bp,D]     Matern52.variance (constrained) */
  const double* query_latent; /* [Bp,2,D]   VGPMP._query_states = joint_sigmoid.inverse(start, goal) */
  const double* Z;            /* [M,D]      inducing inputs without the 2 conditioned timesteps (shared by the batch) */
  const double* X;            /* [N,D]      training inputs (utils/miscellaneous.py:115-127) */
} vgpmp_params;

/* d ELBO / d params, same layouts. */
typedef struct {
  double* d_q_mu;         /* [Bp,M,D] */
  double* d_q_sqrt;       /* [Bp,D,M,M] (strict upper triangle written as 0) */
  double* d_lengthscales; /* [Bp,D] */
  double* d_variances;    /* [Bp,D] */
} vgpmp_grads;

/* Random inputs of one ELBO evaluation (GPflowSampling random_fourier + exact_update). */
typedef struct {
  const double* omega; /* [Bp,D,B,D] Matern-5/2 spectral draws N(0,I)/sqrt(Gamma(5/2,5/2)), before /lengthscale */
  const double* tau;   /* [Bp,D,B]   U(0, 2pi) */
  const double* w;     /* [Bp,D,S,B] N(0,1) prior weights */
  const double* eps_u; /* [Bp,D,S,Mp] N(0,1): u = q_mu + q_sqrt eps_u */
  const double* eps_j; /* [Bp,D,S,Mp] N(0,1): err -= sqrt(jitter) eps_j */
} vgpmp_draws;

/* Optional intermediate outputs of the fused iteration (any may be NULL). */
typedef struct {
  double* f;    /* [Bp,S,N,D] latent path samples (predict_f_samples) */
  double* logp; /* [Bp,S,N]   likelihood.log_prob */
  double* kl;   /* [Bp]       prior_kl */
} vgpmp_aux;

/* Keras Adam state on the UNCONSTRAINED variables (models/vgpmp.py:77; utils/miscellaneous.py:68-84). */
typedef struct {
  double* q_mu;          /* [Bp,M,D]  trained in place (identity transform) */
  double* q_sqrt;        /* [Bp,D,M,M] trained in place (FillTriangular = the lower-triangle entries) */
  double* raw_lengthscales; /* [Bp,D] softplus^-1(lengthscale) */
  double* raw_variances;    /* [Bp,D] softplus^-1(variance - variance_lower) */
  double* lengthscales;  /* [Bp,D] constrained copies refreshed by the step */
  double* variances;     /* [Bp,D] */
  double* m;             /* first moments, packed [Bp, M*D + D*M*M + 2D] */
  double* v;             /* second moments, same packing */
  double variance_lower; /* positive(lower) shift of the variance transform */
  double learning_rate, beta1, beta2, eps;
  int32_t step;          /* number of steps already taken (t = step+1 for this call) */
  int32_t train_q_mu, train_q_sqrt, train_lengthscales, train_variances; /* utils/miscellaneous.py:324-343 */
} vgpmp_adam;

/* ---- lifetime ------------------------------------------------------------------------------- */
int vgpmp_create(vgpmp_handle** out, int device, const vgpmp_robot_desc* robot, const vgpmp_sdf_desc* sdf,
                 const vgpmp_lik_desc* lik);
/* A second handle on the SAME device-resident SDF records as `src` (ref-counted: the 4x-grid-size record array exists once
 * per device and is released with the last handle that uses it).  `robot` and `lik` (each NULL = src's) let the new handle
 * carry its own robot / likelihood constants.  For callers that drive sub-batches of problems on their own streams
 * (the reference builds ONE SignedDistanceField per environment: utils/simulation_manager.py:45-58). */
int vgpmp_create_shared(vgpmp_handle** out, const vgpmp_handle* src, const vgpmp_robot_desc* robot,
                        const vgpmp_lik_desc* lik);
/* identity of the record array a handle reads (equal for handles that share it; tests / diagnostics) */
uint64_t vgpmp_sdf_records_id(const vgpmp_handle* h);
int vgpmp_destroy(vgpmp_handle* h);
const char* vgpmp_last_error(const vgpmp_handle* h); /* NULL handle -> last create() error */
const char* vgpmp_version(void);
size_t vgpmp_workspace_bytes(const vgpmp_handle* h, const vgpmp_dims* dims);
/* number of kernel launches issued by this handle so far (bench.py "gpu_launches") */
uint64_t vgpmp_launch_count(const vgpmp_handle* h);

/* ---- stage kernels (each parity-testable in isolation) -------------------------------------- */
/* Sampler.forward_kinematics (utils/sampler.py:103-120): joints [n,D] -> frames [n,D+1,4,4] */
int vgpmp_fk_frames(vgpmp_handle* h, const double* joints, double* frames, int64_t n, void* stream);
/* Sampler.forward_kinematics_cost (utils/sampler.py:216-244): joints [n,D] -> sphere centres [n,P,3] */
int vgpmp_fk_spheres(vgpmp_handle* h, const double* joints, double* centres, int64_t n, void* stream);
/* SignedDistanceField.get_distance_tf / get_distance_grad_tf (utils/sdf_utils.py:73-76,100-136):
 * pts [n,3] (already relative to the grid, no scene offset) -> dist [n], grad [n,3] (grad may be NULL) */
int vgpmp_sdf_lookup(vgpmp_handle* h, const double* pts, double* dist, double* grad, int64_t n, void* stream);
/* VariationalMonteCarloLikelihood.log_prob (likelihoods/likelihood.py:57-176) fused with its reverse pass.
 * in [n,D]: latent f when squash!=0 (joint_sigmoid applied first, models/vgpmp.py:283) else joint angles.
 * logp [n]; d_in [n,D] = upstream * d logp / d in (NULL to skip the reverse pass). */
int vgpmp_loglik_fwd_bwd(vgpmp_handle* h, const double* in, int squash, double upstream, double* logp, double* d_in,
                         int64_t n, void* stream);
/* min over the P spheres of (sdf(x_p - scene_offset) - r_p) per joint configuration: joints [n,D] -> clearance [n].
 * The reference's "solved" verdict is pybullet motor tracking (utils/robot.py:416-480), which is outside the hot path;
 * SURVEY.md 8f-3 defines the collision-free verdict as min clearance > 0 on the best sample. */
int vgpmp_clearance(vgpmp_handle* h, const double* joints, double* clearance, int64_t n, void* stream);
/* SVGP posterior mean, `self.posterior().predict_f(X)[0]` in VGPMP.sample_from_posterior (models/vgpmp.py:315):
 * mean [Bp,Nq,D] = Kfu (Kuu + jitter I)^-1 q_mu_full at Xq [Nq,D].  Workspace as for vgpmp_gp_prepare. */
int vgpmp_predict_f_mean(vgpmp_handle* h, const vgpmp_dims* dims, const vgpmp_params* p, const double* Xq, int num_query,
                         double* mean, void* ws, size_t ws_bytes, void* stream);
/* K_conditioned / Kuu (covariances/multioutput/Kuus.py:42-53): K [Bp,D,Mp,Mp] = K(Zy,Zy) + jitter I, Zy=[0;1;Z] */
int vgpmp_kuu(vgpmp_handle* h, const double* Z, const double* lengthscales, const double* variances, double jitter,
              double* K, int num_problems, int num_inducing, void* stream);
/* Kuf (covariances/multioutput/Kufs.py:26-34): [Bp,D,Mp,N];  Kfu (covariances/Kfus.py:36-42) is its transpose */
int vgpmp_kuf(vgpmp_handle* h, const double* Z, const double* X, const double* lengthscales, const double* variances,
              double* Kuf, int num_problems, int num_inducing, int num_points, void* stream);
/* Kuu + Cholesky + VGPMP.q_sqrt un-whitening (models/vgpmp.py:208-218) + prior_kl
 * (kullback_leiblers/prior_kl.py:16-35).  Lc, q_sqrt_full: [Bp,D,Mp,Mp]; kl [Bp]; outputs may be NULL. */
int vgpmp_gp_prepare(vgpmp_handle* h, const vgpmp_dims* dims, const vgpmp_params* p, double* Lc, double* q_sqrt_full,
                     double* kl, void* ws, size_t ws_bytes, void* stream);
/* temporary_paths + predict_f_samples (models/vgpmp.py:281-282; GPflowSampling): f [Bp,S,Nq,D] at Xq [Nq,D] */
int vgpmp_pathwise_sample(vgpmp_handle* h, const vgpmp_dims* dims, const vgpmp_params* p, const vgpmp_draws* r,
                          const double* Xq, int num_query, double* f, void* ws, size_t ws_bytes, void* stream);

/* ---- fused iteration ------------------------------------------------------------------------ */
/* VGPMP.elbo (models/vgpmp.py:265-289) + the reverse pass TF autodiff performs in optimization_step
 * (utils/miscellaneous.py:68-84).  elbo [Bp]; grads may be NULL (forward only). */
int vgpmp_elbo_fwd_bwd(vgpmp_handle* h, const vgpmp_dims* dims, const vgpmp_params* p, const vgpmp_draws* r,
                       double* elbo, const vgpmp_grads* g, const vgpmp_aux* aux, void* ws, size_t ws_bytes,
                       void* stream);
/* optimizer.apply_gradients on loss = -ELBO (utils/miscellaneous.py:81-82); increments st->step on the host */
int vgpmp_adam_step(vgpmp_handle* h, const vgpmp_dims* dims, vgpmp_adam* st, const vgpmp_grads* g, void* stream);
/* Device-side draw generator (Philox4x32-10 keyed by seed/iteration): fills a vgpmp_draws-shaped set of buffers.
 * sample_offset / total_samples let several GPUs draw disjoint sample slices of one problem with a shared basis. */
int vgpmp_rng_fill(vgpmp_handle* h, const vgpmp_dims* dims, uint64_t seed, uint64_t iteration, int64_t problem_offset,
                   int64_t sample_offset, double* omega, double* tau, double* w, double* eps_u, double* eps_j,
                   void* stream);

/* Same draws, produced lazily.  eps_u / eps_j are written now; omega, tau and w are NOT: the handle remembers the key
 * (seed, iteration, offsets) and the buffers.  The next vgpmp_elbo_fwd_bwd / vgpmp_pathwise_sample that is handed exactly
 * these buffers either generates the values inside the sampler kernel (equispaced rank-1 inputs: they never touch memory)
 * or writes them into the buffers right before the first kernel that has to read them.  Values are bit-identical to
 * vgpmp_rng_fill.  The contents of omega / tau / w are unspecified for the caller afterwards.  The remembered key is
 * one-shot: the next vgpmp_elbo_fwd_bwd / vgpmp_pathwise_sample consumes it whatever buffers it is given.
 * omega = tau = w = NULL is allowed ("never materialise", nothing of size Bp*D*S*B is ever allocated): the consumer must
 * then be handed NULL for the three as well, and X / Z must be the equispaced rank-1 grids of the reference with
 * N + M + 2 within the in-kernel generating samplers' range (96 points, or 112 with >= 64 samples); if the device-side
 * probe finds otherwise the sample paths, the ELBO and all gradients come out NaN.  Replaces the per-step
 * random_fourier / exact_update draws of GPflowSampling inside tf_optimization_step (utils/miscellaneous.py:68-84). */
int vgpmp_rng_fill_lazy(vgpmp_handle* h, const vgpmp_dims* dims, uint64_t seed, uint64_t iteration, int64_t problem_offset,
                        int64_t sample_offset, double* omega, double* tau, double* w, double* eps_u, double* eps_j,
                        void* stream);

/* Draw prefetch on the handle's internal side stream (two buffer slots), so that generating step t+1's randomness
 * overlaps with step t's kernels.  Protocol per step t with slot = t & 1:
 *   vgpmp_rng_join(h, slot, stream)            stream waits until slot's draws are complete
 *   ... vgpmp_elbo_fwd_bwd / vgpmp_adam_step on `stream` reading slot's buffers ...
 *   vgpmp_rng_release(h, slot, stream)         marks the point after which slot's buffers may be overwritten
 *   vgpmp_rng_fill_async(..., slot ^ 1)        fills the other slot for iteration t+1 (waits for its last release) */
int vgpmp_rng_fill_async(vgpmp_handle* h, const vgpmp_dims* dims, uint64_t seed, uint64_t iteration,
                         int64_t problem_offset, int64_t sample_offset, double* omega, double* tau, double* w,
                         double* eps_u, double* eps_j, int slot);
int vgpmp_rng_join(vgpmp_handle* h, int slot, void* stream);
int vgpmp_rng_release(vgpmp_handle* h, int slot, void* stream);

/* ---- reference-facing step with HOST buffers (training_loop body, utils/miscellaneous.py:87-103) ---
 * Copies X_host [N,D] to the device, draws the step's randomness on the device (when draws_bytes holds two draw sets,
 * 2 * vgpmp_draws_bytes, the next step's draws are generated on a side stream while this step computes), runs ELBO forward+reverse and the Adam
 * update on the handle-resident state registered with vgpmp_adam, copies loss = -ELBO [Bp] back to loss_host and
 * synchronises the stream.  Returns after the loss is readable (like `loss = tf_optimization_step(...)`). */
int vgpmp_train_step_host(vgpmp_handle* h, const vgpmp_dims* dims, vgpmp_adam* st, const double* query_latent,
                          const double* Z, const double* X_host, double* X_dev, uint64_t seed, double* draws_ws,
                          size_t draws_bytes, const vgpmp_grads* g, double* elbo_dev, double* loss_host, void* ws,
                          size_t ws_bytes, void* stream);

/* The same step in two halves, for callers that drive several handles (sub-batches of independent problems on their own
 * streams) and want their steps to overlap: _begin enqueues everything up to the asynchronous copy of the loss and
 * returns without waiting; _end synchronises `stream` and turns the copied ELBO into loss = -ELBO.  problem_offset =
 * global index of this handle's first problem (the Philox keys use global problem indices, so a split batch reproduces
 * the unsplit one bit for bit). */
int vgpmp_train_step_host_begin(vgpmp_handle* h, const vgpmp_dims* dims, vgpmp_adam* st, const double* query_latent,
                                const double* Z, const double* X_host, double* X_dev, uint64_t seed, int64_t problem_offset,
                                double* draws_ws, size_t draws_bytes, const vgpmp_grads* g, double* elbo_dev,
                                double* loss_host, void* ws, size_t ws_bytes, void* stream);
int vgpmp_train_step_host_end(vgpmp_handle* h, const vgpmp_dims* dims, double* loss_host, void* stream);
size_t vgpmp_draws_bytes(const vgpmp_dims* dims, int dof);
/* 1 when, for equispaced rank-1 inputs of this shape (num_timesteps = number of query points), the sampler generates omega /
 * tau / w inside the kernel, i.e. vgpmp_rng_fill_lazy may be given NULL for them */
int vgpmp_sampler_generates_draws(const vgpmp_handle* h, const vgpmp_dims* dims);
/* eps_u / eps_j only: a draws_ws of this size makes vgpmp_train_step_host[_begin] use lazy draws that are never materialised */
size_t vgpmp_draws_bytes_lazy(const vgpmp_dims* dims, int dof);

/* ---- SDF producer (replaces the external SDFGen binary driven by gpflow_vgpmp/utils/gen_sdf.py:16-43) ---------------
 * Exact distance from every grid node origin + (i,j,k)*delta to a triangle soup made of CONVEX pieces, negative inside
 * any piece.  HOST pointers: tri [T,3,3]; plane [T,4] outward face planes (n, d): a point is inside a piece iff
 * n.p - d <= 0 for all faces of the piece; piece_end [num_pieces] exclusive end of each piece's (contiguous) triangle
 * range; out_host [nx,ny,nz] (z fastest), i.e. the `data` of SignedDistanceField. */
int vgpmp_mesh_to_sdf(int device, const double* tri, const double* plane, const int32_t* piece_end, int32_t num_tri,
                      int32_t num_pieces, int32_t nx, int32_t ny, int32_t nz, const double* origin, double delta,
                      double* out_host);

/* Measured FP64 FMA throughput of `device` in TFLOP/s (dependent-free DFMA chains, CUDA events): the roofline
 * denominator bench.py uses for the FP64-bound sampler stage.  Negative on error. */
double vgpmp_probe_fp64_tflops(int device);

/* ---- measurement hooks (bench.py / profiles; no reference counterpart beyond the `timing` decorator of
 * utils/miscellaneous.py:46-56) ---------------------------------------------------------------------------
 * With profiling enabled every stage launch of vgpmp_elbo_fwd_bwd / vgpmp_adam_step / vgpmp_rng_fill is bracketed by
 * CUDA events on the launching stream; vgpmp_profile_collect synchronises, returns the summed milliseconds and launch
 * counts per stage (arrays of VGPMP_NUM_STAGES) and clears the record. */
/* Tuning switches (default on unless noted; every combination is covered by the parity tests).
 *   "grid_fast_path"  X and Z rank-1 equispaced grids (always true for the reference's init_trainset / initialize_Z): the
 *                     Fourier features come from rotation recurrences; 0 forces the general per-point sincos kernel.
 *   "tc_sampler"      num_samples >= 64 and N + M + 2 <= 112: the Fourier contraction runs on the 5th-generation tensor cores
 *                     (tcgen05.mma kind::tf32 as a 3-pass hi/lo split, FP32 partial sums in TMEM folded every 64 bases);
 *                     float32-class accuracy, gated by the tolerance tests (ELBO 1e-4, gradients 1e-3).  0 = float64 DMMA.
 *   "tc_min_samples"  the sample count from which "tc_sampler" applies (default 64; measured cross-over with the float64
 *                     DMMA samplers is between 20 and 32 samples, profiles/r2_tc_sampler_ablations.txt).
 *   "rr_sampler"      register-resident warp-specialised DMMA sampler (points must fit 12 tiles of 8 rows, e.g. N=70, M=24);
 *                     0 = the variant that passes the features through shared memory.
 *   "step_graph"      vgpmp_train_step_host[_begin] captures the step's launches into a CUDA graph once per argument signature
 *                     and replays it (the iteration counter and Adam's bias-corrected rate live in device memory); 0 = plain
 *                     launches every step.  Needs lazy draws and an explicit (non-NULL) stream.
 *   "rrm_min_ctas"    9..63 samples: the register-resident sampler processes up to 4 tiles of 8 samples per CTA pass (basis
 *                     tables produced once per pass) as long as at least this many CTAs remain (default -1 = 2 per SM).
 *   "dmma_sampler"    shared-memory DMMA sampler for up to 192 points; 0 = the general kernel takes those shapes.
 *   "lazy_draws"      vgpmp_rng_fill_lazy / vgpmp_train_step_host leave omega, tau, w to be generated inside the sampler;
 *                     0 = always materialise them (vgpmp_train_step_host then prefetches the next step's set instead). */
int vgpmp_set_option(vgpmp_handle* h, const char* name, int value);
#define VGPMP_NUM_STAGES 7
int vgpmp_profile_enable(vgpmp_handle* h, int on);
int vgpmp_profile_collect(vgpmp_handle* h, double* stage_ms, int64_t* stage_launches);
const char* vgpmp_stage_name(int stage);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* VGPMP_B200_H */
