// Forward kinematics, sphere placement, nearest-voxel SDF lookup with its 7-point stencil, hinge
// likelihood and the fused reverse pass back to the latent path samples.  sm_100a, float64.
//
// Reference semantics restated here (files under /root/reference):
//   utils/sampler.py:103-120,142-244     DH transforms, prefix product, sphere placement
//   utils/sdf_utils.py:62-76,100-136     idx = clip(trunc((x-origin)/delta)), value, central differences, 0 -> 0.1
//   likelihoods/likelihood.py:86-176     d = sdf(x-offset)-r, custom gradient, hinge, -1/2 sum h^2/sigma_obs
//   models/vgpmp.py:283                  joint_sigmoid squash
//
// Design: one thread per (problem, sample, timestep) configuration; the DH chain lives in registers.  The reverse
// pass never materialises a Jacobian: every sphere contributes a wrench (g, x cross g) and joint j receives
//   dtheta_j = z_j . tau_{>=j} - (z_j x o_j) . F_{>=j}
// (z_j, o_j = joint axis / a point on it in world coordinates), evaluated with running prefix sums so the chain is
// walked exactly once.  Lanes of a warp are consecutive timesteps of one sample: neighbouring lanes hit neighbouring
// voxels, robot constants come from the constant bank as warp-uniform loads.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;

struct Frame {
  double r[9];
  double t[3];
};

__device__ __forceinline__ void frame_from_base(const RobotDev& rb, Frame& A) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    A.r[3 * i + 0] = rb.base[4 * i + 0];
    A.r[3 * i + 1] = rb.base[4 * i + 1];
    A.r[3 * i + 2] = rb.base[4 * i + 2];
    A.t[i] = rb.base[4 * i + 3];
  }
}

// A <- A * T_j(theta)   (Spong: sampler.py:159-164, Craig: sampler.py:205-210)
__device__ __forceinline__ void frame_step(const RobotDev& rb, int j, double theta, Frame& A) {
  double st, ct;
  sincos(theta + rb.twist[j], &st, &ct);
  const double d = rb.dh[j][0], a = rb.dh[j][1];
  const double ca = rb.cos_alpha[j], sa = rb.sin_alpha[j];
  double T[12];
  if (rb.craig) {
    T[0] = ct;      T[1] = -st;     T[2] = 0.0;  T[3] = a;
    T[4] = st * ca; T[5] = ct * ca; T[6] = -sa;  T[7] = -d * sa;
    T[8] = st * sa; T[9] = ct * sa; T[10] = ca;  T[11] = d * ca;
  } else {
    T[0] = ct;  T[1] = -st * ca; T[2] = st * sa;  T[3] = a * ct;
    T[4] = st;  T[5] = ct * ca;  T[6] = -ct * sa; T[7] = a * st;
    T[8] = 0.0; T[9] = sa;       T[10] = ca;      T[11] = d;
  }
  Frame B;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double r0 = A.r[3 * i], r1 = A.r[3 * i + 1], r2 = A.r[3 * i + 2];
    B.r[3 * i + 0] = r0 * T[0] + r1 * T[4] + r2 * T[8];
    B.r[3 * i + 1] = r0 * T[1] + r1 * T[5] + r2 * T[9];
    B.r[3 * i + 2] = r0 * T[2] + r1 * T[6] + r2 * T[10];
    B.t[i] = r0 * T[3] + r1 * T[7] + r2 * T[11] + A.t[i];
  }
  A = B;
}

__device__ __forceinline__ int clamp_index(double q, int n) {
  // clip(trunc(q), 0, n-1); the pre-clamp keeps the double->int conversion defined for far-away points
  q = fmin(fmax(q, -1.0), (double)n);
  int i = (int)q;  // truncation toward zero, like tf.cast(float64 -> int64)
  return min(max(i, 0), n - 1);
}

struct Voxel {
  int ix, iy, iz;
};

__device__ __forceinline__ Voxel sdf_voxel(const SdfDev& s, double x, double y, double z) {
  Voxel v;
  v.ix = clamp_index((x - s.origin[0]) / s.delta, s.nx);
  v.iy = clamp_index((y - s.origin[1]) / s.delta, s.ny);
  v.iz = clamp_index((z - s.origin[2]) / s.delta, s.nz);
  return v;
}

__device__ __forceinline__ double sdf_at(const SdfDev& s, int ix, int iy, int iz) {
  return __ldg(s.grid + ((size_t)ix * s.ny + iy) * s.nz + iz);
}

// value + gradient: the 7-point stencil of one sphere-SDF evaluation (7 grid elements = 56 B algorithmic)
__device__ __forceinline__ double sdf_value_grad(const SdfDev& s, double x, double y, double z, double g[3]) {
  const Voxel v = sdf_voxel(s, x, y, z);
  const int xp = min(v.ix + 1, s.nx - 1), xm = max(v.ix - 1, 0);
  const int yp = min(v.iy + 1, s.ny - 1), ym = max(v.iy - 1, 0);
  const int zp = min(v.iz + 1, s.nz - 1), zm = max(v.iz - 1, 0);
  // issue all seven loads before any use
  const double c = sdf_at(s, v.ix, v.iy, v.iz);
  const double ax = sdf_at(s, xp, v.iy, v.iz), bx = sdf_at(s, xm, v.iy, v.iz);
  const double ay = sdf_at(s, v.ix, yp, v.iz), by = sdf_at(s, v.ix, ym, v.iz);
  const double az = sdf_at(s, v.ix, v.iy, zp), bz = sdf_at(s, v.ix, v.iy, zm);
  const double den = 2.0 * s.delta;
  double gx = (ax - bx) / den, gy = (ay - by) / den, gz = (az - bz) / den;
  g[0] = (gx == 0.0) ? 0.1 : gx;  // sdf_utils.py:124,129,135
  g[1] = (gy == 0.0) ? 0.1 : gy;
  g[2] = (gz == 0.0) ? 0.1 : gz;
  return c;
}

__global__ void __launch_bounds__(kThreads) fk_frames_kernel(RobotDev rb, const double* __restrict__ joints,
                                                            double* __restrict__ frames, int64_t n) {
  const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (c >= n) return;
  Frame A;
  frame_from_base(rb, A);
  double* out = frames + c * (rb.dof + 1) * 16;
  for (int j = 0; j <= rb.dof; ++j) {
    if (j > 0) frame_step(rb, j - 1, joints[c * rb.dof + j - 1], A);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      out[4 * i + 0] = A.r[3 * i];
      out[4 * i + 1] = A.r[3 * i + 1];
      out[4 * i + 2] = A.r[3 * i + 2];
      out[4 * i + 3] = A.t[i];
    }
    out[12] = 0.0; out[13] = 0.0; out[14] = 0.0; out[15] = 1.0;
    out += 16;
  }
}

__global__ void __launch_bounds__(kThreads) fk_spheres_kernel(RobotDev rb, const double* __restrict__ joints,
                                                             double* __restrict__ centres, int64_t n) {
  const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (c >= n) return;
  Frame A;
  frame_from_base(rb, A);
  double* out = centres + c * rb.num_spheres * 3;
  int p = 0;
  for (int j = 0; j <= rb.dof; ++j) {
    if (j > 0) frame_step(rb, j - 1, joints[c * rb.dof + j - 1], A);
    while (p < rb.num_spheres && rb.sphere_frame[p] == j) {
      const double ox = rb.sphere_off[p][0], oy = rb.sphere_off[p][1], oz = rb.sphere_off[p][2];
#pragma unroll
      for (int i = 0; i < 3; ++i) out[3 * p + i] = A.r[3 * i] * ox + A.r[3 * i + 1] * oy + A.r[3 * i + 2] * oz + A.t[i];
      ++p;
    }
  }
}

__global__ void __launch_bounds__(kThreads) sdf_lookup_kernel(SdfDev sdf, const double* __restrict__ pts,
                                                             double* __restrict__ dist, double* __restrict__ grad,
                                                             int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
  if (grad != nullptr) {
    double g[3];
    dist[i] = sdf_value_grad(sdf, x, y, z, g);
    grad[3 * i] = g[0]; grad[3 * i + 1] = g[1]; grad[3 * i + 2] = g[2];
  } else {
    const Voxel v = sdf_voxel(sdf, x, y, z);
    dist[i] = sdf_at(sdf, v.ix, v.iy, v.iz);
  }
}

__device__ __forceinline__ double stable_sigmoid(double x) {
  if (x >= 0.0) return 1.0 / (1.0 + exp(-x));
  const double e = exp(x);
  return e / (1.0 + e);
}

// Fused likelihood: squash -> FK -> spheres -> SDF stencil -> hinge -> logp, plus the reverse pass to the input.
template <int D, bool BWD>
__global__ void __launch_bounds__(kThreads) loglik_kernel(RobotDev rb, SdfDev sdf, LikDev lk,
                                                         const double* __restrict__ in, int squash, double upstream,
                                                         double* __restrict__ logp, double* __restrict__ d_in,
                                                         int64_t n) {
  const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (c >= n) return;

  double th[D], dsq[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const double x = in[c * D + j];
    if (squash) {
      const double s = stable_sigmoid(x);
      const double span = rb.hi[j] - rb.lo[j];
      th[j] = rb.lo[j] + span * s;
      dsq[j] = span * s * (1.0 - s);
    } else {
      th[j] = x;
      dsq[j] = 1.0;
    }
  }

  Frame A;
  frame_from_base(rb, A);
  double Fw[3] = {0.0, 0.0, 0.0}, Tw[3] = {0.0, 0.0, 0.0};  // running wrench of the spheres seen so far
  double az[D][3], an[D][3], cj[D];                         // joint axis, axis x point, prefix term
  double lp = 0.0;
  const double inv_sigma = 1.0 / lk.sigma_obs;
  int p = 0;

#pragma unroll
  for (int k = 0; k <= D; ++k) {
    if (k > 0) {
      const int j = k - 1;
      if (BWD && !rb.craig) {  // Spong: joint j turns about z of frame j-1, through its origin
        az[j][0] = A.r[2]; az[j][1] = A.r[5]; az[j][2] = A.r[8];
        an[j][0] = az[j][1] * A.t[2] - az[j][2] * A.t[1];
        an[j][1] = az[j][2] * A.t[0] - az[j][0] * A.t[2];
        an[j][2] = az[j][0] * A.t[1] - az[j][1] * A.t[0];
      }
      frame_step(rb, j, th[j], A);
      if (BWD && rb.craig) {   // Craig: joint j turns about z of frame j (its own frame), through its origin
        az[j][0] = A.r[2]; az[j][1] = A.r[5]; az[j][2] = A.r[8];
        an[j][0] = az[j][1] * A.t[2] - az[j][2] * A.t[1];
        an[j][1] = az[j][2] * A.t[0] - az[j][0] * A.t[2];
        an[j][2] = az[j][0] * A.t[1] - az[j][1] * A.t[0];
      }
      if (BWD)
        cj[j] = az[j][0] * Tw[0] + az[j][1] * Tw[1] + az[j][2] * Tw[2] -
                (an[j][0] * Fw[0] + an[j][1] * Fw[1] + an[j][2] * Fw[2]);
    }
    while (p < rb.num_spheres && rb.sphere_frame[p] == k) {
      const double ox = rb.sphere_off[p][0], oy = rb.sphere_off[p][1], oz = rb.sphere_off[p][2];
      const double x = A.r[0] * ox + A.r[1] * oy + A.r[2] * oz + A.t[0];
      const double y = A.r[3] * ox + A.r[4] * oy + A.r[5] * oz + A.t[1];
      const double z = A.r[6] * ox + A.r[7] * oy + A.r[8] * oz + A.t[2];
      double g[3];
      double dist;
      if (BWD) {
        dist = sdf_value_grad(sdf, x - lk.offset[0], y - lk.offset[1], z - lk.offset[2], g);
      } else {
        const Voxel v = sdf_voxel(sdf, x - lk.offset[0], y - lk.offset[1], z - lk.offset[2]);
        dist = sdf_at(sdf, v.ix, v.iy, v.iz);
      }
      dist -= rb.sphere_rad[p];
      const double hinge = fmax(lk.epsilon - dist, 0.0);
      lp -= 0.5 * (hinge * inv_sigma) * hinge;
      if (BWD && hinge > 0.0) {
        // d logp / d dist = hinge / sigma; the custom gradient defines d dist / d x := stencil gradient
        const double w = hinge * inv_sigma;
        const double gx = w * g[0], gy = w * g[1], gz = w * g[2];
        Fw[0] += gx; Fw[1] += gy; Fw[2] += gz;
        Tw[0] += y * gz - z * gy;
        Tw[1] += z * gx - x * gz;
        Tw[2] += x * gy - y * gx;
      }
      ++p;
    }
  }
  logp[c] = lp;
  if (BWD) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double dth = az[j][0] * Tw[0] + az[j][1] * Tw[1] + az[j][2] * Tw[2] -
                         (an[j][0] * Fw[0] + an[j][1] * Fw[1] + an[j][2] * Fw[2]) - cj[j];
      d_in[c * D + j] = upstream * dth * dsq[j];
    }
  }
}

template <int D>
cudaError_t launch_loglik_d(vgpmp_handle* h, const double* in, int squash, double upstream, double* logp, double* d_in,
                            int64_t n, cudaStream_t s) {
  const unsigned blocks = (unsigned)((n + kThreads - 1) / kThreads);
  if (d_in != nullptr)
    loglik_kernel<D, true><<<blocks, kThreads, 0, s>>>(h->robot, h->sdf, h->lik, in, squash, upstream, logp, d_in, n);
  else
    loglik_kernel<D, false><<<blocks, kThreads, 0, s>>>(h->robot, h->sdf, h->lik, in, squash, upstream, logp, d_in, n);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_fk_frames(vgpmp_handle* h, const double* joints, double* frames, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  fk_frames_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, s>>>(h->robot, joints, frames, n);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_fk_spheres(vgpmp_handle* h, const double* joints, double* centres, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  fk_spheres_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, s>>>(h->robot, joints, centres, n);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_sdf_lookup(vgpmp_handle* h, const double* pts, double* dist, double* grad, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  sdf_lookup_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, s>>>(h->sdf, pts, dist, grad, n);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_loglik(vgpmp_handle* h, const double* in, int squash, double upstream, double* logp, double* d_in,
                          int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  h->launches++;
  switch (h->robot.dof) {
    case 1: return launch_loglik_d<1>(h, in, squash, upstream, logp, d_in, n, s);
    case 2: return launch_loglik_d<2>(h, in, squash, upstream, logp, d_in, n, s);
    case 3: return launch_loglik_d<3>(h, in, squash, upstream, logp, d_in, n, s);
    case 4: return launch_loglik_d<4>(h, in, squash, upstream, logp, d_in, n, s);
    case 5: return launch_loglik_d<5>(h, in, squash, upstream, logp, d_in, n, s);
    case 6: return launch_loglik_d<6>(h, in, squash, upstream, logp, d_in, n, s);
    case 7: return launch_loglik_d<7>(h, in, squash, upstream, logp, d_in, n, s);
    case 8: return launch_loglik_d<8>(h, in, squash, upstream, logp, d_in, n, s);
    default: return cudaErrorInvalidValue;
  }
}
