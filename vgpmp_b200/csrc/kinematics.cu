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

// clip(trunc(a / delta), 0, n-1) exactly as the reference computes it (a correctly rounded float64 division, truncation
// toward zero, clip) without paying for a division per axis per sphere.  q = a * (1/delta) is within 2 ulp of a/delta, so
// both truncate alike unless q sits next to an integer; only then (fraction within 1e-9 of 0 or 1, i.e. ~1e-9 of all
// lookups) is the true division evaluated.  Out-of-range values need no care: anything at or beyond an end clips to
// that end under either rounding.  (cvt.rzi.s32.f64 saturates, so far-away points are safe.)
__device__ __forceinline__ int voxel_index(double a, double delta, double inv_delta, int n) {
  const double q = a * inv_delta;
  int i = __double2int_rz(q);
  if ((unsigned)i < (unsigned)n) {
    const double frac = q - (double)i;
    if (!(frac > 1e-9 && frac < 1.0 - 1e-9)) i = clamp_index(a / delta, n);
    return i;
  }
  return min(max(i, 0), n - 1);
}

struct Voxel {
  int ix, iy, iz;
};

__device__ __forceinline__ Voxel sdf_voxel(const SdfDev& s, double x, double y, double z) {
  Voxel v;
  v.ix = voxel_index(x - s.origin[0], s.delta, s.inv_delta, s.nx);
  v.iy = voxel_index(y - s.origin[1], s.delta, s.inv_delta, s.ny);
  v.iz = voxel_index(z - s.origin[2], s.delta, s.inv_delta, s.nz);
  return v;
}

// The same index for a coordinate that is ALREADY relative to the grid origin, without the double<->int conversions
// (quarter-rate on sm_100): q + 1.5 * 2^52 leaves rint(q) in the low mantissa word, r = q - rint(q) gives floor and the
// distance to the nearest integer in full-rate FP64 adds.  Negative q clips to 0 under trunc and under floor alike.  Returns
// false when q sits within 1e-9 of an integer (or is absurdly large / NaN): the caller then takes the exact path above.
// (5 FP64 instructions + 2 integer min/max per axis; the conversion-based form was 14.)
__device__ __forceinline__ bool voxel_index_fast(double a, double inv_delta, int n, int& idx) {
  const double qh = fma(a, inv_delta, -0.5);           // q - 1/2: rint(q - 1/2) = floor(q) unless q is an integer
  const double t = qh + 6755399441055744.0;            // 1.5 * 2^52: the low mantissa word of t is rint(qh)
  const double r = qh - (t - 6755399441055744.0);      // in [-1/2, 1/2]; |r| -> 1/2 when q approaches an integer
  idx = min(max(__double2loint(t), 0), n - 1);
  return fabs(r) < 0.5 - 1e-9 && fabs(qh) < 1.0e9;
}

// grid-relative position -> voxel (fast path for all three axes, exact path if any of them asks for it)
__device__ __forceinline__ Voxel sdf_voxel_rel(const SdfDev& s, double x, double y, double z) {
  Voxel v;
  const bool ok = voxel_index_fast(x, s.inv_delta, s.nx, v.ix) & voxel_index_fast(y, s.inv_delta, s.ny, v.iy) &
                  voxel_index_fast(z, s.inv_delta, s.nz, v.iz);
  if (!ok) {
    v.ix = voxel_index(x, s.delta, s.inv_delta, s.nx);
    v.iy = voxel_index(y, s.delta, s.inv_delta, s.ny);
    v.iz = voxel_index(z, s.delta, s.inv_delta, s.nz);
  }
  return v;
}

// one branch-free double sincos for |x| < 2^20 (joint angles): three-term Cody-Waite reduction by pi/2, Taylor kernels on
// [-pi/4, pi/4] truncated below 1e-17 (the arithmetic of sincos_bf6 in device_utils.cuh); the library sincos carries a
// Payne-Hanek slow path and ~2x the instructions
__device__ __forceinline__ void sincos_small(double x, double& sn, double& cs) {
  const double nq = rint(x * 0.63661977236758134308);
  const int q = __double2int_rn(nq);
  const double r = fma(-nq, 6.123233995736766e-17, fma(-nq, 1.5707963267948966, x));
  const double z = r * r;
  double ps = 1.0 / 1307674368000.0, pc = 1.0 / 20922789888000.0;
  ps = fma(ps, -z, 1.0 / 6227020800.0);  pc = fma(pc, -z, 1.0 / 87178291200.0);
  ps = fma(ps, -z, 1.0 / 39916800.0);    pc = fma(pc, -z, 1.0 / 479001600.0);
  ps = fma(ps, -z, 1.0 / 362880.0);      pc = fma(pc, -z, 1.0 / 3628800.0);
  ps = fma(ps, -z, 1.0 / 5040.0);        pc = fma(pc, -z, 1.0 / 40320.0);
  ps = fma(ps, -z, 1.0 / 120.0);         pc = fma(pc, -z, 1.0 / 720.0);
  ps = fma(ps, -z, 1.0 / 6.0);           pc = fma(pc, -z, 1.0 / 24.0);
  pc = fma(pc, -z, 0.5);
  const double s = fma(-z * r, ps, r), c = fma(-z, pc, 1.0);
  const double a = (q & 1) ? c : s, b = (q & 1) ? s : c;
  sn = (q & 2) ? -a : a;
  cs = ((q + 1) & 2) ? -b : b;
}

__device__ __forceinline__ size_t sdf_cell(const SdfDev& s, int ix, int iy, int iz) {
  return (size_t)(((unsigned)ix * (unsigned)s.ny + (unsigned)iy) * (unsigned)s.nz + (unsigned)iz);   // vgpmp_create: cells < 2^32
}

__device__ __forceinline__ double sdf_value(const SdfDev& s, const Voxel& v) {
  return __ldg(reinterpret_cast<const double*>(s.rec + sdf_cell(s, v.ix, v.iy, v.iz)));
}

// one sphere-SDF evaluation: value + stencil gradient in a single 256-bit read-only load (LDG.E.256, sm_100+)
__device__ __forceinline__ double4 sdf_record(const SdfDev& s, const Voxel& v) {
  double4 r;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
      : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
      : "l"(s.rec + sdf_cell(s, v.ix, v.iy, v.iz)));
  return r;
}

// Builds the records from the raw grid: sdf_utils.py:100-136 verbatim per voxel (clipped neighbours, /(2 delta), 0 -> 0.1)
__global__ void __launch_bounds__(256) sdf_build_kernel(const double* __restrict__ raw, double4* __restrict__ rec,
                                                       int nx, int ny, int nz, double delta) {
  const size_t cells = (size_t)nx * ny * nz;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cells) return;
  const int iz = (int)(i % nz), iy = (int)((i / nz) % ny), ix = (int)(i / ((size_t)nz * ny));
  auto at = [&](int x, int y, int z) { return raw[((size_t)x * ny + y) * nz + z]; };
  const double den = 2.0 * delta;
  double gx = (at(min(ix + 1, nx - 1), iy, iz) - at(max(ix - 1, 0), iy, iz)) / den;
  double gy = (at(ix, min(iy + 1, ny - 1), iz) - at(ix, max(iy - 1, 0), iz)) / den;
  double gz = (at(ix, iy, min(iz + 1, nz - 1)) - at(ix, iy, max(iz - 1, 0))) / den;
  gx = (gx == 0.0) ? 0.1 : gx;  // sdf_utils.py:124,129,135
  gy = (gy == 0.0) ? 0.1 : gy;
  gz = (gz == 0.0) ? 0.1 : gz;
  rec[i] = make_double4(raw[i], gx, gy, gz);
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
  const Voxel v = sdf_voxel(sdf, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
  if (grad != nullptr) {
    const double4 r = sdf_record(sdf, v);
    dist[i] = r.x;
    grad[3 * i] = r.y; grad[3 * i + 1] = r.z; grad[3 * i + 2] = r.w;
  } else {
    dist[i] = sdf_value(sdf, v);
  }
}

// min over spheres of (sdf(x_p - offset) - r_p) for one joint configuration: the collision-free verdict of SURVEY.md 8f-3
__global__ void __launch_bounds__(kThreads) clearance_kernel(RobotDev rb, SdfDev sdf, LikDev lk,
                                                            const double* __restrict__ joints,
                                                            double* __restrict__ clearance, int64_t n) {
  const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (c >= n) return;
  Frame A;
  frame_from_base(rb, A);
  double best = 1e300;
  int p = 0;
  for (int j = 0; j <= rb.dof; ++j) {
    if (j > 0) frame_step(rb, j - 1, joints[c * rb.dof + j - 1], A);
    for (; p < rb.frame_end[j]; ++p) {
      const double ox = rb.sphere_off[p][0], oy = rb.sphere_off[p][1], oz = rb.sphere_off[p][2];
      const double x = A.r[0] * ox + A.r[1] * oy + A.r[2] * oz + A.t[0];
      const double y = A.r[3] * ox + A.r[4] * oy + A.r[5] * oz + A.t[1];
      const double z = A.r[6] * ox + A.r[7] * oy + A.r[8] * oz + A.t[2];
      const Voxel v = sdf_voxel(sdf, x - lk.offset[0], y - lk.offset[1], z - lk.offset[2]);
      best = fmin(best, sdf_value(sdf, v) - rb.sphere_rad[p]);
    }
  }
  clearance[c] = best;
}

__device__ __forceinline__ double stable_sigmoid(double x) {
  // one path for both signs: e = exp(-|x|) in (0, 1], sigmoid = (x >= 0 ? 1 : e) / (1 + e); the reciprocal of 1 + e in
  // [1, 2] is a MUFU seed + two Newton steps (no IEEE division sequence, no divergent branch)
  const double e = exp(-fabs(x));
  const double d = 1.0 + e;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(r, fma(-d, r, 1.0), r);
  r = fma(r, fma(-d, r, 1.0), r);
  r = fma(r, fma(-d, r, 1.0), r);
  return (x >= 0.0 ? 1.0 : e) * r;
}

// Where the D joint inputs of configuration c live: [n,D] (planar_sn = 0), or latent-major per problem,
// [n / planar_sn][D][planar_sn] (the fused step's layout: a warp's lanes then read consecutive doubles).
struct JointLayout {
  size_t base, stride;
};
__device__ __forceinline__ JointLayout joint_layout(int64_t c, int D, int64_t planar_sn) {
  JointLayout jl;
  if (planar_sn == 0) {
    jl.base = (size_t)c * D; jl.stride = 1;
  } else {
    const int64_t p = (c >> 32) == 0 && (planar_sn >> 32) == 0 ? (int64_t)((unsigned)c / (unsigned)planar_sn) : c / planar_sn;
    jl.base = (size_t)p * planar_sn * D + (size_t)(c - p * planar_sn); jl.stride = (size_t)planar_sn;
  }
  return jl;
}

template <int NB>
struct SphereBatch {
  double x[NB], y[NB], z[NB];
  double4 r[NB];  // {value, gx, gy, gz}
};

// Forward-only likelihood (prediction, best-sample selection, clearance-style queries): squash -> FK -> spheres -> SDF value
// -> hinge -> logp.  Spheres of a frame are handled NB at a time: all NB value loads are issued before the first is consumed.
template <int D, int NB>
__global__ void __launch_bounds__(kThreads, 3) loglik_kernel(RobotDev rb, SdfDev sdf, LikDev lk,
                                                         const double* __restrict__ in, int squash,
                                                         double* __restrict__ logp, int64_t n, int64_t planar_sn) {
  const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (c >= n) return;
  const JointLayout jl = joint_layout(c, D, planar_sn);   // joint k of this configuration: in[jl.base + k * jl.stride]

  Frame A;
  frame_from_base(rb, A);
  double lp = 0.0;
  const double inv_sigma = 1.0 / lk.sigma_obs;
  double xnext = in[jl.base];   // joint inputs are fetched one joint ahead of their use

#pragma unroll 1
  for (int k = 0; k <= D; ++k) {
    if (k > 0) {
      const int j = k - 1;
      const double xin = xnext;
      if (k < D) xnext = in[jl.base + k * jl.stride];
      lp += xin - xin;     // 0 for a finite input; NaN / Inf inputs must not vanish in the voxel clip and the hinge's fmax
      double thj = xin;
      if (squash) thj = rb.lo[j] + (rb.hi[j] - rb.lo[j]) * stable_sigmoid(xin);
      frame_step(rb, j, thj, A);
    }
    const int pend = rb.frame_end[k];
    for (int p = (k == 0 ? 0 : rb.frame_end[k - 1]); p < pend; p += NB) {
      double val[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int q = min(p + i, pend - 1);  // tail lanes of the batch repeat the last sphere (not counted below)
        const double ox = rb.sphere_off[q][0], oy = rb.sphere_off[q][1], oz = rb.sphere_off[q][2];
        const double x = A.r[0] * ox + A.r[1] * oy + A.r[2] * oz + A.t[0];
        const double y = A.r[3] * ox + A.r[4] * oy + A.r[5] * oz + A.t[1];
        const double z = A.r[6] * ox + A.r[7] * oy + A.r[8] * oz + A.t[2];
        val[i] = sdf_value(sdf, sdf_voxel(sdf, x - lk.offset[0], y - lk.offset[1], z - lk.offset[2]));
      }
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        if (p + i < pend) {
          const double hinge = fmax(lk.epsilon - (val[i] - rb.sphere_rad[p + i]), 0.0);
          lp -= 0.5 * (hinge * inv_sigma) * hinge;
        }
      }
    }
  }
  logp[c] = lp;
}

// Fused likelihood WITH its reverse pass, two sweeps over the kinematic chain.  (A one-sweep form has to park 8 doubles per
// joint and thread in shared memory for the final joint-gradient formula: 56 KB per CTA, i.e. 3 CTAs per SM and, with the
// carve-out that takes, little L1 left for the record loads - which matters once the grid lives in HBM.)
//   sweep 1  joints + spheres in chain order: log-probability, total wrench (F, T), and per joint the prefix term
//            p_j = z_j . T_{<j} - (z_j x o_j) . F_{<j} in REGISTERS (the joint loop is unrolled); sin / cos of the joint angle
//            and the squash derivative go to shared memory (3 doubles per joint: 21 KB per CTA at D = 7);
//   sweep 2  the chain again from the stored sin / cos (frame products only, no transcendental, no sphere):
//            d theta_j = z_j . T - (z_j x o_j) . F - p_j.
// Sphere positions come from exactly the same operations as before, so voxel indices (hence parity) are unchanged.
__device__ __forceinline__ void frame_step_sc(const RobotDev& rb, int j, double st, double ct, Frame& A) {
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

template <int D, int NB, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) loglik_bwd_kernel(RobotDev rb, SdfDev sdf, LikDev lk,
                                                                   const double* __restrict__ in, int squash, double upstream,
                                                                   double* __restrict__ logp, double* __restrict__ d_in,
                                                                   int64_t n, int64_t planar_sn) {
  extern __shared__ double trig[];  // [D][4][kThreads]  sin, cos of the joint angle, d theta / d in, prefix term
  const int tid = threadIdx.x;
  const int64_t c = (int64_t)blockIdx.x * kThreads + tid;
  if (c >= n) return;
  const JointLayout jl = joint_layout(c, D, planar_sn);   // joint k of this configuration: in[jl.base + k * jl.stride]

  // The chain is walked in GRID coordinates (base translation shifted by -(scene offset + grid origin)): a sphere centre
  // is then directly the voxel coordinate numerator, and the joint-gradient formula is invariant under a common shift of
  // sphere centres and axis points.
  const double shx = lk.offset[0] + sdf.origin[0], shy = lk.offset[1] + sdf.origin[1], shz = lk.offset[2] + sdf.origin[2];
  Frame A;
  frame_from_base(rb, A);
  A.t[0] -= shx; A.t[1] -= shy; A.t[2] -= shz;
  double Fw[3] = {0.0, 0.0, 0.0}, Tw[3] = {0.0, 0.0, 0.0};  // running wrench of the spheres seen so far
  double lp = 0.0;
  const double inv_sigma = 1.0 / lk.sigma_obs;
  double xnext = in[jl.base];   // joint inputs are fetched one joint ahead of their use

#pragma unroll 1
  for (int k = 0; k <= D; ++k) {
    if (k > 0) {
      const int j = k - 1;
      const double xin = xnext;
      if (k < D) xnext = in[jl.base + k * jl.stride];
      lp += xin - xin;     // 0 for a finite input; NaN / Inf inputs must not vanish in the voxel clip and the hinge's fmax
      double thj = xin, dsq = 1.0;
      if (squash) {
        const double sg = stable_sigmoid(xin), span = rb.hi[j] - rb.lo[j];
        thj = rb.lo[j] + span * sg;
        dsq = span * sg * (1.0 - sg);
      }
      double st, ct;
      const double ang = thj + rb.twist[j];
      if (fabs(ang) < 1048576.0) sincos_small(ang, st, ct);
      else sincos(ang, &st, &ct);
      double* slot = trig + (size_t)j * 4 * kThreads + tid;
      slot[0] = st; slot[kThreads] = ct; slot[2 * kThreads] = dsq;
      double zx, zy, zz, ox, oy, oz;
      if (!rb.craig) {  // Spong: joint j turns about z of frame j-1, through its origin
        zx = A.r[2]; zy = A.r[5]; zz = A.r[8]; ox = A.t[0]; oy = A.t[1]; oz = A.t[2];
      }
      frame_step_sc(rb, j, st, ct, A);
      if (rb.craig) {   // Craig: joint j turns about z of frame j (its own frame), through its origin
        zx = A.r[2]; zy = A.r[5]; zz = A.r[8]; ox = A.t[0]; oy = A.t[1]; oz = A.t[2];
      }
      const double nx = zy * oz - zz * oy, ny = zz * ox - zx * oz, nz = zx * oy - zy * ox;
      slot[3 * kThreads] = zx * Tw[0] + zy * Tw[1] + zz * Tw[2] - (nx * Fw[0] + ny * Fw[1] + nz * Fw[2]);
    }
    const int pend = rb.frame_end[k];
    for (int p = (k == 0 ? 0 : rb.frame_end[k - 1]); p < pend; p += NB) {
      SphereBatch<NB> sb;
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int q = min(p + i, pend - 1);  // tail lanes of the batch repeat the last sphere (weight 0 below)
        const double ox = rb.sphere_off[q][0], oy = rb.sphere_off[q][1], oz = rb.sphere_off[q][2];
        sb.x[i] = A.r[0] * ox + A.r[1] * oy + A.r[2] * oz + A.t[0];
        sb.y[i] = A.r[3] * ox + A.r[4] * oy + A.r[5] * oz + A.t[1];
        sb.z[i] = A.r[6] * ox + A.r[7] * oy + A.r[8] * oz + A.t[2];
        const Voxel v = sdf_voxel_rel(sdf, sb.x[i], sb.y[i], sb.z[i]);
        sb.r[i] = sdf_record(sdf, v);
      }
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        if (p + i < pend) {
          const double dist = sb.r[i].x - rb.sphere_rad[p + i];
          const double hinge = fmax(lk.epsilon - dist, 0.0);
          const double w = hinge * inv_sigma;          // 0 outside the hinge: the wrench update needs no per-lane branch
          lp -= 0.5 * w * hinge;
          if (__any_sync(__activemask(), hinge > 0.0)) {   // lanes = neighbouring timesteps: usually all in or all out
            const double gx = sb.r[i].y * w, gy = sb.r[i].z * w, gz = sb.r[i].w * w;
            Fw[0] += gx; Fw[1] += gy; Fw[2] += gz;
            Tw[0] += sb.y[i] * gz - sb.z[i] * gy;
            Tw[1] += sb.z[i] * gx - sb.x[i] * gz;
            Tw[2] += sb.x[i] * gy - sb.y[i] * gx;
          }
        }
      }
    }
  }
  logp[c] = lp;
  // no sphere of this warp's configurations inside the hinge: every joint gradient is zero, sweep 2 has nothing to do
  if (!__any_sync(__activemask(), Fw[0] != 0.0 || Fw[1] != 0.0 || Fw[2] != 0.0 || Tw[0] != 0.0 || Tw[1] != 0.0 || Tw[2] != 0.0)) {
#pragma unroll 1
    for (int j = 0; j < D; ++j) d_in[jl.base + j * jl.stride] = 0.0;
    return;
  }
  // sweep 2: joint axes again (frame products from the stored sin / cos), now against the TOTAL wrench
  frame_from_base(rb, A);
  A.t[0] -= shx; A.t[1] -= shy; A.t[2] -= shz;
#pragma unroll 1
  for (int j = 0; j < D; ++j) {
    const double* slot = trig + (size_t)j * 4 * kThreads + tid;
    const double st = slot[0], ct = slot[kThreads], dsq = slot[2 * kThreads], pj = slot[3 * kThreads];
    double zx, zy, zz, ox, oy, oz;
    if (!rb.craig) { zx = A.r[2]; zy = A.r[5]; zz = A.r[8]; ox = A.t[0]; oy = A.t[1]; oz = A.t[2]; }
    frame_step_sc(rb, j, st, ct, A);
    if (rb.craig) { zx = A.r[2]; zy = A.r[5]; zz = A.r[8]; ox = A.t[0]; oy = A.t[1]; oz = A.t[2]; }
    const double nx = zy * oz - zz * oy, ny = zz * ox - zx * oz, nz = zx * oy - zy * ox;
    const double dth = zx * Tw[0] + zy * Tw[1] + zz * Tw[2] - (nx * Fw[0] + ny * Fw[1] + nz * Fw[2]) - pj;
    d_in[jl.base + j * jl.stride] = upstream * dth * dsq;
  }
}

constexpr int kSphereBatch = 4;
#ifndef VGPMP_BWD_BATCH
#define VGPMP_BWD_BATCH 3
#endif
#ifndef VGPMP_BWD_MINB
#define VGPMP_BWD_MINB 4
#endif
constexpr int kBwdBatch = VGPMP_BWD_BATCH, kBwdMinBlocks = VGPMP_BWD_MINB;

template <int D>
cudaError_t launch_loglik_d(vgpmp_handle* h, const double* in, int squash, double upstream, double* logp, double* d_in,
                            int64_t n, int64_t planar_sn, cudaStream_t s) {
  const unsigned blocks = (unsigned)((n + kThreads - 1) / kThreads);
  if (d_in != nullptr) {
    const size_t smem = sizeof(double) * D * 4 * kThreads;
    auto kern = loglik_bwd_kernel<D, kBwdBatch, kBwdMinBlocks>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<blocks, kThreads, smem, s>>>(h->robot, h->sdf, h->lik, in, squash, upstream, logp, d_in, n, planar_sn);
  } else {
    loglik_kernel<D, kSphereBatch><<<blocks, kThreads, 0, s>>>(h->robot, h->sdf, h->lik, in, squash, logp, n, planar_sn);
  }
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

cudaError_t launch_sdf_build(vgpmp_handle* h, const double* raw_dev, cudaStream_t s) {
  const size_t cells = (size_t)h->sdf.nx * h->sdf.ny * h->sdf.nz;
  sdf_build_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(raw_dev, h->rec_dev, h->sdf.nx, h->sdf.ny, h->sdf.nz,
                                                                  h->sdf.delta);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_sdf_lookup(vgpmp_handle* h, const double* pts, double* dist, double* grad, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  sdf_lookup_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, s>>>(h->sdf, pts, dist, grad, n);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_clearance(vgpmp_handle* h, const double* joints, double* clearance, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  clearance_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, s>>>(h->robot, h->sdf, h->lik, joints,
                                                                                   clearance, n);
  h->launches++;
  return cudaGetLastError();
}

cudaError_t launch_loglik(vgpmp_handle* h, const double* in, int squash, double upstream, double* logp, double* d_in,
                          int64_t n, int64_t planar_sn, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  h->launches++;
  switch (h->robot.dof) {
    case 1: return launch_loglik_d<1>(h, in, squash, upstream, logp, d_in, n, planar_sn, s);
    case 2: return launch_loglik_d<2>(h, in, squash, upstream, logp, d_in, n, planar_sn, s);
    case 3: return launch_loglik_d<3>(h, in, squash, upstream, logp, d_in, n, planar_sn, s);
    case 4: return launch_loglik_d<4>(h, in, squash, upstream, logp, d_in, n, planar_sn, s);
    case 5: return launch_loglik_d<5>(h, in, squash, upstream, logp, d_in, n, planar_sn, s);
    case 6: return launch_loglik_d<6>(h, in, squash, upstream, logp, d_in, n, planar_sn, s);
    case 7: return launch_loglik_d<7>(h, in, squash, upstream, logp, d_in, n, planar_sn, s);
    case 8: return launch_loglik_d<8>(h, in, squash, upstream, logp, d_in, n, planar_sn, s);
    default: return cudaErrorInvalidValue;
  }
}
