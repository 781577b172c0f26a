// Device-side forward models and costs.  Every translation unit that includes this file is
// compiled with -fmad=false: the expressions below follow the reference's float32 op order
// one rounding at a time, so the particle trajectories are bit-identical
// to the CPU reference.  The pendulum cannot be (its sinf/cosf are not the reference's libm): its
// step and cost use explicit fmaf, staying within a few 1e-7 relative of the reference.
#pragma once
#include "common.cuh"

// DUST_PEND_OPT: which of the pendulum instruction-count reductions are compiled in (bit mask;
// the variants exist so that their effect on parity can be measured one by one)
//   1 two-part reduction + mantissa-trick rounding   2 Horner cosine   4 contracted dynamics
//   8 unweighted running cost sums                   16 sigma*eps score rows (rollout.cu)
// Measured on B200 with profiles/variant_check.py: every mask leaves the cost
// error at 2.5e-7 relative; 8 is left OUT because it abandons the reference's summation order
// (step cost rounded, then added), and with it the soft-min weights of the hardest fixture moved
// from 7e-5 to 2.3e-4 of the reference (weights are exp(-cost differences), one cost ulp = 5e-4).
#ifndef DUST_PEND_OPT
#define DUST_PEND_OPT 23
#endif

namespace dust {

constexpr float kPiF = 3.14159274101257324f;  // (float)math.pi  (pendulum.py:95)

// ---------------------------------------------------------------------------------------
// inverted pendulum  (dust/models/pendulum.py:84-100)
// ---------------------------------------------------------------------------------------
struct PendulumCoef {
  float c1, c2;  // c1 = (-3 g)/(2 l),  c2 = 3/(m l^2)
};

// With sampled (tensor) parameters torch evaluates `scalar / tensor` as reciprocal(tensor) * scalar
// (Tensor.__rtruediv__); with the python-float defaults the quotient is formed in double first.
__device__ __forceinline__ PendulumCoef pendulum_coef_sampled(const ModelParams& m, float length, float mass) {
  PendulumCoef c;
  const float neg3g = (float)(-3.0 * (double)m.g);
  c.c1 = __fmul_rn(__frcp_rn(__fmul_rn(2.0f, length)), neg3g);
  c.c2 = __fmul_rn(__frcp_rn(__fmul_rn(mass, __fmul_rn(length, length))), 3.0f);
  return c;
}
__device__ __forceinline__ PendulumCoef pendulum_coef_default(const ModelParams& m) {
  PendulumCoef c;
  c.c1 = m.pend_c1;   // (float)(-3 g / (2 l)) and (float)(3 / (m l^2)) in double: to_params() (common.cuh)
  c.c2 = m.pend_c2;
  return c;
}

// sin and cos of one argument with a single Cody-Waite reduction (pi/2 in two parts, |x| <= 64)
// and the Cephes single-precision minimax polynomials on [-pi/4, pi/4]; max abs error
// 9.2e-8 (~1.5 ulp), the same class as sinf/cosf.  Larger arguments take the library path.
// (the library path is kept out of line: inlined, its Payne-Hanek code would bloat the rollout
// loop past the instruction cache)
static __device__ __noinline__ void slow_sincosf(float x, float* s, float* c) {
  *s = sinf(x);
  *c = cosf(x);
}
template <bool CHECK = true>
__device__ __forceinline__ void fast_sincosf(float x, float& s, float& c) {
  if (CHECK && __builtin_expect(fabsf(x) > 64.0f, 0)) {
    slow_sincosf(x, &s, &c);
    return;
  }
#if DUST_PEND_OPT & 1
  // k = rint(x * 2/pi) by the 1.5*2^23 trick: the integer sits in the low mantissa bits of t, no
  // FRND / F2I.  |k| <= 41, so the reduction needs only two parts of pi/2: x - k*hi is exact (both
  // are multiples of 2^-24 and the difference is below 1), the low part is carried to 2^-53.
  const float t = fmaf(x, 0.636619772f, 12582912.0f);
  const float kf = t - 12582912.0f;
  const uint32_t q = __float_as_uint(t);
  float r = fmaf(kf, -1.57079637050628662109375f, x);
  r = fmaf(kf, 4.37113900018624283e-8f, r);
#else
  const float kf = rintf(x * 0.636619772f);
  const uint32_t q = (uint32_t)(int)kf;
  float r = fmaf(kf, -1.5703125f, x);
  r = fmaf(kf, -4.837512969970703125e-4f, r);
  r = fmaf(kf, -7.54978995489188216e-8f, r);
#endif
  const float r2 = r * r;
  float ps = fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f);
  ps = fmaf(ps, r2 * r, r);
  float pc = fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f);
#if DUST_PEND_OPT & 2
  pc = fmaf(fmaf(pc, r2, -0.5f), r2, 1.0f);
#else
  pc = fmaf(pc, r2 * r2, fmaf(-0.5f, r2, 1.0f));
#endif
  const bool odd = (q & 1u) != 0u;
  const float sv = odd ? pc : ps;
  const float cv = odd ? ps : pc;
  // sign bits: sin is negative in quadrants 2,3 (bit 1 of q), cos in quadrants 1,2 (bit 1 of q+1)
  s = __uint_as_float(__float_as_uint(sv) ^ ((q << 30) & 0x80000000u));
  c = __uint_as_float(__float_as_uint(cv) ^ (((q << 30) + 0x40000000u) & 0x80000000u));
}

// one step; returns the pre-clamp angular speed through *pre (the adjoint needs the mask).
// If cos_th is given it receives cos(th) of the state BEFORE the step, obtained from the same
// reduction as sin(th + pi): with y = fl(th + pi_f) and the exact rounding error err of that sum
// (2Sum), th = y - pi - e, e = (pi_f - pi) - err, so cos(th) = -cos(y - e) = -(cos y + e sin y) + O(e^2).
// CHECK = false: the caller guarantees |th + pi| <= 64 (no library fall-back in the loop).
template <bool CHECK = true>
__device__ __forceinline__ void pendulum_step(const ModelParams& m, const PendulumCoef& c, float& th, float& om,
                                              float a, float* pre_out = nullptr, float* cos_th = nullptr) {
  const float u = fminf(fmaxf(a, -m.max_torque), m.max_torque);
  const float y = th + kPiF;
  float s, cy;
  fast_sincosf<CHECK>(y, s, cy);
  if (cos_th) {
    const float bb = y - th;
    const float err = (th - (y - bb)) + (kPiF - bb);
    const float ne = err - 8.742278e-8f;  // -e
    *cos_th = fmaf(ne, s, -cy);
  }
  // the trig above is already a few 1e-8 from the reference's libm: the products below are
  // contracted (one rounding instead of two), a deviation of the same order and a shorter loop
#if DUST_PEND_OPT & 4
  const float acc = fmaf(c.c1, s, c.c2 * u);
  const float pre = fmaf(m.dt, acc, om);
#else
  const float acc = c.c1 * s + c.c2 * u;
  const float pre = om + m.dt * acc;
#endif
  if (pre_out) *pre_out = pre;
  om = fminf(fmaxf(pre, -m.max_speed_pend), m.max_speed_pend);
#if DUST_PEND_OPT & 4
  th = fmaf(om, m.dt, th);
#else
  th = th + om * m.dt;
#endif
}

// demo/pendulum_example.py:21-28
__device__ __forceinline__ float pendulum_cost_from_cos(const ModelParams& m, float cos_th, float om) {
  float t = cos_th - 1.0f;
  t = t * t;
  return m.w_angle * t + m.w_speed * (om * om);
}
// running cost of a trajectory: the two quadratic terms are summed unweighted (one fused
// multiply-add each per step) and the weights applied once, sum_t [w_a (cos th_t - 1)^2 + w_s om_t^2]
#if DUST_PEND_OPT & 8
struct PendulumCostSum {
  float ang = 0.f, spd = 0.f;
  __device__ __forceinline__ void add(const ModelParams&, float cos_th, float om) {
    const float t = cos_th - 1.0f;
    ang = fmaf(t, t, ang);
    spd = fmaf(om, om, spd);
  }
  __device__ __forceinline__ float total(const ModelParams& m) const { return fmaf(m.w_angle, ang, m.w_speed * spd); }
};
#else
// the reference's own order: every step cost rounded, then added to the running sum
struct PendulumCostSum {
  float sum = 0.f;
  __device__ __forceinline__ void add(const ModelParams& m, float cos_th, float om) { sum = sum + pendulum_cost_from_cos(m, cos_th, om); }
  __device__ __forceinline__ float total(const ModelParams&) const { return sum; }
};
#endif
template <bool CHECK = true>
__device__ __forceinline__ float pendulum_cost(const ModelParams& m, float th, float om) {
  float s, c;
  fast_sincosf<CHECK>(th, s, c);
  return pendulum_cost_from_cos(m, c, om);
}

// ---------------------------------------------------------------------------------------
// two pendulum trajectories per thread on the packed FP32 instructions of sm_100 (FFMA2 / FADD2 /
// FMUL2: one issue slot, both lanes).  Operation for operation the scalar step above (mask 23),
// each lane rounded exactly as the scalar instruction would: results are bit-identical to it.
// Caller guarantees |th + pi| <= 64 (the SMALL horizon).
// ---------------------------------------------------------------------------------------
#if DUST_PEND_OPT == 23
#define DUST_PEND_PAIR 1
__device__ __forceinline__ float2 bc2(float c) { return make_float2(c, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
// a product that must NOT be contracted with a following add: ptxas (12.9) fuses mul.rn.f32x2 +
// add.rn.f32x2 into FFMA2 even under --fmad=false, which it never does for the scalar .rn forms
__device__ __forceinline__ float2 mul2_unfused(float2 a, float2 b) { return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
__device__ __forceinline__ float flip_sign(float v, uint32_t bits) {
  return __uint_as_float(__float_as_uint(v) ^ (bits & 0x80000000u));
}
struct PendulumCoef2 {
  float2 c1, c2;
};
// one step of both trajectories; `sum` accumulates the cost of the states BEFORE the step
__device__ __forceinline__ void pendulum_step_pair(const ModelParams& m, const PendulumCoef2& c, float2& th, float2& om,
                                                   float2 a, float2& sum) {
  const float2 u = make_float2(fminf(fmaxf(a.x, -m.max_torque), m.max_torque), fminf(fmaxf(a.y, -m.max_torque), m.max_torque));
  const float2 y = add2(th, bc2(kPiF));
  const float2 t = fma2(y, bc2(0.636619772f), bc2(12582912.0f));
  const float2 kf = add2(t, bc2(-12582912.0f));
  float2 r = fma2(kf, bc2(-1.57079637050628662109375f), y);
  r = fma2(kf, bc2(4.37113900018624283e-8f), r);
  const float2 r2 = mul2(r, r);
  float2 ps = fma2(fma2(bc2(-1.9515295891e-4f), r2, bc2(8.3321608736e-3f)), r2, bc2(-1.6666654611e-1f));
  ps = fma2(ps, mul2(r2, r), r);
  float2 pc = fma2(fma2(bc2(2.443315711809948e-5f), r2, bc2(-1.388731625493765e-3f)), r2, bc2(4.166664568298827e-2f));
  pc = fma2(fma2(pc, r2, bc2(-0.5f)), r2, bc2(1.0f));
  const uint32_t qa = __float_as_uint(t.x), qb = __float_as_uint(t.y);
  const bool oa = (qa & 1u) != 0u, ob = (qb & 1u) != 0u;
  float2 s, ncy;  // sin(y) and -cos(y)
  s.x = flip_sign(oa ? pc.x : ps.x, qa << 30);
  s.y = flip_sign(ob ? pc.y : ps.y, qb << 30);
  ncy.x = flip_sign(oa ? ps.x : pc.x, (qa << 30) + 0xC0000000u);
  ncy.y = flip_sign(ob ? ps.y : pc.y, (qb << 30) + 0xC0000000u);
  const float2 bb = sub2(y, th);
  const float2 err = add2(sub2(th, sub2(y, bb)), sub2(bc2(kPiF), bb));
  const float2 ne = add2(err, bc2(-8.742278e-8f));
  const float2 cth = fma2(ne, s, ncy);
  float2 tt = add2(cth, bc2(-1.0f));
  tt = mul2(tt, tt);
  sum = add2(sum, add2(mul2_unfused(bc2(m.w_angle), tt), mul2_unfused(bc2(m.w_speed), mul2(om, om))));
  const float2 acc = fma2(c.c1, s, mul2(c.c2, u));
  const float2 pre = fma2(bc2(m.dt), acc, om);
  om = make_float2(fminf(fmaxf(pre.x, -m.max_speed_pend), m.max_speed_pend),
                   fminf(fmaxf(pre.y, -m.max_speed_pend), m.max_speed_pend));
  th = fma2(om, bc2(m.dt), th);
}
// terminal cost of both trajectories: pendulum_cost<false> lane by lane (the reduction of th itself, the same two
// polynomials, cos selected by the quadrant, weights applied as unfused products): bit-identical to the scalar call
__device__ __forceinline__ float2 pendulum_cost_pair(const ModelParams& m, float2 th, float2 om) {
  const float2 t = fma2(th, bc2(0.636619772f), bc2(12582912.0f));
  const float2 kf = add2(t, bc2(-12582912.0f));
  float2 r = fma2(kf, bc2(-1.57079637050628662109375f), th);
  r = fma2(kf, bc2(4.37113900018624283e-8f), r);
  const float2 r2 = mul2(r, r);
  float2 ps = fma2(fma2(bc2(-1.9515295891e-4f), r2, bc2(8.3321608736e-3f)), r2, bc2(-1.6666654611e-1f));
  ps = fma2(ps, mul2(r2, r), r);
  float2 pc = fma2(fma2(bc2(2.443315711809948e-5f), r2, bc2(-1.388731625493765e-3f)), r2, bc2(4.166664568298827e-2f));
  pc = fma2(fma2(pc, r2, bc2(-0.5f)), r2, bc2(1.0f));
  const uint32_t qa = __float_as_uint(t.x), qb = __float_as_uint(t.y);
  float2 c;   // cos: the sine polynomial in odd quadrants, negative in quadrants 1 and 2
  c.x = flip_sign((qa & 1u) ? ps.x : pc.x, (qa << 30) + 0x40000000u);
  c.y = flip_sign((qb & 1u) ? ps.y : pc.y, (qb << 30) + 0x40000000u);
  float2 tt = add2(c, bc2(-1.0f));
  tt = mul2(tt, tt);
  return add2(mul2_unfused(bc2(m.w_angle), tt), mul2_unfused(bc2(m.w_speed), mul2(om, om)));
}
#else
#define DUST_PEND_PAIR 0
#endif

// ---------------------------------------------------------------------------------------
// 2-D point mass among obstacles  (dust/models/particle.py:136-225, dust/utils/obstacle_map.py:64-93)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float grid_lookup(const ModelParams& m, const uint32_t* __restrict__ bits, float x, float y) {
  float fx = floorf(x * m.inv_cell + m.c_offset[0]);
  float fy = floorf(y * m.inv_cell + m.c_offset[1]);
  fx = fminf(fmaxf(fx, 0.0f), m.grid_xmax);
  fy = fminf(fmaxf(fy, 0.0f), m.grid_ymax);
  const int cell = (int)fx * m.grid_ny + (int)fy;
  return (float)((bits[cell >> 5] >> (cell & 31)) & 1u);
}

struct ParticleState {
  float x, y, vx, vy;
};

// c: occupancy (0/1) of the CURRENT cell (shared by the step and the instantaneous cost)
__device__ __forceinline__ void particle_step(const ModelParams& m, ParticleState& s, float ax, float ay, float mass,
                                              float c, float* vpre = nullptr, float* apre = nullptr) {
  const float qx = ax / mass, qy = ay / mass;
  if (apre) { apre[0] = qx; apre[1] = qy; }   // pre-clamp accelerations (the adjoint needs the masks)
  const float ux = fminf(fmaxf(qx, -m.max_accel), m.max_accel);
  const float uy = fminf(fmaxf(qy, -m.max_accel), m.max_accel);
  float nx, ny, nvx, nvy;
  if (m.can_crash) {
    const float k = 1.0f - c;
    nx = s.x + (s.vx * m.dt) * k;
    ny = s.y + (s.vy * m.dt) * k;
    nvx = s.vx + (ux * m.dt) * k;
    nvy = s.vy + (uy * m.dt) * k;
  } else {
    nx = s.x + s.vx * m.dt;
    ny = s.y + s.vy * m.dt;
    nvx = s.vx + ux * m.dt;
    nvy = s.vy + uy * m.dt;
  }
  if (vpre) { vpre[0] = nvx; vpre[1] = nvy; }
  s.x = nx;
  s.y = ny;
  s.vx = fminf(fmaxf(nvx, -m.max_speed), m.max_speed);
  s.vy = fminf(fmaxf(nvy, -m.max_speed), m.max_speed);
}

__device__ __forceinline__ float particle_quad(const ParticleState& s, const float* tgt, const float* w) {
  const float d0 = s.x - tgt[0], d1 = s.y - tgt[1], d2 = s.vx - tgt[2], d3 = s.vy - tgt[3];
  float acc = (d0 * d0) * w[0];
  acc = acc + (d1 * d1) * w[1];
  acc = acc + (d2 * d2) * w[2];
  acc = acc + (d3 * d3) * w[3];
  return acc;
}

// particle.py:170-198 (raw actions in the control term); c = occupancy of the current cell
__device__ __forceinline__ float particle_inst_cost(const ModelParams& m, const ParticleState& s, float ax, float ay, float c) {
  const float sc = particle_quad(s, m.target, m.w_state);
  const float cc = (ax * ax) * m.w_ctrl[0] + (ay * ay) * m.w_ctrl[1];
  float cost = sc + cc;
  if (m.with_obstacle) cost = cost + m.w_obs * c;
  return cost;
}
// particle.py:202-225
__device__ __forceinline__ float particle_term_cost(const ModelParams& m, const ParticleState& s, float c) {
  float cost = particle_quad(s, m.target, m.w_term);
  if (m.with_obstacle) cost = cost + m.w_obs * c;
  return cost;
}

}  // namespace dust
