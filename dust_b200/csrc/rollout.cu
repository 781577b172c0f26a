// K1: batched rollout + trajectory cost, then the per-policy soft-min / likelihood reductions.
// One thread per trajectory (sample s, policy n) of one MPC instance; the state lives in
// registers, the action tile is staged through shared memory, parameter samples are looped
// inside the thread so the mean over P needs no communication.
//
// Reference: dust/controllers/disco.py:139-209 (rollout), :294-346 (cost), :380-393 (soft-min),
// dust/inference/likelihoods.py:81-135, dust/inference/svmpc.py:46-54.
// Compiled with -fmad=false (see models.cuh).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "models.cuh"
#include "svgd_dev.cuh"

namespace dust {

constexpr int kTile = 128;  // trajectories (threads) per CTA

__host__ __device__ inline int padded_stride(int HA) {
  // row stride (floats) of the action tile in shared memory: a multiple of 4 whose quarter is
  // odd, so that 16-byte row reads by consecutive threads hit distinct bank groups.
  int s4 = (HA + 3) / 4;
  if ((s4 & 1) == 0) s4 += 1;
  return s4 * 4;
}

struct RolloutKParams {
  ModelParams m;
  int B, N, S, P, H, A;
  int SN;          // S*N trajectories per instance
  int HA;          // H*A
  int PC;          // number of parameter chunks (grid.y)
  int Pchunk;      // parameters per chunk
  int parts;       // partial cost rows per instance = PC * (sub-chunks per CTA); 1: the CTA writes final costs
  int p0, p1;      // draws [p0, p1) of the P resident ones are rolled out by this call (a rank's share)
  int interleaved;
  const float* state0;
  const float* theta;
  const float* noise;
  const float* sigma;
  const float* params;
  float* cost_out;     // PC==1: costs [B,S,N] (already divided by P); else partial sums [B,PC,SN]
  float* states;       // optional [B,P,S,N,H+1,ds]
  const float* ut_w;   // [P] unscented-transform weights (sigma-point mode, PC == 1) or nullptr
  const float* ctrl_mat;  // [B,N,HA] a_mat @ a_pre (control regulariser, PC == 1) or nullptr
  const float* a_seq;  // [B,HA] or nullptr (zeros)
  float ctrl_reg;
};

// Weighted cost of the sigma-point mode exactly as disco.py:312-323 groups it: the instantaneous costs of
// one (sample, policy) pair, in the order (sigma point p, step t), are cut into runs of P; every run is
// one dot product with the weights, the runs are summed, and the weighted terminal costs are added.
struct UtCostSum {
  const float* w;
  int P, i = 0;
  float dot = 0.f, inst = 0.f, term = 0.f;
  __device__ __forceinline__ void add(float c) {
    dot = dot + __ldg(w + i) * c;
    if (++i == P) { inst = inst + dot; dot = 0.f; i = 0; }
  }
  __device__ __forceinline__ void add_term(int p, float c) { term = term + __ldg(w + p) * c; }
  __device__ __forceinline__ float total() const { return inst + term; }
};

// cooperative load of a [rows, HA] tile of the noise tensor into padded shared memory, fused
// with actions = theta + sigma * eps (likelihoods.py:85-90; exact: one product, one sum).
// The (row, column, policy) walk of each thread is incremental: no division in the loop.
// th: policy means of this instance (global or shared memory), or nullptr.
template <int A, int NT>
__device__ __forceinline__ void load_action_tile(const RolloutKParams& k, float* tile, int stride, long long inst,
                                                 int j0, int rows, const float* __restrict__ th) {
  const int HA = k.HA, N = k.N;
  const float* __restrict__ src = k.noise + (inst * k.SN + j0) * (long long)HA;
  const bool vec = ((HA & 3) == 0) && ((((uintptr_t)src) & 15) == 0) && (!th || ((((uintptr_t)th) & 15) == 0));
  if (vec) {
    const int HA4 = HA >> 2;
    const int total4 = rows * HA4;
    float sg[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) sg[q] = th ? k.sigma[q % A] : 0.f;
    const int drow = NT / HA4, dc = NT - drow * HA4, dn = drow % N;
    int row = (int)threadIdx.x / HA4;
    int c4 = (int)threadIdx.x - row * HA4;
    int n = (j0 + row) % N;
#pragma unroll 2
    for (int e = threadIdx.x; e < total4; e += NT) {
      float4 v = __ldg(reinterpret_cast<const float4*>(src) + e);
      if (th) {
        const float4 t = *reinterpret_cast<const float4*>(th + n * HA + 4 * c4);
        v.x = t.x + sg[0] * v.x;
        v.y = t.y + sg[1] * v.y;
        v.z = t.z + sg[2] * v.z;
        v.w = t.w + sg[3] * v.w;
      }
      *reinterpret_cast<float4*>(tile + row * stride + 4 * c4) = v;
      c4 += dc;
      row += drow;
      n += dn;
      if (c4 >= HA4) { c4 -= HA4; row += 1; n += 1; }
      if (n >= N) n -= N;
      if (n >= N) n -= N;
    }
  } else {
    const int total = rows * HA;
    const int drow = NT / HA, dc = NT - drow * HA, dn = drow % N;
    int row = (int)threadIdx.x / HA;
    int c = (int)threadIdx.x - row * HA;
    int n = (j0 + row) % N;
    for (int e = threadIdx.x; e < total; e += NT) {
      float v = __ldg(src + e);
      if (th) v = th[n * HA + c] + k.sigma[c % A] * v;
      tile[row * stride + c] = v;
      c += dc;
      row += drow;
      n += dn;
      if (c >= HA) { c -= HA; row += 1; n += 1; }
      if (n >= N) n -= N;
      if (n >= N) n -= N;
    }
  }
}

// SMALL: |theta| stays below the fast trig range for the whole horizon (decided per instance);
// STORE: write the state trajectory (only the non-fused path offers it).
// XFORM: `arow` holds the raw noise; the action is th_row[t] + sigma * eps[t], formed per step
// (one product, one sum: the same float32 value as MultivariateNormal.rsample's loc + L eps).
template <int MODEL, bool SMALL, bool STORE, bool XFORM = false, bool UT = false>
__device__ __forceinline__ float trajectory_cost_sum(const RolloutKParams& k, const float* __restrict__ arow,
                                                     const uint32_t* grid_s, long long inst, int j, int p_begin, int p_end,
                                                     const float* __restrict__ th_row = nullptr, float sg0 = 0.f,
                                                     float sg1 = 0.f) {
  constexpr int DS = (MODEL == DUST_MODEL_PENDULUM) ? 2 : 4;
  constexpr int DP = (MODEL == DUST_MODEL_PENDULUM) ? 2 : 1;
  const float* __restrict__ x0 = k.state0 + inst * DS;
  float csum = 0.f;
  UtCostSum ut{k.ut_w, k.P};
  for (int p = p_begin; p < p_end; ++p) {
    const float* prm = nullptr;
    if (k.params) {
      const int pi = k.interleaved ? (int)(((long long)p * k.SN + j) % k.P) : p;
      prm = k.params + (inst * k.P + pi) * DP;
    }
    float* st_out = nullptr;
    if (STORE) st_out = k.states + (((inst * k.P + p) * k.SN + j) * (long long)(k.H + 1)) * DS;
    float cost = 0.f;
    if (MODEL == DUST_MODEL_PENDULUM) {
      const PendulumCoef cf = prm ? pendulum_coef_sampled(k.m, __ldg(prm), __ldg(prm + 1)) : pendulum_coef_default(k.m);
      float th = __ldg(x0), om = __ldg(x0 + 1);
      if (STORE) { st_out[0] = th; st_out[1] = om; }
      // cost at x_t uses cos(th_t), which the step evaluates from the reduction it needs anyway
      PendulumCostSum run;
      auto one = [&](float a, int t) {
        const float om0 = om;
        float cth;
        pendulum_step<!SMALL>(k.m, cf, th, om, a, nullptr, &cth);
        if (UT) ut.add(pendulum_cost_from_cos(k.m, cth, om0));
        else run.add(k.m, cth, om0);
        if (STORE) { st_out[(t + 1) * 2] = th; st_out[(t + 1) * 2 + 1] = om; }
      };
      int t = 0;
      for (; t + 4 <= k.H; t += 4) {
        float4 a4 = *reinterpret_cast<const float4*>(arow + t);
        if (XFORM) {
          const float4 t4 = *reinterpret_cast<const float4*>(th_row + t);
          a4.x = t4.x + sg0 * a4.x; a4.y = t4.y + sg0 * a4.y; a4.z = t4.z + sg0 * a4.z; a4.w = t4.w + sg0 * a4.w;
        }
        one(a4.x, t); one(a4.y, t + 1); one(a4.z, t + 2); one(a4.w, t + 3);
      }
      for (; t < k.H; ++t) one(XFORM ? th_row[t] + sg0 * arow[t] : arow[t], t);
      if (UT) ut.add_term(p, pendulum_cost<!SMALL>(k.m, th, om));
      else cost = run.total(k.m) + pendulum_cost<!SMALL>(k.m, th, om);
    } else {
      const float mass = prm ? __ldg(prm) : k.m.default_mass;
      ParticleState s{__ldg(x0), __ldg(x0 + 1), __ldg(x0 + 2), __ldg(x0 + 3)};
      if (STORE) { st_out[0] = s.x; st_out[1] = s.y; st_out[2] = s.vx; st_out[3] = s.vy; }
      const bool has_grid = k.m.grid_bits != nullptr;
      auto one = [&](float ax, float ay, int t) {
        const float c = has_grid ? grid_lookup(k.m, grid_s, s.x, s.y) : 0.f;
        if (UT) ut.add(particle_inst_cost(k.m, s, ax, ay, c));
        else cost = cost + particle_inst_cost(k.m, s, ax, ay, c);
        particle_step(k.m, s, ax, ay, mass, c);
        if (STORE) {
          float* o = st_out + (t + 1) * 4;
          o[0] = s.x; o[1] = s.y; o[2] = s.vx; o[3] = s.vy;
        }
      };
      int t = 0;
      for (; t + 2 <= k.H; t += 2) {
        float4 a4 = *reinterpret_cast<const float4*>(arow + 2 * t);
        if (XFORM) {
          const float4 t4 = *reinterpret_cast<const float4*>(th_row + 2 * t);
          a4.x = t4.x + sg0 * a4.x; a4.y = t4.y + sg1 * a4.y; a4.z = t4.z + sg0 * a4.z; a4.w = t4.w + sg1 * a4.w;
        }
        one(a4.x, a4.y, t); one(a4.z, a4.w, t + 1);
      }
      for (; t < k.H; ++t) {
        if (XFORM) one(th_row[2 * t] + sg0 * arow[2 * t], th_row[2 * t + 1] + sg1 * arow[2 * t + 1], t);
        else one(arow[2 * t], arow[2 * t + 1], t);
      }
      const float c = has_grid ? grid_lookup(k.m, grid_s, s.x, s.y) : 0.f;
      if (UT) ut.add_term(p, particle_term_cost(k.m, s, c));
      else cost = cost + particle_term_cost(k.m, s, c);
    }
    csum = csum + cost;
  }
  return UT ? ut.total() : csum;
}

#if DUST_PEND_PAIR
// rows jA and jB (same policy, hence the same theta row) rolled out together on the packed FP32
// instructions; XFORM actions th_row[t] + sigma * eps[t].  Small-angle horizon only.
__device__ __forceinline__ float2 pendulum_pair_cost_sum(const RolloutKParams& k, const float* __restrict__ rowA,
                                                         const float* __restrict__ rowB, long long inst, int jA, int jB,
                                                         const float* __restrict__ th_row, float sg0, float th0, float om0) {
  float2 csum = make_float2(0.f, 0.f);
  const float2 sg = bc2(sg0);
  for (int p = 0; p < k.P; ++p) {
    PendulumCoef ca, cb;
    if (k.params) {
      const int pa = k.interleaved ? (int)(((long long)p * k.SN + jA) % k.P) : p;
      const int pb = k.interleaved ? (int)(((long long)p * k.SN + jB) % k.P) : p;
      const float* qa = k.params + (inst * k.P + pa) * 2;
      const float* qb = k.params + (inst * k.P + pb) * 2;
      ca = pendulum_coef_sampled(k.m, __ldg(qa), __ldg(qa + 1));
      cb = (pb == pa) ? ca : pendulum_coef_sampled(k.m, __ldg(qb), __ldg(qb + 1));
    } else {
      ca = cb = pendulum_coef_default(k.m);
    }
    const PendulumCoef2 cf{make_float2(ca.c1, cb.c1), make_float2(ca.c2, cb.c2)};
    float2 th = bc2(th0), om = bc2(om0), run = make_float2(0.f, 0.f);
    int t = 0;
    for (; t + 4 <= k.H; t += 4) {
      const float4 ea = *reinterpret_cast<const float4*>(rowA + t);
      const float4 eb = *reinterpret_cast<const float4*>(rowB + t);
      const float4 t4 = *reinterpret_cast<const float4*>(th_row + t);
      pendulum_step_pair(k.m, cf, th, om, add2(bc2(t4.x), mul2_unfused(sg, make_float2(ea.x, eb.x))), run);
      pendulum_step_pair(k.m, cf, th, om, add2(bc2(t4.y), mul2_unfused(sg, make_float2(ea.y, eb.y))), run);
      pendulum_step_pair(k.m, cf, th, om, add2(bc2(t4.z), mul2_unfused(sg, make_float2(ea.z, eb.z))), run);
      pendulum_step_pair(k.m, cf, th, om, add2(bc2(t4.w), mul2_unfused(sg, make_float2(ea.w, eb.w))), run);
    }
    for (; t < k.H; ++t)
      pendulum_step_pair(k.m, cf, th, om, add2(bc2(th_row[t]), mul2_unfused(sg, make_float2(rowA[t], rowB[t]))), run);
    const float2 term = pendulum_cost_pair(k.m, th, om);
    csum = add2(csum, add2(run, term));
  }
  return csum;
}
#endif

#if DUST_PEND_PAIR
// Two pairs per thread (rows A, B and C, D; all four of the same policy): the two packed chains are independent, so
// the compiler interleaves them -- one chain's quadrant logic (ALU pipe) under the other's polynomials (FMA pipe).
// Per trajectory the same instructions as pendulum_pair_cost_sum: bit-identical costs.
__device__ __forceinline__ void pendulum_quad_cost_sum(const RolloutKParams& k, const float* __restrict__ const (&row)[4], long long inst,
                                                       const int (&j)[4], const float* __restrict__ th_row, float sg0, float th0,
                                                       float om0, float2 (&out)[2]) {
  float2 csum[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  const float2 sg = bc2(sg0);
  for (int p = 0; p < k.P; ++p) {
    PendulumCoef2 cf[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      PendulumCoef ca, cb;
      if (k.params) {
        const int pa = k.interleaved ? (int)(((long long)p * k.SN + j[2 * q]) % k.P) : p;
        const int pb = k.interleaved ? (int)(((long long)p * k.SN + j[2 * q + 1]) % k.P) : p;
        const float* qa = k.params + (inst * k.P + pa) * 2;
        const float* qb = k.params + (inst * k.P + pb) * 2;
        ca = pendulum_coef_sampled(k.m, __ldg(qa), __ldg(qa + 1));
        cb = (pb == pa) ? ca : pendulum_coef_sampled(k.m, __ldg(qb), __ldg(qb + 1));
      } else {
        ca = cb = pendulum_coef_default(k.m);
      }
      cf[q] = PendulumCoef2{make_float2(ca.c1, cb.c1), make_float2(ca.c2, cb.c2)};
    }
    float2 th[2] = {bc2(th0), bc2(th0)}, om[2] = {bc2(om0), bc2(om0)}, run[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    int t = 0;
    for (; t + 4 <= k.H; t += 4) {
      const float4 t4 = *reinterpret_cast<const float4*>(th_row + t);
      float4 e[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) e[r] = *reinterpret_cast<const float4*>(row[r] + t);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 ea = e[2 * q], eb = e[2 * q + 1];
        pendulum_step_pair(k.m, cf[q], th[q], om[q], add2(bc2(t4.x), mul2_unfused(sg, make_float2(ea.x, eb.x))), run[q]);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 ea = e[2 * q], eb = e[2 * q + 1];
        pendulum_step_pair(k.m, cf[q], th[q], om[q], add2(bc2(t4.y), mul2_unfused(sg, make_float2(ea.y, eb.y))), run[q]);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 ea = e[2 * q], eb = e[2 * q + 1];
        pendulum_step_pair(k.m, cf[q], th[q], om[q], add2(bc2(t4.z), mul2_unfused(sg, make_float2(ea.z, eb.z))), run[q]);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 ea = e[2 * q], eb = e[2 * q + 1];
        pendulum_step_pair(k.m, cf[q], th[q], om[q], add2(bc2(t4.w), mul2_unfused(sg, make_float2(ea.w, eb.w))), run[q]);
      }
    }
    for (; t < k.H; ++t) {
#pragma unroll
      for (int q = 0; q < 2; ++q)
        pendulum_step_pair(k.m, cf[q], th[q], om[q],
                           add2(bc2(th_row[t]), mul2_unfused(sg, make_float2(row[2 * q][t], row[2 * q + 1][t]))), run[q]);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) csum[q] = add2(csum[q], add2(run[q], pendulum_cost_pair(k.m, th[q], om[q])));
  }
  out[0] = csum[0];
  out[1] = csum[1];
}
#endif

// |theta_t + pi| <= |theta_0| + t dt max_speed + pi: decided once per instance (uniform in the CTA)
template <int MODEL>
__device__ __forceinline__ bool small_angle_horizon(const RolloutKParams& k, long long inst) {
  if (MODEL != DUST_MODEL_PENDULUM) return true;
  const float th0 = __ldg(k.state0 + inst * 2);
  return fabsf(th0) + (float)k.H * k.m.dt * k.m.max_speed_pend + 3.2f <= 64.0f;  // NaN -> false
}

// EXT: the instantiation that also knows the sigma-point weights and the control regulariser; the plain one
// (every hot path) carries none of that code
template <int MODEL, bool EXT>
__device__ __forceinline__ float trajectory_cost_dispatch(const RolloutKParams& k, const float* __restrict__ arow,
                                                          const uint32_t* grid_s, long long inst, int j, int p_begin,
                                                          int p_end, bool small) {
  if (EXT && k.ut_w) {
    if (k.states) return trajectory_cost_sum<MODEL, false, true, false, true>(k, arow, grid_s, inst, j, p_begin, p_end);
    return trajectory_cost_sum<MODEL, false, false, false, true>(k, arow, grid_s, inst, j, p_begin, p_end);
  }
  if (k.states) return trajectory_cost_sum<MODEL, false, true>(k, arow, grid_s, inst, j, p_begin, p_end);
  if (small) return trajectory_cost_sum<MODEL, true, false>(k, arow, grid_s, inst, j, p_begin, p_end);
  return trajectory_cost_sum<MODEL, false, false>(k, arow, grid_s, inst, j, p_begin, p_end);
}

// NSUB: sub-chunks per CTA.  The action tile of 128 trajectories (52 KB at H*A = 100) limits an SM to a few
// CTAs; with NSUB > 1 the tile is shared by NSUB groups of 128 threads, each rolling out its own slice of the
// CTA's parameter chunk into its own partial row (fixed order of partial rows: deterministic).
template <int MODEL, bool EXT, int NSUB>
__global__ void __launch_bounds__(kTile * NSUB) rollout_cost_kernel(const RolloutKParams k) {
  constexpr int A = (MODEL == DUST_MODEL_PENDULUM) ? 1 : 2;
  extern __shared__ __align__(16) float smem[];
  const int stride = padded_stride(k.HA);
  float* tile = smem;  // [kTile][stride]
  uint32_t* grid_s = reinterpret_cast<uint32_t*>(smem + kTile * stride);

  const int tiles_per_inst = (k.SN + kTile - 1) / kTile;
  const long long inst = blockIdx.x / tiles_per_inst;
  const int j0 = (blockIdx.x - (int)inst * tiles_per_inst) * kTile;
  const int rows = min(kTile, k.SN - j0);
  const int pc = blockIdx.y;

  load_action_tile<A, kTile * NSUB>(k, tile, stride, inst, j0, rows,
                                    k.theta ? k.theta + inst * (long long)k.N * k.HA : nullptr);
  if (MODEL == DUST_MODEL_PARTICLE && k.m.grid_bits != nullptr) {
    const int words = (k.m.grid_nx * k.m.grid_ny + 31) >> 5;
    for (int w = threadIdx.x; w < words; w += kTile * NSUB) grid_s[w] = __ldg(k.m.grid_bits + w);
  }
  __syncthreads();

  const int row = threadIdx.x % kTile, sub = threadIdx.x / kTile;
  if (row >= rows) return;
  const int j = j0 + row;
  int p_begin = k.p0 + pc * k.Pchunk;
  int p_end = min(k.p1, p_begin + k.Pchunk);
  if (NSUB > 1) {   // this thread group's slice of the chunk (possibly empty: it then contributes an exact 0)
    const int sl = (p_end - p_begin + NSUB - 1) / NSUB;
    p_begin += sub * sl;
    p_end = min(p_end, p_begin + sl);
  }
  const float csum = trajectory_cost_dispatch<MODEL, EXT>(k, tile + row * stride, grid_s, inst, j, p_begin, p_end,
                                                          small_angle_horizon<MODEL>(k, inst));
  if (k.parts == 1) {
    float cost = (EXT && k.ut_w) ? csum : csum / (float)k.P;  // sigma-point weights, or the mean over parameter samples (disco.py:330)
    if (EXT && k.ctrl_mat) {
      // control regulariser (disco.py:334-344): a_reg * sum_{h,a} -(action - a_seq) (a_mat a_pre)[n]
      const float* act = tile + row * stride;
      const float* cm = k.ctrl_mat + (inst * k.N + j % k.N) * (long long)k.HA;
      const float* as = k.a_seq ? k.a_seq + inst * k.HA : nullptr;
      float dot = 0.f;
      for (int e = 0; e < k.HA; ++e) dot = fmaf(-(act[e] - (as ? __ldg(as + e) : 0.f)), __ldg(cm + e), dot);
      cost = cost + k.ctrl_reg * dot;
    }
    k.cost_out[inst * k.SN + j] = cost;
  } else {
    k.cost_out[(inst * k.parts + pc * NSUB + sub) * (long long)k.SN + j] = csum;
  }
}

// ---------------------------------------------------------------------------------------
// fused per-instance kernel: one CTA owns ALL trajectories of one MPC instance.  Every thread
// keeps an online soft-min (running max, normaliser, weighted score row) over the trajectories it
// rolls out, so the noise is read from HBM exactly once and costs / log-likelihood / analytic
// likelihood gradient (svmpc.py:46-54) leave in the same launch.
// ---------------------------------------------------------------------------------------
constexpr int kFusedThreads = 256;

// optional tail of the fused kernel: the rest of the SVGD step (and of SVMPC.forward) for the
// instance this CTA owns -- prior score, phi, SGD update, weights / argmax / shift
// (dust/inference/svmpc.py:38-95, 128-200).
struct TailParams {
  int enabled, do_forward, roll, weighted, aliased;
  const float* mu;       // [B,N,D] prior centres (ignored when aliased: centres = particles)
  const float* mix;      // [B,N] unnormalised mixture weights or null
  const float* inv_var;  // [D]
  float log_norm, gamma, c1, c2, lr;
  float* theta_out;      // [B,N,D] updated particles (before the shift) or null
  float* phi;            // [B,N,D] or null
  float* p_weights;      // [B,N]
  int* i_star;           // [B]
  float* a_seq;          // [B,D]
  float* theta_next;     // [B,N,D] shifted particles
  float* mix_next;       // [B,N]
};

struct FusedOut {
  float* costs;     // [B,S,N] or null
  float* log_lik;   // [B,N] or null
  float* grad_lik;  // [B,N,HA] or null
  int likelihood;
  float alpha;
  TailParams tail;
};

#ifndef DUST_FUSED_MINB
#define DUST_FUSED_MINB 4
#endif
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N_) : "memory"); }

// asynchronous copy of a [rows, HA] block of raw noise into a padded shared-memory tile; the walk
// (row0, c0, drow, dc) is the thread's fixed stride pattern, computed once per kernel.
template <bool VEC, int NT>
__device__ __forceinline__ void stage_noise_async(const float* __restrict__ src, float* tile, int stride, int rows, int HA,
                                                  int row0, int c0, int drow, int dc) {
  const int W = VEC ? (HA >> 2) : HA;  // columns in units of the copy width
  const int total = rows * W;
  int row = row0, c = c0;
  for (int e = threadIdx.x; e < total; e += NT) {
    if (VEC) cp_async16(tile + row * stride + 4 * c, reinterpret_cast<const float4*>(src) + e);
    else cp_async4(tile + row * stride + c, src + e);
    c += dc;
    row += drow;
    if (c >= W) { c -= W; row += 1; }
  }
}

// TPT: trajectories per thread.  2 = the packed pendulum path: NT = 128 threads, thread tid owns rows
// tid and tid + TN/2 of every 256-row tile (same policy: the host checks that TN/2 is a multiple of N).
template <int MODEL, int ACC, int TPT>
__global__ void __launch_bounds__(kFusedThreads / TPT, TPT == 2 ? 5 : DUST_FUSED_MINB) svmpc_instance_kernel(const RolloutKParams k, const FusedOut o) {
  constexpr int A = (MODEL == DUST_MODEL_PENDULUM) ? 1 : 2;
  constexpr int NT = kFusedThreads / TPT;   // threads of the CTA
  static_assert(TPT == 1 || (MODEL == DUST_MODEL_PENDULUM && DUST_PEND_PAIR), "the packed path is the pendulum's");
  extern __shared__ __align__(16) float smem[];
  const int stride = padded_stride(k.HA);
  const int HA = k.HA, N = k.N;
  const int TN = (kFusedThreads / N) * N;  // rows per tile: a multiple of N, so tid % N is this thread's policy
  float* buf0 = smem;                                   // [256][stride] noise tile, double buffered
  float* buf1 = buf0 + kFusedThreads * stride;
  float* th_s = buf1 + kFusedThreads * stride;          // [N][thst] policy means
  const int thst = (HA + 3) & ~3;                       // theta row stride (16-byte aligned rows)
  float* red_m = buf1;                                  // [256] the reductions after the tile loop live in the
  float* red_z = red_m + kFusedThreads;                 // [256] second noise buffer, which is dead by then
  float* red_c = red_z + kFusedThreads;                 // [256] (stride >= 4 floats per row: 1024 >= 768)
  uint32_t* grid_s = reinterpret_cast<uint32_t*>(th_s + N * thst);
  const long long inst = blockIdx.x;
  const int tid = threadIdx.x;
  const int n = tid % N;
  const float* __restrict__ noise = k.noise + inst * (long long)k.SN * HA;
  const bool vec = ((HA & 3) == 0) && ((((uintptr_t)noise) & 15) == 0);
  const int W = vec ? (HA >> 2) : HA;
  const int row0 = tid / W, c0 = tid - row0 * W;
  const int drow = NT / W, dc = NT - drow * W;
  const int ntiles = (k.SN + TN - 1) / TN;

  // unpadded rows (stride == HA): a tile is one contiguous block of global memory, fetched by a
  // single 1-D bulk copy (TMA) that one thread issues and an mbarrier completes -- no per-thread
  // copy instructions or address arithmetic.  Padded rows keep the 16-byte cp.async walk.
  __shared__ __align__(8) uint64_t tile_bar[2];   // "tile landed" (TMA complete_tx)
  const bool bulk = vec && stride == HA;
  auto prefetch = [&](int it) {
    const int j0 = it * TN;
    const int rows = min(TN, k.SN - j0);
    float* dst = (it & 1) ? buf1 : buf0;
    if (bulk) {
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)rows * (uint32_t)HA * 4u;
        mbar_expect_tx(&tile_bar[it & 1], bytes);
        bulk_g2s(dst, noise + (long long)j0 * HA, bytes, &tile_bar[it & 1]);
      }
      return;
    }
    if (vec) stage_noise_async<true, NT>(noise + (long long)j0 * HA, dst, stride, rows, HA, row0, c0, drow, dc);
    else stage_noise_async<false, NT>(noise + (long long)j0 * HA, dst, stride, rows, HA, row0, c0, drow, dc);
    cp_async_commit();
  };
  if (bulk && tid == 0) {
    mbar_init(&tile_bar[0], 1);
    mbar_init(&tile_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  prefetch(0);
  for (int e = tid; e < N * HA; e += NT) {
    const int nn = e / HA;
    th_s[nn * thst + (e - nn * HA)] = k.theta[inst * (long long)N * HA + e];
  }
  if (MODEL == DUST_MODEL_PARTICLE && k.m.grid_bits != nullptr) {
    const int words = (k.m.grid_nx * k.m.grid_ny + 31) >> 5;
    for (int w = tid; w < words; w += NT) grid_s[w] = __ldg(k.m.grid_bits + w);
  }
  __syncthreads();  // theta / grid / barrier initialisation visible to everyone
  const float sg0 = k.sigma[0], sg1 = k.sigma[A - 1];
  const float is0 = 1.0f / (sg0 * sg0), is1 = 1.0f / (sg1 * sg1);
  const bool small = small_angle_horizon<MODEL>(k, inst);
  const float* __restrict__ th_row = th_s + n * thst;

  // online soft-min relative to the running MINIMUM cost: weights are exp(-alpha (c - c_min)) with
  // the difference formed first (exact), as softmax(-alpha c) does after its max shift
  float m_run = INFINITY, z_run = 0.f, c_run = 0.f;
  float acc[ACC];
#pragma unroll
  for (int c = 0; c < ACC; ++c) acc[c] = 0.f;

  for (int it = 0; it < ntiles; ++it) {
    const int j0 = it * TN;
    const int rows = min(TN, k.SN - j0);
    if (it + 1 < ntiles) prefetch(it + 1);  // overlaps this tile's rollouts
    if (bulk) {
      mbar_wait(&tile_bar[it & 1], (uint32_t)(it >> 1) & 1u);
    } else {
      if (it + 1 < ntiles) cp_async_wait<1>();
      else cp_async_wait<0>();
      __syncthreads();
    }
    const float* tile = (it & 1) ? buf1 : buf0;
    // fold one finished trajectory (row `erow`, index j, summed cost csum) into this thread's running
    // soft-min state; both rows of a packed thread belong to the same policy and share that state
    auto fold = [&](float csum, const float* __restrict__ erow, int j) {
      const float cost = (k.P == 1) ? csum : csum / (float)k.P;
      if (o.costs) o.costs[inst * k.SN + j] = cost;
      c_run += cost;
      // exactly one of {rescale of the running sums, weight of this trajectory} differs from 1
      const float dlt = cost - m_run;                 // -inf on the first trajectory
      const float ex = expf(-o.alpha * fabsf(dlt));   // 0 there
      const bool lower = dlt < 0.f;
      const float scale = lower ? ex : 1.f, e = lower ? 1.f : ex;
      if (lower) m_run = cost;
      z_run = z_run * scale + e;
      // weighted score row (a - theta)/sigma^2 = eps/sigma (the reference's a - theta differs from
      // sigma*eps by the rounding of the sum, ~1e-7 relative); skipped when the weight is below
      // 1e-30 of the running maximum weight (invisible in float32)
      if (scale != 1.f || e > 1e-30f) {
#if DUST_PEND_OPT & 16
        const float e0 = e * (is0 * sg0), e1 = e * (is1 * sg1);
#define DUST_SCORE_TERM(w, sg, th, v) ((w) * (v))
#else
        const float e0 = e * is0, e1 = e * is1;
#define DUST_SCORE_TERM(w, sg, th, v) ((w) * (((th) + (sg) * (v)) - (th)))
#endif
        if ((HA & 3) == 0) {
#pragma unroll
          for (int c4 = 0; c4 < ACC / 4; ++c4) {
            if (4 * c4 < HA) {
              const float4 v = *reinterpret_cast<const float4*>(erow + 4 * c4);
#if !(DUST_PEND_OPT & 16)
              const float4 t4 = *reinterpret_cast<const float4*>(th_row + 4 * c4);
#endif
              const float sa = sg0, sb = (A == 1) ? sg0 : sg1;
              const float ea = e0, eb = (A == 1) ? e0 : e1;
              (void)sa; (void)sb;
              acc[4 * c4 + 0] = fmaf(acc[4 * c4 + 0], scale, DUST_SCORE_TERM(ea, sa, t4.x, v.x));
              acc[4 * c4 + 1] = fmaf(acc[4 * c4 + 1], scale, DUST_SCORE_TERM(eb, sb, t4.y, v.y));
              acc[4 * c4 + 2] = fmaf(acc[4 * c4 + 2], scale, DUST_SCORE_TERM(ea, sa, t4.z, v.z));
              acc[4 * c4 + 3] = fmaf(acc[4 * c4 + 3], scale, DUST_SCORE_TERM(eb, sb, t4.w, v.w));
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < ACC; ++c)
            if (c < HA) acc[c] = fmaf(acc[c], scale, DUST_SCORE_TERM((c % A) ? e1 : e0, (c % A) ? sg1 : sg0, th_row[c], erow[c]));
        }
#undef DUST_SCORE_TERM
      }
    };
    if (TPT == 1) {
      if (tid < rows) {
        const float* __restrict__ erow = tile + tid * stride;
        const float csum = small
            ? trajectory_cost_sum<MODEL, true, false, true>(k, erow, grid_s, inst, j0 + tid, 0, k.P, th_row, sg0, sg1)
            : trajectory_cost_sum<MODEL, false, false, true>(k, erow, grid_s, inst, j0 + tid, 0, k.P, th_row, sg0, sg1);
        fold(csum, erow, j0 + tid);
      }
    } else {
#if DUST_PEND_PAIR
      const int half = TN >> 1;
      const int ra = tid, rb = tid + half;
      if (tid < half && ra < rows) {
        const float* __restrict__ rowA = tile + ra * stride;
        const float* __restrict__ rowB = tile + rb * stride;
        const int nv = (rb < rows) ? 2 : 1;
        float csA, csB = 0.f;
        if (small && nv == 2) {
          const float2 cs = pendulum_pair_cost_sum(k, rowA, rowB, inst, j0 + ra, j0 + rb, th_row, sg0,
                                                   __ldg(k.state0 + inst * 2), __ldg(k.state0 + inst * 2 + 1));
          csA = cs.x;
          csB = cs.y;
        } else {
          // ragged last tile, or an angle beyond the fast range: the scalar step, same arithmetic.
          // Rare: kept as ONE copy of the code (no unrolling) so the kernel stays small.
          csA = 0.f;
#pragma unroll 1
          for (int q = 0; q < nv; ++q) {
            const float* __restrict__ row = q ? rowB : rowA;
            const int jj = j0 + (q ? rb : ra);
            const float v = small ? trajectory_cost_sum<MODEL, true, false, true>(k, row, grid_s, inst, jj, 0, k.P, th_row, sg0, sg1)
                                  : trajectory_cost_sum<MODEL, false, false, true>(k, row, grid_s, inst, jj, 0, k.P, th_row, sg0, sg1);
            if (q) csB = v; else csA = v;
          }
        }
#pragma unroll 1
        for (int q = 0; q < nv; ++q) fold(q ? csB : csA, q ? rowB : rowA, j0 + (q ? rb : ra));
      }
#endif
    }
    // everyone is done with this buffer before it is refilled (a per-warp mbarrier release that lets
    // warps run a tile ahead measured 2 % SLOWER than this plain barrier)
    __syncthreads();
  }
  // combine the G = TN/N threads that share a policy
  const int TNT = TN / TPT;  // threads that own trajectories; tid % N is their policy
  red_m[tid] = (tid < TNT) ? m_run : INFINITY;
  red_c[tid] = (tid < TNT) ? c_run : 0.f;
  __syncthreads();
  const int G = TNT / N;
  float m_n = INFINITY;
  for (int g = 0; g < G; ++g) m_n = fminf(m_n, red_m[g * N + n]);
  // every thread rescales its OWN partial normaliser to the policy minimum (one exponential per
  // thread instead of G), then the G slots of a policy are summed in a fixed order
  const float own = (tid < TNT && m_run != INFINITY) ? expf(-o.alpha * (m_run - m_n)) : 0.f;
  red_z[tid] = z_run * own;
  __syncthreads();
  float z_n = 0.f, c_n = 0.f;
  for (int g = 0; g < G; ++g) z_n += red_z[g * N + n];
  if (tid < N)
    for (int g = 0; g < G; ++g) c_n += red_c[g * N + n];
  if (o.log_lik && tid < N) {
    float ll;
    if (o.likelihood == DUST_LIK_EXP_UTILITY) ll = (-o.alpha * m_n + logf(z_n)) - logf((float)k.S);  // likelihoods.py:133-135
    else ll = -o.alpha * (c_n / (float)k.S);                                                          // likelihoods.py:119
    o.log_lik[inst * N + tid] = ll;
  }
  float* tail_s = reinterpret_cast<float*>(grid_s + ((MODEL == DUST_MODEL_PARTICLE && k.m.grid_bits) ? ((k.m.grid_nx * k.m.grid_ny + 31) >> 5) : 0));
  float* gl_s = tail_s;                    // [N][thst] likelihood gradient
  float* sc_s = gl_s + N * thst;           // [N][thst] score = grad_lik + grad_prior
  float* nw_s = sc_s + N * thst;           // [N][thst] updated particles
  float* ll_s = nw_s + N * thst;           // [N] log-likelihood
  float* lmix_s = ll_s + N;                // [N] log mixture weights
  float* logw_s = lmix_s + N;              // [N]
  float* logit_s = logw_s + N;             // [8 warps][N]
  if (o.tail.enabled && tid < N) {
    ll_s[tid] = (o.likelihood == DUST_LIK_EXP_UTILITY) ? (-o.alpha * m_n + logf(z_n)) - logf((float)k.S)
                                                       : -o.alpha * (c_n / (float)k.S);
  }
  if (o.grad_lik || o.tail.enabled) {
    const float f = own / z_n;
    float* crow = buf0 + tid * stride;  // the noise tiles are dead: reuse buffer 0 for the partial rows
#pragma unroll
    for (int c = 0; c < ACC; ++c)
      if (c < HA) crow[c] = acc[c] * f;
    __syncthreads();
    for (int col = tid; col < N * HA; col += NT) {
      const int n2 = col / HA, c = col - n2 * HA;
      float sacc = 0.f;
      for (int g = 0; g < G; ++g) sacc += buf0[(g * N + n2) * stride + c];
      if (o.grad_lik) o.grad_lik[inst * (long long)N * HA + col] = sacc;
      if (o.tail.enabled) gl_s[n2 * thst + c] = sacc;
    }
  }
  if (!o.tail.enabled) return;

  // ------------------------------------------------------------------------------------
  // tail: score = grad_lik + grad log GMM(theta); phi; theta += lr*phi; [weights, argmax, shift]
  // one warp per particle; lanes over the flattened dimension (HA <= 32 here)
  // ------------------------------------------------------------------------------------
  const TailParams& t = o.tail;
  const int warp = tid >> 5, lane = tid & 31, nwarps = NT >> 5;
  const float* mu_g = t.mu ? t.mu + inst * (long long)N * HA : nullptr;
  if (warp == 0) warp_log_mix(t.mix ? t.mix + inst * N : nullptr, N, lmix_s);
  __syncthreads();
  for (int n2 = warp; n2 < N; n2 += nwarps) {
    // HA <= 32 on this path: one lane slot per dimension
    float xr[1], sr[1];
    xr[0] = lane < HA ? th_s[n2 * thst + lane] : 0.f;
    if (t.aliased) warp_gmm_point<1>(xr, th_s, lmix_s, t.inv_var, N, HA, logit_s + warp * N, sr, thst);
    else warp_gmm_point<1>(xr, mu_g, lmix_s, t.inv_var, N, HA, logit_s + warp * N, sr, HA);
    if (lane < HA) sc_s[n2 * thst + lane] = gl_s[n2 * thst + lane] + sr[0];
  }
  __syncthreads();
  for (int i = warp; i < N; i += nwarps) {
    const int d = lane;  // HA <= 32
    const float xi = d < HA ? th_s[i * thst + d] : 0.f;
    float accp = 0.f;
    for (int j = 0; j < N; ++j) {
      const float xj = d < HA ? th_s[j * thst + d] : 0.f;
      const float df = xi - xj;
      const float d2 = warp_sum(df * df);
      const float kij = expf(-t.gamma * d2);
      const float sj = d < HA ? sc_s[j * thst + d] : 0.f;
      accp += kij * (t.c1 * sj + t.c2 * df);
    }
    if (d < HA) {
      const float nv = xi + t.lr * accp;
      nw_s[i * thst + d] = nv;
      const long long oidx = (inst * N + i) * (long long)HA + d;
      if (t.phi) t.phi[oidx] = accp;
      if (t.theta_out) t.theta_out[oidx] = nv;
    }
  }
  if (!t.do_forward) return;
  __syncthreads();
  // weights from the PRE-update costs and the prior evaluated at the POST-update particles
  for (int n2 = warp; n2 < N; n2 += nwarps) {
    float xr[1];
    xr[0] = lane < HA ? nw_s[n2 * thst + lane] : 0.f;
    float lp;
    if (t.aliased) lp = warp_gmm_point<1>(xr, nw_s, lmix_s, t.inv_var, N, HA, logit_s + warp * N, nullptr, thst);
    else lp = warp_gmm_point<1>(xr, mu_g, lmix_s, t.inv_var, N, HA, logit_s + warp * N, nullptr, HA);
    if (lane == 0) logw_s[n2] = ll_s[n2] + (lp + t.log_norm);
  }
  __syncthreads();
  __shared__ int s_istar;
  if (warp == 0) {
    float mx = -INFINITY;
    for (int n2 = lane; n2 < N; n2 += 32) mx = fmaxf(mx, logw_s[n2]);
    mx = warp_max(mx);
    float z = 0.f;
    for (int n2 = lane; n2 < N; n2 += 32) z += expf(logw_s[n2] - mx);
    z = warp_sum(z);
    const float lse = mx + logf(z);
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int n2 = lane; n2 < N; n2 += 32) {
      const float pw = expf(logw_s[n2] - lse);
      t.p_weights[inst * N + n2] = pw;
      if (t.mix_next) t.mix_next[inst * N + n2] = t.weighted ? pw : 1.0f;
      if (pw > best) { best = pw; bi = n2; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      s_istar = bi;
      if (t.i_star) t.i_star[inst] = bi;
    }
  }
  __syncthreads();
  const int is = s_istar;
  if (t.a_seq)
    for (int d = tid; d < HA; d += NT) t.a_seq[inst * HA + d] = nw_s[is * thst + d];
  if (t.theta_next) {
    const int shift_lim = (k.H - 1) * A;
    for (int e = tid; e < N * HA; e += NT) {
      const int n2 = e / HA, d = e - n2 * HA;
      float v;
      if (d < shift_lim) {
        v = nw_s[n2 * thst + d + A];
      } else if (t.roll == DUST_ROLL_REPEAT) {
        v = nw_s[n2 * thst + d];
      } else {
        const int a = d - shift_lim;
        float sm = 0.f;
        for (int h = 0; h < k.H; ++h) sm += nw_s[n2 * thst + h * A + a];
        v = sm / (float)k.H;
      }
      t.theta_next[inst * (long long)N * HA + e] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// The rest of the SVGD step for ONE instance, by all threads of a CTA (blockDim.x a multiple of 32), on data the CTA
// holds in shared memory: th_s [N][thst] the particles, gl_s [N][thst] the likelihood gradient, ll_s [N] the
// log-likelihood.  (svmpc.py:38-95, 128-200: prior score, phi, SGD step, weights / argmax / shift.)  Work is spread
// over (particle, centre) pairs and (particle, dimension) items -- any N <= 32 and any D = H*A that fits shared memory.
// Scratch: sc_s, nw_s [N][thst]; Lg_s, Kx_s [N][N]; lmix_s, logw_s [N].
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void svmpc_tail(const RolloutKParams& k, const TailParams& t, long long inst, int A,
                                           const float* th_s, int thst, const float* gl_s, const float* ll_s, float* sc_s,
                                           float* nw_s, float* Lg_s, float* Kx_s, float* lmix_s, float* logw_s) {
  const int N = k.N, HA = k.HA;
  const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  const float* mu_g = t.mu ? t.mu + inst * (long long)N * HA : nullptr;
  if (warp == 0) warp_log_mix(t.mix ? t.mix + inst * N : nullptr, N, lmix_s);
  __syncthreads();
  // (i, k): iv-weighted distance to centre k (GMM logits) and the kernel among the particles
  for (int p = tid; p < N * N; p += nthr) {
    const int i = p / N, kk = p - i * N;
    const float* xi = th_s + i * thst;
    const float* xk = th_s + kk * thst;
    float q = 0.f, dxx = 0.f;
    if (t.aliased) {
      for (int d = 0; d < HA; ++d) {
        const float df = xi[d] - xk[d];
        const float d2 = df * df;
        dxx += d2;
        q = fmaf(d2, __ldg(t.inv_var + d), q);
      }
    } else {
      const float* ck = mu_g + kk * HA;
      for (int d = 0; d < HA; ++d) {
        const float df = xi[d] - xk[d];
        dxx = fmaf(df, df, dxx);
        const float dc = xi[d] - __ldg(ck + d);
        q = fmaf(dc * dc, __ldg(t.inv_var + d), q);
      }
    }
    Lg_s[p] = lmix_s[kk] - 0.5f * q;
    Kx_s[p] = expf(-t.gamma * dxx);
  }
  __syncthreads();
  // responsibilities r_ik = softmax_k(logits): a warp per particle, lanes over the centres (N <= 32)
  for (int i = warp; i < N; i += nwarps) {
    const float l = lane < N ? Lg_s[i * N + lane] : -INFINITY;
    const float mx = warp_max(l);
    const float e = lane < N ? expf(l - mx) : 0.f;
    const float z = warp_sum(e);
    if (lane < N) Lg_s[i * N + lane] = e / z;
  }
  __syncthreads();
  // (i, d): prior score, total score
  for (int e = tid; e < N * HA; e += nthr) {
    const int i = e / HA, d = e - i * HA;
    const float xi = th_s[i * thst + d];
    float a = 0.f;
    if (t.aliased) {
      for (int kk = 0; kk < N; ++kk) a = fmaf(Lg_s[i * N + kk], th_s[kk * thst + d] - xi, a);
    } else {
      for (int kk = 0; kk < N; ++kk) a = fmaf(Lg_s[i * N + kk], __ldg(mu_g + kk * HA + d) - xi, a);
    }
    sc_s[i * thst + d] = gl_s[i * thst + d] + a * __ldg(t.inv_var + d);
  }
  __syncthreads();
  // (i, d): phi_i = sum_j K_ij (c1 score_j + c2 (x_i - x_j)); SGD step
  for (int e = tid; e < N * HA; e += nthr) {
    const int i = e / HA, d = e - i * HA;
    const float xi = th_s[i * thst + d];
    float accp = 0.f;
    for (int j = 0; j < N; ++j) accp = fmaf(Kx_s[i * N + j], fmaf(t.c2, xi - th_s[j * thst + d], t.c1 * sc_s[j * thst + d]), accp);
    const float nv = xi + t.lr * accp;
    nw_s[i * thst + d] = nv;
    const long long oidx = (inst * N + i) * (long long)HA + d;
    if (t.phi) t.phi[oidx] = accp;
    if (t.theta_out) t.theta_out[oidx] = nv;
  }
  if (!t.do_forward) return;
  __syncthreads();
  // weights from the PRE-update costs and the prior evaluated at the POST-update particles (quirk H19: an
  // aliased prior's centres are the updated particles themselves)
  for (int p = tid; p < N * N; p += nthr) {
    const int i = p / N, kk = p - i * N;
    const float* xi = nw_s + i * thst;
    float q = 0.f;
    if (t.aliased) {
      const float* ck = nw_s + kk * thst;
      for (int d = 0; d < HA; ++d) { const float dc = xi[d] - ck[d]; q = fmaf(dc * dc, __ldg(t.inv_var + d), q); }
    } else {
      const float* ck = mu_g + kk * HA;
      for (int d = 0; d < HA; ++d) { const float dc = xi[d] - __ldg(ck + d); q = fmaf(dc * dc, __ldg(t.inv_var + d), q); }
    }
    Lg_s[p] = lmix_s[kk] - 0.5f * q;
  }
  __syncthreads();
  for (int i = warp; i < N; i += nwarps) {
    const float l = lane < N ? Lg_s[i * N + lane] : -INFINITY;
    const float mx = warp_max(l);
    const float z = warp_sum(lane < N ? expf(l - mx) : 0.f);
    if (lane == 0) logw_s[i] = ll_s[i] + ((mx + logf(z)) + t.log_norm);
  }
  __syncthreads();
  __shared__ int s_istar;
  if (warp == 0) {
    const float lw = lane < N ? logw_s[lane] : -INFINITY;
    const float mx = warp_max(lw);
    const float z = warp_sum(lane < N ? expf(lw - mx) : 0.f);
    const float lse = mx + logf(z);
    float best = -INFINITY;
    int bi = 0x7fffffff;
    if (lane < N) {
      const float pw = expf(lw - lse);
      t.p_weights[inst * N + lane] = pw;
      if (t.mix_next) t.mix_next[inst * N + lane] = t.weighted ? pw : 1.0f;
      best = pw; bi = lane;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      s_istar = bi;
      if (t.i_star) t.i_star[inst] = bi;
    }
  }
  __syncthreads();
  const int is = s_istar;
  if (t.a_seq)
    for (int d = tid; d < HA; d += nthr) t.a_seq[inst * HA + d] = nw_s[is * thst + d];
  if (t.theta_next) {
    const int shift_lim = (k.H - 1) * A;
    for (int e = tid; e < N * HA; e += nthr) {
      const int n2 = e / HA, d = e - n2 * HA;
      float v;
      if (d < shift_lim) {
        v = nw_s[n2 * thst + d + A];
      } else if (t.roll == DUST_ROLL_REPEAT) {
        v = nw_s[n2 * thst + d];
      } else {
        const int a = d - shift_lim;
        float sm = 0.f;
        for (int h = 0; h < k.H; ++h) sm += nw_s[n2 * thst + h * A + a];
        v = sm / (float)k.H;
      }
      t.theta_next[inst * (long long)N * HA + e] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Second generation of the fused per-instance kernel (packed pendulum path; H*A a multiple of 4, N <= 32).
//
// What limited the first one (ncu, profiles/r1_fused_instance_kernel_ncu.md): 96 registers -- 20 of them the
// per-thread score-row accumulators -- held an SM at 20 warps; every 256-row tile ended in a CTA-wide barrier;
// and the tail (one warp per particle, lanes over D, a shuffle reduction per pair) cost 19 % of the
// instructions for a few thousand multiply-adds.  Here
//   * a WARP owns its tiles: 2*La consecutive rows (La = 32/N * N lanes, rows l and l + La of a lane share
//     the policy l % N), fetched by one TMA bulk copy that the warp which frees a buffer issues itself and
//     whose arrival only the consuming warp waits for (an mbarrier per buffer) -- no CTA barrier in the loop.
//     NBUF >= 4 buffers rotate over the CTA's tile sequence: tile t lives in buffer t % NBUF and is requested
//     when tile t - NBUF retires, i.e. (NBUF - 4) tile-quarters ahead of its use;
//   * the weighted score row of a lane lives in shared memory (conflict-free 16-byte rows) and is updated for
//     both trajectories of the lane at once on the packed pipe: no accumulator survives the rollout loop;
//   * the tail works on (particle, centre) pairs and (particle, dimension) items with all threads.
// Costs are produced by the same pendulum_pair_cost_sum / trajectory_cost_sum calls as before: bit-identical.
// ---------------------------------------------------------------------------------------
constexpr int kWarpKernelThreads = 128;
constexpr int kWarpKernelWarps = kWarpKernelThreads / 32;

struct WarpKernelSmem {
  int stride, thst, tile_floats, off_tile, off_acc, off_th, off_tail, off_bar, total_bytes;
};
// NP = pairs of trajectories per lane and tile (1: svmpc_warp_kernel, 2: svmpc_quad_kernel)
__host__ __device__ inline WarpKernelSmem warp_kernel_smem(int N, int HA, int NP = 1) {
  WarpKernelSmem L;
  L.stride = padded_stride(HA);
  L.thst = (HA + 3) & ~3;
  const int La = (32 / N) * N;
  L.tile_floats = 2 * NP * La * L.stride;
  // tail scratch (gl, sc, nw [N][thst]; Lg, Kx [N][N]; ll, lmix, logw, lse [N]; reductions 3 x [threads]) lives in
  // the tile buffers, which are dead once every warp has left the rollout loop
  const int tail = 3 * N * L.thst + 2 * N * N + 4 * N + 3 * kWarpKernelThreads;
  const int ring = kWarpKernelWarps * L.tile_floats;
  L.off_tile = 0;                                              // [warps][2*La][stride]: one buffer per warp
  L.off_tail = 0;
  L.off_acc = ((ring > tail ? ring : tail) + 3) & ~3;          // [threads][stride] weighted score rows
  L.off_th = L.off_acc + kWarpKernelThreads * L.stride;        // [N][thst] policy means
  L.off_bar = (L.off_th + N * L.thst + 1) & ~1;                // 8-byte aligned mbarriers
  L.total_bytes = 4 * L.off_bar + 8 * kWarpKernelWarps;
  return L;
}

#if DUST_PEND_PAIR
__device__ __forceinline__ float exp2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// acc <- acc * scale + eA * rowA + eB * rowB over one score row, both trajectories of the lane at once (packed pipe)
template <int HA4>
__device__ __forceinline__ void fold_score_rows(float* __restrict__ acc_row, const float* __restrict__ rowA,
                                                const float* __restrict__ rowB, float scale, float eA, float eB, int HA) {
  const float2 sc2 = bc2(scale), ea2 = bc2(eA), eb2 = bc2(eB);
  auto chunk = [&](int c) {
    const float4 a4 = *reinterpret_cast<const float4*>(acc_row + c);
    const float4 va = *reinterpret_cast<const float4*>(rowA + c);
    const float4 vb = *reinterpret_cast<const float4*>(rowB + c);   // a missing row: the caller passes rowA and eB = 0
    const float2 lo = fma2(make_float2(a4.x, a4.y), sc2, fma2(ea2, make_float2(va.x, va.y), mul2(eb2, make_float2(vb.x, vb.y))));
    const float2 hi = fma2(make_float2(a4.z, a4.w), sc2, fma2(ea2, make_float2(va.z, va.w), mul2(eb2, make_float2(vb.z, vb.w))));
    *reinterpret_cast<float4*>(acc_row + c) = make_float4(lo.x, lo.y, hi.x, hi.y);
  };
  if (HA4 > 0) {
#pragma unroll
    for (int c4 = 0; c4 < HA4; ++c4) chunk(4 * c4);
  } else {
    for (int c = 0; c < HA; c += 4) chunk(c);
  }
}

// What follows the rollout loop of the warp kernels: the G = 4*La/N lanes that share a policy are combined (soft-min
// statistics through shared memory, weighted score rows summed per policy), then -- dust_svmpc_step -- the tail.
__device__ __forceinline__ void warp_kernel_finish(const RolloutKParams& k, const FusedOut& o, const WarpKernelSmem& L, float* smem,
                                                   long long inst, int HA, int La, bool active, int slot, int n, float* acc_row,
                                                   float* th_s, float m_run, float z_run, float c_run, float sg0) {
  const int N = k.N, stride = L.stride, thst = L.thst, tid = threadIdx.x;
  // ---- combine the G = 4*La/N lanes that share a policy -----------------------------------------------
  float* tail_s = smem + L.off_tail;
  float* gl_s = tail_s;                    // [N][thst] likelihood gradient
  float* sc_s = gl_s + N * thst;           // [N][thst] score = grad_lik + grad_prior
  float* nw_s = sc_s + N * thst;           // [N][thst] updated particles
  float* Lg_s = nw_s + N * thst;           // [N][N] mixture logits / responsibilities
  float* Kx_s = Lg_s + N * N;              // [N][N] kernel matrix among the particles
  float* ll_s = Kx_s + N * N;              // [N] log-likelihood
  float* lmix_s = ll_s + N;                // [N] log mixture weights
  float* logw_s = lmix_s + N;              // [N]
  float* lse_s = logw_s + N;               // [N]
  float* red_m = lse_s + N;                // [threads]
  float* red_z = red_m + kWarpKernelThreads;
  float* red_c = red_z + kWarpKernelThreads;
  const int TNT = kWarpKernelWarps * La;   // lanes that own trajectories
  __syncthreads();                         // every warp is done with the tile ring: the scratch above may overwrite it
  if (active) { red_m[slot] = m_run; red_c[slot] = c_run; }
  __syncthreads();
  const int G = TNT / N;
  float m_n = INFINITY;
  for (int g = 0; g < G; ++g) m_n = fminf(m_n, red_m[g * N + n]);
  const float own = (active && m_run != INFINITY) ? expf(-o.alpha * (m_run - m_n)) : 0.f;
  if (active) red_z[slot] = z_run * own;
  __syncthreads();
  float z_n = 0.f, c_n = 0.f;
  for (int g = 0; g < G; ++g) z_n += red_z[g * N + n];
  if (tid < N) {
    for (int g = 0; g < G; ++g) c_n += red_c[g * N + n];
    const float ll = (o.likelihood == DUST_LIK_EXP_UTILITY) ? (-o.alpha * m_n + logf(z_n)) - logf((float)k.S)   // likelihoods.py:133-135
                                                            : -o.alpha * (c_n / (float)k.S);                    // likelihoods.py:119
    if (o.log_lik) o.log_lik[inst * N + tid] = ll;
    ll_s[tid] = ll;
  }
  if (o.grad_lik || o.tail.enabled) {
    // (a - theta)/sigma^2 = eps/sigma: the factor 1/sigma is applied once per row here
    const float f = (own / z_n) * ((1.0f / (sg0 * sg0)) * sg0);
    if (active)
      for (int c = 0; c < HA; c += 4) {
        float4 v = *reinterpret_cast<float4*>(acc_row + c);
        v.x *= f; v.y *= f; v.z *= f; v.w *= f;
        *reinterpret_cast<float4*>(acc_row + c) = v;
      }
    __syncthreads();
    const float* accs = smem + L.off_acc;
    for (int col = tid; col < N * HA; col += kWarpKernelThreads) {
      const int n2 = col / HA, c = col - n2 * HA;
      float sacc = 0.f;
      for (int g = 0; g < G; ++g) sacc += accs[(g * N + n2) * stride + c];
      if (o.grad_lik) o.grad_lik[inst * (long long)N * HA + col] = sacc;
      gl_s[n2 * thst + c] = sacc;
    }
  }
  if (!o.tail.enabled) return;

  svmpc_tail(k, o.tail, inst, 1, th_s, thst, gl_s, ll_s, sc_s, nw_s, Lg_s, Kx_s, lmix_s, logw_s);
}

// HA4 = H*A/4 when it is a compile-time constant (the bench shape: 5), 0 = any multiple of 4 up to 32
template <int HA4>
__global__ void __launch_bounds__(kWarpKernelThreads, 7) svmpc_warp_kernel(const RolloutKParams k, const FusedOut o) {
  extern __shared__ __align__(16) float smem[];
  const int HA = HA4 > 0 ? 4 * HA4 : k.HA, N = k.N;
  const WarpKernelSmem L = warp_kernel_smem(N, HA);
  const int stride = L.stride, thst = L.thst;
  const int La = (32 / N) * N, WT = 2 * La;
  float* th_s = smem + L.off_th;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  const long long inst = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool active = lane < La;
  const int slot = warp * La + lane;             // dense index of an active lane; slot % N == lane % N
  const int n = lane % N;
  float* acc_row = smem + L.off_acc + (active ? slot : 0) * stride;
  const int ntiles = (k.SN + WT - 1) / WT;
  const bool bulk = stride == HA;                // unpadded rows: a tile is one contiguous block
  // this warp's tile buffer and "tile landed" barrier: tile t of the instance goes to warp t % 4, which requests its
  // next tile itself as soon as it has folded the current one -- no CTA barrier, no index arithmetic in the loop
  float* const my_tile = smem + L.off_tile + warp * L.tile_floats;
  uint64_t* const my_bar = &full_bar[warp];
  const uint32_t bar_a = smem_u32(my_bar), tile_a = smem_u32(my_tile);   // shared-window addresses, converted once
  const float* __restrict__ rowA = my_tile + lane * stride;
  const float* __restrict__ rowB = my_tile + (lane + La) * stride;
  const uint32_t tile_bytes = (uint32_t)WT * (uint32_t)HA * 4u;
  // the warp's next tile in global memory: tiles warp, warp + 4, ... of the instance, one pointer bump per request
  const float* __restrict__ next_src = k.noise + (inst * (long long)k.SN + (long long)warp * WT) * HA;
  int next_j0 = warp * WT;

  auto request = [&]() {                         // the warp's next tile into its buffer; called by the whole warp
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this warp's reads of the buffer are done
    if (bulk) {
      if (lane == 0) {
        const uint32_t bytes = (next_j0 + WT <= k.SN) ? tile_bytes : (uint32_t)(k.SN - next_j0) * (uint32_t)HA * 4u;
        mbar_expect_tx_a(bar_a, bytes);
        bulk_g2s_a(tile_a, next_src, bytes, bar_a);
      }
    } else {
      const int rows = min(WT, k.SN - next_j0);
      if (lane == 0) mbar_expect_tx_a(bar_a, (uint32_t)rows * (uint32_t)HA * 4u);
      __syncwarp();
      for (int r = lane; r < rows; r += 32) bulk_g2s_a(tile_a + (uint32_t)(r * stride) * 4u, next_src + (long long)r * HA, (uint32_t)HA * 4u, bar_a);
    }
    next_src += (long long)kWarpKernelWarps * WT * HA;
    next_j0 += kWarpKernelWarps * WT;
  };
  if (lane == 0) {
    mbar_init(my_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp < ntiles) request();
  for (int e = tid; e < N * HA; e += kWarpKernelThreads) {
    const int nn = e / HA;
    th_s[nn * thst + (e - nn * HA)] = k.theta[inst * (long long)N * HA + e];
  }
  if (active)
    for (int c = 0; c < stride; c += 4) *reinterpret_cast<float4*>(acc_row + c) = make_float4(0.f, 0.f, 0.f, 0.f);
  const float sg0 = k.sigma[0];
  const float th0 = __ldg(k.state0 + inst * 2), om0 = __ldg(k.state0 + inst * 2 + 1);
  const bool small = small_angle_horizon<DUST_MODEL_PENDULUM>(k, inst);
  const float* __restrict__ th_row = th_s + n * thst;
  const bool want_costs = o.costs != nullptr;     // uniform: the pointer below is only dereferenced when set
  uintptr_t cost_ptr = reinterpret_cast<uintptr_t>(o.costs) + sizeof(float) * (size_t)(inst * k.SN + warp * WT + lane);
  __syncthreads();                               // policy means visible to every warp

  float m_run = INFINITY, z_run = 0.f, c_run = 0.f;
  uint32_t parity = 0;
  const float nal2 = -o.alpha * 1.4426950408889634f;
  for (int t = warp; t < ntiles; t += kWarpKernelWarps) {
    const int j0 = t * WT;
    const int rows = min(WT, k.SN - j0);
    mbar_wait_a(bar_a, parity);
    parity ^= 1u;
    if (active && lane < rows) {
      const bool hasB = lane + La < rows;
      float csA, csB = 0.f;
      if (small && hasB) {
        const float2 cs = pendulum_pair_cost_sum(k, rowA, rowB, inst, j0 + lane, j0 + lane + La, th_row, sg0, th0, om0);
        csA = cs.x;
        csB = cs.y;
      } else {
        // ragged last tile, or an angle beyond the fast range: the scalar step, same arithmetic (one copy of the code)
        csA = 0.f;
#pragma unroll 1
        for (int q = 0; q < (hasB ? 2 : 1); ++q) {
          const float* __restrict__ row = q ? rowB : rowA;
          const int jj = j0 + lane + (q ? La : 0);
          const float v = small ? trajectory_cost_sum<DUST_MODEL_PENDULUM, true, false, true>(k, row, nullptr, inst, jj, 0, k.P, th_row, sg0, sg0)
                                : trajectory_cost_sum<DUST_MODEL_PENDULUM, false, false, true>(k, row, nullptr, inst, jj, 0, k.P, th_row, sg0, sg0);
          if (q) csB = v; else csA = v;
        }
      }
      const float costA = (k.P == 1) ? csA : csA / (float)k.P;
      const float costB = hasB ? ((k.P == 1) ? csB : csB / (float)k.P) : INFINITY;
      if (want_costs) {
        float* cp = reinterpret_cast<float*>(cost_ptr);
        cp[0] = costA;
        if (hasB) cp[La] = costB;
      }
      c_run += costA;
      if (hasB) c_run += costB;
      // online soft-min of this lane's rows relative to its running minimum: both new rows are folded at once
      const float m_new = fminf(m_run, fminf(costA, costB));
      // exp(-alpha d) = 2^(-alpha log2(e) d) on the MUFU unit (2 ulp; the argument's rounding adds |d| 6e-8 relative to a
      // weight of e^-|d|): two instructions per weight instead of libm's nine
      const float scale = exp2_fast(nal2 * (m_run - m_new));         // 0 on the first tile (m_run = inf), 1 if the minimum stands
      const float eA = exp2_fast(nal2 * (costA - m_new));
      const float eB = exp2_fast(nal2 * (costB - m_new));            // 0 for a missing row (cost = inf)
      m_run = m_new;
      z_run = z_run * scale + (eA + eB);
      if (scale != 1.f || eA > 1e-30f || eB > 1e-30f)                // else: invisible in float32
        fold_score_rows<HA4>(acc_row, rowA, hasB ? rowB : rowA, scale, eA, eB, HA);
    }
    cost_ptr += sizeof(float) * (size_t)(kWarpKernelWarps * WT);
    __syncwarp();
    if (t + kWarpKernelWarps < ntiles) request();
  }

  warp_kernel_finish(k, o, L, smem, inst, HA, La, active, slot, n, acc_row, th_s, m_run, z_run, c_run, sg0);
}

// acc <- acc * scale + sum_r e[r] * row[r] over one score row: the four trajectories of a lane at once (packed pipe)
template <int HA4>
__device__ __forceinline__ void fold_score_rows4(float* __restrict__ acc_row, const float* __restrict__ const (&row)[4], float scale,
                                                 const float (&e)[4], int HA) {
  const float2 sc2 = bc2(scale), e0 = bc2(e[0]), e1 = bc2(e[1]), e2 = bc2(e[2]), e3 = bc2(e[3]);
  auto chunk = [&](int c) {
    const float4 a4 = *reinterpret_cast<const float4*>(acc_row + c);
    const float4 v0 = *reinterpret_cast<const float4*>(row[0] + c), v1 = *reinterpret_cast<const float4*>(row[1] + c);
    const float4 v2 = *reinterpret_cast<const float4*>(row[2] + c), v3 = *reinterpret_cast<const float4*>(row[3] + c);
    float2 lo = fma2(e0, make_float2(v0.x, v0.y), mul2(e1, make_float2(v1.x, v1.y)));
    float2 hi = fma2(e0, make_float2(v0.z, v0.w), mul2(e1, make_float2(v1.z, v1.w)));
    lo = fma2(e2, make_float2(v2.x, v2.y), lo);
    hi = fma2(e2, make_float2(v2.z, v2.w), hi);
    lo = fma2(e3, make_float2(v3.x, v3.y), lo);
    hi = fma2(e3, make_float2(v3.z, v3.w), hi);
    lo = fma2(make_float2(a4.x, a4.y), sc2, lo);
    hi = fma2(make_float2(a4.z, a4.w), sc2, hi);
    *reinterpret_cast<float4*>(acc_row + c) = make_float4(lo.x, lo.y, hi.x, hi.y);
  };
  if (HA4 > 0) {
#pragma unroll
    for (int c4 = 0; c4 < HA4; ++c4) chunk(4 * c4);
  } else {
    for (int c = 0; c < HA; c += 4) chunk(c);
  }
}

// Two pairs per lane: tiles of 4 * La rows, 16 warps per SM with two independent packed chains each (the per-tile
// overhead -- barrier wait, request, exponentials, fold -- is shared by four trajectories).  Host guarantees
// S*N % (4 * La) == 0 (no ragged tile).  An experiment that lost (see the launch site); opt-in with DUST_B200_QUAD=1.
template <int HA4>
__global__ void __launch_bounds__(kWarpKernelThreads, 4) svmpc_quad_kernel(const RolloutKParams k, const FusedOut o) {
  extern __shared__ __align__(16) float smem[];
  const int HA = HA4 > 0 ? 4 * HA4 : k.HA, N = k.N;
  const WarpKernelSmem L = warp_kernel_smem(N, HA, 2);
  const int stride = L.stride, thst = L.thst;
  const int La = (32 / N) * N, WT = 4 * La;
  float* th_s = smem + L.off_th;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  const long long inst = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool active = lane < La;
  const int slot = warp * La + lane;
  const int n = lane % N;
  float* acc_row = smem + L.off_acc + (active ? slot : 0) * stride;
  const int ntiles = k.SN / WT;
  float* const my_tile = smem + L.off_tile + warp * L.tile_floats;
  uint64_t* const my_bar = &full_bar[warp];
  const uint32_t bar_a = smem_u32(my_bar), tile_a = smem_u32(my_tile);
  const float* row[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) row[r] = my_tile + (lane + r * La) * stride;
  const uint32_t tile_bytes = (uint32_t)WT * (uint32_t)HA * 4u;
  const bool bulk = stride == HA;
  const float* __restrict__ next_src = k.noise + (inst * (long long)k.SN + (long long)warp * WT) * HA;

  auto request = [&]() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (bulk) {
      if (lane == 0) {
        mbar_expect_tx_a(bar_a, tile_bytes);
        bulk_g2s_a(tile_a, next_src, tile_bytes, bar_a);
      }
    } else {
      if (lane == 0) mbar_expect_tx_a(bar_a, tile_bytes);
      __syncwarp();
      for (int r = lane; r < WT; r += 32) bulk_g2s_a(tile_a + (uint32_t)(r * stride) * 4u, next_src + (long long)r * HA, (uint32_t)HA * 4u, bar_a);
    }
    next_src += (long long)kWarpKernelWarps * WT * HA;
  };
  if (lane == 0) {
    mbar_init(my_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (warp < ntiles) request();
  for (int e = tid; e < N * HA; e += kWarpKernelThreads) {
    const int nn = e / HA;
    th_s[nn * thst + (e - nn * HA)] = k.theta[inst * (long long)N * HA + e];
  }
  if (active)
    for (int c = 0; c < stride; c += 4) *reinterpret_cast<float4*>(acc_row + c) = make_float4(0.f, 0.f, 0.f, 0.f);
  const float sg0 = k.sigma[0];
  const float th0 = __ldg(k.state0 + inst * 2), om0 = __ldg(k.state0 + inst * 2 + 1);
  const bool small = small_angle_horizon<DUST_MODEL_PENDULUM>(k, inst);
  const float* __restrict__ th_row = th_s + n * thst;
  const bool want_costs = o.costs != nullptr;
  uintptr_t cost_ptr = reinterpret_cast<uintptr_t>(o.costs) + sizeof(float) * (size_t)(inst * k.SN + warp * WT + lane);
  __syncthreads();

  float m_run = INFINITY, z_run = 0.f, c_run = 0.f;
  uint32_t parity = 0;
  const float nal2 = -o.alpha * 1.4426950408889634f;
  for (int t = warp; t < ntiles; t += kWarpKernelWarps) {
    const int j0 = t * WT;
    mbar_wait_a(bar_a, parity);
    parity ^= 1u;
    if (active) {
      float cost[4];
      if (small) {
        const int j[4] = {j0 + lane, j0 + lane + La, j0 + lane + 2 * La, j0 + lane + 3 * La};
        float2 cs[2];
        pendulum_quad_cost_sum(k, row, inst, j, th_row, sg0, th0, om0, cs);
        cost[0] = cs[0].x; cost[1] = cs[0].y; cost[2] = cs[1].x; cost[3] = cs[1].y;
      } else {
        // an angle beyond the fast range: the scalar step, one copy of the code (no dynamic indexing of row[] / cost[]:
        // the arrays must stay in registers)
        cost[0] = cost[1] = cost[2] = cost[3] = 0.f;
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
          const float v = trajectory_cost_sum<DUST_MODEL_PENDULUM, false, false, true>(k, my_tile + (lane + r * La) * stride, nullptr, inst,
                                                                                      j0 + lane + r * La, 0, k.P, th_row, sg0, sg0);
          if (r == 0) cost[0] = v; else if (r == 1) cost[1] = v; else if (r == 2) cost[2] = v; else cost[3] = v;
        }
      }
      if (k.P != 1) {
#pragma unroll
        for (int r = 0; r < 4; ++r) cost[r] = cost[r] / (float)k.P;
      }
      if (want_costs) {
        float* cp = reinterpret_cast<float*>(cost_ptr);
#pragma unroll
        for (int r = 0; r < 4; ++r) cp[r * La] = cost[r];
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) c_run += cost[r];
      const float m_new = fminf(fminf(m_run, fminf(cost[0], cost[1])), fminf(cost[2], cost[3]));
      const float scale = exp2_fast(nal2 * (m_run - m_new));
      float e[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) e[r] = exp2_fast(nal2 * (cost[r] - m_new));
      m_run = m_new;
      z_run = z_run * scale + ((e[0] + e[1]) + (e[2] + e[3]));
      if (scale != 1.f || fmaxf(fmaxf(e[0], e[1]), fmaxf(e[2], e[3])) > 1e-30f) fold_score_rows4<HA4>(acc_row, row, scale, e, HA);
    }
    cost_ptr += sizeof(float) * (size_t)(kWarpKernelWarps * WT);
    __syncwarp();
    if (t + kWarpKernelWarps < ntiles) request();
  }
  warp_kernel_finish(k, o, L, smem, inst, HA, La, active, slot, n, acc_row, th_s, m_run, z_run, c_run, sg0);
}
#endif

// ---------------------------------------------------------------------------------------
// Few instances (the demo shapes: ONE instance, 384 trajectories x 4-8 parameter draws): the whole control step in one
// launch of a thread-block CLUSTER per instance.  The instance's trajectories are cut into C = 8 row blocks, one per
// CTA; inside a CTA a thread rolls out one (row, parameter draw) pair, so the 3072 / 1536 rollouts of the two demo
// configurations spread over 8 SMs instead of queueing in one CTA.  Two exchanges through distributed shared memory:
//   1. every CTA writes the finished costs of its rows into the cost vector of ALL CTAs (cluster.sync), so each can
//      form the per-policy soft-min statistics on its own;
//   2. every CTA sends its partial weighted column sums (the analytic likelihood gradient, svmpc.py:46-54) to rank 0
//      (cluster.sync), which adds them in rank order and runs the tail (svmpc_tail) for the instance.
// Costs per (row, draw) come from trajectory_cost_sum and are averaged in draw order: bit-identical to the fused
// per-instance kernels.  Reference: disco.py:139-209, 294-346, 380-393; likelihoods.py:81-135; svmpc.py:32-95, 128-200.
// ---------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int kClusterSize = 8;
constexpr int kClusterMaxThreads = 512;

struct ClusterLayout {
  int rows, PT, threads, stride, thst;
  int off_tile, off_th, off_grid, off_cpart, off_costs, off_w, off_stat, off_recv, off_tail, total_floats;
};
static ClusterLayout cluster_layout(int SN, int N, int HA, int P, int grid_words) {
  ClusterLayout L;
  L.rows = (SN + kClusterSize - 1) / kClusterSize;
  L.PT = P < kClusterMaxThreads / L.rows ? P : kClusterMaxThreads / L.rows;
  if (L.PT < 1) L.PT = 1;
  L.threads = (L.rows * L.PT + 31) & ~31;
  if (L.threads < 128) L.threads = 128;
  L.stride = padded_stride(HA);
  L.thst = (HA + 3) & ~3;
  int off = 0;
  auto take = [&](int n) { const int o = off; off += (n + 3) & ~3; return o; };
  L.off_tile = take(L.rows * L.stride);
  L.off_th = take(N * L.thst);
  L.off_grid = take(grid_words);
  L.off_cpart = take(P * L.rows);
  L.off_costs = take(SN);
  L.off_w = take(L.rows);
  L.off_stat = take(4 * N);                       // cmin, za, csum, ll per policy
  L.off_recv = take(kClusterSize * N * HA);       // rank 0: the partial column sums of every CTA
  L.off_tail = take(3 * N * L.thst + 2 * N * N + 2 * N);
  L.total_floats = off;
  return L;
}

template <int MODEL>
__global__ void __launch_bounds__(kClusterMaxThreads) svmpc_cluster_kernel(const RolloutKParams k, const FusedOut o, const ClusterLayout L) {
  constexpr int A = (MODEL == DUST_MODEL_PENDULUM) ? 1 : 2;
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const long long inst = blockIdx.x / kClusterSize;
  const int N = k.N, HA = k.HA, SN = k.SN, P = k.P;
  const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  const int stride = L.stride, thst = L.thst;
  float* tile = smem + L.off_tile;
  float* th_s = smem + L.off_th;
  uint32_t* grid_s = reinterpret_cast<uint32_t*>(smem + L.off_grid);
  float* cpart = smem + L.off_cpart;
  float* costs_all = smem + L.off_costs;
  float* w_s = smem + L.off_w;
  float* cmin_s = smem + L.off_stat;
  float* za_s = cmin_s + N;
  float* ll_s = za_s + 2 * N;
  float* recv = smem + L.off_recv;
  const int r0 = rank * L.rows;
  const int rows = max(0, min(L.rows, SN - r0));

  // ---- stage this CTA's noise rows (padded), the policy means, the occupancy bits ----
  const float* __restrict__ src = k.noise + (inst * SN + r0) * (long long)HA;
  if ((HA & 3) == 0 && ((((uintptr_t)k.noise) & 15) == 0)) {
    const int HA4 = HA >> 2;
    for (int e = tid; e < rows * HA4; e += nthr) {
      const int r = e / HA4, c4 = e - r * HA4;
      *reinterpret_cast<float4*>(tile + r * stride + 4 * c4) = __ldg(reinterpret_cast<const float4*>(src) + e);
    }
  } else {
    for (int e = tid; e < rows * HA; e += nthr) {
      const int r = e / HA, c = e - r * HA;
      tile[r * stride + c] = __ldg(src + e);
    }
  }
  for (int e = tid; e < N * HA; e += nthr) {
    const int nn = e / HA;
    th_s[nn * thst + (e - nn * HA)] = k.theta[inst * (long long)N * HA + e];
  }
  if (MODEL == DUST_MODEL_PARTICLE && k.m.grid_bits != nullptr) {
    const int words = (k.m.grid_nx * k.m.grid_ny + 31) >> 5;
    for (int w = tid; w < words; w += nthr) grid_s[w] = __ldg(k.m.grid_bits + w);
  }
  // every CTA of the cluster is running (its shared memory exists) before anyone writes into a peer; also the CTA barrier
  cluster.sync();

  // ---- rollouts: thread <-> (row, draw group); one trajectory per (row, draw) ----
  const float sg0 = k.sigma[0], sg1 = k.sigma[A - 1];
  {
    const int row = tid % L.rows, pg = tid / L.rows;
    if (row < rows && pg < L.PT) {
      const int j = r0 + row;
      const float* __restrict__ erow = tile + row * stride;
      const float* __restrict__ th_row = th_s + (j % N) * thst;
      const bool small = small_angle_horizon<MODEL>(k, inst);
      for (int p = pg; p < P; p += L.PT) {
        cpart[p * L.rows + row] =
            small ? trajectory_cost_sum<MODEL, true, false, true>(k, erow, grid_s, inst, j, p, p + 1, th_row, sg0, sg1)
                  : trajectory_cost_sum<MODEL, false, false, true>(k, erow, grid_s, inst, j, p, p + 1, th_row, sg0, sg1);
      }
    }
  }
  __syncthreads();
  // mean over the draws in draw order (disco.py:330), then into the cost vector of every CTA of the cluster
  if (tid < rows) {
    float cs = 0.f;
    for (int p = 0; p < P; ++p) cs = cs + cpart[p * L.rows + tid];
    const float cost = (P == 1) ? cs : cs / (float)P;
    if (o.costs) o.costs[inst * SN + r0 + tid] = cost;
    for (int c = 0; c < kClusterSize; ++c) cluster.map_shared_rank(costs_all, c)[r0 + tid] = cost;
  }
  cluster.sync();

  // ---- per-policy soft-min statistics (every CTA, same arithmetic as policy_softmin_kernel) ----
  for (int n = warp; n < N; n += nwarps) {
    float cmin = INFINITY, csum = 0.f;
    for (int s2 = lane; s2 < k.S; s2 += 32) {
      const float c = costs_all[s2 * N + n];
      cmin = fminf(cmin, c);
      csum += c;
    }
    cmin = warp_min(cmin);
    csum = warp_sum(csum);
    float za = 0.f;
    for (int s2 = lane; s2 < k.S; s2 += 32) za += expf(-o.alpha * (costs_all[s2 * N + n] - cmin));
    za = warp_sum(za);
    if (lane == 0) {
      cmin_s[n] = cmin;
      za_s[n] = za;
      const float ll = (o.likelihood == DUST_LIK_EXP_UTILITY) ? (-o.alpha * cmin + logf(za)) - logf((float)k.S)   // likelihoods.py:133-135
                                                              : -o.alpha * (csum / (float)k.S);                   // likelihoods.py:119
      ll_s[n] = ll;
      if (rank == 0 && o.log_lik) o.log_lik[inst * N + n] = ll;
    }
  }
  __syncthreads();
  if (o.grad_lik || o.tail.enabled) {
    if (tid < rows) {
      const int n = (r0 + tid) % N;
      w_s[tid] = expf(-o.alpha * (costs_all[r0 + tid] - cmin_s[n])) / za_s[n];
    }
    __syncthreads();
    // partial weighted column sums over this CTA's rows: (a - theta)/sigma^2 = eps/sigma
    float* dst = cluster.map_shared_rank(recv, 0) + rank * N * HA;
    const int first_n = r0 % N;
    for (int col = tid; col < N * HA; col += nthr) {
      const int n = col / HA, c = col - n * HA;
      const float sg = (c % A) ? sg1 : sg0;
      const float f = (1.0f / (sg * sg)) * sg;
      float acc = 0.f;
      int row = n - first_n;
      if (row < 0) row += N;
      for (; row < rows; row += N) acc = fmaf(w_s[row] * f, tile[row * stride + c], acc);
      dst[col] = acc;
    }
  }
  cluster.sync();
  if (rank != 0) return;
  float* tail_s = smem + L.off_tail;
  float* gl_s = tail_s;
  float* sc_s = gl_s + N * thst;
  float* nw_s = sc_s + N * thst;
  float* Lg_s = nw_s + N * thst;
  float* Kx_s = Lg_s + N * N;
  float* lmix_s = Kx_s + N * N;
  float* logw_s = lmix_s + N;
  if (o.grad_lik || o.tail.enabled) {
    for (int col = tid; col < N * HA; col += nthr) {
      float g = 0.f;
      for (int c = 0; c < kClusterSize; ++c) g += recv[c * N * HA + col];
      const int n = col / HA;
      if (o.grad_lik) o.grad_lik[inst * (long long)N * HA + col] = g;
      gl_s[n * thst + (col - n * HA)] = g;
    }
  }
  if (!o.tail.enabled) return;
  svmpc_tail(k, o.tail, inst, A, th_s, thst, gl_s, ll_s, sc_s, nw_s, Lg_s, Kx_s, lmix_s, logw_s);
}

// ---------------------------------------------------------------------------------------
// per-policy statistics: combine parameter chunks, log-likelihood, soft-min weights, mixture
// one CTA per instance; warp w handles policies w, w+nwarps, ...
// ---------------------------------------------------------------------------------------
struct SoftminKParams {
  int B, N, S, P, PC, SN;
  int likelihood;
  float alpha, inv_temp;
  const float* cost_part;  // [B,PC,SN] when PC>1 (else costs already final)
  float* costs;            // [B,S,N]
  float* log_lik;          // [B,N] or null
  float* lik_w;            // [B,S,N] or null
  float* mppi_w;           // [B,S,N] or null
  float* mix;              // [B,N] or null
  float* eta;              // [B,N] scratch for a_mix when spread (may be null if mix is not requested)
  int spread;              // the policies of an instance are spread over grid.y; chunk sums were combined before
};

constexpr int kSoftminThreads = 256;

// few instances: the chunk sums are combined by a grid over the trajectories (same order of additions as
// the one-CTA loop below), the policies are spread over grid.y, and a_mix is formed from the etas afterwards
__global__ void __launch_bounds__(256) combine_cost_chunks_kernel(const SoftminKParams k) {
  const long long inst = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k.SN) return;
  const float* part = k.cost_part + inst * (long long)k.PC * k.SN;
  float acc = 0.f;
  for (int c = 0; c < k.PC; ++c) acc = acc + part[(long long)c * k.SN + j];
  k.costs[inst * k.SN + j] = acc / (float)k.P;
}

__global__ void __launch_bounds__(32) policy_mix_kernel(const float* __restrict__ eta, float* __restrict__ mix, int N) {
  const long long inst = blockIdx.x;
  const int lane = threadIdx.x;
  const float* e = eta + inst * N;
  float mx = -INFINITY;
  for (int n = lane; n < N; n += 32) mx = fmaxf(mx, e[n]);
  mx = warp_max(mx);
  float z = 0.f;
  for (int n = lane; n < N; n += 32) z += expf(e[n] - mx);
  z = warp_sum(z);
  for (int n = lane; n < N; n += 32) mix[inst * N + n] = expf(e[n] - mx) / z;
}

__global__ void __launch_bounds__(kSoftminThreads) policy_softmin_kernel(const SoftminKParams k) {
  extern __shared__ float sm_eta[];  // [N] eta_n
  const long long inst = blockIdx.x;
  float* costs = k.costs + inst * k.SN;
  if (k.PC > 1 && !k.spread) {
    const float* part = k.cost_part + inst * (long long)k.PC * k.SN;
    for (int j = threadIdx.x; j < k.SN; j += blockDim.x) {
      float acc = 0.f;
      for (int c = 0; c < k.PC; ++c) acc = acc + part[(long long)c * k.SN + j];
      costs[j] = acc / (float)k.P;
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const bool need_mppi = (k.mppi_w != nullptr) || (k.mix != nullptr);
  for (int n = blockIdx.y * nwarps + warp; n < k.N; n += nwarps * gridDim.y) {
    float cmin = INFINITY, csum = 0.f;
    for (int s = lane; s < k.S; s += 32) {
      const float c = costs[s * k.N + n];
      cmin = fminf(cmin, c);
      csum += c;
    }
    cmin = warp_min(cmin);
    csum = warp_sum(csum);
    float za = 0.f, zt = 0.f;
    for (int s = lane; s < k.S; s += 32) {
      const float c = costs[s * k.N + n];
      za += expf(-k.alpha * (c - cmin));
      if (need_mppi) zt += expf(-(c - cmin) * k.inv_temp);
    }
    za = warp_sum(za);
    zt = warp_sum(zt);
    if (k.log_lik && lane == 0) {
      float ll;
      if (k.likelihood == DUST_LIK_EXP_UTILITY)
        ll = (-k.alpha * cmin + logf(za)) - logf((float)k.S);  // likelihoods.py:133-135
      else
        ll = -k.alpha * (csum / (float)k.S);                   // likelihoods.py:119
      k.log_lik[inst * k.N + n] = ll;
    }
    const float inv_za = 1.f / za, inv_zt = need_mppi ? 1.f / zt : 0.f;
    for (int s = lane; s < k.S; s += 32) {
      const float c = costs[s * k.N + n];
      if (k.lik_w) k.lik_w[inst * k.SN + s * k.N + n] = expf(-k.alpha * (c - cmin)) * inv_za;
      if (k.mppi_w) k.mppi_w[inst * k.SN + s * k.N + n] = expf(-(c - cmin) * k.inv_temp) * inv_zt;
    }
    if (need_mppi && lane == 0) {
      const float eta = -cmin * k.inv_temp + logf(zt);  // eta_n + beta/temp
      if (!k.spread) sm_eta[n] = eta;
      else if (k.eta) k.eta[inst * k.N + n] = eta;
    }
  }
  if (k.mix && !k.spread) {
    __syncthreads();
    // a_mix = softmax_n(eta) (disco.py:393); the global shift beta cancels
    if (warp == 0) {
      float mx = -INFINITY;
      for (int n = lane; n < k.N; n += 32) mx = fmaxf(mx, sm_eta[n]);
      mx = warp_max(mx);
      float z = 0.f;
      for (int n = lane; n < k.N; n += 32) z += expf(sm_eta[n] - mx);
      z = warp_sum(z);
      for (int n = lane; n < k.N; n += 32) k.mix[inst * k.N + n] = expf(sm_eta[n] - mx) / z;
    }
  }
}

// ---------------------------------------------------------------------------------------
// weighted column sums over the [S, N*H*A] noise block of one instance:
//   grad_lik[n,c]   = sum_s lik_w[s,n]  * ((theta + sigma*eps) - theta) / sigma^2     (svmpc.py:52-54)
//   mppi_delta[n,c] = sum_s mppi_w[s,n] * pert[s,n,c]                                 (disco.py:387-392)
// grid = (column chunks, B); thread <-> one column (n, c), rows strided over row groups.
// ---------------------------------------------------------------------------------------
struct ColsumKParams {
  int B, N, S, HA, A, W;  // W = N*HA
  const float* theta;     // [B,N,HA] or null
  const float* noise;     // [B,S,W]
  const float* sigma;     // [A]
  const float* a_seq;     // [B,HA] or null
  const float* pert;      // [B,S,W] or null
  const float* lik_w;     // [B,S,N]
  const float* mppi_w;    // [B,S,N]
  float* grad_lik;        // [B,W] or null
  float* mppi_delta;      // [B,W] or null
};

constexpr int kColsumThreads = 256;

__global__ void __launch_bounds__(kColsumThreads) weighted_colsum_kernel(const ColsumKParams k) {
  __shared__ float red_g[kColsumThreads], red_d[kColsumThreads];
  const long long inst = blockIdx.y;
  const int W = k.W;
  // columns handled by this CTA: [col0, col0 + cols)
  const int cols_per_cta = min(W, kColsumThreads);
  const int col0 = blockIdx.x * cols_per_cta;
  const int cols = min(cols_per_cta, W - col0);
  const int groups = kColsumThreads / cols_per_cta;  // row groups when W < threads
  const int g = threadIdx.x / cols_per_cta;
  const int cl = threadIdx.x - g * cols_per_cta;
  float accg = 0.f, accd = 0.f;
  const bool active = (g < groups) && (cl < cols);
  if (active) {
    const int col = col0 + cl;
    const int n = col / k.HA, c = col - n * k.HA;
    const float sg = k.sigma ? k.sigma[c % k.A] : 1.f;
    const float sg2 = sg * sg;
    const float th = k.theta ? k.theta[inst * W + col] : 0.f;
    const float aseq = k.a_seq ? k.a_seq[inst * k.HA + c] : 0.f;
    const float* __restrict__ nz = k.noise + inst * (long long)k.S * W + col;
    const float* __restrict__ pt = k.pert ? k.pert + inst * (long long)k.S * W + col : nullptr;
    const float* __restrict__ lw = k.lik_w ? k.lik_w + inst * (long long)k.S * k.N + n : nullptr;
    const float* __restrict__ mw = k.mppi_w ? k.mppi_w + inst * (long long)k.S * k.N + n : nullptr;
    for (int s = g; s < k.S; s += groups) {
      const float e = __ldg(nz + (long long)s * W);
      const float a = k.theta ? (th + sg * e) : e;
      if (k.grad_lik) accg += __ldg(lw + s * k.N) * ((a - th) / sg2);
      if (k.mppi_delta) {
        const float pv = pt ? __ldg(pt + (long long)s * W) : (a - aseq);
        accd += __ldg(mw + s * k.N) * pv;
      }
    }
  }
  if (groups > 1) {
    red_g[threadIdx.x] = accg;
    red_d[threadIdx.x] = accd;
    __syncthreads();
    if (g == 0 && cl < cols) {
      for (int q = 1; q < groups; ++q) {
        accg += red_g[q * cols_per_cta + cl];
        accd += red_d[q * cols_per_cta + cl];
      }
    }
  }
  if (g == 0 && cl < cols) {
    if (k.grad_lik) k.grad_lik[inst * W + col0 + cl] = accg;
    if (k.mppi_delta) k.mppi_delta[inst * W + col0 + cl] = accd;
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static int choose_param_chunks(long long BSN, int P) {
  const long long target_threads = (long long)kNumSMs * 2048 * 2;
  if (P <= 1 || BSN >= target_threads) return 1;
  long long pc = (target_threads + BSN - 1) / BSN;
  if (pc > P) pc = P;
  return (int)pc;
}

struct RolloutPlan {
  int PC, Pchunk, NSUB, parts;
  size_t off_part, off_likw, off_mppiw, off_costs, off_eta, total;
};

static RolloutPlan plan_rollout(const dust_rollout_args* a, bool single_chunk = false) {
  RolloutPlan pl{};
  const long long SN = (long long)a->S * a->N;
  const int P = a->p_end > a->p_begin ? a->p_end - a->p_begin : (a->params ? a->P : 1);   // draws this call rolls out
  if (a->sigma_weights || a->ctrl_mat) single_chunk = true;   // weighted / regularised costs are formed in one thread
  int pc = single_chunk ? 1 : choose_param_chunks((long long)a->B * SN, P);
  const int chunk = (P + pc - 1) / pc;
  pc = (P + chunk - 1) / chunk;
  pl.PC = pc;
  pl.Pchunk = chunk;
  // a wide action tile leaves room for only a few 128-thread CTAs per SM: let up to 4 thread groups share one
  pl.NSUB = 1;
  if (a->model != nullptr) {
    const int A = model_da(a->model->kind);
    const size_t tile_bytes = sizeof(float) * kTile * padded_stride(a->H * A);
    if (!single_chunk && !a->states && tile_bytes * 6 > 227 * 1024) pl.NSUB = chunk >= 8 ? 4 : (chunk >= 4 ? 2 : 1);
  }
  pl.parts = pl.PC * pl.NSUB;
  size_t off = 0;
  auto take = [&](bool needed, size_t bytes) {
    const size_t o = off;
    if (needed) off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t per_traj = sizeof(float) * (size_t)a->B * (size_t)SN;
  pl.off_part = take(pl.parts > 1, per_traj * pl.parts);
  pl.off_likw = take(a->grad_lik && !a->lik_weights, per_traj);
  pl.off_mppiw = take(a->mppi_delta && !a->mppi_weights, per_traj);
  pl.off_costs = take(!a->costs, per_traj);
  pl.off_eta = take(a->mix != nullptr, sizeof(float) * (size_t)a->B * a->N);
  pl.total = off;
  return pl;
}

// the fused per-instance kernel takes a call when it only asks for what that kernel produces
static bool fused_path_ok(const dust_rollout_args* a, const RolloutPlan& pl, bool tail, bool reduce_only) {
  const bool ranged = a->p_end > a->p_begin;
  const bool fused_outputs_only = !a->lik_weights && !a->mppi_weights && !a->mppi_delta && !a->mix && !a->states;
  const int HA = a->H * model_da(a->model->kind);
  return fused_outputs_only && !ranged && !reduce_only && !a->sigma_weights && !a->ctrl_mat && a->theta && pl.parts == 1 &&
         HA <= 32 && a->N <= kFusedThreads && (long long)a->B * 2 >= kNumSMs && (a->log_lik || a->grad_lik || tail);
}

// few instances: the cluster kernel (one thread-block cluster per instance) produces the same outputs in one launch
static bool cluster_path_ok(const dust_rollout_args* a, bool reduce_only, ClusterLayout* out) {
  static const bool off = getenv("DUST_B200_NO_CLUSTER") != nullptr;
  if (off || reduce_only || a->p_end > a->p_begin) return false;
  if (a->lik_weights || a->mppi_weights || a->mppi_delta || a->mix || a->states || a->sigma_weights || a->ctrl_mat) return false;
  if (!a->theta || (long long)a->B * 2 >= kNumSMs || a->N > 32) return false;
  const int kind = a->model->kind, HA = a->H * model_da(kind), P = a->params ? a->P : 1;
  const long long SN = (long long)a->S * a->N;
  if (P > 64 || SN > (long long)kClusterSize * kClusterMaxThreads || HA > 256) return false;
  const int words = (kind == DUST_MODEL_PARTICLE && a->model->grid_bits) ? (a->model->grid_nx * a->model->grid_ny + 31) / 32 : 0;
  const ClusterLayout L = cluster_layout((int)SN, a->N, HA, P, words);
  if ((size_t)L.total_floats * 4 > 200 * 1024) return false;
  if (out) *out = L;
  return true;
}

}  // namespace dust

using namespace dust;

extern "C" int dust_rollout_plan(const dust_rollout_args* a, int32_t plan[5]) {
  DUST_REQUIRE(a != nullptr && plan != nullptr, DUST_ERR_INVALID_ARG, "dust_rollout_plan: NULL argument");
  int rc = validate_model(a->model);
  if (rc) return rc;
  const RolloutPlan pl = plan_rollout(a);
  plan[0] = cluster_path_ok(a, false, nullptr) && (a->log_lik || a->grad_lik) ? 2 : (fused_path_ok(a, pl, false, false) ? 1 : 0);
  plan[1] = pl.PC; plan[2] = pl.Pchunk; plan[3] = pl.NSUB; plan[4] = pl.parts;
  return DUST_OK;
}

extern "C" size_t dust_rollout_workspace_bytes(const dust_rollout_args* a) {
  if (!a || a->B <= 0 || a->N <= 0 || a->S <= 0) return 0;
  return plan_rollout(a).total;
}

// reduce_only: `costs` is complete on entry (dust_cost_reduce); only the reductions after it run
static int rollout_cost_impl(const dust_rollout_args* a, const TailParams* tail, void* stream_, bool reduce_only = false) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DUST_REQUIRE(a != nullptr, DUST_ERR_INVALID_ARG, "dust_rollout_cost: args is NULL");
  int rc = validate_model(a->model);
  if (rc) return rc;
  DUST_REQUIRE(a->B > 0 && a->N > 0 && a->S > 0 && a->H > 0, DUST_ERR_INVALID_ARG,
               "dust_rollout_cost: B,N,S,H must be positive (got %d,%d,%d,%d)", a->B, a->N, a->S, a->H);
  DUST_REQUIRE(a->state0 && a->noise, DUST_ERR_INVALID_ARG, "dust_rollout_cost: state0 and noise are required");
  DUST_REQUIRE(!a->theta || a->sigma, DUST_ERR_INVALID_ARG, "dust_rollout_cost: sigma is required with theta");
  DUST_REQUIRE(!a->params || a->P > 0, DUST_ERR_INVALID_ARG, "dust_rollout_cost: P must be positive with params");
  DUST_REQUIRE(!a->grad_lik || a->theta, DUST_ERR_INVALID_ARG,
               "dust_rollout_cost: grad_lik needs theta (analytic score of N(theta, sigma^2))");
  DUST_REQUIRE(a->alpha > 0.f && a->temperature > 0.f, DUST_ERR_INVALID_ARG,
               "dust_rollout_cost: alpha and temperature must be positive");
  const int kind = a->model->kind;
  const int A = model_da(kind);
  const int P = a->params ? a->P : 1;
  const long long SN = (long long)a->S * a->N;
  DUST_REQUIRE(SN < (1ll << 30), DUST_ERR_UNSUPPORTED, "dust_rollout_cost: S*N too large");

  const bool ranged = a->p_end > a->p_begin;
  DUST_REQUIRE(a->p_begin >= 0 && a->p_end >= a->p_begin && a->p_end <= P, DUST_ERR_INVALID_ARG,
               "dust_rollout_cost: draw range [%d, %d) outside [0, %d)", a->p_begin, a->p_end, P);
  DUST_REQUIRE(!ranged || (!tail && !reduce_only && a->costs && !a->log_lik && !a->lik_weights && !a->grad_lik &&
                           !a->mppi_weights && !a->mppi_delta && !a->mix && !a->states && !a->sigma_weights && !a->ctrl_mat),
               DUST_ERR_INVALID_ARG, "dust_rollout_cost: a draw range only produces `costs` (its share of the mean)");
  DUST_REQUIRE(!reduce_only || a->costs, DUST_ERR_INVALID_ARG, "dust_cost_reduce: costs (input) is required");
  RolloutPlan pl = plan_rollout(a, tail != nullptr);
  if (reduce_only) { pl.PC = 1; pl.NSUB = 1; pl.parts = 1; }
  DUST_REQUIRE(pl.total == 0 || (a->workspace && a->workspace_bytes >= pl.total), DUST_ERR_WORKSPACE,
               "dust_rollout_cost: workspace needs %zu bytes, got %zu", pl.total, a->workspace_bytes);
  char* ws = (char*)a->workspace;
  float* costs = a->costs ? a->costs : (float*)(ws + pl.off_costs);
  float* part = pl.parts > 1 ? (float*)(ws + pl.off_part) : nullptr;
  float* likw = a->lik_weights ? a->lik_weights : ((a->grad_lik) ? (float*)(ws + pl.off_likw) : nullptr);
  float* mppiw = a->mppi_weights ? a->mppi_weights : ((a->mppi_delta) ? (float*)(ws + pl.off_mppiw) : nullptr);

  RolloutKParams k;
  k.m = to_params(*a->model);
  k.B = a->B; k.N = a->N; k.S = a->S; k.P = P; k.H = a->H; k.A = A;
  k.SN = (int)SN; k.HA = a->H * A; k.PC = pl.PC; k.Pchunk = pl.Pchunk; k.parts = pl.parts;
  k.p0 = ranged ? a->p_begin : 0; k.p1 = ranged ? a->p_end : P;
  k.interleaved = a->param_tiling == DUST_PARAMS_INTERLEAVED;
  k.state0 = a->state0; k.theta = a->theta; k.noise = a->noise; k.sigma = a->sigma; k.params = a->params;
  k.cost_out = pl.parts > 1 ? part : costs;
  k.states = a->states;
  k.ut_w = a->sigma_weights; k.ctrl_mat = a->ctrl_mat; k.a_seq = a->a_seq; k.ctrl_reg = a->ctrl_reg;
  DUST_REQUIRE(!a->sigma_weights || (a->params && a->param_tiling == DUST_PARAMS_BLOCKED), DUST_ERR_INVALID_ARG,
               "dust_rollout_cost: sigma_weights needs the sigma points in params (blocked tiling)");
  DUST_REQUIRE(!(a->sigma_weights || a->ctrl_mat) || !tail, DUST_ERR_UNSUPPORTED,
               "dust_svmpc_step: sigma-point weights and the control regulariser need the staged path");

  const int stride = padded_stride(k.HA);
  const size_t grid_bytes = (kind == DUST_MODEL_PARTICLE && a->model->grid_bits)
                                ? sizeof(uint32_t) * ((a->model->grid_nx * a->model->grid_ny + 31) / 32) : 0;
  // ---- few instances: one thread-block cluster per instance, the whole step in one launch ----
  ClusterLayout CL;
  if (cluster_path_ok(a, reduce_only, &CL) && (a->log_lik || a->grad_lik || tail)) {
    FusedOut o{a->costs, a->log_lik, a->grad_lik, a->likelihood, a->alpha, TailParams{}};
    if (tail) o.tail = *tail;
    k.cost_out = nullptr;
    const size_t csmem = (size_t)CL.total_floats * 4;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a->B * kClusterSize), 1, 1);
    cfg.blockDim = dim3((unsigned)CL.threads, 1, 1);
    cfg.dynamicSmemBytes = csmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kClusterSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
#define DUST_CLUSTER(MODEL)                                                                                               \
  do {                                                                                                                    \
    if (csmem > 48 * 1024)                                                                                                \
      DUST_CUDA_OK(cudaFuncSetAttribute(svmpc_cluster_kernel<MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem)); \
    { DUST_TIMED("svmpc_cluster_kernel", stream); DUST_CUDA_OK(cudaLaunchKernelEx(&cfg, svmpc_cluster_kernel<MODEL>, k, o, CL)); } \
  } while (0)
    if (kind == DUST_MODEL_PENDULUM) DUST_CLUSTER(DUST_MODEL_PENDULUM);
    else DUST_CLUSTER(DUST_MODEL_PARTICLE);
#undef DUST_CLUSTER
    DUST_LAUNCH_OK("svmpc_cluster_kernel");
    return DUST_OK;
  }
  // ---- fused per-instance path: everything the SVGD step needs in one launch -----------------
  const bool fused_ok = fused_path_ok(a, pl, tail != nullptr, reduce_only);
  DUST_REQUIRE(fused_ok || !tail, DUST_ERR_UNSUPPORTED,
               "dust_svmpc_step: the one-launch control step needs N <= 32 and either few instances (B < 74: S*N <= 4096, "
               "P <= 64) or many (B >= 74: H*A <= 32)");
  if (fused_ok) {
    const int thst_h = (k.HA + 3) & ~3;
    const size_t fsmem = sizeof(float) * ((size_t)2 * kFusedThreads * stride + (size_t)a->N * thst_h +
                                          (size_t)3 * a->N * thst_h + (size_t)(3 + 8) * a->N) + grid_bytes;
    FusedOut o{a->costs, a->log_lik, a->grad_lik, a->likelihood, a->alpha, TailParams{}};
    if (tail) o.tail = *tail;
    k.cost_out = nullptr;
#define DUST_FUSED(MODEL, ACC, TPT)                                                                                         \
  do {                                                                                                                      \
    if (fsmem > 48 * 1024)                                                                                                  \
      DUST_CUDA_OK(cudaFuncSetAttribute(svmpc_instance_kernel<MODEL, ACC, TPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem)); \
    { DUST_TIMED("svmpc_instance_kernel", stream); svmpc_instance_kernel<MODEL, ACC, TPT><<<a->B, kFusedThreads / TPT, fsmem, stream>>>(k, o); } \
  } while (0)
#if DUST_PEND_PAIR
    // second-generation kernel: packed pendulum path with 16-byte rows and at most 32 policies; DUST_B200_FUSED_V1=1
    // forces the first one (A/B measurements)
    static const bool force_v1 = getenv("DUST_B200_FUSED_V1") != nullptr || getenv("DUST_B200_NO_PAIR") != nullptr;
    if (kind == DUST_MODEL_PENDULUM && !force_v1 && (k.HA & 3) == 0 && a->N <= 32 && ((((uintptr_t)a->noise) & 15) == 0)) {
      // DUST_B200_QUAD=1: two pairs per lane (svmpc_quad_kernel; needs tiles without a ragged end).  Measured 3 % SLOWER
      // than one pair per lane at the bench shape (0.320 vs 0.309 ms: 16 warps/SM with two chains each lose to 28 with
      // one, profiles/r2_instance_kernel_ncu.md) -- kept as an A/B build of the same arithmetic, off by default.
      const bool use_quad = getenv("DUST_B200_QUAD") != nullptr;   // read per call: the tests toggle it
      const int La_h = (32 / a->N) * a->N;
      const WarpKernelSmem Lq = warp_kernel_smem(a->N, k.HA, 2);
      if (use_quad && k.SN % (4 * La_h) == 0 && Lq.total_bytes <= 227 * 1024) {
#define DUST_QUAD_KERNEL(HA4)                                                                                              \
  do {                                                                                                                     \
    if (Lq.total_bytes > 48 * 1024)                                                                                        \
      DUST_CUDA_OK(cudaFuncSetAttribute(svmpc_quad_kernel<HA4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lq.total_bytes)); \
    { DUST_TIMED("svmpc_instance_kernel", stream); svmpc_quad_kernel<HA4><<<a->B, kWarpKernelThreads, Lq.total_bytes, stream>>>(k, o); } \
  } while (0)
        if (k.HA == 20) DUST_QUAD_KERNEL(5);
        else DUST_QUAD_KERNEL(0);
#undef DUST_QUAD_KERNEL
        DUST_LAUNCH_OK("svmpc_instance_kernel");
        return DUST_OK;
      }
      const WarpKernelSmem Lw = warp_kernel_smem(a->N, k.HA);
      if (Lw.total_bytes <= 227 * 1024) {
#define DUST_WARP_KERNEL(HA4)                                                                                              \
  do {                                                                                                                     \
    if (Lw.total_bytes > 48 * 1024)                                                                                        \
      DUST_CUDA_OK(cudaFuncSetAttribute(svmpc_warp_kernel<HA4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lw.total_bytes)); \
    { DUST_TIMED("svmpc_instance_kernel", stream); svmpc_warp_kernel<HA4><<<a->B, kWarpKernelThreads, Lw.total_bytes, stream>>>(k, o); } \
  } while (0)
        if (k.HA == 20) DUST_WARP_KERNEL(5);
        else DUST_WARP_KERNEL(0);
#undef DUST_WARP_KERNEL
        DUST_LAUNCH_OK("svmpc_instance_kernel");
        return DUST_OK;
      }
    }
#endif
    if (kind == DUST_MODEL_PENDULUM) {
      // two trajectories per thread on the packed FP32 pipe when both rows of a thread share a policy
      // (half a tile is a whole number of policy groups); DUST_B200_NO_PAIR=1 forces the scalar kernel
      static const bool no_pair = getenv("DUST_B200_NO_PAIR") != nullptr;
      const bool pair = DUST_PEND_PAIR && !no_pair && ((kFusedThreads / a->N) % 2 == 0) && a->N <= kFusedThreads / 2;
#if DUST_PEND_PAIR
#define DUST_FUSED_PEND(ACC) do { if (pair) DUST_FUSED(DUST_MODEL_PENDULUM, ACC, 2); else DUST_FUSED(DUST_MODEL_PENDULUM, ACC, 1); } while (0)
#else
#define DUST_FUSED_PEND(ACC) DUST_FUSED(DUST_MODEL_PENDULUM, ACC, 1)
      (void)pair;
#endif
      if (k.HA <= 8) DUST_FUSED_PEND(8);
      else if (k.HA <= 16) DUST_FUSED_PEND(16);
      else if (k.HA <= 20) DUST_FUSED_PEND(20);
      else if (k.HA <= 24) DUST_FUSED_PEND(24);
      else DUST_FUSED_PEND(32);
#undef DUST_FUSED_PEND
    } else {
      if (k.HA <= 16) DUST_FUSED(DUST_MODEL_PARTICLE, 16, 1);
      else DUST_FUSED(DUST_MODEL_PARTICLE, 32, 1);
    }
#undef DUST_FUSED
    DUST_LAUNCH_OK("svmpc_instance_kernel");
    return DUST_OK;
  }
  if (!reduce_only) {
  size_t smem = sizeof(float) * kTile * stride + grid_bytes;
  DUST_REQUIRE(smem <= 227 * 1024, DUST_ERR_UNSUPPORTED, "dust_rollout_cost: H*A=%d needs %zu B of shared memory", k.HA, smem);
  const int tiles = ceil_div(SN, kTile);
  const long long gx = (long long)a->B * tiles;
  DUST_REQUIRE(gx < (1ll << 31) && pl.PC <= 65535, DUST_ERR_UNSUPPORTED, "dust_rollout_cost: grid too large");
  dim3 grid((unsigned)gx, (unsigned)pl.PC, 1);
  const bool ext = a->sigma_weights || a->ctrl_mat;
#define DUST_ROLLOUT(MODEL, EXT, NSUB)                                                                                 \
  do {                                                                                                                 \
    if (smem > 48 * 1024)                                                                                              \
      DUST_CUDA_OK(cudaFuncSetAttribute(rollout_cost_kernel<MODEL, EXT, NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    { DUST_TIMED("rollout_cost_kernel", stream); rollout_cost_kernel<MODEL, EXT, NSUB><<<grid, kTile * NSUB, smem, stream>>>(k); }  \
  } while (0)
#define DUST_ROLLOUT_SUB(MODEL)                                                                                        \
  do {                                                                                                                 \
    if (ext) DUST_ROLLOUT(MODEL, true, 1);                                                                             \
    else if (pl.NSUB == 4) DUST_ROLLOUT(MODEL, false, 4);                                                              \
    else if (pl.NSUB == 2) DUST_ROLLOUT(MODEL, false, 2);                                                              \
    else DUST_ROLLOUT(MODEL, false, 1);                                                                                \
  } while (0)
  if (kind == DUST_MODEL_PENDULUM) DUST_ROLLOUT_SUB(DUST_MODEL_PENDULUM);
  else DUST_ROLLOUT_SUB(DUST_MODEL_PARTICLE);
#undef DUST_ROLLOUT_SUB
#undef DUST_ROLLOUT
  DUST_LAUNCH_OK("rollout_cost_kernel");
  }

  const bool need_stats = pl.parts > 1 || a->log_lik || likw || mppiw || a->mix;
  if (need_stats) {
    SoftminKParams s;
    s.B = a->B; s.N = a->N; s.S = a->S; s.P = P; s.PC = pl.parts; s.SN = (int)SN;
    s.likelihood = a->likelihood; s.alpha = a->alpha; s.inv_temp = 1.0f / a->temperature;
    s.cost_part = part; s.costs = costs; s.log_lik = a->log_lik; s.lik_w = likw; s.mppi_w = mppiw; s.mix = a->mix;
    s.eta = nullptr; s.spread = 0;
    if (a->B * 2 >= kNumSMs) {   // many instances: one CTA each
      { DUST_TIMED("policy_softmin_kernel", stream); policy_softmin_kernel<<<a->B, kSoftminThreads, sizeof(float) * a->N, stream>>>(s); }
      DUST_LAUNCH_OK("policy_softmin_kernel");
    } else {                     // few: spread the trajectories, then the policies (two warps per CTA), over the SMs
      DUST_REQUIRE(a->B <= 65535, DUST_ERR_UNSUPPORTED, "dust_rollout_cost: B > 65535");
      s.spread = 1;
      s.eta = a->mix ? (float*)(ws + pl.off_eta) : nullptr;
      if (pl.parts > 1) {
        { DUST_TIMED("combine_cost_chunks_kernel", stream); combine_cost_chunks_kernel<<<dim3((unsigned)ceil_div(SN, 256), (unsigned)a->B, 1), 256, 0, stream>>>(s); }
        DUST_LAUNCH_OK("combine_cost_chunks_kernel");
      }
      if (a->log_lik || likw || mppiw || a->mix) {
        { DUST_TIMED("policy_softmin_kernel", stream); policy_softmin_kernel<<<dim3((unsigned)a->B, (unsigned)ceil_div(a->N, 2), 1), 64, sizeof(float) * a->N, stream>>>(s); }
        DUST_LAUNCH_OK("policy_softmin_kernel");
      }
      if (a->mix) {
        { DUST_TIMED("policy_mix_kernel", stream); policy_mix_kernel<<<a->B, 32, 0, stream>>>(s.eta, a->mix, a->N); }
        DUST_LAUNCH_OK("policy_mix_kernel");
      }
    }
  }
  if (a->grad_lik || a->mppi_delta) {
    ColsumKParams c;
    c.B = a->B; c.N = a->N; c.S = a->S; c.HA = k.HA; c.A = A; c.W = a->N * k.HA;
    c.theta = a->theta; c.noise = a->noise; c.sigma = a->theta ? a->sigma : nullptr; c.a_seq = a->a_seq; c.pert = a->pert;
    c.lik_w = likw; c.mppi_w = mppiw; c.grad_lik = a->grad_lik; c.mppi_delta = a->mppi_delta;
    const int cols_per_cta = c.W < kColsumThreads ? c.W : kColsumThreads;
    dim3 g2((unsigned)ceil_div(c.W, cols_per_cta), (unsigned)a->B, 1);
    DUST_REQUIRE(a->B <= 65535, DUST_ERR_UNSUPPORTED, "dust_rollout_cost: B > 65535 with gradient outputs");
    { DUST_TIMED("weighted_colsum_kernel", stream); weighted_colsum_kernel<<<g2, kColsumThreads, 0, stream>>>(c); }
    DUST_LAUNCH_OK("weighted_colsum_kernel");
  }
  return DUST_OK;
}

extern "C" int dust_rollout_cost(const dust_rollout_args* a, void* stream_) { return rollout_cost_impl(a, nullptr, stream_); }

extern "C" int dust_cost_reduce(const dust_rollout_args* a, void* stream_) { return rollout_cost_impl(a, nullptr, stream_, true); }

extern "C" int dust_svmpc_step(const dust_svmpc_step_args* s, void* stream_) {
  DUST_REQUIRE(s != nullptr, DUST_ERR_INVALID_ARG, "dust_svmpc_step: args is NULL");
  const dust_rollout_args* a = &s->rollout;
  DUST_REQUIRE(a->theta && a->sigma && s->inv_var, DUST_ERR_INVALID_ARG, "dust_svmpc_step: theta, sigma and inv_var are required");
  DUST_REQUIRE(s->prior_aliased || s->mu, DUST_ERR_INVALID_ARG, "dust_svmpc_step: mu is required unless the prior aliases theta");
  DUST_REQUIRE(s->theta_out || s->do_forward, DUST_ERR_INVALID_ARG, "dust_svmpc_step: no output requested");
  DUST_REQUIRE(!s->do_forward || (s->p_weights && s->theta_next), DUST_ERR_INVALID_ARG,
               "dust_svmpc_step: p_weights and theta_next are required with do_forward");
  DUST_REQUIRE(s->roll_strategy == DUST_ROLL_REPEAT || s->roll_strategy == DUST_ROLL_MEAN, DUST_ERR_INVALID_ARG,
               "dust_svmpc_step: invalid roll strategy %d", s->roll_strategy);
  DUST_REQUIRE(s->theta_next != a->theta && s->theta_out != a->theta, DUST_ERR_INVALID_ARG,
               "dust_svmpc_step: outputs must not alias theta");
  TailParams t;
  t.enabled = 1; t.do_forward = s->do_forward; t.roll = s->roll_strategy; t.weighted = s->weighted_prior;
  t.aliased = s->prior_aliased; t.mu = s->mu; t.mix = s->mix; t.inv_var = s->inv_var; t.log_norm = s->log_norm;
  t.gamma = s->gamma; t.c1 = s->c1; t.c2 = s->c2; t.lr = s->lr; t.theta_out = s->theta_out; t.phi = s->phi;
  t.p_weights = s->p_weights; t.i_star = s->i_star; t.a_seq = s->a_seq; t.theta_next = s->theta_next; t.mix_next = s->mix_next;
  return rollout_cost_impl(a, &t, stream_);
}
