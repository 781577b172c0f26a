// MPF: SVGD over dynamics-parameter particles, every optimisation step inside one launch
// (one CTA per MPC instance; particles, scores and the pairwise work live in shared memory).
// Also the stand-alone model step / cost entry points used by the host-side drivers.
//
// Reference: dust/inference/mpf.py:40-86, dust/inference/likelihoods.py:30-49 (one model step
// per parameter particle; the autograd call of mpf.py:50 is replaced by the closed-form
// Jacobian of each shipped model, SURVEY.md §9 "MPF one-step").
// Compiled with -fmad=false (models.cuh).
#include <cooperative_groups.h>

#include "models.cuh"

namespace dust {

constexpr int kMpfMaxThreads = 1024;
constexpr int kMaxDp = 2;

struct MpfKParams {
  ModelParams m;
  int B, Np, dp, n_steps, log_space;
  float* x;
  const float *obs0, *action, *obs1, *prior_inv_var;
  float inv_obs_var, bw, lr;
  float* grad_norms;
  const float* bw_dev;   // device scalar overriding bw, or nullptr
  float* phi_out;        // [B,Np,dp] phi of the last step, or nullptr
  int lanes;             // lanes per particle in mpf_kernel: 32, or 16 / 8 when all particles then fit one pass
};

// d log N(obs1; f(obs0, a; p), obs_std^2 I) / d x  for one particle (x = p or log p)
template <int MODEL>
__device__ __forceinline__ void mpf_lik_grad(const ModelParams& m, const float* xp, int log_space, const float* o0,
                                             const float* act, const float* o1, float inv_obs_var, float c_cell,
                                             float* g) {
  if (MODEL == DUST_MODEL_PENDULUM) {
    const float l = log_space ? expf(xp[0]) : xp[0];
    const float ms = log_space ? expf(xp[1]) : xp[1];
    const PendulumCoef cf = pendulum_coef_sampled(m, l, ms);
    float th = o0[0], om = o0[1], pre;
    const float a = act[0];
    const float u = fminf(fmaxf(a, -m.max_torque), m.max_torque);
    const float sn = sinf(th + kPiF);
    pendulum_step(m, cf, th, om, a, &pre);
    const float r_th = (o1[0] - th) * inv_obs_var, r_om = (o1[1] - om) * inv_obs_var;
    const bool in8 = pre >= -m.max_speed_pend && pre <= m.max_speed_pend;
    // d om'/d l = dt (3g/(2 l^2) sin(th+pi) - 6u/(m l^3)),  d om'/d m = -dt 3u/(m^2 l^2);  th' = th + dt om'
    const float dom_dl = in8 ? m.dt * (3.0f * m.g / (2.0f * l * l) * sn - 6.0f * u / (ms * l * l * l)) : 0.f;
    const float dom_dm = in8 ? m.dt * (-3.0f * u / (ms * ms * l * l)) : 0.f;
    const float w = r_th * m.dt + r_om;
    g[0] = w * dom_dl * (log_space ? l : 1.0f);
    g[1] = w * dom_dm * (log_space ? ms : 1.0f);
  } else {
    const float ms = log_space ? expf(xp[0]) : xp[0];
    ParticleState s{o0[0], o0[1], o0[2], o0[3]};
    float vpre[2];
    const float ax = act[0], ay = act[1];
    particle_step(m, s, ax, ay, ms, c_cell, vpre);
    const float kk = m.can_crash ? m.dt * (1.0f - c_cell) : m.dt;
    const float amx = ax / ms, amy = ay / ms;
    const bool max_ = amx >= -m.max_accel && amx <= m.max_accel, may_ = amy >= -m.max_accel && amy <= m.max_accel;
    const bool mvx = vpre[0] >= -m.max_speed && vpre[0] <= m.max_speed, mvy = vpre[1] >= -m.max_speed && vpre[1] <= m.max_speed;
    const float r_vx = (o1[2] - s.vx) * inv_obs_var, r_vy = (o1[3] - s.vy) * inv_obs_var;
    float gm = 0.f;
    if (max_ && mvx) gm += r_vx * kk * (-ax / (ms * ms));
    if (may_ && mvy) gm += r_vy * kk * (-ay / (ms * ms));
    g[0] = gm * (log_space ? ms : 1.0f);
  }
}

template <int MODEL>
__global__ void __launch_bounds__(kMpfMaxThreads) mpf_kernel(const MpfKParams k) {
  constexpr int DP = (MODEL == DUST_MODEL_PENDULUM) ? 2 : 1;
  constexpr int DS = (MODEL == DUST_MODEL_PENDULUM) ? 2 : 4;
  constexpr int DA = (MODEL == DUST_MODEL_PENDULUM) ? 1 : 2;
  extern __shared__ float sm[];
  float* xs = sm;                       // [Np][DP]
  float* sc = sm + k.Np * DP;           // [Np][DP] score
  float* ph = sc + k.Np * DP;           // [Np][DP] phi
  __shared__ float red[kMpfMaxThreads / 8];
  __shared__ float s_cell;
  const long long inst = blockIdx.x;
  float* xg = k.x + inst * (long long)k.Np * DP;
  float o0[DS], o1[DS], act[DA], piv[DP];
#pragma unroll
  for (int i = 0; i < DS; ++i) { o0[i] = k.obs0[inst * DS + i]; o1[i] = k.obs1[inst * DS + i]; }
#pragma unroll
  for (int i = 0; i < DA; ++i) act[i] = k.action[inst * DA + i];
#pragma unroll
  for (int i = 0; i < DP; ++i) piv[i] = k.prior_inv_var[i];
  for (int e = threadIdx.x; e < k.Np * DP; e += blockDim.x) xs[e] = xg[e];
  if (threadIdx.x == 0) {
    float c = 0.f;
    if (MODEL == DUST_MODEL_PARTICLE && k.m.grid_bits) c = grid_lookup(k.m, k.m.grid_bits, o0[0], o0[1]);
    s_cell = c;
  }
  __syncthreads();
  const float c_cell = s_cell;
  const float bw = k.bw_dev ? __ldg(k.bw_dev) : k.bw;
  const float inv_bw2 = 1.0f / (bw * bw);
  const float inv_np = 1.0f / (float)k.Np;

  // A group of LP lanes per particle (LP = 32: one warp), lanes over the other particles (shuffle sums): with a
  // thread per particle the 512-particle stress shape was three serial passes of 512 exponentials per thread.
  // Few particles (the demos: 50) take LP = 16 so that ALL of them are worked on in one pass of the CTA.
  const int LP = k.lanes;
  const int warp = threadIdx.x / LP, lane = threadIdx.x % LP, nwarps = blockDim.x / LP;
  // the lanes of THIS group only: the groups of one hardware warp may run different trip counts
  const unsigned gmask = LP == 32 ? 0xffffffffu : (((1u << LP) - 1u) << (((threadIdx.x & 31) / LP) * LP));
  auto group_sum = [&](float v) {
    for (int o = LP >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
  };
  for (int step = 0; step < k.n_steps; ++step) {
    // score_i = grad log-likelihood + grad log GMM(x_i; centres = current particles).  The centres
    // ARE the particles (mpf.py:32-38): the j = i term has exponent 0 and every other one is <= 0,
    // so the max shift of the log-sum-exp is exactly 0 and needs no pass of its own.
    for (int i = warp; i < k.Np; i += nwarps) {
      float xi[DP], g[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) xi[d] = xs[i * DP + d];
      mpf_lik_grad<MODEL>(k.m, xi, k.log_space, o0, act, o1, k.inv_obs_var, c_cell, g);
      float z = 0.f, acc[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) acc[d] = 0.f;
      for (int j = lane; j < k.Np; j += LP) {
        float q = 0.f;
#pragma unroll
        for (int d = 0; d < DP; ++d) { const float df = xi[d] - xs[j * DP + d]; q += df * df * piv[d]; }
        const float e = expf(-0.5f * q);
        z += e;
#pragma unroll
        for (int d = 0; d < DP; ++d) acc[d] += e * (xs[j * DP + d] - xi[d]);
      }
      z = group_sum(z);
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const float a_d = group_sum(acc[d]);
        if (lane == 0) sc[i * DP + d] = g[d] + a_d / z * piv[d];
      }
    }
    __syncthreads();
    // phi_i = (1/Np) sum_j K_ij s_j - (1/bw^2) sum_j K_ij (x_i - x_j)     (mpf.py:53-56)
    float nrm = 0.f;
    for (int i = warp; i < k.Np; i += nwarps) {
      float xi[DP], acc[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) { xi[d] = xs[i * DP + d]; acc[d] = 0.f; }
      for (int j = lane; j < k.Np; j += LP) {
        float d2 = 0.f, df[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) { df[d] = xi[d] - xs[j * DP + d]; d2 += df[d] * df[d]; }
        const float kij = expf(-d2 * inv_bw2 * 0.5f);
#pragma unroll
        for (int d = 0; d < DP; ++d) acc[d] += kij * (inv_np * sc[j * DP + d] - inv_bw2 * df[d]);
      }
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const float a_d = group_sum(acc[d]);
        if (lane == 0) { ph[i * DP + d] = a_d; nrm += a_d * a_d; }
      }
    }
    if (lane == 0) red[warp] = nrm;
    __syncthreads();
    if (threadIdx.x < 32 && k.grad_norms) {     // one warp adds the per-group partial norms (a serial loop in one
      float t = 0.f;                            // thread held every other warp at the barrier below)
      for (int w = threadIdx.x; w < nwarps; w += 32) t += red[w];
      t = warp_sum(t);
      if (threadIdx.x == 0) k.grad_norms[inst * k.n_steps + step] = sqrtf(t);
    }
    if (k.phi_out && step == k.n_steps - 1)
      for (int e = threadIdx.x; e < k.Np * DP; e += blockDim.x) k.phi_out[inst * (long long)k.Np * DP + e] = ph[e];
    for (int e = threadIdx.x; e < k.Np * DP; e += blockDim.x) xs[e] = xs[e] + k.lr * ph[e];  // SGD (mpf.py:59-62)
    __syncthreads();
  }
  for (int e = threadIdx.x; e < k.Np * DP; e += blockDim.x) xg[e] = xs[e];
}

// One large instance (B = 1) spread over many SMs: the same warp-per-particle arithmetic, particles and
// scores in global memory (read through L2: other SMs wrote them), two grid barriers per SVGD step.
// ws: [Np*DP] second particle buffer | [Np*DP] scores | [gridDim.x] partial squared norms.
namespace cg = cooperative_groups;
template <int MODEL>
__global__ void __launch_bounds__(256) mpf_coop_kernel(const MpfKParams k, float* ws) {
  constexpr int DP = (MODEL == DUST_MODEL_PENDULUM) ? 2 : 1;
  constexpr int DS = (MODEL == DUST_MODEL_PENDULUM) ? 2 : 4;
  constexpr int DA = (MODEL == DUST_MODEL_PENDULUM) ? 1 : 2;
  cg::grid_group grid = cg::this_grid();
  __shared__ float red[8];
  float* cur = k.x;
  float* nxt = ws;
  float* sc = ws + k.Np * DP;
  float* part = sc + k.Np * DP;
  float o0[DS], o1[DS], act[DA], piv[DP];
#pragma unroll
  for (int i = 0; i < DS; ++i) { o0[i] = k.obs0[i]; o1[i] = k.obs1[i]; }
#pragma unroll
  for (int i = 0; i < DA; ++i) act[i] = k.action[i];
#pragma unroll
  for (int i = 0; i < DP; ++i) piv[i] = k.prior_inv_var[i];
  float c_cell = 0.f;
  if (MODEL == DUST_MODEL_PARTICLE && k.m.grid_bits) c_cell = grid_lookup(k.m, k.m.grid_bits, o0[0], o0[1]);
  const float bw = k.bw_dev ? __ldg(k.bw_dev) : k.bw;
  const float inv_bw2 = 1.0f / (bw * bw);
  const float inv_np = 1.0f / (float)k.Np;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int gwarp = blockIdx.x * nwarps + warp, total_warps = gridDim.x * nwarps;
  for (int step = 0; step < k.n_steps; ++step) {
    for (int i = gwarp; i < k.Np; i += total_warps) {
      float xi[DP], g[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) xi[d] = __ldcg(cur + i * DP + d);
      mpf_lik_grad<MODEL>(k.m, xi, k.log_space, o0, act, o1, k.inv_obs_var, c_cell, g);
      float z = 0.f, acc[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) acc[d] = 0.f;
      for (int j = lane; j < k.Np; j += 32) {
        float q = 0.f, xj[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) { xj[d] = __ldcg(cur + j * DP + d); const float df = xi[d] - xj[d]; q += df * df * piv[d]; }
        const float e = expf(-0.5f * q);
        z += e;
#pragma unroll
        for (int d = 0; d < DP; ++d) acc[d] += e * (xj[d] - xi[d]);
      }
      z = warp_sum(z);
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const float a_d = warp_sum(acc[d]);
        if (lane == 0) sc[i * DP + d] = g[d] + a_d / z * piv[d];
      }
    }
    grid.sync();
    float nrm = 0.f;
    for (int i = gwarp; i < k.Np; i += total_warps) {
      float xi[DP], acc[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) { xi[d] = __ldcg(cur + i * DP + d); acc[d] = 0.f; }
      for (int j = lane; j < k.Np; j += 32) {
        float d2 = 0.f, df[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) { df[d] = xi[d] - __ldcg(cur + j * DP + d); d2 += df[d] * df[d]; }
        const float kij = expf(-d2 * inv_bw2 * 0.5f);
#pragma unroll
        for (int d = 0; d < DP; ++d) acc[d] += kij * (inv_np * __ldcg(sc + j * DP + d) - inv_bw2 * df[d]);
      }
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const float a_d = warp_sum(acc[d]);
        if (lane == 0) { nxt[i * DP + d] = xi[d] + k.lr * a_d; nrm += a_d * a_d; }   // SGD (mpf.py:59-62)
      }
    }
    if (lane == 0) red[warp] = nrm;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < nwarps; ++w) t += red[w];
      part[blockIdx.x] = t;
    }
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0 && k.grad_norms) {
      float t = 0.f;
      for (int b = 0; b < (int)gridDim.x; ++b) t += __ldcg(part + b);
      k.grad_norms[step] = sqrtf(t);
    }
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  if (cur != k.x)   // an odd number of steps left the result in the workspace
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < k.Np * DP; e += gridDim.x * blockDim.x) k.x[e] = __ldcg(cur + e);
}

template <int MODEL>
__global__ void model_step_kernel(const ModelParams m, int M, const float* __restrict__ states,
                                  const float* __restrict__ actions, const float* __restrict__ params,
                                  float* __restrict__ next) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  if (MODEL == DUST_MODEL_PENDULUM) {
    const PendulumCoef cf = params ? pendulum_coef_sampled(m, params[2 * i], params[2 * i + 1]) : pendulum_coef_default(m);
    float th = states[2 * i], om = states[2 * i + 1];
    pendulum_step(m, cf, th, om, actions[i]);
    next[2 * i] = th;
    next[2 * i + 1] = om;
  } else {
    ParticleState s{states[4 * i], states[4 * i + 1], states[4 * i + 2], states[4 * i + 3]};
    const float c = m.grid_bits ? grid_lookup(m, m.grid_bits, s.x, s.y) : 0.f;
    particle_step(m, s, actions[2 * i], actions[2 * i + 1], params ? params[i] : m.default_mass, c);
    next[4 * i] = s.x; next[4 * i + 1] = s.y; next[4 * i + 2] = s.vx; next[4 * i + 3] = s.vy;
  }
}

template <int MODEL>
__global__ void model_cost_kernel(const ModelParams m, int M, int terminal, const float* __restrict__ states,
                                  const float* __restrict__ actions, float* __restrict__ costs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  if (MODEL == DUST_MODEL_PENDULUM) {
    costs[i] = pendulum_cost(m, states[2 * i], states[2 * i + 1]);
  } else {
    ParticleState s{states[4 * i], states[4 * i + 1], states[4 * i + 2], states[4 * i + 3]};
    const float c = m.grid_bits ? grid_lookup(m, m.grid_bits, s.x, s.y) : 0.f;
    if (terminal) costs[i] = particle_term_cost(m, s, c);
    else costs[i] = particle_inst_cost(m, s, actions ? actions[2 * i] : 0.f, actions ? actions[2 * i + 1] : 0.f, c);
  }
}

}  // namespace dust

using namespace dust;

static int mpf_coop_grid(const dust_mpf_args* a) {
  if (a->B != 1 || a->Np < 128) return 0;
  const int g = (a->Np + 7) / 8;   // 8 warps per CTA, one particle per warp and pass where possible
  return g < kNumSMs ? g : kNumSMs;
}

namespace dust {
// KDEpy's silvermans_rule on n <= 4096 values: bitonic sort in shared memory, the two percentiles by linear
// interpolation and the sample standard deviation in double (numpy works on the float64 copy of the particles)
__global__ void __launch_bounds__(1024) silverman_kernel(const float* __restrict__ x, int n, int npad, float scale, float* bw_out,
                                                         float* inv_var_out, int dp) {
  extern __shared__ float sv[];   // [npad] sorted ascending, padded with +inf
  __shared__ double red_s[32], red_q[32];
  const int tid = threadIdx.x;
  for (int i = tid; i < npad; i += blockDim.x) sv[i] = i < n ? x[i] : INFINITY;
  __syncthreads();
  for (int size = 2; size <= npad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < npad; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const float a = sv[i], b = sv[j];
          if ((a > b) == up) { sv[i] = b; sv[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  // mean, then the sum of squared deviations (two passes, double)
  double s = 0.0;
  for (int i = tid; i < n; i += blockDim.x) s += (double)sv[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((tid & 31) == 0) red_s[tid >> 5] = s;
  __syncthreads();
  double mean = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mean += red_s[w];
  mean /= (double)n;
  double q = 0.0;
  for (int i = tid; i < n; i += blockDim.x) { const double d = (double)sv[i] - mean; q += d * d; }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if ((tid & 31) == 0) red_q[tid >> 5] = q;
  __syncthreads();
  if (tid == 0) {
    double ss = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) ss += red_q[w];
    auto pct = [&](double p) {      // numpy.percentile, linear interpolation
      const double pos = p / 100.0 * (double)(n - 1);
      const int lo = (int)floor(pos);
      const int hi = lo + 1 < n ? lo + 1 : lo;
      const double t = pos - (double)lo;
      const double a = (double)sv[lo], b = (double)sv[hi];
      return t >= 0.5 ? b - (b - a) * (1.0 - t) : a + (b - a) * t;
    };
    double bw = 1.0;
    if (n > 1) {
      const double sd = sqrt(ss / (double)(n - 1));
      const double iqr = (pct(75.0) - pct(25.0)) / 1.349;
      double sigma = iqr > 0.0 ? (sd < iqr ? sd : iqr) : sd;
      const double fac = pow((double)n * 3.0 / 4.0, -0.2);
      if (sigma > 0.0) {
        bw = sigma * fac;
      } else {
        const double iqr2 = (pct(99.0) - pct(1.0)) / 4.6526957480816815;
        bw = iqr2 > 0.0 ? iqr2 * fac : 1.0;
      }
    }
    const float bwf = (float)(bw * (double)scale);
    bw_out[0] = bwf;
    if (inv_var_out)
      for (int d = 0; d < dp; ++d) inv_var_out[d] = 1.0f / (bwf * bwf);
  }
}
}  // namespace dust

extern "C" int dust_silverman_bandwidth(const float* x, int32_t n, float scale, float* bw_out, float* inv_var_out, int32_t dp,
                                        void* stream_) {
  DUST_REQUIRE(x && bw_out && n > 0 && n <= 4096 && dp >= 0, DUST_ERR_INVALID_ARG,
               "dust_silverman_bandwidth: x, bw_out and 1 <= n <= 4096 are required (got n=%d)", n);
  int npad = 2;
  while (npad < n) npad <<= 1;
  cudaStream_t stream = (cudaStream_t)stream_;
  { DUST_TIMED("silverman_kernel", stream); silverman_kernel<<<1, npad < 1024 ? (npad < 32 ? 32 : npad) : 1024, sizeof(float) * npad, stream>>>(x, n, npad, scale, bw_out, inv_var_out, dp); }
  DUST_LAUNCH_OK("silverman_kernel");
  return DUST_OK;
}

extern "C" size_t dust_mpf_workspace_bytes(const dust_mpf_args* a) {
  if (a == nullptr || a->model == nullptr) return 0;
  const int g = mpf_coop_grid(a);
  return g ? sizeof(float) * ((size_t)2 * a->Np * model_dp(a->model->kind) + g) : 0;
}

extern "C" int dust_mpf_optimize(const dust_mpf_args* a, void* stream_) {
  DUST_REQUIRE(a != nullptr, DUST_ERR_INVALID_ARG, "dust_mpf_optimize: args is NULL");
  int rc = validate_model(a->model);
  if (rc) return rc;
  DUST_REQUIRE(a->B > 0 && a->Np > 0 && a->n_steps >= 0, DUST_ERR_INVALID_ARG, "dust_mpf_optimize: sizes must be positive");
  DUST_REQUIRE(a->x && a->obs0 && a->action && a->obs1 && a->prior_inv_var, DUST_ERR_INVALID_ARG,
               "dust_mpf_optimize: x, obs0, action, obs1, prior_inv_var are required");
  DUST_REQUIRE(a->obs_std > 0.f && (a->bw > 0.f || a->bw_dev), DUST_ERR_INVALID_ARG, "dust_mpf_optimize: obs_std and bw must be positive");
  const int kind = a->model->kind, dp = model_dp(kind);
  const size_t smem = sizeof(float) * 3 * (size_t)a->Np * dp;
  DUST_REQUIRE(smem <= 200 * 1024, DUST_ERR_UNSUPPORTED, "dust_mpf_optimize: Np=%d too large for one CTA", a->Np);
  MpfKParams k;
  k.m = to_params(*a->model);
  k.B = a->B; k.Np = a->Np; k.dp = dp; k.n_steps = a->n_steps; k.log_space = a->log_space;
  k.x = a->x; k.obs0 = a->obs0; k.action = a->action; k.obs1 = a->obs1; k.prior_inv_var = a->prior_inv_var;
  k.inv_obs_var = 1.0f / (a->obs_std * a->obs_std); k.bw = a->bw; k.lr = a->lr; k.grad_norms = a->grad_norms;
  k.bw_dev = a->bw_dev; k.lanes = 32; k.phi_out = a->phi_out;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int coop = mpf_coop_grid(a);
  if (coop && !a->phi_out && a->workspace && a->workspace_bytes >= dust_mpf_workspace_bytes(a) && a->n_steps > 0) {
    float* ws = (float*)a->workspace;
    void* kargs[] = {(void*)&k, (void*)&ws};
    const void* fn = kind == DUST_MODEL_PENDULUM ? (const void*)mpf_coop_kernel<DUST_MODEL_PENDULUM>
                                                 : (const void*)mpf_coop_kernel<DUST_MODEL_PARTICLE>;
    { DUST_TIMED("mpf_coop_kernel", stream); DUST_CUDA_OK(cudaLaunchCooperativeKernel(fn, dim3(coop), dim3(256), kargs, 0, stream)); }
    DUST_LAUNCH_OK("mpf_coop_kernel");
    return DUST_OK;
  }
  // a warp per particle: as many warps as particles, up to a full CTA; fewer when many instances share the GPU
  int kMpfThreads = 32 * (a->Np < 32 ? a->Np : 32);
  if ((long long)a->B * kMpfThreads > (long long)kNumSMs * 2048) kMpfThreads = 256;
  if (kMpfThreads < 64) kMpfThreads = 64;
  // 33..64 particles of few instances (the demos' 50): 16 lanes per particle cover them all in ONE pass of a
  // 1024-thread CTA instead of two passes of 32 warps (each pass is latency bound: model step + Jacobian per particle)
  if (a->Np > 32 && a->Np <= 64 && kMpfThreads == 1024) k.lanes = 16;
  if (kind == DUST_MODEL_PENDULUM) {
    if (smem > 48 * 1024) DUST_CUDA_OK(cudaFuncSetAttribute(mpf_kernel<DUST_MODEL_PENDULUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { DUST_TIMED("mpf_kernel", stream); mpf_kernel<DUST_MODEL_PENDULUM><<<a->B, kMpfThreads, smem, stream>>>(k); }
  } else {
    if (smem > 48 * 1024) DUST_CUDA_OK(cudaFuncSetAttribute(mpf_kernel<DUST_MODEL_PARTICLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { DUST_TIMED("mpf_kernel", stream); mpf_kernel<DUST_MODEL_PARTICLE><<<a->B, kMpfThreads, smem, stream>>>(k); }
  }
  DUST_LAUNCH_OK("mpf_kernel");
  return DUST_OK;
}

extern "C" int dust_model_step(const dust_model_desc* model, int32_t M, const float* states, const float* actions,
                               const float* params, float* next_states, void* stream_) {
  int rc = validate_model(model);
  if (rc) return rc;
  DUST_REQUIRE(M > 0 && states && actions && next_states, DUST_ERR_INVALID_ARG, "dust_model_step: bad arguments");
  const ModelParams m = to_params(*model);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (model->kind == DUST_MODEL_PENDULUM)
    { DUST_TIMED("model_step_kernel", stream); model_step_kernel<DUST_MODEL_PENDULUM><<<ceil_div(M, 128), 128, 0, stream>>>(m, M, states, actions, params, next_states); }
  else
    { DUST_TIMED("model_step_kernel", stream); model_step_kernel<DUST_MODEL_PARTICLE><<<ceil_div(M, 128), 128, 0, stream>>>(m, M, states, actions, params, next_states); }
  DUST_LAUNCH_OK("model_step_kernel");
  return DUST_OK;
}

extern "C" int dust_model_cost(const dust_model_desc* model, int32_t M, int32_t terminal, const float* states,
                               const float* actions, float* costs, void* stream_) {
  int rc = validate_model(model);
  if (rc) return rc;
  DUST_REQUIRE(M > 0 && states && costs, DUST_ERR_INVALID_ARG, "dust_model_cost: bad arguments");
  const ModelParams m = to_params(*model);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (model->kind == DUST_MODEL_PENDULUM)
    { DUST_TIMED("model_cost_kernel", stream); model_cost_kernel<DUST_MODEL_PENDULUM><<<ceil_div(M, 128), 128, 0, stream>>>(m, M, terminal, states, actions, costs); }
  else
    { DUST_TIMED("model_cost_kernel", stream); model_cost_kernel<DUST_MODEL_PARTICLE><<<ceil_div(M, 128), 128, 0, stream>>>(m, M, terminal, states, actions, costs); }
  DUST_LAUNCH_OK("model_cost_kernel");
  return DUST_OK;
}
