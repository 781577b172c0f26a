// Small-N side of the SVGD update (N = number of control policies, typically 3..64, or MPF
// particles up to ~1k): GMM prior, fused phi (+SGD step), post-step weights / argmax / shift,
// stand-alone controller step.  One warp per particle row; lanes stride over the flattened
// dimension D = H*A.  Large N goes through svgd_large.cu.
//
// Reference: dust/inference/svgd.py:84-99,127-135; dust/inference/svmpc.py:41,62-85,128-200;
// dust/kernels/base_kernels.py:53-108; dust/kernels/composite_kernels.py:33-64;
// dust/controllers/disco.py:396-417.
#include <math.h>

#include "common.cuh"
#include "svgd_dev.cuh"

namespace dust {

struct GmmKParams {
  int B, M, K, D;
  const float *x, *mu, *mix, *inv_var;
  float log_norm;
  float *log_prob, *score;
};

constexpr int kGmmWarps = 4;

__global__ void __launch_bounds__(kGmmWarps * 32) gmm_kernel(const GmmKParams k) {
  extern __shared__ float sm[];
  float* logmix = sm;                  // [K]
  float* logits = sm + k.K;            // [warps][K]
  const long long inst = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) warp_log_mix(k.mix ? k.mix + inst * k.K : nullptr, k.K, logmix);
  __syncthreads();
  const int i = blockIdx.x * kGmmWarps + warp;
  if (i >= k.M) return;
  const float* x = k.x + (inst * k.M + i) * (long long)k.D;
  float xr[kMaxDPerLane], sr[kMaxDPerLane];
#pragma unroll
  for (int q = 0; q < kMaxDPerLane; ++q) {
    const int d = lane + 32 * q;
    xr[q] = d < k.D ? x[d] : 0.f;
  }
  const float lp = warp_gmm_point(xr, k.mu + inst * (long long)k.K * k.D, logmix, k.inv_var, k.K, k.D,
                                  logits + warp * k.K, k.score ? sr : nullptr);
  if (k.log_prob && lane == 0) k.log_prob[inst * k.M + i] = lp + k.log_norm;
  if (k.score) {
    float* so = k.score + (inst * k.M + i) * (long long)k.D;
#pragma unroll
    for (int q = 0; q < kMaxDPerLane; ++q) {
      const int d = lane + 32 * q;
      if (d < k.D) so[d] = sr[q];
    }
  }
}

// ---------------------------------------------------------------------------------------
// phi, small N: one warp per row i, direct-difference distances (exact zero diagonal)
// ---------------------------------------------------------------------------------------
struct PhiKParams {
  int B, N, D, row_begin, row_end;
  const float *x, *score;
  float gamma, c1, c2;
  const float* gamma_dev;
  float lr;
  float *phi, *x_out;
  const float* h;  // per-dim bandwidths [B, D] (per_dim mode)
};

constexpr int kPhiWarps = 4;

template <bool PER_DIM>
__global__ void __launch_bounds__(kPhiWarps * 32) phi_small_kernel(const PhiKParams k) {
  const long long inst = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = k.row_begin + blockIdx.x * kPhiWarps + warp;
  if (i >= k.row_end) return;
  float gamma = k.gamma, c1 = k.c1, c2 = k.c2;
  if (!PER_DIM && k.gamma_dev) {
    gamma = k.gamma_dev[0];
    c1 = k.gamma_dev[1];
    c2 = k.gamma_dev[2];
  }
  const float* X = k.x + inst * (long long)k.N * k.D;
  const float* Sc = k.score + inst * (long long)k.N * k.D;
  float xi[kMaxDPerLane], acc[kMaxDPerLane], acc2[kMaxDPerLane], hq[kMaxDPerLane], xxi[kMaxDPerLane];
#pragma unroll
  for (int q = 0; q < kMaxDPerLane; ++q) {
    const int d = lane + 32 * q;
    xi[q] = d < k.D ? X[(long long)i * k.D + d] : 0.f;
    acc[q] = 0.f;
    acc2[q] = 0.f;
    hq[q] = (PER_DIM && d < k.D) ? k.h[inst * k.D + d] : 1.f;
    xxi[q] = __fmul_rn(xi[q], xi[q]);
  }
  for (int j = 0; j < k.N; ++j) {
    float xj[kMaxDPerLane], sj[kMaxDPerLane];
    float part = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxDPerLane; ++q) {
      const int d = lane + 32 * q;
      if (d < k.D) {
        xj[q] = __ldg(X + (long long)j * k.D + d);
        sj[q] = __ldg(Sc + (long long)j * k.D + d);
        const float df = xi[q] - xj[q];
        part += df * df;
      } else {
        xj[q] = 0.f;
        sj[q] = 0.f;
      }
    }
    if (!PER_DIM) {
      const float d2 = warp_sum(part);
      const float kij = expf(-gamma * d2);
#pragma unroll
      for (int q = 0; q < kMaxDPerLane; ++q) acc[q] += kij * (c1 * sj[q] + c2 * (xi[q] - xj[q]));
    } else {
      // composite_kernels.py:47-55 -> base_kernels.py:58-62,99-100 on one column:
      // d2 = ((-2 x_i y_j) + x_i^2) + y_j^2 (unclamped), K = exp(-d2/h), dK = ((K (x_i-y_j)) 2)/h
#pragma unroll
      for (int q = 0; q < kMaxDPerLane; ++q) {
        const float xy = __fmul_rn(xi[q], xj[q]);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, xy), xxi[q]), __fmul_rn(xj[q], xj[q]));
        const float kij = expf(-d2 / hq[q]);
        acc[q] += kij * sj[q];
        acc2[q] += ((kij * (xi[q] - xj[q])) * 2.0f) / hq[q];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < kMaxDPerLane; ++q) {
    const int d = lane + 32 * q;
    if (d < k.D) {
      const float ph = PER_DIM ? (acc[q] / (float)k.N + acc2[q] / (float)k.N) : acc[q];
      const long long o = (inst * k.N + i) * (long long)k.D + d;
      if (k.phi) k.phi[o] = ph;
      if (k.x_out) k.x_out[o] = xi[q] + k.lr * ph;
    }
  }
}

// per-dimension lower-median bandwidth for the message-passing kernel: one CTA per (dim, inst);
// the N^2 values are sorted with a shared-memory bitonic network.
struct DimBwKParams {
  int B, N, D, n2, n2pad;
  const float* x;
  float scale;
  float* h;  // [B,D]
};

__global__ void __launch_bounds__(256) dim_bandwidth_kernel(const DimBwKParams k) {
  extern __shared__ float vals[];  // [n2pad]
  const int d = blockIdx.x;
  const long long inst = blockIdx.y;
  const float* X = k.x + inst * (long long)k.N * k.D;
  for (int e = threadIdx.x; e < k.n2pad; e += blockDim.x) {
    float v = INFINITY;
    if (e < k.n2) {
      const int i = e / k.N, j = e - i * k.N;
      const float xi = X[(long long)i * k.D + d], xj = X[(long long)j * k.D + d];
      v = __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, __fmul_rn(xi, xj)), __fmul_rn(xi, xi)), __fmul_rn(xj, xj));
    }
    vals[e] = v;
  }
  __syncthreads();
  for (int size = 2; size <= k.n2pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (k.n2pad >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const float a = vals[lo], b = vals[hi];
        if ((a > b) == up) {
          vals[lo] = b;
          vals[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    const float med = vals[(k.n2 - 1) >> 1];  // torch.median: lower median
    float h = med / (float)log((double)k.N + 1.0);
    h = k.scale * h;
    h = fmaxf(h, 1e-5f);
    k.h[inst * k.D + d] = h;
  }
}

__global__ void bandwidth_from_median_kernel(const float* median, int N, float scale, int mode, float* out) {
  const float med = *median;
  const float logn = (float)log((double)N + 1.0);
  if (mode == 0) {
    // svgd.py:51-52: bw = scale * max(sqrt(0.5 med)/log(N+1), tol);  K = exp(-d2/bw^2/2)
    float h = sqrtf(0.5f * med) / logn;
    h = fmaxf(h, 1e-5f);
    const float bw = scale * h;
    out[0] = 1.0f / (2.0f * bw * bw);
    out[1] = 1.0f / (float)N;
    out[2] = 1.0f / ((float)N * bw * bw);
    out[3] = bw;
  } else {
    // base_kernels.py:64-100: h = clamp(scale * med/log(N+1), tol);  K = exp(-d2/h)
    float h = fmaxf(scale * (med / logn), 1e-5f);
    out[0] = 1.0f / h;
    out[1] = 1.0f / (float)N;
    out[2] = 2.0f / ((float)N * h);
    out[3] = h;
  }
}

// ---------------------------------------------------------------------------------------
// SVMPC.forward: weights, argmax, shift, prior refresh   (svmpc.py:128-200)
// ---------------------------------------------------------------------------------------
struct FwdKParams {
  int B, N, H, A, D, roll, weighted;
  const float *log_lik, *theta, *mu, *mix, *inv_var;
  float log_norm;
  float* p_weights;
  int* i_star;
  float *a_seq, *theta_next, *mix_next;
  const float* resample_noise;
};

constexpr int kFwdWarps = 4;        // many instances: small CTAs, several per SM
constexpr int kFwdWarpsWide = 32;   // a few instances: a warp per policy

__global__ void __launch_bounds__(kFwdWarpsWide * 32) svmpc_forward_kernel(const FwdKParams k) {
  extern __shared__ float sm[];
  float* logmix = sm;                       // [N]
  float* logw = sm + k.N;                   // [N]
  float* logits = sm + 2 * k.N;             // [warps][N]
  __shared__ int s_istar;
  const long long inst = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* theta = k.theta + inst * (long long)k.N * k.D;
  const float* mu = k.mu + inst * (long long)k.N * k.D;
  if (warp == 0) warp_log_mix(k.mix ? k.mix + inst * k.N : nullptr, k.N, logmix);
  __syncthreads();
  for (int n = warp; n < k.N; n += (int)(blockDim.x >> 5)) {
    float xr[kMaxDPerLane];
#pragma unroll
    for (int q = 0; q < kMaxDPerLane; ++q) {
      const int d = lane + 32 * q;
      xr[q] = d < k.D ? theta[(long long)n * k.D + d] : 0.f;
    }
    const float lp = warp_gmm_point(xr, mu, logmix, k.inv_var, k.N, k.D, logits + warp * k.N, nullptr);
    if (lane == 0) logw[n] = k.log_lik[inst * k.N + n] + (lp + k.log_norm);
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int n = lane; n < k.N; n += 32) mx = fmaxf(mx, logw[n]);
    mx = warp_max(mx);
    float z = 0.f;
    for (int n = lane; n < k.N; n += 32) z += expf(logw[n] - mx);
    z = warp_sum(z);
    const float lse = mx + logf(z);
    // p = exp(log_w - logsumexp); argmax over p, first maximum wins (svmpc.py:140,192)
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int n = lane; n < k.N; n += 32) {
      const float p = expf(logw[n] - lse);
      k.p_weights[inst * k.N + n] = p;
      if (k.mix_next) k.mix_next[inst * k.N + n] = k.weighted ? p : 1.0f;
      if (p > best) { best = p; bi = n; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      s_istar = bi;
      if (k.i_star) k.i_star[inst] = bi;
    }
  }
  __syncthreads();
  const int is = s_istar;
  if (k.a_seq)
    for (int d = threadIdx.x; d < k.D; d += blockDim.x) k.a_seq[inst * k.D + d] = theta[(long long)is * k.D + d];
  if (k.theta_next) {
    float* out = k.theta_next + inst * (long long)k.N * k.D;
    const int HA_shift = (k.H - 1) * k.A;
    for (int e = threadIdx.x; e < k.N * k.D; e += blockDim.x) {
      const int n = e / k.D, d = e - n * k.D;
      float v;
      if (d < HA_shift) {
        v = theta[(long long)n * k.D + d + k.A];  // theta.roll(-1, dims=-2)
      } else if (k.roll == DUST_ROLL_REPEAT) {
        v = theta[(long long)n * k.D + d];        // last step repeated (svmpc.py:145-147)
      } else if (k.roll == DUST_ROLL_RESAMPLE) {
        // last step of a sample of the current prior (svmpc.py:148-150): component by inverse CDF
        // over the (clamped, normalised) mixture weights, then centre + sigma_prior * z
        const int a = d - HA_shift;
        const float* z = k.resample_noise + (inst * k.N + n) * (long long)(k.A + 1);
        const float u = 0.5f * erfcf(-z[k.A] * 0.70710678118654752f);
        int comp = k.N - 1;
        float cum = 0.f;
        for (int c = 0; c < k.N; ++c) {
          cum += expf(logmix[c]);
          if (u <= cum) { comp = c; break; }
        }
        v = mu[(long long)comp * k.D + d] + rsqrtf(k.inv_var[d]) * z[a];
      } else {
        const int a = d - HA_shift;               // mean over the horizon (svmpc.py:151-153)
        float s = 0.f;
        for (int h = 0; h < k.H; ++h) s += theta[(long long)n * k.D + h * k.A + a];
        v = s / (float)k.H;
      }
      out[e] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// MultiDISCO.step   (disco.py:396-417)
// ---------------------------------------------------------------------------------------
struct StepKParams {
  int B, N, H, A, D, strategy, steps;
  const float *a_low, *a_high, *a_mix;
  float *a_mat, *a_seq, *next_actions;
};

__global__ void __launch_bounds__(128) disco_step_kernel(const StepKParams k) {
  extern __shared__ float sm[];  // [N*D] copy of a_mat, then [D] a_seq
  float* mat = sm;
  float* seq = sm + k.N * k.D;
  __shared__ int s_best;
  const long long inst = blockIdx.x;
  float* a_mat = k.a_mat + inst * (long long)k.N * k.D;
  const float* mixw = k.a_mix + inst * k.N;
  for (int e = threadIdx.x; e < k.N * k.D; e += blockDim.x) mat[e] = a_mat[e];
  if (threadIdx.x == 0) {
    int bi = 0;
    float best = mixw[0];
    for (int n = 1; n < k.N; ++n)
      if (mixw[n] > best) { best = mixw[n]; bi = n; }
    s_best = bi;
  }
  __syncthreads();
  const int bi = s_best;
  for (int d = threadIdx.x; d < k.D; d += blockDim.x) {
    const int a = d % k.A;
    float v;
    if (k.strategy == DUST_SELECT_ARGMAX) {
      v = mat[bi * k.D + d];
    } else {
      v = 0.f;
      for (int n = 0; n < k.N; ++n) v += mat[n * k.D + d] * mixw[n];
    }
    v = fminf(fmaxf(v, k.a_low[a]), k.a_high[a]);
    seq[d] = v;
    if (k.strategy == DUST_SELECT_ARGMAX) mat[bi * k.D + d] = v;  // a_seq is a view of a_mat[i*]
  }
  __syncthreads();
  const int shift = k.steps * k.A;
  if (k.next_actions)
    for (int d = threadIdx.x; d < shift; d += blockDim.x) k.next_actions[inst * shift + d] = seq[d];
  for (int d = threadIdx.x; d < k.D; d += blockDim.x)
    k.a_seq[inst * k.D + d] = (d + shift < k.D) ? seq[d + shift] : 0.f;
  for (int e = threadIdx.x; e < k.N * k.D; e += blockDim.x) {
    const int n = e / k.D, d = e - n * k.D;
    a_mat[e] = (d + shift < k.D) ? mat[n * k.D + d + shift] : 0.f;
  }
}

}  // namespace dust

using namespace dust;

extern "C" int dust_gmm_score(const dust_gmm_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DUST_REQUIRE(a != nullptr, DUST_ERR_INVALID_ARG, "dust_gmm_score: args is NULL");
  DUST_REQUIRE(a->B > 0 && a->M > 0 && a->K > 0 && a->D > 0, DUST_ERR_INVALID_ARG, "dust_gmm_score: sizes must be positive");
  DUST_REQUIRE(a->x && a->mu && a->inv_var, DUST_ERR_INVALID_ARG, "dust_gmm_score: x, mu and inv_var are required");
  DUST_REQUIRE(a->D <= 32 * kMaxDPerLane, DUST_ERR_UNSUPPORTED, "dust_gmm_score: D=%d > %d", a->D, 32 * kMaxDPerLane);
  DUST_REQUIRE(a->B <= 65535, DUST_ERR_UNSUPPORTED, "dust_gmm_score: B > 65535");
  const size_t smem = sizeof(float) * a->K * (1 + kGmmWarps);
  DUST_REQUIRE(smem <= 48 * 1024, DUST_ERR_UNSUPPORTED, "dust_gmm_score: K=%d too large", a->K);
  GmmKParams k{a->B, a->M, a->K, a->D, a->x, a->mu, a->mix, a->inv_var, a->log_norm, a->log_prob, a->score};
  dim3 grid((unsigned)ceil_div(a->M, kGmmWarps), (unsigned)a->B, 1);
  { DUST_TIMED("gmm_kernel", stream); gmm_kernel<<<grid, kGmmWarps * 32, smem, stream>>>(k); }
  DUST_LAUNCH_OK("gmm_kernel");
  return DUST_OK;
}

namespace dust {
int phi_large(const dust_phi_args* a, cudaStream_t stream);      // svgd_large.cu
size_t phi_large_workspace(const dust_phi_args* a);
constexpr int kSmallPhiMaxN = 512;
static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
}  // namespace dust

extern "C" size_t dust_phi_workspace_bytes(const dust_phi_args* a) {
  if (!a || a->B <= 0 || a->N <= 0 || a->D <= 0) return 0;
  if (a->per_dim) return sizeof(float) * (size_t)a->B * a->D;
  if (a->N > kSmallPhiMaxN || a->D > 32 * kMaxDPerLane) return phi_large_workspace(a);
  return 0;
}

extern "C" int dust_svgd_phi(const dust_phi_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DUST_REQUIRE(a != nullptr, DUST_ERR_INVALID_ARG, "dust_svgd_phi: args is NULL");
  DUST_REQUIRE(a->B > 0 && a->N > 0 && a->D > 0, DUST_ERR_INVALID_ARG, "dust_svgd_phi: sizes must be positive");
  DUST_REQUIRE(a->x && a->score, DUST_ERR_INVALID_ARG, "dust_svgd_phi: x and score are required");
  DUST_REQUIRE(a->phi || a->x_out, DUST_ERR_INVALID_ARG, "dust_svgd_phi: no output requested");
  const int r0 = a->row_begin, r1 = a->row_end > 0 ? a->row_end : a->N;
  DUST_REQUIRE(r0 >= 0 && r1 <= a->N && r0 < r1, DUST_ERR_INVALID_ARG, "dust_svgd_phi: bad row range [%d,%d)", r0, r1);
  DUST_REQUIRE(a->x_out != a->x, DUST_ERR_INVALID_ARG,
               "dust_svgd_phi: x_out must not alias x (rows are read by other warps)");
  DUST_REQUIRE(a->ld == 0 || a->ld >= a->D, DUST_ERR_INVALID_ARG, "dust_svgd_phi: ld=%d < D=%d", a->ld, a->D);
  if (!a->per_dim && (a->N > kSmallPhiMaxN || a->D > 32 * kMaxDPerLane)) return phi_large(a, stream);
  DUST_REQUIRE(a->ld == 0 || a->ld == a->D, DUST_ERR_UNSUPPORTED, "dust_svgd_phi: a row stride (ld=%d) needs the tensor-core path", a->ld);
  DUST_REQUIRE(a->D <= 32 * kMaxDPerLane, DUST_ERR_UNSUPPORTED, "dust_svgd_phi: D=%d > %d in per_dim mode", a->D, 32 * kMaxDPerLane);
  DUST_REQUIRE(a->B <= 65535, DUST_ERR_UNSUPPORTED, "dust_svgd_phi: B > 65535");
  PhiKParams k{a->B, a->N, a->D, r0, r1, a->x, a->score, a->gamma, a->c1, a->c2, a->gamma_dev, a->lr, a->phi, a->x_out, nullptr};
  dim3 grid((unsigned)ceil_div(r1 - r0, kPhiWarps), (unsigned)a->B, 1);
  if (a->per_dim) {
    const int n2 = a->N * a->N, n2pad = next_pow2(n2);
    DUST_REQUIRE(a->N <= 96, DUST_ERR_UNSUPPORTED, "dust_svgd_phi: per_dim kernel supports N <= 96 (got %d)", a->N);
    float* h = a->bandwidths;
    if (!h) {
      DUST_REQUIRE(a->workspace && a->workspace_bytes >= sizeof(float) * (size_t)a->B * a->D, DUST_ERR_WORKSPACE,
                   "dust_svgd_phi: per_dim needs B*D floats of workspace");
      h = (float*)a->workspace;
    }
    DimBwKParams bk{a->B, a->N, a->D, n2, n2pad, a->x, a->bw_scale, h};
    const size_t smem = sizeof(float) * n2pad;
    if (smem > 48 * 1024) DUST_CUDA_OK(cudaFuncSetAttribute(dim_bandwidth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { DUST_TIMED("dim_bandwidth_kernel", stream); dim_bandwidth_kernel<<<dim3((unsigned)a->D, (unsigned)a->B, 1), 256, smem, stream>>>(bk); }
    DUST_LAUNCH_OK("dim_bandwidth_kernel");
    k.h = h;
    { DUST_TIMED("phi_small_kernel", stream); phi_small_kernel<true><<<grid, kPhiWarps * 32, 0, stream>>>(k); }
  } else {
    { DUST_TIMED("phi_small_kernel", stream); phi_small_kernel<false><<<grid, kPhiWarps * 32, 0, stream>>>(k); }
  }
  DUST_LAUNCH_OK("phi_small_kernel");
  return DUST_OK;
}

extern "C" int dust_bandwidth_from_median(const float* median, int32_t N, float scale, int32_t mode, float* out,
                                          void* stream_) {
  DUST_REQUIRE(median && out && N > 0, DUST_ERR_INVALID_ARG, "dust_bandwidth_from_median: bad arguments");
  DUST_REQUIRE(mode == 0 || mode == 1, DUST_ERR_INVALID_ARG, "dust_bandwidth_from_median: mode must be 0 or 1");
  { DUST_TIMED("bandwidth_from_median_kernel", (cudaStream_t)stream_); bandwidth_from_median_kernel<<<1, 1, 0, (cudaStream_t)stream_>>>(median, N, scale, mode, out); }
  DUST_LAUNCH_OK("bandwidth_from_median_kernel");
  return DUST_OK;
}

extern "C" int dust_svmpc_forward(const dust_svmpc_forward_args* a, void* stream_) {
  DUST_REQUIRE(a != nullptr, DUST_ERR_INVALID_ARG, "dust_svmpc_forward: args is NULL");
  DUST_REQUIRE(a->B > 0 && a->N > 0 && a->H > 0 && a->A > 0, DUST_ERR_INVALID_ARG, "dust_svmpc_forward: sizes must be positive");
  DUST_REQUIRE(a->log_lik && a->theta && a->mu && a->inv_var && a->p_weights, DUST_ERR_INVALID_ARG,
               "dust_svmpc_forward: log_lik, theta, mu, inv_var, p_weights are required");
  DUST_REQUIRE(a->theta_next != a->theta, DUST_ERR_INVALID_ARG, "dust_svmpc_forward: theta_next must not alias theta");
  DUST_REQUIRE(a->roll_strategy != DUST_ROLL_RESAMPLE || a->resample_noise != nullptr, DUST_ERR_INVALID_ARG,
               "dust_svmpc_forward: roll strategy 'resample' needs resample_noise [B,N,A+1]");
  DUST_REQUIRE(a->roll_strategy == DUST_ROLL_REPEAT || a->roll_strategy == DUST_ROLL_MEAN || a->roll_strategy == DUST_ROLL_RESAMPLE, DUST_ERR_INVALID_ARG,
               "dust_svmpc_forward: invalid roll strategy %d", a->roll_strategy);
  const int D = a->H * a->A;
  DUST_REQUIRE(D <= 32 * kMaxDPerLane, DUST_ERR_UNSUPPORTED, "dust_svmpc_forward: H*A=%d > %d", D, 32 * kMaxDPerLane);
  // few instances: as many warps as policies (the CTA is alone on its SM anyway)
  int warps = kFwdWarps;
  if (a->B < kNumSMs) { warps = a->N < kFwdWarpsWide ? a->N : kFwdWarpsWide; if (warps < kFwdWarps) warps = kFwdWarps; }
  const size_t smem = sizeof(float) * a->N * (2 + warps);
  DUST_REQUIRE(smem <= 48 * 1024, DUST_ERR_UNSUPPORTED, "dust_svmpc_forward: N=%d too large", a->N);
  FwdKParams k{a->B, a->N, a->H, a->A, D, a->roll_strategy, a->weighted_prior, a->log_lik, a->theta, a->mu, a->mix,
               a->inv_var, a->log_norm, a->p_weights, a->i_star, a->a_seq, a->theta_next, a->mix_next, a->resample_noise};
  { DUST_TIMED("svmpc_forward_kernel", (cudaStream_t)stream_); svmpc_forward_kernel<<<a->B, warps * 32, smem, (cudaStream_t)stream_>>>(k); }
  DUST_LAUNCH_OK("svmpc_forward_kernel");
  return DUST_OK;
}

extern "C" int dust_disco_step(const dust_disco_step_args* a, void* stream_) {
  DUST_REQUIRE(a != nullptr, DUST_ERR_INVALID_ARG, "dust_disco_step: args is NULL");
  DUST_REQUIRE(a->B > 0 && a->N > 0 && a->H > 0 && a->A > 0, DUST_ERR_INVALID_ARG, "dust_disco_step: sizes must be positive");
  DUST_REQUIRE(a->strategy == DUST_SELECT_ARGMAX || a->strategy == DUST_SELECT_AVERAGE, DUST_ERR_INVALID_ARG,
               "Invalid value for strategy.");
  DUST_REQUIRE(a->steps >= 1 && a->steps <= a->H, DUST_ERR_INVALID_ARG, "dust_disco_step: steps out of range");
  DUST_REQUIRE(a->a_low && a->a_high && a->a_mat && a->a_mix && a->a_seq, DUST_ERR_INVALID_ARG,
               "dust_disco_step: a_low, a_high, a_mat, a_mix, a_seq are required");
  const int D = a->H * a->A;
  const size_t smem = sizeof(float) * ((size_t)a->N * D + D);
  DUST_REQUIRE(smem <= 200 * 1024, DUST_ERR_UNSUPPORTED, "dust_disco_step: N*H*A too large");
  if (smem > 48 * 1024) DUST_CUDA_OK(cudaFuncSetAttribute(disco_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  StepKParams k{a->B, a->N, a->H, a->A, D, a->strategy, a->steps, a->a_low, a->a_high, a->a_mix, a->a_mat, a->a_seq, a->next_actions};
  { DUST_TIMED("disco_step_kernel", (cudaStream_t)stream_); disco_step_kernel<<<a->B, 128, smem, (cudaStream_t)stream_>>>(k); }
  DUST_LAUNCH_OK("disco_step_kernel");
  return DUST_OK;
}
