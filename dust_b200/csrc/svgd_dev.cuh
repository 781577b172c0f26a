// Warp-level device functions shared by the small-N SVGD kernels (svgd_small.cu) and the fused
// control-step tail of the instance kernel (rollout.cu): GMM mixture logits, log-density and score
// of one point handled by one warp (lanes stride over the flattened dimension D <= 256).
// Reference: dust/inference/svgd.py:84-89 (get_gmm), dust/inference/svmpc.py:41,138.
#pragma once
#include <math.h>

#include "common.cuh"

namespace dust {

constexpr int kMaxDPerLane = 8;  // D <= 256
constexpr float kF32Eps = 1.1920928955078125e-07f;

// Categorical(probs=mix).logits as MixtureSameFamily uses them (normalise, clamp to
// [eps, 1-eps], log, log_softmax).  Called by one warp; result in out[K] (shared memory).
__device__ __forceinline__ void warp_log_mix(const float* __restrict__ mix, int K, float* out) {
  const int lane = threadIdx.x & 31;
  if (mix == nullptr) {
    const float v = -logf((float)K);
    for (int k = lane; k < K; k += 32) out[k] = v;
    return;
  }
  float tot = 0.f;
  for (int k = lane; k < K; k += 32) tot += mix[k];
  tot = warp_sum(tot);
  float mx = -INFINITY;
  for (int k = lane; k < K; k += 32) {
    float p = mix[k] / tot;
    p = fminf(fmaxf(p, kF32Eps), 1.0f - kF32Eps);
    const float l = logf(p);
    out[k] = l;
    mx = fmaxf(mx, l);
  }
  mx = warp_max(mx);
  float z = 0.f;
  for (int k = lane; k < K; k += 32) z += expf(out[k] - mx);
  z = warp_sum(z);
  const float lse = mx + logf(z);
  for (int k = lane; k < K; k += 32) out[k] = out[k] - lse;
}

// log GMM(x; mu, logmix, inv_var) and (optionally) its score for ONE point handled by a warp.
// xr: this lane's slice of x (d = lane + 32 q).  logits: per-warp scratch [K] in shared memory.
template <int QMAX = kMaxDPerLane>
__device__ __forceinline__ float warp_gmm_point(const float* xr, const float* __restrict__ mu, const float* logmix,
                                const float* __restrict__ inv_var, int K, int D, float* logits, float* score_r,
                                                int mu_stride = 0) {
  if (mu_stride == 0) mu_stride = D;
  const int lane = threadIdx.x & 31;
  float iv[QMAX];
#pragma unroll
  for (int q = 0; q < QMAX; ++q) {
    const int d = lane + 32 * q;
    iv[q] = d < D ? inv_var[d] : 0.f;
  }
  float mx = -INFINITY;
  for (int k = 0; k < K; ++k) {
    float part = 0.f;
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
      const int d = lane + 32 * q;
      if (d < D) {
        const float df = xr[q] - mu[(long long)k * mu_stride + d];
        part += df * df * iv[q];
      }
    }
    part = warp_sum(part);
    const float l = logmix[k] - 0.5f * part;
    if (lane == 0) logits[k] = l;
    mx = fmaxf(mx, l);
  }
  __syncwarp();
  float z = 0.f;
  float acc[QMAX];
#pragma unroll
  for (int q = 0; q < QMAX; ++q) acc[q] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float e = expf(logits[k] - mx);
    z += e;
    if (score_r) {
#pragma unroll
      for (int q = 0; q < QMAX; ++q) {
        const int d = lane + 32 * q;
        if (d < D) acc[q] += e * (mu[(long long)k * mu_stride + d] - xr[q]);
      }
    }
  }
  if (score_r) {
    const float invz = 1.f / z;
#pragma unroll
    for (int q = 0; q < QMAX; ++q) score_r[q] = acc[q] * invz * iv[q];
  }
  __syncwarp();
  return mx + logf(z);
}


}  // namespace dust
