// Standard-normal noise for the action samples (the reference draws it with
// MultivariateNormal.rsample inside CostLikelihood.sample, dust/inference/likelihoods.py:97-103,
// and MultiDISCO._sample_actions, dust/controllers/disco.py:211-230).
//
// Counter-based: element block i of a fill with (seed, offset) is Philox4x32-10(counter = (i, offset),
// key = seed) -> 4 uniforms -> 2 Box-Muller pairs -> 4 normals, so a fill is reproducible, order
// independent and can be sharded over ranks by giving every rank its own offset.  The 671 MB noise
// tensor of the batched benchmark is written once at close to the HBM write rate; the library
// generator it replaces (torch's normal_) took as long as the whole control step.
#include "common.cuh"

namespace dust {

__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
  c1 = (uint32_t)p1;
  c3 = (uint32_t)p0;
  c0 = n0;
  c2 = n2;
}

// the ten round keys (key + r * Weyl constants) are formed on the host and passed as kernel
// parameters: they reach the XORs as constant-bank operands, no instruction spent on the schedule
struct PhiloxKeys {
  uint32_t k0[10], k1[10];
};

__device__ __forceinline__ uint4 philox4x32_10(uint64_t index, uint64_t offset, const PhiloxKeys& key) {
  uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, key.k0[r], key.k1[r]);
  return make_uint4(c0, c1, c2, c3);
}

// two normals from two 32-bit words.  The radius uses all 32 bits of `a` (u in (0, 1], tails out to
// 6.7 sigma as in curand / torch); the angle uses 23 bits of `b` placed in the mantissa (no I2F).
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u = fmaf((float)a, 2.3283064365386963e-10f, 2.3283064365386963e-10f);  // (a + 1) * 2^-32, never denormal
  const float rev = __uint_as_float(0x3f800000u | (b >> 9)) - 1.0f;                   // [0, 1) revolutions
  float l2, r, s, c;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l2));     // sqrt(-2 ln u)
  const float ang = 6.283185307179586f * rev;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(ang));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(ang));
  return make_float2(r * c, r * s);
}

__global__ void __launch_bounds__(256) noise_normal_kernel(float* __restrict__ out, long long n, const PhiloxKeys key, uint64_t offset) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint4 w = philox4x32_10((uint64_t)i, offset, key);
    const float2 p = box_muller(w.x, w.y), q = box_muller(w.z, w.w);
    __stcs(reinterpret_cast<float4*>(out) + i, make_float4(p.x, p.y, q.x, q.y));  // streaming store: written once, read once
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 3)) {   // ragged tail: one more block of four
    const uint4 w = philox4x32_10((uint64_t)n4, offset, key);
    const float2 p = box_muller(w.x, w.y), q = box_muller(w.z, w.w);
    const float v[4] = {p.x, p.y, q.x, q.y};
    for (int e = 0; e < (int)(n & 3); ++e) out[(n4 << 2) + e] = v[e];
  }
}

}  // namespace dust

extern "C" int dust_noise_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream_) {
  using namespace dust;
  cudaStream_t stream = (cudaStream_t)stream_;
  DUST_REQUIRE(n >= 0, DUST_ERR_INVALID_ARG, "dust_noise_normal: n = %lld", (long long)n);
  if (n == 0) return DUST_OK;
  DUST_REQUIRE(out != nullptr, DUST_ERR_INVALID_ARG, "dust_noise_normal: null output");
  DUST_REQUIRE((((uintptr_t)out) & 15) == 0, DUST_ERR_INVALID_ARG, "dust_noise_normal: output must be 16-byte aligned");
  const long long n4 = n >> 2;
  const long long want = (n4 + 255) / 256;
  const int grid = (int)(want < 1 ? 1 : (want > (long long)kNumSMs * 8 ? (long long)kNumSMs * 8 : want));
  PhiloxKeys key;
  for (int r = 0; r < 10; ++r) {
    key.k0[r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
    key.k1[r] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
  }
  {
    DUST_TIMED("noise_normal_kernel", stream);
    noise_normal_kernel<<<grid, 256, 0, stream>>>(out, (long long)n, key, offset);
  }
  DUST_LAUNCH_OK("noise_normal_kernel");
  return DUST_OK;
}
