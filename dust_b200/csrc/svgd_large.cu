// Large-N SVGD: exact median of the N^2 squared distances (two-pass radix select on the float32
// bit pattern) and the flash-style phi that never materialises K (SIMT fp32 tiles; the tcgen05
// 3xTF32 kernel in svgd_tc.cu takes over when the shape qualifies).
//
// Reference: squared_distance / bw_median dust/inference/svgd.py:28-52 (d2 = |x|^2+|y|^2-2xy,
// clamped at 0; torch.median = lower median), SVGD.phi dust/inference/svgd.py:127-135.
#include <stdlib.h>

#include "common.cuh"

namespace dust {

constexpr int kLT = 64;        // tile edge (rows and columns)
constexpr int kLThreads = 256; // 16 x 16 threads, 4 x 4 micro-tile each

// squared row norms
__global__ void row_norms_kernel(const float* __restrict__ x, int N, int D, int ld, float* __restrict__ xn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float s = 0.f;
  for (int d = 0; d < D; ++d) { const float v = x[(long long)i * ld + d]; s += v * v; }
  xn[i] = s;
}

// stage a [kLT, D] block of rows (row stride ld floats) transposed into shared memory: dst[d][r]
__device__ __forceinline__ void stage_rows_T(const float* __restrict__ x, int N, int D, int ld, int r0, float* dst) {
  for (int e = threadIdx.x; e < kLT * D; e += kLThreads) {
    const int r = e / D, d = e - r * D;
    const int gi = r0 + r;
    dst[d * kLT + r] = gi < N ? __ldg(x + (long long)gi * ld + d) : 0.f;
  }
}

// 4x4 micro-tile of d2 for rows (i0 + ty*4 + r), cols (j0 + tx*4 + c)
__device__ __forceinline__ void tile_d2(const float* xi_t, const float* xj_t, const float* xn_i, const float* xn_j,
                                        int D, int ty, int tx, float d2[4][4]) {
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  for (int d = 0; d < D; ++d) {
    const float4 a = *reinterpret_cast<const float4*>(xi_t + d * kLT + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(xj_t + d * kLT + tx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      // svgd.py:36-39: (|y|^2 - 2 x.y) + |x|^2, clamped at zero
      const float v = (xn_j[tx * 4 + c] - 2.0f * acc[r][c]) + xn_i[ty * 4 + r];
      d2[r][c] = fmaxf(v, 0.f);
    }
}

// ---------------------------------------------------------------------------------------
// median: histogram passes
// ---------------------------------------------------------------------------------------
struct HistKParams {
  int N, D, ld, row_begin, row_end, pass;
  const float *x, *xn;
  unsigned long long* hist;
  const uint32_t* selected;
};

__global__ void __launch_bounds__(kLThreads) median_hist_kernel(const HistKParams k) {
  extern __shared__ __align__(16) float sm[];
  if (k.selected[5] != 0u) return;  // the tensor-core window pass already found the median
  float* xi_t = sm;                    // [D][64]
  float* xj_t = xi_t + k.D * kLT;      // [D][64]
  float* xn_i = xj_t + k.D * kLT;      // [64]
  float* xn_j = xn_i + kLT;            // [64]
  uint32_t* lh = reinterpret_cast<uint32_t*>(xn_j + kLT);  // pass 0: [32768] local histogram
  const int i0 = k.row_begin + blockIdx.x * kLT;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  if (k.pass == 0)
    for (int e = threadIdx.x; e < 32768; e += kLThreads) lh[e] = 0u;
  stage_rows_T(k.x, k.row_end, k.D, k.ld, i0, xi_t);
  if (threadIdx.x < kLT) xn_i[threadIdx.x] = (i0 + threadIdx.x < k.row_end) ? k.xn[i0 + threadIdx.x] : 0.f;
  const uint32_t sel_hi = k.pass == 1 ? k.selected[0] : 0u;
  for (int j0 = 0; j0 < k.N; j0 += kLT) {
    __syncthreads();
    stage_rows_T(k.x, k.N, k.D, k.ld, j0, xj_t);
    if (threadIdx.x < kLT) xn_j[threadIdx.x] = (j0 + threadIdx.x < k.N) ? k.xn[j0 + threadIdx.x] : 0.f;
    __syncthreads();
    float d2[4][4];
    tile_d2(xi_t, xj_t, xn_i, xn_j, k.D, ty, tx, d2);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (i0 + ty * 4 + r >= k.row_end || j0 + tx * 4 + c >= k.N) continue;
        const uint32_t bits = __float_as_uint(d2[r][c]);
        if (k.pass == 0) {
          atomicAdd(&lh[bits >> 16], 1u);  // non-negative floats: bits >> 16 < 32768
        } else if ((bits >> 16) == sel_hi) {
          atomicAdd(&k.hist[bits & 0xffffu], 1ull);
        }
      }
  }
  if (k.pass == 0) {
    __syncthreads();
    for (int e = threadIdx.x; e < 32768; e += kLThreads)
      if (lh[e]) atomicAdd(&k.hist[e], (unsigned long long)lh[e]);
  }
}

// single-CTA scan of the 65536-bin histogram for the bin that holds rank `k`
__global__ void __launch_bounds__(1024) median_select_kernel(unsigned long long* hist, uint32_t* selected, int pass,
                                                             long long n_total, float* median_out) {
  if (selected[5] != 0u) return;
  __shared__ unsigned long long part[1024];
  __shared__ int s_chunk;
  __shared__ unsigned long long s_before;
  unsigned long long rank;
  if (pass == 0) rank = (unsigned long long)((n_total - 1) / 2);
  else rank = (unsigned long long)selected[1] | ((unsigned long long)selected[2] << 32);
  const int t = threadIdx.x;
  unsigned long long s = 0;
  for (int b = 0; b < 64; ++b) s += hist[t * 64 + b];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    unsigned long long cum = 0;
    int ch = 1023;
    for (int q = 0; q < 1024; ++q) {
      if (cum + part[q] > rank) { ch = q; break; }
      cum += part[q];
    }
    s_chunk = ch;
    s_before = cum;
  }
  __syncthreads();
  if (t == 0) {
    unsigned long long cum = s_before;
    int bin = s_chunk * 64 + 63;
    for (int b = 0; b < 64; ++b) {
      const unsigned long long h = hist[s_chunk * 64 + b];
      if (cum + h > rank) { bin = s_chunk * 64 + b; break; }
      cum += h;
    }
    if (pass == 0) {
      const unsigned long long rem = rank - cum;
      selected[0] = (uint32_t)bin;
      selected[1] = (uint32_t)(rem & 0xffffffffull);
      selected[2] = (uint32_t)(rem >> 32);
    } else {
      const uint32_t bits = (selected[0] << 16) | (uint32_t)bin;
      selected[3] = bits;
      if (median_out) *median_out = __uint_as_float(bits);
    }
  }
  __syncthreads();
  for (int b = 0; b < 64; ++b) hist[t * 64 + b] = 0ull;  // ready for the next pass
}

// ---------------------------------------------------------------------------------------
// phi: per CTA 64 rows; loop over column tiles; K tile -> shared; out += K [S | X | 1]
// ---------------------------------------------------------------------------------------
struct PhiLKParams {
  int N, D, row_begin, row_end;
  const float *x, *score, *xn;
  float gamma, c1, c2;
  const float* gamma_dev;
  float lr;
  float *phi, *x_out;
};

template <int MAXQ>  // accumulators per thread: ceil((2D+1)/4) <= MAXQ
__global__ void __launch_bounds__(kLThreads) phi_large_kernel(const PhiLKParams k) {
  extern __shared__ __align__(16) float sm[];
  const int D = k.D, C = 2 * D + 1;
  float* xi_t = sm;                       // [D][64]
  float* xj_t = xi_t + D * kLT;           // [D][64]
  float* xn_i = xj_t + D * kLT;           // [64]
  float* xn_j = xn_i + kLT;               // [64]
  float* ks = xn_j + kLT;                 // [64][65]  K tile, row-major padded
  float* vs = ks + kLT * (kLT + 1);       // [64][C]   V tile = [score | x | 1]
  float gamma = k.gamma, c1 = k.c1, c2 = k.c2;
  if (k.gamma_dev) { gamma = k.gamma_dev[0]; c1 = k.gamma_dev[1]; c2 = k.gamma_dev[2]; }
  const int i0 = k.row_begin + blockIdx.x * kLT;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int orow = threadIdx.x >> 2, ocg = threadIdx.x & 3;  // output mapping: row, column group
  float acc[MAXQ];
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) acc[q] = 0.f;
  stage_rows_T(k.x, k.row_end, D, D, i0, xi_t);
  if (threadIdx.x < kLT) xn_i[threadIdx.x] = (i0 + threadIdx.x < k.row_end) ? k.xn[i0 + threadIdx.x] : 0.f;
  for (int j0 = 0; j0 < k.N; j0 += kLT) {
    __syncthreads();
    stage_rows_T(k.x, k.N, D, D, j0, xj_t);
    if (threadIdx.x < kLT) xn_j[threadIdx.x] = (j0 + threadIdx.x < k.N) ? k.xn[j0 + threadIdx.x] : 0.f;
    for (int e = threadIdx.x; e < kLT * C; e += kLThreads) {
      const int r = e / C, c = e - r * C;
      const int gj = j0 + r;
      float v = 0.f;
      if (gj < k.N) v = c < D ? __ldg(k.score + (long long)gj * D + c) : (c < 2 * D ? __ldg(k.x + (long long)gj * D + (c - D)) : 1.f);
      vs[e] = v;
    }
    __syncthreads();
    float d2[4][4];
    tile_d2(xi_t, xj_t, xn_i, xn_j, D, ty, tx, d2);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool valid = (j0 + tx * 4 + c) < k.N;
        ks[(ty * 4 + r) * (kLT + 1) + tx * 4 + c] = valid ? expf(-gamma * d2[r][c]) : 0.f;
      }
    __syncthreads();
    for (int j = 0; j < kLT; ++j) {
      const float kv = ks[orow * (kLT + 1) + j];
#pragma unroll
      for (int q = 0; q < MAXQ; ++q) {
        const int c = ocg + 4 * q;
        if (c < C) acc[q] = fmaf(kv, vs[j * C + c], acc[q]);
      }
    }
  }
  // epilogue: phi_i = c1 (K S)_i + c2 (rowsum_i x_i - (K X)_i); gather the row's pieces via shared memory
  __syncthreads();
  float* outs = vs;  // reuse: [64][C]
#pragma unroll
  for (int q = 0; q < MAXQ; ++q) {
    const int c = ocg + 4 * q;
    if (c < C) outs[orow * C + c] = acc[q];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kLT * D; e += kLThreads) {
    const int r = e / D, d = e - r * D;
    const int gi = i0 + r;
    if (gi >= k.row_end) continue;
    const float ksum = outs[r * C + 2 * D];
    const float xid = xi_t[d * kLT + r];
    const float ph = c1 * outs[r * C + d] + c2 * (ksum * xid - outs[r * C + D + d]);
    if (k.phi) k.phi[(long long)gi * D + d] = ph;
    if (k.x_out) k.x_out[(long long)gi * D + d] = xid + k.lr * ph;
  }
}

int phi_tc(const dust_phi_args* a, cudaStream_t stream);  // svgd_tc.cu (tcgen05 3xTF32)
bool phi_tc_supported(const dust_phi_args* a);
size_t phi_tc_workspace(const dust_phi_args* a);

size_t phi_large_workspace(const dust_phi_args* a) {
  const size_t simt = sizeof(float) * (size_t)a->B * a->N;
  if (phi_tc_supported(a)) { const size_t tc = phi_tc_workspace(a); return tc > simt ? tc : simt; }
  return simt;
}

int phi_large(const dust_phi_args* a, cudaStream_t stream) {
  const int D = a->D, C = 2 * D + 1;
  DUST_REQUIRE(a->workspace && a->workspace_bytes >= phi_large_workspace(a), DUST_ERR_WORKSPACE,
               "dust_svgd_phi: large-N path needs %zu bytes of workspace", phi_large_workspace(a));
  if (phi_tc_supported(a) && !getenv("DUST_B200_NO_TC")) return phi_tc(a, stream);
  DUST_REQUIRE(a->ld == 0 || a->ld == a->D, DUST_ERR_UNSUPPORTED, "dust_svgd_phi: a row stride (ld=%d) needs the tensor-core path", a->ld);
  const int r0 = a->row_begin, r1 = a->row_end > 0 ? a->row_end : a->N;
  const size_t smem = sizeof(float) * ((size_t)2 * D * kLT + 2 * kLT + kLT * (kLT + 1) + (size_t)kLT * C);
  DUST_REQUIRE(smem <= 227 * 1024 && C <= 4 * 68, DUST_ERR_UNSUPPORTED, "dust_svgd_phi: D=%d too large for the tiled kernel", D);
  for (int b = 0; b < a->B; ++b) {
    const float* x = a->x + (size_t)b * a->N * D;
    float* xn = (float*)a->workspace + (size_t)b * a->N;
    { DUST_TIMED("row_norms_kernel", stream); row_norms_kernel<<<ceil_div(a->N, 256), 256, 0, stream>>>(x, a->N, D, D, xn); }
    DUST_LAUNCH_OK("row_norms_kernel");
    PhiLKParams k{a->N, D, r0, r1, x, a->score + (size_t)b * a->N * D, xn, a->gamma, a->c1, a->c2, a->gamma_dev, a->lr,
                  a->phi ? a->phi + (size_t)b * a->N * D : nullptr, a->x_out ? a->x_out + (size_t)b * a->N * D : nullptr};
    const int grid = ceil_div(r1 - r0, kLT);
#define DUST_PHI_LARGE(Q)                                                                                             \
  do {                                                                                                                \
    if (smem > 48 * 1024)                                                                                             \
      DUST_CUDA_OK(cudaFuncSetAttribute(phi_large_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    { DUST_TIMED("phi_large_kernel", stream); phi_large_kernel<Q><<<grid, kLThreads, smem, stream>>>(k); }                                                        \
  } while (0)
    if (C <= 4 * 12) DUST_PHI_LARGE(12);
    else if (C <= 4 * 21) DUST_PHI_LARGE(21);
    else if (C <= 4 * 34) DUST_PHI_LARGE(34);
    else DUST_PHI_LARGE(68);
#undef DUST_PHI_LARGE
    DUST_LAUNCH_OK("phi_large_kernel");
  }
  return DUST_OK;
}

}  // namespace dust

using namespace dust;

extern "C" int dust_median_hist_pass(const dust_median_args* a, int32_t pass, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DUST_REQUIRE(a != nullptr, DUST_ERR_INVALID_ARG, "dust_median_hist_pass: args is NULL");
  DUST_REQUIRE(a->ld == 0 || a->ld >= a->D, DUST_ERR_INVALID_ARG, "dust_median_hist_pass: ld=%d < D=%d", a->ld, a->D);
  DUST_REQUIRE(a->N > 0 && a->D > 0 && a->x && a->hist && a->selected && a->row_norms, DUST_ERR_INVALID_ARG,
               "dust_median_hist_pass: N, D, x, hist, selected, row_norms are required");
  DUST_REQUIRE(pass == 0 || pass == 1, DUST_ERR_INVALID_ARG, "dust_median_hist_pass: pass must be 0 or 1");
  const int r0 = a->row_begin, r1 = a->row_end > 0 ? a->row_end : a->N;
  DUST_REQUIRE(r0 >= 0 && r1 <= a->N && r0 < r1, DUST_ERR_INVALID_ARG, "dust_median_hist_pass: bad row range");
  float* xn = a->row_norms;
  const int ld = a->ld > 0 ? a->ld : a->D;
  if (pass == 0) {
    { DUST_TIMED("row_norms_kernel", stream); row_norms_kernel<<<ceil_div(a->N, 256), 256, 0, stream>>>(a->x, a->N, a->D, ld, xn); }
    DUST_LAUNCH_OK("row_norms_kernel");
  }
  HistKParams k{a->N, a->D, ld, r0, r1, pass, a->x, xn, a->hist, a->selected};
  size_t smem = sizeof(float) * ((size_t)2 * a->D * kLT + 2 * kLT) + (pass == 0 ? sizeof(uint32_t) * 32768 : 0);
  DUST_REQUIRE(smem <= 227 * 1024, DUST_ERR_UNSUPPORTED, "dust_median_hist_pass: D=%d too large", a->D);
  DUST_CUDA_OK(cudaFuncSetAttribute(median_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  { DUST_TIMED("median_hist_kernel", stream); median_hist_kernel<<<ceil_div(r1 - r0, kLT), kLThreads, smem, stream>>>(k); }
  DUST_LAUNCH_OK("median_hist_kernel");
  return DUST_OK;
}

namespace dust {
bool median_tc_supported(int N, int D);
size_t median_tc_workspace(int N, int D);
int median_tc_prepare(const dust_median_args* a, void* workspace, cudaStream_t stream);
int median_tc_window(const dust_median_args* a, void* workspace, cudaStream_t stream);
size_t median_tc_sample_hist_offset(int N, int D);
int median_tc_count(const dust_median_args* a, void* workspace, cudaStream_t stream);
int median_tc_select(const dust_median_args* a, float* median_out, cudaStream_t stream);
}  // namespace dust

extern "C" int dust_median_fast_supported(int32_t N, int32_t D) { return median_tc_supported(N, D) ? 1 : 0; }
extern "C" size_t dust_median_fast_workspace_bytes(int32_t N, int32_t D) { return median_tc_workspace(N, D); }

static int check_fast(const dust_median_args* a, const void* workspace, size_t bytes, const char* who) {
  DUST_REQUIRE(a != nullptr && a->x && a->hist && a->selected, DUST_ERR_INVALID_ARG, "%s: x, hist, selected are required", who);
  DUST_REQUIRE(median_tc_supported(a->N, a->D), DUST_ERR_UNSUPPORTED, "%s: shape N=%d D=%d not supported by the tensor-core pass", who, a->N, a->D);
  DUST_REQUIRE(workspace && bytes >= median_tc_workspace(a->N, a->D), DUST_ERR_WORKSPACE, "%s: workspace needs %zu bytes", who,
               median_tc_workspace(a->N, a->D));
  const int r0 = a->row_begin, r1 = a->row_end > 0 ? a->row_end : a->N;
  DUST_REQUIRE(r0 >= 0 && r1 <= a->N && r0 < r1 && r0 % 128 == 0 && r1 % 128 == 0, DUST_ERR_INVALID_ARG,
               "%s: row range must be 128-aligned", who);
  DUST_REQUIRE(a->ld == 0 || a->ld >= a->D, DUST_ERR_INVALID_ARG, "%s: ld=%d < D=%d", who, a->ld, a->D);
  DUST_REQUIRE(a->sample_begin >= 0 && a->sample_end >= a->sample_begin && a->sample_end <= (1 << 20), DUST_ERR_INVALID_ARG,
               "%s: sample share [%d, %d) outside [0, 2^20)", who, a->sample_begin, a->sample_end);
  return DUST_OK;
}

extern "C" int dust_median_fast_prepare(const dust_median_args* a, void* workspace, size_t bytes, void* stream_) {
  const int rc = check_fast(a, workspace, bytes, "dust_median_fast_prepare");
  if (rc) return rc;
  return median_tc_prepare(a, workspace, (cudaStream_t)stream_);
}
extern "C" size_t dust_median_fast_sample_hist_offset(int32_t N, int32_t D) { return median_tc_sample_hist_offset(N, D); }
extern "C" int dust_median_fast_window(const dust_median_args* a, void* workspace, size_t bytes, void* stream_) {
  const int rc = check_fast(a, workspace, bytes, "dust_median_fast_window");
  if (rc) return rc;
  return median_tc_window(a, workspace, (cudaStream_t)stream_);
}
extern "C" int dust_median_fast_count(const dust_median_args* a, void* workspace, size_t bytes, void* stream_) {
  const int rc = check_fast(a, workspace, bytes, "dust_median_fast_count");
  if (rc) return rc;
  return median_tc_count(a, workspace, (cudaStream_t)stream_);
}
extern "C" int dust_median_fast_select(const dust_median_args* a, float* median_out, void* stream_) {
  DUST_REQUIRE(a != nullptr && a->hist && a->selected, DUST_ERR_INVALID_ARG, "dust_median_fast_select: hist and selected are required");
  return median_tc_select(a, median_out, (cudaStream_t)stream_);
}

extern "C" int dust_median_select(const dust_median_args* a, int32_t pass, float* median_out, void* stream_) {
  DUST_REQUIRE(a != nullptr && a->hist && a->selected, DUST_ERR_INVALID_ARG, "dust_median_select: hist and selected are required");
  DUST_REQUIRE(pass == 0 || pass == 1, DUST_ERR_INVALID_ARG, "dust_median_select: pass must be 0 or 1");
  { DUST_TIMED("median_select_kernel", (cudaStream_t)stream_); median_select_kernel<<<1, 1024, 0, (cudaStream_t)stream_>>>(a->hist, a->selected, pass, (long long)a->N * a->N, median_out); }
  DUST_LAUNCH_OK("median_select_kernel");
  return DUST_OK;
}
