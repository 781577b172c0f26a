// Large-N SVGD phi on the 5th-generation tensor cores (tcgen05 / TMEM), flash style: the Gram
// tile S = X_i X_j^T and the products K [score | X] run as 3xTF32 MMAs (hi*hi + hi*lo + lo*hi, fp32
// accumulation in TMEM; the third term of the second GEMM, 2^-11 of the sum, on bf16 copies) so float32
// accuracy holds; exp and the hi/lo split of K happen in registers between the two GEMMs and K is never
// materialised outside TMEM.
//
//   phi_i = c1 * sum_j K_ij s_j + c2 * (x_i sum_j K_ij - sum_j K_ij x_j),  K_ij = exp(-gamma d2_ij),
//   d2_ij = max(|x_i|^2 + |x_j|^2 - 2 x_i.x_j, 0)          (dust/inference/svgd.py:28-39, 92-99, 127-135)
//
// At most one CTA per SM, resident for the whole launch; it works off an equal contiguous range of
// (128-row tile, 64-column tile) pairs, segment by segment when the range crosses row tiles.  16 warps:
//   warp 0      producer   cp.async.bulk (1-D TMA) of pre-tiled hi/lo operand images into smem rings
//   warp 1      GEMM1 issuer (one elected thread): S[j % 3] = X_i X_j^T as soon as GEMM2(j-3) has consumed that buffer
//   warp 2      TMEM allocation, then GEMM2 issuer: O += P(j) V_j as soon as the softmax warps have written P(j)
//   warp 3      producer of the V^T tiles
//   warps 4-7   "softmax" warpgroup A: columns 0..31 of every tile  (tcgen05.ld S, exp, split, tcgen05.st P)
//   warps 8-11  "softmax" warpgroup B: columns 32..63 of every tile
//   warps 12-15 flush warpgroup: stages the row tile of a segment in TMEM; every kTcChunk column tiles the O accumulator is drained into
//               round-to-nearest fp32 registers (the tensor core's own fp32 accumulation truncates:
//               measured bias ~2e-8 per accumulation step, i.e. 5e-4 over the 24576 steps of
//               N = 65536 if left in TMEM) and added to the segment's slot of a global scratch;
//               phi_tc_finish_kernel sums the slots of a row tile and forms the phi rows.
// TMEM columns (512 allocated): S/P_hi ring of three (P_hi overwrites the S it was computed from), P_lo[2] (32 columns
// each: two bf16 per column), O[2] (2*NV columns) and the row tile itself, A_hi | A_lo (2*Dp columns): every MMA reads
// only its B operand from shared memory.  (phi_tc_kernel<false>: row tile in shared memory, P_lo in TF32, 64 columns.)  An SS-form 128x64x8 TF32 MMA fetches 6 KB of operands in its 32 cycles,
// more than the 128 B/cycle the shared-memory port delivers (profiles/r2_phi_a_in_tmem.md).
// Operands live in shared memory in the canonical K-major no-swizzle UMMA layout (8x16B core
// matrices); the prep kernel writes global memory already in that order, so every tile is one
// contiguous bulk copy.
#include <stdlib.h>

#include "common.cuh"

namespace dust {

constexpr int kTcBM = 128;      // rows per CTA (MMA M)
constexpr int kTcBN = 64;       // columns per tile (GEMM1 N, GEMM2 K)
constexpr int kTcThreads = 512;
constexpr int kTcChunk = 32;    // column tiles accumulated in TMEM between two flushes
constexpr int kXbStages = 3;
constexpr int kVbStages = 3;
constexpr int kMedXbStages = 6;   // operand ring of the median kernel (no V tiles there)
constexpr int kXnStages = 16;     // |x_j|^2 ring of the median kernel

struct TcParams {
  int N, D, Dp, NV, T, row_begin;
  int ld;   // row stride of x (floats)
  int a_tmem;  // 1: the row tile (A operand of GEMM1, hi and lo) lives in TMEM, written by the flush warps (TS form)
  int bf16lo;  // 1: GEMM2's P_lo V term as kind::f16 on bf16 copies (P_lo: 32 TMEM columns per buffer; needs a_tmem)
  const void* vb_bf;   // bf16 V^T images (bf16lo)
  // Partition of the (row tile, column tile) pairs of a launch over its CTAs (at most one per SM):
  //  * the first n_chunks * R CTAs (R = row tiles of the launch) each own ONE segment: CTA c = k * R + rt takes row
  //    tile rt and the column chunk [k * chunk_w, min((k + 1) * chunk_w, T)).  All CTAs of a chunk start at the same
  //    column tile and sweep the operand images together (a tile fetched by one is an L2 hit for the others);
  //  * the columns [rem0, T) that the chunks leave over (rem_w of them per row tile) are numbered row-major,
  //    u = rt * rem_w + (j - rem0), and cut into equal contiguous ranges of units_per_cta: CTA n_chunks * R + c' owns
  //    [c' * units_per_cta, (c' + 1) * units_per_cta).  A range that crosses a row-tile boundary is worked off as
  //    consecutive SEGMENTS (one per row tile), each with its own slot of the scratch.
  // n_chunks = 0, rem0 = 0: plain row-major ranges (whole row tiles per CTA when units_per_cta = k * T).
  int R, n_chunks, chunk_w, rem0, rem_w;
  int units_per_cta, total_units, max_seg;
  const float *xa_hi, *xa_lo, *xb_hi, *xb_lo, *vb_hi, *vb_lo, *xn, *x;
  float gamma, c1, c2;
  const float* gamma_dev;
  float lr;
  float *phi, *x_out;
  float* oacc;  // [grid][max_seg][NV+2][128] running sums of the drained O chunks (column-major per slot) + 2 row-sum halves
};

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// float index of element (r, k) inside a [R x K] K-major core-matrix tile (K % 4 == 0)
__host__ __device__ inline int core_index(int r, int k, int K) {
  return ((r >> 3) * (K >> 2) + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3);
}

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ float tf32_lo(float v, float hi) { return __uint_as_float(__float_as_uint(v - hi) & 0xffffe000u); }

// GEMM2's correction term P_lo V carries 2^-11 of the result: it runs as kind::f16 on bf16 copies (P_lo = bf16(K - K_hi),
// V as bf16), 16 instead of 8 elements of K per instruction and two values per TMEM column -- P_lo takes 32 columns
// instead of 64, which is what lets a third S buffer fit beside the row tile (192 + 64 + 160 + 80 = 496 columns).
// Precision of the term: 8 + 8 mantissa bits on 2^-11 of the sum (~2^-19 relative, unbiased) instead of 2^-21.
#ifndef DUST_TC_BF16LO
#define DUST_TC_BF16LO 1
#endif
// two floats -> one 32-bit word of two bf16, element `even` in the low half (round to nearest even)
__device__ __forceinline__ uint32_t pack_bf16x2(float even, float odd) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(odd), "f"(even));
  return d;
}

// ---------------------------------------------------------------------------------------
// prep: squared norms and the tiled hi / lo operand images
// ---------------------------------------------------------------------------------------
// One image serves both operand roles: element (row, k) sits at
//   ((row >> 3) * (Dp / 4) + (k >> 2)) * 32 + (row & 7) * 4 + (k & 3),
// which is core_index() inside a 128-row tile AND inside a 64-row tile (both are whole 8-row groups
// laid out in row order).  Thread <-> one 16-byte core-matrix row, in image order: a warp writes 512
// contiguous bytes per image and reads 64-byte pieces of 8 rows.
__device__ __forceinline__ void tc_prep_x_block(int block, const float* __restrict__ x, int N, int D, int ld, int Dp,
                                                float* __restrict__ x_hi, float* __restrict__ x_lo) {
  const long long e = (long long)block * blockDim.x + threadIdx.x;   // index of the float4 in the image
  const int kq = Dp >> 2;
  if (e >= (long long)N * kq) return;
  const int r7 = (int)(e & 7);
  const long long t = e >> 3;
  const int kg = (int)(t % kq);
  const int row = (int)(t / kq) * 8 + r7, k = kg * 4;
  float v[4];
  const float* xr = x + (long long)row * ld;
  if ((D & 3) == 0 && (ld & 3) == 0 && ((((uintptr_t)x) & 15) == 0) && k + 3 < D) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(xr + k));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = (k + c < D) ? xr[k + c] : 0.f;
  }
  float4 hi, lo;
  hi.x = tf32_hi(v[0]); hi.y = tf32_hi(v[1]); hi.z = tf32_hi(v[2]); hi.w = tf32_hi(v[3]);
  lo.x = tf32_lo(v[0], hi.x); lo.y = tf32_lo(v[1], hi.y); lo.z = tf32_lo(v[2], hi.z); lo.w = tf32_lo(v[3], hi.w);
  reinterpret_cast<float4*>(x_hi)[e] = hi;
  reinterpret_cast<float4*>(x_lo)[e] = lo;
}

// |x_i|^2 as one fused-multiply-add chain over d = 0 .. D-1 per row (a warp reads 32 consecutive rows)
__device__ __forceinline__ void tc_norms_block(int block, const float* __restrict__ x, int N, int D, int ld, float* __restrict__ xn) {
  const int row = block * blockDim.x + threadIdx.x;
  if (row >= N) return;
  const float* xr = x + (long long)row * ld;
  float s = 0.f;
  if ((D & 3) == 0 && (ld & 3) == 0 && ((((uintptr_t)x) & 15) == 0)) {
    for (int d = 0; d < D; d += 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(xr + d));
      s = fmaf(q.x, q.x, s); s = fmaf(q.y, q.y, s); s = fmaf(q.z, q.z, s); s = fmaf(q.w, q.w, s);
    }
  } else {
    for (int d = 0; d < D; ++d) { const float t = xr[d]; s = fmaf(t, t, s); }
  }
  xn[row] = s;
}

// V^T tiles: rows n2 in [0, NV) = [score dims | x dims | 0...], K = the 64 columns j of the tile
// (the row sums of K are accumulated exactly in the softmax warps' registers instead of a ones column).
// Thread <-> one 16-byte core-matrix row = 4 consecutive particles of one dimension, in image order:
// element (n2, jj) of tile jt sits at jt * NV * 64 + ((n2 >> 3) * 16 + (jj >> 2)) * 32 + (n2 & 7) * 4 + (jj & 3).
// The bf16 copy (GEMM2's P_lo V term) uses the same K-major core-matrix order with 8 elements per 16-byte row:
// element (n2, jj) of tile jt at bf16 index jt * NV * 64 + ((n2 >> 3) * 8 + (jj >> 3)) * 64 + (n2 & 7) * 8 + (jj & 7).
__device__ __forceinline__ void tc_prep_v_block(int block, const float* __restrict__ x, const float* __restrict__ score, int N,
                                                int D, int ld, int NV, float* __restrict__ vb_hi, float* __restrict__ vb_lo,
                                                uint2* __restrict__ vb_bf) {
  const long long e = (long long)block * blockDim.x + threadIdx.x;   // index of the float4 in the image
  if (e >= (long long)(N >> 2) * NV) return;
  const int per_tile = NV * (kTcBN >> 2);
  const int jt = (int)(e / per_tile), w = (int)(e - (long long)jt * per_tile);
  const int n7 = w & 7, jq = (w >> 3) & 15, ng = w >> 7;
  const int n2 = ng * 8 + n7, j = jt * kTcBN + jq * 4;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (n2 < 2 * D) {
    const float* src = n2 < D ? score + n2 : x + (n2 - D);
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = __ldg(src + (long long)(j + c) * ld);
  }
  float4 hi, lo;
  hi.x = tf32_hi(v[0]); hi.y = tf32_hi(v[1]); hi.z = tf32_hi(v[2]); hi.w = tf32_hi(v[3]);
  lo.x = tf32_lo(v[0], hi.x); lo.y = tf32_lo(v[1], hi.y); lo.z = tf32_lo(v[2], hi.z); lo.w = tf32_lo(v[3], hi.w);
  reinterpret_cast<float4*>(vb_hi)[e] = hi;
  reinterpret_cast<float4*>(vb_lo)[e] = lo;
  if (vb_bf)
    vb_bf[(long long)jt * NV * 16 + (ng * 8 + (jq >> 1)) * 16 + n7 * 2 + (jq & 1)] =
        make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}

// ONE launch prepares everything a pass needs: CTAs [0, nbx) write the X images, [nbx, nbx + nbv) the
// V^T images (nbv = 0 for the median pass, which has no second GEMM), the rest the squared norms --
// three small memory-bound jobs that overlap instead of queueing behind each other's launch
__global__ void __launch_bounds__(256) tc_prep_kernel(const float* __restrict__ x, const float* __restrict__ score, int N, int D, int ld,
                                                      int Dp, int NV, float* __restrict__ x_hi, float* __restrict__ x_lo,
                                                      float* __restrict__ vb_hi, float* __restrict__ vb_lo, uint2* __restrict__ vb_bf,
                                                      float* __restrict__ xn, int nbx, int nbv) {
  const int b = blockIdx.x;
  if (b < nbx) tc_prep_x_block(b, x, N, D, ld, Dp, x_hi, x_lo);
  else if (b < nbx + nbv) tc_prep_v_block(b - nbx, x, score, N, D, ld, NV, vb_hi, vb_lo, vb_bf);
  else tc_norms_block(b - nbx - nbv, x, N, D, ld, xn);
}

static int tc_prep(const float* x, const float* score, int N, int D, int ld, int Dp, int NV, float* x_hi, float* x_lo, float* vb_hi,
                   float* vb_lo, float* xn, cudaStream_t stream, bool x_prepared = false, void* vb_bf = nullptr) {
  // x_prepared: the X images and the norms are already in place (median pass on the same workspace): V^T images only
  const int nbx = x_prepared ? 0 : ceil_div((long long)N * (Dp / 4), 256);
  const int nbv = score ? ceil_div((long long)(N / 4) * NV, 256) : 0;
  const int nbn = x_prepared ? 0 : ceil_div(N, 256);
  if (nbx + nbv + nbn == 0) return DUST_OK;
  {
    DUST_TIMED("tc_prep_kernel", stream);
    tc_prep_kernel<<<nbx + nbv + nbn, 256, 0, stream>>>(x, score, N, D, ld, Dp, NV, x_hi, x_lo, vb_hi, vb_lo, (uint2*)vb_bf, xn, nbx, nbv);
  }
  DUST_LAUNCH_OK("tc_prep_kernel");
  return DUST_OK;
}

// ---------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------
// smem_u32, mbar_init / mbar_expect_tx / mbar_arrive / mbar_wait and bulk_g2s live in common.cuh
// one elected lane of a fully converged warp (the compiler then keeps descriptors / addresses in
// uniform registers and issues UTCHMMA directly, instead of an election loop per instruction)
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: LBO = byte distance between K-adjacent core matrices, SBO = between 8-row groups
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// descriptors of one operand differ only in the 14-bit start address: keep the high word, bump the low
__device__ __forceinline__ uint64_t desc_lo_hi(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4)                    // D format F32
         | (2u << 7) | (2u << 10)     // A, B format TF32
         | ((uint32_t)(N >> 3) << 17) // N
         | ((uint32_t)(M >> 4) << 24);// M       (a_major = b_major = 0: K-major)
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4)                    // D format F32
         | (1u << 7) | (1u << 10)     // A, B format BF16
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------
struct TcSmem {
  // byte offsets into dynamic shared memory
  uint32_t a_hi, a_lo, xb, vb, bars, tmem_slot, total;
  uint32_t xb_stage_bytes, vb_stage_bytes, xb_half, vb_half, vb_bf_bytes;
};

// a_tmem: the row tile lives in TMEM, no shared memory for it; bf16lo: a V stage also carries the bf16 copy
__host__ __device__ inline TcSmem tc_smem_layout(int Dp, int NV, bool a_tmem = false, bool bf16lo = false) {
  TcSmem s;
  uint32_t off = 0;
  s.a_hi = off; off += a_tmem ? 0 : kTcBM * Dp * 4;
  s.a_lo = off; off += a_tmem ? 0 : kTcBM * Dp * 4;
  s.xb_half = kTcBN * Dp * 4;
  s.xb_stage_bytes = 2 * s.xb_half;
  s.xb = off; off += kXbStages * s.xb_stage_bytes;
  s.vb_half = NV * kTcBN * 4;
  s.vb_bf_bytes = bf16lo ? NV * kTcBN * 2 : 0;
  s.vb_stage_bytes = 2 * s.vb_half + kTcBN * 4 + s.vb_bf_bytes;  // hi, lo, |x_j|^2, bf16 copy
  s.vb = off; off += kVbStages * s.vb_stage_bytes;
  off = (off + 7) & ~7u;
  s.bars = off; off += 40 * 8;
  s.tmem_slot = off; off += 16;
  s.total = off;
  return s;
}

// S / P_hi ring of kSBufs = 3 (P_hi overwrites the S it came from), P_lo ring of 2: GEMM1 runs TWO tiles ahead of
// GEMM2, so the softmax stage of a tile has two tile periods (not one) before the tensor pipe waits for it
#ifndef DUST_TC_SBUFS
#define DUST_TC_SBUFS 3          // S/P_hi ring: GEMM1 may run three tiles ahead of GEMM2.  With ONE issuing warp a ring of
                                 // three measured slower (5.43 vs 5.04 ms at N = 65536); with two issuers the S -> exp -> P ->
                                 // GEMM2 chain through two buffers is what binds (profiles/r2_phi_a_in_tmem.md)
#endif
constexpr int kSBufs = DUST_TC_SBUFS;
enum { BAR_A = 0, BAR_XB_FULL = 1, BAR_XB_EMPTY = 5, BAR_VB_FULL = 9, BAR_VB_EMPTY = 13, BAR_S_FULL = 17, BAR_P_FULL = 20,
       BAR_P_EMPTY = 23, BAR_O_FULL = 25, BAR_O_EMPTY = 27, BAR_A_EMPTY = 29, BAR_S_EMPTY = 30, BAR_COUNT = 33 };

// the segments of a CTA's unit range: `for (TcSegIter s(p); s.valid(); s.next())` gives row tile s.rt,
// first column tile s.j0 and tile count s.len of segment s.seg
struct TcSegIter {
  int u, u1, W, j_off, seg, rt, j0, len;
  __device__ __forceinline__ explicit TcSegIter(const TcParams& p) : seg(0) {
    const int c = blockIdx.x - p.n_chunks * p.R;
    if (c < 0) {                       // one (row tile, column chunk) segment
      const int k = blockIdx.x / p.R;
      rt = blockIdx.x - k * p.R;
      j0 = k * p.chunk_w;
      len = min(p.chunk_w, p.T - j0);
      u = 0; u1 = len; W = 0; j_off = 0;
    } else {                           // a contiguous range of the row-major numbered remainder columns
      W = p.rem_w; j_off = p.rem0;
      u = c * p.units_per_cta;
      u1 = min(u + p.units_per_cta, p.total_units);
      split();
    }
  }
  __device__ __forceinline__ void split() {
    rt = u / W;
    j0 = j_off + (u - rt * W);
    len = min(W - (j0 - j_off), u1 - u);
  }
  __device__ __forceinline__ bool valid() const { return u < u1; }
  __device__ __forceinline__ void next() {
    u += len; ++seg;
    if (W > 0 && u < u1) split();
  }
};

// FAST: the row tile of GEMM1 in TMEM (TS form) and GEMM2's P_lo V term as kind::f16 on bf16 copies (tc_mode)
template <bool FAST>
__global__ void __launch_bounds__(kTcThreads, 1) phi_tc_kernel(const TcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const TcSmem L = tc_smem_layout(p.Dp, p.NV, FAST, FAST);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem_slot);
  // the warp index as a value the compiler knows to be warp-uniform: the role branches below become uniform
  // branches and the MMA issuer's loop state (ring indices, descriptors, TMEM addresses) stays in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const size_t slot_stride = (size_t)(p.NV + 2) * kTcBM;
  float* const oacc_cta = p.oacc + (size_t)blockIdx.x * p.max_seg * slot_stride;

  if (threadIdx.x == 0) {
    mbar_init(&bars[BAR_A], FAST ? 4 : 1);   // TMEM: one arrival per flush warp; shared memory: the TMA transaction
    mbar_init(&bars[BAR_A_EMPTY], 1);
    for (int s = 0; s < kXbStages; ++s) { mbar_init(&bars[BAR_XB_FULL + s], 1); mbar_init(&bars[BAR_XB_EMPTY + s], 1); }
    for (int s = 0; s < kVbStages; ++s) { mbar_init(&bars[BAR_VB_FULL + s], 1); mbar_init(&bars[BAR_VB_EMPTY + s], 1); }
    for (int b = 0; b < kSBufs; ++b) {
      mbar_init(&bars[BAR_S_FULL + b], 1);
      mbar_init(&bars[BAR_S_EMPTY + b], 1);
      mbar_init(&bars[BAR_P_FULL + b], 8);   // one arrival per softmax warp (both groups work on every tile)
    }
    for (int b = 0; b < 2; ++b) mbar_init(&bars[BAR_P_EMPTY + b], 1);   // P_lo ring
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[BAR_O_FULL + b], 1);
      mbar_init(&bars[BAR_O_EMPTY + b], 4);  // one arrival per flush warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  constexpr uint32_t plo_w = FAST ? kTcBN / 2 : kTcBN;      // TMEM columns of one P_lo buffer (two bf16 per column)
  const uint32_t tS = tmem, tPhi = tmem, tPlo = tmem + kSBufs * kTcBN, tO = tPlo + 2 * plo_w;
  const uint32_t tAhi = tO + 2 * p.NV, tAlo = tAhi + p.Dp;                                       // a_tmem only
  // Ring stages, S/P buffers and O buffers are indexed by counters that run on ACROSS segments
  // (g: tiles, cg: flushed chunks), so the pipelines never drain at a row-tile boundary; only the
  // A operand (the row tile itself) is exchanged there.

  if (warp == 0) {
    // ------------------------------ producer ------------------------------------------
    if (lane == 0) {
      const uint32_t a_bytes = kTcBM * p.Dp * 4;
      int g = 0;
      for (TcSegIter s(p); s.valid(); s.next()) {
        if (!FAST) {
          // the GEMM1s of the previous segment must have retired before their A operand is overwritten
          if (s.seg > 0) mbar_wait(&bars[BAR_A_EMPTY], (s.seg - 1) & 1);
          const long long arow = (long long)(p.row_begin / kTcBM + s.rt) * kTcBM * p.Dp;
          mbar_expect_tx(&bars[BAR_A], 2 * a_bytes);
          bulk_g2s(smem + L.a_hi, p.xa_hi + arow, a_bytes, &bars[BAR_A]);
          bulk_g2s(smem + L.a_lo, p.xa_lo + arow, a_bytes, &bars[BAR_A]);
        }
        for (int j = 0; j < s.len; ++j, ++g) {
          const int sx = g % kXbStages;
          mbar_wait(&bars[BAR_XB_EMPTY + sx], ((g / kXbStages) & 1) ^ 1);
          unsigned char* xb = smem + L.xb + sx * L.xb_stage_bytes;
          mbar_expect_tx(&bars[BAR_XB_FULL + sx], L.xb_stage_bytes);
          bulk_g2s(xb, p.xb_hi + (long long)(s.j0 + j) * kTcBN * p.Dp, L.xb_half, &bars[BAR_XB_FULL + sx]);
          bulk_g2s(xb + L.xb_half, p.xb_lo + (long long)(s.j0 + j) * kTcBN * p.Dp, L.xb_half, &bars[BAR_XB_FULL + sx]);
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------ producer of the V^T tiles --------------------------
    if (lane == 0) {
      int g = 0;
      for (TcSegIter s(p); s.valid(); s.next()) {
        for (int j = 0; j < s.len; ++j, ++g) {
          const int sv = g % kVbStages;
          mbar_wait(&bars[BAR_VB_EMPTY + sv], ((g / kVbStages) & 1) ^ 1);
          unsigned char* vb = smem + L.vb + sv * L.vb_stage_bytes;
          mbar_expect_tx(&bars[BAR_VB_FULL + sv], L.vb_stage_bytes);
          bulk_g2s(vb, p.vb_hi + (long long)(s.j0 + j) * p.NV * kTcBN, L.vb_half, &bars[BAR_VB_FULL + sv]);
          bulk_g2s(vb + L.vb_half, p.vb_lo + (long long)(s.j0 + j) * p.NV * kTcBN, L.vb_half, &bars[BAR_VB_FULL + sv]);
          bulk_g2s(vb + 2 * L.vb_half, p.xn + (long long)(s.j0 + j) * kTcBN, kTcBN * 4, &bars[BAR_VB_FULL + sv]);
          if (FAST)
            bulk_g2s(vb + 2 * L.vb_half + kTcBN * 4, reinterpret_cast<const unsigned char*>(p.vb_bf) + (long long)(s.j0 + j) * L.vb_bf_bytes,
                     L.vb_bf_bytes, &bars[BAR_VB_FULL + sv]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ GEMM1 issuer --------------------------------------
    // Two issuing warps, one per GEMM: a single warp spent ~2500 cycles per tile on its ~290 instructions (descriptor
    // moves, waits, commits; ncu: the issuer never waited for P, the softmax warps waited for S) against 1440 cycles of
    // tensor-pipe work.  Every dependency between the two GEMMs is an mbarrier (S_EMPTY, P_FULL, P_EMPTY), none relies
    // on issue order.  The whole warp runs the loop converged (waits included); one elected lane issues.
    const uint32_t idesc1 = make_idesc_tf32(kTcBM, kTcBN);
    const uint32_t sbo1 = (uint32_t)(p.Dp / 4) * 128u;     // 8-row group stride of the X tiles
    const int ks1 = p.Dp / 8;
    // Every integer op on the way to an MMA costs the issuing warp ~5 cycles: descriptors are built once, per MMA
    // only the low word (start address >> 4) moves.
    const uint64_t dA = make_desc(smem_u32(smem + L.a_hi), 128, sbo1), dX = make_desc(smem_u32(smem + L.xb), 128, sbo1);
    const uint32_t hiA = (uint32_t)(dA >> 32);
    const uint32_t loA_hi = (uint32_t)dA, loA_lo = loA_hi + ((kTcBM * p.Dp * 4) >> 4);
    const uint32_t loX0 = (uint32_t)dX;
    const uint32_t xb_stage16 = L.xb_stage_bytes >> 4, xb_half16 = L.xb_half >> 4;
    int g = 0;                               // running tile counter
    for (TcSegIter s(p); s.valid(); s.next()) {
      mbar_wait(&bars[BAR_A], s.seg & 1);
      for (int j = 0; j < s.len; ++j, ++g) {
        const int b = g % kSBufs, sx = g % kXbStages;
        mbar_wait(&bars[BAR_XB_FULL + sx], (g / kXbStages) & 1);
        mbar_wait(&bars[BAR_S_EMPTY + b], ((g / kSBufs) & 1) ^ 1);   // GEMM2(g - kSBufs) has consumed the P_hi in this buffer
        tc_fence_after();
        const uint32_t xh = loX0 + sx * xb_stage16, xl = xh + xb_half16;
        const uint32_t tSb = tS + b * kTcBN;
        if (elect_one_sync()) {
          if (FAST) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              if (kk < ks1) {
                const uint64_t bh = desc_lo_hi(xh + kk * 16, hiA), bl = desc_lo_hi(xl + kk * 16, hiA);
                mma_ts(tSb, tAhi + kk * 8, bh, idesc1, kk > 0 ? 1u : 0u);
                mma_ts(tSb, tAhi + kk * 8, bl, idesc1, 1u);
                mma_ts(tSb, tAlo + kk * 8, bh, idesc1, 1u);
              }
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              if (kk < ks1) {
                const uint64_t ah = desc_lo_hi(loA_hi + kk * 16, hiA), al = desc_lo_hi(loA_lo + kk * 16, hiA);
                const uint64_t bh = desc_lo_hi(xh + kk * 16, hiA), bl = desc_lo_hi(xl + kk * 16, hiA);
                mma_ss(tSb, ah, bh, idesc1, kk > 0 ? 1u : 0u);
                mma_ss(tSb, ah, bl, idesc1, 1u);
                mma_ss(tSb, al, bh, idesc1, 1u);
              }
            }
          }
          tc_commit(&bars[BAR_S_FULL + b]);
          tc_commit(&bars[BAR_XB_EMPTY + sx]);
          if (j == s.len - 1) tc_commit(&bars[BAR_A_EMPTY]);   // last reader of this segment's A operand
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ------------------------------ GEMM2 issuer (after the TMEM allocation above) ------
    // O[chunk & 1] (+)= P[g] V_g, tile by tile; first / last tile of a flush chunk open / close an O buffer
    const uint32_t idesc2 = make_idesc_tf32(kTcBM, p.NV), idesc2b = make_idesc_bf16(kTcBM, p.NV);
    const uint32_t sbo2 = (uint32_t)(kTcBN / 4) * 128u;    // 8-row group stride of the V^T tiles
    const uint32_t sbo2b = (uint32_t)(kTcBN / 8) * 128u;   // ... of their bf16 copies (8 elements per core-matrix row)
    const uint32_t hiVb = (uint32_t)(make_desc(0, 128, sbo2b) >> 32);
    const uint32_t vb_bf16 = (2 * L.vb_half + kTcBN * 4) >> 4;   // the bf16 copy inside a V stage, in 16-byte units
    constexpr int ks2 = kTcBN / 8;
    const uint64_t dV = make_desc(smem_u32(smem + L.vb), 128, sbo2);
    const uint32_t hiV = (uint32_t)(dV >> 32), loV0 = (uint32_t)dV;
    const uint32_t vb_stage16 = L.vb_stage_bytes >> 4, vb_half16 = L.vb_half >> 4;
    int g = 0, cg = 0;                       // running tile / chunk counters
    for (TcSegIter s(p); s.valid(); s.next()) {
      for (int j = 0; j < s.len; ++j, ++g) {
        const bool first = (j % kTcChunk) == 0;
        const bool last = (j % kTcChunk) == kTcChunk - 1 || j == s.len - 1;
        const int b = g & 1, b3 = g % kSBufs, sv = g % kVbStages, ob = cg & 1;
        mbar_wait(&bars[BAR_VB_FULL + sv], (g / kVbStages) & 1);
        if (first) mbar_wait(&bars[BAR_O_EMPTY + ob], ((cg >> 1) & 1) ^ 1);   // the flush warps drained this O buffer (two chunks ago)
        mbar_wait(&bars[BAR_P_FULL + b3], (g / kSBufs) & 1);
        tc_fence_after();
        const uint32_t tOb = tO + ob * p.NV;
        const uint32_t vh = loV0 + sv * vb_stage16, vl = vh + vb_half16;
        const uint32_t ph = tPhi + b3 * kTcBN, pl = tPlo + b * plo_w;
        if (elect_one_sync()) {
          if (FAST) {
#pragma unroll
            for (int kk = 0; kk < ks2; ++kk) {
              const uint64_t bh = desc_lo_hi(vh + kk * 16, hiV), bl = desc_lo_hi(vl + kk * 16, hiV);
              mma_ts(tOb, ph + kk * 8, bh, idesc2, (!first || kk > 0) ? 1u : 0u);
              mma_ts(tOb, ph + kk * 8, bl, idesc2, 1u);
            }
            const uint32_t vbf = vh + vb_bf16;
#pragma unroll
            for (int kk = 0; kk < ks2 / 2; ++kk)     // 16 bf16 of K per instruction = two core matrices = 8 TMEM columns
              mma_ts_f16(tOb, pl + kk * 8, desc_lo_hi(vbf + kk * 16, hiVb), idesc2b, 1u);
          } else {
#pragma unroll
            for (int kk = 0; kk < ks2; ++kk) {
              const uint64_t bh = desc_lo_hi(vh + kk * 16, hiV), bl = desc_lo_hi(vl + kk * 16, hiV);
              mma_ts(tOb, ph + kk * 8, bh, idesc2, (!first || kk > 0) ? 1u : 0u);
              mma_ts(tOb, ph + kk * 8, bl, idesc2, 1u);
              mma_ts(tOb, pl + kk * 8, bh, idesc2, 1u);
            }
          }
          tc_commit(&bars[BAR_S_EMPTY + b3]);
          tc_commit(&bars[BAR_P_EMPTY + b]);
          tc_commit(&bars[BAR_VB_EMPTY + sv]);
          if (last) tc_commit(&bars[BAR_O_FULL + ob]);
        }
        __syncwarp();
        if (last) ++cg;
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------ softmax warpgroups ---------------------------------
    // Both warpgroups work on EVERY tile, one 32-column half each: with two S/P buffers the tile
    // period is (softmax latency + MMA time) / 2, so the latency of this stage is what matters.
    const int half = (warp - 4) >> 2;        // column half of the tile handled by this warpgroup
    const int q = warp & 3;                  // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float gamma = p.gamma;
    if (p.gamma_dev) gamma = p.gamma_dev[0];
    const float g2 = gamma * 1.4426950408889634f;  // exp(-g d2) = 2^(-g log2(e) d2)
    int g = 0;
    for (TcSegIter s(p); s.valid(); s.next()) {
      const int i0 = p.row_begin + s.rt * kTcBM;   // first global row of this segment's row tile
      const float xn_i = p.xn[i0 + row];
      const int jdiag = (i0 + row) / kTcBN;
      const int cdiag_all = (i0 + row) & (kTcBN - 1);
      float ksum = 0.f;  // sum_j K_ij over this warpgroup's column halves (exact fp32, no ones column in V)
      for (int j = 0; j < s.len; ++j, ++g) {
        const int b = g & 1, it = g >> 1, b3 = g % kSBufs, it3 = g / kSBufs, sv = g % kVbStages;
        mbar_wait(&bars[BAR_VB_FULL + sv], (g / kVbStages) & 1);   // |x_j|^2 of this tile
        const float* xnj = reinterpret_cast<const float*>(smem + L.vb + sv * L.vb_stage_bytes + 2 * L.vb_half) + half * 32;
        mbar_wait(&bars[BAR_S_FULL + b3], it3 & 1);
        tc_fence_after();
        uint32_t r[32], lo[32];
        tmem_ld32(tS + lane_base + b3 * kTcBN + half * 32, r);
        tmem_wait_ld();
        // the tile that holds column i itself: d2_ii is exactly 0 (the 3xTF32 Gram entry only gives
        // |x_i|^2 to ~1e-6 relative, which a narrow kernel would amplify)
        if (s.j0 + j == jdiag && (cdiag_all >> 5) == half) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c == (cdiag_all & 31)) r[c] = __float_as_uint(0.5f * (xnj[c] + xn_i));  // => d2 = 0
        }
        // d2 = max((x_j - 2s) + x_i, 0) with 2s exact inside the fma; K = 2^(-g2 d2); hi = its TF32 part, lo = K - hi.
        // Scalar on purpose: the same loop on the packed FP32 pipe (FFMA2 / FADD2 / FMUL2, 23 % fewer instructions)
        // measured 13 % SLOWER -- 128 registers, loads of |x_j|^2 no longer hoisted, the softmax warps latency-bound
        // (profiles/r2_phi_a_in_tmem.md, column 5).
        uint32_t lo2[16];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float d2 = fmaxf(fmaf(-2.0f, __uint_as_float(r[c]), xnj[c]) + xn_i, 0.f);
          const float kv = ex2_approx(-g2 * d2);
          ksum += kv;
          const float hi = tf32_hi(kv);
          r[c] = __float_as_uint(hi);
          lo[c] = FAST ? __float_as_uint(kv - hi) : __float_as_uint(tf32_lo(kv, hi));
        }
        // P_hi overwrites the S columns it came from (this tile's own buffer); GEMM2(g-2) must be done with P_lo[b]
        mbar_wait(&bars[BAR_P_EMPTY + b], (it & 1) ^ 1);
        tc_fence_after();
        tmem_st32(tPhi + lane_base + b3 * kTcBN + half * 32, r);
        if (FAST) {
          // two bf16 per TMEM column, the even element of a pair in the low half (the other order is 1.7e-4 off float64
          // instead of 4e-6: measured, profiles/r2_run18.sh)
#pragma unroll
          for (int c = 0; c < 16; ++c) lo2[c] = pack_bf16x2(__uint_as_float(lo[2 * c]), __uint_as_float(lo[2 * c + 1]));
          tmem_st16(tPlo + lane_base + b * plo_w + half * 16, lo2);
        } else {
          tmem_st32(tPlo + lane_base + b * kTcBN + half * 32, lo);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_P_FULL + b3]);
      }
      oacc_cta[(size_t)s.seg * slot_stride + (size_t)(p.NV + half) * kTcBM + row] = ksum;
    }
  } else if (warp >= 12) {
    // ------------------------------ flush warpgroup + epilogue -------------------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    // running sums live in a per-segment global scratch, column-major ([NV][128]: a warp touches 128
    // contiguous bytes per column); only 16 columns are in registers at any time
    int cg = 0;
    // a_tmem: this warpgroup also stages the row tile of a segment in TMEM (thread <-> row: hi / lo split of its Dp
    // values, the same bits tc_prep_kernel writes into the images).  seg > 0: once the previous segment's last GEMM1
    // has retired (BAR_A_EMPTY) -- which does not depend on the chunk drained below, so it goes first
    auto stage_a = [&](int seg, int rt) {
      const float* __restrict__ xr = p.x + (long long)(p.row_begin + rt * kTcBM + row) * p.ld;
      const bool vec = (p.D & 3) == 0 && (p.ld & 3) == 0 && ((((uintptr_t)p.x) & 15) == 0);
      if (seg > 0) {
        mbar_wait(&bars[BAR_A_EMPTY], (seg - 1) & 1);
        tc_fence_after();
      }
      for (int k0 = 0; k0 < p.Dp; k0 += 8) {
        float v[8];
        if (vec && k0 + 8 <= p.D) {
          const float4 q0 = __ldg(reinterpret_cast<const float4*>(xr + k0)), q1 = __ldg(reinterpret_cast<const float4*>(xr + k0 + 4));
          v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = (k0 + c < p.D) ? __ldg(xr + k0 + c) : 0.f;
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float h = tf32_hi(v[c]);
          hi[c] = __float_as_uint(h);
          lo[c] = __float_as_uint(tf32_lo(v[c], h));
        }
        tmem_st8(tAhi + lane_base + k0, hi);
        tmem_st8(tAlo + lane_base + k0, lo);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_A]);
    };
    if (FAST) {
      TcSegIter s0(p);
      if (s0.valid()) stage_a(0, s0.rt);
    }
    for (TcSegIter s(p); s.valid(); s.next()) {
      float* og = oacc_cta + (size_t)s.seg * slot_stride + row;
      const int n_chunks = (s.len + kTcChunk - 1) / kTcChunk;
      for (int ch = 0; ch < n_chunks; ++ch, ++cg) {
        if (FAST && ch == n_chunks - 1) {
          TcSegIter nx = s;
          nx.next();
          if (nx.valid()) stage_a(nx.seg, nx.rt);
        }
        const int ob = cg & 1;
        mbar_wait_relaxed(&bars[BAR_O_FULL + ob], (cg >> 1) & 1);   // 32 tile times apart: do not spin on the softmax warps' scheduler
        tc_fence_after();
        for (int c0 = 0; c0 < p.NV; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tO + lane_base + ob * p.NV + c0, r);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float v = __uint_as_float(r[c]);
            if (ch > 0) v += og[(size_t)(c0 + c) * kTcBM];
            og[(size_t)(c0 + c) * kTcBM] = v;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_O_EMPTY + ob]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// combine the partial sums of one row tile (every segment that worked on it, in CTA order) and form
// phi; grid (row tiles, ceil(D / 8)): thread <-> row, 8 dimensions per CTA
__global__ void __launch_bounds__(kTcBM) phi_tc_finish_kernel(const TcParams p) {
  const int rt = blockIdx.x, row = threadIdx.x;
  const int gi = p.row_begin + rt * kTcBM + row;
  float c1 = p.c1, c2 = p.c2;
  if (p.gamma_dev) { c1 = p.gamma_dev[1]; c2 = p.gamma_dev[2]; }
  const size_t slot_stride = (size_t)(p.NV + 2) * kTcBM;
  // the segments that worked on row tile rt, in a fixed order: the chunk CTAs k * R + rt (slot 0 each), then the
  // remainder CTAs c' whose range meets [rt * rem_w, (rt + 1) * rem_w) (slot = rt - first row tile of c')
  const long long W = p.units_per_cta, Tw = p.rem_w;
  const int nc = p.n_chunks, cbase = p.n_chunks * p.R;
  int c_first = 0, c_last = -1;
  if (Tw > 0) { c_first = (int)((rt * Tw) / W); c_last = (int)(((rt + 1) * Tw - 1) / W); }
  const int n_src = nc + (c_last - c_first + 1);
  auto slot = [&](int i) {
    if (i < nc) return p.oacc + (size_t)(i * p.R + rt) * p.max_seg * slot_stride + row;
    const int c = c_first + (i - nc);
    return p.oacc + ((size_t)(cbase + c) * p.max_seg + (size_t)(rt - (c * W) / Tw)) * slot_stride + row;
  };
  float ksum = 0.f;
  for (int i = 0; i < n_src; ++i) {
    const float* og = slot(i);
    ksum += og[(size_t)p.NV * kTcBM] + og[(size_t)(p.NV + 1) * kTcBM];
  }
  const int d0 = blockIdx.y * 8, d1 = min(d0 + 8, p.D);
  for (int d = d0; d < d1; ++d) {
    float ks = 0.f, kx = 0.f;  // sum_j K s_j, sum_j K x_j
    for (int i = 0; i < n_src; ++i) {
      const float* og = slot(i);
      ks += og[(size_t)d * kTcBM];
      kx += og[(size_t)(p.D + d) * kTcBM];
    }
    const float xv = p.x[(long long)gi * p.ld + d];
    const float ph = c1 * ks + c2 * (ksum * xv - kx);
    if (p.phi) p.phi[(long long)gi * p.D + d] = ph;
    if (p.x_out) p.x_out[(long long)gi * p.D + d] = xv + p.lr * ph;
  }
}

// Partition of one launch (see TcParams).  A floor on the range length keeps the per-CTA prologue (TMEM allocation,
// A operand, pipeline fill, O drain: ~5 tile times) amortised.
struct TcPlan { int grid, R, n_chunks, chunk_w, rem0, rem_w, units_per_cta, total_units, max_seg, row_tile0, row_tiles; };
constexpr int kSegOverhead = 3;   // tile times a further segment costs its CTA (A exchange, O drain, scratch slot)

// fewer row tiles than SMs: R * T units over up to 148 CTAs
static TcPlan tc_plan(int R, int T) {
  TcPlan pl{};
  pl.row_tile0 = 0; pl.row_tiles = R; pl.R = R;
  const long long units = (long long)R * T;
  int W = (int)ceil_div(units, (long long)kNumSMs);
  if (W < 16) {                       // little work: plain row-major ranges of at least 16 units
    W = T < 16 ? T : 16;
    pl.n_chunks = 0; pl.chunk_w = 0; pl.rem0 = 0; pl.rem_w = T;
    pl.units_per_cta = W; pl.total_units = (int)units;
    pl.grid = ceil_div(pl.total_units, W);
    pl.max_seg = (W + T - 2) / T + 1;
    return pl;
  }
  // (A) chunks of exactly W columns + the left-over columns as row-major ranges of W units;  (B) floor(148 / R) chunks
  // that cover all T columns, some SMs idle.  The cheaper one by the longest CTA (+ segment overheads) is taken.
  const int nfA = T / W, remA = T - nfA * W;
  const int segA = remA > 0 ? (W + remA - 2) / remA + 1 : 1;
  const long long costA = W + (long long)kSegOverhead * (segA - 1);
  const int nfB = kNumSMs / R, wB = ceil_div(T, nfB);
  if (costA <= wB) {
    pl.n_chunks = nfA; pl.chunk_w = W; pl.rem0 = nfA * W; pl.rem_w = remA;
    pl.units_per_cta = W; pl.total_units = R * remA;
    pl.grid = nfA * R + ceil_div(pl.total_units, W);
    pl.max_seg = segA;
  } else {
    pl.n_chunks = ceil_div(T, wB); pl.chunk_w = wB; pl.rem0 = T; pl.rem_w = 0;
    pl.units_per_cta = wB; pl.total_units = 0;
    pl.grid = pl.n_chunks * R;
    pl.max_seg = 1;
  }
  return pl;
}
// A call is cut into at most two launches.  First k = row_tiles / 148 WHOLE row tiles per SM: every CTA
// starts at column tile 0 and they sweep the column images together, so a tile fetched from HBM by one
// CTA is an L2 hit for the other 147 (with ranges that start at scattered column phases the 63 MB of
// operand images are streamed by every CTA on its own schedule and no longer stay resident: 2.1 GB of
// DRAM reads per launch instead of ~0.4).  Then the remaining row tiles, in column chunks (tc_plan).
static int tc_launch_plans(int row_tiles, int T, TcPlan out[2]) {
  int n = 0;
  const int k = row_tiles / kNumSMs;
  if (k >= 1) {
    TcPlan& a = out[n++];
    a = TcPlan{};
    a.row_tile0 = 0; a.row_tiles = k * kNumSMs; a.R = a.row_tiles;
    a.n_chunks = 0; a.chunk_w = 0; a.rem0 = 0; a.rem_w = T;
    a.total_units = a.row_tiles * T; a.units_per_cta = k * T; a.grid = kNumSMs; a.max_seg = k;
  }
  const int rest = row_tiles - k * kNumSMs;
  if (rest > 0) {
    out[n] = tc_plan(rest, T);
    out[n].row_tile0 = k * kNumSMs;
    ++n;
  }
  return n;
}

// =======================================================================================
// exact median of the N^2 squared distances on the same pipeline
//   1. a deterministic sample of 2^20 pairs locates the hi-16-bit bin h of the median;
//   2. ONE full Gram pass (GEMM1 only) counts, per thread in registers, the values below the window
//      [(h-1)<<16, (h+2)<<16) of float bit patterns and histograms only the values inside it
//      (~3 % of them) with 64-bit reductions in global memory;
//   3. a single CTA finds rank (N^2-1)/2 inside the window.  If the rank falls outside (sample
//      off by > 65536 ulps, never observed) a flag is left clear and the robust two-pass radix
//      select of svgd_large.cu runs instead (its kernels exit at once when the flag is set).
// =======================================================================================
constexpr int kMedWindowBins = 3 * 65536;
constexpr int kMedSBufs = 6;        // S ring in TMEM (6 x 64 columns)
constexpr int kMedSample = 1 << 20;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

constexpr int kMedSampleGrid = kNumSMs, kMedSampleThreads = 1024;

// first index q of arr[0 .. 32 * per_lane) whose running sum exceeds `rank` (cum = the sum before q), found
// by one whole warp: a serial pass over per_lane consecutive entries per lane, a shuffle scan across
// the lanes, and a second serial pass inside the lane that holds the rank.  q = -1 if the total <= rank.
template <typename T>
__device__ __forceinline__ void warp_find_rank(const T* arr, int per_lane, unsigned long long rank, int& q, unsigned long long& cum) {
  const int lane = threadIdx.x & 31;
  unsigned long long s = 0;
  for (int i = 0; i < per_lane; ++i) s += (unsigned long long)arr[lane * per_lane + i];
  unsigned long long incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned long long excl = incl - s;
  const unsigned int hit = __ballot_sync(0xffffffffu, excl <= rank && rank < incl);
  q = -1; cum = 0;
  if (hit == 0) return;
  const int src = __ffs(hit) - 1;
  int qq = 0;
  unsigned long long cc = excl;
  if (lane == src) {
    for (int i = 0; i < per_lane; ++i) {
      const unsigned long long v = (unsigned long long)arr[lane * per_lane + i];
      if (cc + v > rank) { qq = lane * per_lane + i; break; }
      cc += v;
    }
  }
  q = __shfl_sync(0xffffffffu, qq, src);
  cum = __shfl_sync(0xffffffffu, cc, src);
}

// hist32[32768]: hi-16-bit histogram of the sample (non-negative floats: bits >> 16 < 32768).
// One CTA per SM keeps the whole histogram in shared memory (128 KB) and adds its non-empty bins to
// the global one at the end (the sample spreads over a few hundred bins: per-sample global atomics
// queue up on them).
__global__ void __launch_bounds__(kMedSampleThreads) med_sample_kernel(const float* __restrict__ x, const float* __restrict__ xn,
                                                                       int N, int D, int ld, unsigned int* __restrict__ hist32,
                                                                       int k_begin, int k_end) {
  extern __shared__ unsigned int hs[];   // [32768]
  for (int b = threadIdx.x; b < 32768; b += blockDim.x) hs[b] = 0u;
  __syncthreads();
  const bool vec = (D & 3) == 0 && (ld & 3) == 0 && ((((uintptr_t)x) & 15) == 0);   // rows are 16-byte aligned: independent 128-bit loads
  for (int k = k_begin + blockIdx.x * blockDim.x + threadIdx.x; k < k_end; k += gridDim.x * blockDim.x) {
    const uint32_t i = hash32(2u * k + 1u) % (uint32_t)N, j = hash32(2u * k + 0x9e3779b9u) % (uint32_t)N;
    const float* __restrict__ xi = x + (long long)i * ld;
    const float* __restrict__ xj = x + (long long)j * ld;
    float dot = 0.f;
    if (vec) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int d = 0; d < D; d += 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(xi + d)), b = __ldg(reinterpret_cast<const float4*>(xj + d));
        acc.x = fmaf(a.x, b.x, acc.x); acc.y = fmaf(a.y, b.y, acc.y); acc.z = fmaf(a.z, b.z, acc.z); acc.w = fmaf(a.w, b.w, acc.w);
      }
      dot = (acc.x + acc.y) + (acc.z + acc.w);
    } else {
      for (int d = 0; d < D; ++d) dot = fmaf(xi[d], xj[d], dot);
    }
    const float d2 = (i == j) ? 0.f : fmaxf((xn[j] - 2.0f * dot) + xn[i], 0.f);
    atomicAdd(&hs[__float_as_uint(d2) >> 16], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < 32768; b += blockDim.x) {
    const unsigned int c = hs[b];
    if (c) atomicAdd(&hist32[b], c);
  }
}

// state[0] = window start (bit pattern), state[1] = ok flag (cleared here), state[2] = window width.
// The median's position inside its hi-16 bin is interpolated from the sample counts; the window is
// that position +- the bit-pattern distance that holds 8 standard deviations of the sample median's
// quantile (0.5/sqrt(m)), clamped to [4096, kMedWindowBins/2] patterns.
__global__ void __launch_bounds__(1024) med_sample_select_kernel(unsigned int* hist32, uint32_t* state) {
  __shared__ unsigned int part[1024];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int q = warp; q < 1024; q += 32) {   // one warp per chunk of 32 bins: coalesced reads
    const unsigned int s = __reduce_add_sync(0xffffffffu, hist32[q * 32 + lane]);
    if (lane == 0) part[q] = s;
  }
  __syncthreads();
  if (warp == 0) {
    const unsigned long long rank = (kMedSample - 1) / 2;
    int q, b;
    unsigned long long cum, cum_b;
    warp_find_rank(part, 32, rank, q, cum);
    int bin = 32767;
    unsigned int in_bin = 1;
    if (q >= 0) {
      warp_find_rank(hist32 + q * 32, 1, rank - cum, b, cum_b);
      if (b >= 0) { bin = q * 32 + b; in_bin = hist32[bin]; cum += cum_b; }
    }
    if (lane == 0) {
      const double frac = ((double)(rank - cum) + 0.5) / (double)in_bin;               // position inside the bin
      const double centre = (double)bin * 65536.0 + frac * 65536.0;
      const double mass_per_pattern = (double)in_bin / (double)kMedSample / 65536.0;   // local density estimate
      double hw = 8.0 * (0.5 / sqrt((double)kMedSample)) / mass_per_pattern;
      hw = fmin(fmax(hw, 4096.0), (double)(kMedWindowBins / 2));
      double lo = centre - hw;
      if (lo < 0.0) lo = 0.0;
      state[0] = (uint32_t)lo;
      state[1] = 0u;
      state[2] = (uint32_t)(2.0 * hw);
    }
  }
  __syncthreads();
  for (int b = t; b < 32768; b += 1024) hist32[b] = 0u;
}

constexpr int kMedConsumers = kTcThreads - 128;   // threads of the three consumer warpgroups (warps 4..15)
struct MedTcParams {
  int N, Dp, T, row_begin, ksplit;
  int xb_stages;                     // operand ring depth (<= kMedXbStages), as many as fit beside the hit queues
  int kcount, kchunk;                // circular half band: T/2 + 1 column tiles per row tile, kchunk of them per CTA
  const float *xa_hi, *xa_lo, *xb_hi, *xb_lo, *xn;
  const float* x;                    // the particles themselves (row stride ld): the row tile is staged in TMEM from them
  int D, ld;
  const uint32_t* state;             // [0] window start
  unsigned long long* hist;          // [kMedWindowBins] window histogram, then [kMedWindowBins] = count below
};

enum { MB_A = 0, MB_XB_FULL = 1, MB_XB_EMPTY = 7, MB_XN_FULL = 13, MB_XN_EMPTY = 29, MB_S_FULL = 45, MB_S_EMPTY = 51, MB_COUNT = 57 };

// operand ring, |x_j|^2 ring, barriers + TMEM slot, then the consumers' hit queues: 32 window hits per thread (one per
// element of a 32-column half), interleaved by thread.  The row tile (A operand) lives in TMEM behind the S ring
// (6 x 64 + 2 Dp <= 512 columns): as in phi_tc_kernel, an SS-form MMA would fetch 6 KB per 32 cycles from shared memory.
__host__ __device__ inline size_t med_smem_bytes(int Dp, int stages) {
  return (size_t)stages * 2 * kTcBN * Dp * 4 + (size_t)kXnStages * kTcBN * 4 + 64 * 8 + 64 + (size_t)kMedConsumers * 32 * 4;
}
__host__ __device__ inline int med_xb_stages(int Dp) {
  for (int s = kMedXbStages; s >= 3; --s)
    if (med_smem_bytes(Dp, s) <= 227 * 1024) return s;
  return 0;
}

__global__ void __launch_bounds__(kTcThreads, 1) median_tc_kernel(const MedTcParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t xb_half = kTcBN * p.Dp * 4, xb_stage = 2 * xb_half;
  const uint32_t off_xb = 0;
  const int nstage = p.xb_stages;
  const uint32_t off_xn = off_xb + nstage * xb_stage;
  const uint32_t off_bars = (off_xn + kXnStages * kTcBN * 4 + 7) & ~7u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + off_bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + off_bars + 64 * 8);
  const uint32_t off_hitq = off_bars + 64 * 8 + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = p.row_begin + (blockIdx.x / p.ksplit) * kTcBM;
  // d2 is symmetric: in units of 128-point blocks (Tb = N/128 of them; a block is one row tile and
  // kBlk = 2 column tiles), row block i visits the circular half band of column blocks i, i+1, ...,
  // i+Tb/2 (mod Tb).  Every unordered block pair is met exactly once -- except the diagonal (offset 0)
  // and, for even Tb, the opposite block (offset Tb/2), which both of its rows meet -- so those two
  // count once and every other block twice.  All row tiles carry the same work, whatever row block
  // a rank owns.  p.kcount = kBlk * (Tb/2 + 1) column tiles per row tile, p.kchunk of them per CTA.
  constexpr int kBlk = kTcBM / kTcBN;
  const int itile = i0 / kTcBM;
  const int Tb = p.T / kBlk;
  const int kbase = (blockIdx.x % p.ksplit) * p.kchunk;
  const int T = max(0, min(p.kchunk, p.kcount - kbase));   // column tiles of this CTA
  auto col_tile = [&](int j) { return (kBlk * itile + kbase + j) % p.T; };

  if (threadIdx.x == 0) {
    mbar_init(&bars[MB_A], 4);       // the four warps of consumer group 0 stage the row tile in TMEM
    for (int s = 0; s < kMedXbStages; ++s) { mbar_init(&bars[MB_XB_FULL + s], 1); mbar_init(&bars[MB_XB_EMPTY + s], 1); }   // nstage of them are used
    // the |x_j|^2 slices ride in their own deep ring, recycled by the 4 consumer warps of a tile, so an
    // operand stage is free as soon as its MMAs retire
    for (int s = 0; s < kXnStages; ++s) { mbar_init(&bars[MB_XN_FULL + s], 1); mbar_init(&bars[MB_XN_EMPTY + s], 4); }
    for (int b = 0; b < kMedSBufs; ++b) { mbar_init(&bars[MB_S_FULL + b], 1); mbar_init(&bars[MB_S_EMPTY + b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tAhi = tmem + kMedSBufs * kTcBN, tAlo = tAhi + p.Dp;

  if (warp == 0) {
    if (lane == 0 && T > 0) {
      for (int j = 0; j < T; ++j) {
        const int sx = j % nstage, sn = j % kXnStages;
        mbar_wait(&bars[MB_XB_EMPTY + sx], ((j / nstage) & 1) ^ 1);
        unsigned char* xb = smem + off_xb + sx * xb_stage;
        mbar_expect_tx(&bars[MB_XB_FULL + sx], xb_stage);
        const long long jt = col_tile(j);
        bulk_g2s(xb, p.xb_hi + jt * kTcBN * p.Dp, xb_half, &bars[MB_XB_FULL + sx]);
        bulk_g2s(xb + xb_half, p.xb_lo + jt * kTcBN * p.Dp, xb_half, &bars[MB_XB_FULL + sx]);
        mbar_wait(&bars[MB_XN_EMPTY + sn], ((j / kXnStages) & 1) ^ 1);
        mbar_expect_tx(&bars[MB_XN_FULL + sn], kTcBN * 4);
        bulk_g2s(smem + off_xn + sn * kTcBN * 4, p.xn + jt * kTcBN, kTcBN * 4, &bars[MB_XN_FULL + sn]);
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc1 = make_idesc_tf32(kTcBM, kTcBN);
      const uint32_t sbo1 = (uint32_t)(p.Dp / 4) * 128u;
      const int ks1 = p.Dp / 8;
      const uint64_t dX = make_desc(smem_u32(smem + off_xb), 128, sbo1);
      const uint32_t hiA = (uint32_t)(dX >> 32), loX0 = (uint32_t)dX;
      const uint32_t xb_stage16 = xb_stage >> 4, xb_half16 = xb_half >> 4;
      if (T > 0) mbar_wait(&bars[MB_A], 0);
      for (int j = 0; j < T; ++j) {
        const int b = j % kMedSBufs, sx = j % nstage;
        mbar_wait(&bars[MB_XB_FULL + sx], (j / nstage) & 1);
        mbar_wait(&bars[MB_S_EMPTY + b], ((j / kMedSBufs) & 1) ^ 1);
        tc_fence_after();
        const uint32_t xh = loX0 + sx * xb_stage16, xl = xh + xb_half16;
        const uint32_t tSb = tmem + b * kTcBN;
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if (kk < ks1) {
              const uint64_t bh = desc_lo_hi(xh + kk * 16, hiA), bl = desc_lo_hi(xl + kk * 16, hiA);
              mma_ts(tSb, tAhi + kk * 8, bh, idesc1, kk > 0 ? 1u : 0u);
              mma_ts(tSb, tAhi + kk * 8, bl, idesc1, 1u);
              mma_ts(tSb, tAlo + kk * 8, bh, idesc1, 1u);
            }
          }
          tc_commit(&bars[MB_S_FULL + b]);
          tc_commit(&bars[MB_XB_EMPTY + sx]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // three consumer warpgroups: tile j is handled by group j % 3, from S buffer j % 6
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float xn_i = p.xn[i0 + row];
    const int jdiag = (i0 + row) / kTcBN;
    const uint32_t win_lo = p.state[0], win_n = p.state[2];
    // this thread's hit queue: entry i at hitq[i * kMedConsumers + consumer index] (conflict-free across a warp)
    const uint32_t q0 = smem_u32(smem + off_hitq) + (uint32_t)(threadIdx.x - 128) * 4u;
    if (wg == 0 && T > 0) {
      // stage the row tile: thread <-> row, the hi / lo TF32 split of its Dp values (the bits tc_prep_kernel writes)
      const float* __restrict__ xr = p.x + (long long)(i0 + row) * p.ld;
      const bool vec = (p.D & 3) == 0 && (p.ld & 3) == 0 && ((((uintptr_t)p.x) & 15) == 0);
      for (int k0 = 0; k0 < p.Dp; k0 += 8) {
        float v[8];
        if (vec && k0 + 8 <= p.D) {
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(xr + k0)), a1 = __ldg(reinterpret_cast<const float4*>(xr + k0 + 4));
          v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = (k0 + c < p.D) ? __ldg(xr + k0 + c) : 0.f;
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float h = tf32_hi(v[c]);
          hi[c] = __float_as_uint(h);
          lo[c] = __float_as_uint(tf32_lo(v[c], h));
        }
        tmem_st8(tAhi + lane_base + k0, hi);
        tmem_st8(tAlo + lane_base + k0, lo);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[MB_A]);
    }
    unsigned int below = 0;
    for (int j = wg; j < T; j += 3) {
      const int b = j % kMedSBufs, sn = j % kXnStages;
      const float* xnj = reinterpret_cast<const float*>(smem + off_xn + sn * kTcBN * 4);
      mbar_wait(&bars[MB_XN_FULL + sn], (j / kXnStages) & 1);
      mbar_wait(&bars[MB_S_FULL + b], (j / kMedSBufs) & 1);
      tc_fence_after();
      const int kb = (kbase + j) / kBlk;   // block offset inside the band
      const int cdiag = (col_tile(j) == jdiag) ? ((i0 + row) & (kTcBN - 1)) : -1;
      const uint32_t wgt = (kb == 0 || 2 * kb == Tb) ? 1u : 2u;
      const unsigned long long wgt64 = wgt;
      uint32_t nbelow = 0;                 // number of values below the window in this tile
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_base + b * kTcBN + half * 32, r);
        tmem_wait_ld();
        if (cdiag >= half * 32 && cdiag < half * 32 + 32) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (half * 32 + c == cdiag) r[c] = __float_as_uint(0.5f * (xnj[half * 32 + c] + xn_i));  // => d2 = 0
        }
        // Per distance: d2 (two at a time on the packed FP32 pipe, each lane rounded like the scalar fmaf / add), its
        // position relative to the window (rel; bit 31 <=> below it, both bit patterns being < 2^31: one IMAD.HI adds
        // it to one of four independent counters), and -- for the ~3 % inside the window -- a predicated store of rel
        // into the thread's queue.  The 64-bit histogram updates (address arithmetic, descriptor, reduction) run
        // afterwards, for the hits only: per-element predicated `red.global` compiled to a branch + descriptor moves
        // around EVERY element, and the earlier carry-chain form (sub.cc / subc) serialised the 32 elements.
        uint32_t qa = q0;
        uint32_t nb[4] = {0u, 0u, 0u, 0u};
        const float4* __restrict__ xq = reinterpret_cast<const float4*>(xnj + half * 32);
        const float2 m2 = make_float2(-2.0f, -2.0f), xi2 = make_float2(xn_i, xn_i);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 xv = xq[c4];
          const float2 s01 = make_float2(__uint_as_float(r[4 * c4]), __uint_as_float(r[4 * c4 + 1]));
          const float2 s23 = make_float2(__uint_as_float(r[4 * c4 + 2]), __uint_as_float(r[4 * c4 + 3]));
          const float2 e01 = __fadd2_rn(__ffma2_rn(m2, s01, make_float2(xv.x, xv.y)), xi2);   // (x_j - 2s) + x_i, 2s exact
          const float2 e23 = __fadd2_rn(__ffma2_rn(m2, s23, make_float2(xv.z, xv.w)), xi2);
          const float d2v[4] = {fmaxf(e01.x, 0.f), fmaxf(e01.y, 0.f), fmaxf(e23.x, 0.f), fmaxf(e23.y, 0.f)};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t rel = __float_as_uint(d2v[k]) - win_lo;
            nb[k] = __umulhi(rel, 2u) + nb[k];
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.lt.u32 p, %1, %2;\n\t"
                "@p st.shared.u32 [%0], %1;\n\t"
                "@p add.u32 %0, %0, %3;\n\t}"
                : "+r"(qa)
                : "r"(rel), "r"(win_n), "n"(kMedConsumers * 4)
                : "memory");
          }
        }
        nbelow += (nb[0] + nb[1]) + (nb[2] + nb[3]);
        for (uint32_t a = q0; a < qa; a += kMedConsumers * 4) {
          uint32_t rel;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rel) : "r"(a) : "memory");
          asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p.hist + rel), "l"(wgt64) : "memory");
        }
      }
      below += nbelow * wgt;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars[MB_S_EMPTY + b]);
        mbar_arrive(&bars[MB_XN_EMPTY + sn]);
      }
    }
    below = __reduce_add_sync(0xffffffffu, below);
    if (lane == 0 && below) atomicAdd(&p.hist[kMedWindowBins], (unsigned long long)below);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// rank (N^2-1)/2 inside the window; state[1] = 1 and the median on success
__global__ void __launch_bounds__(1024) med_window_select_kernel(const unsigned long long* hist, uint32_t* state,
                                                                 long long n_total, float* median_out) {
  __shared__ unsigned long long part[1024];
  constexpr int PER = kMedWindowBins / 1024;  // 192 bins per chunk
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int q = warp; q < 1024; q += 32) {   // one warp per chunk: coalesced reads
    unsigned long long s = 0;
#pragma unroll
    for (int b = 0; b < PER / 32; ++b) s += hist[q * PER + b * 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) part[q] = s;
  }
  __syncthreads();
  if (warp == 0) {
    const unsigned long long below = hist[kMedWindowBins];
    const unsigned long long k = (unsigned long long)((n_total - 1) / 2);
    int ok = 0;
    if (k >= below) {
      int ch, b;
      unsigned long long cum, cum_b;
      warp_find_rank(part, 32, k - below, ch, cum);          // -1: the rank lies above the window
      if (ch >= 0) {
        warp_find_rank(hist + ch * PER, PER / 32, k - below - cum, b, cum_b);
        if (b >= 0) {
          ok = 1;
          if (lane == 0) {
            const uint32_t bits = state[0] + (uint32_t)(ch * PER + b);
            state[3] = bits;
            if (median_out) *median_out = __uint_as_float(bits);
          }
        }
      }
    }
    if (lane == 0) state[1] = ok ? 1u : 0u;
  }
}

bool median_tc_supported(int N, int D) {
  const int Dp = round_up(D, 8);
  if (N % kTcBM || N < 1024 || Dp > 64) return false;
  return med_xb_stages(Dp) > 0;
}
size_t median_tc_workspace(int N, int D) {
  const size_t Dp = round_up(D, 8);
  return sizeof(float) * ((size_t)N * (2 * Dp + 1) + 64) + sizeof(unsigned int) * 32768;
}

size_t median_tc_sample_hist_offset(int N, int D) {
  const size_t Dp = round_up(D, 8);
  return sizeof(float) * ((size_t)(N + 63) / 64 * 64 + 2 * (size_t)N * Dp);
}

// operand images + this rank's share of the sample (sample_begin/end; every rank holds the same gathered X)
int median_tc_prepare(const dust_median_args* a, void* workspace, cudaStream_t stream) {
  const int N = a->N, D = a->D, Dp = round_up(D, 8);
  float* ws = (float*)workspace;
  float* xn = ws;    ws += (N + 63) / 64 * 64;
  float* x_hi = ws;  ws += (size_t)N * Dp;
  float* x_lo = ws;  ws += (size_t)N * Dp;
  unsigned int* hist32 = (unsigned int*)ws;
  DUST_CUDA_OK(cudaMemsetAsync(hist32, 0, sizeof(unsigned int) * 32768, stream));
  const int ld = a->ld > 0 ? a->ld : D;
  int rc = tc_prep(a->x, nullptr, N, D, ld, Dp, 0, x_hi, x_lo, nullptr, nullptr, xn, stream);
  if (rc != DUST_OK) return rc;
  DUST_CUDA_OK(cudaFuncSetAttribute(med_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(unsigned int) * 32768)));
  {
    DUST_TIMED("med_sample_kernel", stream);
    const int k0 = a->sample_end > a->sample_begin ? a->sample_begin : 0;
    const int k1 = a->sample_end > a->sample_begin ? a->sample_end : kMedSample;
    const int grid = min(kMedSampleGrid, max(1, ceil_div(k1 - k0, kMedSampleThreads)));
    med_sample_kernel<<<grid, kMedSampleThreads, sizeof(unsigned int) * 32768, stream>>>(a->x, xn, N, D, ld, hist32, k0, k1);
  }
  DUST_LAUNCH_OK("med_sample_kernel");
  return DUST_OK;
}

// window from the (summed) sample histogram
int median_tc_window(const dust_median_args* a, void* workspace, cudaStream_t stream) {
  unsigned int* hist32 = (unsigned int*)((char*)workspace + median_tc_sample_hist_offset(a->N, a->D));
  {
    DUST_TIMED("med_sample_select_kernel", stream);
    med_sample_select_kernel<<<1, 1024, 0, stream>>>(hist32, a->selected + 4);
  }
  DUST_LAUNCH_OK("med_sample_select_kernel");
  return DUST_OK;
}

int median_tc_count(const dust_median_args* a, void* workspace, cudaStream_t stream) {
  const int N = a->N, D = a->D, Dp = round_up(D, 8);
  float* ws = (float*)workspace;
  float* xn = ws;    ws += (N + 63) / 64 * 64;
  float* x_hi = ws;  ws += (size_t)N * Dp;
  float* x_lo = ws;
  const int r0 = a->row_begin, r1 = a->row_end > 0 ? a->row_end : N;
  const int row_tiles = (r1 - r0) / kTcBM;
  if (row_tiles == 0) return DUST_OK;
  // half band of (N/128)/2 + 1 column blocks (two 64-wide tiles each) per row tile, cut into ksplit chunks
  // (the chunks only feed integer counters: any split gives the same histogram)
  const int T = N / kTcBN, kcount = (kTcBM / kTcBN) * ((N / kTcBM) / 2 + 1);
  int ksplit = 1, kchunk = kcount;
  double best_cost = 1e30;
  for (int ks = 1; ks <= 64 && ks <= kcount; ++ks) {
    const int chunk = ceil_div(kcount, ks);
    if ((ks - 1) * chunk >= kcount) continue;   // would leave a CTA without tiles
    const double cost = (double)ceil_div((long long)row_tiles * ks, kNumSMs) * (chunk + 3.0);  // + per-CTA prologue
    if (cost < best_cost - 1e-9) { best_cost = cost; ksplit = ks; kchunk = chunk; }
  }
  const int stages = med_xb_stages(Dp);
  MedTcParams p{N, Dp, T, r0, ksplit, stages, kcount, kchunk, x_hi, x_lo, x_hi, x_lo, xn, a->x, a->D, a->ld > 0 ? a->ld : a->D,
                a->selected + 4, a->hist};
  const size_t smem = med_smem_bytes(Dp, stages);
  DUST_CUDA_OK(cudaFuncSetAttribute(median_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    DUST_TIMED("median_tc_kernel", stream);
    median_tc_kernel<<<row_tiles * ksplit, kTcThreads, smem, stream>>>(p);
  }
  DUST_LAUNCH_OK("median_tc_kernel");
  return DUST_OK;
}

int median_tc_select(const dust_median_args* a, float* median_out, cudaStream_t stream) {
  {
    DUST_TIMED("med_window_select_kernel", stream);
    med_window_select_kernel<<<1, 1024, 0, stream>>>(a->hist, a->selected + 4, (long long)a->N * a->N, median_out);
  }
  DUST_LAUNCH_OK("med_window_select_kernel");
  // leave the buffer clean for the fallback / next call
  DUST_CUDA_OK(cudaMemsetAsync(a->hist, 0, sizeof(unsigned long long) * (kMedWindowBins + 1), stream));
  return DUST_OK;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// Which form the kernel takes for a shape: (row tile in TMEM + bf16 P_lo term) when the TMEM columns
// (S/P_hi ring, two P_lo buffers of 32, two O buffers, A_hi | A_lo) and the shared memory (no A, a bf16 V copy per
// stage) allow it, else (row tile in shared memory + TF32 P_lo term).  mode < 0: neither fits.
struct TcMode { int a_tmem, bf16lo; bool ok; };
static TcMode tc_mode(int Dp, int NV) {
  const size_t cap = 227 * 1024;
  const bool fast_ok = DUST_TC_BF16LO && kSBufs * kTcBN + 2 * (kTcBN / 2) + 2 * NV + 2 * Dp <= 512 &&
                       tc_smem_layout(Dp, NV, true, true).total <= cap;
  if (fast_ok && !getenv("DUST_B200_TC_A_SMEM")) return TcMode{1, 1, true};
  const bool plain_ok = (kSBufs + 2) * kTcBN + 2 * NV <= 512 && tc_smem_layout(Dp, NV, false, false).total <= cap;
  return TcMode{0, 0, plain_ok};
}

bool phi_tc_supported(const dust_phi_args* a) {
  if (a->B != 1 || a->per_dim) return false;
  const int r0 = a->row_begin, r1 = a->row_end > 0 ? a->row_end : a->N;
  if (a->N % kTcBM || r0 % kTcBM || r1 % kTcBM || a->N < 1024) return false;
  const int Dp = round_up(a->D, 8), NV = round_up(2 * a->D, 16);
  if (Dp > 64) return false;
  return tc_mode(Dp, NV).ok;
}

// head (shared with the median pass): |x|^2, X images hi / lo; then V^T images hi / lo, their bf16 copy, the scratch
size_t phi_tc_workspace(const dust_phi_args* a) {
  const size_t N = a->N, Dp = round_up(a->D, 8), NV = round_up(2 * a->D, 16);
  const int rows = (a->row_end > 0 ? a->row_end : a->N) - a->row_begin;
  TcPlan pl[2];
  const int n = tc_launch_plans(rows / kTcBM, a->N / kTcBN, pl);
  size_t slots = 0;
  for (int i = 0; i < n; ++i) slots += (size_t)pl[i].grid * pl[i].max_seg;
  return sizeof(float) * (N * (2 * Dp + 2 * NV + 1) + N * NV / 2 + 64 + slots * (NV + 2) * kTcBM);
}

int phi_tc(const dust_phi_args* a, cudaStream_t stream) {
  const int N = a->N, D = a->D, Dp = round_up(D, 8), NV = round_up(2 * D, 16);
  DUST_REQUIRE(a->workspace && a->workspace_bytes >= phi_tc_workspace(a), DUST_ERR_WORKSPACE,
               "dust_svgd_phi: tensor-core path needs %zu bytes of workspace", phi_tc_workspace(a));
  float* ws = (float*)a->workspace;
  float* xn = ws;              ws += (N + 63) / 64 * 64;
  float* x_hi = ws;            ws += (size_t)N * Dp;
  float* x_lo = ws;            ws += (size_t)N * Dp;
  float* vb_hi = ws;           ws += (size_t)N * NV;
  float* vb_lo = ws;           ws += (size_t)N * NV;
  float* vb_bf = ws;           ws += (size_t)N * NV / 2;
  float* oacc = ws;
  const TcMode mode = tc_mode(Dp, NV);
  DUST_REQUIRE(mode.ok, DUST_ERR_UNSUPPORTED, "dust_svgd_phi: D=%d does not fit the tensor-core kernel", D);
  const int ld = a->ld > 0 ? a->ld : D;
  int rc = tc_prep(a->x, a->score, N, D, ld, Dp, NV, x_hi, x_lo, vb_hi, vb_lo, xn, stream, a->x_prepared != 0, mode.bf16lo ? vb_bf : nullptr);
  if (rc != DUST_OK) return rc;
  const int r0 = a->row_begin, r1 = a->row_end > 0 ? a->row_end : N;
  const int row_tiles = (r1 - r0) / kTcBM;
  if (row_tiles == 0) return DUST_OK;
  TcParams p;
  p.N = N; p.D = D; p.Dp = Dp; p.NV = NV; p.T = N / kTcBN; p.row_begin = r0; p.ld = ld;
  p.a_tmem = mode.a_tmem; p.bf16lo = mode.bf16lo; p.vb_bf = vb_bf;
  // a 128-row tile and a 64-row tile of the core-matrix image are both whole 8-row groups in row order:
  // ONE image serves as the A operand (row tiles) and as the B operand (column tiles)
  p.xa_hi = x_hi; p.xa_lo = x_lo; p.xb_hi = x_hi; p.xb_lo = x_lo; p.vb_hi = vb_hi; p.vb_lo = vb_lo; p.xn = xn; p.x = a->x;
  p.gamma = a->gamma; p.c1 = a->c1; p.c2 = a->c2; p.gamma_dev = a->gamma_dev; p.lr = a->lr; p.phi = a->phi; p.x_out = a->x_out; p.oacc = oacc;
  const TcSmem L = tc_smem_layout(Dp, NV, mode.a_tmem != 0, mode.bf16lo != 0);
  const bool fast = mode.a_tmem != 0;
  DUST_CUDA_OK(cudaFuncSetAttribute(fast ? phi_tc_kernel<true> : phi_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  TcPlan plans[2];
  const int n_plans = tc_launch_plans(row_tiles, p.T, plans);
  for (int i = 0; i < n_plans; ++i) {
    const TcPlan& pl = plans[i];
    p.row_begin = r0 + pl.row_tile0 * kTcBM;
    p.units_per_cta = pl.units_per_cta; p.total_units = pl.total_units; p.max_seg = pl.max_seg;
    p.R = pl.R; p.n_chunks = pl.n_chunks; p.chunk_w = pl.chunk_w; p.rem0 = pl.rem0; p.rem_w = pl.rem_w;
    {
      DUST_TIMED("phi_tc_kernel", stream);
      if (fast) phi_tc_kernel<true><<<pl.grid, kTcThreads, L.total, stream>>>(p);
      else phi_tc_kernel<false><<<pl.grid, kTcThreads, L.total, stream>>>(p);
    }
    DUST_LAUNCH_OK("phi_tc_kernel");
    {
      DUST_TIMED("phi_tc_finish_kernel", stream);
      phi_tc_finish_kernel<<<dim3(pl.row_tiles, ceil_div(D, 8)), kTcBM, 0, stream>>>(p);
    }
    DUST_LAUNCH_OK("phi_tc_finish_kernel");
    p.oacc += (size_t)pl.grid * pl.max_seg * (NV + 2) * kTcBM;
  }
  return DUST_OK;
}

}  // namespace dust

// Which form of phi_tc_kernel a dimension D takes (no device needed): bit 0 = the row tile of GEMM1 lives in TMEM,
// bit 1 = GEMM2's P_lo V correction term runs as kind::f16 on bf16 copies; -1 = D does not fit the tensor-core kernel.
extern "C" int dust_phi_tc_mode(int32_t D) {
  if (D <= 0) return -1;
  const int Dp = dust::round_up(D, 8), NV = dust::round_up(2 * D, 16);
  if (Dp > 64) return -1;
  const dust::TcMode m = dust::tc_mode(Dp, NV);
  return m.ok ? (m.a_tmem | (m.bf16lo << 1)) : -1;
}

// How dust_svgd_phi's tensor-core path partitions `row_tiles` x `col_tiles` tile pairs over its launches (no launch,
// no device needed): plan[i] = {grid, R, n_chunks, chunk_w, rem0, rem_w, units_per_cta, total_units, max_seg,
// row_tile0, row_tiles} of launch i; returns the number of launches.  Tests emulate the kernel's segment iterator on
// it to prove every tile pair is visited exactly once and the finish kernel finds every scratch slot.
extern "C" int dust_phi_tc_plan(int32_t row_tiles, int32_t col_tiles, int32_t plan[2][11]) {
  if (row_tiles <= 0 || col_tiles < 16 || plan == nullptr) return 0;
  dust::TcPlan pl[2];
  const int n = dust::tc_launch_plans(row_tiles, col_tiles, pl);
  for (int i = 0; i < n; ++i) {
    const dust::TcPlan& q = pl[i];
    const int v[11] = {q.grid, q.R, q.n_chunks, q.chunk_w, q.rem0, q.rem_w, q.units_per_cta, q.total_units, q.max_seg,
                       q.row_tile0, q.row_tiles};
    for (int c = 0; c < 11; ++c) plan[i][c] = v[c];
  }
  return n;
}
