// Microbenchmark: issue cost of packed FP32 (FFMA2/FADD2/FMUL2) on sm_100a, alone and mixed with ALU-pipe work.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb_ffma2 mb_ffma2.cu && ./mb_ffma2
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096, CH = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int alu) {
  float x[CH]; float2 y[CH]; unsigned u[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x + i; y[i] = make_float2(x[i], x[i] + 1.f); u[i] = threadIdx.x * 7 + i; }
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0 || MODE == 2) x[i] = fmaf(x[i], a, b);                 // scalar FFMA
      if (MODE == 1 || MODE == 3) y[i] = __ffma2_rn(y[i], a2, b2);         // packed FFMA2
      if (MODE == 4) { x[i] = fmaf(x[i], a, b); y[i].x = fmaf(y[i].x, a, b); }  // 2 scalar FFMA (same flops as MODE 1)
      if (MODE == 2 || MODE == 3) u[i] = (u[i] ^ (u[i] >> 3)) + alu;      // ALU-pipe work beside it (SHF+LOP3+IADD)
      if (MODE == 5) { y[i] = __ffma2_rn(y[i], a2, b2); x[i] = fminf(fmaxf(x[i], a), b) ; }  // FFMA2 + 2 FMNMX
      if (MODE == 6) { x[i] = fmaf(x[i], a, b); y[i].x = fmaf(y[i].x, a, b); u[i] = __float_as_uint(fminf(fmaxf(__uint_as_float(u[i]), a), b)); }
    }
  }
  float s = 0; unsigned v = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) { s += x[i] + y[i].x + y[i].y; v ^= u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + v;
}

template <int MODE> void run(const char* name, double flops_per_iter_thread) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, 3);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<MODE><<<148 * 8, 256>>>(out, 1.0001f, 0.5f, 3);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double threads = 148.0 * 8 * 256;
  // cycles per loop body per SMSP-warp slot: total warp-iterations / (SMSPs * clock)
  const double warp_iters = threads / 32 * ITER;   // each does CH ops
  const double clk = 1.965e9;
  const double cyc_per_iter_per_smsp = ms * 1e-3 * clk / (warp_iters / (148.0 * 4));
  printf("%-44s %8.3f ms  %7.2f TFLOP/s  %6.2f cycles per %d-op group per SMSP (%.2f / op)\n", name, ms,
         flops_per_iter_thread * threads * ITER / (ms * 1e-3) / 1e12, cyc_per_iter_per_smsp, CH, cyc_per_iter_per_smsp / CH);
  cudaFree(out);
}
int main() {
  run<0>("8 FFMA", 16);
  run<1>("8 FFMA2", 32);
  run<4>("16 FFMA (same flops as 8 FFMA2)", 32);
  run<2>("8 FFMA + 8x(SHF,LOP3,IADD)", 16);
  run<3>("8 FFMA2 + 8x(SHF,LOP3,IADD)", 32);
  run<5>("8 FFMA2 + 16 FMNMX", 32);
  run<6>("16 FFMA + 16 FMNMX", 32);
  return 0;
}
