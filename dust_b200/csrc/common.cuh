// Shared helpers for libdust_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dust_b200.h"

namespace dust {

// thread-local last-error message (dust_last_error)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define DUST_REQUIRE(cond, code, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      ::dust::set_error(__VA_ARGS__);      \
      return (code);                       \
    }                                      \
  } while (0)

#define DUST_CUDA_OK(call)                                         \
  do {                                                             \
    cudaError_t _e = (call);                                       \
    if (_e != cudaSuccess) return ::dust::cuda_fail(_e, #call);    \
  } while (0)

#define DUST_LAUNCH_OK(name)                                       \
  do {                                                             \
    ::dust::count_launch();                                        \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return ::dust::cuda_fail(_e, name);     \
  } while (0)

// Optional per-kernel timing (dust_profiler_enable): a KernelTimer placed in the scope of a
// launch records CUDA events on the launch stream before and after it.  Off by default.
void count_launch();
struct KernelTimer {
  KernelTimer(const char* name, cudaStream_t stream);
  ~KernelTimer();
  int slot;
  cudaStream_t stream;
};
#define DUST_TIMED(name, stream) ::dust::KernelTimer _dust_timer_##__LINE__(name, stream)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200

// device copy of the model description (passed by value as a kernel parameter)
struct ModelParams {
  int kind;
  float dt;
  float g, max_torque, max_speed_pend, w_angle, w_speed, default_length, default_mass;
  float max_accel, max_speed;
  float target[4], w_state[4], w_term[4], w_ctrl[2];
  float w_obs, inv_cell, c_offset[2];
  int grid_nx, grid_ny, can_crash, with_obstacle;
  float grid_xmax, grid_ymax;   // (float)(grid_nx - 1), (float)(grid_ny - 1): the clamp bounds of the cell lookup
  float pend_c1, pend_c2;       // default-parameter dynamics coefficients, formed in double on the host (models.cuh)
  const uint32_t* grid_bits;
};

inline ModelParams to_params(const dust_model_desc& d) {
  ModelParams m;
  m.kind = d.kind; m.dt = d.dt; m.g = d.g; m.max_torque = d.max_torque;
  m.max_speed_pend = d.max_speed_pend; m.w_angle = d.w_angle; m.w_speed = d.w_speed;
  m.default_length = d.default_length; m.default_mass = d.default_mass;
  m.max_accel = d.max_accel; m.max_speed = d.max_speed;
  for (int i = 0; i < 4; ++i) { m.target[i] = d.target[i]; m.w_state[i] = d.w_state[i]; m.w_term[i] = d.w_term[i]; }
  m.w_ctrl[0] = d.w_ctrl[0]; m.w_ctrl[1] = d.w_ctrl[1];
  m.w_obs = d.w_obs; m.inv_cell = d.inv_cell; m.c_offset[0] = d.c_offset[0]; m.c_offset[1] = d.c_offset[1];
  m.grid_nx = d.grid_nx; m.grid_ny = d.grid_ny; m.can_crash = d.can_crash; m.with_obstacle = d.with_obstacle;
  m.grid_xmax = (float)(d.grid_nx - 1); m.grid_ymax = (float)(d.grid_ny - 1);
  {
    // python-float defaults: the quotients are formed in double before the cast (pendulum.py:93-97); done here once
    // instead of with FP64 instructions in every thread
    const double l = (double)d.default_length, ms = (double)d.default_mass;
    m.pend_c1 = (float)(-3.0 * (double)d.g / (2.0 * l));
    m.pend_c2 = (float)(3.0 / (ms * l * l));
  }
  m.grid_bits = d.grid_bits;
  return m;
}

int validate_model(const dust_model_desc* d);
inline int model_ds(int kind) { return kind == DUST_MODEL_PENDULUM ? 2 : 4; }
inline int model_da(int kind) { return kind == DUST_MODEL_PENDULUM ? 1 : 2; }
inline int model_dp(int kind) { return kind == DUST_MODEL_PENDULUM ? 2 : 1; }

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------
// mbarrier + 1-D bulk copy (TMA) helpers shared by the rollout and the tensor-core kernels
// ---------------------------------------------------------------------------------------
constexpr uint32_t kSpinCap = 1u << 28;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  for (uint32_t spin = 0; spin < kSpinCap; ++spin) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();  // a lost arrival must surface as a launch failure, never as a hang
}
// the same on shared-window addresses the caller converted once (a generic -> shared conversion of dynamic shared
// memory costs a few instructions each time; hot loops keep the 32-bit addresses)
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; spin < kSpinCap; ++spin) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// a wait that is allowed to be late (a warp that drains a buffer every few dozen tile times): sleeps between polls so
// that its try_wait / branch pairs do not take issue slots from the warps that share its scheduler
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t ns = 256) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  for (uint32_t spin = 0; spin < kSpinCap; ++spin) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    __nanosleep(ns);
  }
  __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
#endif

}  // namespace dust
