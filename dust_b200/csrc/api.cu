// Library plumbing: thread-local error message, model validation, build info.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dust {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return DUST_ERR_CUDA;
}

int validate_model(const dust_model_desc* d) {
  DUST_REQUIRE(d != nullptr, DUST_ERR_INVALID_ARG, "model description is NULL");
  DUST_REQUIRE(d->kind == DUST_MODEL_PENDULUM || d->kind == DUST_MODEL_PARTICLE, DUST_ERR_UNSUPPORTED,
               "unknown model kind %d: only the pendulum and the 2-D particle have device kernels", d->kind);
  DUST_REQUIRE(d->dt > 0.f, DUST_ERR_INVALID_ARG, "Delta t must be greater than zero.");
  if (d->kind == DUST_MODEL_PARTICLE && (d->with_obstacle || d->can_crash)) {
    DUST_REQUIRE(d->grid_bits != nullptr && d->grid_nx > 0 && d->grid_ny > 0, DUST_ERR_INVALID_ARG,
                 "particle model with obstacles needs an occupancy grid");
    DUST_REQUIRE((size_t)d->grid_nx * d->grid_ny <= (size_t)1 << 20, DUST_ERR_UNSUPPORTED,
                 "occupancy grid larger than 2^20 cells does not fit in shared memory");
  }
  return DUST_OK;
}

// ---- launch counter and optional per-kernel event timing ---------------------------------
static unsigned long long g_launches = 0;
void count_launch() { ++g_launches; }

constexpr int kMaxSlots = 8192;
static bool g_prof_on = false;
static int g_prof_n = 0;
static cudaEvent_t g_ev0[kMaxSlots], g_ev1[kMaxSlots];
static const char* g_name[kMaxSlots];
static bool g_ev_init = false;

KernelTimer::KernelTimer(const char* name, cudaStream_t s) : slot(-1), stream(s) {
  if (!g_prof_on || g_prof_n >= kMaxSlots) return;
  if (!g_ev_init) {
    for (int i = 0; i < kMaxSlots; ++i) { cudaEventCreate(&g_ev0[i]); cudaEventCreate(&g_ev1[i]); }
    g_ev_init = true;
  }
  slot = g_prof_n++;
  g_name[slot] = name;
  cudaEventRecord(g_ev0[slot], stream);
}
KernelTimer::~KernelTimer() {
  if (slot >= 0) cudaEventRecord(g_ev1[slot], stream);
}

}  // namespace dust

extern "C" void dust_profiler_enable(int on) { dust::g_prof_on = on != 0; }
extern "C" void dust_profiler_reset(void) { dust::g_prof_n = 0; }
extern "C" unsigned long long dust_launch_count(void) { return dust::g_launches; }
// Synchronises the device and writes one line per kernel name: "<name> <launches> <total_ms>\n".
extern "C" int dust_profiler_report(char* buf, size_t cap) {
  using namespace dust;
  if (!buf || cap == 0) return DUST_ERR_INVALID_ARG;
  DUST_CUDA_OK(cudaDeviceSynchronize());
  const char* names[64];
  int counts[64];
  double totals[64];
  int nn = 0;
  for (int i = 0; i < g_prof_n; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_ev0[i], g_ev1[i]) != cudaSuccess) continue;
    int j = 0;
    for (; j < nn; ++j)
      if (strcmp(names[j], g_name[i]) == 0) break;
    if (j == nn) {
      if (nn == 64) continue;
      names[nn] = g_name[i]; counts[nn] = 0; totals[nn] = 0.0; ++nn;
    }
    counts[j] += 1;
    totals[j] += ms;
  }
  size_t off = 0;
  buf[0] = 0;
  for (int j = 0; j < nn; ++j) {
    const int w = snprintf(buf + off, cap - off, "%s %d %.6f\n", names[j], counts[j], totals[j]);
    if (w < 0 || (size_t)w >= cap - off) break;
    off += (size_t)w;
  }
  return DUST_OK;
}

extern "C" int dust_abi_version(void) { return DUST_B200_ABI_VERSION; }
extern "C" const char* dust_last_error(void) { return dust::g_err; }
extern "C" const char* dust_build_info(void) {
  return "libdust_b200 sm_100a (nvcc " __DATE__ " " __TIME__ ")";
}
