// The exchange step of row-block sharded SVGD over NVLink peer memory (one process per GPU on one NVSwitch box).
//
// Reference: the all-gather of particles and scores that precedes phi when X is split by row blocks
// (dust/inference/svgd.py:127-135 evaluates phi against ALL particles).  Instead of a library collective every
// rank keeps its rows [X | score] in a slab other ranks can read (CUDA IPC mapping), announces a new version with
// one flag store per peer, and PULLS the other ranks' rows straight into its gathered [N, 2D] buffer:
//
//   rank r:   copy rows -> slab[e & 1]        (its own stream order)
//             dust_peer_signal: flags_on_rank_q[r] = e   for every q     (system-scope stores after a system fence)
//             dust_peer_gather: CTA c waits until flags_here[owner(c)] >= e, then copies that owner's rows
//                               slab_owner[e & 1] -> gathered   (16-byte loads over NVLink, coalesced)
//
// Two slabs alternate by epoch parity: a rank overwrites slab[e & 1] only at epoch e + 2, after its own gather of
// epoch e + 1 has seen every peer's flag e + 1 -- which a peer raises only after its gather of epoch e (the last
// reader of the old slab) is behind it in stream order.  No second barrier, no host synchronisation.
// A lost peer must not hang the GPU: the wait is capped (~2 s) and then traps, surfacing as a launch failure.
//
// PUSH form (dust_peer_push / dust_peer_wait; the one ShardedSVGD uses): the gathered buffers themselves are the
// peer-mapped allocations.  Every rank writes its rows straight from X and score into the gathered buffer of EVERY
// rank (posted 16-byte stores over NVLink; measured on 8 B200: pulling 18 MB with loads ran at 390 GB/s), fences,
// and the last CTA per destination raises this rank's flag there; a one-CTA wait kernel then holds the stream until
// all flags of the epoch are up.  Two gathered buffers alternate by parity for the same reason as the slabs.
#include <string.h>

#include "common.cuh"

namespace dust {

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerThreads = 256;

struct PeerPtrs {
  const float* slab[kPeerMaxWorld];   // this epoch's slab of every rank (own entry: the local one)
  int* flags[kPeerMaxWorld];          // the flag array [world] living on every rank
};

__global__ void peer_signal_kernel(PeerPtrs p, int world, int rank, int epoch) {
  const int q = threadIdx.x;
  if (q >= world) return;
  __threadfence_system();             // the slab written by the kernels before this one, visible system-wide first
  volatile int* f = p.flags[q] + rank;
  *f = epoch;
}

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// grid: (chunks per owner, world).  Every CTA copies a contiguous share of ONE owner's rows, after that owner's flag.
__global__ void __launch_bounds__(kPeerThreads) peer_gather_kernel(PeerPtrs p, int rank, int epoch, long long vec_per_rank,
                                                                   float4* __restrict__ dst) {
  const int owner = blockIdx.y;
  if (threadIdx.x == 0 && owner != rank) {
    const int* f = p.flags[rank] + owner;
    long long spin = 0;
    while (ld_acquire_sys(f) < epoch) {
      __nanosleep(64);
      if (++spin > (1ll << 24)) __trap();   // ~2 s: a peer that never arrives is an error, not a hang
    }
  }
  __syncthreads();
  const float4* __restrict__ src = reinterpret_cast<const float4*>(p.slab[owner]);
  float4* __restrict__ out = dst + (long long)owner * vec_per_rank;
  const long long per_cta = (vec_per_rank + gridDim.x - 1) / gridDim.x;
  const long long v0 = (long long)blockIdx.x * per_cta, v1 = min(v0 + per_cta, vec_per_rank);
  // four independent 16-byte loads in flight per thread: NVLink round trips are ~2 us
  long long v = v0 + threadIdx.x;
  for (; v + 3 * kPeerThreads < v1; v += 4 * kPeerThreads) {
    const float4 a = src[v], b = src[v + kPeerThreads], c = src[v + 2 * kPeerThreads], d = src[v + 3 * kPeerThreads];
    out[v] = a; out[v + kPeerThreads] = b; out[v + 2 * kPeerThreads] = c; out[v + 3 * kPeerThreads] = d;
  }
  for (; v < v1; v += kPeerThreads) out[v] = src[v];
}

struct PushParams {
  float4* dst[kPeerMaxWorld];         // this epoch's gathered buffer of every rank
  int* flags[kPeerMaxWorld];
  const float4* part[4];              // local row pieces (e.g. X and score), each [rows, part_vec] contiguous
  int part_vec[4];
  int n_parts, world, rank, epoch, rows, vec_per_row;
  int* counters;                      // local [world], zero between launches
};

// grid: (chunks, world).  blockIdx.y = destination rank (the own one included: the local copy).
__global__ void __launch_bounds__(kPeerThreads) peer_push_kernel(PushParams p) {
  const int q = blockIdx.y;
  float4* __restrict__ out = p.dst[q] + (long long)p.rank * p.rows * p.vec_per_row;
  const int total = p.rows * p.vec_per_row;            // < 2^31 (checked on the host)
  const int per_cta = (total + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per_cta, i1 = min(i0 + per_cta, total);
  auto fetch = [&](int i) {
    const int row = i / p.vec_per_row;
    int v = i - row * p.vec_per_row;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < p.n_parts) {
        if (v >= 0 && v < p.part_vec[k]) val = __ldg(p.part[k] + (long long)row * p.part_vec[k] + v);
        v -= p.part_vec[k];
      }
    }
    return val;
  };
  // four independent loads in flight per thread, then four posted stores
  int i = i0 + threadIdx.x;
  for (; i + 3 * kPeerThreads < i1; i += 4 * kPeerThreads) {
    const float4 a = fetch(i), b = fetch(i + kPeerThreads), c = fetch(i + 2 * kPeerThreads), d = fetch(i + 3 * kPeerThreads);
    out[i] = a; out[i + kPeerThreads] = b; out[i + 2 * kPeerThreads] = c; out[i + 3 * kPeerThreads] = d;
  }
  for (; i < i1; i += kPeerThreads) out[i] = fetch(i);
  // CTA barrier, then ONE system-scope fence by the thread that counts the CTA in: the barrier orders every thread's
  // stores before it, the fence is cumulative (a fence per thread -- 150 000 of them -- held the kernel at 0.064 ms)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const int done = atomicAdd(&p.counters[q], 1);
    if (done == (int)gridDim.x - 1) {  // the last CTA for this destination: every block of rows has landed
      p.counters[q] = 0;
      __threadfence_system();
      volatile int* f = p.flags[q] + p.rank;
      *f = p.epoch;
    }
  }
}

__global__ void peer_wait_kernel(const int* flags, int world, int epoch) {
  const int q = threadIdx.x;
  if (q >= world) return;
  long long spin = 0;
  while (ld_acquire_sys(flags + q) < epoch) {
    __nanosleep(64);
    if (++spin > (1ll << 24)) __trap();
  }
}

static int fill_ptrs(const dust_peer_args* a, PeerPtrs* out, const char* who) {
  DUST_REQUIRE(a != nullptr, DUST_ERR_INVALID_ARG, "%s: args is NULL", who);
  DUST_REQUIRE(a->world >= 1 && a->world <= kPeerMaxWorld && a->rank >= 0 && a->rank < a->world, DUST_ERR_INVALID_ARG,
               "%s: world %d (<= %d) / rank %d", who, a->world, kPeerMaxWorld, a->rank);
  DUST_REQUIRE(a->slabs && a->flags && a->epoch > 0, DUST_ERR_INVALID_ARG, "%s: slabs, flags and a positive epoch are required", who);
  for (int r = 0; r < a->world; ++r) {
    DUST_REQUIRE(a->slabs[r] && a->flags[r], DUST_ERR_INVALID_ARG, "%s: rank %d has no mapping", who, r);
    out->slab[r] = a->slabs[r];
    out->flags[r] = a->flags[r];
  }
  return DUST_OK;
}

}  // namespace dust

using namespace dust;

extern "C" int dust_peer_alloc(size_t bytes, void** ptr) {
  DUST_REQUIRE(ptr != nullptr && bytes > 0, DUST_ERR_INVALID_ARG, "dust_peer_alloc: ptr / bytes");
  DUST_CUDA_OK(cudaMalloc(ptr, bytes));
  DUST_CUDA_OK(cudaMemset(*ptr, 0, bytes));
  DUST_CUDA_OK(cudaDeviceSynchronize());
  return DUST_OK;
}

extern "C" int dust_peer_free(void* ptr) {
  if (ptr) DUST_CUDA_OK(cudaFree(ptr));
  return DUST_OK;
}

extern "C" int dust_peer_export(const void* ptr, unsigned char handle[64]) {
  DUST_REQUIRE(ptr && handle, DUST_ERR_INVALID_ARG, "dust_peer_export: ptr / handle");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  DUST_CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle, &h, 64);
  return DUST_OK;
}

extern "C" int dust_peer_open(const unsigned char handle[64], void** ptr) {
  DUST_REQUIRE(ptr && handle, DUST_ERR_INVALID_ARG, "dust_peer_open: ptr / handle");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  DUST_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return DUST_OK;
}

extern "C" int dust_peer_close(void* ptr) {
  if (ptr) DUST_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return DUST_OK;
}

extern "C" int dust_peer_signal(const dust_peer_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PeerPtrs p{};
  int rc = fill_ptrs(a, &p, "dust_peer_signal");
  if (rc != DUST_OK) return rc;
  {
    DUST_TIMED("peer_signal_kernel", stream);
    peer_signal_kernel<<<1, 32, 0, stream>>>(p, a->world, a->rank, a->epoch);
  }
  DUST_LAUNCH_OK("peer_signal_kernel");
  return DUST_OK;
}

extern "C" int dust_peer_gather(const dust_peer_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PeerPtrs p{};
  int rc = fill_ptrs(a, &p, "dust_peer_gather");
  if (rc != DUST_OK) return rc;
  DUST_REQUIRE(a->gathered && a->rows_per_rank > 0 && a->row_floats > 0, DUST_ERR_INVALID_ARG,
               "dust_peer_gather: gathered, rows_per_rank, row_floats are required");
  const long long floats = (long long)a->rows_per_rank * a->row_floats;
  DUST_REQUIRE(floats % 4 == 0 && ((uintptr_t)a->gathered & 15) == 0, DUST_ERR_INVALID_ARG,
               "dust_peer_gather: a rank's block must be a whole number of 16-byte vectors (%lld floats)", floats);
  for (int r = 0; r < a->world; ++r)
    DUST_REQUIRE(((uintptr_t)a->slabs[r] & 15) == 0, DUST_ERR_INVALID_ARG, "dust_peer_gather: slab of rank %d is not 16-byte aligned", r);
  const long long vec = floats / 4;
  // enough CTAs to keep every SM busy with loads in flight, at least ~4 KB per CTA
  int chunks = (int)((vec + 255) / 256);
  const int want = (4 * kNumSMs + a->world - 1) / a->world;
  if (chunks > want) chunks = want;
  if (chunks < 1) chunks = 1;
  {
    DUST_TIMED("peer_gather_kernel", stream);
    peer_gather_kernel<<<dim3(chunks, a->world), kPeerThreads, 0, stream>>>(p, a->rank, a->epoch, vec, reinterpret_cast<float4*>(a->gathered));
  }
  DUST_LAUNCH_OK("peer_gather_kernel");
  return DUST_OK;
}

extern "C" int dust_peer_push(const dust_peer_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DUST_REQUIRE(a != nullptr, DUST_ERR_INVALID_ARG, "dust_peer_push: args is NULL");
  DUST_REQUIRE(a->world >= 1 && a->world <= kPeerMaxWorld && a->rank >= 0 && a->rank < a->world && a->epoch > 0, DUST_ERR_INVALID_ARG,
               "dust_peer_push: world %d (<= %d) / rank %d / epoch %d", a->world, kPeerMaxWorld, a->rank, a->epoch);
  DUST_REQUIRE(a->gathered_peers && a->flags && a->counters && a->rows_per_rank > 0, DUST_ERR_INVALID_ARG,
               "dust_peer_push: gathered_peers, flags, counters, rows_per_rank are required");
  DUST_REQUIRE(a->n_parts >= 1 && a->n_parts <= 4, DUST_ERR_INVALID_ARG, "dust_peer_push: 1..4 row pieces, got %d", a->n_parts);
  PushParams p{};
  int width = 0;
  for (int k = 0; k < a->n_parts; ++k) {
    DUST_REQUIRE(a->parts[k] && a->part_floats[k] > 0 && a->part_floats[k] % 4 == 0 && ((uintptr_t)a->parts[k] & 15) == 0, DUST_ERR_INVALID_ARG,
                 "dust_peer_push: piece %d must be 16-byte aligned with a multiple of 4 floats per row", k);
    p.part[k] = reinterpret_cast<const float4*>(a->parts[k]);
    p.part_vec[k] = a->part_floats[k] / 4;
    width += a->part_floats[k];
  }
  DUST_REQUIRE(width == a->row_floats, DUST_ERR_INVALID_ARG, "dust_peer_push: pieces give %d floats per row, row_floats = %d", width, a->row_floats);
  for (int r = 0; r < a->world; ++r) {
    DUST_REQUIRE(a->gathered_peers[r] && a->flags[r] && ((uintptr_t)a->gathered_peers[r] & 15) == 0, DUST_ERR_INVALID_ARG,
                 "dust_peer_push: rank %d has no (aligned) mapping", r);
    p.dst[r] = reinterpret_cast<float4*>(a->gathered_peers[r]);
    p.flags[r] = a->flags[r];
  }
  p.n_parts = a->n_parts; p.world = a->world; p.rank = a->rank; p.epoch = a->epoch; p.rows = a->rows_per_rank;
  p.vec_per_row = a->row_floats / 4; p.counters = a->counters;
  const long long total = (long long)p.rows * p.vec_per_row;
  DUST_REQUIRE(total < (1ll << 31), DUST_ERR_UNSUPPORTED, "dust_peer_push: a rank's block has %lld 16-byte vectors (limit 2^31)", total);
  int chunks = (int)((total + 4 * kPeerThreads - 1) / (4 * kPeerThreads));
  const int want = (4 * kNumSMs + a->world - 1) / a->world;
  if (chunks > want) chunks = want;
  if (chunks < 1) chunks = 1;
  {
    DUST_TIMED("peer_push_kernel", stream);
    peer_push_kernel<<<dim3(chunks, a->world), kPeerThreads, 0, stream>>>(p);
  }
  DUST_LAUNCH_OK("peer_push_kernel");
  return DUST_OK;
}

extern "C" int dust_peer_wait(const dust_peer_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DUST_REQUIRE(a != nullptr && a->flags && a->world >= 1 && a->world <= kPeerMaxWorld && a->rank >= 0 && a->rank < a->world && a->epoch > 0,
               DUST_ERR_INVALID_ARG, "dust_peer_wait: flags, world (<= %d), rank and a positive epoch are required", kPeerMaxWorld);
  DUST_REQUIRE(a->flags[a->rank], DUST_ERR_INVALID_ARG, "dust_peer_wait: the local flag array is missing");
  {
    DUST_TIMED("peer_wait_kernel", stream);
    peer_wait_kernel<<<1, 32, 0, stream>>>(a->flags[a->rank], a->world, a->epoch);
  }
  DUST_LAUNCH_OK("peer_wait_kernel");
  return DUST_OK;
}
