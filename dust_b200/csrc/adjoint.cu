// K2: pathwise likelihood gradient by hand-derived reverse-time adjoints of the two shipped
// models (in place of torch.autograd through rsample -> rollout -> cost,
// dust/inference/svmpc.py:58-60).  Forward sweep stores the trajectory in per-thread local
// memory, the reverse sweep propagates the cost cotangent; clamp sub-gradients are inclusive
// and floor / occupancy terms carry none (torch semantics).
// Compiled with -fmad=false so the forward sweep reproduces K1's trajectories exactly.
#include "models.cuh"

namespace dust {

constexpr int kAdjTile = 128;

__host__ __device__ inline int adj_stride(int HA) {
  int s4 = (HA + 3) / 4;
  if ((s4 & 1) == 0) s4 += 1;
  return s4 * 4;
}

struct AdjKParams {
  ModelParams m;
  int B, N, S, P, H, A, SN, HA, PC, Pchunk, interleaved, likelihood, tiles;
  int p0, p1;      // draws [p0, p1) of the P resident ones are covered by this call (a rank's share)
  const float *state0, *theta, *noise, *sigma, *params, *lik_w;
  float alpha;
  float* partial;  // [B, tiles, PC, N*HA]
};

template <int MODEL, int MAXH>
__global__ void __launch_bounds__(kAdjTile) rollout_adjoint_kernel(const AdjKParams k) {
  constexpr int A = (MODEL == DUST_MODEL_PENDULUM) ? 1 : 2;
  constexpr int DS = (MODEL == DUST_MODEL_PENDULUM) ? 2 : 4;
  constexpr int DP = (MODEL == DUST_MODEL_PENDULUM) ? 2 : 1;
  extern __shared__ __align__(16) float smem[];
  const int stride = adj_stride(k.HA);
  float* tile = smem;                        // actions  [kAdjTile][stride]
  float* gacc = smem + kAdjTile * stride;    // gradient [kAdjTile][stride]
  uint32_t* grid_s = reinterpret_cast<uint32_t*>(gacc + kAdjTile * stride);

  const long long inst = blockIdx.x / k.tiles;
  const int tile_idx = blockIdx.x - (int)inst * k.tiles;
  const int j0 = tile_idx * kAdjTile;
  const int rows = min(kAdjTile, k.SN - j0);
  const int pc = blockIdx.y;
  const int HA = k.HA;

  // The cotangent of a trajectory is its soft-min weight: rows whose weight is EXACTLY zero in float32
  // (cost more than ~100/alpha above the policy's best: with the demos' cost scales that is nearly all
  // of them) contribute exactly nothing and are not rolled out; a tile without any live row returns
  // before staging anything.  ExpectedCost weighs every row alike and skips none.
  const int row = threadIdx.x;
  float coef = 0.f;
  if (row < rows) {
    if (k.likelihood == DUST_LIK_EXP_UTILITY)
      coef = -k.alpha * __ldg(k.lik_w + inst * k.SN + j0 + row) / (float)k.P;
    else
      coef = -k.alpha / ((float)k.S * (float)k.P);
  }
  const bool live = (row < rows) && (coef != 0.f);   // NaN weights stay live and propagate
  if (!__syncthreads_or(live ? 1 : 0)) {
    float* out0 = k.partial + ((inst * k.tiles + tile_idx) * (long long)k.PC + pc) * (long long)(k.N * HA);
    for (int col = threadIdx.x; col < k.N * HA; col += kAdjTile) out0[col] = 0.f;
    return;
  }
  // Deterministic compaction of the live rows (ballot + prefix over the 4 warps).  When at most half
  // of the tile is live, the idle threads take slices of the live rows' parameter loops: thread t
  // serves live row t / G with the draws p_begin + Q*(t % G), stepping Q*G, into its OWN gradient
  // slot, and the slots are summed per policy in slot order afterwards.
  __shared__ int s_live[kAdjTile], s_slot_row[kAdjTile], s_wcnt[kAdjTile / 32];   // s_slot_row: policy of the slot, -1 = unused
  const int warp_ = threadIdx.x >> 5, lane_ = threadIdx.x & 31;
  const unsigned live_mask = __ballot_sync(0xffffffffu, live);
  if (lane_ == 0) s_wcnt[warp_] = __popc(live_mask);
  __syncthreads();
  int n_live = 0, live_base = 0;
#pragma unroll
  for (int w = 0; w < kAdjTile / 32; ++w) {
    if (w < warp_) live_base += s_wcnt[w];
    n_live += s_wcnt[w];
  }
  if (live) s_live[live_base + __popc(live_mask & ((1u << lane_) - 1u))] = row;
  const bool spread = 2 * n_live <= kAdjTile;

  {  // stage actions = theta + sigma*eps
    const float* __restrict__ src = k.noise + (inst * k.SN + j0) * (long long)HA;
    const float* __restrict__ th = k.theta ? k.theta + inst * (long long)k.N * HA : nullptr;
    const int NHA = k.N * HA;
    const int wrap0 = (int)(((long long)j0 * HA) % NHA);
    for (int e = threadIdx.x; e < rows * HA; e += kAdjTile) {
      const int row = e / HA, c = e - row * HA;
      float v = __ldg(src + e);
      if (th) v = __ldg(th + (wrap0 + e) % NHA) + k.sigma[c % A] * v;
      tile[row * stride + c] = v;
    }
    for (int e = threadIdx.x; e < kAdjTile * stride; e += kAdjTile) gacc[e] = 0.f;
    if (MODEL == DUST_MODEL_PARTICLE && k.m.grid_bits != nullptr) {
      const int words = (k.m.grid_nx * k.m.grid_ny + 31) >> 5;
      for (int w = threadIdx.x; w < words; w += kAdjTile) grid_s[w] = __ldg(k.m.grid_bits + w);
    }
  }
  __syncthreads();

  constexpr int Q = (MODEL == DUST_MODEL_PENDULUM) ? 1 : 2;   // parameter draws advanced together per thread
  int my_row = live ? row : -1, p_first = 0, p_step = Q;
  const int chunk_len = min(k.p1, k.p0 + pc * k.Pchunk + k.Pchunk) - (k.p0 + pc * k.Pchunk);   // draws of this CTA
  const int G = spread ? kAdjTile / n_live : 1;                               // thread slots per live row
  const int g_used = min(G, (chunk_len + Q - 1) / Q);                         // ... of which this many get draws
  if (spread) {
    const int l = threadIdx.x / G, g = threadIdx.x % G;
    my_row = (l < n_live && g < g_used) ? s_live[l] : -1;
    p_first = Q * g;
    p_step = Q * G;
    if (my_row >= 0 && k.likelihood == DUST_LIK_EXP_UTILITY)
      coef = -k.alpha * __ldg(k.lik_w + inst * k.SN + j0 + my_row) / (float)k.P;
  }
  s_slot_row[threadIdx.x] = my_row >= 0 ? (j0 + my_row) % k.N : -1;   // the slot's policy
  if (my_row >= 0) {
    const int j = j0 + my_row;
    const float* __restrict__ arow = tile + my_row * stride;
    float* __restrict__ grow = gacc + threadIdx.x * stride;
    const float* __restrict__ x0 = k.state0 + inst * DS;
    const int p_begin = k.p0 + pc * k.Pchunk + p_first, p_end = min(k.p1, k.p0 + pc * k.Pchunk + k.Pchunk);
    for (int p = p_begin; p < p_end; p += p_step) {
      const float* prm = nullptr;
      if (k.params) {
        const int pi = k.interleaved ? (int)(((long long)p * k.SN + j) % k.P) : p;
        prm = k.params + (inst * k.P + pi) * DP;
      }
      if (MODEL == DUST_MODEL_PENDULUM) {
        const PendulumCoef cf = prm ? pendulum_coef_sampled(k.m, __ldg(prm), __ldg(prm + 1)) : pendulum_coef_default(k.m);
        float ths[MAXH + 1], oms[MAXH + 1];
        uint32_t m8[(MAXH + 31) / 32];
#pragma unroll
        for (int w = 0; w < (MAXH + 31) / 32; ++w) m8[w] = 0u;
        float th = __ldg(x0), om = __ldg(x0 + 1);
        ths[0] = th; oms[0] = om;
        for (int t = 0; t < k.H; ++t) {
          float pre;
          pendulum_step(k.m, cf, th, om, arow[t], &pre);
          if (pre >= -k.m.max_speed_pend && pre <= k.m.max_speed_pend) m8[t >> 5] |= 1u << (t & 31);
          ths[t + 1] = th; oms[t + 1] = om;
        }
        float sH, cH;
        sincosf(ths[k.H], &sH, &cH);
        float lam_th = -2.0f * k.m.w_angle * (cH - 1.0f) * sH;
        float lam_om = 2.0f * k.m.w_speed * oms[k.H];
        for (int t = k.H - 1; t >= 0; --t) {
          const float a = arow[t];
          const float tht = ths[t], omt = oms[t];
          const bool in8 = (m8[t >> 5] >> (t & 31)) & 1u;
          const bool in2 = (a >= -k.m.max_torque) && (a <= k.m.max_torque);
          const float gom = lam_om + k.m.dt * lam_th;
          const float g = in8 ? gom : 0.f;
          const float ga = in2 ? g * k.m.dt * cf.c2 : 0.f;
          grow[t] += coef * ga;
          float st, ct;
          sincosf(tht, &st, &ct);
          const float cpi = cosf(tht + kPiF);
          lam_th = lam_th + g * k.m.dt * cf.c1 * cpi - 2.0f * k.m.w_angle * (ct - 1.0f) * st;
          lam_om = g + 2.0f * k.m.w_speed * omt;
        }
      } else {
        // Checkpointed reverse sweep: the forward pass keeps the state every SEG steps (a handful of
        // 16-byte entries); each segment is then rolled out again into REGISTERS and reversed.  Three
        // model-step sweeps instead of two, but a tenth of the local-memory traffic of storing the
        // whole trajectory (which went through L2: the shared-memory tiles leave almost no L1).
        // Q = 2 parameter draws run side by side in one thread: the tiles cost 832 B of shared memory
        // per thread, so an SM holds only 8 warps and the second, independent chain fills the gaps.
        constexpr int SEG = 10;
        constexpr int NSEG = (MAXH + SEG - 1) / SEG;
        const bool has_grid = k.m.grid_bits != nullptr;
        const int nseg = (k.H + SEG - 1) / SEG;
        float mass[Q], rmass[Q], cq[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const bool live = p + q < p_end;
          float mq = k.m.default_mass;
          if (k.params) {
            const int pq = live ? p + q : p;
            const int pi = k.interleaved ? (int)(((long long)pq * k.SN + j) % k.P) : pq;
            mq = __ldg(k.params + (inst * k.P + pi) * DP);
          }
          mass[q] = mq;
          rmass[q] = 1.0f / mq;
          cq[q] = live ? coef : 0.f;   // a padding chain contributes nothing
        }
        float4 ck[NSEG][Q];
        uint32_t occ[NSEG][Q];   // occupancy of the cell the state sits in, step by step (first sweep -> recompute)
        ParticleState s[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) s[q] = ParticleState{__ldg(x0), __ldg(x0 + 1), __ldg(x0 + 2), __ldg(x0 + 3)};
        for (int sg = 0; sg < nseg; ++sg) {
#pragma unroll
          for (int q = 0; q < Q; ++q) ck[sg][q] = make_float4(s[q].x, s[q].y, s[q].vx, s[q].vy);
          uint32_t ob[Q];
#pragma unroll
          for (int q = 0; q < Q; ++q) ob[q] = 0u;
#pragma unroll
          for (int i = 0; i < SEG; ++i) {
            const int t = sg * SEG + i;
            if (t < k.H) {
              const float ax = arow[2 * t], ay = arow[2 * t + 1];
#pragma unroll
              for (int q = 0; q < Q; ++q) {
                const float c = has_grid ? grid_lookup(k.m, grid_s, s[q].x, s[q].y) : 0.f;
                if (c != 0.f) ob[q] |= 1u << i;
                particle_step(k.m, s[q], ax, ay, mass[q], c);
              }
            }
          }
#pragma unroll
          for (int q = 0; q < Q; ++q) occ[sg][q] = ob[q];
        }
        float lam[Q][4];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          lam[q][0] = 2.0f * k.m.w_term[0] * (s[q].x - k.m.target[0]);
          lam[q][1] = 2.0f * k.m.w_term[1] * (s[q].y - k.m.target[1]);
          lam[q][2] = 2.0f * k.m.w_term[2] * (s[q].vx - k.m.target[2]);
          lam[q][3] = 2.0f * k.m.w_term[3] * (s[q].vy - k.m.target[3]);
        }
        for (int sg = nseg - 1; sg >= 0; --sg) {
          float xs[Q][SEG][4];
          uint32_t cbit[Q], mvx[Q], mvy[Q], max_[Q], may_[Q];
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const float4 c4 = ck[sg][q];
            s[q] = ParticleState{c4.x, c4.y, c4.z, c4.w};
            cbit[q] = occ[sg][q];
            mvx[q] = mvy[q] = max_[q] = may_[q] = 0u;
          }
#pragma unroll
          for (int i = 0; i < SEG; ++i) {
            const int t = sg * SEG + i;
            if (t < k.H) {
              const float ax = arow[2 * t], ay = arow[2 * t + 1];
#pragma unroll
              for (int q = 0; q < Q; ++q) {
                xs[q][i][0] = s[q].x; xs[q][i][1] = s[q].y; xs[q][i][2] = s[q].vx; xs[q][i][3] = s[q].vy;
                const float c = (float)((cbit[q] >> i) & 1u);   // same state as in the first sweep: same cell
                float vpre[2], apre[2];
                particle_step(k.m, s[q], ax, ay, mass[q], c, vpre, apre);
                if (vpre[0] >= -k.m.max_speed && vpre[0] <= k.m.max_speed) mvx[q] |= 1u << i;
                if (vpre[1] >= -k.m.max_speed && vpre[1] <= k.m.max_speed) mvy[q] |= 1u << i;
                if (apre[0] >= -k.m.max_accel && apre[0] <= k.m.max_accel) max_[q] |= 1u << i;
                if (apre[1] >= -k.m.max_accel && apre[1] <= k.m.max_accel) may_[q] |= 1u << i;
              }
            }
          }
#pragma unroll
          for (int i = SEG - 1; i >= 0; --i) {
            const int t = sg * SEG + i;
            if (t < k.H) {
              const float ax = arow[2 * t], ay = arow[2 * t + 1];
              float gx = 0.f, gy = 0.f;   // this row's gradient contributions, draw after draw (the order of the scalar loop)
#pragma unroll
              for (int q = 0; q < Q; ++q) {
                const bool crashed = (cbit[q] >> i) & 1u;
                const float kk = (k.m.can_crash && crashed) ? 0.f : k.m.dt;
                const float gvx = ((mvx[q] >> i) & 1u) ? lam[q][2] : 0.f;
                const float gvy = ((mvy[q] >> i) & 1u) ? lam[q][3] : 0.f;
                const float gax = (((max_[q] >> i) & 1u) ? kk * gvx * rmass[q] : 0.f) + 2.0f * k.m.w_ctrl[0] * ax;
                const float gay = (((may_[q] >> i) & 1u) ? kk * gvy * rmass[q] : 0.f) + 2.0f * k.m.w_ctrl[1] * ay;
                gx += cq[q] * gax;
                gy += cq[q] * gay;
                const float l0 = lam[q][0] + 2.0f * k.m.w_state[0] * (xs[q][i][0] - k.m.target[0]);
                const float l1 = lam[q][1] + 2.0f * k.m.w_state[1] * (xs[q][i][1] - k.m.target[1]);
                const float l2 = gvx + kk * lam[q][0] + 2.0f * k.m.w_state[2] * (xs[q][i][2] - k.m.target[2]);
                const float l3 = gvy + kk * lam[q][1] + 2.0f * k.m.w_state[3] * (xs[q][i][3] - k.m.target[3]);
                lam[q][0] = l0; lam[q][1] = l1; lam[q][2] = l2; lam[q][3] = l3;
              }
              grow[2 * t] += gx;
              grow[2 * t + 1] += gy;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // column sums over the rows of this tile that belong to the same policy n
  float* out = k.partial + ((inst * k.tiles + tile_idx) * (long long)k.PC + pc) * (long long)(k.N * HA);
  for (int col = threadIdx.x; col < k.N * HA; col += kAdjTile) {
    const int n = col / HA, c = col - n * HA;
    int r0 = (n - (j0 % k.N)) % k.N;
    if (r0 < 0) r0 += k.N;
    float acc = 0.f;
    if (!spread) {
      for (int r = r0; r < rows; r += k.N) acc += gacc[r * stride + c];
    } else {
      for (int l = 0; l < n_live; ++l) {            // live rows in row order, their slots in draw order
        if (s_slot_row[l * G] != n) continue;
        for (int g = 0; g < g_used; ++g) acc += gacc[(l * G + g) * stride + c];
      }
    }
    out[col] = acc;
  }
}

// out[inst][col] = sum over the `parts` partial rows, in a fixed order: a CTA owns 32 columns, its 8 warps
// take the rows r = w, w + 8, ... (128-byte coalesced reads), and the 8 sub-sums are added in warp order
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, int W,
                                                              int parts) {
  __shared__ float sub[8][33];
  const long long inst = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (col < W) {
    const float* p = partial + inst * (long long)parts * W + col;
    for (int q = warp; q < parts; q += 8) acc += p[(long long)q * W];
  }
  sub[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && col < W) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sub[w][lane];
    out[inst * W + col] = t;
  }
}

struct AdjPlan {
  int PC, Pchunk, tiles;
  size_t total;
};

static AdjPlan plan_adjoint(const dust_adjoint_args* a) {
  AdjPlan pl{};
  const long long SN = (long long)a->S * a->N;
  const int P = a->p_end > a->p_begin ? a->p_end - a->p_begin : (a->params ? a->P : 1);   // draws this call covers
  const long long target_threads = (long long)kNumSMs * 2048 * 2;
  long long pc = 1;
  if (P > 1 && a->B * SN < target_threads) pc = (target_threads + a->B * SN - 1) / (a->B * SN);
  if (pc > P) pc = P;
  const int chunk = (int)((P + pc - 1) / pc);
  pl.PC = (P + chunk - 1) / chunk;
  pl.Pchunk = chunk;
  pl.tiles = ceil_div(SN, kAdjTile);
  const int A = model_da(a->model->kind);
  pl.total = sizeof(float) * (size_t)a->B * pl.tiles * pl.PC * a->N * a->H * A;
  return pl;
}

template <int MODEL, int MAXH>
static int launch_adjoint(const AdjKParams& k, dim3 grid, size_t smem, cudaStream_t stream) {
  if (smem > 48 * 1024)
    DUST_CUDA_OK(cudaFuncSetAttribute(rollout_adjoint_kernel<MODEL, MAXH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  { DUST_TIMED("rollout_adjoint_kernel", stream); rollout_adjoint_kernel<MODEL, MAXH><<<grid, kAdjTile, smem, stream>>>(k); }
  DUST_LAUNCH_OK("rollout_adjoint_kernel");
  return DUST_OK;
}

}  // namespace dust

using namespace dust;

extern "C" size_t dust_adjoint_workspace_bytes(const dust_adjoint_args* a) {
  if (!a || !a->model || a->B <= 0 || a->N <= 0 || a->S <= 0 || a->H <= 0) return 0;
  return plan_adjoint(a).total;
}

extern "C" int dust_adjoint_plan(const dust_adjoint_args* a, int32_t plan[5]) {
  DUST_REQUIRE(a != nullptr && plan != nullptr, DUST_ERR_INVALID_ARG, "dust_adjoint_plan: NULL argument");
  int rc = validate_model(a->model);
  if (rc) return rc;
  DUST_REQUIRE(a->B > 0 && a->N > 0 && a->S > 0 && a->H > 0 && a->H <= 128, DUST_ERR_INVALID_ARG, "dust_adjoint_plan: bad sizes");
  const AdjPlan pl = plan_adjoint(a);
  plan[0] = pl.PC; plan[1] = pl.Pchunk; plan[2] = pl.tiles;
  plan[3] = a->model->kind == DUST_MODEL_PARTICLE ? (a->H + 9) / 10 : 1;   // SEG = 10 in rollout_adjoint_kernel
  plan[4] = a->H <= 32 ? 32 : (a->H <= 64 ? 64 : 128);
  return DUST_OK;
}

extern "C" int dust_rollout_adjoint(const dust_adjoint_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DUST_REQUIRE(a != nullptr, DUST_ERR_INVALID_ARG, "dust_rollout_adjoint: args is NULL");
  int rc = validate_model(a->model);
  if (rc) return rc;
  DUST_REQUIRE(a->B > 0 && a->N > 0 && a->S > 0 && a->H > 0, DUST_ERR_INVALID_ARG, "dust_rollout_adjoint: sizes must be positive");
  DUST_REQUIRE(a->state0 && a->noise && a->grad_theta, DUST_ERR_INVALID_ARG,
               "dust_rollout_adjoint: state0, noise and grad_theta are required");
  DUST_REQUIRE(!a->theta || a->sigma, DUST_ERR_INVALID_ARG, "dust_rollout_adjoint: sigma is required with theta");
  DUST_REQUIRE(a->likelihood != DUST_LIK_EXP_UTILITY || a->lik_weights, DUST_ERR_INVALID_ARG,
               "dust_rollout_adjoint: lik_weights (from dust_rollout_cost) are required for the exponentiated utility");
  DUST_REQUIRE(a->grad_params == nullptr, DUST_ERR_UNSUPPORTED,
               "dust_rollout_adjoint: grad_params is not implemented in this build");
  DUST_REQUIRE(a->H <= 128, DUST_ERR_UNSUPPORTED, "dust_rollout_adjoint: H=%d > 128", a->H);
  DUST_REQUIRE(a->B <= 65535, DUST_ERR_UNSUPPORTED, "dust_rollout_adjoint: B > 65535");
  const AdjPlan pl = plan_adjoint(a);
  DUST_REQUIRE(a->workspace && a->workspace_bytes >= pl.total, DUST_ERR_WORKSPACE,
               "dust_rollout_adjoint: workspace needs %zu bytes, got %zu", pl.total, a->workspace_bytes);
  const int kind = a->model->kind, A = model_da(kind);
  AdjKParams k;
  k.m = to_params(*a->model);
  k.B = a->B; k.N = a->N; k.S = a->S; k.P = a->params ? a->P : 1; k.H = a->H; k.A = A;
  k.SN = a->S * a->N; k.HA = a->H * A; k.PC = pl.PC; k.Pchunk = pl.Pchunk;
  k.p0 = a->p_end > a->p_begin ? a->p_begin : 0; k.p1 = a->p_end > a->p_begin ? a->p_end : k.P;
  DUST_REQUIRE(a->p_begin >= 0 && a->p_end >= a->p_begin && a->p_end <= k.P, DUST_ERR_INVALID_ARG,
               "dust_rollout_adjoint: draw range [%d, %d) outside [0, %d)", a->p_begin, a->p_end, k.P);
  k.interleaved = a->param_tiling == DUST_PARAMS_INTERLEAVED; k.likelihood = a->likelihood; k.tiles = pl.tiles;
  k.state0 = a->state0; k.theta = a->theta; k.noise = a->noise; k.sigma = a->sigma; k.params = a->params;
  k.lik_w = a->lik_weights; k.alpha = a->alpha; k.partial = (float*)a->workspace;
  size_t smem = sizeof(float) * 2 * kAdjTile * adj_stride(k.HA);
  if (kind == DUST_MODEL_PARTICLE && a->model->grid_bits) smem += sizeof(uint32_t) * ((a->model->grid_nx * a->model->grid_ny + 31) / 32);
  DUST_REQUIRE(smem <= 227 * 1024, DUST_ERR_UNSUPPORTED, "dust_rollout_adjoint: H*A=%d needs %zu B of shared memory", k.HA, smem);
  const long long gx = (long long)a->B * pl.tiles;
  DUST_REQUIRE(gx < (1ll << 31), DUST_ERR_UNSUPPORTED, "dust_rollout_adjoint: grid too large");
  dim3 grid((unsigned)gx, (unsigned)pl.PC, 1);
  if (kind == DUST_MODEL_PENDULUM) {
    if (a->H <= 32) rc = launch_adjoint<DUST_MODEL_PENDULUM, 32>(k, grid, smem, stream);
    else if (a->H <= 64) rc = launch_adjoint<DUST_MODEL_PENDULUM, 64>(k, grid, smem, stream);
    else rc = launch_adjoint<DUST_MODEL_PENDULUM, 128>(k, grid, smem, stream);
  } else {
    if (a->H <= 32) rc = launch_adjoint<DUST_MODEL_PARTICLE, 32>(k, grid, smem, stream);
    else if (a->H <= 64) rc = launch_adjoint<DUST_MODEL_PARTICLE, 64>(k, grid, smem, stream);
    else rc = launch_adjoint<DUST_MODEL_PARTICLE, 128>(k, grid, smem, stream);
  }
  if (rc) return rc;
  const int W = a->N * k.HA;
  { DUST_TIMED("reduce_partials_kernel", stream); reduce_partials_kernel<<<dim3((unsigned)ceil_div(W, 32), (unsigned)a->B, 1), 256, 0, stream>>>(
      k.partial, a->grad_theta, W, pl.tiles * pl.PC); }
  DUST_LAUNCH_OK("reduce_partials_kernel");
  return DUST_OK;
}
