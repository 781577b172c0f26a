/*
 * dust_b200.h -- C ABI of libdust_b200.so: the DuSt-MPC inner loop as sm_100a CUDA kernels.
 *
 * The reference (lubaroli/dust) is pure Python/torch: it has NO FFI boundary of its own
 * (SURVEY.md §2.1, §8b).  This header is therefore the boundary a binding for the reference's
 * hot path would target; every entry point cites the reference code it replaces
 * (paths relative to the reference repository root).  INTEGRATION.md shows the ctypes
 * binding that `dust_b200/_lib.py` uses and how the reference classes map onto it.
 *
 * Conventions
 *   - all tensors are device pointers to contiguous row-major float32 unless noted;
 *     every entry point has a leading "instance" (batch) dimension B of independent MPC
 *     problems (B = 1 reproduces the reference's single-problem API);
 *   - no entry point allocates, synchronises or touches the host: work is enqueued on
 *     `stream` (a cudaStream_t passed as void*), scratch is supplied by the caller
 *     (sizes from the *_workspace_bytes functions);
 *   - return value: 0 on success, a negative dust_status otherwise; the message for the
 *     last error on the calling thread is available from dust_last_error();
 *   - optional pointers may be NULL where stated.
 *
 * Index names: B instances, N policies (= SVGD particles), S action samples per policy,
 * P dynamics-parameter samples, H horizon, A action dim, ds state dim, dp parameter dim,
 * D = H*A flattened particle dimension.
 */
#ifndef DUST_B200_H
#define DUST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DUST_B200_ABI_VERSION 3

typedef enum dust_status {
  DUST_OK = 0,
  DUST_ERR_INVALID_ARG = -1,   /* NULL / out-of-range / inconsistent sizes             */
  DUST_ERR_UNSUPPORTED = -2,   /* a shape or mode this build has no kernel for         */
  DUST_ERR_WORKSPACE = -3,     /* caller-supplied scratch too small                    */
  DUST_ERR_CUDA = -4           /* a CUDA runtime call failed (message has the detail)  */
} dust_status;

/* dust/models/{pendulum,particle}.py */
typedef enum dust_model_kind { DUST_MODEL_PENDULUM = 0, DUST_MODEL_PARTICLE = 1 } dust_model_kind;

/* how the P sampled parameter vectors are tiled over the rollouts (dust/controllers/disco.py:171-179):
 * BLOCKED     row (p,s,n) uses params[p]                 (params_dist with a vector event)
 * INTERLEAVED row (p,s,n) uses params[(p*S*N+s*N+n) % P] (scalar-event params_dist: `repeat` quirk) */
typedef enum dust_param_tiling { DUST_PARAMS_BLOCKED = 0, DUST_PARAMS_INTERLEAVED = 1 } dust_param_tiling;

/* dust/inference/likelihoods.py:104-135 */
typedef enum dust_likelihood_kind { DUST_LIK_EXP_UTILITY = 0, DUST_LIK_EXPECTED_COST = 1 } dust_likelihood_kind;

/* dust/inference/svmpc.py:142-158 */
typedef enum dust_roll_strategy { DUST_ROLL_REPEAT = 0, DUST_ROLL_MEAN = 1, DUST_ROLL_RESAMPLE = 2 } dust_roll_strategy;

/* dust/controllers/disco.py:396-417 */
typedef enum dust_select_strategy { DUST_SELECT_ARGMAX = 0, DUST_SELECT_AVERAGE = 1 } dust_select_strategy;

/* Static description of the forward model and its cost (host struct, passed by pointer).
 * Pendulum: dust/models/pendulum.py:61-100 + demo/pendulum_example.py:21-28 cost
 *           cost = w_angle*(cos(th)-1)^2 + w_speed*thd^2 (instantaneous and terminal).
 * Particle: dust/models/particle.py:117-225, dust/utils/obstacle_map.py:64-93. */
typedef struct dust_model_desc {
  int32_t kind;            /* dust_model_kind                                           */
  float dt;                /* pendulum 0.05, particle 0.015                            */
  /* pendulum */
  float g;                 /* 9.8                                                      */
  float max_torque;        /* 2.0                                                      */
  float max_speed_pend;    /* 8.0                                                      */
  float w_angle;           /* 50.0                                                     */
  float w_speed;           /* 1.0                                                      */
  float default_length;    /* used when params == NULL                                 */
  float default_mass;      /* used when params == NULL (both models)                   */
  /* particle */
  float max_accel;         /* 10                                                       */
  float max_speed;         /* 5                                                        */
  float target[4];
  float w_state[4];        /* [w_qpos,w_qpos,w_qvel,w_qvel]                            */
  float w_term[4];
  float w_ctrl[2];
  float w_obs;             /* 1e6                                                      */
  float inv_cell;          /* (float)(1.0 / cell_size)                                 */
  float c_offset[2];       /* map origin in cells                                      */
  int32_t grid_nx, grid_ny;/* occupancy grid dims (cells)                              */
  int32_t can_crash;       /* can_crash && with_obstacle: freeze state while in a cell */
  int32_t with_obstacle;   /* obstacle term in the costs                               */
  const uint32_t* grid_bits; /* DEVICE pointer: bit (ix*grid_ny+iy) of a packed, LSB-first
                                uint32 array, ceil(nx*ny/32) words; NULL if !with_obstacle */
} dust_model_desc;

/* ------------------------------------------------------------------------------------------
 * K1  rollout + cost + per-policy soft-min / likelihood reduction
 * replaces: MultiDISCO._rollout / _compute_cost / forward  dust/controllers/disco.py:139-209,
 *           294-346, 380-393; CostLikelihood.sample + log_prob  dust/inference/likelihoods.py:81-135;
 *           analytic likelihood gradient  dust/inference/svmpc.py:46-54.
 * ------------------------------------------------------------------------------------------ */
typedef struct dust_rollout_args {
  const dust_model_desc* model;
  int32_t B, N, S, P, H;
  int32_t param_tiling;      /* dust_param_tiling                                      */
  int32_t likelihood;        /* dust_likelihood_kind                                   */
  /* inputs */
  const float* state0;       /* [B, ds]                                                */
  const float* theta;        /* [B, N, H, A] policy means, or NULL when `noise` already
                                holds the action sequences (ext_actions)               */
  const float* noise;        /* [B, S, N, H, A]: standard-normal eps (actions = theta +
                                sigma*eps, likelihoods.py:85-90) or actions if theta==NULL */
  const float* sigma;        /* [A] sqrt(diag(a_cov)); required iff theta != NULL      */
  const float* params;       /* [B, P, dp] physical parameters (already exp'd when the
                                controller samples in log space) or NULL => defaults, P=1 */
  const float* a_seq;        /* [B, H, A] or NULL: the controller's selected plan; the
                                MPPI perturbation is actions - a_seq (disco.py:164)    */
  const float* pert;         /* [B, S, N, H, A] or NULL: explicit MPPI perturbation
                                (internal sampling, disco.py:157-160); overrides a_seq */
  float alpha;               /* likelihood inverse temperature (likelihoods.py:105,123)*/
  float temperature;         /* controller temperature (disco.py:89)                   */
  const float* sigma_weights;/* [P] or NULL: unscented-transform mode (MultiDISCO._sigma_rollout,
                                disco.py:211-292): `params` holds the P = 2n+1 sigma points and the
                                trajectory cost is the reference's weighted sum (disco.py:312-323,
                                INCLUDING its grouping of the instantaneous costs: the flat
                                (sigma point, step) index m = p*H + t takes weight [m mod P], and
                                each run of P consecutive m is one dot product) instead of the
                                mean over P                                            */
  const float* ctrl_mat;     /* [B, N, H, A] or NULL: a_mat @ a_pre; with ctrl_reg it adds the control
                                regulariser ctrl_reg * sum_{h,a} (a_seq - action) * ctrl_mat[n]
                                (disco.py:334-344) to every trajectory cost           */
  float ctrl_reg;            /* a_reg = temperature * (1 - ctrl_penalty) (disco.py:90)   */
  int32_t p_begin, p_end;    /* 0, 0: all P draws.  Otherwise (multi-GPU sharding of ONE instance over its
                                parameter draws) this call rolls out draws [p_begin, p_end) of the P resident
                                ones and `costs` receives their share of the mean, (1/P) * sum over the range;
                                no other output may be requested.  Sum the shares over the ranks (all-reduce)
                                and finish with dust_cost_reduce                        */
  /* outputs (any may be NULL) */
  float* costs;              /* [B, S, N]      trajectory costs, mean over P           */
  float* log_lik;            /* [B, N]         likelihoods.py:113-135                  */
  float* lik_weights;        /* [B, S, N]      softmax_S(-alpha C)   (svmpc.py:49-51)  */
  float* grad_lik;           /* [B, N, H, A]   svmpc.py:52-54 (needs theta, sigma)     */
  float* mppi_weights;       /* [B, S, N]      exp(omega)            (disco.py:385,394)*/
  float* mppi_delta;         /* [B, N, H, A]   sum_s exp(omega) pert (disco.py:387-392)*/
  float* mix;                /* [B, N]         a_mix                 (disco.py:393)    */
  float* states;             /* [B, P, S, N, H+1, ds] rollouts (disco.py:202-205)      */
  /* scratch */
  void* workspace;
  size_t workspace_bytes;
} dust_rollout_args;

size_t dust_rollout_workspace_bytes(const dust_rollout_args* args);
int dust_rollout_cost(const dust_rollout_args* args, void* stream);
/* How dust_rollout_cost will run these arguments (no launch): plan[0] = 1 if the fused per-instance kernel
 * takes them, plan[1] = parameter chunks over grid.y, plan[2] = draws per chunk, plan[3] = thread groups
 * that share one action tile (NSUB), plan[4] = partial cost rows per instance.  Tests use it to prove that
 * a shape reaches the variant it is meant to exercise. */
int dust_rollout_plan(const dust_rollout_args* args, int32_t plan[5]);
/* The reductions that follow the costs, on their own: `costs` [B,S,N] is an INPUT (complete trajectory
 * costs, e.g. all-reduced shares); log_lik / lik_weights / grad_lik / mppi_weights / mppi_delta / mix are
 * written as dust_rollout_cost would (disco.py:380-393, likelihoods.py:113-135, svmpc.py:46-54).
 * Same workspace as dust_rollout_cost. */
int dust_cost_reduce(const dust_rollout_args* args, void* stream);

/* The whole SVGD step of SVMPC (and optionally SVMPC.forward) for B instances in ONE launch:
 * K1 with its fused per-instance kernel, then -- in the tail of the same CTA -- the GMM prior score,
 * phi (gamma, c1, c2 as for dust_svgd_phi), the SGD update theta_out = theta + lr*phi and, with
 * do_forward, the weights / argmax / shift / mixture refresh of dust_svmpc_forward.
 * replaces: SVMPC.step + SVMPC.forward  dust/inference/svmpc.py:87-95, 172-200 (one call each).
 * `rollout` carries the K1 inputs (theta, noise, sigma, state0, params ...) and the optional K1
 * outputs costs / log_lik / grad_lik; its weights / MPPI / states outputs must be NULL.
 * Returns DUST_ERR_UNSUPPORTED when the shape does not qualify (B < 74, H*A > 32, parameter loop
 * split over chunks): call the staged entry points instead. */
typedef struct dust_svmpc_step_args {
  dust_rollout_args rollout;
  int32_t do_forward, roll_strategy, weighted_prior, prior_aliased;
  const float* mu;           /* [B, N, H, A] prior centres (unused when prior_aliased)  */
  const float* mix;          /* [B, N] or NULL                                          */
  const float* inv_var;      /* [H*A]                                                   */
  float log_norm;
  float gamma, c1, c2, lr;
  float* theta_out;          /* [B, N, H, A] updated particles before the shift, or NULL */
  float* phi;                /* [B, N, H, A] or NULL                                    */
  float* p_weights;          /* [B, N]        (do_forward)                              */
  int32_t* i_star;           /* [B]                                                     */
  float* a_seq;              /* [B, H, A]                                               */
  float* theta_next;         /* [B, N, H, A]                                            */
  float* mix_next;           /* [B, N]                                                  */
} dust_svmpc_step_args;

int dust_svmpc_step(const dust_svmpc_step_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2  pathwise likelihood gradient by a hand-derived reverse-time adjoint
 * replaces: torch.autograd.grad(log_l.sum(), x) through rsample -> rollout -> cost,
 *           dust/inference/svmpc.py:58-60 (the commented alternative), for the two shipped models.
 * grad_theta[b,n,h,a] = sum_s -alpha*w[s,n] * mean_p dC_p[s,n]/da[h,a]   (exp-utility)
 *                     = -alpha/S * sum_s mean_p dC_p/da                 (expected cost)
 * `lik_weights` is K1's output (ignored for expected cost).  grad_params (optional) is
 * d sum_n log_l_n / d params[b,p,:] w.r.t. the physical parameters (blocked tiling only).
 * ------------------------------------------------------------------------------------------ */
typedef struct dust_adjoint_args {
  const dust_model_desc* model;
  int32_t B, N, S, P, H;
  int32_t param_tiling;
  int32_t likelihood;
  const float* state0;       /* [B, ds]            */
  const float* theta;        /* [B, N, H, A] or NULL (noise = actions) */
  const float* noise;        /* [B, S, N, H, A]    */
  const float* sigma;        /* [A]                */
  const float* params;       /* [B, P, dp] or NULL */
  const float* lik_weights;  /* [B, S, N]          */
  float alpha;
  float* grad_theta;         /* [B, N, H, A]       */
  float* grad_params;        /* [B, P, dp] or NULL */
  void* workspace;
  size_t workspace_bytes;
  int32_t p_begin, p_end;    /* 0, 0: all draws; else only draws [p_begin, p_end): grad_theta receives their
                                share of the gradient (the 1/P of the mean included) -- sum over the ranks */
} dust_adjoint_args;

size_t dust_adjoint_workspace_bytes(const dust_adjoint_args* args);
int dust_rollout_adjoint(const dust_adjoint_args* args, void* stream);
/* plan[0] = parameter chunks, plan[1] = draws per chunk, plan[2] = trajectory tiles per instance,
 * plan[3] = checkpointed segments of the reverse sweep (particle model; 1 = stored trajectory),
 * plan[4] = compiled horizon bound (32 / 64 / 128). */
int dust_adjoint_plan(const dust_adjoint_args* args, int32_t plan[5]);

/* ------------------------------------------------------------------------------------------
 * K3  Gaussian-mixture prior: log-density and score
 * replaces: prior.log_prob(x) and its autograd gradient, dust/inference/svmpc.py:41,138;
 *           get_gmm dust/inference/svgd.py:84-89; MPF prior dust/inference/mpf.py:26-38,45.
 * log p(x_i) = logsumexp_k( logmix_k - 0.5*sum_d (x_id-mu_kd)^2*inv_var_d ) + const
 * score_i    = sum_k r_ik (mu_k - x_i) * inv_var
 * `mix` holds UNNORMALISED weights; they are normalised, clamped to [eps,1-eps] and
 * re-normalised exactly as torch.distributions.Categorical does.
 * ------------------------------------------------------------------------------------------ */
typedef struct dust_gmm_args {
  int32_t B, M, K, D;        /* M evaluation points, K components                      */
  const float* x;            /* [B, M, D]                                              */
  const float* mu;           /* [B, K, D]                                              */
  const float* mix;          /* [B, K] or NULL (uniform)                               */
  const float* inv_var;      /* [D]  diagonal precision                                */
  float log_norm;            /* -0.5*(D*log(2*pi) + sum_d log var_d)                   */
  float* log_prob;           /* [B, M] or NULL                                         */
  float* score;              /* [B, M, D] or NULL                                      */
} dust_gmm_args;

int dust_gmm_score(const dust_gmm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * K4  exact median of all N^2 squared pairwise distances (median-heuristic bandwidth)
 * replaces: bw_median dust/inference/svgd.py:42-52 (torch.median = LOWER median, rank
 *           (N^2-1)/2 of the clamped addmm distances) and RBF.compute_bandwidth
 *           dust/kernels/base_kernels.py:58-66.
 * Two-pass radix select over the float32 bit pattern of d2 = max(|x_i|^2+|x_j|^2-2 x_i.x_j, 0):
 * pass 0 histograms bits [31:16], pass 1 bits [15:0] inside the selected bin.  Rows
 * [row_begin,row_end) x all N columns are visited, so R ranks can each histogram their own row
 * block, all-reduce `hist` (uint64[65536]) and call dust_median_select on the sum.
 * ------------------------------------------------------------------------------------------ */
typedef struct dust_median_args {
  int32_t N, D;
  int32_t row_begin, row_end;
  const float* x;            /* [N, D]                                                 */
  unsigned long long* hist;  /* [196616] device (65536 used by the two-pass select, 3*65536+1 by the
                                fast path); kernels ADD into it (caller zeroes it once) */
  uint32_t* selected;        /* [8] device, zeroed by the caller: {hi16, rank lo, rank hi, bits,
                                window start, fast-path ok flag, -, bits (fast path)} */
  float* row_norms;          /* [N] device scratch: |x_i|^2, written by pass 0, read by pass 1 */
  int32_t sample_begin, sample_end;  /* fast path: the share [begin, end) of the 2^20 sampled pairs this rank draws
                                (0, 0: all of them); sum the sample histogram over the ranks before `window` */
  int32_t ld;                /* row stride of x in floats (0 = D: contiguous rows) */
} dust_median_args;

int dust_median_hist_pass(const dust_median_args* args, int32_t pass, void* stream);
/* pass 0: picks the hi-16 bin holding rank k=(N*N-1)/2 and the residual rank; pass 1: picks the
 * lo-16 value, writes the median's bit pattern to selected[3] and the float to *median_out. */
int dust_median_select(const dust_median_args* args, int32_t pass, float* median_out, void* stream);

/* Fast path of K4 on the tensor cores (N % 128 == 0, N >= 1024): a deterministic sample of 2^20 pairs
 * picks a window of 3*65536 float bit patterns around the median (`prepare`, identical on every rank);
 * ONE tcgen05 Gram pass over rows [row_begin,row_end) counts the values below the window in registers
 * and histograms only those inside it (`count`; all-reduce `hist[0..196608]` when rows are sharded);
 * `select` writes the median and sets selected[5] = 1, or leaves it 0 if the rank fell outside the
 * window -- the two-pass entry points above then do the work (their kernels return at once when
 * selected[5] is set, so they can always be enqueued behind the fast path without a host sync). */
int dust_median_fast_supported(int32_t N, int32_t D);
size_t dust_median_fast_workspace_bytes(int32_t N, int32_t D);
int dust_median_fast_prepare(const dust_median_args* args, void* workspace, size_t workspace_bytes, void* stream);
/* `prepare` writes the operand images and histograms this rank's share of the sample into the workspace at byte offset
 * dust_median_fast_sample_hist_offset (uint32[32768]; all-reduce it when the sample is sharded); `window` then
 * interpolates the median's position in the sample and chooses the window (selected[4..6]). */
size_t dust_median_fast_sample_hist_offset(int32_t N, int32_t D);
int dust_median_fast_window(const dust_median_args* args, void* workspace, size_t workspace_bytes, void* stream);
int dust_median_fast_count(const dust_median_args* args, void* workspace, size_t workspace_bytes, void* stream);
int dust_median_fast_select(const dust_median_args* args, float* median_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5  SVGD direction  phi = c1 * K S + c2 * (rowsum(K) o X - K X),  K_ij = exp(-gamma*|x_i-x_j|^2)
 * replaces: SVGD.phi dust/inference/svgd.py:127-135 (+ default_kernel :92-99);
 *           SVMPC.phi kernel branches dust/inference/svmpc.py:62-85;  MPF.phi dust/inference/mpf.py:53-57;
 *           RBF.eval dust/kernels/base_kernels.py:91-108; iid_mp.eval dust/kernels/composite_kernels.py:33-64.
 * (gamma, c1, c2) per variant: SURVEY.md §8(a) table P.  When `gamma_dev` != NULL the three
 * scalars are read from device memory (gamma_dev[3] = {gamma, c1, c2}; lets the bandwidth
 * come from dust_median_select without a host round trip).
 * per_dim != 0 selects the message-passing kernel: one scalar RBF per flattened dimension with
 * its own lower-median bandwidth h_d = max(scale*med_d/log(N+1), 1e-5), c1 = 1/N, c2 = 2/(N h_d).
 * If x_out != NULL the SGD update x_out = x + lr * phi is fused (svmpc.py:93-94).
 * Rows [row_begin,row_end) of phi are produced (all N columns are read).
 * ------------------------------------------------------------------------------------------ */
typedef struct dust_phi_args {
  int32_t B, N, D;
  int32_t row_begin, row_end;
  int32_t per_dim;
  const float* x;            /* [B, N, D]                                              */
  const float* score;        /* [B, N, D]                                              */
  float gamma, c1, c2;
  const float* gamma_dev;    /* optional device {gamma, c1, c2}                        */
  float bw_scale;            /* per_dim only: RBF.ell_scale                            */
  float lr;
  float* phi;                /* [B, N, D] (rows outside the range untouched) or NULL   */
  float* x_out;              /* [B, N, D] or NULL                                      */
  float* bandwidths;         /* per_dim only, optional [B, D]                          */
  void* workspace;
  size_t workspace_bytes;
  int32_t x_prepared;        /* tensor-core path: the head of `workspace` already holds |x|^2 and the hi/lo operand images
                                of this x -- written there by dust_median_fast_prepare on the SAME buffer (its workspace
                                has the same head layout) -- so only the [score | x] images are prepared               */
  int32_t ld;                /* row stride of x AND score in floats (0 = D: contiguous rows).  ld > D lets both live in
                                one [N, ld] buffer -- e.g. the all-gathered [X | score] with score = x + D -- without
                                slicing copies.  Tensor-core path only (other paths require ld = 0 or D)               */
} dust_phi_args;

size_t dust_phi_workspace_bytes(const dust_phi_args* args);
/* How the tensor-core path of dust_svgd_phi cuts `row_tiles` (128 rows each) x `col_tiles` (64 columns each) tile
 * pairs into launches and CTAs (no launch, no device needed).  plan[i] = {grid, R, n_chunks, chunk_w, rem0, rem_w,
 * units_per_cta, total_units, max_seg, row_tile0, row_tiles} of launch i: CTA c < n_chunks * R owns row tile c % R and
 * the column chunk [(c / R) * chunk_w, +chunk_w); the others own equal contiguous ranges of the row-major numbered
 * left-over columns [rem0, rem0 + rem_w).  Returns the number of launches (<= 2).  Tests use it to prove coverage. */
int dust_phi_tc_plan(int32_t row_tiles, int32_t col_tiles, int32_t plan[2][11]);
/* Which form of the tensor-core kernel dimension D takes (no device needed): bit 0 = the row tile of the Gram GEMM
 * lives in TMEM (TS form), bit 1 = the P_lo V correction term of the second GEMM runs as kind::f16 on bf16 copies
 * (2^-11 of the sum carried with 16 mantissa bits); -1 = D does not fit (the SIMT kernel takes it).  bench_phi.py
 * uses it to count the tensor work actually issued. */
int dust_phi_tc_mode(int32_t D);
int dust_svgd_phi(const dust_phi_args* args, void* stream);

/* bandwidth -> (gamma,c1,c2) on the device, from the median written by dust_median_select:
 *   mode 0 (svgd.py:51,98,131-133): bw = scale*max(sqrt(med/2)/log(N+1),1e-5); gamma=1/(2bw^2), c1=1/N, c2=1/(N bw^2)
 *   mode 1 (base_kernels.py:64-100): h = max(scale*med/log(N+1),1e-5);          gamma=1/h,      c1=1/N, c2=2/(N h)
 * out[4] = {gamma, c1, c2, bw_or_h}. */
int dust_bandwidth_from_median(const float* median, int32_t N, float scale, int32_t mode,
                               float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K7  after the SVGD step: particle weights, best particle, time shift, prior refresh
 * replaces: SVMPC.get_weights / forward / roll / update_prior dust/inference/svmpc.py:128-200.
 * log_w_n = log_lik_n + log GMM(theta_n; mu, mix); p = softmax_N(log_w); i* = first argmax;
 * a_seq = theta[i*]; theta <- shift(theta) with the last step repeated / averaged;
 * mu <- theta (new prior centres), mix <- p if weighted_prior else 1.
 * ------------------------------------------------------------------------------------------ */
typedef struct dust_svmpc_forward_args {
  int32_t B, N, H, A;
  int32_t roll_strategy;     /* dust_roll_strategy                                     */
  int32_t weighted_prior;
  const float* log_lik;      /* [B, N] from K1 (pre-update costs, fast_pred)           */
  const float* theta;        /* [B, N, H, A] updated particles                         */
  const float* mu;           /* [B, N, H, A] prior centres (pass theta when the prior aliases it) */
  const float* mix;          /* [B, N] or NULL                                         */
  const float* inv_var;      /* [H*A]                                                  */
  float log_norm;
  float* p_weights;          /* [B, N]                                                 */
  int32_t* i_star;           /* [B]                                                    */
  float* a_seq;              /* [B, H, A]                                              */
  float* theta_next;         /* [B, N, H, A] rolled particles (must not alias theta)   */
  float* mix_next;           /* [B, N]                                                 */
  const float* resample_noise; /* [B, N, A+1] standard normals (dust_noise_normal), required for
                              * DUST_ROLL_RESAMPLE (svmpc.py:148-150: the new last action of particle n is
                              * the last step of a draw from the CURRENT prior): columns 0..A-1 perturb
                              * the chosen centre, column A picks the component (through the normal CDF) */
} dust_svmpc_forward_args;

int dust_svmpc_forward(const dust_svmpc_forward_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * stand-alone controller step
 * replaces: MultiDISCO.step dust/controllers/disco.py:396-417 (argmax | average, clamp to the
 * action box, emit the first `steps` actions, shift a_seq and a_mat with zero fill).  With
 * argmax the clamped row is also written back into a_mat before the shift (the reference's
 * a_seq is a view of a_mat[i*]).
 * ------------------------------------------------------------------------------------------ */
typedef struct dust_disco_step_args {
  int32_t B, N, H, A;
  int32_t strategy;          /* dust_select_strategy                                   */
  int32_t steps;
  const float* a_low;        /* [A] */
  const float* a_high;       /* [A] */
  float* a_mat;              /* [B, N, H, A] in/out                                    */
  const float* a_mix;        /* [B, N]                                                 */
  float* a_seq;              /* [B, H, A] out                                          */
  float* next_actions;       /* [B, steps, A] out                                      */
} dust_disco_step_args;

int dust_disco_step(const dust_disco_step_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * MPF: SVGD over dynamics-parameter particles, all n_steps in one launch
 * replaces: MPF.optimize / step / phi dust/inference/mpf.py:40-86 and
 *           GaussianLikelihood.sample / log_prob dust/inference/likelihoods.py:30-49
 *           (one model step per particle + autograd w.r.t. the parameters, mpf.py:50).
 * Per step: score = J^T (obs1 - f(obs0, action; x))/obs_std^2 + grad log GMM(x; centres = x,
 * prior_inv_var) ; phi = K score / Np - (1/bw^2) (rowsum(K) o x - K x), K = exp(-d2/(2 bw^2));
 * x += lr * phi.  log_space: physical parameter = exp(x).
 * ------------------------------------------------------------------------------------------ */
typedef struct dust_mpf_args {
  const dust_model_desc* model;
  int32_t B, Np, n_steps;
  int32_t log_space;
  float* x;                  /* [B, Np, dp] in/out                                     */
  const float* obs0;         /* [B, ds] previous observation                           */
  const float* action;       /* [B, A]  action applied                                 */
  const float* obs1;         /* [B, ds] new observation                                */
  const float* prior_inv_var;/* [dp]                                                   */
  float obs_std, bw, lr;
  float* grad_norms;         /* [B, n_steps] or NULL                                   */
  void* workspace;           /* optional, dust_mpf_workspace_bytes(): a single instance with many particles
                              * (B = 1, Np >= 128) is then spread over many SMs by a cooperative launch
                              * (two grid barriers per SVGD step, same per-particle arithmetic)            */
  size_t workspace_bytes;
  const float* bw_dev;       /* optional device scalar: the kernel bandwidth is read from it (bw is then ignored), e.g.
                              * the output of dust_silverman_bandwidth -- no host round trip between the two          */
  float* phi_out;            /* optional [B, Np, dp]: phi of the LAST step.  With lr = 0 and n_steps = 1 the call only
                              * evaluates phi at x -- the host then applies any optimiser (mpf.py:59-62: Adam is the
                              * reference's default; plain SGD stays fused in the kernel)                                */
} dust_mpf_args;

size_t dust_mpf_workspace_bytes(const dust_mpf_args* args);   /* 0: the one-CTA-per-instance kernel is used */
int dust_mpf_optimize(const dust_mpf_args* args, void* stream);

/* Silverman's rule as KDEpy 1.1.0 `bw_selection.silvermans_rule` applies it to the flattened particles
 * (dust/inference/mpf.py:72, svmpc.py:105): sigma = min(std(ddof=1), IQR/1.349) [std if IQR = 0],
 * bw = scale * sigma * (0.75 n)^(-1/5); percentiles by linear interpolation, arithmetic in double.
 * x: n <= 4096 device floats.  bw_out[0] = bw; inv_var_out[0..dp) = 1/bw^2 (may be NULL). */
int dust_silverman_bandwidth(const float* x, int32_t n, float scale, float* bw_out, float* inv_var_out, int32_t dp, void* stream);

/* one model step for M (state, action, params) triples: BaseModel.step
 * dust/models/pendulum.py:61-100, dust/models/particle.py:117-166.  params may be NULL. */
int dust_model_step(const dust_model_desc* model, int32_t M, const float* states, const float* actions,
                    const float* params, float* next_states, void* stream);

/* trajectory-free cost evaluation used by the drivers (inst_cost_fn on the plant state):
 * demo/pendulum_example.py:21-28, dust/models/particle.py:170-225.  actions may be NULL (0). */
/* One transition of the reference's other two forward models (no cost function / demo ships for them, so only
 * `model.step` exists): skid-steer robot dust/models/skid_steer_robot.py:73-122 (states [M,5], actions [M,2], params
 * [M,3] = x_icr, wheel_radius, axial_distance or NULL; cfg = {x_icr, wheel_radius, axial_distance, min_right, max_right,
 * min_left, max_left, -}) and cart-pole dust/models/cartpole.py:127-172 (states [M,4], actions [M,1], params [M,7] =
 * g, mass_cart, mass_pole, length, mu_c, mu_p, f_mag or NULL; cfg = the same seven defaults).  cfg is HOST memory. */
enum { DUST_AUX_SKID_STEER = 0, DUST_AUX_CARTPOLE = 1 };
int dust_aux_model_step(int32_t kind, float dt, const float* cfg, int32_t M, const float* states, const float* actions,
                        const float* params, float* next_states, void* stream);

int dust_model_cost(const dust_model_desc* model, int32_t M, int32_t terminal, const float* states,
                    const float* actions, float* costs, void* stream);

/* Standard-normal action noise, replacing the rsample draws of dust/inference/likelihoods.py:97-103
 * and dust/controllers/disco.py:211-230.  Counter-based (Philox4x32-10 + Box-Muller): block i of four
 * values depends only on (seed, offset, i), so a fill is reproducible, and ranks / successive steps
 * draw independent streams by using distinct offsets.  out: n floats, 16-byte aligned. */
int dust_noise_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream);

/* Exchange of row-block sharded SVGD over NVLink peer memory, replacing the all-gather of particles and scores
 * that row-block sharding of dust/inference/svgd.py:127-135 needs (the reference has one process and no exchange).
 * One process per GPU on one NVSwitch box.  Every rank allocates a slab with dust_peer_alloc (plain cudaMalloc: CUDA
 * IPC cannot export pool memory), exports it (64-byte handle, exchanged by the host over any channel) and opens the
 * other ranks' handles.  Per exchange `epoch` (1, 2, ...): write the local rows into the slab of parity epoch & 1,
 * dust_peer_signal (stores `epoch` into entry `rank` of the flag array ON EVERY RANK, after a system fence), then
 * dust_peer_gather: every CTA waits for the flag of the rank whose rows it copies and pulls them with 16-byte loads
 * into `gathered` [world * rows_per_rank, row_floats].  Two slabs alternating by parity make a second barrier
 * unnecessary (csrc/peer.cu).  A peer that never signals traps the kernel after ~2 s (launch failure, no hang). */
typedef struct dust_peer_args {
  int32_t world, rank;       /* ranks on the box (<= 16), this rank */
  int32_t epoch;             /* > 0, the same on every rank, increasing by one per exchange */
  int32_t rows_per_rank;     /* rows every rank contributes */
  int32_t row_floats;        /* floats per row (2D for [X | score]); rows_per_rank * row_floats % 4 == 0 */
  const float* const* slabs; /* HOST array [world]: this epoch's slab of every rank as mapped in this process */
  int32_t* const* flags;     /* HOST array [world]: the int32 flag array [world] of every rank as mapped here */
  float* gathered;           /* [world * rows_per_rank, row_floats] local output of dust_peer_gather */
  /* push form (dust_peer_push / dust_peer_wait): no slabs; the gathered buffers themselves are peer-mapped */
  float* const* gathered_peers; /* HOST array [world]: this epoch's gathered buffer of every rank as mapped here */
  int32_t* counters;         /* LOCAL device int32 [world], zero-filled once (arrival counters of the push CTAs) */
  int32_t n_parts;           /* 1..4 pieces a local row is made of, e.g. X and score */
  int32_t part_floats[4];    /* floats per row of each piece (multiples of 4; their sum = row_floats) */
  const float* parts[4];     /* local [rows_per_rank, part_floats[k]] contiguous, 16-byte aligned */
} dust_peer_args;

int dust_peer_alloc(size_t bytes, void** ptr);                      /* zero-filled, 256-byte aligned */
int dust_peer_free(void* ptr);
int dust_peer_export(const void* ptr, unsigned char handle[64]);     /* ptr from dust_peer_alloc */
int dust_peer_open(const unsigned char handle[64], void** ptr);      /* maps a peer's allocation, enables peer access */
int dust_peer_close(void* ptr);
int dust_peer_signal(const dust_peer_args* args, void* stream);
int dust_peer_gather(const dust_peer_args* args, void* stream);
/* Push form: every rank stores its rows (pieces concatenated per row) into block `rank` of gathered_peers[q] for every
 * q and raises flags_on_q[rank] = epoch once all of them have landed; dust_peer_wait holds the stream until all flags
 * of `epoch` are up in the local flag array.  Buffers of two parities alternate, as the slabs do. */
int dust_peer_push(const dust_peer_args* args, void* stream);
int dust_peer_wait(const dust_peer_args* args, void* stream);

/* Optional per-kernel timing for benchmarks: when enabled every kernel launch of the library is
 * bracketed by CUDA events on its stream.  dust_profiler_report synchronises the device and
 * writes "<kernel> <launches> <total_ms>" lines.  dust_launch_count: kernels launched so far. */
void dust_profiler_enable(int on);
void dust_profiler_reset(void);
int dust_profiler_report(char* buf, size_t cap);
unsigned long long dust_launch_count(void);

/* library info */
int dust_abi_version(void);
const char* dust_last_error(void);
const char* dust_build_info(void);    /* "sm_100a ..., built <date>" */

#ifdef __cplusplus
}
#endif
#endif /* DUST_B200_H */
