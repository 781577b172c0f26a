"""Functional wrappers over the C ABI: allocate outputs / scratch as torch CUDA tensors,
fill the argument struct, enqueue on the current stream.  All tensors carry a leading
instance dimension B."""
import ctypes as C
import math

import torch

from . import _lib as L


def _f32(t, device):
    """float32, contiguous, on `device` (host tensors are copied over)."""
    if t is None:
        return None
    t = torch.as_tensor(t)
    if t.dtype != torch.float32 or t.device != device or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.float32).contiguous()
    return t


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def rollout_cost(spec, state0, noise, theta=None, sigma=None, params=None, param_tiling=L.PARAMS_BLOCKED,
                 likelihood=L.LIK_EXP_UTILITY, a_seq=None, pert=None, alpha=1.0, temperature=1.0,
                 want=("costs", "log_lik"), out=None, sigma_weights=None, ctrl_mat=None, ctrl_reg=0.0, p_range=None,
                 reduce_only=False, plan_only=False):
    """K1.  noise [B,S,N,H,A]; theta [B,N,H,A] or None (noise = actions); params [B,P,dp] or None.
    `want` subset of {costs, log_lik, lik_weights, grad_lik, mppi_weights, mppi_delta, mix, states}.
    sigma_weights [P]: unscented-transform mode (params = the sigma points, disco.py:211-323);
    ctrl_mat [B,N,H,A] + ctrl_reg: control regulariser (disco.py:334-344).
    p_range (p0, p1): roll out only those parameter draws; `costs` is then their share of the mean
    (sum the shares over the ranks).  reduce_only: `out["costs"]` is complete, only the reductions after
    the costs run (dust_cost_reduce).
    plan_only: launch nothing; return dict(fused, param_chunks, chunk, nsub, parts) -- how the library would run it.
    Returns a dict of CUDA tensors."""
    L.require_cuda()
    dev = noise.device
    B, S, N, H, A = noise.shape
    assert A == spec.da, f"action dim {A} != model's {spec.da}"
    P = 1 if params is None else params.shape[1]
    out = {} if out is None else out
    shapes = {
        "costs": (B, S, N), "log_lik": (B, N), "lik_weights": (B, S, N), "grad_lik": (B, N, H, A),
        "mppi_weights": (B, S, N), "mppi_delta": (B, N, H, A), "mix": (B, N),
        "states": (B, P, S, N, H + 1, spec.ds),
    }
    for k in want:
        if k not in out:
            out[k] = torch.empty(shapes[k], dtype=torch.float32, device=dev)
    a = L.RolloutArgs()
    a.model = C.pointer(spec.desc)
    a.B, a.N, a.S, a.P, a.H = B, N, S, P, H
    a.param_tiling, a.likelihood = param_tiling, likelihood
    a.state0, a.theta, a.noise = L.ptr(state0), L.ptr(theta), L.ptr(noise)
    a.sigma, a.params, a.a_seq, a.pert = L.ptr(sigma), L.ptr(params), L.ptr(a_seq), L.ptr(pert)
    a.alpha, a.temperature = float(alpha), float(temperature)
    if sigma_weights is not None:
        assert sigma_weights.numel() == P and params is not None
    a.sigma_weights, a.ctrl_mat, a.ctrl_reg = L.ptr(sigma_weights), L.ptr(ctrl_mat), float(ctrl_reg)
    a.p_begin, a.p_end = (0, 0) if p_range is None else (int(p_range[0]), int(p_range[1]))
    for k in shapes:
        setattr(a, k, L.ptr(out.get(k)) if k in want else None)
    if reduce_only:
        a.costs = L.ptr(out["costs"])
    if plan_only:
        plan = (C.c_int32 * 5)()
        L.check(L.load().dust_rollout_plan(C.byref(a), C.byref(plan)))
        return dict(zip(("fused", "param_chunks", "chunk", "nsub", "parts"), (int(v) for v in plan)))
    nbytes = L.load().dust_rollout_workspace_bytes(C.byref(a))
    ws = _ws(nbytes, dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
    L.call("dust_cost_reduce" if reduce_only else "dust_rollout_cost", C.byref(a), L.stream(), launches=3)
    return out


def svmpc_step(spec, state0, noise, theta, sigma, mu, mix, inv_var, log_norm, gamma, c1, c2, lr, params=None,
               param_tiling=L.PARAMS_BLOCKED, likelihood=L.LIK_EXP_UTILITY, alpha=1.0, temperature=1.0,
               aliased=False, do_forward=True, roll_strategy=L.ROLL_REPEAT, weighted_prior=False,
               want=("costs", "log_lik", "theta_out")):
    """The whole SVGD step (+ SVMPC.forward) in ONE launch (dust_svmpc_step).  Raises
    NotImplementedError when the shape does not qualify (the caller then uses the staged path).
    `want` subset of {costs, log_lik, grad_lik, theta_out, phi}."""
    L.require_cuda()
    dev = noise.device
    B, S, N, H, A = noise.shape
    P = 1 if params is None else params.shape[1]
    f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)  # noqa: E731
    out = {}
    for k, shape in (("costs", (B, S, N)), ("log_lik", (B, N)), ("grad_lik", (B, N, H, A)), ("theta_out", (B, N, H, A)),
                     ("phi", (B, N, H, A))):
        if k in want:
            out[k] = f32(*shape)
    s = L.SvmpcStepArgs()
    a = s.rollout
    a.model = C.pointer(spec.desc)
    a.B, a.N, a.S, a.P, a.H = B, N, S, P, H
    a.param_tiling, a.likelihood = param_tiling, likelihood
    a.state0, a.theta, a.noise, a.sigma, a.params = L.ptr(state0), L.ptr(theta), L.ptr(noise), L.ptr(sigma), L.ptr(params)
    a.alpha, a.temperature = float(alpha), float(temperature)
    a.costs, a.log_lik, a.grad_lik = L.ptr(out.get("costs")), L.ptr(out.get("log_lik")), L.ptr(out.get("grad_lik"))
    s.do_forward, s.roll_strategy = int(bool(do_forward)), int(roll_strategy)
    s.weighted_prior, s.prior_aliased = int(bool(weighted_prior)), int(bool(aliased))
    s.mu, s.mix, s.inv_var, s.log_norm = L.ptr(None if aliased else mu), L.ptr(mix), L.ptr(inv_var), float(log_norm)
    s.gamma, s.c1, s.c2, s.lr = float(gamma), float(c1), float(c2), float(lr)
    s.theta_out, s.phi = L.ptr(out.get("theta_out")), L.ptr(out.get("phi"))
    if do_forward:
        out.update(p_weights=f32(B, N), i_star=torch.empty((B,), dtype=torch.int32, device=dev), a_seq=f32(B, H, A),
                   theta_next=f32(B, N, H, A), mix_next=f32(B, N))
        s.p_weights, s.i_star, s.a_seq = L.ptr(out["p_weights"]), L.ptr(out["i_star"]), L.ptr(out["a_seq"])
        s.theta_next, s.mix_next = L.ptr(out["theta_next"]), L.ptr(out["mix_next"])
    nbytes = L.load().dust_rollout_workspace_bytes(C.byref(a))
    ws = _ws(nbytes, dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
    L.call("dust_svmpc_step", C.byref(s), L.stream(), launches=1)
    return out


def rollout_adjoint(spec, state0, noise, lik_weights, theta=None, sigma=None, params=None,
                    param_tiling=L.PARAMS_BLOCKED, likelihood=L.LIK_EXP_UTILITY, alpha=1.0, p_range=None, plan_only=False):
    """K2.  Returns grad_theta [B,N,H,A] = d sum_n log_l_n / d theta (pathwise); with p_range (p0, p1) the
    share of those parameter draws (sum the shares over the ranks)."""
    L.require_cuda()
    dev = noise.device
    B, S, N, H, A = noise.shape
    P = 1 if params is None else params.shape[1]
    g = torch.empty((B, N, H, A), dtype=torch.float32, device=dev)
    a = L.AdjointArgs()
    a.model = C.pointer(spec.desc)
    a.B, a.N, a.S, a.P, a.H = B, N, S, P, H
    a.param_tiling, a.likelihood = param_tiling, likelihood
    a.state0, a.theta, a.noise, a.sigma = L.ptr(state0), L.ptr(theta), L.ptr(noise), L.ptr(sigma)
    a.params, a.lik_weights, a.alpha = L.ptr(params), L.ptr(lik_weights), float(alpha)
    a.grad_theta, a.grad_params = L.ptr(g), None
    a.p_begin, a.p_end = (0, 0) if p_range is None else (int(p_range[0]), int(p_range[1]))
    if plan_only:
        plan = (C.c_int32 * 5)()
        L.check(L.load().dust_adjoint_plan(C.byref(a), C.byref(plan)))
        return dict(zip(("param_chunks", "chunk", "tiles", "segments", "max_h"), (int(v) for v in plan)))
    nbytes = L.load().dust_adjoint_workspace_bytes(C.byref(a))
    ws = _ws(nbytes, dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
    L.call("dust_rollout_adjoint", C.byref(a), L.stream(), launches=2)
    return g


def gmm(x, mu, mix, inv_var, log_norm, want_log_prob=True, want_score=True):
    """K3.  x [B,M,D], mu [B,K,D], mix [B,K] | None, inv_var [D].  -> (log_prob [B,M], score [B,M,D])."""
    L.require_cuda()
    B, M, D = x.shape
    K = mu.shape[1]
    lp = torch.empty((B, M), dtype=torch.float32, device=x.device) if want_log_prob else None
    sc = torch.empty((B, M, D), dtype=torch.float32, device=x.device) if want_score else None
    a = L.GmmArgs(B, M, K, D, L.ptr(x), L.ptr(mu), L.ptr(mix), L.ptr(inv_var), float(log_norm),
                  L.ptr(lp), L.ptr(sc))
    L.call("dust_gmm_score", C.byref(a), L.stream())
    return lp, sc


def gmm_log_norm(var_full):
    """-0.5 (D log 2pi + sum log var) for a diagonal covariance given as a [D] tensor."""
    var_full = torch.as_tensor(var_full, dtype=torch.float64)
    return float(-0.5 * (var_full.numel() * math.log(2 * math.pi) + var_full.log().sum()))


def svgd_phi(x, score, gamma=0.0, c1=0.0, c2=0.0, gamma_dev=None, per_dim=False, bw_scale=1.0, lr=0.0,
             want_phi=True, want_update=False, rows=None, want_bandwidths=False, workspace=None, x_prepared=False):
    """K5.  x, score [B,N,D].  Returns dict(phi=, x_out=, bandwidths=).
    workspace: a caller-owned uint8 buffer (at least dust_phi_workspace_bytes) instead of a fresh one; with
    x_prepared its head already holds the operand images of x (the median pass ran on the same buffer)."""
    L.require_cuda()
    B, N, D = x.shape
    dev = x.device
    phi = torch.empty(x.shape, dtype=torch.float32, device=dev) if want_phi else None
    xo = torch.empty(x.shape, dtype=torch.float32, device=dev) if want_update else None
    bws = torch.empty((B, D), dtype=torch.float32, device=dev) if (per_dim and want_bandwidths) else None
    a = L.PhiArgs()
    a.B, a.N, a.D = B, N, D
    a.row_begin, a.row_end = (0, N) if rows is None else rows
    a.per_dim = int(per_dim)
    if B == 1 and not x.is_contiguous():
        # X and score as column slices of ONE [N, ld] buffer (the all-gathered [X | score]): no slicing copies
        (px, ldx), (ps, lds) = L.ptr_rows(x[0]), L.ptr_rows(score[0])
        if ldx != lds:
            raise ValueError("dust_b200: x and score must share their row stride")
        a.x, a.score, a.ld = px, ps, ldx
    else:
        a.x, a.score = L.ptr(x), L.ptr(score)
    a.gamma, a.c1, a.c2, a.gamma_dev = float(gamma), float(c1), float(c2), L.ptr(gamma_dev)
    a.bw_scale, a.lr = float(bw_scale), float(lr)
    a.phi, a.x_out, a.bandwidths = L.ptr(phi), L.ptr(xo), L.ptr(bws)
    nbytes = L.load().dust_phi_workspace_bytes(C.byref(a))
    if workspace is not None and workspace.numel() >= nbytes:
        ws = workspace
        a.x_prepared = int(bool(x_prepared))
    else:
        ws = _ws(nbytes, dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), max(nbytes, ws.numel())
    L.call("dust_svgd_phi", C.byref(a), L.stream(), launches=2 if per_dim or N > 512 else 1)
    return dict(phi=phi, x_out=xo, bandwidths=bws)


class MedianWorkspace:
    """Device scratch for the median select: 3*65536+8 u64 bins, 8 u32 of state, N row norms, and
    (when the tensor-core pass applies) the tiled operand images."""

    def __init__(self, N, D, device, fast_ws=None):
        """fast_ws: a caller-owned uint8 buffer for the tensor-core pass (shared with `svgd_phi(workspace=...)`: both
        keep |x|^2 and the hi/lo operand images of x at its head, so phi need not prepare them again)."""
        self.hist = torch.zeros(3 * 65536 + 8, dtype=torch.int64, device=device)
        self.selected = torch.zeros(8, dtype=torch.int32, device=device)
        self.row_norms = torch.empty(N, dtype=torch.float32, device=device)
        self.median = torch.zeros(1, dtype=torch.float32, device=device)
        lib = L.load()
        self.fast = bool(lib.dust_median_fast_supported(N, D))
        self.fast_bytes = lib.dust_median_fast_workspace_bytes(N, D) if self.fast else 0
        if self.fast and fast_ws is not None and fast_ws.numel() >= self.fast_bytes:
            self.fast_ws = fast_ws
        else:
            self.fast_ws = _ws(self.fast_bytes, device) if self.fast else None
        self.sample_hist = None
        if self.fast:      # the 32768-bin histogram of the sampled pairs inside the workspace (summed over ranks when sharded)
            off = int(lib.dust_median_fast_sample_hist_offset(N, D))
            self.sample_hist = self.fast_ws[off:off + 4 * 32768].view(torch.int32)
        self.flag_host, self.flag_event = None, None      # pinned success flag of the deferred (sharded) form


def _median_args(x, ws, rows, sample=None):
    N, D = x.shape
    a = L.MedianArgs()
    a.N, a.D = N, D
    a.row_begin, a.row_end = (0, N) if rows is None else rows
    a.sample_begin, a.sample_end = (0, 0) if sample is None else sample
    if x.is_contiguous():
        px = L.ptr(x)
    else:
        px, a.ld = L.ptr_rows(x)           # a column slice of a wider buffer
    a.x, a.hist, a.selected, a.row_norms = px, ws.hist.data_ptr(), ws.selected.data_ptr(), L.ptr(ws.row_norms)
    return a


def _median_fast(a, ws, all_reduce):
    L.call("dust_median_fast_prepare", C.byref(a), ws.fast_ws.data_ptr(), ws.fast_bytes, L.stream(), launches=2)
    if all_reduce is not None and a.sample_end > a.sample_begin:
        all_reduce(ws.sample_hist)         # every rank drew its own share of the sampled pairs
    L.call("dust_median_fast_window", C.byref(a), ws.fast_ws.data_ptr(), ws.fast_bytes, L.stream())
    L.call("dust_median_fast_count", C.byref(a), ws.fast_ws.data_ptr(), ws.fast_bytes, L.stream())
    if all_reduce is not None:
        all_reduce(ws.hist)
    L.call("dust_median_fast_select", C.byref(a), ws.median.data_ptr(), L.stream())


def _median_radix(a, ws, all_reduce):
    for p in (0, 1):
        L.call("dust_median_hist_pass", C.byref(a), p, L.stream(), launches=2 if p == 0 else 1)
        if all_reduce is not None:
            all_reduce(ws.hist)
        L.call("dust_median_select", C.byref(a), p, ws.median.data_ptr(), L.stream())


def median_sq_dist(x, ws=None, rows=None, all_reduce=None, allow_fast=True):
    """K4.  Exact lower median of the N^2 clamped squared distances of x [N,D] (device scalar).
    `rows` restricts the histogrammed row block; `all_reduce(hist)` (e.g. an NCCL sum) is called
    between the histogram and the select of each pass when the rows are sharded over ranks.
    Qualifying shapes take the tensor-core window pass first; the two-pass radix select is always
    enqueued behind it and does nothing when the fast path succeeded (no host synchronisation)."""
    L.require_cuda()
    N, D = x.shape
    ws = ws or MedianWorkspace(N, D, x.device)
    a = _median_args(x, ws, rows)
    ws.hist.zero_()
    ws.selected.zero_()
    if allow_fast and ws.fast and a.row_begin % 128 == 0 and a.row_end % 128 == 0:
        _median_fast(a, ws, all_reduce)
    _median_radix(a, ws, all_reduce)
    return ws.median


def median_sq_dist_deferred(x, ws=None, rows=None, all_reduce=None, sample=None):
    """The sharded form of `median_sq_dist`: when the tensor-core window pass applies it runs ALONE -- the radix
    kernels behind it would drag two more histogram all-reduces along that do nothing in the common case -- and
    the success flag travels to pinned host memory behind an event.  -> (median, check): `check()` waits for that
    event only (kernels queued after it keep the GPU busy) and returns False when the rank fell outside the
    window; the caller then runs `median_sq_dist(..., allow_fast=False)`.  The flag is computed from the
    all-reduced counts: every rank takes the same branch."""
    L.require_cuda()
    N, D = x.shape
    ws = ws or MedianWorkspace(N, D, x.device)
    a = _median_args(x, ws, rows, sample if all_reduce is not None else None)
    ws.hist.zero_()
    ws.selected.zero_()
    if not (ws.fast and a.row_begin % 128 == 0 and a.row_end % 128 == 0):
        _median_radix(a, ws, all_reduce)
        return ws.median, None
    _median_fast(a, ws, all_reduce)
    if ws.flag_host is None:
        ws.flag_host, ws.flag_event = torch.empty(1, dtype=torch.int32).pin_memory(), torch.cuda.Event()
    flag, ev = ws.flag_host, ws.flag_event
    flag.copy_(ws.selected[5:6], non_blocking=True)
    ev.record()

    def check():
        ev.synchronize()
        return bool(int(flag[0]))

    return ws.median, check


def noise_normal(out, seed, offset):
    """Fill the float32 CUDA tensor `out` with standard-normal noise: Philox4x32-10 keyed by `seed`,
    stream selected by `offset` (one offset per draw / per rank), Box-Muller.  Returns `out`."""
    L.require_cuda()
    if out.dtype != torch.float32:
        raise ValueError("dust_b200: noise_normal fills float32 tensors")
    L.call("dust_noise_normal", L.ptr(out), out.numel(), int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1), L.stream())
    return out


def bandwidth_from_median(median, N, scale=1.0, mode=0):
    """-> device tensor {gamma, c1, c2, bw|h}."""
    out = torch.empty(4, dtype=torch.float32, device=median.device)
    L.call("dust_bandwidth_from_median", median.data_ptr(), int(N), float(scale), int(mode), out.data_ptr(),
           L.stream())
    return out


def svmpc_forward(log_lik, theta, mu, mix, inv_var, log_norm, roll_strategy=L.ROLL_REPEAT, weighted_prior=False,
                  resample_noise=None):
    """K7.  theta, mu [B,N,H,A].  Returns dict(p_weights, i_star, a_seq, theta_next, mix_next)."""
    L.require_cuda()
    B, N, H, A = theta.shape
    dev = theta.device
    out = dict(
        p_weights=torch.empty((B, N), dtype=torch.float32, device=dev),
        i_star=torch.empty((B,), dtype=torch.int32, device=dev),
        a_seq=torch.empty((B, H, A), dtype=torch.float32, device=dev),
        theta_next=torch.empty_like(theta),
        mix_next=torch.empty((B, N), dtype=torch.float32, device=dev),
    )
    a = L.SvmpcForwardArgs(B, N, H, A, int(roll_strategy), int(bool(weighted_prior)), L.ptr(log_lik), L.ptr(theta),
                           L.ptr(mu), L.ptr(mix), L.ptr(inv_var), float(log_norm), L.ptr(out["p_weights"]),
                           L.ptr(out["i_star"]), L.ptr(out["a_seq"]), L.ptr(out["theta_next"]), L.ptr(out["mix_next"]),
                           L.ptr(resample_noise))
    L.call("dust_svmpc_forward", C.byref(a), L.stream())
    return out


def disco_step(a_mat, a_mix, a_low, a_high, strategy=L.SELECT_ARGMAX, steps=1):
    """MultiDISCO.step.  a_mat [B,N,H,A] is updated IN PLACE.  -> (next_actions [B,steps,A], a_seq [B,H,A])."""
    L.require_cuda()
    B, N, H, A = a_mat.shape
    dev = a_mat.device
    a_seq = torch.empty((B, H, A), dtype=torch.float32, device=dev)
    nxt = torch.empty((B, steps, A), dtype=torch.float32, device=dev)
    a = L.DiscoStepArgs(B, N, H, A, int(strategy), int(steps), L.ptr(a_low), L.ptr(a_high), L.ptr(a_mat), L.ptr(a_mix),
                        L.ptr(a_seq), L.ptr(nxt))
    L.call("dust_disco_step", C.byref(a), L.stream())
    return nxt, a_seq


def silverman_bandwidth(x, scale=1.0, dp=0):
    """KDEpy's Silverman rule on the flattened device tensor x (n <= 4096) -> (bw [1], inv_var [dp] = 1/bw^2), both
    on the device (no host round trip: `mpf_optimize(bw=bw_tensor)` reads the bandwidth there)."""
    L.require_cuda()
    flat = x.reshape(-1)
    if not flat.is_contiguous():
        flat = flat.contiguous()
    bw = torch.empty(1, dtype=torch.float32, device=x.device)
    iv = torch.empty(max(int(dp), 1), dtype=torch.float32, device=x.device)
    L.call("dust_silverman_bandwidth", L.ptr(flat), flat.numel(), float(scale), L.ptr(bw), L.ptr(iv) if dp else None, int(dp),
           L.stream())
    return bw, (iv if dp else None)


def mpf_optimize(spec, x, obs0, action, obs1, prior_inv_var, obs_std, bw, lr, n_steps, log_space, cooperative=True,
                 phi_out=None):
    """x [B,Np,dp] is updated IN PLACE.  bw: python float, or a 1-element device tensor (read by the kernel).
    phi_out [B,Np,dp]: receives phi of the last step (lr = 0, n_steps = 1: evaluate phi only).
    -> grad_norms [B,n_steps]."""
    L.require_cuda()
    B, Np, dp = x.shape
    assert dp == spec.dp, f"parameter dim {dp} != model's {spec.dp}"
    gn = torch.empty((B, n_steps), dtype=torch.float32, device=x.device)
    a = L.MpfArgs()
    a.model = C.pointer(spec.desc)
    a.B, a.Np, a.n_steps, a.log_space = B, Np, int(n_steps), int(bool(log_space))
    a.x, a.obs0, a.action, a.obs1 = L.ptr(x), L.ptr(obs0), L.ptr(action), L.ptr(obs1)
    a.prior_inv_var = L.ptr(prior_inv_var)
    bw_dev = bw if torch.is_tensor(bw) else None
    a.obs_std, a.bw, a.lr = float(obs_std), 0.0 if bw_dev is not None else float(bw), float(lr)
    a.bw_dev = L.ptr(bw_dev)
    a.grad_norms = L.ptr(gn)
    a.phi_out = L.ptr(phi_out)
    if phi_out is not None:
        cooperative = False
    # > 0: one large instance, cooperative multi-SM kernel (cooperative=False keeps the one-CTA kernel)
    nbytes = L.load().dust_mpf_workspace_bytes(C.byref(a)) if cooperative else 0
    ws = _ws(nbytes, x.device) if nbytes else None
    a.workspace, a.workspace_bytes = (ws.data_ptr() if nbytes else None), nbytes
    L.call("dust_mpf_optimize", C.byref(a), L.stream())
    return gn


def model_step(spec, states, actions, params=None):
    """states [M,ds], actions [M,A], params [M,dp] | None -> next states [M,ds]."""
    L.require_cuda()
    M = states.shape[0]
    nxt = torch.empty_like(states)
    L.call("dust_model_step", C.byref(spec.desc), M, L.ptr(states), L.ptr(actions), L.ptr(params), L.ptr(nxt),
           L.stream())
    return nxt


def aux_model_step(kind, dt, cfg, states, actions, params=None):
    """One transition of the skid-steer robot (kind 0) / cart-pole (kind 1).  cfg: up to 8 python floats (defaults and
    action bounds); states [M,ds], actions [M,A], params [M,np] | None -> next states [M,ds]."""
    L.require_cuda()
    M = states.shape[0]
    nxt = torch.empty_like(states)
    c = (C.c_float * 8)(*([float(v) for v in cfg] + [0.0] * (8 - len(cfg))))
    L.call("dust_aux_model_step", int(kind), float(dt), C.byref(c), M, L.ptr(states), L.ptr(actions), L.ptr(params), L.ptr(nxt),
           L.stream())
    return nxt


def model_cost(spec, states, actions=None, terminal=False):
    L.require_cuda()
    M = states.shape[0]
    c = torch.empty((M,), dtype=torch.float32, device=states.device)
    L.call("dust_model_cost", C.byref(spec.desc), M, int(bool(terminal)), L.ptr(states), L.ptr(actions), L.ptr(c),
           L.stream())
    return c
