"""Round-2 hardware parity: the shapes and paths the first suite never reached on a B200 --
N = 65536 phi / median (BASELINE configs[3]), the non-SGD optimiser path and `fast_pred=False`
through the `SVMPC` class (svgd.py:115, svmpc.py:128-140), the NSUB = 2 / 4 rollout variants and the
five-segment checkpointed adjoint of the dual-stress shape (configs[4], scaled)."""
import functools
import math
import time

import pytest
import torch

from oracle import dust_oracle as O
from tests.test_gpu_api import ENV, FixedParams, demo_inst_cost, demo_term_cost
from tests.util import RTOL_COST, RTOL_PHI, golden_grid, load, record_parity, rel_elem, rel_max

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(t):
    return None if t is None else torch.as_tensor(t, dtype=torch.float32).to(DEV).contiguous()


def ulp_distance(a, b):
    """distance in float32 representable steps between two positive floats"""
    ia = int(torch.tensor([a], dtype=torch.float32).view(torch.int32)[0])
    ib = int(torch.tensor([b], dtype=torch.float32).view(torch.int32)[0])
    return abs(ia - ib)


# ---------------------------------------------------------------------------------------------
# configs[3] at its full size
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cloud", ["isotropic", "anisotropic"])
def test_phi_and_median_at_65536(cloud):
    """N = 65536, d = 40 (SURVEY 8(d) cfg 4): the exact median against the tiled radix oracle, sampled phi rows
    against float64 (the reference itself cannot run here: K alone is 17 GB).  O is drained from TMEM every 32
    column tiles (svgd_tc.cu), an accumulation-depth effect that only shows at this size."""
    from dust_b200 import _lib as L
    from dust_b200 import ops

    N, D = 65536, 40
    g = torch.Generator().manual_seed(0 if cloud == "isotropic" else 1)
    X = torch.randn(N, D, generator=g)
    if cloud == "anisotropic":
        X = X * (torch.arange(1, D + 1).float() / 10)
    S = -X
    x = cu(X)
    lib = L.load()
    lib.dust_profiler_reset()
    lib.dust_profiler_enable(1)
    ws = ops.MedianWorkspace(N, D, x.device)
    med = ops.median_sq_dist(x, ws=ws)
    coef = ops.bandwidth_from_median(med, N, 1.0, 0)
    out = ops.svgd_phi(x.unsqueeze(0), cu(S).unsqueeze(0), gamma_dev=coef)
    prof = L.profiler_report()
    lib.dust_profiler_enable(0)
    assert "phi_tc_kernel" in prof and "median_tc_kernel" in prof and "phi_large_kernel" not in prof
    assert int(ws.selected[5]) == 1, "a well-behaved cloud must not need the radix fallback"
    med_dev = float(med[0])
    robust = float(ops.median_sq_dist(x, allow_fast=False)[0])      # SIMT two-pass radix select, own distance code
    t0 = time.time()
    med_ref = float(O.median_sq_dist_tiled(X))                       # exact rank over torch's float32 distances
    t_oracle = time.time() - t0
    u_ref, u_rob = ulp_distance(med_dev, med_ref), ulp_distance(med_dev, robust)
    record_parity(f"median N=65536 {cloud}", device=med_dev, oracle=med_ref, radix_simt=robust, ulp_vs_oracle=u_ref,
                  ulp_vs_radix_simt=u_rob, oracle_seconds=t_oracle)
    # three exact rank selections over distances rounded three different ways (3xTF32 Gram, fp32 FMA chain,
    # MKL sgemm): the order statistic of 4.3e9 values moves by far less than one distance rounding
    assert u_ref <= 4 and u_rob <= 4, (med_dev, med_ref, robust)
    gam, c1, c2, bw = [float(v) for v in coef.cpu()]
    bw_ref = max(math.sqrt(0.5 * med_ref) / math.log(N + 1), 1e-5)   # svgd.py:42-52
    assert abs(bw - bw_ref) <= 1e-6 * bw_ref
    # phi rows: 256 of them, spread over every row tile position, against float64 over all N columns
    idx = torch.cat([torch.arange(0, N, N // 128), torch.randint(0, N, (128,), generator=g)])
    Xd, Sd = X.double(), S.double()
    xi = Xd[idx]
    d2 = ((xi * xi).sum(-1, keepdim=True) + (Xd * Xd).sum(-1)[None, :] - 2 * xi @ Xd.t()).clamp(min=0)
    d2[torch.arange(len(idx)), idx] = 0.0
    K = (-gam * d2).exp()
    ref = c1 * (K @ Sd) + c2 * (K.sum(1, keepdim=True) * xi - K @ Xd)
    got = out["phi"][0].cpu()[idx].double()
    err = float((got - ref).abs().max() / ref.abs().max())
    err_rows = float(((got - ref).norm(dim=1) / ref.norm(dim=1)).max())
    record_parity(f"phi N=65536 {cloud}", err_vs_float64_max=err, err_vs_float64_rowwise=err_rows, rtol=RTOL_PHI)
    assert err <= RTOL_PHI and err_rows <= RTOL_PHI
    # one rank's row block of 8 equals the same rows of the full call to rounding
    blk = ops.svgd_phi(x.unsqueeze(0), cu(S).unsqueeze(0), gamma_dev=coef, rows=(3 * N // 8, 4 * N // 8))["phi"][0]
    e_blk = rel_max(blk[3 * N // 8: 4 * N // 8].cpu(), out["phi"][0, 3 * N // 8: 4 * N // 8].cpu())
    record_parity(f"phi N=65536 {cloud} row block 3/8", err_vs_full_call=e_blk)
    assert e_blk <= 2e-6


@pytest.mark.parametrize("N,D,rows", [(8192, 40, None), (8192, 40, (1024, 2048)), (8192, 40, (0, 128)),
                                      (8192, 24, (384, 8192)), (24576, 40, None), (20480, 32, None), (4096, 37, None),
                                      (9344 * 2, 16, (0, 9344))])
def test_phi_partitions_and_operand_modes(N, D, rows):
    """The tensor-core phi over every class of partition dust_phi_tc_plan produces (plain ranges for little work, column
    chunks with and without left-over columns, whole row tiles per SM plus a chunked rest) and both forms of the kernel
    (row tile in TMEM + bf16 P_lo term; row tile in shared memory + TF32 P_lo term with DUST_B200_TC_A_SMEM): sampled
    rows against float64 for both, and the two forms against each other.  svgd.py:127-135."""
    import os

    from dust_b200 import _lib as L
    from dust_b200 import ops

    g = torch.Generator().manual_seed(N + D)
    X = torch.randn(N, D, generator=g) * 0.7
    S = torch.randn(N, D, generator=g)
    x, s = cu(X).unsqueeze(0), cu(S).unsqueeze(0)
    gam, c1, c2 = 1.0 / (2.0 * D), 1.0 / N, 1.0 / (N * D)
    r0, r1 = rows or (0, N)
    lib = L.load()
    lib.dust_profiler_reset(); lib.dust_profiler_enable(1)
    out = ops.svgd_phi(x, s, gamma=gam, c1=c1, c2=c2, rows=rows)["phi"][0, r0:r1]
    prof = L.profiler_report()
    lib.dust_profiler_enable(0)
    assert "phi_tc_kernel" in prof and "phi_large_kernel" not in prof
    idx = torch.cat([torch.arange(r0, r1, max(1, (r1 - r0) // 96))[:96], torch.randint(r0, r1, (96,), generator=g)])
    Xd, Sd = X.double(), S.double()
    xi = Xd[idx]
    d2 = ((xi * xi).sum(-1, keepdim=True) + (Xd * Xd).sum(-1)[None, :] - 2 * xi @ Xd.t()).clamp(min=0)
    d2[torch.arange(len(idx)), idx] = 0.0
    K = (-gam * d2).exp()
    ref = c1 * (K @ Sd) + c2 * (K.sum(1, keepdim=True) * xi - K @ Xd)
    got = out.cpu()[idx - r0].double()
    err = float((got - ref).abs().max() / ref.abs().max())
    record_parity(f"phi partitions N={N} D={D} rows={rows}", err_vs_float64_max=err, rtol=RTOL_PHI)
    assert err <= 0.25 * RTOL_PHI, err     # wide kernel (gamma d2 ~ 0.5): ksum * x_i - K X cancels to a few percent
    os.environ["DUST_B200_TC_A_SMEM"] = "1"
    try:
        other = ops.svgd_phi(x, s, gamma=gam, c1=c1, c2=c2, rows=rows)["phi"][0, r0:r1]
    finally:
        del os.environ["DUST_B200_TC_A_SMEM"]
    # the second form (row tile in shared memory, all three GEMM2 terms in TF32) differs from the first only in the
    # P_lo V correction term, 2^-11 of the sum, carried with 16 instead of 21 mantissa bits
    e_modes = rel_max(out.cpu(), other.cpu())
    err2 = float((other.cpu()[idx - r0].double() - ref).abs().max() / ref.abs().max())
    record_parity(f"phi operand modes N={N} D={D} rows={rows}", tmem_bf16lo_vs_smem_tf32=e_modes, smem_tf32_err_vs_float64=err2)
    assert e_modes <= 4e-6 and err2 <= 0.25 * RTOL_PHI, (e_modes, err2)


def test_peer_gather_kernels_two_ranks_on_one_device():
    """dust_peer_signal / dust_peer_gather and dust_peer_push / dust_peer_wait (csrc/peer.cu; the exchange that row-block sharding of svgd.py:127-135 needs)
    with both ranks played by one process on one device -- the IPC mapping aside, the same launches a two-rank job
    makes: slabs alternate by epoch parity, a gather waits for the flags of its epoch, the result is the rank-major
    concatenation.  (The mapping itself runs under torchrun in bench.py, which compares with the single-GPU rows.)"""
    import ctypes as C

    from dust_b200 import _lib as L

    lib = L.load()
    world, rows, width = 2, 384, 80
    stride = (4 * rows * width + 255) // 256 * 256
    bases = []
    for _ in range(world):
        p = C.c_void_p()
        L.check(lib.dust_peer_alloc(2 * stride + 256, C.byref(p)))
        bases.append(p.value)
    handle = (C.c_ubyte * 64)()
    L.check(lib.dust_peer_export(bases[0], C.byref(handle)))
    assert any(handle)
    try:
        g = torch.Generator().manual_seed(3)
        outs = [torch.zeros(world * rows, width, device=DEV) for _ in range(world)]
        for epoch in (1, 2, 3):
            par = epoch & 1
            blocks = [torch.randn(rows, width, generator=g) for _ in range(world)]
            slabs = (C.c_void_p * world)(*[b + par * stride for b in bases])
            flags = (C.c_void_p * world)(*[b + 2 * stride for b in bases])

            def args(rank):
                a = L.PeerArgs()
                a.world, a.rank, a.epoch, a.rows_per_rank, a.row_floats = world, rank, epoch, rows, width
                a.slabs, a.flags = C.cast(slabs, C.POINTER(C.c_void_p)), C.cast(flags, C.POINTER(C.c_void_p))
                a.gathered = outs[rank].data_ptr()
                return a

            for r in range(world):      # every rank: rows into its slab, then its flag on every rank
                dst = torch.as_tensor(_Dev(bases[r] + par * stride, (rows, width)), device=DEV)
                dst.copy_(blocks[r].to(DEV))
                L.call("dust_peer_signal", C.byref(args(r)), L.stream())
            for r in range(world):
                L.call("dust_peer_gather", C.byref(args(r)), L.stream())
            torch.cuda.synchronize()
            want = torch.cat(blocks, 0)
            for r in range(world):
                assert torch.equal(outs[r].cpu(), want), (epoch, r)
            fl = torch.as_tensor(_Dev(bases[0] + 2 * stride, (world,), "<i4"), device=DEV)
            assert fl.cpu().tolist() == [epoch] * world
        a = args(0)
        a.epoch = 0
        with pytest.raises(ValueError):
            L.call("dust_peer_gather", C.byref(a), L.stream())
        # push form: the gathered buffers are the shared allocations (parity p at bases[r] + p * gstride), the rows come
        # straight from X and score; flags and arrival counters sit behind the two buffers
        gstride = (4 * world * rows * width + 255) // 256 * 256
        pbases = []
        for _ in range(world):
            p = C.c_void_p()
            L.check(lib.dust_peer_alloc(2 * gstride + 512, C.byref(p)))
            pbases.append(p.value)
        bases += pbases
        for epoch in (1, 2, 3):
            par = epoch & 1
            xs = [torch.randn(rows, width // 2, generator=g).to(DEV) for _ in range(world)]
            ss = [torch.randn(rows, width // 2, generator=g).to(DEV) for _ in range(world)]
            bufs = (C.c_void_p * world)(*[b + par * gstride for b in pbases])
            flags = (C.c_void_p * world)(*[b + 2 * gstride for b in pbases])
            for r in range(world):
                a = L.PeerArgs()
                a.world, a.rank, a.epoch, a.rows_per_rank, a.row_floats = world, r, epoch, rows, width
                a.gathered_peers, a.flags = C.cast(bufs, C.POINTER(C.c_void_p)), C.cast(flags, C.POINTER(C.c_void_p))
                a.counters = pbases[r] + 2 * gstride + 256
                a.n_parts = 2
                a.parts[0], a.parts[1] = xs[r].data_ptr(), ss[r].data_ptr()
                a.part_floats[0] = a.part_floats[1] = width // 2
                L.call("dust_peer_push", C.byref(a), L.stream())
            for r in range(world):
                a = L.PeerArgs()
                a.world, a.rank, a.epoch = world, r, epoch
                a.flags = C.cast(flags, C.POINTER(C.c_void_p))
                L.call("dust_peer_wait", C.byref(a), L.stream())
            torch.cuda.synchronize()
            want = torch.cat([torch.cat([xs[r], ss[r]], 1) for r in range(world)], 0)
            for r in range(world):
                got = torch.as_tensor(_Dev(pbases[r] + par * gstride, (world * rows, width)), device=DEV)
                assert torch.equal(got, want), (epoch, r)
            cnt = torch.as_tensor(_Dev(pbases[0] + 2 * gstride + 256, (world,), "<i4"), device=DEV)
            assert cnt.cpu().tolist() == [0] * world          # the last CTA per destination re-arms the counter
    finally:
        torch.cuda.synchronize()
        for b in bases:
            lib.dust_peer_free(b)


class _Dev:
    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2,
                                         "strides": None}


def test_phi_and_median_read_a_packed_buffer_in_place():
    """X and score as column slices of ONE [N, 2D] buffer (the all-gathered [X | score] of ShardedSVGD, `ld` in the
    ABI): the median, the bandwidth and the phi rows are bit-identical to the contiguous call (same launches, same
    partition -- only the addresses the prepare kernels read differ).  svgd.py:42-52, 127-135."""
    from dust_b200 import ops

    N, D = 4096, 40
    g = torch.Generator().manual_seed(5)
    X, S = torch.randn(N, D, generator=g), torch.randn(N, D, generator=g)
    packed = cu(torch.cat([X, S], dim=1))
    xs, ss = packed[:, :D], packed[:, D:]
    assert not xs.is_contiguous()
    x, s = cu(X), cu(S)
    ws_a, ws_b = ops.MedianWorkspace(N, D, x.device), ops.MedianWorkspace(N, D, x.device)
    med_a, med_b = ops.median_sq_dist(x, ws=ws_a), ops.median_sq_dist(xs, ws=ws_b)
    assert int(ws_a.selected[5]) == 1 and int(ws_b.selected[5]) == 1
    assert torch.equal(med_a, med_b)
    coef = ops.bandwidth_from_median(med_a, N, 1.0, 0)
    for rows in (None, (1024, 1536)):
        a = ops.svgd_phi(x.unsqueeze(0), s.unsqueeze(0), gamma_dev=coef, rows=rows, want_update=True, lr=0.5)
        b = ops.svgd_phi(xs.unsqueeze(0), ss.unsqueeze(0), gamma_dev=coef, rows=rows, want_update=True, lr=0.5)
        r0, r1 = rows or (0, N)
        assert torch.equal(a["phi"][0, r0:r1], b["phi"][0, r0:r1])
        assert torch.equal(a["x_out"][0, r0:r1], b["x_out"][0, r0:r1])
        assert b["phi"].is_contiguous() and b["x_out"].is_contiguous()
    # the paths without a row stride say so instead of reading the wrong rows
    with pytest.raises(Exception):
        ops.svgd_phi(xs[:256].unsqueeze(0), ss[:256].unsqueeze(0), gamma=0.5, c1=1.0, c2=1.0)


# ---------------------------------------------------------------------------------------------
# non-SGD optimisers and fast_pred=False through the SVMPC class, on the device
# ---------------------------------------------------------------------------------------------
def build_svmpc(d, n_steps=1, **opt):
    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.inference.likelihoods import ExponentiatedUtility
    from dust_b200.inference.svgd import get_gmm
    from dust_b200.inference.svmpc import SVMPC
    from dust_b200.kernels.base_kernels import RBFKernel
    from dust_b200.models.pendulum import PendulumModel

    model = PendulumModel(uncertain_params=("length", "mass"))
    N, H, A = d["theta_init"].shape
    S = d["t0_eps"].shape[1]
    P = d["t0_params"].shape[1]
    pv = float(d["prior_var"])
    ctrl = MultiDISCO(observation_space=model.observation_space, action_space=model.action_space, hz_len=H, n_policies=N,
                      action_samples=S, params_samples=P, temperature=1.0, a_cov=torch.diag(d["sigma"] ** 2),
                      inst_cost_fn=demo_inst_cost, term_cost_fn=demo_term_cost, params_sampling=True)
    prior = get_gmm(d["mu_init"], torch.ones(N), pv * torch.eye(A))
    lik = ExponentiatedUtility(1.0, n_samples=S, controller=ctrl, model=model)
    return SVMPC(init_particles=d["theta_init"].clone(), prior=prior, likelihood=lik, kernel=RBFKernel(), n_particles=N,
                 bw_scale=1.0, n_steps=n_steps, weighted_prior=False, **opt)


@pytest.mark.parametrize("name,opt", [("svmpc_pendulum_adam", dict(optimizer_class=torch.optim.Adam)),
                                      ("svmpc_pendulum_momentum", dict(optimizer_class=torch.optim.SGD, momentum=0.9))])
def test_svmpc_class_with_torch_optimizers_on_device(name, opt):
    """The reference stepping Adam (its default, svgd.py:115) and momentum SGD, two SVGD steps per control step:
    phi from the kernel, `torch.optim` on the device particles, optimiser state renewed after every roll."""
    d = load(name)
    sv = build_svmpc(d, n_steps=2, lr=float(d["lr"]), **opt)
    assert sv._core._make_opt is not None, "a non-plain optimiser must not take the fused SGD update"
    worst = {}
    for t in range(int(d["n_ctrl"])):
        state = d[f"t{t}_state"]
        for k in range(2):
            sv.step(state, FixedParams(d[f"t{t}_params"][k], torch.Size([2])), eps=d[f"t{t}_eps"][k])
        e_c = rel_elem(sv.likelihood.last_costs.cpu(), d[f"t{t}_costs"])
        e_t1 = rel_max(sv.theta.cpu(), d[f"t{t}_theta1"])
        a_seq, p_w = sv.forward(state, FixedParams(d[f"t{t}_params"][1], torch.Size([2])))
        e_a, e_t2 = rel_max(a_seq.cpu(), d[f"t{t}_a_seq"]), rel_max(sv.theta.cpu(), d[f"t{t}_theta2"])
        e_w = float((p_w.cpu() - d[f"t{t}_p_weights"]).abs().max())
        for key, v in (("costs", e_c), ("theta1", e_t1), ("a_seq", e_a), ("theta2", e_t2), ("p_weights_abs", e_w)):
            worst[key] = max(worst.get(key, 0.0), v)
        assert int(sv.i_star) == int(d[f"t{t}_i_star"])
        # free running over three control steps: the inputs of later steps carry earlier rounding
        assert e_c <= 2e-4 and e_t1 <= 2e-4 and e_a <= 2e-4 and e_t2 <= 2e-4 and e_w <= 1e-3, (t, e_c, e_t1, e_a, e_t2, e_w)
    record_parity(f"SVMPC class {name}", **worst)


def test_svmpc_class_fresh_likelihood_weights_on_device():
    """get_weights(fast_pred=False) / forward(fast_pred=False), svmpc.py:128-140: the likelihood sampled again at the
    updated particles (three noise + parameter draws per control step)."""
    d = load("svmpc_pendulum_slow_pred")
    sv = build_svmpc(d, n_steps=1, optimizer_class=torch.optim.SGD, lr=float(d["lr"]))
    worst = {}
    for t in range(int(d["n_ctrl"])):
        state, eps, params = d[f"t{t}_state"], d[f"t{t}_eps"], d[f"t{t}_params"]
        pd = [FixedParams(params[k], torch.Size([2])) for k in range(3)]
        sv.optimize(state, pd[0], eps=eps[0])
        peek = sv.get_weights(state, pd[1], fast_pred=False, eps=eps[1])
        e_peek = float((peek.cpu() - d[f"t{t}_peek"]).abs().max())
        a_seq, p_w = sv.forward(state, pd[2], fast_pred=False, eps=eps[2])
        e_c = rel_elem(sv.likelihood.last_costs.cpu(), d[f"t{t}_costs_fwd"])
        e_w = float((p_w.cpu() - d[f"t{t}_p_weights"]).abs().max())
        e_a, e_t2 = rel_max(a_seq.cpu(), d[f"t{t}_a_seq"]), rel_max(sv.theta.cpu(), d[f"t{t}_theta2"])
        for key, v in (("peek_abs", e_peek), ("costs_fwd", e_c), ("p_weights_abs", e_w), ("a_seq", e_a), ("theta2", e_t2)):
            worst[key] = max(worst.get(key, 0.0), v)
        assert e_peek <= 2e-3 and e_w <= 2e-3 and e_c <= 2e-4 and e_a <= 5e-4 and e_t2 <= 5e-4, (t, e_peek, e_w, e_c, e_a, e_t2)
    record_parity("SVMPC class svmpc_pendulum_slow_pred", **worst)


# ---------------------------------------------------------------------------------------------
# configs[4], scaled: the variants only the dual-stress shape selects
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def particle_env():
    from dust_b200.models.particle import Particle

    part = Particle(**ENV, uncertain_params=["mass"], mass=2.0)
    return dict(spec=part.device_spec(part.default_inst_cost, part.default_term_cost, DEV), cfg=O.ParticleCfg(golden_grid()))


@pytest.mark.parametrize("B,nsub", [(8, 2), (16, 4)])
def test_stress_shape_variants_are_reached_and_match_the_oracle(particle_env, B, nsub):
    """Particle model, H = 50 (H*A = 100: a 51 KB action tile, three CTAs per SM), 296 parameter draws over 1024
    trajectories per instance: the plan must select `rollout_cost_kernel<.., NSUB = 2 | 4>` (thread groups sharing one
    action tile) and the adjoint its five checkpointed segments of ten steps; costs and the pathwise gradient are
    then compared with the oracle on the first and last instance."""
    from dust_b200 import ops

    spec, cfg = particle_env["spec"], particle_env["cfg"]
    torch.manual_seed(100 + B)
    S, N, H, P, A = 32, 32, 50, 296, 2
    state = torch.tensor([-9.0, -9.0, 0.0, 0.0]).repeat(B, 1) + torch.randn(B, 4) * torch.tensor([2.0, 2.0, 0.5, 0.5])
    theta = torch.randn(B, N, H, A) * 3
    eps = torch.randn(B, S, N, H, A)
    sigma = torch.tensor([5.0, 5.0])
    params = (torch.randn(B, P, 1) * 0.1 + math.log(2.0)).exp()
    alpha = 1e-5          # dense soft-min weights: every rollout is reversed by the adjoint
    kw = dict(theta=cu(theta), sigma=cu(sigma), params=cu(params), param_tiling=0, alpha=alpha)
    want = ("costs", "log_lik", "lik_weights", "grad_lik")
    plan = ops.rollout_cost(spec, cu(state), cu(eps), want=want, plan_only=True, **kw)
    assert plan["nsub"] == nsub and plan["fused"] == 0 and plan["chunk"] >= 2 * nsub, plan
    out = ops.rollout_cost(spec, cu(state), cu(eps), want=want, **kw)
    aplan = ops.rollout_adjoint(spec, cu(state), cu(eps), out["lik_weights"], plan_only=True, **kw)
    assert aplan["segments"] == 5 and aplan["max_h"] == 64, aplan
    grad = ops.rollout_adjoint(spec, cu(state), cu(eps), out["lik_weights"], **kw)
    model = O.Model("particle", cfg)
    worst = {}
    for b in (0, B - 1):
        acts = theta[b] + sigma * eps[b]
        ref = O.disco_forward(model, state[b], acts, params[b])
        e_c = rel_elem(out["costs"][b].cpu(), ref["costs"])
        assert e_c <= RTOL_COST, (b, e_c)
        costs = out["costs"][b].cpu().double()
        g_dev = O.analytic_lik_grad(costs, acts.double(), theta[b].double(), sigma.double(), alpha)
        e_g = rel_max(out["grad_lik"][b].cpu(), g_dev)
        assert e_g <= 1e-5, (b, e_g)
        # pathwise gradient with the device's soft-min weights held fixed (the adjoint is linear in them)
        g64 = O.pathwise_lik_grad_adjoint(model, state[b].double(), theta[b].double(), eps[b].double(), sigma.double(),
                                          params[b].double(), False, alpha, weights=out["lik_weights"][b].cpu().double())[0]
        e_p = rel_max(grad[b].cpu(), g64)
        assert e_p <= RTOL_PHI, (b, e_p)
        for key, v in (("costs", e_c), ("grad_lik_vs_own_costs", e_g), ("pathwise_grad", e_p)):
            worst[key] = max(worst.get(key, 0.0), v)
    record_parity(f"stress-shape particle H=50 P=296 NSUB={nsub}", **worst)


# ---------------------------------------------------------------------------------------------
# few instances: the whole control step in one launch of a thread-block cluster per instance
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,B,N,S,H,P,wp", [("pendulum", 1, 3, 128, 30, 8, False),      # configs[0]
                                               ("particle", 1, 6, 64, 40, 4, True),        # configs[1]
                                               ("pendulum", 3, 8, 33, 20, 0, True),        # ragged row blocks, no draws
                                               ("particle", 2, 5, 7, 9, 3, False),         # fewer rows than CTAs x warps
                                               ("pendulum", 1, 32, 100, 12, 20, False)])   # draws looped inside a thread
def test_cluster_kernel_control_step_matches_staged_path(particle_env, kind, B, N, S, H, P, wp):
    """`dust_svmpc_step` for B < 74 (svmpc_cluster_kernel: 8 CTAs per instance, costs and partial gradients exchanged
    through distributed shared memory) against the staged kernels: costs bit for bit, everything else to rounding;
    first with a free-standing prior, then with the prior aliasing the particles; and against the float64 oracle."""
    from dust_b200 import _lib as L
    from dust_b200.inference.core import SvmpcCore
    from dust_b200.models.pendulum import PendulumModel, inst_cost, term_cost

    torch.manual_seed(N * 11 + S)
    if kind == "pendulum":
        spec = PendulumModel(uncertain_params=("length", "mass")).device_spec(inst_cost, term_cost, DEV)
        ds, A, dp, model = 2, 1, 2, O.Model("pendulum")
    else:
        spec, ds, A, dp, model = particle_env["spec"], 4, 2, 1, O.Model("particle", particle_env["cfg"])
    state = torch.randn(B, ds, device=DEV) * (torch.tensor([6.0, 6.0, 1.0, 1.0], device=DEV) if kind == "particle" else 1.5)
    theta, mu = torch.randn(B, N, H, A, device=DEV) * 2, torch.randn(B, N, H, A, device=DEV)
    mix = torch.rand(B, N, device=DEV) + 0.1
    params = (torch.rand(B, P, dp, device=DEV) + 0.8) if P else None
    alpha = 1.0 if kind == "pendulum" else 1e-3
    mk = lambda: SvmpcCore(spec, theta.clone(), mu.clone(), mix.clone(), torch.full((A,), 4.0), torch.full((A,), 2.0),  # noqa: E731
                           alpha=alpha, lr=0.5, kernel="gpytorch", weighted_prior=wp)
    fused, staged = mk(), mk()
    staged._fused_ok = False
    lib = L.load()
    worst = {}
    for step in range(2):
        eps = torch.randn(B, S, N, H, A, device=DEV)
        theta_in, mu_in, mix_in, aliased = fused.theta.clone(), fused.mu.clone(), fused.mix.clone(), fused.aliased
        lib.dust_profiler_reset(); lib.dust_profiler_enable(1)
        fused.optimize_step(state, eps, params)
        prof = L.profiler_report(); lib.dust_profiler_enable(0)
        assert list(prof) == ["svmpc_cluster_kernel"] and prof["svmpc_cluster_kernel"][0] == 1, prof
        theta1 = fused.theta.clone()
        a1, p1, i1 = fused.forward_step()
        staged.optimize_step(state, eps, params)
        theta1s = staged.theta.clone()
        a2, p2, i2 = staged.forward_step()
        assert torch.equal(fused.last["costs"], staged.last["costs"])
        e = dict(log_lik=rel_max(fused.last["log_lik"].cpu(), staged.last["log_lik"].cpu()),
                 grad_lik=rel_max(fused.last["grad_lik"].cpu(), staged.last["grad_lik"].cpu()),
                 phi=rel_max(fused.last["phi"].cpu(), staged.last["phi"].cpu()), theta1=rel_max(theta1.cpu(), theta1s.cpu()),
                 p_weights=float((p1 - p2).abs().max()), a_seq=rel_max(a1.cpu(), a2.cpu()),
                 theta_next=rel_max(fused.theta.cpu(), staged.theta.cpu()), mix=rel_max(fused.mix.cpu(), staged.mix.cpu()))
        assert torch.equal(i1, i2)
        assert e["log_lik"] <= 1e-6 and e["grad_lik"] <= 1e-5 and e["phi"] <= 1e-4 and e["theta1"] <= 1e-5, e
        assert e["p_weights"] <= 1e-4 and e["a_seq"] <= 1e-5 and e["theta_next"] <= 1e-5 and e["mix"] <= 1e-4, e
        # float64 restatement of the first instance
        st = O.SvmpcState(theta_in[0].cpu().double(), mu_in[0].cpu().double(), mix_in[0].cpu().double(), 4.0, aliased=aliased)
        ref = O.svmpc_optimize(model, st, state[0].cpu().double(), eps[0].cpu().double(), torch.full((A,), 2.0).double(),
                               None if params is None else params[0].cpu().double(), False, alpha, 0.5, kernel="rbf")
        e["costs_vs_oracle"] = rel_elem(fused.last["costs"][0].cpu(), ref["costs"])
        e["theta1_vs_oracle"] = rel_max(theta1[0].cpu(), ref["theta1"])
        assert e["costs_vs_oracle"] <= RTOL_COST and e["theta1_vs_oracle"] <= RTOL_PHI, e
        for k_, v in e.items():
            worst[k_] = max(worst.get(k_, 0.0), v)
        staged.theta = fused.theta.clone()
        staged.mu, staged.mix = staged.theta, fused.mix.clone()
    record_parity(f"cluster kernel {kind} B={B} N={N} S={S} H={H} P={P}", **worst)


# ---------------------------------------------------------------------------------------------
# the parameter filter without host round trips: device Silverman rule, the belief object
# ---------------------------------------------------------------------------------------------
def test_device_silverman_rule_matches_kdepy_restatement():
    """`dust_silverman_bandwidth` against the host restatement of KDEpy's rule (checked against the reference's own
    values in the CPU suite): recorded particle clouds, sizes that are not powers of two, degenerate data."""
    from dust_b200 import ops
    from dust_b200.inference.mpf import silvermans_rule

    d = load("dual_pendulum_silverman")
    clouds = [d[f"t{t}_in_mpf_x0"] for t in range(int(d["n_steps"]))]
    g = torch.Generator().manual_seed(5)
    clouds += [torch.randn(n, generator=g) * s for n, s in ((2, 1.0), (3, 0.1), (100, 2.0), (513, 1e-3), (4096, 30.0))]
    clouds += [torch.cat([torch.zeros(90), torch.randn(10, generator=g)]),     # IQR = 0: falls back to the std
               torch.full((64,), 0.7), torch.tensor([1.5])]                       # all equal -> 1.0; a single value -> 1.0
    worst = 0.0
    for x in clouds:
        ref = silvermans_rule(x.numpy()) * 0.8
        bw, iv = ops.silverman_bandwidth(cu(x), 0.8, dp=2)
        got = float(bw[0])
        worst = max(worst, abs(got - ref) / ref)
        assert abs(got - ref) <= 2e-7 * ref, (x.shape, got, ref)
        assert torch.allclose(iv.cpu(), torch.full((2,), 1.0 / got ** 2), rtol=1e-6)
    record_parity("device Silverman rule vs host restatement", max_rel_err=worst)


def test_mpf_class_with_device_bandwidth_matches_reference():
    """MPF.optimize(bw=None): the pendulum configuration's setting (mpf_bandwidth: None) -- Silverman on the device,
    the kernel reading the bandwidth from device memory, 16 lanes per particle (Np = 50) -- against the reference's
    recording, teacher forced on the particles (the pendulum filter amplifies rounding)."""
    from dust_b200.inference.likelihoods import GaussianLikelihood
    from dust_b200.inference.mpf import MPF
    from dust_b200.models.pendulum import PendulumModel
    d = load("dual_pendulum_silverman")
    model = PendulumModel(uncertain_params=("length", "mass"))
    lik = GaussianLikelihood(initial_obs=d["t0_in_state"], obs_std=float(d["obs_std"]), model=model, log_space=False)
    mpf = MPF(init_particles=d["t0_in_mpf_x0"].clone(), likelihood=lik, optimizer_class=torch.optim.SGD,
              lr=float(d["mpf_lr"]), bw=float(d["mpf_prior_bw0"]), bw_scale=1.0)
    worst = 0.0
    prior_bw = float(d["mpf_prior_bw0"])
    for t in range(int(d["n_steps"])):
        x0 = d[f"t{t}_in_mpf_x0"]
        mpf.x.copy_(cu(x0))
        mpf.update_prior(prior_bw)
        lik.condition(None, d[f"t{t}_in_state"])                  # past observation of this step
        gn, bw = mpf.optimize(d[f"t{t}_out_a_seq"][0], d[f"t{t}_out_next_state"], bw=None, n_steps=20)
        assert torch.is_tensor(bw) and bw.is_cuda, "the bandwidth must stay on the device"
        bw_ref = float(d[f"t{t}_out_mpf_bw"])
        worst = max(worst, abs(float(bw) - bw_ref) / bw_ref)
        assert abs(float(bw) - bw_ref) <= 1e-6 * bw_ref
        x64, _ = O.mpf_optimize(O.Model("pendulum"), x0.double(), d[f"t{t}_in_state"].double(), d[f"t{t}_out_a_seq"][0].double(),
                                d[f"t{t}_out_next_state"].double(), float(d["obs_std"]), prior_bw ** 2, bw_ref, float(d["mpf_lr"]), 20, False)
        # the pendulum filter at the demo's settings amplifies rounding ~1.4x per SVGD step: after 20 of them the
        # reference itself sits 4e-2 .. 0.17 from its float64 restatement.  Bar: no further from float64 than the reference.
        err_truth, err_gold, ref_noise = rel_max(mpf.x.cpu(), x64), rel_max(mpf.x.cpu(), d[f"t{t}_out_mpf_x1"]), rel_max(d[f"t{t}_out_mpf_x1"], x64)
        record_parity(f"MPF class device bandwidth t{t} x1", err_truth=err_truth, err_gold=err_gold, ref_noise=ref_noise, rtol=RTOL_PHI)
        assert err_truth <= max(RTOL_PHI, 1.05 * ref_noise), (t, err_truth, err_gold, ref_noise)
        assert gn.shape == (20,)
        # the prior of the next step carries this step's bandwidth (mpf.py:84)
        assert torch.allclose(mpf._prior_inv_var.cpu(), torch.full((2,), 1.0 / bw_ref ** 2), rtol=1e-5)
        prior_bw = bw_ref
    record_parity("MPF class, Silverman bandwidth on the device", bw_rel_err=worst)


def test_particle_belief_samples_and_log_prob():
    """`MPF.prior` (ParticleBelief): log-density equal to the torch MixtureSameFamily it stands for, samples with the
    mixture's moments, centres aliasing the particles."""
    from dust_b200.inference.belief import ParticleBelief

    torch.manual_seed(3)
    x = torch.randn(50, 2, device=DEV) * 0.2 + torch.tensor([0.9, 1.1], device=DEV)
    b = ParticleBelief(x, torch.tensor([0.01, 0.04]))
    ref = b.as_torch()
    v = torch.randn(37, 2, device=DEV) * 0.3 + 1.0
    assert rel_max(b.log_prob(v).cpu(), ref.log_prob(v).cpu()) <= 1e-5
    s = b.sample([20000])
    assert s.shape == (20000, 2) and s.is_cuda
    assert float((s.mean(0) - ref.mean).abs().max()) <= 0.01
    assert float((s.var(0) - ref.variance).abs().max()) <= 0.01
    assert rel_max(b.mean.cpu(), ref.mean.cpu()) <= 1e-6 and rel_max(b.variance.cpu(), ref.variance.cpu()) <= 1e-5
    x += 1.0                                   # in-place move of the particles: the belief follows (quirk H24)
    assert float((b.sample([4000]).mean(0) - x.mean(0)).abs().max()) <= 0.03
    assert b.sample([8]).shape == (8, 2) and b.event_shape == torch.Size([2])


@pytest.mark.parametrize("name,opt", [("mpf_particle_adam", dict(optimizer_class=torch.optim.Adam, lr=0.01)),
                                      ("mpf_particle_momentum", dict(optimizer_class=torch.optim.SGD, lr=0.01, momentum=0.9))])
def test_mpf_class_with_torch_optimizers_on_device(name, opt):
    """MPF with Adam (the SVGD default, svgd.py:115) and momentum SGD: phi from the kernel one step at a time, the
    optimiser -- built once, state kept across optimize() calls (mpf.py:23) -- on the device particles."""
    from dust_b200.inference.likelihoods import GaussianLikelihood
    from dust_b200.inference.mpf import MPF
    from dust_b200.models.particle import Particle

    d = load(name)
    model = Particle(**ENV, uncertain_params=["mass"], mass=2.0)
    lik = GaussianLikelihood(initial_obs=d["obs0"], obs_std=float(d["obs_std"]), model=model, log_space=True)
    mpf = MPF(init_particles=d["x0"].clone(), likelihood=lik, bw=float(d["prior_bw"]), bw_scale=1.0, **opt)
    assert mpf.optimizer is not None
    worst = {}
    for c in range(2):
        gn, bw = mpf.optimize(d[f"c{c}_action"], d[f"c{c}_obs1"], bw=float(d["bw"]), n_steps=10)
        e_x, e_g = rel_max(mpf.x.cpu(), d[f"c{c}_x1"]), rel_max(gn.cpu(), d[f"c{c}_grad_norms"])
        worst[f"x_call{c}"], worst[f"grad_norms_call{c}"] = e_x, e_g
        assert e_x <= RTOL_PHI and e_g <= 1e-3, (c, e_x, e_g)
    record_parity(f"MPF class {name}", **worst)


def test_skid_steer_and_cartpole_step_on_device():
    """The reference's two other forward models (step only): the skid-steer robot against the reference's recording,
    the cart-pole against the restatement of its method body (the reference's own step raises: parity unpinned)."""
    from dust_b200.models.cartpole import CartPoleModel
    from dust_b200.models.skid_steer_robot import SkidSteerRobot

    d = load("skid_steer_step")
    m = SkidSteerRobot(delta_t=float(d["dt"]))
    nd = m.step(cu(d["states"]), cu(d["actions"]), None).cpu()
    ns = m.step(cu(d["states"]), cu(d["actions"]), {k: cu(d[k]) for k in ("x_icr", "wheel_radius", "axial_distance")}).cpu()
    e1, e2 = rel_elem(nd, d["next_default"], floor=1e-6), rel_elem(ns, d["next_sampled"], floor=1e-6)
    assert e1 <= RTOL_COST and e2 <= RTOL_COST, (e1, e2)       # sinf / cosf differ from libm in the last bits
    torch.manual_seed(31)
    M = 200
    x = torch.randn(M, 4) * torch.tensor([1.0, 2.0, 0.3, 1.5])
    a = torch.randn(M, 1) * 0.8
    cp = CartPoleModel()
    got = cp.step(cu(x), cu(a)).cpu()
    ref = O.cartpole_step(x.double(), a.double())
    e3 = rel_max(got, ref)                       # norm-wise: theta_dd cancels to ~0 for some rows
    pd = {"mass_pole": 0.05 + 0.1 * torch.rand(M, 1), "length": 0.5 + torch.rand(M, 1)}
    got_p = cp.step(cu(x), cu(a), {k: cu(v) for k, v in pd.items()}).cpu()
    ref_p = O.cartpole_step(x.double(), a.double(), m_p=pd["mass_pole"].double(), length=pd["length"].double())
    e4 = rel_max(got_p, ref_p)
    assert e3 <= 2e-6 and e4 <= 2e-6, (e3, e4)
    assert rel_elem(got, ref, floor=1e-2) <= RTOL_COST and rel_elem(got_p, ref_p, floor=1e-2) <= RTOL_COST
    record_parity("skid-steer / cart-pole step", skid_default=e1, skid_sampled=e2, cartpole_default=e3, cartpole_sampled=e4)


def test_captured_control_step_replays_the_eager_sequence():
    """dust_b200.utils.graphs.CapturedStep: one dual control step of the pendulum demo shape (the cluster-kernel
    control step, the plant step, 20 MPF steps; particle_example.py:177-207) captured into a CUDA graph with the
    controller state in persistent buffers; k replays leave the policy particles, the mixture weights and the
    parameter particles exactly where k eager steps leave them."""
    from bench_configs import CONFIGS, build, dual_step, dual_step_inplace
    from dust_b200.utils.graphs import CapturedStep

    cfg = CONFIGS["pendulum_demo"]
    dev = torch.device(DEV)
    eager, graphed = build(cfg, dev, seed=4), build(cfg, dev, seed=4)
    for q in (eager, graphed):
        dual_step(cfg, q)                      # reach the steady state: the prior aliases the particles from here on
    assert torch.equal(eager["core"].theta, graphed["core"].theta)
    step = CapturedStep(lambda: dual_step_inplace(cfg, graphed), warmup=3)   # 3 eager steps; the capture pass runs nothing
    for _ in range(3):
        dual_step(cfg, eager)
    assert torch.equal(eager["core"].theta, graphed["theta_p"])
    for _ in range(4):
        dual_step(cfg, eager)
        step()
    torch.cuda.synchronize()
    assert torch.equal(eager["core"].theta, graphed["theta_p"])
    assert torch.equal(eager["core"].mix, graphed["mix_p"])
    assert torch.equal(eager["mpf_x"], graphed["mpf_x"])
    assert not torch.equal(graphed["theta_p"], build(cfg, dev, seed=4)["core"].theta)


def test_two_pairs_per_lane_kernel_equals_the_default():
    """svmpc_quad_kernel (DUST_B200_QUAD=1; an A/B build of the batched pendulum step with two trajectory pairs per
    lane, disco.py:139-209 + svmpc.py:46-54 in one launch): costs bit-equal to the default kernel, likelihood and its
    gradient to summation order."""
    import os

    from dust_b200 import _lib as L
    from dust_b200 import ops
    from dust_b200.models.pendulum import PendulumModel, inst_cost, term_cost

    dev = torch.device(DEV)
    spec = PendulumModel().device_spec(inst_cost, term_cost, dev)
    B, N, S, H = 80, 8, 64, 20                   # S*N = 512 rows = 4 tiles of 128: no ragged tile
    g = torch.Generator().manual_seed(7)
    theta = cu(torch.randn(B, N, H, 1, generator=g) * 2)
    eps = cu(torch.randn(B, S, N, H, 1, generator=g))
    state = cu((torch.rand(B, 2, generator=g) * 2 - 1) * torch.tensor([3.0, 1.0]))
    sigma = cu(torch.tensor([2.0]))
    kw = dict(theta=theta, sigma=sigma, alpha=1.0, want=("costs", "log_lik", "grad_lik"))
    lib = L.load()

    def run():
        lib.dust_profiler_reset(); lib.dust_profiler_enable(1)
        out = ops.rollout_cost(spec, state, eps, **kw)
        prof = L.profiler_report()
        lib.dust_profiler_enable(0)
        assert list(prof) == ["svmpc_instance_kernel"], prof
        return {k: v.clone() for k, v in out.items() if v is not None}

    ref = run()
    os.environ["DUST_B200_QUAD"] = "1"
    try:
        quad = run()
    finally:
        del os.environ["DUST_B200_QUAD"]
    assert torch.equal(quad["costs"], ref["costs"])
    assert rel_max(quad["log_lik"].cpu(), ref["log_lik"].cpu()) <= 1e-6
    assert rel_max(quad["grad_lik"].cpu(), ref["grad_lik"].cpu()) <= 1e-5


def test_batched_controller_noise_prefetch_is_the_same_sequence():
    """BatchedSVMPC(prefetch_noise=True): the action noise of step k + 1 is drawn on a side stream while step k runs
    (likelihoods.py:97-103 draws it inside every call); same draws in the same order, hence the same actions."""
    from dust_b200.batched import BatchedSVMPC
    from dust_b200.models.pendulum import PendulumModel, inst_cost, term_cost

    dev = torch.device(DEV)
    mk = lambda pf: BatchedSVMPC(PendulumModel(), 96, 8, 64, 20, 2.0, 2.0, alpha=1.0, learning_rate=2.0, kernel="gpytorch",  # noqa: E731
                                 inst_cost_fn=inst_cost, term_cost_fn=term_cost, device=dev, seed=11, prefetch_noise=pf)
    a, b = mk(False), mk(True)
    g = torch.Generator().manual_seed(2)
    for step in range(5):
        state = cu((torch.rand(96, 2, generator=g) * 2 - 1) * torch.tensor([3.0, 1.0]))
        ua, ub = a.control_step(state).clone(), b.control_step(state).clone()
        torch.cuda.synchronize()
        assert torch.equal(ua, ub), step
    assert torch.equal(a.theta, b.theta)
    assert a.draws == 5 and b.draws == 6          # one draw ahead
