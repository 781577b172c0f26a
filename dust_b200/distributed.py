"""Multi-GPU sharding of the hot path on one NVSwitch box (one process per GPU, torch.distributed).

* Independent MPC instances (`shard_instances`): contiguous split of the batch, NO communication.
* ONE large instance (`ShardedRollout`): the P dynamics-parameter draws are split over the ranks; the
  trajectory cost is a mean over the draws (disco.py:330), so every rank rolls out its share and one
  all-reduce (sum) of the [S, N] cost shares (and, for the pathwise gradient, of the [N, H, A] gradient
  shares) completes it.  The soft-min / likelihood reductions then run replicated on every rank.
* Large-N SVGD (`ShardedSVGD`): rank r owns a row block of X, score and phi.  One exchange per
  evaluation: all-gather of [X | score] (N*2D*4 bytes in total), plus -- for the exact median
  bandwidth -- an all-reduce (sum) of the 65536-bin radix histogram after each of the two passes.
  Every rank then computes its own phi rows against all N columns; nothing else is reduced.

The collectives go through torch.distributed (NCCL on GPUs; gloo in the CPU tests, where the
compute callbacks are replaced by the oracle)."""
import torch
import torch.distributed as dist

from . import ops as _ops


def row_block(n, rank, world):
    """Contiguous, balanced [begin, end) of `n` rows for `rank` of `world`."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_instances(n_instances, rank=None, world=None):
    """Instance range of this rank for the batched controller (no data-path collective)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return row_block(n_instances, rank, world)


class PeerExchange:
    """All-gather of equal row blocks over NVLink peer memory (csrc/peer.cu, push form): every rank owns two gathered
    buffers [world * rows_per_rank, row_floats] that the other ranks map through CUDA IPC.  `gather(x, score)` stores
    the local rows into block `rank` of the current-parity buffer of EVERY rank (posted 16-byte stores), raises this
    rank's flag there, and makes the current stream wait for all flags of the exchange -- two small launches, no host
    synchronisation, no library collective.  One process per GPU on one NVSwitch box; the 64-byte handles travel
    once, through `group`.  Buffers alternate by parity: the tensor `gather` returns stays valid until the exchange
    after next."""

    def __init__(self, rows_per_rank, row_floats, group=None, device=None):
        import ctypes as C

        from . import _lib as L
        self.L, self.C = L, C
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.rows, self.width = int(rows_per_rank), int(row_floats)
        if self.width % 4:
            raise ValueError("PeerExchange: rows must be whole 16-byte vectors")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        lib = L.load()
        self.buf_stride = (4 * self.world * self.rows * self.width + 255) // 256 * 256
        total = 2 * self.buf_stride + 512                       # two gathered buffers, flags [<= 16], counters [<= 16]
        base = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(lib.dust_peer_alloc(total, C.byref(base)))
            handle = (C.c_ubyte * 64)()
            L.check(lib.dust_peer_export(base, C.byref(handle)))
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
            every = torch.empty(self.world * 64, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(every, mine, group=group)
            every = every.cpu().reshape(self.world, 64)
            self.bases = []
            for r in range(self.world):
                if r == self.rank:
                    self.bases.append(base.value)
                    continue
                h = (C.c_ubyte * 64)(*[int(v) for v in every[r]])
                p = C.c_void_p()
                L.check(lib.dust_peer_open(C.byref(h), C.byref(p)))
                self.bases.append(p.value)
        self._own = base.value
        self.epoch = 0
        self._buf_arrays = [(C.c_void_p * self.world)(*[b + par * self.buf_stride for b in self.bases]) for par in (0, 1)]
        self._flag_array = (C.c_void_p * self.world)(*[b + 2 * self.buf_stride for b in self.bases])
        self._counters = self._own + 2 * self.buf_stride + 256
        self._local = [_as_tensor(self._own + par * self.buf_stride, (self.world * self.rows, self.width), self.device, self)
                       for par in (0, 1)]
        dist.barrier(group=group)                                # every mapping exists before the first store

    def gather(self, *parts):
        """parts: contiguous float32 tensors [rows_per_rank, w_i], w_i % 4 == 0, sum(w_i) = row_floats (X and score)
        -> [N, row_floats] (rank-major rows)."""
        L, C = self.L, self.C
        if not 1 <= len(parts) <= 4:
            raise ValueError("PeerExchange.gather: 1..4 row pieces")
        self.epoch += 1
        par = self.epoch & 1
        a = L.PeerArgs()
        a.world, a.rank, a.epoch, a.rows_per_rank, a.row_floats = self.world, self.rank, self.epoch, self.rows, self.width
        a.gathered_peers = C.cast(self._buf_arrays[par], C.POINTER(C.c_void_p))
        a.flags = C.cast(self._flag_array, C.POINTER(C.c_void_p))
        a.counters = self._counters
        a.n_parts = len(parts)
        for k, t in enumerate(parts):
            if t.shape[0] != self.rows or not t.is_contiguous():
                raise ValueError("PeerExchange.gather: pieces must be contiguous [rows_per_rank, w]")
            a.parts[k] = L.ptr(t)
            a.part_floats[k] = int(t.shape[1])
        L.call("dust_peer_push", C.byref(a), L.stream())
        L.call("dust_peer_wait", C.byref(a), L.stream())
        return self._local[par]

    def close(self):
        if getattr(self, "_own", None) is None:
            return
        lib = self.L.load()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)                           # nobody still writes into this rank's buffers
        for r, b in enumerate(self.bases):
            if r != self.rank:
                lib.dust_peer_close(b)
        lib.dust_peer_free(self._own)
        self._own = None


class _DevMem:
    """__cuda_array_interface__ view of library-owned device memory (kept alive by `owner`)."""

    def __init__(self, ptr, shape, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2,
                                         "strides": None}


def _as_tensor(ptr, shape, device, owner):
    return torch.as_tensor(_DevMem(ptr, shape, owner), device=device)


class ShardedSVGD:
    """phi = (K grad log p + sum grad k)/N for N particles split by row blocks over the ranks of
    `group` (dust/inference/svgd.py:127-135 with the bw_median bandwidth, svgd.py:42-52)."""

    def __init__(self, n_total, dim, group=None, ops=_ops, device=None, gather="packed"):
        """gather: "packed" = ONE all-gather of [X | score] into an [N, 2D] buffer that the kernels read in place through
        a row stride (X = buffer[:, :D], score = buffer[:, D:]: no slicing copies); "separate" = X and score gathered
        into their own [N, D] buffers by two collectives (the score gather then overlaps the bandwidth pass);
        "peer" = no library collective: every rank stores its rows over NVLink peer memory into the packed buffer of
        every rank (`PeerExchange`, csrc/peer.cu; CUDA devices of one box only).
        Measured on 8 B200 (profiles/r2_scale8.md): the NCCL all-gather of 21 MB costs ~90 us of a 0.6 ms step."""
        if gather not in ("packed", "separate", "peer"):
            raise ValueError(gather)
        self.gather_mode = gather
        self._gathered2 = {}
        self.N, self.D, self.group, self.ops = n_total, dim, group, ops
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if n_total % self.world:
            raise ValueError(f"N={n_total} must be a multiple of the world size {self.world}")
        self.rows = row_block(n_total, self.rank, self.world)
        self.device = device
        self._gathered = None
        self._median_ws = None
        self._peer = None

    def _all_gather(self, local):
        """[n_loc, C] -> [N, C] (rank-major row order)."""
        if self.world == 1:
            return local
        if self._gathered is None or self._gathered.shape[1] != local.shape[1]:
            self._gathered = torch.empty((self.N, local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(self._gathered, local.contiguous(), group=self.group)
        return self._gathered

    def _all_reduce_hist(self, hist):
        if self.world > 1:
            dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=self.group)

    def gather_async(self, x_local, score_local):
        """-> (x_all, score_all, wait_x, wait_score): the two collectives are enqueued back to back; `wait_x()` /
        `wait_score()` make the current stream wait for one of them, so the score gather overlaps whatever only
        needs X (the bandwidth pass)."""
        nothing = lambda: None  # noqa: E731
        if self.world == 1:
            return x_local, score_local, nothing, nothing
        if self.gather_mode == "separate":
            out, waits = [], []
            for key, loc in (("x", x_local), ("score", score_local)):
                buf = self._gathered2.get(key)
                if buf is None or buf.dtype != loc.dtype or buf.device != loc.device:
                    buf = self._gathered2[key] = torch.empty((self.N, self.D), dtype=loc.dtype, device=loc.device)
                work = dist.all_gather_into_tensor(buf, loc.contiguous(), group=self.group, async_op=True)
                out.append(buf)
                waits.append(work.wait)
            return out[0], out[1], waits[0], waits[1]
        if self.gather_mode == "peer":
            if self._peer is None:
                self._peer = PeerExchange(self.N // self.world, 2 * self.D, group=self.group, device=x_local.device)
            both = self._peer.gather(x_local.contiguous(), score_local.contiguous())
            return both[:, : self.D], both[:, self.D:], nothing, nothing
        local = torch.cat([x_local, score_local], dim=1)
        if self._gathered is None or self._gathered.shape[1] != 2 * self.D or self._gathered.device != local.device:
            self._gathered = torch.empty((self.N, 2 * self.D), dtype=local.dtype, device=local.device)
        work = dist.all_gather_into_tensor(self._gathered, local, group=self.group, async_op=True)
        return self._gathered[:, : self.D], self._gathered[:, self.D:], work.wait, nothing

    def gather(self, x_local, score_local):
        x_all, s_all, wait_x, wait_s = self.gather_async(x_local, score_local)
        wait_x()
        wait_s()
        return x_all, s_all

    def _shared_bytes(self, device):
        import ctypes as C

        from . import _lib as L
        a = L.PhiArgs()
        a.B, a.N, a.D = 1, self.N, self.D
        a.row_begin, a.row_end = self.rows
        lib = L.load()
        return int(max(lib.dust_phi_workspace_bytes(C.byref(a)), lib.dust_median_fast_workspace_bytes(self.N, self.D), 1))

    def median(self, x_all, defer_fallback=False):
        """Exact lower median of all N^2 squared distances: each rank histograms its row block.
        defer_fallback (library ops only): run the tensor-core window pass alone and return (median, check);
        `check()` tells -- without stalling the kernels queued after it -- whether the rank fell inside the window
        (the decision is taken on the all-reduced counts, so every rank sees the same answer)."""
        if getattr(self.ops, "MedianWorkspace", None) is None:     # stand-in ops (host-logic tests)
            return self.ops.median_sq_dist(x_all, rows=self.rows, all_reduce=self._all_reduce_hist), None
        if self._median_ws is None:
            # ONE buffer for the median pass and for phi: both keep |x|^2 and the operand images of X at its head
            self._shared_ws = torch.empty(self._shared_bytes(x_all.device), dtype=torch.uint8, device=x_all.device)
            self._median_ws = self.ops.MedianWorkspace(self.N, self.D, x_all.device, fast_ws=self._shared_ws)
        if defer_fallback:
            # every rank draws 1/world of the 2^20 sampled pairs that place the window; the sample histogram (128 KB) is summed
            return self.ops.median_sq_dist_deferred(x_all, ws=self._median_ws, rows=self.rows, all_reduce=self._all_reduce_hist,
                                                    sample=row_block(1 << 20, self.rank, self.world))
        return self.ops.median_sq_dist(x_all, ws=self._median_ws, rows=self.rows, all_reduce=self._all_reduce_hist), None

    def phi(self, x_local, score_local, bw=None, bw_scale=1.0):
        """Returns (phi_local [n_loc, D], coef) with coef = device {gamma, c1, c2, bw}."""
        x_all, s_all, wait_x, wait_s = self.gather_async(x_local, score_local)
        b, e = self.rows
        if bw is not None:
            wait_x()
            wait_s()
            out = self.ops.svgd_phi(x_all.unsqueeze(0), s_all.unsqueeze(0), gamma=1.0 / (2.0 * bw * bw),
                                    c1=1.0 / self.N, c2=1.0 / (self.N * bw * bw), rows=self.rows)
            return out["phi"][0, b:e], None
        wait_x()
        med, check = self.median(x_all, defer_fallback=True)
        coef = self.ops.bandwidth_from_median(med, self.N, bw_scale, 0)
        wait_s()
        shared = getattr(self, "_shared_ws", None)
        kw = {} if shared is None else dict(workspace=shared, x_prepared=self._median_ws.fast and self.rows[0] % 128 == 0
                                            and self.rows[1] % 128 == 0)
        out = self.ops.svgd_phi(x_all.unsqueeze(0), s_all.unsqueeze(0), gamma_dev=coef, rows=self.rows, **kw)
        if check is not None and not check():
            # the rank fell outside the sampled window (ties, clusters): the two-pass radix select, then phi again
            med = self.ops.median_sq_dist(x_all, ws=self._median_ws, rows=self.rows, all_reduce=self._all_reduce_hist,
                                          allow_fast=False)
            coef = self.ops.bandwidth_from_median(med, self.N, bw_scale, 0)
            out = self.ops.svgd_phi(x_all.unsqueeze(0), s_all.unsqueeze(0), gamma_dev=coef, rows=self.rows)
        return out["phi"][0, b:e], coef


class ShardedRollout:
    """Rollout, cost, likelihood and likelihood gradient of ONE instance with its P parameter draws split
    over the ranks of `group` ("rollouts of one instance"; disco.py:139-209, 294-346;
    likelihoods.py:113-135; svmpc.py:46-60).  Inputs are replicated (state, theta, noise, all P draws:
    the draws are a few KB); each rank rolls out draws [p0, p1)."""

    def __init__(self, n_params, group=None, ops=_ops):
        self.P, self.group, self.ops = int(n_params), group, ops
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.p_range = row_block(self.P, self.rank, self.world)

    def _all_reduce(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def evaluate(self, spec, state0, noise, theta, sigma, params, param_tiling=0, likelihood=0, alpha=1.0, temperature=1.0,
                 grad="analytic"):
        """state0 [1,ds], noise [1,S,N,H,A], theta [1,N,H,A], params [1,P,dp].
        -> dict(costs [1,S,N], log_lik [1,N], grad_lik [1,N,H,A][, lik_weights])."""
        assert params.shape[1] == self.P
        kw = dict(theta=theta, sigma=sigma, params=params, param_tiling=param_tiling, likelihood=likelihood, alpha=alpha)
        p0, p1 = self.p_range
        if p1 > p0:
            out = self.ops.rollout_cost(spec, state0, noise, temperature=temperature, want=("costs",), p_range=(p0, p1), **kw)
        else:   # more ranks than draws: this rank contributes nothing
            out = {"costs": noise.new_zeros(noise.shape[:3])}
        self._all_reduce(out["costs"])
        want = ("log_lik", "grad_lik") if grad == "analytic" else ("log_lik", "lik_weights")
        res = self.ops.rollout_cost(spec, state0, noise, temperature=temperature, want=want, out={"costs": out["costs"]},
                                    reduce_only=True, **kw)
        if grad != "analytic":
            if p1 > p0:
                g = self.ops.rollout_adjoint(spec, state0, noise, res["lik_weights"], p_range=(p0, p1), **kw)
            else:
                g = theta.new_zeros(theta.shape)
            res["grad_lik"] = self._all_reduce(g)
        return res
