"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, host
logic (maps, registries, sharding, Silverman), and that the product never touches the oracle."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import dust_oracle as O
from tests.util import golden_grid, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from dust_b200 import _lib

    lib = _lib.load()  # dlopen works without a GPU (no compute call is made)
    header = open(os.path.join(ROOT, "include", "dust_b200.h")).read()
    declared = set(re.findall(r"\b(dust_[a-z0-9_]+)\s*\(", header))
    declared -= {"dust_status"}
    assert declared, "no prototypes found in the header"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/dust_b200.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes prototype in dust_b200/_lib.py"
    assert lib.dust_abi_version() == 3
    assert b"sm_100a" in lib.dust_build_info()


def test_struct_layouts_match_the_header():
    """sizeof of every argument struct, computed by gcc from the header, equals ctypes'."""
    import ctypes

    from dust_b200 import _lib

    names = {"dust_model_desc": _lib.ModelDesc, "dust_rollout_args": _lib.RolloutArgs, "dust_svmpc_step_args": _lib.SvmpcStepArgs, "dust_adjoint_args": _lib.AdjointArgs,
             "dust_gmm_args": _lib.GmmArgs, "dust_median_args": _lib.MedianArgs, "dust_phi_args": _lib.PhiArgs,
             "dust_svmpc_forward_args": _lib.SvmpcForwardArgs, "dust_disco_step_args": _lib.DiscoStepArgs,
             "dust_mpf_args": _lib.MpfArgs, "dust_peer_args": _lib.PeerArgs}
    src = '#include <stdio.h>\n#include "dust_b200.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    exe = "/tmp/dust_sizeof"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    out = subprocess.run([exe], capture_output=True, check=True).stdout.decode()
    for line in out.strip().splitlines():
        n, sz = line.split()
        assert ctypes.sizeof(names[n]) == int(sz), n
    # ... and every field sits at the offset the C compiler gives it (same names on both sides)
    fields = [(n, f[0]) for n, cls in names.items() for f in cls._fields_]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "dust_b200.h"\nint main(){' + "".join(
        f'printf("{n} {f} %zu\\n", offsetof({n}, {f}));' for n, f in fields) + "return 0;}"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    out = subprocess.run([exe], capture_output=True, check=True).stdout.decode()
    for line in out.strip().splitlines():
        n, f, off = line.split()
        assert getattr(names[n], f).offset == int(off), (n, f)


def test_workspace_queries_tolerate_incomplete_arguments():
    """The *_workspace_bytes functions are called before validation: zeroed / model-less structs must not crash."""
    import ctypes as C

    from dust_b200 import _lib

    lib = _lib.load()
    a = _lib.RolloutArgs()
    assert lib.dust_rollout_workspace_bytes(C.byref(a)) == 0
    a.B, a.N, a.S, a.P, a.H = 1, 4, 64, 8, 50          # sizes set, model still NULL
    assert lib.dust_rollout_workspace_bytes(C.byref(a)) >= 0
    adj = _lib.AdjointArgs()
    assert lib.dust_adjoint_workspace_bytes(C.byref(adj)) == 0
    m = _lib.MpfArgs()
    assert lib.dust_mpf_workspace_bytes(C.byref(m)) == 0


def test_no_cpu_fallback_without_a_device():
    from dust_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.require_cuda()
    from dust_b200.models.pendulum import PendulumModel

    with pytest.raises(RuntimeError):
        PendulumModel().step(torch.zeros(1, 2), torch.zeros(1, 1))


def test_product_never_imports_the_oracle_or_the_reference():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dust_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M) or "/root/reference" in txt or "refshim" in txt:
                    bad.append(f)
    assert not bad, bad


@pytest.mark.parametrize("preset,width,name", [("grid_4x4", 2.1, "map_grid4x4"), ("staggered_3-2-3", 2.0, "map_staggered_3-2-3"),
                                               ("grid_3x3", 1.5, "map_grid_3x3"), ("single_centred", 3.0, "map_single_centred"),
                                               ("staggered_4-3-4-3-4", 1.0, "map_staggered_4-3-4-3-4"), ("grid_6x6", 1.3, "map_grid_6x6")])
def test_obstacle_map_matches_reference_generator(preset, width, name):
    from dust_b200.utils.obstacle_map import generate_obstacle_map, get_obst_preset

    om = generate_obstacle_map([22, 22], get_obst_preset(preset, width), 0.1, map_type="direct")
    assert np.array_equal(om.map.astype(np.uint8), golden_grid(name))
    # bit packing the kernels read: bit (ix*ny+iy), LSB first
    words = om.device_bits("cpu").numpy().view(np.uint32)
    cells = np.arange(om.map.size)
    assert np.array_equal((words[cells >> 5] >> (cells & 31)) & 1, om.map.reshape(-1).astype(np.uint32))
    d = load("map_collisions")
    if preset == "grid_4x4":
        assert torch.equal(om.get_collisions(d["X"]), d["coll"])
    with pytest.raises(IOError):
        get_obst_preset("nope", 1.0)


def test_cost_registry_and_kernel_modes():
    from dust_b200.inference.svmpc import _kernel_mode
    from dust_b200.kernels.base_kernels import RBF, RBFKernel
    from dust_b200.kernels.composite_kernels import iid_mp
    from dust_b200.models.pendulum import SwingUpCost, _match_swingup

    def demo(states, controls=None, n_pol=1, debug=None):
        th, thd = states.chunk(2, dim=1)
        return 50.0 * (th.cos() - 1) ** 2 + 1.0 * thd ** 2

    assert _match_swingup(demo) == pytest.approx((50.0, 1.0), rel=1e-5)
    assert _match_swingup(SwingUpCost(3.0, 0.5)) == (3.0, 0.5)
    assert _match_swingup(lambda s, *a, **k: (s ** 2).sum(-1)) is None
    mode, ell, _ = _kernel_mode(RBFKernel())
    assert mode == "gpytorch" and abs(ell - np.log(2.0)) < 1e-6
    assert _kernel_mode(iid_mp(base_kernel=RBF(bandwidth=-1), ctrl_dim=2))[0] == "mp"
    with pytest.raises(NotImplementedError):
        _kernel_mode(RBF())


def test_silverman_and_box():
    from dust_b200.inference.mpf import silvermans_rule
    from dust_b200.utils.spaces import Box

    d = load("dual_pendulum_silverman")
    for t in range(int(d["n_steps"])):
        bw = silvermans_rule(d[f"t{t}_in_mpf_x0"].numpy())
        assert abs(bw - float(d[f"t{t}_out_mpf_bw"])) < 1e-6 * bw
    b = Box(dim=2, low=-1.0, high=torch.tensor([1.0, 2.0]))
    assert b.dim == 2 and b.shape == torch.Size([2]) and torch.equal(b.low, torch.tensor([-1.0, -1.0]))
    with pytest.raises(AssertionError):
        Box(dim=0)


def _phi_segments(pl, cta):
    """The kernel's TcSegIter (svgd_tc.cu) restated: the (row tile, first column tile, length) segments of CTA `cta`."""
    grid, R, n_chunks, chunk_w, rem0, rem_w, W, total, max_seg, _, _ = pl
    T = rem0 + rem_w
    c = cta - n_chunks * R
    if c < 0:
        k, rt = divmod(cta, R)
        return [(rt, k * chunk_w, min(chunk_w, T - k * chunk_w))]
    out, u, u1 = [], c * W, min((c + 1) * W, total)
    while u < u1:
        rt = u // rem_w
        jj = u - rt * rem_w
        ln = min(rem_w - jj, u1 - u)
        out.append((rt, rem0 + jj, ln))
        u += ln
    return out


@pytest.mark.parametrize("col_tiles", [16, 64, 128, 1024])
def test_phi_tensor_core_partition_covers_every_tile_pair_once(col_tiles):
    """dust_phi_tc_plan (the partition phi_tc_kernel works off, dust/inference/svgd.py:127-135 by row blocks): for row
    blocks of every size class -- fewer row tiles than SMs, an exact multiple, a multiple plus a rest -- each (row tile,
    column tile) pair belongs to exactly one segment, no CTA exceeds its scratch slots, and the finish kernel's slot
    lookup (chunk CTAs k * R + rt, then the remainder CTAs that meet the row tile) names exactly the segments that
    worked on a row tile."""
    import ctypes as C

    from dust_b200 import _lib as L
    lib = L.load()
    T = col_tiles
    for row_tiles in [1, 2, 3, 7, 8, 16, 31, 64, 68, 73, 74, 75, 100, 147, 148, 149, 200, 296, 300, 512]:
        if row_tiles * 128 > T * 64 * 4 and T < 1024:      # keep the emulation small
            continue
        buf = (C.c_int32 * 22)()
        n = lib.dust_phi_tc_plan(row_tiles, T, C.byref(buf))
        assert 1 <= n <= 2
        plans = [list(buf[i * 11:(i + 1) * 11]) for i in range(n)]
        assert sum(pl[10] for pl in plans) == row_tiles and plans[0][9] == 0
        for pl in plans:
            grid, R, n_chunks, chunk_w, rem0, rem_w, W, total, max_seg, rt0, rts = pl
            assert 1 <= grid <= 148 and R == rts and rem0 + rem_w == T
            seen = np.zeros((R, T), dtype=np.int32)
            by_rt = {}
            longest = 0
            for cta in range(grid):
                segs = _phi_segments(pl, cta)
                assert 1 <= len(segs) <= max_seg, (row_tiles, T, cta, segs, pl)
                longest = max(longest, sum(s[2] for s in segs))
                for slot, (rt, j0, ln) in enumerate(segs):
                    assert ln >= 1 and 0 <= rt < R and j0 + ln <= T
                    seen[rt, j0:j0 + ln] += 1
                    by_rt.setdefault(rt, []).append((cta, slot))
            assert (seen == 1).all(), (row_tiles, T, pl)
            # balance: the longest CTA is within a few percent (+ the 16-unit floor) of an even split
            assert longest <= max(16, int(np.ceil(R * T / 148) * 1.16) + 1), (row_tiles, T, longest, pl)
            # the finish kernel's view of row tile rt
            for rt in range(R):
                srcs = [(k * R + rt, 0) for k in range(n_chunks)]
                if rem_w > 0:
                    for c in range((rt * rem_w) // W, ((rt + 1) * rem_w - 1) // W + 1):
                        srcs.append((n_chunks * R + c, rt - (c * W) // rem_w))
                assert sorted(srcs) == sorted(by_rt[rt]), (row_tiles, T, rt, pl)


def test_row_blocks_partition():
    from dust_b200.distributed import row_block

    for n, w in ((65536, 8), (10, 3), (7, 8), (4096, 1)):
        blocks = [row_block(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert max(e - b for b, e in blocks) - min(e - b for b, e in blocks) <= 1


# ---- world_size-2 gloo run of the sharded SVGD host logic (compute callbacks = the oracle) --------
_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DUST_ROOT"])
from oracle import dust_oracle as O
from dust_b200.distributed import ShardedSVGD, row_block

class OracleOps:
    """CPU stand-ins with the ABI's semantics: per-rank row-block histogram + all-reduce."""
    @staticmethod
    def median_sq_dist(x, rows=None, all_reduce=None, **kw):
        d2 = O.sq_dists_addmm(x[rows[0]:rows[1]], x)
        bits = d2.reshape(-1).view(torch.int32).to(torch.int64)
        k = (x.shape[0] ** 2 - 1) // 2
        hist = torch.bincount(bits >> 16, minlength=65536)
        all_reduce(hist)
        cum = hist.cumsum(0); hi = int((cum > k).nonzero()[0]); below = int(cum[hi - 1]) if hi else 0
        hist2 = torch.bincount(bits[(bits >> 16) == hi] & 0xFFFF, minlength=65536)
        all_reduce(hist2)
        lo = int((hist2.cumsum(0) > (k - below)).nonzero()[0])
        return torch.tensor([(hi << 16) | lo], dtype=torch.int32).view(torch.float32)
    @staticmethod
    def bandwidth_from_median(med, N, scale, mode):
        bw = scale * max(float(torch.sqrt(0.5 * med[0])) / float(torch.tensor(N + 1.0).log()), 1e-5)
        return torch.tensor([1 / (2 * bw * bw), 1 / N, 1 / (N * bw * bw), bw])
    @staticmethod
    def svgd_phi(x, s, gamma_dev=None, rows=None, **kw):
        g, c1, c2 = [float(v) for v in gamma_dev[:3]]
        full = O.phi_unified(x[0].double(), s[0].double(), g, c1, c2).float()
        out = torch.zeros_like(x); out[0, rows[0]:rows[1]] = full[rows[0]:rows[1]]
        return dict(phi=out)

dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
torch.manual_seed(0)
N, D = 96, 7
X = torch.randn(N, D); S = -X
b, e = row_block(N, r, w)
sh = ShardedSVGD(N, D, ops=OracleOps)
phi_loc, coef = sh.phi(X[b:e].clone(), S[b:e].clone())
sh2 = ShardedSVGD(N, D, ops=OracleOps, gather="separate")      # X and score gathered separately: same result
phi_loc2, coef2 = sh2.phi(X[b:e].clone(), S[b:e].clone())
assert torch.equal(phi_loc2, phi_loc) and torch.equal(coef2, coef)
bw_ref, med_ref = O.bw_median(X)
ref = O.phi_svgd(X.double(), S.double(), float(bw_ref)).float()
assert abs(float(coef[3]) - float(bw_ref)) < 1e-6, (float(coef[3]), float(bw_ref))
assert float((phi_loc - ref[b:e]).abs().max()) < 1e-5
gathered = [torch.empty_like(phi_loc) for _ in range(w)]
dist.all_gather(gathered, phi_loc)
assert float((torch.cat(gathered) - ref).abs().max()) < 1e-5
dist.destroy_process_group()
print("rank", r, "ok")
'''


def test_sharded_svgd_host_logic_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, DUST_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                         env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == 2


# ---- world_size-2 gloo run of the parameter-draw sharding of one instance (compute callbacks = the oracle) ----
_WORKER_ROLLOUT = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DUST_ROOT"])
from oracle import dust_oracle as O
from dust_b200.distributed import ShardedRollout

MODEL = O.Model("pendulum")

class OracleOps:
    """CPU stand-ins with the ABI's semantics: a draw range yields its share of the mean cost / gradient."""
    @staticmethod
    def rollout_cost(spec, state0, noise, theta=None, sigma=None, params=None, alpha=1.0, want=(), out=None, p_range=None,
                     reduce_only=False, **kw):
        actions = theta[0] + sigma * noise[0]
        if not reduce_only:
            P = params.shape[1]
            st = O.rollout(MODEL, state0[0], actions, params[0, p_range[0]:p_range[1]])
            c = O.trajectory_costs(MODEL, st, actions) * (p_range[1] - p_range[0]) / P
            return {"costs": c.unsqueeze(0)}
        costs = out["costs"][0]
        res = {"costs": out["costs"], "log_lik": O.exp_utility_log_prob(costs, alpha).unsqueeze(0)}
        w = torch.softmax(-alpha * costs, 0)
        res["lik_weights"] = w.unsqueeze(0)
        res["grad_lik"] = O.analytic_lik_grad(costs, actions, theta[0], sigma, alpha).unsqueeze(0)
        return res
    @staticmethod
    def rollout_adjoint(spec, state0, noise, lik_w, theta=None, sigma=None, params=None, alpha=1.0, p_range=None, **kw):
        P = params.shape[1]
        th = theta[0].clone().requires_grad_(True)
        actions = th + sigma * noise[0]
        st = O.rollout(MODEL, state0[0], actions, params[0, p_range[0]:p_range[1]])
        c = O.trajectory_costs(MODEL, st, actions) * (p_range[1] - p_range[0]) / P
        (g,) = torch.autograd.grad((-alpha * lik_w[0] * c).sum(), th)     # d log_l / d theta through the costs
        return g.unsqueeze(0)

dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
torch.manual_seed(0)
S, N, H, P = 16, 3, 8, 5
state0, theta, sigma = torch.tensor([[2.5, 0.3]]), torch.randn(1, N, H, 1), torch.tensor([2.0])
noise, params = torch.randn(1, S, N, H, 1), torch.rand(1, P, 2) * 0.7 + 0.6
sh = ShardedRollout(P, ops=OracleOps)
assert sh.p_range == ((0, 3) if r == 0 else (3, 5))
full = ShardedRollout.__new__(ShardedRollout); full.P, full.group, full.ops, full.rank, full.world, full.p_range = P, None, OracleOps, 0, 1, (0, P)
for grad in ("analytic", "pathwise"):
    a = sh.evaluate(None, state0, noise, theta, sigma, params, alpha=0.01, grad=grad)
    b = full.evaluate(None, state0, noise, theta, sigma, params, alpha=0.01, grad=grad)
    for k in ("costs", "log_lik", "grad_lik"):
        err = float((a[k] - b[k]).abs().max() / b[k].abs().max())
        assert err < 1e-5, (grad, k, err)
dist.destroy_process_group()
print("rank", r, "ok")
'''


def test_sharded_rollout_host_logic_gloo_world2(tmp_path):
    script = tmp_path / "worker_rollout.py"
    script.write_text(_WORKER_ROLLOUT)
    env = dict(os.environ, DUST_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29633", str(script)],
                         env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == 2


def test_demo_default_configs_equal_the_reference_yaml():
    """demo/configs.py carries the reference's two configurations key for key; where the reference tree is
    mounted (the build container) its yaml files must load to the same dictionaries."""
    import os

    import pytest
    import yaml

    from demo import configs

    ref = "/root/reference/demo"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not mounted")
    for name, default in (("pendulum", configs.PENDULUM), ("particle", configs.PARTICLE)):
        path = os.path.join(ref, name + "_config.yaml")
        with open(path) as f:
            assert yaml.load(f, yaml.FullLoader) == default, name
        assert configs.load(path, None) == default
    assert configs.load(None, configs.PENDULUM) is not configs.PENDULUM


def test_merwe_transformer_matches_reference():
    import pytest
    import torch

    from tests.util import load, rel_max
    from dust_b200.utils.utf import MerweScaledUTF

    d = load("utf_points")
    tf = MerweScaledUTF(n=2, alpha=0.5)
    assert tf.pts == 5
    assert torch.equal(tf.loc_weights, d["loc_weights"]) and torch.equal(tf.cov_weights, d["cov_weights"])
    sig = tf.compute_sigma_points(d["mean"], d["cov"])
    assert torch.equal(sig, d["sigmas"])
    mu, K = tf.unscented_transform(sig)
    # the reference's transform of its own points (their spread uses the ROWS of the upper Cholesky factor, so
    # the covariance that comes back is not the one that went in -- reproduced, not corrected)
    assert rel_max(mu, d["ut_mean"]) <= 1e-6 and rel_max(K, d["ut_cov"]) <= 1e-6
    with pytest.raises(ValueError):
        tf.compute_sigma_points(torch.zeros(3), torch.eye(3))
