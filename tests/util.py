"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# BASELINE.json tolerances
RTOL_COST = 1e-5   # rollouts / costs, fp32, relative
RTOL_PHI = 1e-4    # phi and updated particles, relative (norm-wise)


def load(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(np.asarray(d[k])) for k in d.files}


def golden_grid(name="map_grid4x4"):
    m = load(name)
    shape = tuple(int(v) for v in m["shape"])
    return np.unpackbits(m["packed"].numpy())[: shape[0] * shape[1]].reshape(shape)


def rel_max(a, b):
    """max |a-b| / max |b|  (norm-wise relative error)."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def rel_elem(a, b, floor=0.0):
    """max element-wise relative error."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float(((a - b).abs() / (b.abs() + floor + 1e-300)).max())


def assert_close_to_reference(actual, golden, truth64, rtol, what=""):
    """`actual` must match the reference's `golden` output up to `rtol` PLUS the reference's
    own float32 deviation from the float64 restatement `truth64` (triangle inequality): where
    the reference is itself rounding-noise limited (matmul-form distances, unstable MPF
    iteration) it cannot be matched more closely than it matches exact arithmetic."""
    ref_noise = rel_max(golden, truth64)
    err_truth = rel_max(actual, truth64)
    err_gold = rel_max(actual, golden)
    record_parity(what or "unnamed", err_truth=err_truth, err_gold=err_gold, ref_noise=ref_noise, rtol=rtol)
    assert err_truth <= rtol or err_gold <= rtol, (
        f"{what}: |actual-truth64|={err_truth:.3e}, |actual-golden|={err_gold:.3e}, "
        f"reference noise |golden-truth64|={ref_noise:.3e}, rtol={rtol:.1e}")
    assert err_gold <= rtol + 2.0 * ref_noise, (
        f"{what}: |actual-golden|={err_gold:.3e} > rtol + 2*refnoise ({ref_noise:.3e})")


def record_parity(case, **values):
    """Append the observed errors of a parity check to gpurun_out/parity_table.jsonl (scratch; summarised into
    profiles/r2_parity_table.md by profiles/parity_table.py).  Never fails a test."""
    import json

    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "parity_table.jsonl"), "a") as f:
            f.write(json.dumps({"case": case, **{k: (float(v) if isinstance(v, (int, float)) else v) for k, v in values.items()}}) + "\n")
    except Exception:
        pass
