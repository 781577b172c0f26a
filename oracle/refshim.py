"""Import shim for the UNMODIFIED reference package (test infrastructure only).

The reference (`/root/reference/dust`, pure Python) imports gpytorch, KDEpy, gym and
matplotlib, none of which exist in this image.  This module registers minimal
stand-ins for exactly the symbols the reference touches and
then loads the reference under the module name ``dust_ref`` so that it can never
be confused with the product package ``dust_b200``.

It is used ONLY by ``tests/golden/make_golden.py`` (golden-vector generation, run in
the build container where /root/reference is mounted) and by the optional
"live reference" cross-checks in ``tests/`` (skipped when /root/reference is absent,
e.g. on the GPU box).  Nothing in ``dust_b200/`` imports it.

The gpytorch / KDEpy bodies are restatements of those packages' *documented*
behaviour (gpytorch 1.5.0 ``RBFKernel`` with default ``raw_lengthscale = 0``;
KDEpy 1.1.0 ``silvermans_rule``); goldens that pass through them are labelled
"shim-dependent" in ``tests/golden/MANIFEST.json``.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("DUST_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dust", "__init__.py"))


def silvermans_rule(data):
    """KDEpy 1.1.0 ``bw_selection.silvermans_rule`` for 1-D data of shape [n, 1]."""
    obs, dims = data.shape
    assert dims == 1
    if obs == 1:
        return 1
    iqr = (np.percentile(data, 75) - np.percentile(data, 25)) / 1.349
    std = np.std(data, ddof=1)
    sigma = min(std, iqr) if iqr > 0 else std
    if sigma > 0:
        return sigma * (obs * 3 / 4.0) ** (-1 / 5)
    iqr = (np.percentile(data, 99) - np.percentile(data, 1)) / 4.6526957480816815
    return iqr * (obs * 3 / 4.0) ** (-1 / 5) if iqr > 0 else 1.0


class _Lazy:
    def __init__(self, t):
        self._t = t

    def evaluate(self):
        return self._t


class RBFKernel(torch.nn.Module):
    """gpytorch 1.5 ``RBFKernel()`` with its default hyper-parameters.

    lengthscale = softplus(raw_lengthscale = 0) = ln 2; the covariance is
    exp(-0.5 * ||(x1 - x2) / l||^2) with gpytorch's mean-centring and clamp.
    """

    def __init__(self):
        super().__init__()
        self.raw_lengthscale = torch.nn.Parameter(torch.zeros(1, 1))

    @property
    def lengthscale(self):
        return torch.nn.functional.softplus(self.raw_lengthscale)

    def forward(self, x1, x2):
        a, b = x1 / self.lengthscale, x2 / self.lengthscale
        adj = a.mean(-2, keepdim=True)
        a, b = a - adj, b - adj
        d = (
            a.pow(2).sum(-1, keepdim=True)
            + b.pow(2).sum(-1, keepdim=True).transpose(-2, -1)
            - 2 * a @ b.transpose(-2, -1)
        ).clamp_min(0)
        return _Lazy(d.div(-2).exp())


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_third_party_stubs():
    if not hasattr(np, "float"):
        np.float = float  # reference helper.py:90 (numpy >= 1.24 removed it)
    if "matplotlib" not in sys.modules:
        plt = _mod("matplotlib.pyplot")
        cm = _mod("matplotlib.cm")
        _mod("matplotlib", pyplot=plt, cm=cm)
    if "gym" not in sys.modules:
        def _no_gym(*a, **k):
            raise RuntimeError("gym is not installed (stub)")
        _mod("gym", make=_no_gym)
    if "KDEpy" not in sys.modules:
        bw = _mod("KDEpy.bw_selection", silvermans_rule=silvermans_rule)
        _mod("KDEpy", bw_selection=bw)
    if "gpytorch" not in sys.modules:
        gk = _mod("gpytorch.kernels", RBFKernel=RBFKernel)
        _mod("gpytorch.kernels.rbf_kernel", RBFKernel=RBFKernel)
        _mod("gpytorch", kernels=gk)


def load_reference(name: str = "dust_ref"):
    """Load /root/reference/dust as package ``name`` (relative imports keep working)."""
    if name in sys.modules:
        return sys.modules[name]
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_third_party_stubs()
    pkg_dir = os.path.join(REFERENCE_ROOT, "dust")
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir]
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    # anomaly mode is switched on globally by the reference at import
    # (svgd.py:7); the callers decide whether to keep it.
    return mod


def ref_import(path: str, name: str = "dust_ref"):
    """``ref_import('inference.svmpc')`` -> reference module object."""
    load_reference(name)
    return importlib.import_module(f"{name}.{path}")
