"""dust_b200: the DuSt-MPC inner loop (rollout/cost -> grad log p -> RBF-SVGD update) as
hand-written sm_100a CUDA behind the reference's controller / inference API.

    from dust_b200.controllers.disco import MultiDISCO
    from dust_b200.inference.svmpc import SVMPC
    from dust_b200.inference.mpf import MPF

`install_as_dust()` aliases the package as `dust` so that code written against the reference
(`from dust.inference.svmpc import SVMPC`) imports this implementation unchanged.
"""
import importlib
import sys

__version__ = "0.1.0"

_SUBMODULES = [
    "controllers", "controllers.base", "controllers.disco",
    "inference", "inference.svgd", "inference.likelihoods", "inference.svmpc", "inference.mpf",
    "kernels", "kernels.base_kernels", "kernels.composite_kernels",
    "models", "models.base", "models.pendulum", "models.particle",
    "utils", "utils.spaces", "utils.obstacle_map", "utils.utf", "utils.simulations",
]


def install_as_dust():
    """Register `dust` / `dust.*` as aliases of this package in sys.modules."""
    pkg = sys.modules[__name__]
    sys.modules["dust"] = pkg
    for name in _SUBMODULES:
        sys.modules["dust." + name] = importlib.import_module(__name__ + "." + name)
    return pkg
