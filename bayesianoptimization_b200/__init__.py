"""bayesianoptimization_b200 - a B200-native GP-surrogate + acquisition engine that drops in
behind ``bayes_opt.BayesianOptimization.suggest()`` and the ``bayes_opt.acquisition`` classes.

Hot path: GP fit -> batched posterior predict -> UCB/EI/PoI (x constraint probability) ->
argmin/top-k, in hand-written sm_100a CUDA behind the C ABI declared in include/b200bo.h.
No CPU fallback: importing the compute classes without the built library raises ImportError.

Two layers:
  * GP seam / C ABI (needs numpy, scipy, sklearn):  B200GaussianProcessRegressor, FusedAcquisition
  * acquisition seam (a plug-in for the ``bayes_opt`` package, which must be importable):
    UpperConfidenceBound, ExpectedImprovement, ProbabilityOfImprovement, ConstantLiar, GPHedge,
    AcquisitionFunction, ConstraintModel, enable(optimizer) - resolved lazily on first access.
"""
from . import _lib
from ._build import build_library
from .dropin import accelerate_acquisition, enable
from .fused import FusedAcquisition
from .gpr import B200GaussianProcessRegressor, to_b200_gp

__version__ = "0.2.0"

_PLUGIN = {
    "AcquisitionFunction": "acquisition", "UpperConfidenceBound": "acquisition",
    "ExpectedImprovement": "acquisition", "ProbabilityOfImprovement": "acquisition",
    "ConstantLiar": "acquisition", "GPHedge": "acquisition", "DeviceHooks": "acquisition",
    "ConstraintModel": "constraint",
}


def __getattr__(name):
    mod = _PLUGIN.get(name)
    if mod is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    import importlib

    return getattr(importlib.import_module(f"{__name__}.{mod}"), name)


__all__ = [
    "B200GaussianProcessRegressor", "FusedAcquisition", "enable", "accelerate_acquisition", "to_b200_gp",
    "build_library", "__version__", *_PLUGIN,
]
