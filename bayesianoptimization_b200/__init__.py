"""bayesianoptimization_b200 - a B200-native GP-surrogate + acquisition engine that drops in
behind ``bayes_opt.BayesianOptimization.suggest()`` and the ``bayes_opt.acquisition`` classes.

Hot path: GP fit -> batched posterior predict -> UCB/EI/PoI (x constraint probability) ->
argmin/top-k, in hand-written sm_100a CUDA behind the C ABI declared in include/b200bo.h.
No CPU fallback: importing the compute classes without the built library raises ImportError.
"""
from . import _lib
from ._build import build_library
from .acquisition import (
    AcquisitionFunction,
    ConstantLiar,
    ExpectedImprovement,
    FusedAcquisition,
    GPHedge,
    ProbabilityOfImprovement,
    UpperConfidenceBound,
)
from .constraint import ConstraintModel
from .dropin import accelerate_acquisition, enable, to_b200_gp
from .gpr import B200GaussianProcessRegressor
from .space import TargetSpace

__version__ = "0.1.0"

__all__ = [
    "AcquisitionFunction", "ConstantLiar", "ExpectedImprovement", "FusedAcquisition", "GPHedge",
    "ProbabilityOfImprovement", "UpperConfidenceBound", "ConstraintModel",
    "B200GaussianProcessRegressor", "TargetSpace", "enable", "accelerate_acquisition",
    "to_b200_gp", "build_library", "__version__",
]
