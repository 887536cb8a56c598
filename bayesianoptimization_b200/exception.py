"""Exception types with the reference's names (R/bayes_opt/exception.py).  When the reference
package is importable its own classes are re-used so ``except bayes_opt.exception.X`` keeps
working for drop-in users."""
try:  # pragma: no cover - depends on the environment
    from bayes_opt.exception import (  # type: ignore
        BayesianOptimizationError,
        ConstraintNotSupportedError,
        NoValidPointRegisteredError,
        NotUniqueError,
        TargetSpaceEmptyError,
    )
except Exception:  # reference not installed (e.g. on the GPU box)

    class BayesianOptimizationError(Exception):
        """Base class for exceptions in this package."""

    class NotUniqueError(BayesianOptimizationError):
        """A point is non-unique."""

    class ConstraintNotSupportedError(BayesianOptimizationError):
        """Constrained optimization is not supported by this acquisition function."""

    class NoValidPointRegisteredError(BayesianOptimizationError):
        """No registered point satisfies the constraints."""

    class TargetSpaceEmptyError(BayesianOptimizationError):
        """The target space is empty."""


__all__ = [
    "BayesianOptimizationError",
    "NotUniqueError",
    "ConstraintNotSupportedError",
    "NoValidPointRegisteredError",
    "TargetSpaceEmptyError",
]
