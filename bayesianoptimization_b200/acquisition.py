"""Acquisition seam of the drop-in boundary.

The reference's plugin point is ``BayesianOptimization(acquisition_function=...)``: a subclass of
``bayes_opt.acquisition.AcquisitionFunction`` controls ``_get_acq``, ``_random_sample_minimize`` and
``_smart_minimize`` (R/bayes_opt/acquisition.py:171-219, :274-320, :322-420).  This module plugs the B200
engine in at exactly those three hooks and nothing else: the classes below ARE the reference's classes
(``bayes_opt`` is imported, not restated - constructors, ``suggest``, decay schedules, parameter
get/set, ConstantLiar's dummy bookkeeping, GPHedge's portfolio logic and every error message are
inherited), with ``DeviceHooks`` mixed in:

  _get_acq                 -> FusedAcquisition: the whole closure is ONE fused kernel launch
  _random_sample_minimize  -> candidates from the caller's RandomState exactly as the reference draws
                              them, evaluation + argmin + top-n_smart selection on the device(s)
  _smart_minimize          -> the n_smart L-BFGS-B runs advanced in lockstep, one device call per round
                              (mixed-integer spaces: the reference's own DE branch, unchanged, calling
                              the device closure)

``bayes_opt`` must be importable (this package is a plug-in for it).  The GP seam
(gpr.B200GaussianProcessRegressor), ``fused.FusedAcquisition`` and the C ABI do not need it.
"""
from __future__ import annotations

import abc

import numpy as np

try:
    from bayes_opt import acquisition as _ref
except ImportError as e:  # pragma: no cover - depends on the environment
    raise ImportError(
        "bayesianoptimization_b200.acquisition plugs into the bayes_opt package "
        "(bayesian-optimization >= 3.0), which is not importable here: install it, or use "
        "B200GaussianProcessRegressor / FusedAcquisition / the C ABI directly") from e

from . import _lib as B
from .fused import FusedAcquisition, _as_b200_gp, lockstep_lbfgsb

_STOCK = {
    _ref.UpperConfidenceBound: B.ACQ_UCB,
    _ref.ExpectedImprovement: B.ACQ_EI,
    _ref.ProbabilityOfImprovement: B.ACQ_POI,
}


def _device_kind(obj):
    """Device epilogue code if ``obj.base_acq`` is one of the reference's three formulas, else None
    (a user subclass that overrides base_acq keeps its override: mu/sigma come from the device, the
    formula runs where the user wrote it)."""
    for cls, kind in _STOCK.items():
        if isinstance(obj, cls) and type(obj).base_acq is cls.base_acq:
            return kind
    return None


class DeviceHooks(abc.ABC):
    """Mixin: the three hooks of the acquisition seam on the B200.  Must precede the reference class in
    the MRO.  (Derives from abc.ABC like bayes_opt's AcquisitionFunction so that both have the same
    instance layout: a live reference object can then be re-classed in place, see ``accelerate``.)"""

    def _get_acq(self, gp, constraint=None):
        _as_b200_gp(gp)
        kind = _device_kind(self)
        if kind is None:
            acq = super()._get_acq(gp=gp, constraint=constraint)  # gp.predict / constraint.predict on device
            acq.b200_vectorized = True  # maps (M,d) -> (M,): lockstep batching is safe
            return acq
        return FusedAcquisition(kind, gp, constraint, owner=self)

    # "host_rng": the candidates are TargetSpace.random_sample's MT19937 stream (parity with the reference);
    # "device_philox": throughput mode - they are generated inside the fused kernel (continuous spaces only;
    # ONE 64-bit seed is drawn from the caller's RandomState per call).  Set by enable(candidate_source=...).
    b200_candidate_source = "host_rng"

    def _random_sample_minimize(self, acq, space, random_state, n_random, n_x_seeds=0):
        if n_random == 0 or not isinstance(acq, FusedAcquisition) or n_x_seeds > B.MAX_TOPK:
            # (n_smart beyond the device's top-k capacity: evaluate on the device, select with numpy)
            return super()._random_sample_minimize(acq, space, random_state, n_random, n_x_seeds)
        if self.b200_candidate_source == "device_philox" and all(space.continuous_dimensions):
            seed = int(random_state.randint(0, 2**32, dtype=np.uint64)) << 32 | int(random_state.randint(0, 2**32, dtype=np.uint64))
            _, min_acq, x_min, _, x_seeds = acq.argmin_topk_philox(seed, space.bounds, n_random, n_x_seeds)
            return x_min, min_acq, (x_seeds if n_x_seeds != 0 else [])
        x_tries = space.random_sample(n_random, random_state=random_state)  # the reference's RNG stream
        idx, min_acq, top = acq.argmin_topk(x_tries, n_x_seeds)
        return x_tries[idx], min_acq, (x_tries[top] if n_x_seeds != 0 else [])

    def _smart_minimize(self, acq, space, x_seeds, random_state):
        batched = isinstance(acq, FusedAcquisition) or getattr(acq, "b200_vectorized", False)
        refine = acq.refine_mode() if isinstance(acq, FusedAcquisition) else _null()
        with refine:
            if not batched or len(x_seeds) == 0 or not all(space.continuous_dimensions):
                return super()._smart_minimize(acq, space, x_seeds, random_state)
            runs = [r for r in lockstep_lbfgsb(acq, x_seeds, space.bounds) if r.success]
        if not runs:
            return np.full(space.bounds.shape[0], np.nan), np.inf
        best = min(runs, key=lambda r: float(np.squeeze(r.fun)))  # first of equal minima, like the loop
        return np.clip(best.x, space.bounds[:, 0], space.bounds[:, 1]), np.squeeze(best.fun)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class AcquisitionFunction(DeviceHooks, _ref.AcquisitionFunction):
    """Base for user-defined acquisitions on the device GP: implement ``base_acq(mean, std)``."""


class UpperConfidenceBound(DeviceHooks, _ref.UpperConfidenceBound):
    """bayes_opt.acquisition.UpperConfidenceBound with the device hooks."""


class ProbabilityOfImprovement(DeviceHooks, _ref.ProbabilityOfImprovement):
    """bayes_opt.acquisition.ProbabilityOfImprovement with the device hooks."""


class ExpectedImprovement(DeviceHooks, _ref.ExpectedImprovement):
    """bayes_opt.acquisition.ExpectedImprovement with the device hooks."""


_HOOKED = {
    _ref.UpperConfidenceBound: UpperConfidenceBound,
    _ref.ProbabilityOfImprovement: ProbabilityOfImprovement,
    _ref.ExpectedImprovement: ExpectedImprovement,
}


def accelerate(acq, candidate_source=None):
    """Give an existing reference acquisition object the device hooks IN PLACE (all state kept: kappa/xi,
    decay counters, dummies, gains).  ConstantLiar / GPHedge only orchestrate: their base acquisitions are
    accelerated, the wrappers stay what they are."""
    if candidate_source not in (None, "host_rng", "device_philox"):
        raise ValueError("candidate_source must be 'host_rng' or 'device_philox'")
    if isinstance(acq, _ref.ConstantLiar):
        acq.base_acquisition = accelerate(acq.base_acquisition, candidate_source)
        return acq
    if isinstance(acq, _ref.GPHedge):
        acq.base_acquisitions = [accelerate(a, candidate_source) for a in acq.base_acquisitions]
        return acq
    if candidate_source is not None:
        acq.b200_candidate_source = candidate_source
    if isinstance(acq, DeviceHooks):
        return acq
    cls = type(acq)
    hooked = _HOOKED.get(cls)
    if hooked is None:  # user subclass (possibly of UCB/EI/PoI, possibly with its own base_acq)
        hooked = type("B200" + cls.__name__, (DeviceHooks, cls), {"__module__": __name__})
    acq.__class__ = hooked
    return acq


class ConstantLiar(_ref.ConstantLiar):
    """bayes_opt.acquisition.ConstantLiar whose base acquisition runs on the device."""

    def __init__(self, base_acquisition, *args, **kwargs):
        super().__init__(accelerate(base_acquisition), *args, **kwargs)


class GPHedge(_ref.GPHedge):
    """bayes_opt.acquisition.GPHedge over device base acquisitions."""

    def __init__(self, base_acquisitions, *args, **kwargs):
        super().__init__([accelerate(a) for a in base_acquisitions], *args, **kwargs)


# isinstance(x, b200.AcquisitionFunction) holds for every acquisition of this module, as
# isinstance(x, bayes_opt.acquisition.AcquisitionFunction) does in the reference (abc virtual subclasses:
# the concrete classes keep the reference's MRO).
for _cls in (UpperConfidenceBound, ProbabilityOfImprovement, ExpectedImprovement, ConstantLiar, GPHedge):
    AcquisitionFunction.register(_cls)
del _cls
