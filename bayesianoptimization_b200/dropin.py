"""``enable(optimizer)``: switch an existing ``bayes_opt.BayesianOptimization`` onto the B200
engine.  The reference constructs its GP objects privately (R/bayes_opt/bayesian_optimization.py:
124-130, R/bayes_opt/constraint.py:72-81) and offers no constructor injection, so the swap happens
on the three attributes the hot path reads:

  optimizer._gp                       -> B200GaussianProcessRegressor (same kernel / params / RNG)
  optimizer._space._constraint._model -> list of B200GaussianProcessRegressor
  optimizer._acquisition_function     -> same object, re-classed so that _get_acq /
                                         _random_sample_minimize run on the device
"""
from __future__ import annotations

from . import acquisition as A
from .gpr import B200GaussianProcessRegressor


def to_b200_gp(gp, device=0):
    """Same hyper-parameters, kernel object and RandomState, device numerics."""
    if isinstance(gp, B200GaussianProcessRegressor):
        return gp
    p = gp.get_params(deep=False)
    return B200GaussianProcessRegressor(device=device, **p)


_DEVICE_KINDS = {
    "UpperConfidenceBound": A.UpperConfidenceBound,
    "ProbabilityOfImprovement": A.ProbabilityOfImprovement,
    "ExpectedImprovement": A.ExpectedImprovement,
}


def accelerate_acquisition(acq):
    """Re-class a reference acquisition object in place: keeps all state (kappa/xi, decay, i,
    dummies), swaps in the device hooks."""
    if isinstance(acq, A.AcquisitionFunction):
        return acq
    name = type(acq).__name__
    if name == "ConstantLiar":
        acq.base_acquisition = accelerate_acquisition(acq.base_acquisition)
        hooks = A.AcquisitionFunction
        kind = None
    elif name == "GPHedge":
        acq.base_acquisitions = [accelerate_acquisition(a) for a in acq.base_acquisitions]
        return acq  # GPHedge itself only orchestrates; its bases and gp.predict run on the device
    elif name in _DEVICE_KINDS:
        hooks = _DEVICE_KINDS[name]
        kind = hooks._b200_kind
    else:
        hooks = A.AcquisitionFunction  # custom subclass: host base_acq, device mu/sigma
        kind = None
    ns = {
        "_b200_kind": kind,
        "_get_acq": A.AcquisitionFunction._get_acq,
        "_get_acq_generic": A.AcquisitionFunction._get_acq_generic,
        "_random_sample_minimize": A.AcquisitionFunction._random_sample_minimize,
        "_smart_minimize": A.AcquisitionFunction._smart_minimize,  # same algorithm, batched device calls
        "_acq_params": getattr(hooks, "_acq_params", A.AcquisitionFunction._acq_params),
    }
    acq.__class__ = type("B200" + name, (type(acq),), ns)
    return acq


def enable(optimizer, device=0):
    """Make ``optimizer.suggest()`` / ``maximize()`` / ``predict()`` run on the B200."""
    optimizer._gp = to_b200_gp(optimizer._gp, device)
    space = optimizer._space
    cm = getattr(space, "_constraint", None)
    if cm is not None:
        cm._model = [to_b200_gp(g, device) for g in cm._model]
    optimizer._acquisition_function = accelerate_acquisition(optimizer._acquisition_function)
    return optimizer
