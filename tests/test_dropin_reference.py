"""Drop-in mechanics against the REAL reference package (only where /root/reference exists, i.e.
the build container; skipped on the GPU box).  No device work: checks that ``enable`` swaps the
three seams and that the re-classed objects keep the reference's state and types."""
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
SHIMS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shims")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "bayes_opt")), reason="reference not present")


@pytest.fixture(scope="module")
def ref():
    for p in (REF, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib.metadata as md

    orig = md.version
    md.version = lambda n: "3.3.0" if n == "bayesian-optimization" else orig(n)
    import bayes_opt

    yield bayes_opt
    md.version = orig


def test_enable_swaps_gp_constraint_and_acquisition(ref):
    import __graft_entry__ as g

    g.build()
    import bayesianoptimization_b200 as bo
    from scipy.optimize import NonlinearConstraint

    con = NonlinearConstraint(lambda x, y: x + y, -np.inf, 4.0)
    opt = ref.BayesianOptimization(f=lambda x, y: -x**2 - (y - 1) ** 2 + 1, pbounds={"x": (2, 4), "y": (-3, 3)},
                                   constraint=con, random_state=1, verbose=0)
    rs = opt._random_state
    acq0 = opt._acquisition_function
    bo.enable(opt)
    assert isinstance(opt._gp, bo.B200GaussianProcessRegressor)
    assert opt._gp.random_state is rs and opt._gp.alpha == 1e-6 and opt._gp.n_restarts_optimizer == 5
    assert type(opt._gp.kernel).__name__ == "WrappedKernel"
    assert all(isinstance(m, bo.B200GaussianProcessRegressor) for m in opt._space._constraint._model)
    a = opt._acquisition_function
    assert a is acq0 and isinstance(a, ref.acquisition.ExpectedImprovement)
    assert type(a).__name__ == "B200ExpectedImprovement" and a.xi == 0.01
    assert type(a)._b200_kind == bo._lib.ACQ_EI
    # the reference's own machinery still drives the object (get/set params, decay)
    assert a.get_acquisition_params()["xi"] == 0.01
    # transform of a float-only space is the identity
    from bayesianoptimization_b200.gpr import probe_transform

    assert probe_transform(opt._gp.kernel, 2) is None


def test_enable_constant_liar_and_custom(ref):
    import bayesianoptimization_b200 as bo

    cl = ref.acquisition.ConstantLiar(ref.acquisition.UpperConfidenceBound(kappa=1.3))
    opt = ref.BayesianOptimization(f=None, pbounds={"x": (0, 1)}, acquisition_function=cl, verbose=0)
    bo.enable(opt)
    assert type(opt._acquisition_function.base_acquisition).__name__ == "B200UpperConfidenceBound"
    assert opt._acquisition_function.base_acquisition.kappa == 1.3

    class Custom(ref.acquisition.AcquisitionFunction):
        def base_acq(self, mean, std):
            return mean + std

    opt = ref.BayesianOptimization(f=None, pbounds={"x": (0, 1)}, acquisition_function=Custom(), verbose=0)
    bo.enable(opt)
    assert type(opt._acquisition_function)._b200_kind is None


def test_mirror_classes_match_reference_signatures(ref):
    """Same constructor parameters and hook names as the reference's classes."""
    import inspect

    import bayesianoptimization_b200 as bo

    for name in ("UpperConfidenceBound", "ProbabilityOfImprovement", "ExpectedImprovement", "ConstantLiar", "GPHedge"):
        r, m = getattr(ref.acquisition, name), getattr(bo, name)
        assert list(inspect.signature(r.__init__).parameters) == list(inspect.signature(m.__init__).parameters), name
        assert list(inspect.signature(r.suggest).parameters) == list(inspect.signature(m.suggest).parameters), name
    for hook in ("_fit_gp", "_get_acq", "_acq_min", "_random_sample_minimize", "_smart_minimize",
                 "get_acquisition_params", "set_acquisition_params", "base_acq", "suggest"):
        assert hasattr(bo.AcquisitionFunction, hook)
        rp = list(inspect.signature(getattr(ref.acquisition.AcquisitionFunction, hook)).parameters)
        mp = list(inspect.signature(getattr(bo.AcquisitionFunction, hook)).parameters)
        assert rp == mp, hook
    rc = inspect.signature(ref.constraint.ConstraintModel.__init__).parameters
    mc = inspect.signature(bo.ConstraintModel.__init__).parameters
    assert list(rc) == list(mc)[: len(rc)]


def test_space_mirror_matches_reference_targetspace(ref):
    import bayesianoptimization_b200 as bo

    pb = {"b": (0.0, 2.0), "a": (-1.0, 1.0), "c": (3.0, 9.0)}
    r = ref.target_space.TargetSpace(None, pb)
    m = bo.TargetSpace(None, pb)
    assert r.keys == m.keys and np.array_equal(r.bounds, m.bounds)
    assert np.array_equal(r.random_sample(1000, np.random.RandomState(4)), m.random_sample(1000, np.random.RandomState(4)))
    assert np.array_equal(r.random_sample(0, np.random.RandomState(4)), m.random_sample(0, np.random.RandomState(4)))
    for x, t in [([0.5, 0.1, 4.0], 1.0), ([1.5, -0.9, 8.0], 3.0), ([5.0, 0.0, 4.0], 9.0)]:
        r.register(np.array(x), t)
        m.register(np.array(x), t)
    assert r._target_max() == m._target_max() == 3.0
    assert np.array_equal(r.mask, m.mask)
    assert np.array_equal(r.continuous_dimensions, m.continuous_dimensions)
