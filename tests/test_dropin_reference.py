"""Drop-in mechanics against the REAL reference package (vendored, unmodified: tools/vendor_ref.py).
No device work here: checks that ``enable`` swaps the three seams, that the hooked objects ARE the
reference's objects (state and types kept), and that the product package restates nothing of them.
The device-side behaviour of the same objects is tests/test_gpu_dropin_live.py."""
import inspect
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bo():
    import __graft_entry__ as g

    g.build()
    import bayesianoptimization_b200 as bo

    return bo


def test_enable_swaps_gp_constraint_and_acquisition(bo, ref):
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
    assert a is acq0 and isinstance(a, ref.acquisition.ExpectedImprovement) and isinstance(a, bo.DeviceHooks)
    assert type(a) is bo.ExpectedImprovement and a.xi == 0.01
    # the reference's own machinery still drives the object (get/set params, decay)
    assert a.get_acquisition_params()["xi"] == 0.01
    # transform of a float-only space is the identity
    from bayesianoptimization_b200.gpr import probe_transform

    assert probe_transform(opt._gp.kernel, 2) is None
    # multi-device option reaches every GP of the optimizer
    opt2 = ref.BayesianOptimization(f=None, pbounds={"x": (2, 4)}, constraint=con, random_state=1, verbose=0)
    bo.enable(opt2, devices=[0, 1], precision="fp32")
    assert opt2._gp.device_list() == [0, 1] and opt2._gp.precision == "fp32"
    assert opt2._space._constraint._model[0].device_list() == [0, 1]


def test_enable_constant_liar_gphedge_and_custom(bo, ref):
    from bayesianoptimization_b200.acquisition import _device_kind
    from bayesianoptimization_b200 import _lib as B

    A = ref.acquisition
    cl = A.ConstantLiar(A.UpperConfidenceBound(kappa=1.3))
    opt = ref.BayesianOptimization(f=None, pbounds={"x": (0, 1)}, acquisition_function=cl, verbose=0)
    bo.enable(opt)
    assert opt._acquisition_function is cl and type(cl) is A.ConstantLiar  # the wrapper only orchestrates
    assert type(cl.base_acquisition) is bo.UpperConfidenceBound and cl.base_acquisition.kappa == 1.3
    assert _device_kind(cl.base_acquisition) == B.ACQ_UCB

    hedge = A.GPHedge([A.UpperConfidenceBound(kappa=2.0), A.ExpectedImprovement(xi=0.01)])
    opt = ref.BayesianOptimization(f=None, pbounds={"x": (0, 1)}, acquisition_function=hedge, verbose=0)
    bo.enable(opt)
    assert [type(b) for b in hedge.base_acquisitions] == [bo.UpperConfidenceBound, bo.ExpectedImprovement]

    class Custom(A.AcquisitionFunction):
        def base_acq(self, mean, std):
            return mean + std

    class TweakedEI(A.ExpectedImprovement):  # overrides the formula: must NOT get the built-in device epilogue
        def base_acq(self, mean, std):
            return super().base_acq(mean, std) + 1.0

    class PlainSubEI(A.ExpectedImprovement):  # does not: keeps the fused path
        pass

    for obj, kind in ((Custom(), None), (TweakedEI(xi=0.1), None), (PlainSubEI(xi=0.1), B.ACQ_EI)):
        opt = ref.BayesianOptimization(f=None, pbounds={"x": (0, 1)}, acquisition_function=obj, verbose=0)
        bo.enable(opt)
        a = opt._acquisition_function
        assert a is obj and isinstance(a, bo.DeviceHooks) and isinstance(a, type(obj).__mro__[2])
        assert _device_kind(a) == kind, type(a).__mro__


def test_hooked_classes_are_the_reference_classes(bo, ref):
    """Nothing is restated: constructors, suggest(), decay, get/set params, ConstantLiar / GPHedge logic and
    ConstraintModel are inherited from bayes_opt; only the three hooks are defined here."""
    A = ref.acquisition
    for name in ("UpperConfidenceBound", "ProbabilityOfImprovement", "ExpectedImprovement", "ConstantLiar", "GPHedge",
                 "AcquisitionFunction"):
        r, m = getattr(A, name), getattr(bo, name)
        assert issubclass(m, r), name
        for meth in ("suggest", "get_acquisition_params", "set_acquisition_params", "base_acq", "_acq_min", "_fit_gp"):
            assert getattr(m, meth) is getattr(r, meth), (name, meth)
    for name in ("UpperConfidenceBound", "ProbabilityOfImprovement", "ExpectedImprovement", "AcquisitionFunction"):
        m = getattr(bo, name)
        assert m.__init__ is getattr(A, name).__init__
        for hook in ("_get_acq", "_random_sample_minimize", "_smart_minimize"):
            assert getattr(m, hook) is getattr(bo.DeviceHooks, hook)
            rp = list(inspect.signature(getattr(A.AcquisitionFunction, hook)).parameters)
            mp = list(inspect.signature(getattr(m, hook)).parameters)
            assert rp == mp, hook
    assert bo.ConstraintModel.predict is ref.constraint.ConstraintModel.predict
    assert bo.ConstraintModel.fit is ref.constraint.ConstraintModel.fit


def _code_tokens(path):
    import io
    import tokenize

    toks = []
    with open(path, "rb") as f:
        for t in tokenize.tokenize(io.BytesIO(f.read()).readline):
            if t.type in (tokenize.COMMENT, tokenize.NL, tokenize.NEWLINE, tokenize.INDENT, tokenize.DEDENT,
                          tokenize.ENCODING, tokenize.ENDMARKER):
                continue
            if t.type == tokenize.STRING and t.string.lstrip("rRbBuU")[:3] in ('"""', "'''"):
                continue  # docstrings
            toks.append(t.string)
    return toks


def test_product_package_does_not_transcribe_the_reference(ref):
    """Token-stream similarity (comments/docstrings stripped) of every product file against every reference
    module stays below 0.3 (acquisition.py was 0.64, constraint.py 0.70), and less than 30 % of a
    product file's tokens sit in runs of >= 6 tokens shared with a reference module (shorter matches are
    punctuation noise)."""
    import difflib

    pkg = os.path.join(ROOT, "bayesianoptimization_b200")
    refdir = os.path.dirname(ref.__file__)
    ref_files = [os.path.join(refdir, f) for f in os.listdir(refdir) if f.endswith(".py")]
    ref_toks = {f: _code_tokens(f) for f in ref_files}
    for fn in sorted(os.listdir(pkg)):
        if not fn.endswith(".py"):
            continue
        mine = _code_tokens(os.path.join(pkg, fn))
        if len(mine) < 50:
            continue
        for rf, rt in ref_toks.items():
            sm = difflib.SequenceMatcher(None, mine, rt, autojunk=False)
            runs = sum(b.size for b in sm.get_matching_blocks() if b.size >= 6)
            frac = runs / len(mine)
            assert sm.ratio() < 0.3 and frac < 0.3, (fn, os.path.basename(rf), round(sm.ratio(), 3), round(frac, 3))
