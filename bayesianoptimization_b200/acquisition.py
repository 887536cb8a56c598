"""Acquisition seam of the drop-in boundary: B200 versions of
bayes_opt.acquisition.{AcquisitionFunction, UpperConfidenceBound, ProbabilityOfImprovement,
ExpectedImprovement, ConstantLiar} (R/bayes_opt/acquisition.py).

Same class names, constructor arguments, hooks (``_fit_gp``, ``_get_acq``, ``_acq_min``,
``_random_sample_minimize``, ``_smart_minimize``, ``base_acq``, ``get/set_acquisition_params``),
decay schedules and exception types as the reference.  What changes is where the arithmetic
runs: ``_get_acq`` returns a closure that evaluates  -base_acq(mu, sigma) [* p_constraint]  for a
whole candidate batch in ONE fused sm_100a kernel launch, and ``_random_sample_minimize`` does
evaluate + argmin + top-k seeds on the device.  Instances duck-type with the reference's
``BayesianOptimization(acquisition_function=...)``; ``dropin.enable`` swaps
them into an existing optimizer.
"""
from __future__ import annotations

import abc
import ctypes as C
import warnings
from copy import deepcopy

import numpy as np
from packaging import version
from scipy import __version__ as scipy_version
from scipy.optimize import minimize
from scipy.optimize._differentialevolution import DifferentialEvolutionSolver
from scipy.special import softmax
from scipy.stats import norm

from . import _lib as B
from .exception import (
    ConstraintNotSupportedError,
    NoValidPointRegisteredError,
    TargetSpaceEmptyError,
)
from .gpr import B200GaussianProcessRegressor
from .space import ensure_rng


def _as_b200_gp(gp):
    if not isinstance(gp, B200GaussianProcessRegressor):
        raise TypeError(
            "the B200 acquisition functions need a B200GaussianProcessRegressor (got "
            f"{type(gp).__name__}); use bayesianoptimization_b200.enable(optimizer) or construct "
            "the GP with B200GaussianProcessRegressor - there is no CPU fallback")
    return gp


class FusedAcquisition:
    """Callable closure over fitted device GPs: x (M,d)|(d,) -> (M,) negated acquisition.
    Replaces the closure built at R/bayes_opt/acquisition.py:196-219."""

    def __init__(self, kind, gp, constraint=None, kappa=0.0, xi=0.0, y_max=None):
        gp = _as_b200_gp(gp)
        gp._ensure_device_fit()
        self.dim = gp.X_train_.shape[1]
        self._keep = [gp]
        spec = B.AcqSpec()
        spec.kind = kind
        spec.kappa = float(kappa)
        spec.xi = float(xi)
        spec.y_max = float(y_max) if y_max is not None else 0.0
        spec.gps[0] = gp._handle().ptr.value
        n = 1
        if constraint is not None:
            models = constraint.model
            if len(models) + 1 > B.MAX_GPS:
                raise NotImplementedError(f"at most {B.MAX_GPS - 1} constraint GPs are supported")
            for j, cgp in enumerate(models):
                cgp = _as_b200_gp(cgp)
                cgp._ensure_device_fit()
                self._keep.append(cgp)
                spec.gps[n] = cgp._handle().ptr.value
                spec.lb[n] = float(constraint.lb[j])
                spec.ub[n] = float(constraint.ub[j])
                n += 1
        spec.n_gps = n
        self.spec = spec

    def _candidates(self, x):
        for g in self._keep:  # an LML evaluation in between re-uses the factor buffers: refit lazily
            g._ensure_device_fit()
        x = B.c_f64(np.asarray(x, dtype=np.float64).reshape(-1, self.dim))
        if not np.isfinite(x).all():  # sklearn's predict raises the same way (validate_data)
            raise ValueError("Input X contains NaN or infinity.")
        # kernels with a host-side input transform (categorical one-hot): every GP of the call must
        # see the same transformed batch, as in the reference where they share space.kernel_transform
        modes = [g.__dict__.get("_b200_xform", ("device", None)) for g in self._keep]
        host = [a for m, a in modes if m == "host"]
        if host:
            if len(host) != len(modes) or any(h is not host[0] for h in host):
                raise NotImplementedError("GPs of one acquisition call use different host-side input transforms")
            x = self._keep[0]._device_candidates(x)
        return x

    def __call__(self, x):
        x = self._candidates(x)
        out = np.empty(x.shape[0])
        B.check(B.lib().b200bo_acq_eval(C.byref(self.spec), B.as_dp(x), x.shape[0], B.as_dp(out)))
        return out

    def argmin_topk(self, x, k):
        """Evaluate + np.argmin + k smallest (value, index) on the device
        (R/bayes_opt/acquisition.py:312-317)."""
        x = self._candidates(x)
        best_val = C.c_double()
        best_idx = C.c_int64()
        tv = np.empty(max(k, 1))
        ti = np.empty(max(k, 1), dtype=np.int64)
        B.check(B.lib().b200bo_acq_argmin_topk(
            C.byref(self.spec), B.as_dp(x), x.shape[0], int(k), C.byref(best_val), C.byref(best_idx),
            B.as_dp(tv), ti.ctypes.data_as(C.POINTER(C.c_int64)), None))
        ti = ti[:k]
        return best_idx.value, best_val.value, ti[ti >= 0]


class AcquisitionFunction(abc.ABC):
    """Mirror of bayes_opt.acquisition.AcquisitionFunction (R/bayes_opt/acquisition.py:56-420)."""

    _b200_kind = None  # subclasses with a device epilogue set B.ACQ_*

    def __init__(self, random_state=None):
        if random_state is not None:
            msg = ("Providing a random_state to an acquisition function during initialization is deprecated "
                   "and will be ignored. The random_state is instead provided automatically during the "
                   "suggest() call.")
            warnings.warn(msg, DeprecationWarning, stacklevel=2)
        self.i = 0

    @abc.abstractmethod
    def base_acq(self, *args, **kwargs):
        """Provide access to the base acquisition function."""

    def _fit_gp(self, gp, target_space):
        """R/bayes_opt/acquisition.py:79-86."""
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gp.fit(target_space.params, target_space.target)
            if target_space.constraint is not None:
                target_space.constraint.fit(target_space.params, target_space._constraint_values)

    def get_acquisition_params(self):
        raise NotImplementedError(
            "Custom AcquisitionFunction subclasses must implement their own get_acquisition_params method.")

    def set_acquisition_params(self, params):
        raise NotImplementedError(
            "Custom AcquisitionFunction subclasses must implement their own set_acquisition_params method.")

    def suggest(self, gp, target_space, n_random=10_000, n_smart=10, fit_gp=True, random_state=None):
        """R/bayes_opt/acquisition.py:116-169."""
        random_state = ensure_rng(random_state)
        if len(target_space) == 0:
            msg = ("Cannot suggest a point without previous samples. Use "
                   " target_space.random_sample() to generate a point and "
                   " target_space.probe(*) to evaluate it.")
            raise TargetSpaceEmptyError(msg)
        self.i += 1
        if fit_gp:
            self._fit_gp(gp=gp, target_space=target_space)
        acq = self._get_acq(gp=gp, constraint=target_space.constraint)
        return self._acq_min(acq, target_space, n_random=n_random, n_smart=n_smart, random_state=random_state)

    # -- device closure ------------------------------------------------------------------
    def _acq_params(self):
        return {}

    def _get_acq(self, gp, constraint=None):
        """R/bayes_opt/acquisition.py:171-219, fused on the device."""
        if self._b200_kind is None:
            return self._get_acq_generic(gp, constraint)
        return FusedAcquisition(self._b200_kind, gp, constraint, **self._acq_params())

    def _get_acq_generic(self, gp, constraint=None):
        """Custom ``base_acq`` written in numpy by a user subclass: mu/sigma (and the constraint
        probabilities) still come from the device; only the user's O(M) formula runs on host."""
        gp = _as_b200_gp(gp)
        dim = gp.X_train_.shape[1]

        def acq(x):
            x = np.asarray(x, dtype=float).reshape(-1, dim)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                mean, std = gp.predict(x, return_std=True)
                if constraint is not None:
                    return -1 * self.base_acq(mean, std) * constraint.predict(x)
            return -1 * self.base_acq(mean, std)

        acq._b200_vectorized = True
        return acq

    def _acq_min(self, acq, space, random_state, n_random=10_000, n_smart=10):
        """R/bayes_opt/acquisition.py:221-272."""
        if n_random == 0 and n_smart == 0:
            raise ValueError("Either n_random or n_smart needs to be greater than 0.")
        x_min_r, min_acq_r, x_seeds = self._random_sample_minimize(
            acq, space, random_state, n_random=max(n_random, n_smart), n_x_seeds=n_smart)
        if n_smart:
            x_min_s, min_acq_s = self._smart_minimize(acq, space, x_seeds=x_seeds, random_state=random_state)
            if min_acq_r > min_acq_s:
                return x_min_s
        return x_min_r

    def _random_sample_minimize(self, acq, space, random_state, n_random, n_x_seeds=0):
        """R/bayes_opt/acquisition.py:274-320; evaluation + argmin + top-k on the device."""
        if n_random == 0:
            return None, np.inf, space.random_sample(n_x_seeds, random_state=random_state)
        x_tries = space.random_sample(n_random, random_state=random_state)
        if isinstance(acq, FusedAcquisition):
            idx, min_acq, top = acq.argmin_topk(x_tries, n_x_seeds)
            x_min = x_tries[idx]
            x_seeds = x_tries[top] if n_x_seeds != 0 else []
            return x_min, min_acq, x_seeds
        ys = acq(x_tries)
        x_min = x_tries[ys.argmin()]
        min_acq = ys.min()
        if n_x_seeds != 0:
            idxs = np.argsort(ys)[:n_x_seeds]
            x_seeds = x_tries[idxs]
        else:
            x_seeds = []
        return x_min, min_acq, x_seeds

    def _smart_minimize(self, acq, space, x_seeds, random_state):
        """R/bayes_opt/acquisition.py:322-420, continuous branch (:364-374): n_smart L-BFGS-B runs
        with SciPy's finite-difference gradient; every objective evaluation is a device call."""
        continuous_dimensions = space.continuous_dimensions
        continuous_bounds = space.bounds[continuous_dimensions]
        min_acq = None
        x_min = None
        if all(continuous_dimensions):
            # the n_smart runs are independent: advance them in lockstep so that every round of
            # objective / stencil requests of ALL seeds is one device call (same iterates per seed)
            # (only for closures known to map (M,d) -> (M,); arbitrary user callables run one by one)
            batched = isinstance(acq, FusedAcquisition) or getattr(acq, "_b200_vectorized", False)
            for res in _lockstep_lbfgsb(acq, x_seeds, continuous_bounds, lockstep=batched):
                if not res.success:
                    continue
                if min_acq is None or np.squeeze(res.fun) < min_acq:
                    x_min = res.x
                    min_acq = np.squeeze(res.fun)
        else:
            # mixed-integer branch (R/bayes_opt/acquisition.py:376-412): SciPy's differential
            # evolution over all dimensions seeded with x_seeds, then an L-BFGS-B polish of the
            # continuous ones.  Host driver as in the reference; every objective call is a device
            # call (small-batch kernels); DE's immediate-updating semantics are kept.
            xinit = space.random_sample(15 * len(space.bounds), random_state=random_state)
            if len(x_seeds) > 0:
                n_seeds = min(len(x_seeds), len(xinit))
                xinit[:n_seeds] = x_seeds[:n_seeds]
            de_parameters = {"func": acq, "bounds": space.bounds, "polish": False, "init": xinit}
            if version.parse(scipy_version) < version.parse("1.15.0"):
                de_parameters["seed"] = random_state
            else:
                de_parameters["rng"] = random_state
            de = DifferentialEvolutionSolver(**de_parameters)
            res_de = de.solve()
            if not res_de.success:
                raise RuntimeError(f"Differential evolution optimization failed. Message: {res_de.message}")
            x_min = res_de.x
            min_acq = np.squeeze(res_de.fun)
            if any(continuous_dimensions):
                x_try = x_min.copy()

                def continuous_acq(x, x_try=x_try):
                    x_try[continuous_dimensions] = x
                    return acq(x_try)

                res = minimize(continuous_acq, x_min[continuous_dimensions], bounds=continuous_bounds)
                if res.success and np.squeeze(res.fun) < min_acq:
                    x_try[continuous_dimensions] = res.x
                    x_min = x_try
                    min_acq = np.squeeze(res.fun)
        if min_acq is None:
            min_acq = np.inf
            x_min = np.array([np.nan] * space.bounds.shape[0])
        return np.clip(x_min, space.bounds[:, 0], space.bounds[:, 1]), min_acq


def _stencil_options(acq):
    """SciPy's L-BFGS-B (jac=None) builds its 2-point forward-difference gradient by mapping the
    objective over the d stencil points x + h_i e_i (SP/optimize/_numdiff.py:693-712).  Since SciPy
    1.16 that map is pluggable (``workers``): evaluate the whole stencil in ONE device call instead
    of d single-row calls.  Same points, same differences, same iterates as the reference."""
    try:
        from packaging import version
        from scipy import __version__ as scipy_version

        if version.parse(scipy_version) < version.parse("1.16.0"):
            return None
    except Exception:  # pragma: no cover
        return None

    def batched_map(fun, iterable):
        xs = [np.asarray(x, dtype=float) for x in iterable]
        if not xs:
            return []
        ys = np.asarray(acq(np.vstack(xs)), dtype=float)
        return [np.atleast_1d(y) for y in ys]

    return {"workers": batched_map}


class _LockstepEvaluator:
    """Serves the pending objective requests of several concurrently running SciPy minimisations
    with ONE call of the (device) closure.  Each minimisation runs in its own thread and blocks in
    ``evaluate`` until every still-active run has submitted its request; the last one to arrive
    evaluates the concatenated batch.  Per-candidate results of the fused kernels do not depend on
    what else is in the batch, so each run sees exactly the values it would see alone."""

    def __init__(self, acq, n_active):
        import threading

        self.acq = acq
        self.cv = threading.Condition()
        self.pending, self.results = {}, {}
        self.active = n_active
        self.error = None

    def _flush(self):
        keys = list(self.pending)
        xs = [self.pending[k] for k in keys]
        try:
            ys = np.asarray(self.acq(np.vstack(xs)), dtype=float)
            off = 0
            for k, x in zip(keys, xs):
                self.results[k] = ys[off:off + len(x)]
                off += len(x)
        except BaseException as e:  # propagate to every waiting run
            self.error = e
        self.pending.clear()
        self.cv.notify_all()

    def evaluate(self, key, x):
        with self.cv:
            if self.error is not None:
                raise self.error
            self.pending[key] = np.atleast_2d(np.asarray(x, dtype=float))
            if len(self.pending) >= self.active:
                self._flush()
            while key not in self.results and self.error is None:
                self.cv.wait()
            if self.error is not None:
                raise self.error
            return self.results.pop(key)

    def finish(self, key):
        with self.cv:
            self.active -= 1
            if self.pending and len(self.pending) >= self.active:
                self._flush()


def _lockstep_lbfgsb(acq, x_seeds, bounds, lockstep=True):
    """``[minimize(acq, seed, bounds=bounds, method="L-BFGS-B") for seed in x_seeds]`` (the loop at
    R/bayes_opt/acquisition.py:365-366) with the runs advanced in lockstep.  B200BO_LOCKSTEP=0 (or a
    single seed) falls back to the plain sequential loop."""
    import os
    import threading

    seeds = [np.asarray(s, dtype=float) for s in x_seeds]
    if len(seeds) <= 1 or not lockstep or os.environ.get("B200BO_LOCKSTEP", "1") == "0":
        options = _stencil_options(acq) if lockstep else None
        return [minimize(acq, s, bounds=bounds, method="L-BFGS-B", options=options) for s in seeds]
    ev = _LockstepEvaluator(acq, len(seeds))
    results, errors = [None] * len(seeds), [None] * len(seeds)
    use_workers = _stencil_options(acq) is not None

    def run(i):
        try:
            def fun(x):
                return ev.evaluate(i, x)

            options = None
            if use_workers:
                def stencil_map(_f, iterable):
                    xs = [np.asarray(x, dtype=float) for x in iterable]
                    return [np.atleast_1d(y) for y in ev.evaluate(i, np.vstack(xs))] if xs else []

                options = {"workers": stencil_map}
            results[i] = minimize(fun, seeds[i], bounds=bounds, method="L-BFGS-B", options=options)
        except BaseException as e:
            errors[i] = e
        finally:
            ev.finish(i)

    threads = [threading.Thread(target=run, args=(i,), daemon=True) for i in range(len(seeds))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    return results


def _check_decay(exploration_decay, exploration_decay_delay):
    if exploration_decay is not None and not (0 < exploration_decay <= 1):
        raise ValueError("exploration_decay must be greater than 0 and less than or equal to 1.")
    if exploration_decay_delay is not None and (
        not isinstance(exploration_decay_delay, int) or exploration_decay_delay < 0
    ):
        raise ValueError("exploration_decay_delay must be an integer greater than or equal to 0.")


class UpperConfidenceBound(AcquisitionFunction):
    """mu + kappa*sigma (R/bayes_opt/acquisition.py:423-580)."""

    _b200_kind = B.ACQ_UCB

    def __init__(self, kappa=2.576, exploration_decay=None, exploration_decay_delay=None, random_state=None):
        if kappa < 0:
            raise ValueError("kappa must be greater than or equal to 0.")
        _check_decay(exploration_decay, exploration_decay_delay)
        super().__init__(random_state=random_state)
        self.kappa = kappa
        self.exploration_decay = exploration_decay
        self.exploration_decay_delay = exploration_decay_delay

    def base_acq(self, mean, std):
        return mean + self.kappa * std

    def _acq_params(self):
        return dict(kappa=self.kappa)

    def suggest(self, gp, target_space, n_random=10_000, n_smart=10, fit_gp=True, random_state=None):
        if target_space.constraint is not None:
            msg = (f"Received constraints, but acquisition function {type(self)} "
                   "does not support constrained optimization.")
            raise ConstraintNotSupportedError(msg)
        x_max = super().suggest(gp=gp, target_space=target_space, n_random=n_random, n_smart=n_smart,
                                fit_gp=fit_gp, random_state=random_state)
        self.decay_exploration()
        return x_max

    def decay_exploration(self):
        if self.exploration_decay is not None and (
            self.exploration_decay_delay is None or self.exploration_decay_delay <= self.i
        ):
            self.kappa = self.kappa * self.exploration_decay

    def get_acquisition_params(self):
        return {"kappa": self.kappa, "exploration_decay": self.exploration_decay,
                "exploration_decay_delay": self.exploration_decay_delay}

    def set_acquisition_params(self, params):
        self.kappa = params["kappa"]
        self.exploration_decay = params["exploration_decay"]
        self.exploration_decay_delay = params["exploration_decay_delay"]


class _XiAcquisition(AcquisitionFunction):
    """Shared body of PoI / EI (R/bayes_opt/acquisition.py:583-760, :763-949)."""

    def __init__(self, xi, exploration_decay=None, exploration_decay_delay=None, random_state=None):
        if xi < 0:
            raise ValueError("xi must be greater than or equal to 0.")
        _check_decay(exploration_decay, exploration_decay_delay)
        super().__init__(random_state=random_state)
        self.xi = xi
        self.exploration_decay = exploration_decay
        self.exploration_decay_delay = exploration_decay_delay
        self.y_max = None

    def _acq_params(self):
        if self.y_max is None:
            raise ValueError("y_max is not set. If you are calling this method outside "
                             "of suggest(), you must set y_max manually.")
        return dict(xi=self.xi, y_max=self.y_max)

    def suggest(self, gp, target_space, n_random=10_000, n_smart=10, fit_gp=True, random_state=None):
        y_max = target_space._target_max()
        if y_max is None and not target_space.empty:
            msg = ("Cannot suggest a point without an allowed point. Use "
                   "target_space.random_sample() to generate a point until "
                   " at least one point that satisfies the constraints is found.")
            raise NoValidPointRegisteredError(msg)
        self.y_max = y_max
        x_max = super().suggest(gp=gp, target_space=target_space, n_random=n_random, n_smart=n_smart,
                                fit_gp=fit_gp, random_state=random_state)
        self.decay_exploration()
        return x_max

    def decay_exploration(self):
        if self.exploration_decay is not None and (
            self.exploration_decay_delay is None or self.exploration_decay_delay <= self.i
        ):
            self.xi = self.xi * self.exploration_decay

    def get_acquisition_params(self):
        return {"xi": self.xi, "exploration_decay": self.exploration_decay,
                "exploration_decay_delay": self.exploration_decay_delay}

    def set_acquisition_params(self, params):
        self.xi = params["xi"]
        self.exploration_decay = params["exploration_decay"]
        self.exploration_decay_delay = params["exploration_decay_delay"]


class ProbabilityOfImprovement(_XiAcquisition):
    """Phi((mu - y_max - xi)/sigma) (R/bayes_opt/acquisition.py:633-661)."""

    _b200_kind = B.ACQ_POI

    def base_acq(self, mean, std):
        if self.y_max is None:
            raise ValueError("y_max is not set. If you are calling this method outside "
                             "of suggest(), you must set y_max manually.")
        z = (mean - self.y_max - self.xi) / std
        return norm.cdf(z)


class ExpectedImprovement(_XiAcquisition):
    """a*Phi(a/sigma) + sigma*phi(a/sigma), a = mu - y_max - xi (R/bayes_opt/acquisition.py:820-849)."""

    _b200_kind = B.ACQ_EI

    def base_acq(self, mean, std):
        if self.y_max is None:
            raise ValueError("y_max is not set. If you are calling this method outside "
                             "of suggest(), ensure y_max is set, or set it manually.")
        a = mean - self.y_max - self.xi
        z = a / std
        return a * norm.cdf(z) + std * norm.pdf(z)


class ConstantLiar(AcquisitionFunction):
    """R/bayes_opt/acquisition.py:952-1178: re-fit on a copy of the space that contains the
    pending suggestions with a lied-about target, then delegate to the base acquisition."""

    def __init__(self, base_acquisition, strategy="max", random_state=None, atol=1e-5, rtol=1e-8):
        super().__init__(random_state)
        self.base_acquisition = base_acquisition
        self.dummies = []
        if not isinstance(strategy, float) and strategy not in ["min", "mean", "max"]:
            raise ValueError(f"Received invalid argument {strategy} for strategy.")
        self.strategy = strategy
        self.atol = atol
        self.rtol = rtol

    def base_acq(self, *args, **kwargs):
        return self.base_acquisition.base_acq(*args, **kwargs)

    def _copy_target_space(self, target_space):
        """R/bayes_opt/acquisition.py:1013-1037."""
        keys = target_space.keys
        pbounds = {key: bound for key, bound in zip(keys, target_space.bounds)}
        target_space_copy = type(target_space)(
            None, pbounds=pbounds, allow_duplicate_points=target_space._allow_duplicate_points)
        if target_space._constraint is not None:
            target_space_copy.set_constraint(deepcopy(target_space.constraint))
        target_space_copy._params = deepcopy(target_space._params)
        target_space_copy._target = deepcopy(target_space._target)
        return target_space_copy

    def _remove_expired_dummies(self, target_space):
        """R/bayes_opt/acquisition.py:1039-1056."""
        dummies = []
        for dummy in self.dummies:
            close = np.isclose(dummy, target_space.params, rtol=self.rtol, atol=self.atol)
            if not close.all(axis=1).any():
                dummies.append(dummy)
        self.dummies = dummies

    def suggest(self, gp, target_space, n_random=10_000, n_smart=10, fit_gp=True, random_state=None):
        if len(target_space) == 0:
            msg = ("Cannot suggest a point without previous samples. Use "
                   " target_space.random_sample() to generate a point and "
                   " target_space.probe(*) to evaluate it.")
            raise TargetSpaceEmptyError(msg)
        if target_space.constraint is not None:
            msg = (f"Received constraints, but acquisition function {type(self)} "
                   "does not support constrained optimization.")
            raise ConstraintNotSupportedError(msg)
        self._remove_expired_dummies(target_space)
        dummy_target_space = self._copy_target_space(target_space)
        if isinstance(self.strategy, float):
            dummy_target = self.strategy
        elif self.strategy == "min":
            dummy_target = target_space.target.min()
        elif self.strategy == "mean":
            dummy_target = target_space.target.mean()
        elif self.strategy != "max":
            raise ValueError(f"Received invalid argument {self.strategy} for strategy.")
        else:
            dummy_target = target_space.target.max()
        for dummy in self.dummies:
            dummy_target_space.register(dummy, dummy_target)
        self._fit_gp(gp=gp, target_space=dummy_target_space)
        x_max = self.base_acquisition.suggest(gp, dummy_target_space, n_random=n_random, n_smart=n_smart,
                                              fit_gp=False, random_state=random_state)
        self.dummies.append(x_max)
        return x_max

    def get_acquisition_params(self):
        return {"dummies": [dummy.tolist() for dummy in self.dummies],
                "base_acquisition_params": self.base_acquisition.get_acquisition_params(),
                "strategy": self.strategy, "atol": self.atol, "rtol": self.rtol}

    def set_acquisition_params(self, params):
        self.dummies = [np.array(dummy) for dummy in params["dummies"]]
        self.base_acquisition.set_acquisition_params(params["base_acquisition_params"])
        self.strategy = params["strategy"]
        self.atol = params["atol"]
        self.rtol = params["rtol"]


class GPHedge(AcquisitionFunction):
    """Portfolio of base acquisitions chosen by softmax of cumulative rewards
    (R/bayes_opt/acquisition.py:1181-1360).  Host logic; the rewards (posterior means of the previous
    candidates) and every base ``suggest`` run on the device."""

    def __init__(self, base_acquisitions, random_state=None):
        super().__init__(random_state)
        self.base_acquisitions = list(base_acquisitions)
        self.n_acq = len(self.base_acquisitions)
        self.gains = np.zeros(self.n_acq)
        self.previous_candidates = None

    def base_acq(self, *args, **kwargs):
        msg = ("GPHedge base acquisition function is ambiguous."
               " You may use self.base_acquisitions[i].base_acq(mean, std)"
               " to get the base acquisition function for the i-th acquisition.")
        raise TypeError(msg)

    def _sample_idx_from_softmax_gains(self, random_state):
        cumsum_softmax_g = np.cumsum(softmax(self.gains))
        r = random_state.rand()
        return np.argmax(r <= cumsum_softmax_g)

    def _update_gains(self, gp):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rewards = gp.predict(self.previous_candidates)
        self.gains += rewards
        self.previous_candidates = None

    def suggest(self, gp, target_space, n_random=10_000, n_smart=10, fit_gp=True, random_state=None):
        if len(target_space) == 0:
            msg = ("Cannot suggest a point without previous samples. Use "
                   " target_space.random_sample() to generate a point and "
                   " target_space.probe(*) to evaluate it.")
            raise TargetSpaceEmptyError(msg)
        self.i += 1
        random_state = ensure_rng(random_state)
        if fit_gp:
            self._fit_gp(gp=gp, target_space=target_space)
        if self.previous_candidates is not None:
            self._update_gains(gp)
        x_max = [
            base_acq.suggest(gp=gp, target_space=target_space, n_random=n_random // self.n_acq,
                             n_smart=n_smart // self.n_acq, fit_gp=False, random_state=random_state)
            for base_acq in self.base_acquisitions
        ]
        self.previous_candidates = np.array(x_max)
        idx = self._sample_idx_from_softmax_gains(random_state=random_state)
        if not target_space._allow_duplicate_points and x_max[idx] in target_space:
            non_duplicate_idx = [i for i, x in enumerate(x_max) if x not in target_space]
            if len(non_duplicate_idx) > 0:
                cumsum_softmax_g = np.cumsum(softmax(self.gains[non_duplicate_idx]))
                r = random_state.rand()
                idx = non_duplicate_idx[np.argmax(r <= cumsum_softmax_g)]
        return x_max[idx]

    def get_acquisition_params(self):
        return {
            "base_acquisitions_params": [acq.get_acquisition_params() for acq in self.base_acquisitions],
            "gains": self.gains.tolist(),
            "previous_candidates": self.previous_candidates.tolist()
            if self.previous_candidates is not None else None,
        }

    def set_acquisition_params(self, params):
        for acq, acq_params in zip(self.base_acquisitions, params["base_acquisitions_params"]):
            acq.set_acquisition_params(acq_params)
        self.gains = np.array(params["gains"])
        self.previous_candidates = (
            np.array(params["previous_candidates"]) if params["previous_candidates"] is not None else None)
