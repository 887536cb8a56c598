"""B200 ConstraintModel: one device GP per constraint component
(mirror of R/bayes_opt/constraint.py:16-263)."""
from __future__ import annotations

import numpy as np
from scipy.special import ndtr
from sklearn.gaussian_process.kernels import Matern

from .gpr import B200GaussianProcessRegressor
from .kernels import wrap_kernel


def _wrap(kernel, transform):
    return kernel if transform is None else wrap_kernel(kernel, transform)


class ConstraintModel:
    """GP models of the constraint function(s); ``predict`` = probability of fulfilment."""

    def __init__(self, fun, lb, ub, transform=None, random_state=None, device=0):
        self.fun = fun
        self._lb = np.atleast_1d(lb).astype(float)
        self._ub = np.atleast_1d(ub).astype(float)
        if np.any(self._lb >= self._ub):
            raise ValueError("Lower bounds must be less than upper bounds.")
        # same GP configuration as R/bayes_opt/constraint.py:72-81
        self._model = [
            B200GaussianProcessRegressor(
                kernel=_wrap(Matern(nu=2.5), transform), alpha=1e-6, normalize_y=True,
                n_restarts_optimizer=5, random_state=random_state, device=device)
            for _ in range(len(self._lb))
        ]

    @property
    def lb(self):
        return self._lb

    @property
    def ub(self):
        return self._ub

    @property
    def model(self):
        return self._model

    def eval(self, **kwargs):
        if self.fun is None:
            raise ValueError("No constraint function was provided.")
        try:
            return self.fun(**kwargs)
        except TypeError as e:
            msg = ("Encountered TypeError when evaluating constraint function. This could be because "
                   "your constraint function doesn't use the same keyword arguments as the target "
                   f"function. Original error message:\n\n{e}")
            e.args = (msg,)
            raise

    def fit(self, X, Y):
        """R/bayes_opt/constraint.py:132-151."""
        if len(self._model) == 1:
            self._model[0].fit(X, Y)
        else:
            for i, gp in enumerate(self._model):
                gp.fit(X, Y[:, i])

    def predict(self, X):
        """R/bayes_opt/constraint.py:153-221.  Stand-alone use only: inside the acquisition closure
        the same product is evaluated in the fused kernel (acquisition.FusedAcquisition)."""
        X = np.asarray(X, dtype=float)
        X_shape = X.shape
        X = X.reshape((-1, self._model[0].n_features_in_))
        result = None
        with np.errstate(divide="ignore", invalid="ignore"):
            for j, gp in enumerate(self._model):
                mu, sd = gp.predict(X, return_std=True)
                ok = sd > 0
                p_lo = np.where(ok, ndtr((self._lb[j] - mu) / sd), np.nan) if self._lb[j] != -np.inf else 0.0
                p_hi = np.where(ok, ndtr((self._ub[j] - mu) / sd), np.nan) if self._ub[j] != np.inf else 1.0
                term = p_hi - p_lo
                result = term if result is None else result * term
        return np.asarray(result).reshape(X_shape[:-1])

    def approx(self, X):
        """R/bayes_opt/constraint.py:223-243."""
        X = np.asarray(X, dtype=float)
        X_shape = X.shape
        X = X.reshape((-1, self._model[0].n_features_in_))
        if len(self._model) == 1:
            return self._model[0].predict(X).reshape(X_shape[:-1])
        result = np.column_stack([gp.predict(X) for gp in self._model])
        return result.reshape(*X_shape[:-1], len(self._lb))

    def allowed(self, constraint_values):
        """R/bayes_opt/constraint.py:245-263."""
        if self._lb.size == 1:
            return np.less_equal(self._lb, constraint_values) & np.less_equal(constraint_values, self._ub)
        return np.all(constraint_values <= self._ub, axis=-1) & np.all(constraint_values >= self._lb, axis=-1)
