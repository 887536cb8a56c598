"""Minimal host-side mirror of bayes_opt.target_space.TargetSpace (R/bayes_opt/target_space.py)
for float and int parameters: just what the acquisition hot path reads.

It exists so that the parity tests, bench.py and smoke() can drive the acquisition classes on
the GPU box, where the reference package is not installed.  With the reference installed, pass
its own TargetSpace - the acquisition classes only use the attributes mirrored here.
"""
from __future__ import annotations

import numpy as np

from .exception import NotUniqueError


def ensure_rng(random_state=None):
    """R/bayes_opt/util.py:8-30."""
    if random_state is None:
        return np.random.RandomState()
    if isinstance(random_state, int):
        return np.random.RandomState(random_state)
    if isinstance(random_state, np.random.RandomState):
        return random_state
    raise TypeError("random_state should be None, int or RandomState")


class TargetSpace:
    """Float-only TargetSpace: params (N,d) float64, target (N,), bounds (d,2)."""

    def __init__(self, target_func, pbounds, constraint=None, random_state=None,
                 allow_duplicate_points=False):
        self.target_func = target_func
        self._keys = list(pbounds.keys())  # R/bayes_opt/target_space.py:88 (insertion order)
        # (lo, hi) or (lo, hi, float) -> float parameter; (lo, hi, int) -> int parameter
        # (R/bayes_opt/target_space.py:252-262); categorical parameters are not mirrored
        self._is_int = np.array([len(pbounds[k]) == 3 and pbounds[k][2] is int for k in self._keys])
        for k in self._keys:
            b = pbounds[k]
            if not (len(b) == 2 or (len(b) == 3 and b[2] in (int, float))):
                raise NotImplementedError("only float and int parameters are mirrored here")
        self._bounds = np.array([pbounds[k][:2] for k in self._keys], dtype=float)
        self._dim = len(self._keys)
        self._params = np.empty((0, self._dim))
        self._target = np.empty((0,))
        self._cache = {}
        self._allow_duplicate_points = allow_duplicate_points or False
        self.n_duplicate_points = 0
        self._constraint = constraint  # a (B200)ConstraintModel or None
        if constraint is not None:
            if constraint.lb.size == 1:
                self._constraint_values = np.empty((0,), dtype=float)
            else:
                self._constraint_values = np.empty((0, constraint.lb.size), dtype=float)

    def __len__(self):
        return len(self._target)

    def __contains__(self, x):
        return tuple(float(v) for v in np.asarray(x, dtype=float).ravel()) in self._cache

    @property
    def empty(self):
        return len(self) == 0

    @property
    def params(self):
        return self._params

    @property
    def target(self):
        return self._target

    @property
    def dim(self):
        return self._dim

    @property
    def keys(self):
        return self._keys

    @property
    def bounds(self):
        return self._bounds

    @property
    def constraint(self):
        return self._constraint

    @property
    def constraint_values(self):
        if self._constraint is None:
            raise AttributeError("Available only if a constraint was passed.")
        return self._constraint_values

    @property
    def continuous_dimensions(self):
        return ~self._is_int

    def kernel_transform(self, value):
        """Identity for floats, np.round for ints (R/bayes_opt/parameter.py:222-234, :308-320)."""
        value = np.array(np.atleast_2d(value), dtype=float)
        if self._is_int.any():
            value[:, self._is_int] = np.round(value[:, self._is_int])
        return value

    def array_to_params(self, x):
        x = np.asarray(x, dtype=float)
        return {k: (int(np.round(v)) if i else float(v)) for k, v, i in zip(self._keys, x, self._is_int)}

    def params_to_array(self, params):
        return np.asarray([params[k] for k in self._keys], dtype=float)

    @property
    def mask(self):
        """R/bayes_opt/target_space.py:387-410."""
        mask = np.ones_like(self.target, dtype=bool)
        if self._constraint is not None:
            mask &= self._constraint.allowed(self._constraint_values)
        within = np.all((self._bounds[:, 0] <= self._params) & (self._params <= self._bounds[:, 1]), axis=1)
        mask &= within
        return mask

    def _as_array(self, x):
        if isinstance(x, dict):
            x = self.params_to_array(x)
        x = np.asarray(x, dtype=float).ravel()
        if x.size != self._dim:
            raise ValueError(f"Size of array ({x.size}) is different than the expected number of "
                             f"parameters ({self._dim}).")
        return x

    def register(self, params, target, constraint_value=None):
        """R/bayes_opt/target_space.py:424-518 (float parameters)."""
        x = self._as_array(params)
        key = tuple(float(v) for v in x)
        if key in self._cache:
            if self._allow_duplicate_points:
                self.n_duplicate_points += 1
            else:
                raise NotUniqueError(f"Data point {x} is not unique.")
        if self._constraint is None:
            self._cache[key] = target
        else:
            if constraint_value is None:
                raise ValueError("When registering a point to a constrained TargetSpace a constraint "
                                 "value needs to be present.")
            self._cache[key] = (target, constraint_value)
            self._constraint_values = np.concatenate(
                [self._constraint_values, np.atleast_1d(constraint_value).reshape((1,) + self._constraint_values.shape[1:])]
            )
        self._params = np.concatenate([self._params, x.reshape(1, -1)])
        self._target = np.concatenate([self._target, [target]])

    def probe(self, params):
        x = self._as_array(params)
        kw = self.array_to_params(x)
        target = self.target_func(**kw)
        if self._constraint is None:
            self.register(x, target)
            return target
        cv = self._constraint.eval(**kw)
        self.register(x, target, cv)
        return target, cv

    def random_sample(self, n_samples=0, random_state=None):
        """R/bayes_opt/target_space.py:565-603 + parameter.py:68-87: one RandomState.uniform call
        PER PARAMETER (column), in key order - the draw order the candidates' parity depends on."""
        random_state = ensure_rng(random_state)
        flatten = n_samples == 0
        n_samples = max(1, n_samples)
        data = np.empty((n_samples, self._dim))
        for j in range(self._dim):
            if self._is_int[j]:  # IntParameter.random_sample, R/bayes_opt/parameter.py:260-278
                data[:, j] = random_state.randint(int(self._bounds[j, 0]), int(self._bounds[j, 1]) + 1,
                                                  n_samples).astype(float)
            else:
                data[:, j] = random_state.uniform(self._bounds[j, 0], self._bounds[j, 1], n_samples)
        if flatten:
            return data.ravel()
        return data

    def _target_max(self):
        """R/bayes_opt/target_space.py:605-622."""
        if len(self.target) == 0:
            return None
        if len(self.target[self.mask]) == 0:
            return None
        return self.target[self.mask].max()

    def max(self):
        t = self._target_max()
        if t is None:
            return None
        idx = np.where(self.target == t)[0][0]
        return {"target": t, "params": self.array_to_params(self.params[idx])}

    def set_bounds(self, new_bounds):
        for j, k in enumerate(self._keys):
            if k in new_bounds:
                self._bounds[j] = new_bounds[k]
