"""The reference is imported from its source tree (not pip-installed), so
importlib.metadata.version("bayesian-optimization") (R/bayes_opt/__init__.py:14) needs an answer."""
import importlib.metadata as _m

_orig = _m.version


def _version(name):
    if name == "bayesian-optimization":
        return "3.3.0"
    return _orig(name)


_m.version = _version
