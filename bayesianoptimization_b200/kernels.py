"""``wrap_kernel``: mirror of R/bayes_opt/parameter.py:457-495 - a sklearn kernel whose
``__call__`` first applies an input transform (identity for floats, np.round for ints).  The
reference captures ``transform`` in the closure of the generated class (the ``_transform``
attribute it also sets does NOT survive ``sklearn.base.clone``), so ``find_transform`` looks in
both places."""
from __future__ import annotations

from inspect import signature

from sklearn.gaussian_process import kernels


def wrap_kernel(kernel: kernels.Kernel, transform):
    kernel_type = type(kernel)

    class WrappedKernel(kernel_type):
        def __init__(self, **kwargs):
            super().__init__(**kwargs)

        def __call__(self, X, Y=None, eval_gradient=False):
            X = transform(X)
            Y = transform(Y) if Y is not None else None
            return super().__call__(X, Y, eval_gradient)

        def __reduce__(self):
            return (wrap_kernel, (kernel, transform))

    WrappedKernel.__init__.__signature__ = signature(
        getattr(kernel_type.__init__, "deprecated_original", kernel_type.__init__))
    wrapped = WrappedKernel.__new__(WrappedKernel)
    wrapped.__dict__.update(kernel.__dict__)
    wrapped._transform = transform
    return wrapped


def find_transform(kernel):
    """The input transform of a wrapped kernel (this module's or the reference's), else None."""
    t = getattr(kernel, "_transform", None)
    if t is not None:
        return t
    call = type(kernel).__dict__.get("__call__")
    code = getattr(call, "__code__", None)
    if code is not None and call.__closure__:
        for name, cell in zip(code.co_freevars, call.__closure__):
            if name == "transform":
                return cell.cell_contents
    if type(kernel).__name__ == "WrappedKernel":
        raise NotImplementedError("wrapped kernel whose transform cannot be located")
    return None
