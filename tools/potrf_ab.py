"""A/B of the diagonal-block Cholesky kernels: seconds per LML+gradient evaluation (the objective of
the hyper-parameter search) with the blocked kernel (default) and B200BO_POTRF=legacy."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesianoptimization_b200 as bo  # noqa: E402
from sklearn.gaussian_process.kernels import Matern  # noqa: E402

out = {}
for n, d in [(1024, 8), (4096, 16)]:
    rs = np.random.RandomState(0)
    X = rs.uniform(size=(n, d))
    y = np.sin(X.sum(1)) + 0.1 * rs.randn(n)
    gp = bo.B200GaussianProcessRegressor(kernel=Matern(nu=2.5, length_scale=0.7), alpha=1e-6, normalize_y=True,
                                         optimizer=None).fit(X, y)
    for mode in ("blocked", "legacy", "blocked"):
        if mode == "legacy":
            os.environ["B200BO_POTRF"] = "legacy"
        else:
            os.environ.pop("B200BO_POTRF", None)
        vals = [gp.log_marginal_likelihood(np.log([0.9]), eval_gradient=True) for _ in range(2)]
        t0 = time.perf_counter()
        for _ in range(8):
            v = gp.log_marginal_likelihood(np.log([0.9]), eval_gradient=True)
        dt = (time.perf_counter() - t0) / 8
        out.setdefault(f"n{n}_d{d}", {})[mode] = {"lml_grad_eval_ms": 1e3 * dt, "lml": float(v[0]), "grad": float(v[1][0])}
print(json.dumps(out))
