"""LML + gradient wall time at N in {1024, 4096} with the factorisation issued as a CUDA graph vs launch by launch."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesianoptimization_b200 as bo
from sklearn.gaussian_process.kernels import Matern
out = {}
for n, d in [(1024, 8), (4096, 16)]:
    rs = np.random.RandomState(0)
    X = rs.uniform(size=(n, d)); y = np.sin(X.sum(1)) + 0.1 * rs.randn(n)
    for mode in ("graph", "graph_trail_gemm128", "direct", "direct_serial_gemm64"):
        os.environ["B200BO_GRAPH"] = "1" if mode.startswith("graph") else "0"
        os.environ["B200BO_TRAIL"] = "gemm" if mode == "graph_trail_gemm128" else "tile64"
        if mode == "direct_serial_gemm64":
            os.environ["B200BO_POTRF"], os.environ["B200BO_GEMM"] = "serial", "64"
        else:
            os.environ.pop("B200BO_POTRF", None); os.environ.pop("B200BO_GEMM", None)
        gp = bo.B200GaussianProcessRegressor(kernel=Matern(nu=2.5, length_scale=0.7), alpha=1e-6, normalize_y=True, optimizer=None).fit(X, y)
        th = np.log([0.9])
        gp.log_marginal_likelihood(th, eval_gradient=True)
        ts = []
        for _ in range(10):
            t0 = time.perf_counter(); v = gp.log_marginal_likelihood(th, eval_gradient=True); ts.append(time.perf_counter() - t0)
        out[f"n{n}_{mode}"] = {"lml_grad_ms_median": 1e3 * float(np.median(ts)), "min": 1e3 * float(np.min(ts)), "lml": float(v[0]), "grad": float(v[1][0])}
print(json.dumps(out))
