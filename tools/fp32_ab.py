"""A/B of the fp32-mode kernels at C3 (d=16, N=4096, EI, 2^20 candidates resident in HBM): variant 2 (row-block
pairs x 128 candidates) vs variant 3 (N = 256 per MMA) vs variant 4 (N = 256, row-block pairs), interleaved so both see the same clocks / power state."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bayesianoptimization_b200 as bo
from bayesianoptimization_b200 import _lib as B
from sklearn.gaussian_process.kernels import Matern
rs = np.random.RandomState(0)
n, d, m = 4096, 16, 1 << 20
X = rs.uniform(size=(n, d)); y = np.sin(X.sum(1)) + 0.1 * rs.randn(n)
gp = bo.B200GaussianProcessRegressor(kernel=Matern(nu=2.5, length_scale=0.7), alpha=1e-6, normalize_y=True, optimizer=None, precision="fp32").fit(X, y)
acq = bo.FusedAcquisition(B.ACQ_EI, gp, xi=0.01, y_max=float(y.max()))
dev = torch.device("cuda", 0)
bufs = [torch.from_numpy(np.random.RandomState(100 + b).uniform(size=(m, d))).to(dev) for b in range(3)]
sel = torch.zeros((11, 2), dtype=torch.int64, device=dev)
L = B.lib(); stream = torch.cuda.current_stream()
out = {"2": [], "3": [], "4": []}; res = {}
for rep in range(6):
    for v in ("2", "3", "4"):
        os.environ["B200BO_TC_VARIANT"] = v
        B.check(L.b200bo_acq_eval_dev(C.byref(acq.spec), bufs[rep % 3].data_ptr(), m, None, None, None, 10, sel.data_ptr(), 0, stream.cuda_stream))
        ms = C.c_float(); B.check(L.b200bo_last_kernel_ms(C.byref(ms)))
        if rep >= 2:
            out[v].append(ms.value)
        res[v] = sel.cpu().numpy()[:, 1].tolist()
print(json.dumps({"kernel_ms": {k: float(np.mean(v)) for k, v in out.items()}, "all": out,
                  "cand_per_s": {k: m / (np.mean(v) * 1e-3) for k, v in out.items()},
                  "same_selection": res["2"] == res["3"] == res["4"], "sel2": res["2"][:4], "sel4": res["4"][:4]}))
