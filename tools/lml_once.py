"""One LML+gradient evaluation at N=4096, d=16 (for an ncu launch list of the fit-side kernels)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesianoptimization_b200 as bo
from sklearn.gaussian_process.kernels import Matern
rs = np.random.RandomState(0)
X = rs.uniform(size=(4096, 16)); y = np.sin(X.sum(1)) + 0.1 * rs.randn(4096)
gp = bo.B200GaussianProcessRegressor(kernel=Matern(nu=2.5, length_scale=0.7), alpha=1e-6, normalize_y=True, optimizer=None).fit(X, y)
print(gp.log_marginal_likelihood(np.log([0.9]), eval_gradient=True))
