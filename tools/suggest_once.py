"""One acquisition-side suggest() without refit at N=4096, d=16 (10 000 random candidates + 10 L-BFGS-B
refinements in lockstep) - for an ncu launch list of the small-batch kernels."""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import bayesianoptimization_b200 as bo
from bayes_opt.target_space import TargetSpace
from sklearn.gaussian_process.kernels import Matern
warnings.simplefilter("ignore")
n, d = int(os.environ.get("N", 4096)), 16
rs = np.random.RandomState(0)
X = rs.uniform(size=(n, d)); y = np.sin(X.sum(1)) + 0.1 * rs.randn(n)
gp = bo.B200GaussianProcessRegressor(kernel=Matern(nu=2.5, length_scale=1.1), alpha=1e-6, normalize_y=True, optimizer=None).fit(X, y)
space = TargetSpace(None, {f"x{i:02d}": (0.0, 1.0) for i in range(d)})
for a, b in zip(X[:3], y[:3]):
    space.register(a, b)
ei = bo.ExpectedImprovement(xi=0.01); ei.y_max = float(y.max())
for rep in range(int(os.environ.get("REPS", 2))):
    t0 = time.perf_counter(); l0 = bo._lib.lib().b200bo_launch_count()
    acq = ei._get_acq(gp=gp)
    x = ei._acq_min(acq, space, random_state=np.random.RandomState(5), n_random=10_000, n_smart=10)
    print("suggest_nofit_s", time.perf_counter() - t0, "launches", bo._lib.lib().b200bo_launch_count() - l0)
