"""The reference's OWN tests as the judge of the drop-in: test_acquisition.py, test_constraint.py,
test_bayesian_optimization.py, test_target_space.py, test_seq_domain_red.py, test_parameter.py, test_util.py and
test_logger.py of the reference (everything but the notebook runner) - vendored unmodified
into the git-ignored oracle/_ref/ref_tests - collected with tests/ref_suite_plugin.py, which rebinds
GaussianProcessRegressor, the bayes_opt.acquisition classes and ConstraintModel to the B200 classes.  Every
suggest()/maximize()/predict() of that suite then runs on the device; the assertions are the reference's."""
import os
import re

import pytest

from test_reference_suite_cpu import SUITE, run_reference_suite

pytestmark = pytest.mark.gpu


def test_reference_test_suite_passes_on_the_drop_in():
    if not os.path.isdir(SUITE):
        pytest.skip("reference tests not vendored (tools/vendor_ref.py needs /root/reference)")
    r = run_reference_suite(("--tb=short",))
    tail = r.stdout[-6000:]
    m = re.search(r"(\d+) passed", r.stdout)
    assert r.returncode == 0 and m, tail
    assert int(m.group(1)) == 167, tail  # the count the unmodified reference passes on the CPU (R/tests, 8 modules)
    assert "b200 drop-in: rebound" in r.stdout
    calls = dict(re.findall(r"(\w+\.\w+) x(\d+)", r.stdout))
    for name in ("B200GaussianProcessRegressor.fit", "B200GaussianProcessRegressor.predict",
                 "FusedAcquisition.argmin_topk", "FusedAcquisition.__call__"):
        assert int(calls[name]) > 0, (name, calls)
