#!/usr/bin/env python
"""BASELINE configs[0] through the reference's driver, stock vs enable()d (the bench.py leg, stand-alone)."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)
import bayesianoptimization_b200 as bo  # noqa: E402

print(json.dumps(bench.c1_live_leg(bo, 0)))
