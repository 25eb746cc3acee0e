"""Ad-hoc: lane-change parity over random corridor shapes and seeds:  python tools/fuzz_lane_change.py [n_cases] [first_seed]"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import fuzz_cases  # noqa: E402
from tests.parity import run_parity  # noqa: E402

synth = importlib.import_module("open-symuvia_b200.synth")
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 10
first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bad = 0
for seed in range(first, first + n_cases):
    kw = fuzz_cases.corridor_case(seed, n_steps=250)
    try:
        st = run_parity(synth.corridor(**kw), 250)
        ok = st["worst"] == 0.0 and st["overflow"] == 0
        print(seed, "ok" if ok else "DIFF", st["worst"], "n", st["n_final"], "ties", st["ties_oracle"], flush=True)
        bad += 0 if ok else 1
    except AssertionError as e:
        bad += 1
        print(seed, "FAIL", str(e)[:200], kw, flush=True)
print("failures:", bad)
