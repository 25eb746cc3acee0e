"""Ad-hoc: parity over random small grids (signals, demand, pre-seeded population, shuffled list order, 1-2 vehicle types):
python tools/fuzz_grid.py [n_cases] [first_seed]"""
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
    kw = fuzz_cases.grid_case(seed, n_steps=200)
    try:
        st = run_parity(synth.grid(**kw), 200)
        ok = st["worst"] == 0.0 and st["overflow"] == 0 and st["loops_gpu"] == st["loops_oracle"]
        print(seed, "ok" if ok else "DIFF", st["worst"], "n", st["n_final"], "loops", st["loops_gpu"], "ties", st["ties_oracle"], flush=True)
        bad += 0 if ok else 1
    except AssertionError as e:
        bad += 1
        print(seed, "FAIL", str(e)[:220], {k: kw[k] for k in ("n", "ny", "n_lanes", "demand_per_lane", "preseed_per_lane", "shuffle")}, flush=True)
print("failures:", bad)
