"""Ad-hoc GPU bring-up driver for the merge points: runs small grids through tests.parity and prints the first difference."""
import importlib, os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
synth = importlib.import_module("open-symuvia_b200.synth")
from tests.parity import run_parity
cases = [
    ("grid3 light", lambda: synth.grid(n=3, n_steps=200, demand_per_lane=0.15, demand_duration=150), 200),
    ("grid4 demand", lambda: synth.grid(n=4, n_steps=500, demand_per_lane=0.2, demand_duration=400), 500),
    ("grid4 saturated", lambda: synth.grid(n=4, n_steps=400, demand_per_lane=0.45, demand_duration=600), 400),
    ("grid6 preseed", lambda: synth.grid(n=6, n_steps=300, demand_per_lane=0.0, preseed_per_lane=12.6, shuffle=True), 300),
    ("grid4 nomerge", lambda: synth.grid(n=4, n_steps=300, demand_per_lane=0.3, demand_duration=400, merges=False), 300),
    ("c1", lambda: synth.single_link(n_steps=300), 300),
]
only = sys.argv[1:] 
for name, mk, n in cases:
    if only and not any(o in name for o in only):
        continue
    try:
        s = run_parity(mk(), n, check_every=1)
        print("OK  ", name, s, flush=True)
    except Exception as e:
        print("FAIL", name, str(e)[:600], flush=True)
