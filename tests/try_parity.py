"""Ad-hoc driver used during bring-up: python tests/try_parity.py <case> <steps>"""
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.parity import run_parity  # noqa: E402

synth = importlib.import_module("open-symuvia_b200.synth")
case = sys.argv[1] if len(sys.argv) > 1 else "c1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
if case == "c1":
    b = synth.single_link(n_steps=steps)
elif case == "c1s":
    b = synth.single_link(n_steps=steps, signal=(20, 3, 17), vit_reg=15.0, demand=0.4)
elif case == "g4":
    b = synth.grid(n=4, n_steps=steps, demand_per_lane=0.2, demand_duration=400)
elif case == "g4hi":
    b = synth.grid(n=4, n_steps=steps, demand_per_lane=0.45, demand_duration=600)
elif case == "g6seed":
    b = synth.grid(n=6, n_steps=steps, demand_per_lane=0.0, preseed_per_lane=12.6, shuffle=True)
elif case == "g6both":
    b = synth.grid(n=6, n_steps=steps, demand_per_lane=0.3, preseed_per_lane=8, shuffle=True)
elif case == "c4":
    b = synth.config("C4")
elif case == "g30":
    b = synth.grid(n=30, n_steps=steps, demand_per_lane=0.0, preseed_per_lane=12.6, shuffle=False)
elif case.startswith("lc:"):
    from tests.test_gpu_lanechange import CASES
    kw = dict(CASES[case.split(":")[1]])
    if case.endswith(":seq"):
        os.environ["SYMGPU_LC"] = "seq"
    b = synth.corridor(**kw)
else:
    raise SystemExit("unknown case")
t0 = time.time()
print(case, run_parity(b, steps, verbose=True), f"{time.time() - t0:.1f}s")
