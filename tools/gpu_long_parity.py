"""Long oracle parity run of a scope config: python tools/gpu_long_parity.py C4 300 25 [ties]   (ties: leader ids may differ among vehicles standing at one place)"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
synth = importlib.import_module("open-symuvia_b200.synth")
from tests.parity import run_parity
name, steps, every = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
t0 = time.time()
s = run_parity(synth.config(name), steps, check_every=every, verbose=True, leader_ties_ok="ties" in sys.argv[4:], check_all_from=int(os.environ.get("CHECK_ALL_FROM", "0")) or None)
print(name, steps, "steps:", s, "in %.0f s" % (time.time() - t0))
