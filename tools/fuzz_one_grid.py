"""One case of the grid campaign with every differing row printed: python tools/fuzz_one_grid.py [seed]   (3036: the one known difference)"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import fuzz_cases
from tests.parity import run_parity
synth = importlib.import_module("open-symuvia_b200.synth")
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 3036
kw = fuzz_cases.grid_case(seed, n_steps=200)
print(kw)
try:
    print(run_parity(synth.grid(**kw), 200, verbose=True))
except AssertionError as e:
    print("FAIL", str(e)[:400])
