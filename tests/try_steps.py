"""Bring-up: per-step device time on a pre-seeded grid: python tests/try_steps.py <n> <steps>"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("open-symuvia_b200")
synth = importlib.import_module("open-symuvia_b200.synth")
n = int(sys.argv[1]); steps = int(sys.argv[2])
b = synth.grid(n=n, n_steps=steps + 10, demand_per_lane=0.0, preseed_per_lane=12.6); b.write("/tmp/s.symflat")
g = pkg.Engine("/tmp/s.symflat")
ms = []
for k in range(steps):
    N = g.step(); ms.append(g.last_step_ms)
    if k < 6 or k % 10 == 0: print(k, N, f"{g.last_step_ms:.3f} ms", "loops", g.counter(pkg.CNT_LOOPS), "rounds", g.counter(pkg.CNT_PASSES))
print("mean", sum(ms)/len(ms), "min", min(ms), "last10", sum(ms[-10:])/10)
