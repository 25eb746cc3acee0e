"""Bring-up timing: python tests/try_speed.py <n> <steps>"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("open-symuvia_b200")
synth = importlib.import_module("open-symuvia_b200.synth")
n = int(sys.argv[1]); steps = int(sys.argv[2])
t0 = time.time(); b = synth.grid(n=n, n_steps=steps + 10, demand_per_lane=0.0, preseed_per_lane=12.6); b.write("/tmp/s.symflat"); print("gen", time.time() - t0)
g = pkg.Engine("/tmp/s.symflat")
g.run(3)
t0 = time.time(); N = g.run(steps); dt = time.time() - t0
print(f"n={N} steps={steps} wall={dt:.3f}s dev={g.last_step_ms:.1f}ms per-step={g.last_step_ms/steps:.3f}ms veh-steps/s={N*steps/dt:.3e} passes={g.counter(pkg.CNT_PASSES)} loops={g.counter(pkg.CNT_LOOPS)} launches={g.counter(pkg.CNT_LAUNCHES)}")
