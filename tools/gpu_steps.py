"""Runs a few steps of a pre-seeded grid (for ncu launch lists): python tools/gpu_steps.py NX NY STEPS [merges=1]"""
import importlib, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("open-symuvia_b200")
synth = importlib.import_module("open-symuvia_b200.synth")
nx, ny, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
merges = int(sys.argv[4]) if len(sys.argv) > 4 else 1
path = os.path.join(tempfile.gettempdir(), f"steps_{nx}x{ny}_{merges}.symflat")
n, tot = synth.fast_grid(path, nx, ny, 12.6, n_steps=steps + 10, merges=bool(merges))
g = pkg.Engine(path)
t0 = time.time()
for k in range(steps):
    g.step()
print("vehicles", g.n, "steps", steps, "last_step_ms", g.last_step_ms, "wall", time.time() - t0,
      {k: g.counter(getattr(pkg, k)) for k in ("CNT_MRG_INS", "CNT_MRG_REFUS", "CNT_MRG_SERIAL", "CNT_MRG_VALUES", "CNT_MRG_DRAWS", "CNT_PASSES", "CNT_LOOPS", "CNT_TIES")}, "rng", g.rng_count)
