"""Step time series of a bench workload: python tools/gpu_series.py C4|C3 STEPS [EVERY]   (device ms per step and the merge / lane-change counters)"""
import importlib, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("open-symuvia_b200")
name, steps = sys.argv[1], int(sys.argv[2])
every = int(sys.argv[3]) if len(sys.argv) > 3 else 10
path = os.path.join(tempfile.gettempdir(), f"series_{name}.symflat")
info = bench.make_workload(name, 0, 1, path)
g = pkg.Engine(path)
names = ("CNT_MRG_INS", "CNT_MRG_REFUS", "CNT_MRG_SERIAL", "CNT_MRG_VALUES", "CNT_MRG_DRAWS", "CNT_PASSES", "CNT_LC_EVENTS", "CNT_LC_REEVAL")
prev = {k: 0 for k in names}
acc = 0.0
for k in range(steps):
    g.step()
    acc += g.last_step_ms
    if (k + 1) % every == 0:
        cur = {n: g.counter(getattr(pkg, n)) for n in names if hasattr(pkg, n)}
        print(f"step {k + 1:5d} n {g.n:8d} avg_ms {acc / every:8.3f} last_ms {g.last_step_ms:8.3f}", {n[4:]: (cur[n] - prev.get(n, 0)) // every for n in cur}, flush=True)
        prev = cur
        acc = 0.0
