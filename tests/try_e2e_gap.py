"""Ad-hoc: where the end-to-end loop loses time against symgpu_run (C4).  python tests/try_e2e_gap.py"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("open-symuvia_b200")
synth = importlib.import_module("open-symuvia_b200.synth")
path = "/tmp/e2e_gap.symflat"
synth.fast_grid(path, 100, 100, per_lane=12.6, seed=42, route_len=24, n_steps=400)
g = pkg.Engine(path)
g.run(10)
K = 40
t0 = time.perf_counter(); g.run(K); t1 = time.perf_counter()
print(f"symgpu_run      : {1e3 * (t1 - t0) / K:.4f} ms/step wall, {g.last_step_ms / K:.4f} device")
t0 = time.perf_counter()
for _ in range(K):
    g.step()
t1 = time.perf_counter()
print(f"python g.step() : {1e3 * (t1 - t0) / K:.4f} ms/step wall")
cap = g.n + 65536
bufs = [{"id": np.zeros(cap, np.int32), "lane": np.zeros(cap, np.int32), "pos": np.zeros(cap), "vit": np.zeros(cap), "acc": np.zeros(cap)} for _ in range(2)]
g.traj_bind(0, bufs[0]); g.traj_bind(1, bufs[1])
t0 = time.perf_counter()
for k in range(K):
    w = k & 1
    g.step_traj_async(w)
    if k > 0:
        g.traj_wait(w ^ 1)
g.traj_wait((K - 1) & 1)
t1 = time.perf_counter()
print(f"traj pipelined  : {1e3 * (t1 - t0) / K:.4f} ms/step wall")
t0 = time.perf_counter()
for k in range(K):
    g.step()
    if k % 2 == 0:
        g.traj_copy_async(0); g.traj_wait(0)
t1 = time.perf_counter()
print(f"traj every 2nd, waited at once: {1e3 * (t1 - t0) / K:.4f} ms/step wall")
g.close()
