"""Where the end-to-end time goes: the same 30 steps of C4 (steps 34-63) resident and through symgpu_step_traj_async: python tools/gpu_e2e_gap.py"""
import importlib, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("open-symuvia_b200")
synth = importlib.import_module("open-symuvia_b200.synth")
path = os.path.join(tempfile.gettempdir(), "gap_c4.symflat")
synth.fast_grid(path, 100, 100, 12.6, seed=42, route_len=24, n_steps=300, merges=True)
for mode in ("resident", "e2e", "resident", "e2e"):
    g = pkg.Engine(path)
    for _ in range(33):
        g.step()
    dev = 0.0
    t0 = time.perf_counter()
    if mode == "resident":
        for _ in range(30):
            g.step(); dev += g.last_step_ms
    else:
        cap = g.n + 65536
        bufs = [{"id": np.zeros(cap, np.int32), "lane": np.zeros(cap, np.int32), "pos": np.zeros(cap), "vit": np.zeros(cap), "acc": np.zeros(cap)} for _ in range(2)]
        for w in range(2):
            g.traj_bind(w, bufs[w])
        for k in range(30):
            g.step_traj_async(k % 2); dev += g.last_step_ms
            if k >= 1:
                g.traj_wait((k - 1) % 2)
        g.traj_wait(29 % 2)
    wall = (time.perf_counter() - t0) * 1e3
    print(f"{mode:9s} wall {wall / 30:.4f} ms/step, device (events around each step) {dev / 30:.4f} ms/step", flush=True)
    g.close()
