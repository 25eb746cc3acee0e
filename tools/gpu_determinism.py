"""State digests of a bench workload every K steps (to compare runs under different launch timings): python tools/gpu_determinism.py STEPS EVERY [C4|C3]"""
import hashlib, importlib, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("open-symuvia_b200")
synth = importlib.import_module("open-symuvia_b200.synth")
steps, every = int(sys.argv[1]), int(sys.argv[2])
wl = sys.argv[3] if len(sys.argv) > 3 else "C4"
path = os.path.join(tempfile.gettempdir(), f"det_{wl}.symflat")
if wl == "C4":
    synth.fast_grid(path, 100, 100, 12.6, n_steps=steps + 10, merges=True)
else:
    import bench
    bench.make_workload(wl, 0, 1, path)
g = pkg.Engine(path)
for k in range(steps):
    g.step()
    if (k + 1) % every == 0:
        st = g.state()
        h = hashlib.sha1()
        for key in ("id", "lane", "pos", "vit"):
            h.update(st[key].tobytes())
        print(k + 1, g.n, g.rng_count, g.counter(pkg.CNT_MRG_INS), g.counter(pkg.CNT_MRG_REFUS), g.counter(pkg.CNT_MRG_SERIAL), h.hexdigest()[:16], flush=True)
