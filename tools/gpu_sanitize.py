"""Small merge + lane-change scenarios for compute-sanitizer: compute-sanitizer --tool memcheck|racecheck|synccheck python tools/gpu_sanitize.py"""
import importlib, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("open-symuvia_b200")
synth = importlib.import_module("open-symuvia_b200.synth")
cases = [("grid", synth.grid(n=5, n_steps=80, demand_per_lane=0.45, demand_duration=600, preseed_per_lane=10, shuffle=True), 60),
         ("corridor", synth.corridor(n_links=8, length=400.0, n_lanes=3, demand=0.5, offramp_every=2, onramp_every=3, onramp_demand=0.25, preseed_density=0.08, shuffle=True, n_steps=80), 60)]
for name, b, steps in cases:
    fd, path = tempfile.mkstemp(suffix=".symflat"); os.close(fd)
    b.write(path)
    g = pkg.Engine(path)
    os.unlink(path)
    for _ in range(steps):
        g.step()
    print(name, "vehicles", g.n, "merge ins", g.counter(pkg.CNT_MRG_INS), "serial", g.counter(pkg.CNT_MRG_SERIAL), "lc events", g.counter(pkg.CNT_LC_EVENTS), flush=True)
    g.close()
