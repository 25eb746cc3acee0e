import importlib, os, sys, tempfile
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests/golden")
import make_golden
pkg = importlib.import_module("open-symuvia_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "c1_600"
make, steps = make_golden.CASES[name]
b = make(); path = "/tmp/dbg.symflat"; b.write(path)
g = pkg.Engine(path)
for k in range(steps):
    try:
        n = g.step()
    except Exception as e:
        print("step", k, "failed:", e); break
else:
    print("ok", steps)
