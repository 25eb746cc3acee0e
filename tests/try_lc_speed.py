"""Ad-hoc: lane-changing corridor at scale, GPU only: python tests/try_lc_speed.py <n_links> <steps> [seq]"""
import importlib
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
synth = importlib.import_module("open-symuvia_b200.synth")
pkg = importlib.import_module("open-symuvia_b200")
n_links = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
if len(sys.argv) > 3 and sys.argv[3] == "seq":
    os.environ["SYMGPU_LC"] = "seq"
b = synth.corridor(n_links=n_links, length=1000.0, n_lanes=3, demand=1.5, offramp_every=2, preseed_density=0.06, shuffle=True, n_steps=steps)
path = tempfile.mktemp(suffix=".symflat")
b.write(path)
g = pkg.Engine(path)
import ctypes as C
L = pkg.lib()
L.symgpu_profile.argtypes = [g.h.__class__, C.c_int]
L.symgpu_kernel_ms.argtypes = [g.h.__class__, C.c_int]
L.symgpu_kernel_ms.restype = C.c_double
L.symgpu_kernel_name.restype = C.c_char_p
L.symgpu_profile(g.h, 1)
for _ in range(5):
    n = g.step()
c0 = g.counter(pkg.CNT_LANE_CHANGES)
t0 = time.time()
for _ in range(steps):
    n = g.step()
dt = time.time() - t0
print(f"n={n} steps={steps} {dt / steps * 1e3:.2f} ms/step, lane changes/step {(g.counter(pkg.CNT_LANE_CHANGES) - c0) / steps:.1f}, "
      f"events walked/step {g.counter(pkg.CNT_LC_EVENTS) / (steps + 5):.0f}, records re-evaluated/step {g.counter(pkg.CNT_LC_REEVAL) / (steps + 5):.0f}")
print({L.symgpu_kernel_name(k).decode(): round(L.symgpu_kernel_ms(g.h, k) / (steps + 5), 3) for k in range(L.symgpu_kernel_count())})
os.unlink(path)
