"""Two engines on the same C4 input with different SYMGPU_SIDE masks, compared every step: python tools/gpu_diff_modes.py STEPS MASK_A MASK_B"""
import importlib, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("open-symuvia_b200")
synth = importlib.import_module("open-symuvia_b200.synth")
steps, ma, mb = int(sys.argv[1]), sys.argv[2], sys.argv[3]
path = os.path.join(tempfile.gettempdir(), "diff_c4.symflat")
synth.fast_grid(path, 100, 100, 12.6, n_steps=steps + 10, merges=True)
os.environ["SYMGPU_SIDE"] = ma
a = pkg.Engine(path)
os.environ["SYMGPU_SIDE"] = mb
b = pkg.Engine(path)
for k in range(steps):
    a.step(); b.step()
    sa, sb = a.state(), b.state()
    if a.n != b.n or any(not np.array_equal(sa[key], sb[key]) for key in ("id", "lane", "pos", "vit")):
        print("first difference at step", k + 1, "n", a.n, b.n, "rng", a.rng_count, b.rng_count)
        if a.n == b.n:
            bad = np.nonzero((sa["lane"] != sb["lane"]) | (sa["pos"] != sb["pos"]) | (sa["vit"] != sb["vit"]))[0]
            print(len(bad), "vehicles differ")
            for i in bad[:12]:
                print(" id", sa["id"][i], "| A lane", sa["lane"][i], "pos %.6f vit %.6f" % (sa["pos"][i], sa["vit"][i]), "| B lane", sb["lane"][i], "pos %.6f vit %.6f" % (sb["pos"][i], sb["vit"][i]), "leader", sa["leader"][i], sb["leader"][i])
        for name, xa, xb in zip(("last passage", "rand numbers"), a.merge_state(), b.merge_state()):
            d = np.nonzero(xa != xb)[0]
            if len(d): print(" merge state", name, "differs at points", d[:10], xa[d[:10]], xb[d[:10]])
        break
else:
    print("no difference in", steps, "steps")
