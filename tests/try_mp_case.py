"""Ad-hoc: one partitioned fuzz case with a diagnosis of the first difference: python tests/try_mp_case.py <seed> <world>"""
import importlib
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import fuzz_cases  # noqa: E402
from oracle.binding import Oracle  # noqa: E402

pkg = importlib.import_module("open-symuvia_b200")
part_mod = importlib.import_module("open-symuvia_b200.partition")
seed, world = int(sys.argv[1]), int(sys.argv[2])
kw = fuzz_cases.grid_case(seed, n_steps=150)
kw["n"] = max(kw["n"], 4, world)
b = fuzz_cases.synth.grid(**kw)
path = tempfile.mktemp(suffix=".symflat")
b.write(path)
node_part = part_mod.strip_partition(np.asarray(b.node_xy), world)
o = Oracle(path, len(b.links), len(b.classes)); o.set_partition(node_part)
g = part_mod.LocalGroup(pkg, path, node_part, world)
prev = None
for k in range(150):
    no, ng = o.step(), g.step()
    so, sg = o.state(), g.state()
    order = np.argsort(so["id"], kind="stable")
    so = {key: v[order] for key, v in so.items()}
    diff = [key for key in ("id", "lane", "route_pos", "flags", "leader", "pos", "vit", "acc", "deltaN") if not np.array_equal(so[key], sg[key])]
    if diff:
        print("step", k + 1, "differs in", diff, "ties oracle", o.counter(0))
        i = np.nonzero(so["leader"] != sg["leader"])[0]
        for r in i[:3]:
            print(" vehicle id", so["id"][r], "lane", so["lane"][r], "pos", so["pos"][r], "oracle leader", so["leader"][r], "gpu leader", sg["leader"][r])
            for lid in (so["leader"][r], sg["leader"][r]):
                j = np.nonzero(prev["id"] == lid)[0]
                if len(j):
                    print("   candidate", lid, "at start of step: lane", prev["lane"][j[0]], "pos", repr(prev["pos"][j[0]]), "vit", prev["vit"][j[0]])
        break
    prev = so
else:
    print("no difference in 150 steps")
os.unlink(path)
