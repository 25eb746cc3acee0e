"""Ad-hoc: partitioned runs (all ranks in one process, LocalGroup) against the oracle's partitioned mode over random grids:
python tools/fuzz_partition.py [n_cases] [first_seed]"""
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
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 10
first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bad = 0
for seed in range(first, first + n_cases):
    kw = fuzz_cases.grid_case(seed, n_steps=150)
    kw["n"] = max(kw["n"], 4)
    world = 2 + seed % 3
    kw["n"] = max(kw["n"], world)
    b = fuzz_cases.synth.grid(**kw)
    path = tempfile.mktemp(suffix=".symflat")
    b.write(path)
    try:
        node_part = part_mod.strip_partition(np.asarray(b.node_xy), world)
        if len(set(node_part.tolist())) != world:
            print(seed, "skip (degenerate partition)"); continue
        o = Oracle(path, len(b.links), len(b.classes)); o.set_partition(node_part)
        g = part_mod.LocalGroup(pkg, path, node_part, world)
        mig = 0
        for k in range(150):
            no, ng = o.step(), g.step()
            assert no == ng, f"step {k + 1}: n {no} vs {ng}"
            so, sg = o.state(), g.state()
            order = np.argsort(so["id"], kind="stable")
            for key in ("id", "lane", "route_pos", "flags", "leader", "pos", "vit", "acc", "deltaN"):
                assert np.array_equal(so[key][order], sg[key]), f"step {k + 1}: {key}"
            assert o.rng_count == g.rng_count(), f"step {k + 1}: rng {o.rng_count} vs {g.rng_count()}"
            mig += sum(int(r.send_counts.sum()) for r in g.ranks)
        print(seed, "ok world", world, "n", no, "migrated", mig, flush=True)
        g.close(); o.close()
    except AssertionError as e:
        bad += 1
        print(seed, "FAIL world", world, str(e)[:200], {k: kw[k] for k in ("n", "ny", "n_lanes", "demand_per_lane", "preseed_per_lane", "shuffle")}, flush=True)
    finally:
        os.unlink(path)
print("failures:", bad)
