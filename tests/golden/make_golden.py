"""Regenerates the golden fixtures of tests/golden/ from the CPU oracle:  python tests/golden/make_golden.py

The reference itself cannot be built or run here (parity unpinned), so these vectors are *oracle* outputs,
frozen: they pin the oracle (and through it the CUDA engine) against drift, and give anyone with a SymuVia build concrete
numbers to compare with — vehicles present per step, the random-draw count, exit instants and the complete final state of
small seeded cases.  All cases use Newell (no libm dependence): the fp64 values are compared bit for bit."""
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
synth = importlib.import_module("open-symuvia_b200.synth")

CASES = {
    "c1_600": (lambda: synth.single_link(n_steps=600), 600),
    "c1_signal_400": (lambda: synth.single_link(n_steps=400, signal=(20, 3, 17), vit_reg=15.0, demand=0.4), 400),
    "grid4_demand_300": (lambda: synth.grid(n=4, n_steps=300, demand_per_lane=0.2, demand_duration=250), 300),
    "grid6_preseed_loops_120": (lambda: synth.grid(n=6, n_steps=120, demand_per_lane=0.0, preseed_per_lane=12.6, shuffle=True), 120),
    "corridor_lanechange_200": (lambda: synth.corridor(n_links=6, length=500.0, n_lanes=3, demand=1.2, offramp_every=2, preseed_density=0.03,
                                                       shuffle=True, n_steps=200), 200),
}


def run_case(name, variant="faithful"):
    from oracle.binding import Oracle
    make, steps = CASES[name]
    b = make()
    fd, path = tempfile.mkstemp(suffix=".symflat")
    os.close(fd)
    try:
        b.write(path)
        o = Oracle(path, len(b.links), len(b.classes), variant=variant)
        n_per_step = np.zeros(steps, np.int32)
        ex_id, ex_t = [], []
        for k in range(steps):
            n_per_step[k] = o.step()
            i, t = o.exited()
            ex_id.extend(int(x) for x in i)
            ex_t.extend(float(x) for x in t)
        st = o.state()
        ent, sor = o.link_counts()
        out = dict(n_per_step=n_per_step, rng_count=np.int64(o.rng_count), exit_id=np.asarray(ex_id, np.int32), exit_t=np.asarray(ex_t, np.float64),
                   link_entered=np.asarray(ent), link_left=np.asarray(sor), loops=np.int64(o.counter(1)),
                   **{k: np.asarray(st[k]) for k in ("id", "lane", "route_pos", "flags", "leader", "pos", "vit", "acc", "deltaN")})
        o.close()
        return out
    finally:
        os.unlink(path)


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        out = run_case(name)
        np.savez_compressed(os.path.join(here, name + ".npz"), **out)
        print(name, "steps", len(out["n_per_step"]), "vehicles at end", len(out["id"]), "exits", len(out["exit_id"]), "draws", int(out["rng_count"]))
