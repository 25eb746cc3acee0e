"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle): the oracle must still reproduce
them (CPU test), and the CUDA engine must reproduce them through the C-ABI without the oracle in the loop (GPU test)."""
import importlib
import os
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

INT_KEYS = ("id", "lane", "route_pos", "flags", "leader")
F64_KEYS = ("pos", "vit", "acc", "deltaN")


def _load(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


@pytest.mark.parametrize("variant", ["faithful", "flat"])          # flat: the "sorted lanes" CPU baseline of bench.py (flat per-lane arrays), same results
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_reproduces_golden(name, variant):
    want, got = _load(name), make_golden.run_case(name, variant)
    assert sorted(want) == sorted(got)
    for k in want:
        assert np.array_equal(np.asarray(want[k]), np.asarray(got[k])), k       # integers and fp64 bit for bit


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_engine_reproduces_golden(name):
    pkg = importlib.import_module("open-symuvia_b200")
    want = _load(name)
    make, steps = make_golden.CASES[name]
    b = make()
    fd, path = tempfile.mkstemp(suffix=".symflat")
    os.close(fd)
    try:
        b.write(path)
        g = pkg.Engine(path)
        n_per_step = np.zeros(steps, np.int32)
        ex_id, ex_t = [], []
        for k in range(steps):
            n_per_step[k] = g.step()
            i, t = g.exited()
            ex_id.extend(int(x) for x in i)
            ex_t.extend(float(x) for x in t)
        st = g.state()
        ent, sor = g.link_counts()
        assert np.array_equal(n_per_step, want["n_per_step"])
        assert g.rng_count == int(want["rng_count"])
        assert np.array_equal(np.asarray(ex_id, np.int32), want["exit_id"])
        assert np.array_equal(np.asarray(ex_t, np.float64), want["exit_t"])
        assert np.array_equal(np.asarray(ent), want["link_entered"]) and np.array_equal(np.asarray(sor), want["link_left"])
        assert g.counter(pkg.CNT_LOOPS) == int(want["loops"])
        for k in INT_KEYS + F64_KEYS:
            assert np.array_equal(np.asarray(st[k]), want[k]), k
        g.close()
    finally:
        os.unlink(path)
