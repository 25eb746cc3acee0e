"""Parity of the CUDA path (through the symgpu C-ABI) against the CPU oracle — needs a B200.

Integer state (list order, ids, lanes, route cursor, regime flags, leader ids, link counters, exits, RNG count)
must be bit-exact; fp64 state within 1e-9 relative (BASELINE.json north_star).  In practice the device arithmetic
is contraction-free and ordered like the reference, so the tests also assert bit-exact fp64 (`worst == 0`).
"""
import importlib

import numpy as np
import pytest

from tests.parity import run_parity

pytestmark = pytest.mark.gpu


def _check(stats, expect_loops=None):
    assert stats["overflow"] == 0
    assert stats["loops_gpu"] == stats["loops_oracle"]
    assert stats["worst"] == 0.0, stats
    if expect_loops:
        assert stats["loops_gpu"] > 0


def test_c1_single_link_full_hour(synth):
    """BASELINE config C1: 2 km one-lane link, Newell, 1800 veh/h, 3600 steps of 1 s."""
    s = run_parity(synth.config("C1"), 3600, check_every=1)
    _check(s)
    assert s["veh_steps"] > 100000


def test_signalised_link(synth):
    _check(run_parity(synth.single_link(n_steps=600, signal=(20, 3, 17), vit_reg=15.0, demand=0.4), 600))


def test_saturated_origin_queue(synth):
    """Origin above capacity: ready / waiting vehicle logic (usage/SymuViaTripNode.cpp:1642-1851)."""
    s = run_parity(synth.single_link(length=1500.0, vit_reg=10.0, demand=0.9, n_steps=300), 300)
    _check(s)


def test_deceleration_at_red(synth):
    cls = dict(synth.VL, decel=2.5)
    _check(run_parity(synth.single_link(n_steps=400, signal=(15, 3, 22), vit_reg=18.0, demand=0.35, classes=[cls]), 400))


def test_exit_capacity(synth):
    _check(run_parity(synth.single_link(length=800.0, n_steps=400, vit_reg=20.0, demand=0.5, exit_capacity=0.3), 400))


def test_accbornee(synth):
    _check(run_parity(synth.single_link(length=800.0, n_steps=300, vit_reg=20.0, demand=0.4, accbornee=True,
                                        signal=(20, 3, 17)), 300))


@pytest.mark.parametrize("pow_mode", ["libm", "cr"])
@pytest.mark.parametrize("law", [1, 2])
def test_gipps_idm(synth, law, pow_mode, monkeypatch):
    """The device evaluates the laws' three powers (x^0.5, x^4, x^1.5) correctly rounded.  Against the oracle with the host libm's
    std::pow (glibc: within 0.52 ulp, differs from the correctly rounded value for ~0.08 % of the arguments): 1e-9 relative,
    integers exact.  Against the oracle with the same correctly rounded powers (ORACLE_POW=cr): bit-exact."""
    if pow_mode == "cr":
        monkeypatch.setenv("ORACLE_POW", "cr")
    else:
        monkeypatch.delenv("ORACLE_POW", raising=False)
    cls = dict(synth.VL, law=law, decel=3.0)
    s = run_parity(synth.single_link(length=800.0, vit_reg=15.0, demand=0.4, n_steps=300, classes=[cls], signal=(10, 2, 28)), 300)
    assert s["overflow"] == 0 and s["worst"] <= (1e-9 if pow_mode == "libm" else 0.0)


def test_two_classes_on_a_link(synth):
    s = run_parity(synth.single_link(length=1200.0, vit_reg=20.0, demand=0.5, n_steps=400, classes=[synth.VL, synth.PL],
                                     type_coefs=[0.85, 0.15], signal=(20, 3, 17)), 400)
    _check(s)


def test_exponential_headways(synth):
    _check(run_parity(synth.single_link(length=1200.0, vit_reg=20.0, demand=0.4, n_steps=400, exponential=True), 400))


def test_grid_with_demand(synth):
    """Small C2: 4x4 signalised grid, boundary demand, OD routes."""
    _check(run_parity(synth.grid(n=4, n_steps=500, demand_per_lane=0.2, demand_duration=400), 500))


def test_grid_saturated_demand(synth):
    _check(run_parity(synth.grid(n=4, n_steps=400, demand_per_lane=0.45, demand_duration=600), 400))


def test_grid_preseeded_shuffled_with_loops(synth):
    """Pre-seeded congested grid, list order shuffled: exercises leader-recursion loop breaking."""
    s = run_parity(synth.grid(n=6, n_steps=300, demand_per_lane=0.0, preseed_per_lane=12.6, shuffle=True), 300)
    _check(s, expect_loops=True)


def test_grid_rectangular_mixed(synth):
    _check(run_parity(synth.grid(n=5, ny=3, n_steps=300, demand_per_lane=0.3, preseed_per_lane=8, shuffle=True), 300))


def test_c2_config(synth):
    """BASELINE config C2 (10x10 signalised grid with its merge points, ~20k vehicles at peak): the full 1800 steps."""
    s = run_parity(synth.config("C2"), 1800, check_every=10)
    _check(s)
    assert s["merge_ins"] == s["merge_ins_oracle"] > 10000


# ---- merge points (row A10) ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kw", [
    dict(),                                                            # ramp + main road, one lane each
    dict(tf=6.0, demand_ramp=0.5, demand_main=0.0),                    # follow-up time: imposed approach speed of the branch head
    dict(demand_main=0.75, demand_ramp=0.3, vit_reg=14.0),             # dense priority stream: refusals, ramp queue
    dict(lanes_main=2, lanes_down=2, demand_main=0.8),                 # multi-lane merge: inner-lane follower draw, one point per downstream lane
    dict(lanes_main=3, lanes_down=2, demand_main=1.0, demand_ramp=0.25),
    dict(exit_capacity=0.25, demand_main=0.3),                         # congested downstream: sensor-based supply / demand test with draws
    dict(exit_capacity=0.3, demand_main=0.4, demand_ramp=0.3, exponential=True),
    dict(prio_main=2, prio_ramp=1),                                    # the ramp has the right of way
    dict(accbornee=True, demand_main=0.5),
], ids=lambda kw: "-".join(f"{k}={v}" for k, v in kw.items()) or "default")
def test_plain_merge(synth, kw):
    classes = [synth.VL, synth.PL] if kw.get("lanes_main", 1) > 1 else None
    s = run_parity(synth.ymerge(n_steps=500, classes=classes, type_coefs=[0.8, 0.2] if classes else None, **kw), 500)
    _check(s)
    assert s["merge_ins"] == s["merge_ins_oracle"] > 0


def test_corridor_with_on_and_off_ramps(synth):
    """Small C3: three lanes, two vehicle types, lane changing, on-ramp Convergents and off-ramp distributors."""
    b = synth.corridor(n_links=8, length=400.0, n_lanes=3, demand=1.1, offramp_every=2, onramp_every=2, onramp_demand=0.25, preseed_density=0.05,
                       shuffle=True, n_steps=400)
    s = run_parity(b, 400)
    _check(s)
    assert s["merge_ins"] == s["merge_ins_oracle"] > 100


def test_c3_config_prefix(synth):
    """BASELINE config C3 geometry (10 km, on / off ramps every km, lane changing, two types), first 300 steps."""
    s = run_parity(synth.config("C3"), 300, check_every=5)
    _check(s)


def test_lane_changing_without_origins(synth):
    """Pre-seeded corridor, no demand: the random stream is advanced on the device by the next-lane memo draws of every step, and the
    lane-change acceptance draws of the next step must come after them (the stream lives on the device)."""
    b = synth.corridor(n_links=6, length=500.0, n_lanes=3, demand=0.0, offramp_every=2, preseed_density=0.06, shuffle=True, n_steps=200)
    s = run_parity(b, 200)
    _check(s)


def test_side_streams_do_not_change_results(synth, pkg, monkeypatch):
    """The merge phase settles the points that share a vehicle beside the clean ones, the merge sensors beside the next step's lane
    ordering, the count-only draws beside the relaxation: with everything on one stream (SYMGPU_SIDE=0) the state must be the same,
    step after step (a point settled by k_mrg_dirty used to be taken again by k_mrg_eval when it happened to run later)."""
    import os
    import tempfile
    fd, path = tempfile.mkstemp(suffix=".symflat")
    os.close(fd)
    synth.fast_grid(path, 40, 40, 12.6, n_steps=100, merges=True)
    monkeypatch.setenv("SYMGPU_SIDE", "15")
    a = pkg.Engine(path)
    monkeypatch.setenv("SYMGPU_SIDE", "0")
    b = pkg.Engine(path)
    os.unlink(path)
    for k in range(80):
        a.step()
        b.step()
        assert a.n == b.n and a.rng_count == b.rng_count, k
        if k % 10 == 9:
            sa, sb = a.state(), b.state()
            for key in ("id", "lane", "pos", "vit", "acc"):
                assert np.array_equal(sa[key], sb[key]), (k, key)
    assert a.counter(pkg.CNT_MRG_INS) == b.counter(pkg.CNT_MRG_INS) > 1000


def test_c4_full_size_prefix_and_invariants(synth, pkg):
    """BASELINE config C4 at full size (100x100, ~1M vehicles): oracle parity on the first 100 steps (tools/gpu_long_parity.py runs it
    further: bit-exact over 600 steps), then
    size-independent invariants over a longer run: no follower passes its leader on a lane, positions stay
    within the lane, the population only shrinks by exits."""
    import os
    import tempfile
    b = synth.config("C4")
    s = run_parity(b, 100, check_every=20)
    _check(s)
    assert s["merge_ins"] == s["merge_ins_oracle"] > 300000
    fd, path = tempfile.mkstemp(suffix=".symflat")
    os.close(fd)
    b.write(path)
    g = pkg.Engine(path)
    os.unlink(path)
    lane_len = np.asarray(b.lane_f)
    n_prev = g.n
    for _ in range(4):
        n = g.run(5)
        assert n <= n_prev
        n_prev = n
        st = g.state()
        assert (st["pos"] >= 0).all() and (st["pos"] <= lane_len[st["lane"]] + 1e-9).all()
        order = np.lexsort((st["pos"], st["lane"]))
        same = st["lane"][order][1:] == st["lane"][order][:-1]
        assert (np.diff(st["pos"][order])[same] >= 0).all()
    assert g.counter(pkg.CNT_CTX_OVERFLOW) == 0
    g.close()


def test_trajectory_rows_sync_and_async(synth, pkg, tmp_path):
    """symgpu_step_traj / symgpu_step_traj_async deliver exactly the state rows, in m_LstVehicles order."""
    b = synth.grid(n=4, n_steps=100, demand_per_lane=0.3, preseed_per_lane=6, shuffle=True)
    p = str(tmp_path / "t.symflat")
    b.write(p)
    g = pkg.Engine(p)
    cap = 20000
    mk = lambda: {"id": np.zeros(cap, np.int32), "lane": np.zeros(cap, np.int32), "pos": np.zeros(cap), "vit": np.zeros(cap), "acc": np.zeros(cap)}
    bufs = [mk(), mk()]
    g.traj_bind(0, bufs[0]); g.traj_bind(1, bufs[1])
    for k in range(40):
        w = k & 1
        rows = g.step_traj_async(w)
        g.traj_wait(w)
        st = g.state()
        assert rows == len(st["id"])
        for key in ("id", "lane", "pos", "vit", "acc"):
            assert np.array_equal(bufs[w][key][:rows], st[key]), (k, key)
    one = mk()
    rows = g.step_traj(one)
    st = g.state()
    assert rows == len(st["id"]) and np.array_equal(one["pos"][:rows], st["pos"]) and np.array_equal(one["id"][:rows], st["id"])
    g.close()


@pytest.mark.parametrize("nbuf", [2, 3])
def test_trajectory_rows_pipelined(synth, pkg, tmp_path, nbuf):
    """The bench's pattern: the rows of step k are picked up while step k+1 runs (the pack and the copies are enqueued from the
    next step, on the side stream) — they must still be the state at the end of step k, exits and arrivals included."""
    b = synth.grid(n=4, n_steps=100, demand_per_lane=0.3, preseed_per_lane=6, shuffle=True)
    p = str(tmp_path / "t.symflat")
    b.write(p)
    g, ref = pkg.Engine(p), pkg.Engine(p)
    cap = 20000
    mk = lambda: {"id": np.zeros(cap, np.int32), "lane": np.zeros(cap, np.int32), "pos": np.zeros(cap), "vit": np.zeros(cap), "acc": np.zeros(cap)}
    bufs = [mk() for _ in range(nbuf)]
    for w in range(nbuf):
        g.traj_bind(w, bufs[w])
    states, counts = [], []

    def check(j):
        g.traj_wait(j % nbuf)
        st, rows = states[j], counts[j]
        assert rows == len(st["id"])
        for key in ("id", "lane", "pos", "vit", "acc"):
            assert np.array_equal(bufs[j % nbuf][key][:rows], st[key]), (j, key)

    for k in range(60):
        counts.append(g.step_traj_async(k % nbuf))
        ref.step(); states.append(ref.state())
        if k >= nbuf - 1:
            check(k - (nbuf - 1))
    for j in range(60 - (nbuf - 1), 60):
        check(j)
    g.close(); ref.close()


def test_grid_without_context_memo(synth, monkeypatch):
    """SYMGPU_CTXCACHE=0: every vehicle walks the itinerary / movement / stop-line tables every step (the path the per-vehicle
    memo of micro.cuh: CtxCache replaces for the vehicles that stayed on their lane); same bits against the oracle."""
    monkeypatch.setenv("SYMGPU_CTXCACHE", "0")
    _check(run_parity(synth.grid(n=4, n_steps=150, demand_per_lane=0.3, preseed_per_lane=6, shuffle=True), 150))
