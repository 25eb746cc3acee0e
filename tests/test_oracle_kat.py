"""Known-answer tests pinning the CPU oracle (CPU only).

The reference ships no tests or golden vectors, so the oracle is pinned by answers that can be
derived by hand from the cited reference code and by the published MT19937 known answer.
"""
import math
import os
import tempfile

import numpy as np
import pytest

from oracle.binding import Oracle, lib


def _oracle(builder):
    fd, path = tempfile.mkstemp(suffix=".symflat")
    os.close(fd)
    builder.write(path)
    o = Oracle(path, len(builder.links), len(builder.classes))
    os.unlink(path)
    return o


def test_mt19937_known_answers():
    L = lib()
    r = L.orc_rng_new(5489)                      # ISO C++ [rand.predef]: 10000th output of mt19937 is 4123659995
    for _ in range(9999):
        L.orc_rng_raw(r)
    assert L.orc_rng_raw(r) == 4123659995
    L.orc_rng_free(r)
    r = L.orc_rng_new(42)
    ref = np.random.RandomState(42).randint(0, 2 ** 32, size=64, dtype=np.uint64)   # legacy init_genrand(42)
    assert [L.orc_rng_raw(r) for _ in range(64)] == [int(x) for x in ref]
    L.orc_rng_free(r)


def test_uniform_int_bucket_rejection():
    """RandManager::myRand = uniform_int<>(0, 0x7FFE) over mt19937: r = x / (0xFFFFFFFF / 32767), redraw if r > 32766."""
    L = lib()
    a, b = L.orc_rng_new(7), L.orc_rng_new(7)
    for _ in range(2000):
        x = L.orc_rng_raw(a)
        while x // 131076 > 32766:
            x = L.orc_rng_raw(a)
        v = L.orc_rng_next(b)
        assert v == x // 131076 and 0 <= v <= 32766
    L.orc_rng_free(a)
    L.orc_rng_free(b)


def test_free_flow_and_deterministic_demand(synth):
    """NewellCarFollowing.cpp:86-100: a free vehicle advances min(vx, vit_reg)*dt; usage/SymuViaTripNode.cpp:283-285:
    demand q creates a vehicle every 1/q s at intra-step instant dt - (crit-1)*dt/(q*dt)."""
    o = _oracle(synth.single_link(length=2000.0, vit_reg=20.0, demand=0.5, n_steps=100))
    created_at = {}
    prev = {}
    for k in range(1, 61):
        o.step()
        s = o.state()
        for i, vid in enumerate(s["id"]):
            created_at.setdefault(int(vid), k)
            if int(vid) in prev and prev[int(vid)] >= 0:
                assert s["pos"][i] - prev[int(vid)] == pytest.approx(20.0, abs=1e-12)
                assert s["vit"][i] == pytest.approx(20.0, abs=1e-12)
            prev[int(vid)] = s["pos"][i] if s["pos"][i] > -1e-300 else -1.0
    # crit accumulates 0.5 per step: creations at steps 2, 4, 6, ... with t_creation = dt (crit == 1 exactly)
    assert [created_at[i] for i in range(5)] == [2, 4, 6, 8, 10]
    assert o.rng_count == 5 * len(created_at)    # lane, type, destination, itinerary, Jorge class per vehicle


def test_fractional_creation_instant(synth):
    """q = 0.4 veh/s: crit = 0.4, 0.8, 1.2 -> created at step 3 with t_creation = 1 - 0.2/0.4 = 0.5 s, so the
    vehicle travels (dt - 0.5) * v in its first step."""
    o = _oracle(synth.single_link(vit_reg=20.0, demand=0.4, n_steps=20))
    for _ in range(3):
        o.step()
    s = o.state()
    assert len(s["id"]) == 1
    assert s["pos"][0] == pytest.approx(0.5 * 20.0, abs=1e-9)
    assert s["vit"][0] == pytest.approx(20.0, abs=1e-9) and s["acc"][0] == 0.0


def test_red_light_stops_at_stop_line_and_queue_spacing(synth):
    """vehicule.cpp:1888-1893: red during the whole step => the vehicle stops exactly at the stop line;
    at standstill Newell followers queue at spacing 1/Kx (deltaN = 1)."""
    b = synth.single_link(length=1000.0, vit_reg=15.0, demand=0.5, n_steps=400, signal=(5, 0, 195))
    o = _oracle(b)
    for _ in range(150):
        o.step()
    s = o.state()
    on_l1 = s["lane"] == 0
    pos = np.sort(s["pos"][on_l1])[::-1]
    assert pos[0] == 500.0                       # stop line at the end of T1
    stopped = s["vit"][on_l1] == 0.0
    assert stopped.sum() >= 5
    q = np.sort(s["pos"][on_l1][stopped])[::-1]
    assert np.allclose(np.diff(q), -1 / 0.17, atol=1e-9)


def test_steady_state_spacing(synth):
    """NewellCarFollowing.cpp:311,364: following a leader at constant speed v the spacing converges to 1/Kx + v*tau,
    tau = -1/(w*Kx).  Saturated entry (demand above capacity at vit_reg) produces such a platoon."""
    w, kx, v = -5.8823, 0.17, 10.0
    o = _oracle(synth.single_link(length=3000.0, vit_reg=v, demand=0.9, n_steps=400))
    for _ in range(300):
        o.step()
    s = o.state()
    pos = np.sort(s["pos"][s["pos"] > 200])[::-1]
    gaps = -np.diff(pos)
    expected = 1 / kx + v * (-1 / (w * kx))
    g = gaps[5:60]
    assert g.min() == pytest.approx(expected, rel=1e-9), (g[:8], expected)       # never closer than equilibrium
    assert (np.abs(g - expected) < 1e-6).mean() > 0.7                             # most followers sit exactly on it
    assert o.waiting() > 0                       # entry is saturated: vehicles queue at the origin


def test_exit_and_link_counters(synth):
    """vehicule.cpp:4337-4442: entering the first link increments nb_entre; the exit link's nb_sorti is not
    incremented when leaving the network (link[0] is reset after the counters)."""
    b = synth.single_link(length=300.0, vit_reg=25.0, demand=0.5, n_steps=100)
    o = _oracle(b)
    n_exit = 0
    for _ in range(60):
        o.step()
        n_exit += len(o.exited()[0])
    e, x = o.link_counts()
    assert e[0] == o.counter(2) and x[0] == 0 and n_exit > 10
    ids, inst = o.exited()
    assert o.n + n_exit + o.waiting() == o.counter(2)


def test_signal_colour_cycle(synth):
    """ControleurDeFeux.cpp:1517-1541 with integer-second cycle arithmetic (:250-251): vehicles pass only during
    green+amber of a 20/3/17 plan; flow through the stop line is periodic with the 40 s cycle."""
    b = synth.single_link(length=600.0, vit_reg=15.0, demand=0.3, n_steps=400, signal=(20, 3, 17))
    o = _oracle(b)
    crossings = []
    prev_lane = {}
    for k in range(1, 321):
        o.step()
        s = o.state()
        for vid, ln in zip(s["id"], s["lane"]):
            if prev_lane.get(int(vid)) == 0 and ln == 1:
                crossings.append(k)
            prev_lane[int(vid)] = int(ln)
    phase = np.array(crossings) % 40
    # a crossing recorded at the end of step k happened during (k-1, k]: the light was green/amber at some instant of it
    assert ((phase >= 1) & (phase <= 24)).all(), sorted(set(phase.tolist()))
    assert len(crossings) > 40


@pytest.mark.parametrize("law", [1, 2])
def test_gipps_idm_run_and_respect_leader(synth, law):
    """GippsCarFollowing.cpp:50-122 / IDMCarFollowing.cpp:61-211: followers never pass their leader on the link."""
    cls = dict(synth.VL, law=law, decel=3.0)
    o = _oracle(synth.single_link(length=800.0, vit_reg=15.0, demand=0.4, n_steps=200, classes=[cls], signal=(10, 2, 28)))
    for _ in range(150):
        o.step()
        s = o.state()
        for ln in (0, 1):
            m = s["lane"] == ln
            order = np.argsort(s["id"][m])
            p = s["pos"][m][order]
            p = p[p > -1e-300]
            assert (np.diff(p) <= 1e-9).all()     # ids increase upstream: positions must not increase with id
    assert o.n > 5


# ---- merge points (row A10): answers derivable by hand from convergent.cpp:333-990 and AbstractCarFollowing.cpp:336-664 -------
def test_merge_alone_on_the_ramp_inserts_at_the_crossing_instant(synth):
    """No traffic on the priority branch, free flow downstream: every ramp vehicle inserts (no refusal), keeps the position the car
    following law gave it (free flow over ta = dt - tr from the merge = what was left of the step), and the point records the
    crossing instant t - dt + tr with tr = dt * (L - pos1) / travelled distance (AbstractCarFollowing.cpp:544-547, convergent.cpp:818)."""
    v, Lr = 16.0, 300.0
    b = synth.ymerge(demand_main=0.0, demand_ramp=0.2, vit_reg=v, len_ramp=Lr, n_steps=200)
    o = _oracle(b)
    ramp_lane = b.lane_of(1, 0)
    prev = {}
    last_seen = 0.0
    for k in range(1, 151):
        o.step()
        s = o.state()
        t_last, _ = o.merge_state()
        for i, vid in enumerate(s["id"]):
            if vid in prev and prev[vid][0] == ramp_lane and s["lane"][i] != ramp_lane:       # crossed the merge during this step
                p1 = prev[vid][1]
                tr = (Lr - p1) / v
                assert abs(t_last[0] - (k - 1.0 + tr)) < 1e-9
                assert abs(s["pos"][i] - (p1 + v - Lr)) < 1e-9 and abs(s["vit"][i] - v) < 1e-9
                last_seen = t_last[0]
        prev = {vid: (s["lane"][i], s["pos"][i]) for i, vid in enumerate(s["id"])}
    assert last_seen > 0 and o.counter(5) == 0 and o.counter(4) > 20                            # no refusal, every crossing is an insertion


def test_merge_follow_up_time_spaces_the_insertions(synth):
    """Downstream condition (AbstractCarFollowing.cpp:385-391): t - last passage >= max(Tf, 1/Qmax).  With Tf = 6 s and a ramp vehicle every
    2 s the insertions are exactly Tf apart once the ramp queues: the head of the non-priority branch is held to the speed that brings
    it to the merge when the follow-up time has elapsed (Convergent::CalculVitesseLeader, convergent.cpp:1121-1168) and inserts at
    tr = dt - (t - (last + Tf)) (AbstractCarFollowing.cpp:589-593)."""
    tf, Lr = 6.0, 300.0
    b = synth.ymerge(demand_main=0.0, demand_ramp=0.5, tf=tf, vit_reg=16.0, len_ramp=Lr, n_steps=300)
    o = _oracle(b)
    ramp_lane = b.lane_of(1, 0)
    instants = []
    for k in range(1, 251):
        o.step()
        t_last, _ = o.merge_state()
        if not instants or t_last[0] != instants[-1]:
            instants.append(float(t_last[0]))
        s = o.state()
        on_ramp = s["lane"] == ramp_lane
        if on_ramp.any():
            assert s["pos"][on_ramp].max() <= Lr + 1e-12
    gaps = np.diff([t for t in instants if t > 0])
    assert len(gaps) > 15 and gaps.min() >= tf - 1e-9
    assert np.allclose(gaps[-10:], tf, atol=1e-9)               # saturated ramp: one insertion per follow-up time
    assert o.n > 40                                               # the queue on the ramp (2 s headways in, 6 s out)


def test_merge_priority_stream_blocks_the_ramp(synth):
    """A dense priority stream leaves no gap of `dlag` upstream of the merge: ramp vehicles are refused and queue; without the stream
    the same ramp flows freely.  Dense = followers less than (vx_am - w) / (-Kx w) apart (AbstractCarFollowing.cpp:529-585)."""
    dense = _oracle(synth.ymerge(demand_main=0.75, demand_ramp=0.3, vit_reg=14.0, n_steps=400))
    free = _oracle(synth.ymerge(demand_main=0.0, demand_ramp=0.3, vit_reg=14.0, n_steps=400))
    for _ in range(300):
        dense.step()
        free.step()
    assert free.counter(5) == 0
    assert dense.counter(5) > 3 * dense.counter(4) > 0          # far more refusals than insertions
    assert dense.n > free.n + 20                                   # the ramp queue
