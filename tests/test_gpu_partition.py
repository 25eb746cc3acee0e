"""Partitioned runs.  All ranks of a 2/3/4-way partition run in one process on the test GPU
(LocalGroup: same symgpu_mp_* calls, device-to-device copies instead of NCCL) and are compared step by step with
the oracle in its partitioned mode.  Needs a B200."""
import importlib

import numpy as np
import pytest

from oracle.binding import Oracle

pytestmark = pytest.mark.gpu


def _run(pkg, synth, tmp_path, builder, world, steps):
    part_mod = importlib.import_module("open-symuvia_b200.partition")
    p = str(tmp_path / "p.symflat")
    builder.write(p)
    node_part = part_mod.strip_partition(np.asarray(builder.node_xy), world)
    assert len(set(node_part.tolist())) == world
    o = Oracle(p, len(builder.links), len(builder.classes))
    o.set_partition(node_part)
    g = part_mod.LocalGroup(pkg, p, node_part, world)
    migrated = 0
    for k in range(steps):
        no = o.step()
        ng = g.step()
        assert no == ng, f"step {k + 1}: {no} vs {ng}"
        so, sg = o.state(), g.state()
        order = np.argsort(so["id"], kind="stable")
        for key in ("id", "lane", "route_pos", "flags", "leader"):
            assert np.array_equal(so[key][order], sg[key]), f"step {k + 1}: {key}"
        for key in ("pos", "vit", "acc", "deltaN"):
            assert np.array_equal(so[key][order], sg[key]), f"step {k + 1}: {key}"
        assert o.rng_count == g.rng_count()
        eo, xo = o.link_counts()
        eg, xg = g.link_counts()
        assert np.array_equal(eo, eg) and np.array_equal(xo, xg)
        migrated += sum(int(r.send_counts.sum()) for r in g.ranks)
    loops = sum(r.e.counter(pkg.CNT_LOOPS) for r in g.ranks)
    g.close()
    return migrated, loops, o.counter(1)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_partitioned_grid_matches_partitioned_oracle(pkg, synth, tmp_path, world):
    b = synth.grid(n=8, ny=4, n_steps=200, demand_per_lane=0.0, preseed_per_lane=9.0, shuffle=False, merges=False)
    migrated, loops_gpu, loops_oracle = _run(pkg, synth, tmp_path, b, world, 150)
    assert migrated > 50            # vehicles did cross the cuts
    assert loops_gpu == loops_oracle


def test_partition_of_one_is_the_plain_engine(pkg, synth, tmp_path):
    """world = 1: no halo, no migration — must reproduce the unpartitioned oracle."""
    part_mod = importlib.import_module("open-symuvia_b200.partition")
    b = synth.grid(n=4, n_steps=100, demand_per_lane=0.0, preseed_per_lane=10.0, merges=False)
    p = str(tmp_path / "p.symflat")
    b.write(p)
    o = Oracle(p, len(b.links))
    g = part_mod.LocalGroup(pkg, p, np.zeros(len(b.nodes), np.int32), 1)
    for k in range(80):
        assert o.step() == g.step()
        so, sg = o.state(), g.state()
        assert np.array_equal(so["id"], sg["id"]) and np.array_equal(so["pos"], sg["pos"]) and np.array_equal(so["lane"], sg["lane"])
    g.close()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_grid_with_origins(pkg, synth, tmp_path, world):
    """Boundary demand (deterministic creation, queues at the origins, draws per created vehicle) in a partitioned run: every
    rank replays all origins' draws and keeps the vehicles of its own entry lanes; the device draws are summed over the ranks."""
    b = synth.grid(n=8, ny=4, n_steps=300, demand_per_lane=0.25, demand_duration=200, preseed_per_lane=4.0, shuffle=False, merges=False)
    migrated, loops_gpu, loops_oracle = _run(pkg, synth, tmp_path, b, world, 220)
    assert migrated > 50
    assert loops_gpu == loops_oracle


def test_nccl_two_ranks(tmp_path):
    """The real multi-process path (partition.DistributedRun: one process per GPU, NCCL all_to_all of boundary tails and migrating
    vehicles) against the oracle's partitioned mode, every step, bit for bit.  Needs two GPUs (skipped otherwise)."""
    import json
    import os
    import subprocess
    import sys

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "mp.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(root, "tools", "mp_parity_worker.py"), out, "80"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.load(open(out))
    assert res["ok"], res
    assert res["migrated"] > 20
