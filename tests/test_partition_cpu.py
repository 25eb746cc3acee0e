"""Host-side logic of the partitioned runs on CPU: the strip partitioner and the torch.distributed exchange helpers
(world_size 2, gloo) that move tail records and migrating vehicle rows between ranks."""
import importlib
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_strip_partition(synth):
    part = importlib.import_module("open-symuvia_b200.partition")
    b = synth.grid(n=8, ny=3, n_steps=10, demand_per_lane=0.0, preseed_per_lane=1.0)
    xy = np.asarray(b.node_xy)
    for world in (1, 2, 4):
        p = part.strip_partition(xy, world)
        assert p.min() == 0 and p.max() == world - 1
        # monotone in x: a node further east never belongs to a lower strip
        order = np.argsort(xy[:, 0], kind="stable")
        assert (np.diff(p[order]) >= 0).all()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = importlib.import_module("open-symuvia_b200.partition")
    T, R = part.TAIL_D, part.ROW_D
    # rank r exports (r + 2) lanes to the other rank; halo sizes mirror that
    export_cnt = np.zeros(world, np.int32)
    halo_cnt = np.zeros(world, np.int32)
    other = 1 - rank
    export_cnt[other] = rank + 2
    halo_cnt[other] = other + 2
    send = torch.arange(int(export_cnt.sum()) * T, dtype=torch.float64) + 1000 * rank
    recv = torch.zeros(int(halo_cnt.sum()) * T, dtype=torch.float64)
    part.exchange_tails_dist(dist, send, recv, export_cnt, halo_cnt)
    ok = bool(torch.equal(recv, torch.arange(int(halo_cnt.sum()) * T, dtype=torch.float64) + 1000 * other))
    # migrating rows: rank 0 sends 3 rows, rank 1 sends none
    cap = 8
    send_rows = torch.zeros(world * cap * R, dtype=torch.float64)
    recv_rows = torch.zeros(world * cap * R, dtype=torch.float64)
    counts = np.zeros(world, np.int32)
    if rank == 0:
        counts[1] = 3
        send_rows[1 * cap * R: (1 * cap + 3) * R] = torch.arange(3 * R, dtype=torch.float64) + 7
    n = part.exchange_rows_dist(dist, torch, send_rows, recv_rows, counts, cap, rank, world)
    if rank == 1:
        ok = ok and n == 3 and bool(torch.equal(recv_rows[: 3 * R], torch.arange(3 * R, dtype=torch.float64) + 7))
    else:
        ok = ok and n == 0
    # same rows through the single fixed-size all_to_all (count in the spare double of the region's first row)
    send2 = send_rows.clone()
    for peer in range(world):
        send2[peer * cap * R + 15] = float(counts[peer])
    regions = torch.zeros(world * cap * R, dtype=torch.float64)
    recv2 = torch.zeros(world * cap * R, dtype=torch.float64)
    n2 = part.exchange_rows_fixed(dist, torch, send2, regions, recv2, cap, rank, world)
    if rank == 1:
        want = torch.arange(3 * R, dtype=torch.float64) + 7
        want[15] = 3.0
        ok = ok and n2 == 3 and bool(torch.equal(recv2[: 3 * R], want))
    else:
        ok = ok and n2 == 0
    q.put((rank, ok))
    dist.destroy_process_group()


def test_exchange_helpers_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
