"""Ad-hoc (torchrun, N GPUs): where a partitioned step spends its time.  torchrun ... tests/try_mp_timing.py [steps]"""
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

pkg = importlib.import_module("open-symuvia_b200")
part = importlib.import_module("open-symuvia_b200.partition")
flat = importlib.import_module("open-symuvia_b200.flat")
synth = importlib.import_module("open-symuvia_b200.synth")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
path = f"/tmp/mp_timing_{rank}.symflat"
synth.fast_grid(path, 100 * world, 100, per_lane=12.6, seed=42, route_len=24, n_steps=300, rank=rank, world=world)
run = part.DistributedRun(pkg, path, part.strip_partition(flat.read(path)["node_xy"], world), local, cap_per_peer=4096)
r = run.r
MAXC = 0
T = {k: 0.0 for k in ("front", "tails", "back", "rows", "import")}


def step():
    t0 = time.perf_counter(); r.front(); t1 = time.perf_counter()
    part.exchange_tails_dist(dist, r.send_tails, r.recv_tails, r.export_cnt, r.halo_cnt); torch.cuda.synchronize(); t2 = time.perf_counter()
    r.back(); t3 = time.perf_counter()
    if os.environ.get("MP_OLD"):
        n = part.exchange_rows_dist(dist, torch, r.send_rows, r.recv_rows, r.send_counts, r.cap, rank, world); torch.cuda.synchronize()
    else:
        n = part.exchange_rows_fixed(dist, torch, r.send_rows, r.recv_regions, r.recv_rows, r.cap, rank, world)
    t4 = time.perf_counter()
    global MAXC; MAXC = max(MAXC, int(r.send_counts.max()))
    r.import_rows(n); t5 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4


if os.environ.get("MP_RUN"):          # the production path (DistributedRun.step: device-ordered, no host sync between kernels and collectives)
    def step():
        run.step()
        return (0.0,) * 5
for _ in range(10):
    step()
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    for k, v in zip(T, step()):
        T[k] += v
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f"rank {rank}: max rows to one peer {MAXC}; {1e3 * tot / steps:.3f} ms/step; " + ", ".join(f"{k} {1e3 * v / steps:.3f}" for k, v in T.items()), flush=True)
st = r.e.state()
chk = torch.tensor([float(len(st["id"])), float(st["id"].astype("float64").sum()), float(st["pos"].sum()), float(st["vit"].sum())], dtype=torch.float64, device="cuda")
dist.all_reduce(chk)
import numpy as np
key = (2 * st["id"].astype(np.int64) + 1)
hx = int(np.bitwise_xor.reduce(st["vit"].view(np.int64) * key)) ^ int(np.bitwise_xor.reduce(st["pos"].view(np.int64) * key)) if len(key) else 0
allh = [None] * world
dist.all_gather_object(allh, hx)
if rank == 0:
    tot = 0
    for v in allh:
        tot ^= v
    print("rank 0 checksum", [repr(float(x)) for x in chk], "exact state hash", hex(tot & 0xffffffffffffffff), flush=True)
dist.destroy_process_group()
