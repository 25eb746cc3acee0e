"""One rank of the NCCL partition parity test (tests/test_gpu_partition.py::test_nccl_two_ranks): torchrun --nproc-per-node 2 tools/mp_parity_worker.py OUT.

Every rank steps its strip of one merge-free pre-seeded grid through open-symuvia_b200.partition.DistributedRun (the real exchange:
all_to_all of the boundary tails and of the migrating vehicles over NCCL); rank 0 also steps the CPU oracle in its partitioned mode and
compares the union of the ranks' vehicles with it after every step (ids, lanes, fp64 state bit for bit)."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 80
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("open-symuvia_b200")
    synth = importlib.import_module("open-symuvia_b200.synth")
    part_mod = importlib.import_module("open-symuvia_b200.partition")
    b = synth.grid(n=8, ny=4, n_steps=steps + 10, demand_per_lane=0.0, preseed_per_lane=9.0, shuffle=False, merges=False)
    path = f"/tmp/mp_parity_{rank}.symflat"
    b.write(path)
    node_part = part_mod.strip_partition(np.asarray(b.node_xy), world)
    run = part_mod.DistributedRun(pkg, path, node_part, local, cap_per_peer=2048)
    o = None
    if rank == 0:
        from oracle.binding import Oracle
        o = Oracle(path, len(b.links), len(b.classes))
        o.set_partition(node_part)
    keys_i, keys_d = ("id", "lane", "route_pos", "flags", "leader"), ("pos", "vit", "acc", "deltaN")
    ok, msg, migrated = True, "", 0
    for k in range(steps):
        run.step()
        migrated += int(run.r.send_counts.sum())
        st = run.r.e.state()
        n = torch.tensor([len(st["id"])], dtype=torch.int64, device="cuda")
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, n)
        sizes = [int(s.item()) for s in sizes]
        cap = max(sizes + [1])
        packed = torch.zeros(cap, 9, dtype=torch.float64, device="cuda")
        if len(st["id"]):
            packed[: len(st["id"])] = torch.from_numpy(np.stack([st[q].astype(np.float64) for q in keys_i + keys_d], axis=1)).cuda()
        allp = [torch.zeros_like(packed) for _ in range(world)]
        dist.all_gather(allp, packed)
        if rank == 0:
            no = o.step()
            rows = np.concatenate([allp[r][: sizes[r]].cpu().numpy() for r in range(world)])
            rows = rows[np.argsort(rows[:, 0], kind="stable")]
            so = o.state()
            order = np.argsort(so["id"], kind="stable")
            if no != len(rows):
                ok, msg = False, f"step {k + 1}: vehicle count oracle {no} gpus {len(rows)}"
            else:
                for j, q in enumerate(keys_i + keys_d):
                    if not np.array_equal(so[q][order].astype(np.float64), rows[:, j]):
                        ok, msg = False, f"step {k + 1}: {q} differs"
                        break
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda")
        dist.broadcast(flag, 0)
        if not int(flag.item()):
            break
    tot = torch.tensor([migrated], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    if rank == 0:
        json.dump({"ok": ok, "msg": msg, "steps": steps, "world": world, "migrated": int(tot.item())}, open(out, "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
