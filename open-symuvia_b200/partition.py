"""Partitioned (multi-GPU) runs: one process per GPU, road graph cut into regions, boundary exchange every step.

Host-side plumbing only (torch.distributed over NCCL / gloo); the computation stays behind the C-ABI
(`symgpu_mp_*`, include/symgpu.h).  Per step and rank:

    symgpu_mp_step_front   origins-free first half: lane membership + order, then the tails of the lanes other ranks look
                           into are packed (4 doubles per lane)                                      ── all_to_all ──►
    symgpu_mp_step_back    received tails become ghost leaders; context, car following, placement, finalisation;
                           vehicles that ended on a foreign lane are packed with their whole state (16 doubles)
                                                                                       ── counts, then send/recv ──►
    symgpu_mp_import       arrivals join the local vehicle list

SURVEY.md §8(e): a link belongs to the partition of its downstream node; messages are tiny (one record per boundary lane,
a few migrating vehicles), so the exchange is latency bound.  A follower whose leader is a ghost uses the reference's
estimated leader speed (NewellCarFollowing.cpp:217-230); the oracle has the same switch for the parity tests.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence

import numpy as np

TAIL_D = 4      # doubles per tail record
ROW_D = 16      # doubles per migrating vehicle


def strip_partition(node_xy: np.ndarray, world: int) -> np.ndarray:
    """Vertical strips of equal width over the x range of the nodes (grid networks: BASELINE config C5)."""
    x = np.asarray(node_xy, np.float64).reshape(-1, 2)[:, 0]
    lo, hi = float(x.min()), float(x.max())
    if hi <= lo or world == 1:
        return np.zeros(len(x), np.int32)
    part = np.floor((x - lo) / (hi - lo) * world).astype(np.int32)
    return np.clip(part, 0, world - 1)


def bind_mp(L):
    vp, ci = C.c_void_p, C.c_int
    L.symgpu_mp_configure.argtypes = [vp, ci, ci, vp]
    L.symgpu_mp_sizes.argtypes = [vp, vp, vp]
    L.symgpu_mp_step_front.argtypes = [vp, vp]
    L.symgpu_mp_step_back.argtypes = [vp, vp, vp, ci, vp]
    L.symgpu_mp_import.argtypes = [vp, vp, ci]
    L.symgpu_mp_import_regions.argtypes = [vp, vp, ci]
    L.symgpu_set_stream.argtypes = [vp, vp]
    L.symgpu_mp_p2p_handle.argtypes = [vp, vp]
    L.symgpu_mp_p2p_connect.argtypes = [vp, vp, vp, vp]
    L.symgpu_mp_has_origins.argtypes = [vp]
    L.symgpu_mp_local_draws.argtypes = [vp]
    L.symgpu_mp_local_draws.restype = C.c_long
    L.symgpu_mp_commit_draws.argtypes = [vp, C.c_long]


class RankEngine:
    """One rank's engine + its exchange buffers (torch tensors on the rank's device)."""

    def __init__(self, pkg, flat_path: str, rank: int, world: int, node_part: np.ndarray, device: int = 0, cap_per_peer: int = 4096):
        import torch
        self.torch = torch
        self.pkg, self.rank, self.world, self.cap = pkg, rank, world, cap_per_peer
        self.e = pkg.Engine(flat_path, device=device)
        self.L = self.e.L
        bind_mp(self.L)
        part = np.ascontiguousarray(node_part, np.int32)
        n = self.L.symgpu_mp_configure(self.e.h, rank, world, part.ctypes.data_as(C.c_void_p))
        if n < 0:
            raise pkg.SymGpuError(f"symgpu_mp_configure failed ({n}): {self.L.symgpu_last_error().decode()}")
        self.export_cnt = np.zeros(world, np.int32)
        self.halo_cnt = np.zeros(world, np.int32)
        self.e._chk(self.L.symgpu_mp_sizes(self.e.h, self.export_cnt.ctypes.data_as(C.c_void_p), self.halo_cnt.ctypes.data_as(C.c_void_p)))
        dev = torch.device("cuda", device)
        self.send_tails = torch.zeros(max(1, int(self.export_cnt.sum())) * TAIL_D, dtype=torch.float64, device=dev)
        self.recv_tails = torch.zeros(max(1, int(self.halo_cnt.sum())) * TAIL_D, dtype=torch.float64, device=dev)
        self.send_rows = torch.zeros(world * cap_per_peer * ROW_D, dtype=torch.float64, device=dev)
        self.recv_rows = torch.zeros(world * cap_per_peer * ROW_D, dtype=torch.float64, device=dev)
        self.recv_regions = torch.zeros(world * cap_per_peer * ROW_D, dtype=torch.float64, device=dev)
        self.send_counts = np.zeros(world, np.int32)
        self.has_origins = self.L.symgpu_mp_has_origins(self.e.h) == 1

    def front(self):
        self.e._chk(self.L.symgpu_mp_step_front(self.e.h, C.c_void_p(self.send_tails.data_ptr())))

    def back(self) -> int:
        return self.e._chk(self.L.symgpu_mp_step_back(self.e.h, C.c_void_p(self.recv_tails.data_ptr()), C.c_void_p(self.send_rows.data_ptr()),
                                                      self.cap, self.send_counts.ctypes.data_as(C.c_void_p)))

    def local_draws(self) -> int:
        return int(self.L.symgpu_mp_local_draws(self.e.h))

    def commit_draws(self, total: int):
        self.e._chk(self.L.symgpu_mp_commit_draws(self.e.h, int(total)))

    def import_regions(self) -> int:
        return self.e._chk(self.L.symgpu_mp_import_regions(self.e.h, C.c_void_p(self.recv_regions.data_ptr()), self.cap))

    def import_rows(self, n_rows: int) -> int:
        return self.e._chk(self.L.symgpu_mp_import(self.e.h, C.c_void_p(self.recv_rows.data_ptr()), n_rows))


def split_offsets(cnt: Sequence[int], width: int) -> List[int]:
    off = [0]
    for c in cnt:
        off.append(off[-1] + int(c) * width)
    return off


def exchange_tails_dist(dist, send, recv, export_cnt, halo_cnt):
    """all_to_all of the tail records (works for NCCL/cuda and gloo/cpu tensors alike)."""
    insz = [int(c) * TAIL_D for c in export_cnt]
    outsz = [int(c) * TAIL_D for c in halo_cnt]
    dist.all_to_all_single(recv[: sum(outsz)], send[: sum(insz)], output_split_sizes=outsz, input_split_sizes=insz)


def exchange_rows_dist(dist, torch, send_rows, recv_rows, send_counts, cap, rank, world):
    """Counts first (fixed size), then the used part of each per-peer region; returns rows received (packed at the front)."""
    dev = send_rows.device
    sc = torch.as_tensor(np.asarray(send_counts, np.int64), device=dev)
    rc = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc)
    rc_h = [int(x) for x in rc.cpu()]
    ops, off = [], 0
    for q in range(world):
        if q == rank:
            continue
        if send_counts[q] > 0:
            ops.append(dist.P2POp(dist.isend, send_rows[q * cap * ROW_D: (q * cap + int(send_counts[q])) * ROW_D], q))
        if rc_h[q] > 0:
            ops.append(dist.P2POp(dist.irecv, recv_rows[off: off + rc_h[q] * ROW_D], q))
            off += rc_h[q] * ROW_D
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return off // ROW_D


def exchange_rows_fixed(dist, torch, send_rows, recv_regions, recv_rows, cap, rank, world):
    """One fixed-size all_to_all of the per-peer regions; the row count of a region travels in the spare double (index 15) of
    its first row (k_mp_row_counts), so no count exchange is needed.  Returns rows received (packed at the front of recv_rows)."""
    dist.all_to_all_single(recv_regions, send_rows)
    cnt = [int(x) for x in recv_regions.view(world, cap * ROW_D)[:, 15].cpu()]
    off = 0
    for q in range(world):
        c = cnt[q]
        if q == rank or c <= 0:
            continue
        recv_rows[off: off + c * ROW_D].copy_(recv_regions[q * cap * ROW_D: (q * cap + c) * ROW_D])
        off += c * ROW_D
    return off // ROW_D


class DistributedRun:
    """The run as seen by one rank of a torch.distributed job (bench.py --gpus N, torchrun)."""

    def __init__(self, pkg, flat_path, node_part, device, cap_per_peer=4096):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.r = RankEngine(pkg, flat_path, self.rank, self.world, node_part, device, cap_per_peer)
        # the engine runs on torch's current stream: its kernels and the NCCL collectives are ordered on the device, no host sync between them
        self.r.e._chk(self.r.L.symgpu_set_stream(self.r.e.h, C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
        # SYMGPU_P2P=1: tails travel over peer memory (NVLink) instead of an NCCL all_to_all: every rank exports its halo area through
        # CUDA IPC, the pack kernel of a peer stores straight into it and raises a flag (engine.cu: k_mp_pack_tails_p2p /
        # k_mp_wait_tails).  Bit-identical state (tests/try_mp_timing.py: exact state hash, N=2) and the same step time as the collective
        # (N=2 0.97 vs 0.97 ms, N=4 1.09 vs 1.06): on the caller's stream the collective's latency already hides behind the lane
        # ordering, so the collective stays the default until the migration exchange goes over peer memory too.
        self.p2p = os.environ.get("SYMGPU_P2P", "0") == "1" and self.world > 1
        if self.p2p:
            r = self.r
            h = C.create_string_buffer(64)
            ok = r.L.symgpu_mp_p2p_handle(r.e.h, h) >= 0
            gathered = [None] * self.world
            dist.all_gather_object(gathered, (h.raw, [int(c) for c in r.halo_cnt], ok))
            if all(g[2] for g in gathered):
                handles = b"".join(g[0] for g in gathered)
                peer_off = np.array([sum(g[1][: self.rank]) for g in gathered], np.int32)   # where this rank's lanes start in peer q's halo
                peer_n = np.array([sum(g[1]) for g in gathered], np.int32)
                ok = r.L.symgpu_mp_p2p_connect(r.e.h, handles, peer_off.ctypes.data_as(C.c_void_p), peer_n.ctypes.data_as(C.c_void_p)) >= 0
            else:
                ok = False
            flags = [None] * self.world
            dist.all_gather_object(flags, ok)
            if not all(flags):                         # a peer could not be mapped (no peer access between two GPUs): everybody uses the collective
                raise pkg.SymGpuError("tail exchange over peer memory could not be set up on every rank (set SYMGPU_P2P=0 to use the NCCL all_to_all): "
                                      + r.L.symgpu_last_error().decode())
            dist.barrier()
        self._tail_in = [int(c) * TAIL_D for c in self.r.export_cnt]
        self._tail_out = [int(c) * TAIL_D for c in self.r.halo_cnt]
        self._tail_in_n, self._tail_out_n = sum(self._tail_in), sum(self._tail_out)

    def step(self) -> int:
        r, torch, dist = self.r, self.torch, self.dist
        r.front()
        if not self.p2p:
            dist.all_to_all_single(r.recv_tails[: self._tail_out_n], r.send_tails[: self._tail_in_n], output_split_sizes=self._tail_out, input_split_sizes=self._tail_in)
        r.back()
        dist.all_to_all_single(r.recv_regions, r.send_rows)     # per-peer regions of migrating vehicles, row counts inside (k_mp_row_counts)
        if r.has_origins:                             # one global random sequence: the ranks' device draws of this step, summed
            t = torch.tensor([r.local_draws()], dtype=torch.int64, device=r.send_rows.device)
            dist.all_reduce(t)
            r.commit_draws(int(t.item()))
        return r.import_regions()


class LocalGroup:
    """All ranks of a partitioned run inside ONE process on one GPU (tests on a single-GPU box): same engine calls, the
    exchange is a device-to-device copy instead of NCCL."""

    def __init__(self, pkg, flat_path, node_part, world, device=0, cap_per_peer=4096):
        import torch
        self.torch = torch
        self.world = world
        self.ranks = [RankEngine(pkg, flat_path, r, world, node_part, device, cap_per_peer) for r in range(world)]

    def step(self) -> int:
        torch, W = self.torch, self.world
        for r in self.ranks:
            r.front()
        for dst in self.ranks:                       # tails: sender's export region for dst -> dst's halo region for sender
            hoff = split_offsets(dst.halo_cnt, TAIL_D)
            for src in self.ranks:
                if src is dst:
                    continue
                eoff = split_offsets(src.export_cnt, TAIL_D)
                n = hoff[src.rank + 1] - hoff[src.rank]
                assert n == eoff[dst.rank + 1] - eoff[dst.rank]
                if n:
                    dst.recv_tails[hoff[src.rank]: hoff[src.rank + 1]].copy_(src.send_tails[eoff[dst.rank]: eoff[dst.rank + 1]])
        torch.cuda.synchronize()
        for r in self.ranks:
            r.back()
        if self.ranks[0].has_origins:
            draws = sum(r.local_draws() for r in self.ranks)
            for r in self.ranks:
                r.commit_draws(draws)
        total = 0
        for dst in self.ranks:                       # migrating vehicles: the sender's region for dst -> dst's region for the sender, as the
            W = dst.cap * ROW_D                      # fixed-size all_to_all delivers them (row counts inside); imported like DistributedRun does
            for src in self.ranks:
                if src is not dst:
                    dst.recv_regions[src.rank * W: (src.rank + 1) * W].copy_(src.send_rows[dst.rank * W: (dst.rank + 1) * W])
            torch.cuda.synchronize()
            total += dst.import_regions()
        return total

    def state(self):
        """Union of the ranks' vehicles ordered by id (= m_LstVehicles order for pre-seeded networks)."""
        parts = [r.e.state() for r in self.ranks]
        out = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
        order = np.argsort(out["id"], kind="stable")
        return {k: v[order] for k, v in out.items()}

    def rng_count(self) -> int:
        if self.ranks[0].has_origins:                 # every rank carries the whole sequence
            counts = {r.e.rng_count for r in self.ranks}
            assert len(counts) == 1, counts
            return counts.pop()
        return sum(r.e.rng_count for r in self.ranks)

    def link_counts(self):
        e = sum(r.e.link_counts()[0] for r in self.ranks)
        s = sum(r.e.link_counts()[1] for r in self.ranks)
        return e, s

    def close(self):
        for r in self.ranks:
            r.e.close()
