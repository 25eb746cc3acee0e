#!/usr/bin/env python
"""bench.py — vehicle-steps/s of the per-time-step microscopic update on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W     the CPU oracle on the host cores (reference arm)

A "step" is one simulated time step (Reseau::SimuTrafic, reseau.cpp:2028-2428) over every vehicle of the workload.
Workload at N=1: BASELINE config C4 — 100x100 signalised grid, ~1M pre-seeded vehicles (synthetic, seed 42).
`value`  : whole-job vehicle-steps/s, network + vehicles resident in HBM, timed with CUDA events on the step stream.
`e2e`    : same metric through the host-facing call symgpu_step_traj(): every step ends with the device->host copy
           of the trajectory rows (id, lane, pos, speed, acc) the XML writers consume; C4 has no origins, so the
           per-step host->device input (created vehicles) is 0 bytes — stated, not hidden.
`roofline`: dominant kernel, algorithmic bytes = 128 B per vehicle-step over its measured device time.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
B_ALG = 128.0   # algorithmic bytes per vehicle-step


def make_workload(name: str, rank: int, world: int, path: str, partitioned: bool = False):
    """C4: the 100x100 signalised grid with its merge points (csrc/gen_grid.cpp), ~1M pre-seeded vehicles.  N GPUs: one such network per
    GPU (replicas: the merge points' random draws are ordered over the whole network, so a network with merges is not partitioned yet).
    `partitioned` (merge-free networks only): ONE (100*world) x 100 grid cut into strips, the file holds the whole network and the
    vehicles of strip `rank` (global ids); boundary tails and migrating vehicles are exchanged every step (NCCL)."""
    synth = importlib.import_module("open-symuvia_b200.synth")
    if name == "C5":        # BASELINE configs[4]: ONE 300x300 grid, ~10M vehicles, cut into `world` vertical strips (strong scaling)
        n_own, n_tot = synth.fast_grid(path, 300, 300, per_lane=14.0, seed=42, route_len=24, n_steps=100, rank=rank, world=world, merges=(world == 1))
        info = dict(n_init=n_own, n_total=n_tot, n_classes=1)
        hdr = flat_section_counts(path)
        info["n_links"] = hdr["link_i"] // 8
        info["n_lanes"] = hdr["lane_i"] // 4
        return info
    if name == "C3":        # BASELINE configs[2]: multi-lane multi-class arterial with lane changing, on-ramp Convergents and off-ramp distributors, ~100 k vehicles
        b = synth.config("C3", km=1200, seed=42 + rank)
        b.write(path)
        return dict(n_init=len(b.iv_i), n_total=len(b.iv_i), n_classes=len(b.classes), n_links=len(b.links), n_lanes=len(b.lanes))
    tile = {"C4": 100, "C4small": 20}.get(name)
    if tile is None:
        raise SystemExit(f"unknown workload {name}")
    if partitioned and world > 1:
        n_own, n_tot = synth.fast_grid(path, tile * world, tile, per_lane=12.6, seed=42, route_len=24, n_steps=300, rank=rank, world=world, merges=False)
    else:
        n_own, n_tot = synth.fast_grid(path, tile, tile, per_lane=12.6, seed=42 + rank, route_len=24, n_steps=300, rank=0, world=1, merges=True)
    info = dict(n_init=n_own, n_total=n_tot, n_classes=1)
    hdr = flat_section_counts(path)
    info["n_links"] = hdr["link_i"] // 8
    info["n_lanes"] = hdr["lane_i"] // 4
    return info


def flat_section_counts(path: str):
    """{section: element count} of a SYMFLAT1 file, reading only the headers."""
    import struct
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == b"SYMFLAT1"
        (n,) = struct.unpack("<I", f.read(4))
        for _ in range(n):
            name = f.read(24).rstrip(b"\0").decode()
            dt, cnt = struct.unpack("<IQ", f.read(12))
            nbytes = cnt * (4 if dt == 0 else 8 if dt == 1 else 1)
            f.seek(nbytes + ((-nbytes) % 8), 1)
            out[name] = cnt
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 2 and r[1].isdigit())
        mx = max([int(r[2]) for r in self.rows if len(r) > 2 and r[2].isdigit()] or [0])
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_oracle_rate(path, info, steps, warmup=0, repeats=1, variant="faithful"):
    """vehicle-steps/s of the CPU oracle on ONE host core (the reference steps one network on one thread): `repeats` consecutive
    samples of `steps` time steps each, the median rate is returned (median of 5).  variant "flat" = the
    "sorted lanes" baseline (flat per-lane arrays instead of the reference's std::set; same results)."""
    from oracle.binding import Oracle
    o = Oracle(path, info["n_links"], info["n_classes"], variant=variant)
    for _ in range(warmup):
        o.step()
    rates, tot_all, dt_all = [], 0, 0.0
    for _ in range(repeats):
        tot = 0
        t0 = time.perf_counter()
        for _ in range(steps):
            tot += o.step()
        dt = time.perf_counter() - t0
        rates.append(tot / dt)
        tot_all += tot
        dt_all += dt
    o.close()
    rates.sort()
    return rates[len(rates) // 2], dt_all, tot_all


def measured_dram_bytes():
    """{kernel group: DRAM bytes (read + write) per time step of C4} from the committed ncu capture (profiles/r2_dram_bytes.json, made by
    tools/ncu_dram_summary.py from the CSV of `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`); {} when the file is absent."""
    p = os.path.join(ROOT, "profiles", "r2_dram_bytes.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--cpu-steps", type=int, default=15, help="oracle steps for cpu_baseline (bounded sample: 5 samples of cpu_steps / 5 steps)")
    ap.add_argument("--partitioned", action="store_true", help="N > 1: one merge-free network cut into N strips with NCCL exchange instead of N replicas")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    warm = max(a.warmup, 3)
    base_cfg = {"workload": (f"{a.workload}: synthetic 100x100 signalised Manhattan grid per GPU, 200 m blocks, 2 lanes/dir, ~1M pre-seeded vehicles per GPU, "
                             "Newell, dt=1 s, chgtvoie off, merge points at every junction exit (BASELINE.json configs[3])") if a.workload not in ("C5", "C3") else
                            ("C5: ONE synthetic 300x300 signalised Manhattan grid, ~10M pre-seeded vehicles in total, Newell, dt=1 s (BASELINE.json configs[4])" if a.workload == "C5" else
                             "C3: synthetic 1200 km 3-lane arterial, VL 85 % / PL 15 %, lane changing on, on-ramp Convergents and off-ramp distributors every km, ~100k vehicles pre-seeded + demand "
                             "(BASELINE.json configs[2])"),
                "working_set": "vehicle SoA + tables ~0.4 GB per GPU > 126 MB L2 (no L2 flush needed)"}
    path = f"/tmp/symuvia_bench_{a.workload}_{rank}.symflat"

    if a.impl == "reference":
        # Reference arm: the reference's own CPU algorithm (oracle port; the C++ reference cannot be built here)
        if rank != 0:
            return
        info = make_workload(a.workload, 0, 1, path)
        rate, dt, tot = cpu_oracle_rate(path, info, a.steps, warmup=warm)
        line = {"impl": "reference", "metric": "vehicle-steps/sec", "value": rate, "unit": "vehicle-steps/s", "n_gpus": a.gpus, "steps": a.steps,
                "warmup": warm, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": dict(base_cfg, parallelism="1 GPU" if a.gpus == 1 else f"{a.gpus} replicas", merge_points=True, vehicles_per_gpu=info["n_init"], lanes_per_gpu=info["n_lanes"],
                               reference_arm="CPU oracle (C++ restatement of the reference's algorithm: the reference itself cannot be built here), ONE host thread as the "
                                             "reference steps a network, one C4 network whatever --gpus says: a RATE to compare with the per-GPU rate"),
                "cpu_baseline": {"value": rate, "unit": "vehicle-steps/s", "cores": 1, "kind": "port", "cpu_model": cpu_model(), "host_cores": os.cpu_count(),
                                 "sample": f"{a.steps} full time steps of {a.workload} after {warm} warm-up steps ({tot} vehicle-steps, {dt:.1f} s)"},
                "e2e": {"value": rate, "unit": "vehicle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("open-symuvia_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # N > 1: ONE network (same file on every rank) cut into N vertical strips; every rank owns the vehicles of its strip
    partitioned = world > 1 and (a.partitioned or a.workload == "C5")
    info = make_workload(a.workload, rank, world, path, partitioned)
    run = None
    if partitioned:
        part_mod = importlib.import_module("open-symuvia_b200.partition")
        flat = importlib.import_module("open-symuvia_b200.flat")
        node_xy = flat.read(path)["node_xy"]
        run = part_mod.DistributedRun(pkg, path, part_mod.strip_partition(node_xy, world), local, cap_per_peer=4096)
        g = run.r.e
    else:
        g = pkg.Engine(path, device=local)
    L = g.L
    L.symgpu_profile.argtypes = [g.h.__class__, pkg.C.c_int]
    L.symgpu_kernel_ms.argtypes = [g.h.__class__, pkg.C.c_int]
    L.symgpu_kernel_ms.restype = pkg.C.c_double
    L.symgpu_kernel_launches.argtypes = [g.h.__class__, pkg.C.c_int]
    L.symgpu_kernel_launches.restype = pkg.C.c_long
    L.symgpu_kernel_name.restype = pkg.C.c_char_p
    L.symgpu_vehicle_steps.argtypes = [g.h.__class__]
    L.symgpu_vehicle_steps.restype = pkg.C.c_longlong

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def run_steps(k):
        if run is None:
            g.run(k)
            return g.last_step_ms
        t0 = time.perf_counter()
        tot_ms = 0.0
        for _ in range(k):
            run.step()
            tot_ms += g.last_step_ms        # CUDA events: front of the step .. import of the arrivals (exchange included)
        return tot_ms

    run_steps(warm)
    # ---- value: state resident in HBM, CUDA events around exactly K steps ------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    L.symgpu_profile(g.h, 2)            # events around the relaxation sweeps only (the dominant kernel), inside the timed region
    vs0, l0 = L.symgpu_vehicle_steps(g.h), g.counter(pkg.CNT_LAUNCHES)
    barrier()
    t0 = time.perf_counter()
    dev_ms = run_steps(a.steps)
    barrier()
    wall = time.perf_counter() - t0
    vsteps = L.symgpu_vehicle_steps(g.h) - vs0
    launches = g.counter(pkg.CNT_LAUNCHES) - l0
    def kernel_table():
        return {L.symgpu_kernel_name(k).decode(): (L.symgpu_kernel_ms(g.h, k), L.symgpu_kernel_launches(g.h, k)) for k in range(L.symgpu_kernel_count())}
    kms_timed = kernel_table()
    L.symgpu_profile(g.h, 0)
    t_max = allmax(dev_ms / 1e3)
    total_vsteps = allsum(float(vsteps))
    value = total_vsteps / t_max
    # ---- e2e: host-facing call, trajectory rows copied to host buffers every step ---------------------------------
    cap = g.n + 65536
    NB = 2                                   # two host buffers: the rows of step k are consumed while step k+1 runs
    bufs = [{"id": np.zeros(cap, np.int32), "lane": np.zeros(cap, np.int32), "pos": np.zeros(cap), "vit": np.zeros(cap), "acc": np.zeros(cap)}
            for _ in range(NB)]
    for w in range(NB):
        g.traj_bind(w, bufs[w])
    e2e_steps = max(3, a.steps)
    barrier()
    t0 = time.perf_counter()
    ev = 0
    rows = 0
    checksum = 0.0
    copy_ms = 0.0
    L.symgpu_traj_copy_ms.argtypes = [g.h.__class__, pkg.C.c_int]
    L.symgpu_traj_copy_ms.restype = pkg.C.c_double

    def consume(j):                          # rows of step j are in host memory: read them
        nonlocal checksum, copy_ms
        g.traj_wait(j % NB)
        checksum += float(bufs[j % NB]["vit"][0])
        copy_ms += L.symgpu_traj_copy_ms(g.h, j % NB)

    for k in range(e2e_steps):
        if run is None:
            rows = g.step_traj_async(k % NB)  # step k; its rows travel to host buffer k % NB on the side stream
        else:
            run.step()                       # partitioned: exchange included, then this rank's rows travel to the host
            rows = g.traj_copy_async(k % NB)
        ev += rows
        if k >= NB - 1:
            consume(k - (NB - 1))
    for j in range(max(0, e2e_steps - (NB - 1)), e2e_steps):
        consume(j)
    barrier()
    e2e_t = allmax(time.perf_counter() - t0)
    e2e = allsum(float(ev)) / e2e_t
    clocks = sampler.stop()
    # per-kernel breakdown: a last pass of K steps with every kernel group bracketed by events (not part of `value`)
    L.symgpu_profile(g.h, 1)
    vs1 = L.symgpu_vehicle_steps(g.h)
    run_steps(a.steps)
    vsteps_bd = L.symgpu_vehicle_steps(g.h) - vs1
    kms = kernel_table()
    L.symgpu_profile(g.h, 0)
    # ---- roofline of the dominant kernel -----------------------------------------------------------------------------
    peak, peak_src = peak_hbm()
    dom = max(kms, key=lambda k: kms[k][0])
    if kms_timed.get(dom, (0.0, 0))[0] > 0:      # timed inside the measured region
        dom_ms, dom_launches, dom_vsteps, dom_where = kms_timed[dom][0], kms_timed[dom][1], vsteps, "timed region"
    else:
        dom_ms, dom_launches, dom_vsteps, dom_where = kms[dom][0], kms[dom][1], vsteps_bd, "breakdown pass"
    achieved = B_ALG * dom_vsteps / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
    roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            # DRAM bytes (read + write) of the dominant kernel's launches of one step, from the ncu --set full capture summarised in
            # profiles/r1g_summary.md (C4: k_relax 350 B, k_context 529 B per vehicle-step); other cases: not captured
            "traffic": (measured_dram_bytes().get(dom) if a.workload == "C4" and not partitioned else None),
            "traffic_note": "DRAM bytes read + written by the group's launches of one C4 step, ncu capture committed as profiles/r2_dram_bytes.json (per GPU; null: not captured for this workload)",
            "peak_source": peak_src, "launches_in_region": dom_launches, "avg_launch_ms": dom_ms / max(1, dom_launches),
            "alg_bytes_per_vehicle_step": B_ALG, "whole_step_frac": B_ALG * vsteps / (dev_ms / 1e3) / 1e9 / peak,
            "kernel_timed_in": dom_where,
            "kernel_ms": {k: round(v[0], 3) for k, v in kms.items()},
            "kernel_ms_note": f"per-kernel CUDA-event totals of a separate pass of {a.steps} steps with every kernel group bracketed (the timed region only brackets the relaxation sweeps)"}
    line = {"metric": "vehicle-steps/sec", "value": value, "unit": "vehicle-steps/s", "n_gpus": world, "steps": a.steps, "warmup": warm,
            "ms_per_step": 1e3 * t_max / a.steps, "higher_is_better": True, "scaling": ("strong" if a.workload == "C5" else "weak"), "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": dict(base_cfg, parallelism=("1 GPU" if world == 1 else
                                                  f"{world} replicas, one C4 network (with merge points) per GPU, different seeds, no data-path collective (merge points are not partitioned yet: DESIGN.md section 7)" if not partitioned else
                                                  (f"one MERGE-FREE {100 * world}x100 grid cut into {world} vertical strips of 100x100 junctions" if a.workload != "C5" else f"the merge-free 300x300 grid cut into {world} vertical strips") +
                                                  ", one per GPU; per step: all_to_all of boundary-lane tails + one fixed-size all_to_all of migrating vehicles (NCCL)"),
                           merge_points=not partitioned,
                           vehicles_per_gpu=info["n_init"], lanes_per_gpu=info["n_lanes"]),
            "e2e": {"value": e2e, "unit": "vehicle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(rows) * 32,
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_t / e2e_steps, "d2h_copy_ms_per_step": copy_ms / max(1, e2e_steps),
                    "note": "symgpu_step_traj_async: rows of step k packed and copied to page-locked host buffers on a side stream while step k+1 runs; "
                            + ("the origins' new vehicles (about 100 bytes each, a few per step) are uploaded by the engine inside the step and not counted in h2d_bytes_per_step" if a.workload == "C3"
                               else "the workload has no origins, so the per-step host input is empty")},
            "gpu_launches": int(launches), "wall_s": wall, "clocks": clocks, "roofline": roof,
            "loops_broken": g.counter(pkg.CNT_LOOPS), "follow_passes": g.counter(pkg.CNT_PASSES),
            # which time steps of the run each pass covered (the simulation goes on from pass to pass; later steps of C4 are more congested and slower)
            "step_windows": {"warmup": [1, warm], "value": [warm + 1, warm + a.steps], "e2e": [warm + a.steps + 1, warm + a.steps + e2e_steps],
                             "kernel_breakdown": [warm + a.steps + e2e_steps + 1, warm + 2 * a.steps + e2e_steps]}}
    if rank == 0 and world == 1 and a.cpu_steps > 0:
        per = max(1, a.cpu_steps // 5)
        rate, dt, tot = cpu_oracle_rate(path, info, per, repeats=5)
        rate2, dt2, tot2 = cpu_oracle_rate(path, info, per, repeats=3, variant="flat")
        line["cpu_baseline"] = {"value": rate, "unit": "vehicle-steps/s", "cores": 1, "kind": "port", "cpu_model": cpu_model(), "host_cores": os.cpu_count(),
                                "sample": f"median of 5 consecutive samples of {per} time steps of the same {a.workload} input ({tot} vehicle-steps, {dt:.1f} s in all); "
                                          "the oracle keeps the reference's structure (id-ordered lane sets, linear neighbour scans, leader-first recursion), one thread",
                                # BASELINE.md section 3 asks for a second variant: flat per-lane arrays in id order instead of std::set (same results, better memory behaviour)
                                "sorted_lanes": {"value": rate2, "unit": "vehicle-steps/s", "cores": 1, "sample": f"median of 3 samples of {per} time steps ({tot2} vehicle-steps, {dt2:.1f} s)"}}
    if rank == 0:
        print(json.dumps(line))
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
