"""DRAM traffic per kernel group and time step from an ncu CSV (profiles/): python tools/ncu_dram_summary.py launches.csv out.json

Only whole time steps count: the launches between the first and the last k_reset (the first kernel of a step) of the capture.

The CSV comes from
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --launch-skip S -c N --csv ...
over whole time steps of the C4 workload (tools/gpu_steps.py).  Kernels are mapped to the groups bench.py reports (symgpu_kernel_name)."""
import csv
import json
import sys

GROUP = [("k_relax_done", "k_relax"), ("k_relax", "k_relax"), ("k_context", "k_context"), ("k_ctx_fill", "k_context"), ("k_mrg_", "k_merge"), ("k_rng_", "k_merge"), ("k_scan64", "k_merge"),
         ("k_signals", "k_signals"), ("k_prepare", "k_prepare"), ("k_reset", "k_prepare"), ("k_scan", "k_scan"), ("k_scatter", "k_scatter"), ("k_lane_", "k_lane_order"),
         ("k_seg", "k_seg"), ("k_cyc", "k_cyc_loops"), ("k_place", "k_place"), ("k_entries", "k_entries"), ("k_final", "k_final"), ("k_exit", "k_final"), ("k_lc", "k_lane_change"),
         ("k_alive", "k_traj"), ("k_pack", "k_traj")]


def group_of(name):
    for pat, g in GROUP:
        if pat in name:
            return g
    return "other"


def main():
    path, out = sys.argv[1], sys.argv[-1]
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[hdr]
    ki, mi, vi, ui = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit")
    idi = H.index("ID")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "nsecond": 1.0, "msecond": 1e6}
    per = {}
    marks = sorted({int(r[idi]) for r in rows[hdr + 1:] if len(r) > vi and r[idi].isdigit() and "k_reset" in r[ki]})
    if len(marks) < 2:
        raise SystemExit("the capture holds no whole time step")
    n_steps = len(marks) - 1
    for r in rows[hdr + 1:]:
        if len(r) <= vi or not r[idi].isdigit() or not (marks[0] <= int(r[idi]) < marks[-1]):
            continue
        g = group_of(r[ki])
        v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
        d = per.setdefault(g, {"dram_bytes": 0.0, "time_ns": 0.0, "launches": set()})
        if r[mi].startswith("dram__bytes"):
            d["dram_bytes"] += v
        elif r[mi].startswith("gpu__time"):
            d["time_ns"] += v
        d["launches"].add(r[idi])
    res = {g: d["dram_bytes"] / n_steps for g, d in per.items()}
    detail = {g: {"dram_bytes_per_step": d["dram_bytes"] / n_steps, "ncu_time_us_per_step": d["time_ns"] / 1e3 / n_steps, "launches_per_step": len(d["launches"]) / n_steps}
              for g, d in per.items()}
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)
    json.dump(detail, open(out.replace(".json", "_detail.json"), "w"), indent=1, sort_keys=True)
    print(f"{n_steps} whole steps, launches {marks[0]}..{marks[-1] - 1}")
    for g in sorted(detail, key=lambda k: -detail[k]["ncu_time_us_per_step"]):
        print(f"{g:16s} {detail[g]['dram_bytes_per_step'] / 1e6:9.1f} MB/step {detail[g]['ncu_time_us_per_step']:8.1f} us/step {detail[g]['launches_per_step']:5.1f} launches/step")


if __name__ == "__main__":
    main()
