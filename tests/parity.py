"""Step-by-step comparison of the CUDA engine (through the C-ABI) with the CPU oracle on the same SYMFLAT1 input.

Gates: vehicle list order, ids, lane, route cursor, regime flags, leader ids, link entry/exit
counters, exits, RNG draw count: bit-exact.  pos / speed / acc / deltaN: <= 1e-9 relative (the tolerance
BASELINE.json states); the maximum difference seen is returned so tests can also assert bit-exactness.
"""
import importlib
import os
import tempfile

import numpy as np

RTOL = 1e-9


def _close(a, b):
    d = np.abs(a - b)
    scale = np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
    return float((d / scale).max()) if len(a) else 0.0


def run_parity(builder, n_steps, check_every=1, verbose=False, leader_ties_ok=False, check_all_from=None):
    pkg = importlib.import_module("open-symuvia_b200")
    from oracle.binding import Oracle
    fd, path = tempfile.mkstemp(suffix=".symflat")
    os.close(fd)
    try:
        builder.write(path)
        o = Oracle(path, len(builder.links), len(builder.classes))
        g = pkg.Engine(path)
        worst = 0.0
        veh_steps = 0
        tie_diffs = 0
        for k in range(n_steps):
            no = o.step()
            ng = g.step()
            assert no == ng, f"step {k + 1}: vehicle count oracle {no} != gpu {ng}"
            veh_steps += no
            if (k + 1) % check_every and k + 1 != n_steps and not (check_all_from is not None and k + 1 >= check_all_from):
                continue
            so, sg = o.state(), g.state()
            for key in ("id", "lane", "route_pos", "flags", "leader"):
                if key == "leader" and leader_ties_ok and not np.array_equal(so[key], sg[key]):
                    # diagnostic mode (tools/gpu_long_parity.py): a differing leader id is tolerated when both candidates stand at the same place
                    bad = np.nonzero(so[key] != sg[key])[0]
                    row_of = {int(i): r for r, i in enumerate(so["id"])}
                    for r in bad:
                        a, b = row_of.get(int(so[key][r]), -1), row_of.get(int(sg[key][r]), -1)
                        assert a >= 0 and b >= 0 and so["lane"][a] == so["lane"][b] and so["pos"][a] == so["pos"][b], f"step {k + 1}: leader of id {so['id'][r]} differs and the candidates are not at one place"
                    tie_diffs += len(bad)
                    print(f"step {k + 1}: {len(bad)} leader ids differ among vehicles standing at one place", flush=True)
                    continue
                if not np.array_equal(so[key], sg[key]):
                    bad = np.nonzero(so[key] != sg[key])[0][:5]
                    if verbose:
                        allbad = np.nonzero((so["lane"] != sg["lane"]) | (so["pos"] != sg["pos"]) | (so["vit"] != sg["vit"]) | (so["leader"] != sg["leader"]))[0]
                        print(f"step {k + 1}: {len(allbad)} rows differ")
                        for r in allbad[:24]:
                            print(f"  id {so['id'][r]} oracle lane {so['lane'][r]} pos {so['pos'][r]!r} vit {so['vit'][r]!r} leader {so['leader'][r]} flags {so['flags'][r]} | gpu lane {sg['lane'][r]} pos {sg['pos'][r]!r} vit {sg['vit'][r]!r} leader {sg['leader'][r]} flags {sg['flags'][r]}")
                    raise AssertionError(f"step {k + 1}: {key} differs at rows {bad}: oracle {so[key][bad]} gpu {sg[key][bad]} "
                                         f"(ids {so['id'][bad]})")
            for key in ("pos", "vit", "acc", "deltaN"):
                e = _close(so[key], sg[key])
                if e > RTOL:
                    bad = int(np.argmax(np.abs(so[key] - sg[key])))
                    raise AssertionError(f"step {k + 1}: {key} rel err {e:.3e} at row {bad} id {so['id'][bad]}: "
                                         f"oracle {so[key][bad]!r} gpu {sg[key][bad]!r}")
                worst = max(worst, e)
            assert o.rng_count == g.rng_count, (f"step {k + 1}: rng count oracle {o.rng_count} gpu {g.rng_count} (merge insertions oracle {o.counter(4)} gpu "
                                                f"{g.counter(pkg.CNT_MRG_INS)}, refusals {o.counter(5)} / {g.counter(pkg.CNT_MRG_REFUS)}, congested tests {o.counter(7)} / {g.counter(pkg.CNT_MRG_VALUES)})")
            eo, xo = o.link_counts()
            eg, xg = g.link_counts()
            assert np.array_equal(eo, eg) and np.array_equal(xo, xg), f"step {k + 1}: link counters differ"
            io, to = o.exited()
            ig, tg = g.exited()
            assert np.array_equal(io, ig), f"step {k + 1}: exited ids differ {io} {ig}"
            assert _close(to, tg) <= RTOL, f"step {k + 1}: exit instants differ"
            assert o.waiting() == g.waiting(), f"step {k + 1}: waiting queues differ {o.waiting()} {g.waiting()}"
            (mo_t, mo_r), (mg_t, mg_r) = o.merge_state(), g.merge_state()
            assert np.array_equal(mo_r, mg_r) and np.array_equal(mo_t, mg_t), f"step {k + 1}: merge point state differs at {np.nonzero((mo_r != mg_r) | (mo_t != mg_t))[0][:5]}"
            if verbose and (k + 1) % 50 == 0:
                print(f"step {k + 1}: n={no} worst={worst:.2e} rng={o.rng_count} passes={g.counter(pkg.CNT_PASSES)}")
        stats = dict(worst=worst, veh_steps=veh_steps, ties_gpu=g.counter(pkg.CNT_TIES), ties_oracle=o.counter(0),
                     loops_gpu=g.counter(pkg.CNT_LOOPS), loops_oracle=o.counter(1), passes=g.counter(pkg.CNT_PASSES),
                     overflow=g.counter(pkg.CNT_CTX_OVERFLOW), n_final=no, merge_ins=g.counter(pkg.CNT_MRG_INS), merge_ins_oracle=o.counter(4),
                     merge_refus=g.counter(pkg.CNT_MRG_REFUS), merge_serial=g.counter(pkg.CNT_MRG_SERIAL), merge_values=g.counter(pkg.CNT_MRG_VALUES),
                     merge_cong_oracle=o.counter(7), leader_tie_diffs=tie_diffs)
        g.close()
        o.close()
        return stats
    finally:
        os.unlink(path)
