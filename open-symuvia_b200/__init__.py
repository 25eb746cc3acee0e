"""open-symuvia_b200 — B200-native drop-in for SymuVia's per-time-step microscopic vehicle update.

The product is `libSymuViaB200.so` (csrc/, hand-written sm_100a CUDA behind the C-ABI of
include/symgpu.h and the SymuVia exports of include/symuvia_exports.h).  This Python package is
only a thin ctypes mirror for tests and benchmarks; it never computes anything itself and has no
CPU fallback: if the shared library is missing or no GPU is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SYMB200_LIB", os.path.join(_HERE, "libSymuViaB200.so"))   # override only for A/B experiments
_LIB = None

(CNT_TIES, CNT_LOOPS, CNT_CREATED, CNT_PASSES, CNT_LAUNCHES, CNT_SLOTS, CNT_CTX_OVERFLOW, CNT_LANE_CHANGES, CNT_LC_EVENTS, CNT_LC_REEVAL,
 CNT_MRG_INS, CNT_MRG_REFUS, CNT_MRG_SERIAL, CNT_MRG_VALUES, CNT_MRG_DRAWS) = range(15)


class SymGpuError(RuntimeError):
    pass


def lib():
    """Loads libSymuViaB200.so (raises if it was not built: there is no fallback path)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise SymGpuError(f"{LIB_PATH} not built — run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        L.symgpu_create.argtypes = [C.c_char_p, ci, C.POINTER(vp)]
        L.symgpu_destroy.argtypes = [vp]
        L.symgpu_last_error.restype = C.c_char_p
        L.symgpu_step.argtypes = [vp]
        L.symgpu_run.argtypes = [vp, ci]
        L.symgpu_step_traj.argtypes = [vp, ci, vp, vp, vp, vp, vp]
        L.symgpu_traj_bind.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp]
        L.symgpu_step_traj_async.argtypes = [vp, ci]
        L.symgpu_traj_wait.argtypes = [vp, ci]
        L.symgpu_traj_copy_async.argtypes = [vp, ci]
        L.symgpu_num_vehicles.argtypes = [vp]
        L.symgpu_time.argtypes = [vp]
        L.symgpu_time.restype = cd
        L.symgpu_rng_count.argtypes = [vp]
        L.symgpu_rng_count.restype = C.c_uint
        L.symgpu_counter.argtypes = [vp, ci]
        L.symgpu_counter.restype = C.c_long
        L.symgpu_last_step_ms.argtypes = [vp]
        L.symgpu_last_step_ms.restype = cd
        L.symgpu_get_state.argtypes = [vp, ci] + [vp] * 9
        L.symgpu_get_link_counts.argtypes = [vp, vp, vp]
        L.symgpu_get_exited.argtypes = [vp, ci, vp, vp]
        L.symgpu_num_links.argtypes = [vp]
        L.symgpu_num_classes.argtypes = [vp]
        L.symgpu_waiting.argtypes = [vp]
        L.symgpu_get_merge_state.argtypes = [vp, ci, vp, vp]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """One network on one B200: mirrors the oracle binding's interface for the parity tests."""

    def __init__(self, flat_path: str, device: int = 0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.symgpu_create(flat_path.encode(), device, C.byref(h))
        if rc != 0:
            raise SymGpuError(f"symgpu_create failed ({rc}): {self.L.symgpu_last_error().decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.symgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _chk(self, rc):
        if rc < 0:
            raise SymGpuError(f"symgpu call failed ({rc}): {self.L.symgpu_last_error().decode()}")
        return rc

    def step(self) -> int:
        return self._chk(self.L.symgpu_step(self.h))

    def run(self, n_steps: int) -> int:
        return self._chk(self.L.symgpu_run(self.h, n_steps))

    @property
    def n(self) -> int:
        return self.L.symgpu_num_vehicles(self.h)

    @property
    def time(self) -> float:
        return self.L.symgpu_time(self.h)

    @property
    def rng_count(self) -> int:
        return self.L.symgpu_rng_count(self.h)

    @property
    def last_step_ms(self) -> float:
        return self.L.symgpu_last_step_ms(self.h)

    def counter(self, which: int) -> int:
        return self.L.symgpu_counter(self.h, which)

    def waiting(self) -> int:
        return self.L.symgpu_waiting(self.h)

    def state(self):
        n = self.n
        i = {k: np.zeros(n, np.int32) for k in ("id", "lane", "route_pos", "flags", "leader")}
        d = {k: np.zeros(n, np.float64) for k in ("pos", "vit", "acc", "deltaN")}
        got = self._chk(self.L.symgpu_get_state(self.h, n, _p(i["id"]), _p(i["lane"]), _p(i["route_pos"]), _p(i["flags"]),
                                                _p(d["pos"]), _p(d["vit"]), _p(d["acc"]), _p(d["deltaN"]), _p(i["leader"])))
        assert got == n, (got, n)
        i.update(d)
        return i

    def step_traj(self, bufs):
        """One step + trajectory rows into preallocated host arrays (id, lane, pos, vit, acc)."""
        return self._chk(self.L.symgpu_step_traj(self.h, len(bufs["id"]), _p(bufs["id"]), _p(bufs["lane"]), _p(bufs["pos"]),
                                                 _p(bufs["vit"]), _p(bufs["acc"])))

    def traj_bind(self, which: int, bufs):
        """Registers host arrays (id, lane, pos, vit, acc) as trajectory buffer 0 or 1 (page-locked in place)."""
        self._keep = getattr(self, "_keep", {})
        self._keep[which] = bufs
        return self._chk(self.L.symgpu_traj_bind(self.h, which, len(bufs["id"]), _p(bufs["id"]), _p(bufs["lane"]), _p(bufs["pos"]),
                                                 _p(bufs["vit"]), _p(bufs["acc"])))

    def step_traj_async(self, which: int) -> int:
        return self._chk(self.L.symgpu_step_traj_async(self.h, which))

    def traj_copy_async(self, which: int) -> int:
        return self._chk(self.L.symgpu_traj_copy_async(self.h, which))

    def traj_wait(self, which: int):
        self._chk(self.L.symgpu_traj_wait(self.h, which))

    def link_counts(self):
        n = self.L.symgpu_num_links(self.h) * self.L.symgpu_num_classes(self.h)
        e = np.zeros(n, np.int32)
        s = np.zeros(n, np.int32)
        self._chk(self.L.symgpu_get_link_counts(self.h, _p(e), _p(s)))
        return e, s

    def merge_state(self):
        """(last passage instant, draw counter) of every merge point, in Liste_convergents order."""
        n = self._chk(self.L.symgpu_get_merge_state(self.h, 0, None, None))
        t = np.zeros(n, np.float64)
        r = np.zeros(n, np.int32)
        if n:
            self._chk(self.L.symgpu_get_merge_state(self.h, n, _p(t), _p(r)))
        return t, r

    def exited(self, cap: int = 4096):
        ids = np.zeros(cap, np.int32)
        inst = np.zeros(cap, np.float64)
        n = self._chk(self.L.symgpu_get_exited(self.h, cap, _p(ids), _p(inst)))
        return ids[:n], inst[:n]


class SymuVia:
    """ctypes view of the SymuVia C exports (include/symuvia_exports.h) exactly as a user of the reference's
    libSymuVia.so binds them (SymExpC.h:15-53): one global current network, caller-allocated XML buffer."""

    def __init__(self, xml_buffer_bytes: int = 1 << 24):
        self.L = lib()
        L = self.L
        L.SymLoadNetworkEx.argtypes = [C.c_char_p]
        L.SymLoadNetworkEx.restype = C.c_int
        L.SymRunEx.argtypes = [C.c_char_p]
        L.SymRunEx.restype = C.c_int
        L.SymRunNextStepEx.argtypes = [C.c_char_p, C.c_bool, C.POINTER(C.c_bool)]
        L.SymRunNextStepEx.restype = C.c_bool
        L.SymRunNextStepLiteEx.argtypes = [C.c_bool, C.POINTER(C.c_bool)]
        L.SymRunNextStepLiteEx.restype = C.c_bool
        L.SymCreateVehicleEx.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_double]
        L.SymCreateVehicleEx.restype = C.c_int
        L.SymUnloadCurrentNetworkEx.argtypes = []
        self.buf = C.create_string_buffer(xml_buffer_bytes)

    def load(self, path: str) -> bool:
        return bool(self.L.SymLoadNetworkEx(path.encode()))

    def run(self, path: str) -> int:
        return self.L.SymRunEx(path.encode())

    def step(self, lite: bool = False):
        end = C.c_bool(False)
        if lite:
            ok = self.L.SymRunNextStepLiteEx(True, C.byref(end))
            return bool(ok), bool(end.value), None
        ok = self.L.SymRunNextStepEx(self.buf, True, C.byref(end))
        return bool(ok), bool(end.value), (self.buf.value.decode() if ok else None)

    def create_vehicle(self, typ: str, origin: str, dest: str, lane: int, dbt: float) -> int:
        return self.L.SymCreateVehicleEx(typ.encode(), origin.encode(), dest.encode(), lane, dbt)

    def unload(self):
        self.L.SymUnloadCurrentNetworkEx()
