"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = {}


def build(force: bool = False, variant: str = "faithful") -> str:
    """variant "flat": the "sorted lanes" CPU baseline of BASELINE.md section 3 (flat per-lane arrays in id order instead of std::set; same results)."""
    name = "liboracle.so" if variant == "faithful" else "liboracle_flat.so"
    so = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", name], stdout=subprocess.DEVNULL)
    return so


def lib(variant: str = "faithful"):
    if variant not in _LIB:
        L = C.CDLL(build(variant=variant))
        L.orc_load.restype = C.c_void_p
        L.orc_load.argtypes = [C.c_char_p]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p]
        L.orc_nveh.argtypes = [C.c_void_p]
        L.orc_time.argtypes = [C.c_void_p]
        L.orc_time.restype = C.c_double
        L.orc_rng_count.argtypes = [C.c_void_p]
        L.orc_rng_count.restype = C.c_uint
        L.orc_counter.argtypes = [C.c_void_p, C.c_int]
        L.orc_counter.restype = C.c_long
        L.orc_dump.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        L.orc_leaders.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_link_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_exited.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_waiting.argtypes = [C.c_void_p]
        L.orc_merge_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_set_partition.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_create_vehicle.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.orc_rng_new.restype = C.c_void_p
        L.orc_rng_new.argtypes = [C.c_uint]
        L.orc_rng_next.argtypes = [C.c_void_p]
        L.orc_rng_raw.argtypes = [C.c_void_p]
        L.orc_rng_raw.restype = C.c_uint
        L.orc_rng_free.argtypes = [C.c_void_p]
        _LIB[variant] = L
    return _LIB[variant]


class Oracle:
    """One simulation on the CPU oracle, loaded from a SYMFLAT1 file."""

    def __init__(self, flat_path: str, n_links: int, n_classes: int = 1, variant: str = "faithful"):
        self.L = lib(variant)
        self.h = self.L.orc_load(flat_path.encode())
        if not self.h:
            raise RuntimeError(f"oracle: cannot load {flat_path}")
        self.n_links, self.n_classes = n_links, n_classes

    def close(self):
        if self.h:
            self.L.orc_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def step(self) -> int:
        return self.L.orc_step(self.h)

    @property
    def n(self) -> int:
        return self.L.orc_nveh(self.h)

    @property
    def time(self) -> float:
        return self.L.orc_time(self.h)

    @property
    def rng_count(self) -> int:
        return self.L.orc_rng_count(self.h)

    def counter(self, which: int) -> int:
        return self.L.orc_counter(self.h, which)

    def create_vehicle(self, cls: int, entry: int, dest: int, lane: int, dbt: float) -> int:
        return self.L.orc_create_vehicle(self.h, cls, entry, dest, lane, dbt)

    def set_partition(self, node_part):
        part = np.ascontiguousarray(node_part, np.int32)
        self.L.orc_set_partition(self.h, part.ctypes.data_as(C.c_void_p))

    def waiting(self) -> int:
        return self.L.orc_waiting(self.h)

    def merge_state(self):
        n = self.L.orc_merge_state(self.h, None, None, 0)
        t = np.zeros(n, np.float64)
        r = np.zeros(n, np.int32)
        if n:
            self.L.orc_merge_state(self.h, t.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), n)
        return t, r

    def state(self):
        n = self.n
        i = {k: np.zeros(n, np.int32) for k in ("id", "lane", "route_pos", "flags")}
        d = {k: np.zeros(n, np.float64) for k in ("pos", "vit", "acc", "deltaN")}
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.L.orc_dump(self.h, p(i["id"]), p(i["lane"]), p(i["route_pos"]), p(i["flags"]),
                        p(d["pos"]), p(d["vit"]), p(d["acc"]), p(d["deltaN"]))
        i.update(d)
        ld = np.zeros(n, np.int32)
        self.L.orc_leaders(self.h, ld.ctypes.data_as(C.c_void_p))
        i["leader"] = ld
        return i

    def link_counts(self):
        e = np.zeros(self.n_links * self.n_classes, np.int32)
        s = np.zeros_like(e)
        self.L.orc_link_counts(self.h, e.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p))
        return e, s

    def exited(self, cap: int = 4096):
        ids = np.zeros(cap, np.int32)
        inst = np.zeros(cap, np.float64)
        n = self.L.orc_exited(self.h, ids.ctypes.data_as(C.c_void_p), inst.ctypes.data_as(C.c_void_p), cap)
        return ids[:n], inst[:n]
