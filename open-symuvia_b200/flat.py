"""SYMFLAT1 container: the flattened network/demand tables shared by the host library and the test oracle.

The reference keeps the network as a pointer graph built by its XML loader
(`symuvia/src/reseau.cpp:4931-9612`).  The B200 path needs flat tables; this module is the
Python writer/reader of the binary container both the product (`csrc/host/flatnet.cpp`) and
the CPU checker used by the tests load.  Layout (little endian):

    "SYMFLAT1" | u32 n_sections | { char name[24] | u32 dtype | u64 count | data, padded to 8 B }*

dtype: 0 = int32, 1 = float64, 2 = uint8.  Section names and row strides are listed in
`include/symflat.h` (authoritative) and mirrored in `SECTIONS` below.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

MAGIC = b"SYMFLAT1"
_DT = {0: np.int32, 1: np.float64, 2: np.uint8}
_DTI = {np.dtype(np.int32): 0, np.dtype(np.float64): 1, np.dtype(np.uint8): 2}

# params indices (f64[32])
P_DT, P_NSTEPS, P_SEED, P_ACCBORNEE, P_RELAX, P_CHGTVOIE, P_DSTCHGT, P_TRAVERSEES, P_GHOST = range(9)
P_PHI_FORCE = 9      # chgtvoie mandatory: probability when forced near link end
P_DST_FORCE = 10     # chgtvoie mandatory: forcing distance before link end
P_NPARAMS = 32

# class row (f64[CLS_STRIDE])
CLS_STRIDE = 24
C_W, C_KX, C_VX, C_ESP, C_DECEL, C_TAULC, C_PIRAB, C_LAW, C_NPLAGE = range(9)
C_PLAGE0 = 9         # (ax, vit_sup) x 4 -> 9..16
LAW_NEWELL, LAW_GIPPS, LAW_IDM = 0, 1, 2

LINK_I, LINK_F, LANE_I, SUCC_I, SL_I = 8, 2, 4, 3, 6
ENTRY_I, DEST_I, EXIT_I, IV_I, IV_F, NODE_I = 8, 3, 2, 4, 3, 4


def write(path: str, sections: Dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(sections)))
        for name, arr in sections.items():
            a = np.ascontiguousarray(arr)
            if a.dtype not in _DTI:
                raise TypeError(f"section {name}: dtype {a.dtype}")
            nm = name.encode()
            assert len(nm) < 24, name
            f.write(nm.ljust(24, b"\0"))
            f.write(struct.pack("<IQ", _DTI[a.dtype], a.size))
            raw = a.tobytes()
            f.write(raw)
            f.write(b"\0" * ((-len(raw)) % 8))


def read(path: str) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("not a SYMFLAT1 file")
        (n,) = struct.unpack("<I", f.read(4))
        for _ in range(n):
            name = f.read(24).rstrip(b"\0").decode()
            dt, cnt = struct.unpack("<IQ", f.read(12))
            nbytes = cnt * np.dtype(_DT[dt]).itemsize
            raw = f.read(nbytes)
            f.read((-nbytes) % 8)
            out[name] = np.frombuffer(raw, dtype=_DT[dt]).copy()
    return out
