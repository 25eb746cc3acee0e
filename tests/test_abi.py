"""The C-ABI library loads on a CPU-only host and exports every symbol the headers declare (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(symgpu_\w+|Sym\w+Ex|SymRunDeleteStr)\s*\(", txt)))


def test_library_exports_all_declared_symbols(pkg):
    L = ctypes.CDLL(pkg.LIB_PATH)
    names = _declared("symgpu.h") + (_declared("symuvia_exports.h") if os.path.exists(os.path.join(ROOT, "include", "symuvia_exports.h")) else [])
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"missing export {n}"


def test_no_cpu_fallback(pkg, synth, tmp_path):
    """Without a CUDA device the product refuses to run (it must never route through a CPU path)."""
    import pytest
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    p = str(tmp_path / "n.symflat")
    synth.single_link(n_steps=5).write(p)
    with pytest.raises(pkg.SymGpuError):
        pkg.Engine(p)


def test_product_never_touches_oracle():
    """oracle/ is test infrastructure: no product source may include, import or link it."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "open-symuvia_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"oracle/|liboracle|from oracle|import oracle", txt):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
