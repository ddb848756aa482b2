"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/edgpu.h declares, struct layouts agree between the oracle/product/ctypes mirrors, and
the product fails loudly (no CPU fallback) when no B200 is visible."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "edgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(edgpu_[a-z0-9_]+)\s*\(", src)))


def test_build_and_exports():
    from edipack_b200 import build as B
    import edipack_b200 as E

    B.build()
    L = E.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"libedgpu.so does not export {s}"
    assert sorted(E.ABI_SYMBOLS) == syms


def test_param_struct_layout(oracle):
    import edipack_b200 as E

    assert C.sizeof(E._abi.NormalParams) == C.sizeof(oracle.OraParams)
    # sizeof from the C side: 8 int32 + xmu + arrays
    n = 5
    expect = 8 * 4 + 8 + 8 * (2 * n * n + n + 4 + n + 4 * n * n + 2 * 2 * n * 32 + 2 * n * n * 32) + 4 * n * 32
    assert C.sizeof(E._abi.NormalParams) == expect


def test_no_cpu_fallback():
    """Without a GPU every entry point refuses to work instead of computing on the host."""
    import torch
    import edipack_b200 as E

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(E.EdgpuError, match="no CPU fallback"):
        E.ed_init(0)
    m = E.EDModel(Norb=1, Nbath=3)
    with pytest.raises(E.EdgpuError):
        E.build_Hv_sector_normal(m, 2, 2)
    import numpy as np
    with pytest.raises(E.EdgpuError):
        E.spHtimesV_p(np.zeros(36))


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (parity would be void)."""
    pkg = os.path.join(ROOT, "edipack_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "edipack_oracle" not in txt and "ed_oracle" not in txt, f


def test_model_mirrors_agree(oracle):
    """The oracle-side and product-side model builders produce identical parameter blocks."""
    import edipack_b200 as E
    from models import messy_kwargs, normal_normal_kwargs, replica_kwargs, star_kwargs

    for kw in (normal_normal_kwargs(), messy_kwargs(), messy_kwargs("hybrid"), replica_kwargs(),
               star_kwargs(7)):
        a = oracle.Model(**kw).params()
        b = E.EDModel(**kw).params()
        assert bytes(a) == bytes(b)
