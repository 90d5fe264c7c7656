"""CPU-side checks of the C ABI: the library loads, exports every symbol include/cwr.h declares,
and rejects use without a GPU loudly (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from clearwater_riverine_b200 import build
    build.build()
    from clearwater_riverine_b200.backend import load_library
    return load_library()


def test_every_declared_symbol_is_exported(lib):
    header = (ROOT / "include" / "cwr.h").read_text()
    declared = set(re.findall(r"\b(cwr_[a-z_]+)\s*\(", header))
    assert len(declared) >= 20
    raw = ctypes.CDLL(str(ROOT / "clearwater_riverine_b200" / "libcwr_b200.so"))
    missing = [s for s in sorted(declared) if not hasattr(raw, s)]
    assert not missing, missing
    assert set(lib._cwr_symbols) == declared


def test_default_options(lib):
    from clearwater_riverine_b200.backend import CwrOptions
    o = CwrOptions()
    assert lib.cwr_default_options(ctypes.byref(o)) == 0
    assert o.rtol == 1e-13 and o.precond_steps == 0 and o.precond_sweep == 1 and o.precond_precision == 32 and o.reorder == 1 and o.mass_flux == 1 and o.keep_history == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_fails_loudly_without_gpu(lib):
    from clearwater_riverine_b200 import CwrError, TransportBackend
    f1 = np.array([0, 0, 1], dtype=np.int32); f2 = np.array([1, 2, 3], dtype=np.int32)
    with pytest.raises(CwrError):
        TransportBackend(f1, f2, 4, 3, 1, 0.1)


def test_product_never_imports_the_oracle():
    pkg = ROOT / "clearwater_riverine_b200"
    for py in pkg.rglob("*.py"):
        text = py.read_text()
        assert "import oracle" not in text and "from oracle" not in text, py
