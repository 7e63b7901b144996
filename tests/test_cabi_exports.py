"""The C-ABI library loads and exports every symbol include/remo3d_b200.h declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

from remo3d_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "remo3d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(remo_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    declared = _declared()
    assert len(declared) >= 20
    assert sorted(_cabi.SIGNATURES) == declared


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_cabi.LIB_PATH), "build the library first: python build.py"
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert _cabi.load() is _cabi.load()


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch

    if torch.cuda.is_available():
        return
    import pytest

    with pytest.raises(_cabi.RemoError):
        _cabi.Context(0)
    from remo3d_b200 import Model

    m = Model(["A2.0M0.5N"])
    with pytest.raises(_cabi.RemoError):
        m.initialize_workers(cpu_workers=1, gpu_workers=1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "remo3d_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), fn
