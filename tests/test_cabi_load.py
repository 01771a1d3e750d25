"""The C-ABI library builds, loads and exports every symbol include/lasso_b200.h declares.
No compute calls succeed without a GPU -- and they must fail loudly, not fall back."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "lasso_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lasso_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(build_extension):
    cabi = build_extension._cabi
    assert declared_functions() == sorted(cabi.EXPORTS)


def test_library_exports_every_declared_symbol(build_extension):
    cabi = build_extension._cabi
    lib = ctypes.CDLL(cabi.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert cabi.load().lasso_b200_version() >= 1000
    assert cabi.load().lasso_b200_last_error() == b""


def test_no_torch_types_in_the_abi():
    code = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)   # declarations only
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code
    assert "#include <stdint.h>" in code and code.count("#include") == 1


def test_select_path_is_callable_without_gpu(build_extension):
    cabi = build_extension._cabi
    assert cabi.select_path(65536, 64, 256) == cabi.PATH_RESIDENT
    assert cabi.select_path(128, 10, 50) == cabi.PATH_RESIDENT
    assert cabi.select_path(128, 150, 90) == cabi.PATH_FFMA


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without GPU")
def test_compute_fails_loudly_without_gpu(build_extension):
    from lasso_b200.linear import sparse_encode
    x, w = torch.randn(8, 4), torch.randn(4, 6)
    with pytest.raises(Exception) as info:
        sparse_encode(x, w, alpha=0.1, lr=0.1, maxiter=3)
    assert not isinstance(info.value, AssertionError)
    # and no silent CPU result either way
    with pytest.raises(Exception):
        build_extension.linear.solvers.lipschitz_constant(w)


def test_missing_library_raises(build_extension, monkeypatch):
    cabi = build_extension._cabi
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", os.path.join(ROOT, "does", "not", "exist.so"))
    with pytest.raises(cabi.LassoB200Error):
        cabi.load()


def test_integration_stub_compiles_and_binds_declared_symbols():
    """The ctypes stub of INTEGRATION.md section B must at least compile and only name exported symbols
    (it is executed on the GPU by tests/test_gpu_parity.py::test_integration_stub_runs_as_written)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    blocks = [b for b in re.findall(r"```python\n(.*?)```", text, flags=re.S) if "ctypes.CDLL" in b]
    assert len(blocks) == 1
    compile(blocks[0], "INTEGRATION.md", "exec")
    from lasso_b200 import _cabi
    for sym in set(re.findall(r"_lib\.(lasso_b200_\w+)", blocks[0])):
        assert sym in _cabi.EXPORTS, sym
