"""CPU tests: the C-ABI library builds, loads and exports every symbol include/fsr_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("fsr_build", os.path.join(ROOT, "fast-srgan_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    path = mod.build()
    return ctypes.CDLL(path)


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "fsr_b200.h")).read()
    names = set(re.findall(r"\b(fsr_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 10
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/fsr_b200.h but not exported"


def test_binding_covers_header():
    from fast_srgan_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "fsr_b200.h")).read()
    names = set(re.findall(r"\b(fsr_[a-z0-9_]+)\s*\(", hdr))
    assert names == set(_lib.EXPORTED_SYMBOLS)


def test_abi_version_and_errors(lib):
    lib.fsr_abi_version.restype = ctypes.c_int
    assert lib.fsr_abi_version() == 2
    lib.fsr_error_string.restype = ctypes.c_char_p
    assert lib.fsr_error_string(0) == b"ok"
    assert b"workspace" in lib.fsr_error_string(-4)


def test_generator_state_dict_contract():
    """Same 36 keys / shapes as the reference Generator; `_orig_mod.` prefix tolerated."""
    import types
    import srgan_oracle as O
    from fast_srgan_b200.model import Generator
    g = Generator(types.SimpleNamespace(n_filters=64, n_layers=8))
    sd = O.make_generator_state(64, 8)
    assert set(g.state_dict().keys()) == set(sd.keys()) and len(sd) == 36
    for k, v in g.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    g.load_state_dict({"_orig_mod." + k: v for k, v in sd.items()})


def test_cpu_input_raises():
    import types
    import torch
    from fast_srgan_b200.model import Generator
    g = Generator(types.SimpleNamespace(n_filters=64, n_layers=1))
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            g(torch.zeros(1, 3, 8, 8))


def test_missing_library_fails_loudly(monkeypatch):
    """No CPU / PyTorch fallback: without the built .so the product path raises instead of computing elsewhere."""
    from fast_srgan_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libfsr_b200.so")
    with pytest.raises(RuntimeError, match="no CPU"):
        _lib.load()


def test_product_package_never_imports_the_oracle():
    import glob
    for f in glob.glob(os.path.join(ROOT, "fast-srgan_b200", "*.py")):
        src = open(f).read()
        assert "srgan_oracle" not in src and "import oracle" not in src, f
