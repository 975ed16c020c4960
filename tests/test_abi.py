"""CPU tests: the C-ABI library loads and exports exactly what include/r2l_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "r2l_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(r2l_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol(E):
    import importlib.util
    spec = importlib.util.spec_from_file_location("r2l_build", os.path.join(ROOT, "efficient-nerf_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    path = b.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/r2l_b200.h but not exported"


def test_python_binding_covers_the_header(E):
    assert sorted(E._lib.SIGNATURES) == header_symbols()
    lib = E._lib.load()
    assert lib.r2l_abi_version() == E._lib.ABI_VERSION == 7


def test_argument_validation_needs_no_gpu(E):
    """Bad arguments are rejected on the host, before any CUDA call."""
    lib = E._lib.load()
    rc = lib.r2l_get_rays(0, 10, 1.0, None, None, None, None)
    assert rc != 0 and b"bad H/W/focal" in lib.r2l_last_error()
    rc = lib.r2l_embed(4, 3, 10, 1, 7, None, None, None)
    assert rc != 0 and b"layout" in lib.r2l_last_error()
    rc = lib.r2l_sample_pdf(4, 1, 8, None, 1, None, 0, None, 0, None, None, None)
    assert rc != 0
    rc = lib.r2l_sample_pdf(0, 63, 128, None, 63, None, 62, None, 0, None, None, None)
    assert rc == 0  # empty input is a no-op


def test_no_cpu_fallback(E):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.get_rays(4, 4, 5.0, torch.eye(4)[:3])
    net = E.NeRF(8, 256, 63, 27, 5, [4], True)
    with pytest.raises(RuntimeError, match="CUDA only"):
        with torch.no_grad():
            net(torch.zeros(2, 90))


def test_graft_entry_build_passes_its_own_checks():
    """The driver's 'does it build' entry: compiles (a no-op when the sources are unchanged), imports the package and
    checks the ABI version against the binding's constant."""
    import __graft_entry__
    __graft_entry__.build()
