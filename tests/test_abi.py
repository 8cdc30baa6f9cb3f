"""The C-ABI shared library loads in a GPU-less process and exports every symbol include/cfbpe.h declares.
No compute call is made here (there is no device); creating a context must FAIL, not fall back."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from cfbpe import _native
    if not os.path.exists(_native.SO_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("cfbpe_build", os.path.join(ROOT, "cyberfabric-core_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return _native.load()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cfbpe.h")).read()
    return sorted(set(re.findall(r"CFBPE_API[^;(]*?\b(cfbpe_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    from cfbpe import _native
    assert declared_symbols() == sorted(_native.EXPORTS)


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert getattr(lib, name) is not None
    assert lib.cfbpe_abi_version() == 1


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "cfbpe.h")).read()
    assert "torch" not in src and "at::" not in src and "#include <cuda" not in src


def test_struct_layouts_match_the_header():
    from cfbpe import _native as N
    assert ctypes.sizeof(N.Config) == 24 + 4 * 8 + 4 + 4          # the ABI-1 fields, then devices[8], n_devices, n_workspaces
    assert N.Config.devices.offset == 24 and N.Config.n_devices.offset == 56 and N.Config.n_workspaces.offset == 60
    assert ctypes.sizeof(N.VocabInfo) == 24
    assert ctypes.sizeof(N.Profile) == 4 * 10 + 4 * 10 + 12 + 4 + 72


def test_create_fails_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cfbpe import _native as N
    with pytest.raises(N.NativeError) as ei:
        N.Context(0, 1 << 20, 16)
    assert ei.value.code == N.ENODEV       # no CPU fallback


def test_old_and_bad_configs(lib):
    """struct_size versions the config: too small a struct is refused before any CUDA call; so are duplicate devices"""
    from cfbpe import _native as N
    h = ctypes.c_void_p()
    cfg = N.Config(8, 0, 1 << 20, 16, 0)
    assert lib.cfbpe_create(ctypes.byref(cfg), ctypes.byref(h)) == N.EINVAL
    assert lib.cfbpe_create(None, ctypes.byref(h)) == N.EINVAL


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under cyberfabric-core_b200/ may reference it"""
    pkg = os.path.join(ROOT, "cyberfabric-core_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, os.path.join(d, f)
