"""The C-ABI shared library: builds in tree, loads, exports every symbol include/*.h declares,
and refuses to run without CUDA.  No compute calls here (CPU suite)."""
import ctypes
import os
import re

import pytest
import torch

from projected_lmc_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "plmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(plmc_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    path = build.build()
    assert path.exists() and path.parent.name == "projected_lmc_b200"


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(str(_cabi.lib_path()))
    names = header_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/plmc_b200.h but not exported"
    assert set(names) == set(_cabi.EXPORTED_SYMBOLS), set(names) ^ set(_cabi.EXPORTED_SYMBOLS)
    bound = _cabi.load()
    assert bound.plmc_version() >= 100
    assert bound.plmc_npad(44484) == 44544 and bound.plmc_npad(128) == 128
    assert bound.plmc_dinv_bytes(256, 3) == 3 * (256 * 128 + 512 * 512) * 8       # leaf inverses + one dense 512-slot
    assert bound.plmc_dinv_bytes(1152, 1) == (1152 * 128 + 3 * 512 * 512) * 8
    # above 2048: one 2048-slot per diagonal 2048-block for its explicit inverse (panel solves of potrf)
    assert bound.plmc_dinv_bytes(4224, 1) == (4224 * 128 + 9 * 512 * 512 + 3 * 2048 * 2048) * 8


def test_sass_has_fp64_tensor_core_and_async_copy_instructions():
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", str(_cabi.lib_path())], capture_output=True, text=True).stdout
    assert "DMMA.8x8x4" in sass, "FP64 tensor-core MMA missing from the sm_100a build"
    assert "LDGSTS" in sass, "cp.async staging missing"
    # the dominant kernels: tcgen05 INT8 MMA (single CTA and CTA pair), TMEM loads, bulk async copies
    for mnemonic in ("UTCIMMA", "UTCIMMA.2CTA", "LDTM", "UBLKCP", "UTCBAR", "IDP.4A"):
        assert mnemonic in sass, f"{mnemonic} missing from the sm_100a build"
    assert "sm_100a" in sass or "SM100" in sass.upper()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(_cabi.PlmcError):
        _cabi.lib()
    with pytest.raises(_cabi.PlmcError):
        _cabi.ptr(torch.zeros(3, dtype=torch.float64))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "_LIB_PATH", tmp_path / "libplmc_b200.so")
    with pytest.raises(_cabi.PlmcError, match="no CPU fallback"):
        _cabi.load()
