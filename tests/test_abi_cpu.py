"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol of include/ga_b200.h."""
import ctypes
import os

import pytest
import torch

from gen_adversarial_b200 import _lib, ops


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert set(_lib._PROTOS) == set(names), set(_lib._PROTOS) ^ set(names)
    assert L.ga_abi_version() == 11


def test_error_convention_without_gpu():
    """bad arguments -> non-zero return + message (no compute, no CUDA call)."""
    L = _lib.lib()
    rc = L.ga_pgd_linf_step(None, None, None, 0.1, 0.1, 16, None)
    assert rc != 0
    assert b"ga_pgd_linf_step" in L.ga_last_error()


def test_ops_refuse_cpu_tensors():
    x = torch.zeros(1, 4, 4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.maxpool2x2(x)


def test_sass_contains_blackwell_tensor_core_and_tma_ops():
    """cuobjdump evidence that the conv kernel is tcgen05 (UTCHMMA), TMEM (LDTM) and TMA (UTMALDG) code."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
