"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol the header
declares; without a GPU the product refuses to run instead of falling back."""
import ctypes as C

import pytest


def test_library_exports_every_declared_symbol():
    import twoarmy_b200 as pkg
    so = pkg._capi.build()
    L = C.CDLL(str(so))
    declared = pkg._capi.declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), name
    assert pkg._capi.lib().ta_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    import twoarmy_b200 as pkg
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.TwoarmyLibraryError):
        pkg.TwoarmyVecEnv("MiniGrid-twoarmy-17x17-v4", 4)
    h = C.c_void_p()
    rc = pkg._capi.lib().ta_create(C.byref(h), 4, 64, 17, 0, 0, 0)
    assert rc == -2  # TA_E_CUDA


def test_product_does_not_import_the_oracle():
    """The package must never reach into oracle/ (parity claims depend on it)."""
    import pathlib
    import twoarmy_b200 as pkg
    root = pathlib.Path(pkg.__file__).parent
    for p in list(root.rglob("*.py")) + list(root.rglob("*.cu")) + list(root.rglob("*.cuh")):
        text = p.read_text()
        assert "oracle" not in text.lower().replace("# oracle", ""), p


def test_argument_validation():
    import twoarmy_b200 as pkg
    L = pkg._capi.lib()
    h = C.c_void_p()
    assert L.ta_create(C.byref(h), 5, 64, 17, 0, 0, 0) == -1   # bad version
    assert L.ta_create(C.byref(h), 4, 64, 8, 0, 0, 0) == -1    # even view
    assert L.ta_create(C.byref(h), 4, 0, 17, 0, 0, 0) == -1    # no envs
    assert L.ta_strerror(-1) == b"invalid argument"
