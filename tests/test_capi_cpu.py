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


def test_host_decode_of_the_packed_transfer_form():
    """The host stage of ta_step_host (pure CPU code in the library): 2-bit cell codes -> Grid.encode bytes, status byte ->
    reward / terminated / truncated, against a numpy restatement of the documented layout; ragged tile, unaligned output."""
    import numpy as np
    import twoarmy_b200 as pkg
    L = pkg._capi.lib()
    enc = np.array([[1, 0, 0], [2, 5, 0], [6, 4, 0], [8, 1, 0]], np.uint8)
    rlut = np.array([-0.01, -0.1, -0.9, 0.2, 0.9], np.float64).astype(np.float32)
    rng = np.random.default_rng(0)
    for V, n, shift in [(17, 96, 0), (17, 45, 0), (7, 33, 0), (3, 70, 0), (9, 31, 1), (17, 64, 3)]:
        ntiles = (n + 31) // 32
        words = rng.integers(0, 2 ** 32, size=(ntiles, 2 * V * V), dtype=np.uint64).astype(np.uint32)
        status = (rng.integers(0, 5, size=ntiles * 32) | (rng.integers(0, 2, size=ntiles * 32) << 3) | (rng.integers(0, 2, size=ntiles * 32) << 4)).astype(np.uint8)
        raw = np.full(n * V * V * 3 + 80, 0xEE, np.uint8)
        off = (-raw.ctypes.data) % 16 + shift
        obs = raw[off:off + n * V * V * 3]
        rew = np.empty(n, np.float32); te = np.empty(n, np.uint8); tr = np.empty(n, np.uint8)
        p = lambda a: C.c_void_p(a.ctypes.data)
        assert L.ta_decode_packed_host(p(words), p(status), n, V, p(obs), p(rew), p(te), p(tr)) == 0
        cells = ((words.reshape(-1)[:, None] >> (2 * np.arange(16, dtype=np.uint32))) & 3).reshape(ntiles * 32, V * V)[:n]
        assert np.array_equal(obs.reshape(n, V * V, 3), enc[cells]), (V, n, shift)
        assert np.array_equal(rew, rlut[status[:n] & 7]) and np.array_equal(te, (status[:n] >> 3) & 1) and np.array_equal(tr, (status[:n] >> 4) & 1)
        assert (raw[:off] == 0xEE).all() and (raw[off + n * V * V * 3:] == 0xEE).all()   # nothing outside the caller's array
    assert L.ta_decode_packed_host(None, None, 4, 17, None, None, None, None) == -1
