"""CPU-only: the C-ABI library builds, loads and exports every symbol of include/gact_b200.h."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "gact_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gact_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(pygact):
    if not os.path.exists(pygact.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(pygact.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gact_b200.h but not exported"
    assert sorted(pygact.EXPORTS) == syms


def test_abi_version_and_status_strings(pygact):
    L = pygact.load()
    assert L.gact_abi_version() == 1
    assert L.gact_status_string(0) == b"ok"
    assert L.gact_status_string(-1) == b"bad argument"


def test_struct_layouts(pygact):
    assert pygact.TILE_DESC_DTYPE.itemsize == 32
    assert pygact.TILE_RESULT_DTYPE.itemsize == 24
    assert ctypes.sizeof(pygact.Params) == 28
    assert ctypes.sizeof(pygact.Stats) == 64


def test_no_cpu_fallback_without_device(pygact):
    """On a box without a GPU the engine must refuse to come up (no silent CPU path)."""
    if pygact.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(pygact.GactError):
        pygact.GactEngine()


def test_bad_params_rejected_before_touching_cuda(pygact):
    L = pygact.load()
    h = ctypes.c_void_p()
    p = pygact.Params(1, -1, 1, -1, 320, 120, 35)          # positive gap_open
    assert L.gact_engine_create(ctypes.byref(h), 0, ctypes.byref(p), 16, None) == -1
    p = pygact.Params(1, -1, -1, -1, 4096, 120, 35)        # tile too large
    assert L.gact_engine_create(ctypes.byref(h), 0, ctypes.byref(p), 16, None) == -1
    p = pygact.Params(1, -1, -1, -1, 320, 320, 35)         # overlap >= tile
    assert L.gact_engine_create(ctypes.byref(h), 0, ctypes.byref(p), 16, None) == -1
    assert b"tile_size" in L.gact_last_error(None)


def test_unpack_states_roundtrip(pygact):
    rng = np.random.default_rng(0)
    st = rng.integers(1, 4, size=77)
    words = np.zeros(8, dtype=np.uint32)
    for k, s in enumerate(st):
        words[k >> 4] |= np.uint32(int(s) << (2 * (k & 15)))
    assert pygact.unpack_states(words, 77) == st.tolist()
