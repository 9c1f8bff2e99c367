"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the CPU checkers:

* ``libgact_oracle.so``      -- our plain-C restatement (oracle/gact_oracle.c) of
  the reference's ``AlignWithBT`` (align.cpp:60-233) and ``GACT`` (gact.cpp:48-228).
* ``_ref/libalign_ref.so``   -- the reference's own ``align.cpp`` compiled in place
  (oracle/Makefile, oracle/ref_shim.cpp); optional, only where it was built.

Only tests/, ``__graft_entry__.smoke()`` and the CPU-baseline legs of bench.py may
import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libgact_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libalign_ref.so")
REF_DARWIN = os.path.join(HERE, "_ref", "darwin_ref")


class TileResult(C.Structure):
    _fields_ = [("score", C.c_int32), ("max_i", C.c_int32), ("max_j", C.c_int32),
                ("n_states", C.c_int32), ("i_steps", C.c_int32), ("j_steps", C.c_int32)]


class TileDesc(C.Structure):
    _fields_ = [("ref_off", C.c_int64), ("query_off", C.c_int64),
                ("ref_len", C.c_int32), ("query_len", C.c_int32),
                ("reverse", C.c_int32), ("first", C.c_int32)]


class GactResult(C.Structure):
    _fields_ = [("ab", C.c_int32), ("ae", C.c_int32), ("bb", C.c_int32), ("be", C.c_int32),
                ("score", C.c_int32), ("first_tile_score", C.c_int32),
                ("n_tiles", C.c_int32), ("n_cells", C.c_int64), ("n_columns", C.c_int32)]


class TileLog(C.Structure):
    _fields_ = [("ref_start", C.c_int32), ("query_start", C.c_int32), ("ref_len", C.c_int32),
                ("query_len", C.c_int32), ("reverse", C.c_int32), ("first", C.c_int32)]


TILE_DESC_DTYPE = np.dtype([("ref_off", "<i8"), ("query_off", "<i8"), ("ref_len", "<i4"),
                            ("query_len", "<i4"), ("reverse", "<i4"), ("first", "<i4")])
TILE_RESULT_DTYPE = np.dtype([("score", "<i4"), ("max_i", "<i4"), ("max_j", "<i4"),
                              ("n_states", "<i4"), ("i_steps", "<i4"), ("j_steps", "<i4")])


def build(ref=True):
    """Compile the checkers (oracle always; oracle/_ref when /root/reference exists)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


_oracle = None
_ref = None


def lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.oracle_align_tile.restype = C.c_int
        L.oracle_align_tile.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.POINTER(TileResult), C.c_void_p, C.c_int]
        L.oracle_align_batch.restype = C.c_int
        L.oracle_align_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_gact_extend.restype = C.c_int
        L.oracle_gact_extend.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(GactResult), C.c_void_p, C.c_int]
        _oracle = L
    return _oracle


def ref_available():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.ref_align_with_bt.restype = C.c_int
        L.ref_align_with_bt.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.ref_align_batch.restype = C.c_longlong
        L.ref_align_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p]
        _ref = L
    return _ref


REF_GPU_SO = os.path.join(HERE, "_ref", "libalign_ref_gpu.so")
_ref_gpu = None


def ref_gpu_available():
    return os.path.exists(REF_GPU_SO)


def ref_gpu_lib():
    """The reference's own GPU tile path (cuda_host.cu + cuda_header.h) built for sm_100a (oracle/ref_gpu_shim.cu)."""
    global _ref_gpu
    if _ref_gpu is None:
        L = C.CDLL(REF_GPU_SO)
        L.ref_gpu_init.restype = C.c_int
        L.ref_gpu_init.argtypes = [C.c_int] * 8
        L.ref_gpu_close.restype = None
        L.ref_gpu_align_batch.restype = C.c_longlong
        L.ref_gpu_align_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _ref_gpu = L
    return _ref_gpu


def align_tile(ref, query, scores=(1, -1, -1, -1), reverse=0, first=0, et=200):
    """Oracle for one tile.  Returns the reference's queue layout as a list:
    [score, (max_i, max_j if first), states...]  (align.cpp:190-199,208)."""
    ref = bytes(ref)
    query = bytes(query)
    res = TileResult()
    cap = 2 * max(et, 1) + 8
    st = np.zeros(cap, dtype=np.uint8)
    rc = lib().oracle_align_tile(ref, len(ref), query, len(query), *scores,
                                 int(reverse), int(first), int(et), None,
                                 C.byref(res), st.ctypes.data, cap)
    if rc != 0:
        raise RuntimeError("oracle_align_tile failed")
    out = [res.score]
    if first:
        out += [res.max_i, res.max_j]
    return out + st[:res.n_states].tolist(), res


def ref_align_tile(ref, query, scores=(1, -1, -1, -1), reverse=0, first=0, et=200):
    """The reference's own AlignWithBT through oracle/_ref (queue as a list)."""
    ref = bytes(ref)
    query = bytes(query)
    cap = 2 * max(et, 1) + 16
    out = np.zeros(cap, dtype=np.int32)
    n = ref_lib().ref_align_with_bt(ref, len(ref), query, len(query), *scores,
                                    int(reverse), int(first), int(et), out.ctypes.data, cap)
    return out[:n].tolist()


def align_batch(ref_buf, query_buf, descs, scores=(1, -1, -1, -1), et=200, max_len=320,
                n_threads=1, states_pitch=None):
    """Oracle over a batch of tiles.  ref_buf/query_buf: uint8 arrays; descs: TILE_DESC_DTYPE.
    Returns (results[TILE_RESULT_DTYPE], states[n, pitch] uint8)."""
    ref_buf = np.ascontiguousarray(ref_buf, dtype=np.uint8)
    query_buf = np.ascontiguousarray(query_buf, dtype=np.uint8)
    descs = np.ascontiguousarray(descs, dtype=TILE_DESC_DTYPE)
    n = len(descs)
    pitch = states_pitch or (2 * et + 8)
    res = np.zeros(n, dtype=TILE_RESULT_DTYPE)
    st = np.zeros((n, pitch), dtype=np.uint8)
    rc = lib().oracle_align_batch(ref_buf.ctypes.data, query_buf.ctypes.data, descs.ctypes.data, n,
                                  *scores, int(et), int(max_len), int(n_threads),
                                  res.ctypes.data, st.ctypes.data, pitch)
    if rc != 0:
        raise RuntimeError("oracle_align_batch failed")
    return res, st


def gact_extend(ref, query, ref_pos, query_pos, tile_size=320, tile_overlap=120, thr=35,
                scores=(1, -1, -1, -1), log_cap=0):
    ref = bytes(ref)
    query = bytes(query)
    out = GactResult()
    log = (TileLog * log_cap)() if log_cap else None
    rc = lib().oracle_gact_extend(ref, len(ref), query, len(query), tile_size, tile_overlap,
                                  ref_pos, query_pos, thr, *scores, C.byref(out),
                                  C.cast(log, C.c_void_p) if log_cap else None, log_cap)
    if rc != 0:
        raise RuntimeError("oracle_gact_extend failed")
    tiles = [(l.ref_start, l.query_start, l.ref_len, l.query_len, l.reverse, l.first)
             for l in (log[:min(out.n_tiles, log_cap)] if log_cap else [])]
    return out, tiles
