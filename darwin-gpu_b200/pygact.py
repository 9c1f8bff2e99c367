"""pygact -- thin ctypes binding of libgact_b200.so (include/gact_b200.h).

The product's host side is C++ (darwin-gpu_b200/host); this module exists so that the
parity tests, ``__graft_entry__.smoke()`` and ``bench.py`` can drive the very same C ABI
from Python.  It contains no alignment logic and no fallback: if the shared library is
missing or no B200 is present, calls raise.

Reference interface mirrored here (file:line in the reference):
  * ``GactEngine(...)``            <- GPU_init / GPU_close        cuda_host.cu:193-258
  * ``GactEngine.align_tiles``     <- Align_Batch_GPU             cuda_host.cu:23-190
  * ``align_with_bt`` result list  <- AlignWithBT's queue layout  align.cpp:190-199,208
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GACT_LIB") or os.path.join(HERE, "libgact_b200.so")     # GACT_LIB: profiling builds (tools/)

GACT_OK = 0
SET_REF, SET_READS, SET_READS_RC, SET_AUX = 0, 1, 2, 3
MAX_INFLIGHT = 3          # GACT_MAX_INFLIGHT: async batches between submit() and wait()

TILE_DESC_DTYPE = np.dtype([("ref_off", "<i8"), ("query_off", "<i8"), ("ref_len", "<i4"),
                            ("query_len", "<i4"), ("ref_set", "u1"), ("query_set", "u1"),
                            ("reverse", "u1"), ("first", "u1"), ("reserved", "<u4")])
TILE_RESULT_DTYPE = np.dtype([("score", "<i4"), ("max_i", "<i4"), ("max_j", "<i4"),
                              ("n_states", "<i4"), ("i_steps", "<i4"), ("j_steps", "<i4")])
assert TILE_DESC_DTYPE.itemsize == 32 and TILE_RESULT_DTYPE.itemsize == 24


class Params(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32),
                ("gap_extend", C.c_int32), ("tile_size", C.c_int32), ("tile_overlap", C.c_int32),
                ("first_tile_score_threshold", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("tiles", C.c_uint64), ("cells", C.c_uint64), ("first_tiles", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("batches", C.c_uint64),
                ("kernel_ms", C.c_double), ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double)]


class GactError(RuntimeError):
    pass


EXPORTS = [
    "gact_abi_version", "gact_status_string", "gact_device_count", "gact_engine_create",
    "gact_engine_destroy", "gact_last_error", "gact_engine_upload", "gact_engine_seq_start",
    "gact_engine_set_length", "gact_engine_set_bits", "gact_engine_seq_has_exceptions", "gact_engine_states_pitch_words",
    "gact_engine_max_tiles", "gact_engine_align_tiles", "gact_engine_submit", "gact_engine_wait", "gact_engine_wait_view",
    "gact_engine_stage", "gact_engine_run_staged", "gact_engine_fetch_staged", "gact_engine_sync",
    "gact_engine_last_kernel_ms", "gact_engine_stats", "gact_engine_reset_stats",
    "gact_engine_set_kernel", "gact_engine_get_kernel", "gact_int_peak",
    "gact_engine_extend", "gact_engine_extend_supported", "gact_engine_extend_reserve", "gact_dsoft_reserve",
    "gact_engine_extend_submit", "gact_engine_extend_wait", "gact_engine_reserve_tiles", "gact_engine_tile_path_info", "gact_engine_set_chain_mode", "gact_engine_chain_info",
    "gact_dsoft_create", "gact_dsoft_destroy", "gact_dsoft_run", "gact_dsoft_last_kernel_ms", "gact_dsoft_submit", "gact_dsoft_wait",
    "gact_seed_table_build", "gact_seed_table_destroy", "gact_seed_table_info", "gact_seed_table_download",
    "gact_dsoft_create_from_table",
]

CALL_DTYPE = np.dtype([("ref_seq", "<i4"), ("query_seq", "<i4"), ("ref_pos", "<i4"), ("query_pos", "<i4"),
                       ("query_set", "u1"), ("reserved", "u1", (3,))])
ALIGNMENT_DTYPE = np.dtype([("ab", "<i4"), ("ae", "<i4"), ("bb", "<i4"), ("be", "<i4"), ("score", "<i4"),
                            ("first_tile_score", "<i4"), ("n_tiles", "<i4"), ("reserved", "<i4"), ("n_cells", "<i8")])
DSOFT_CAND_DTYPE = np.dtype([("query", "<i4"), ("seq", "<i4"), ("hit", "<u4"), ("offset", "<u4")])

_lib = None


def load():
    """Load libgact_b200.so; raises GactError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GactError(f"{LIB_PATH} not found: build it with __graft_entry__.build() "
                        "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.gact_abi_version.restype = i32
    L.gact_status_string.restype = C.c_char_p
    L.gact_status_string.argtypes = [i32]
    L.gact_device_count.restype = i32
    L.gact_engine_create.restype = i32
    L.gact_engine_create.argtypes = [C.POINTER(vp), i32, C.POINTER(Params), i32, vp]
    L.gact_engine_destroy.restype = None
    L.gact_engine_destroy.argtypes = [vp]
    L.gact_last_error.restype = C.c_char_p
    L.gact_last_error.argtypes = [vp]
    L.gact_engine_upload.restype = i32
    L.gact_engine_upload.argtypes = [vp, i32, i64, C.POINTER(C.c_char_p), C.POINTER(i64)]
    L.gact_engine_seq_start.restype = i64
    L.gact_engine_seq_start.argtypes = [vp, i32, i64]
    L.gact_engine_set_length.restype = i64
    L.gact_engine_set_length.argtypes = [vp, i32]
    L.gact_engine_set_bits.restype = i32
    L.gact_engine_set_bits.argtypes = [vp, i32]
    L.gact_engine_seq_has_exceptions.restype = i32
    L.gact_engine_seq_has_exceptions.argtypes = [vp, i32, i64]
    L.gact_engine_states_pitch_words.restype = i32
    L.gact_engine_states_pitch_words.argtypes = [vp]
    L.gact_engine_max_tiles.restype = i32
    L.gact_engine_max_tiles.argtypes = [vp]
    L.gact_engine_align_tiles.restype = i32
    L.gact_engine_align_tiles.argtypes = [vp, i32, vp, vp, vp]
    L.gact_engine_submit.restype = i32
    L.gact_engine_submit.argtypes = [vp, i32, vp]
    L.gact_engine_wait.restype = i32
    L.gact_engine_wait.argtypes = [vp, vp, vp]
    L.gact_engine_wait_view.restype = i32
    L.gact_engine_wait_view.argtypes = [vp, C.POINTER(i32), C.POINTER(vp), C.POINTER(vp)]
    L.gact_engine_stage.restype = i32
    L.gact_engine_stage.argtypes = [vp, i32, vp]
    L.gact_engine_run_staged.restype = i32
    L.gact_engine_run_staged.argtypes = [vp]
    L.gact_engine_fetch_staged.restype = i32
    L.gact_engine_fetch_staged.argtypes = [vp, vp, vp]
    L.gact_engine_sync.restype = i32
    L.gact_engine_sync.argtypes = [vp]
    L.gact_engine_last_kernel_ms.restype = C.c_double
    L.gact_engine_last_kernel_ms.argtypes = [vp]
    L.gact_engine_stats.restype = i32
    L.gact_engine_stats.argtypes = [vp, C.POINTER(Stats)]
    L.gact_engine_reset_stats.restype = i32
    L.gact_engine_reset_stats.argtypes = [vp]
    L.gact_engine_set_kernel.restype = i32
    L.gact_engine_set_kernel.argtypes = [vp, i32]
    L.gact_engine_get_kernel.restype = i32
    L.gact_engine_get_kernel.argtypes = [vp]
    L.gact_engine_extend.restype = i32
    L.gact_engine_extend.argtypes = [vp, i32, vp, vp]
    L.gact_engine_extend_supported.restype = i32
    L.gact_engine_extend_supported.argtypes = [vp]
    L.gact_engine_extend_submit.restype = i32
    L.gact_engine_extend_submit.argtypes = [vp, i32, vp]
    L.gact_engine_extend_wait.restype = i32
    L.gact_engine_extend_wait.argtypes = [vp, vp]
    L.gact_engine_set_chain_mode.restype = i32
    L.gact_engine_set_chain_mode.argtypes = [vp, i32]
    L.gact_engine_chain_info.restype = i32
    L.gact_engine_chain_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.gact_dsoft_create.restype = i32
    L.gact_dsoft_create.argtypes = [C.POINTER(vp), vp, vp, C.c_uint64, vp, C.c_uint64, i32, i32, C.c_uint32,
                                    C.c_uint32, i32, i32, i32]
    L.gact_dsoft_destroy.restype = None
    L.gact_dsoft_destroy.argtypes = [vp]
    L.gact_dsoft_run.restype = i32
    L.gact_dsoft_run.argtypes = [vp, i32, vp, vp, vp, i64, C.POINTER(i64)]
    L.gact_dsoft_submit.restype = i32
    L.gact_dsoft_submit.argtypes = [vp, i32, vp, vp, i64]
    L.gact_dsoft_wait.restype = i32
    L.gact_dsoft_wait.argtypes = [vp, vp, i64, C.POINTER(i64)]
    L.gact_dsoft_last_kernel_ms.restype = C.c_double
    L.gact_dsoft_last_kernel_ms.argtypes = [vp]
    L.gact_seed_table_build.restype = i32
    L.gact_seed_table_build.argtypes = [C.POINTER(vp), vp, C.c_char_p, C.c_uint32, i32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.gact_seed_table_destroy.restype = None
    L.gact_seed_table_destroy.argtypes = [vp]
    L.gact_seed_table_info.restype = i32
    L.gact_seed_table_info.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                       C.POINTER(C.c_double)]
    L.gact_seed_table_download.restype = i32
    L.gact_seed_table_download.argtypes = [vp, vp, vp]
    L.gact_dsoft_create_from_table.restype = i32
    L.gact_dsoft_create_from_table.argtypes = [C.POINTER(vp), vp, vp, i32, i32, i32]
    L.gact_int_peak.restype = i32
    L.gact_int_peak.argtypes = [i32, i32, C.POINTER(C.c_double)]
    _lib = L
    return L


def device_count():
    return load().gact_device_count()


def int_peak(kind, device=0):
    out = C.c_double(0)
    rc = load().gact_int_peak(device, kind, C.byref(out))
    if rc != GACT_OK:
        raise GactError(f"gact_int_peak: {load().gact_status_string(rc).decode()}")
    return out.value


def make_descs(n):
    return np.zeros(n, dtype=TILE_DESC_DTYPE)


def unpack_states(words, n_states):
    """2-bit packed states of one tile -> list of ints (1=D, 2=I, 3=M)."""
    w = np.asarray(words, dtype=np.uint32)
    k = np.arange(n_states)
    return ((w[k >> 4] >> (2 * (k & 15)).astype(np.uint32)) & 3).astype(np.int32).tolist()


class GactEngine:
    """One engine = one device + one stream (the reference's GPU_storage, gact.h:51-67)."""

    def __init__(self, match=1, mismatch=-1, gap_open=-1, gap_extend=-1, tile_size=320,
                 tile_overlap=120, first_tile_score_threshold=35, device=0, max_tiles=4096,
                 stream=None):
        self.L = load()
        self.params = Params(match, mismatch, gap_open, gap_extend, tile_size, tile_overlap,
                             first_tile_score_threshold)
        self.h = C.c_void_p()
        rc = self.L.gact_engine_create(C.byref(self.h), device, C.byref(self.params), max_tiles,
                                       C.c_void_p(stream) if stream else None)
        if rc != GACT_OK:
            msg = self.L.gact_last_error(None).decode()
            self.h = None
            raise GactError(f"gact_engine_create: {self.L.gact_status_string(rc).decode()}: {msg}")
        self.pitch = self.L.gact_engine_states_pitch_words(self.h)
        self.max_tiles = max_tiles
        self.et = tile_size - tile_overlap

    def close(self):
        if getattr(self, "h", None):
            self.L.gact_engine_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc, what):
        if rc != GACT_OK:
            raise GactError(f"{what}: {self.L.gact_status_string(rc).decode()}: "
                            f"{self.L.gact_last_error(self.h).decode()}")

    def upload(self, set_id, seqs):
        """seqs: list of bytes objects.  Returns the start offset of each sequence."""
        n = len(seqs)
        arr = (C.c_char_p * n)(*[bytes(s) for s in seqs])
        lens = (C.c_int64 * n)(*[len(s) for s in seqs])
        self._ck(self.L.gact_engine_upload(self.h, set_id, n, arr, lens), "gact_engine_upload")
        return [self.L.gact_engine_seq_start(self.h, set_id, i) for i in range(n)]

    def set_bits(self, set_id):
        return self.L.gact_engine_set_bits(self.h, set_id)

    def tile_path_info(self):
        a, b = C.c_int(0), C.c_int(0)
        self.L.gact_engine_tile_path_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self._ck(self.L.gact_engine_tile_path_info(self.h, C.byref(a), C.byref(b)), "gact_engine_tile_path_info")
        return {"inter_task": a.value, "handed_back": b.value}

    def seq_has_exceptions(self, set_id, i):
        return self.L.gact_engine_seq_has_exceptions(self.h, set_id, i)

    def set_kernel(self, variant):
        self._ck(self.L.gact_engine_set_kernel(self.h, variant), "gact_engine_set_kernel")

    def get_kernel(self):
        return self.L.gact_engine_get_kernel(self.h)

    def _bufs(self, n, want_states=True):
        res = np.zeros(n, dtype=TILE_RESULT_DTYPE)
        st = np.zeros((n, self.pitch), dtype=np.uint32) if want_states else None
        return res, st

    def align_tiles(self, descs, want_states=True):
        descs = np.ascontiguousarray(descs, dtype=TILE_DESC_DTYPE)
        n = len(descs)
        res, st = self._bufs(n, want_states)
        self._ck(self.L.gact_engine_align_tiles(self.h, n, descs.ctypes.data, res.ctypes.data,
                                                st.ctypes.data if want_states else None),
                 "gact_engine_align_tiles")
        return res, st

    def submit(self, descs):
        descs = np.ascontiguousarray(descs, dtype=TILE_DESC_DTYPE)
        self._last_n = getattr(self, "_last_n", [])
        self._ck(self.L.gact_engine_submit(self.h, len(descs), descs.ctypes.data), "gact_engine_submit")
        self._last_n.append(len(descs))

    def wait(self, out_res=None, out_st=None):
        """Results of the oldest outstanding batch; optionally into caller-provided arrays."""
        n = self._last_n.pop(0) if getattr(self, "_last_n", None) else 0
        if out_res is None:
            out_res, out_st = self._bufs(n)
        assert len(out_res) == n and out_res.flags.c_contiguous and out_st.flags.c_contiguous
        self._ck(self.L.gact_engine_wait(self.h, out_res.ctypes.data, out_st.ctypes.data), "gact_engine_wait")
        return out_res, out_st

    def stage(self, descs):
        descs = np.ascontiguousarray(descs, dtype=TILE_DESC_DTYPE)
        self._staged_n = len(descs)
        self._ck(self.L.gact_engine_stage(self.h, len(descs), descs.ctypes.data), "gact_engine_stage")

    def run_staged(self):
        self._ck(self.L.gact_engine_run_staged(self.h), "gact_engine_run_staged")

    def sync(self):
        self._ck(self.L.gact_engine_sync(self.h), "gact_engine_sync")

    def last_kernel_ms(self):
        return self.L.gact_engine_last_kernel_ms(self.h)

    def fetch_staged(self, want_states=True):
        res, st = self._bufs(self._staged_n, want_states)
        self._ck(self.L.gact_engine_fetch_staged(self.h, res.ctypes.data,
                                                 st.ctypes.data if want_states else None),
                 "gact_engine_fetch_staged")
        return res, st

    def extend_supported(self):
        return bool(self.L.gact_engine_extend_supported(self.h))

    def extend(self, calls):
        """Whole GACT() extensions on the device; calls: CALL_DTYPE array -> ALIGNMENT_DTYPE array."""
        calls = np.ascontiguousarray(calls, dtype=CALL_DTYPE)
        out = np.zeros(len(calls), dtype=ALIGNMENT_DTYPE)
        self._ck(self.L.gact_engine_extend(self.h, len(calls), calls.ctypes.data, out.ctypes.data), "gact_engine_extend")
        return out

    def extend_submit(self, calls):
        calls = np.ascontiguousarray(calls, dtype=CALL_DTYPE)
        self._ck(self.L.gact_engine_extend_submit(self.h, len(calls), calls.ctypes.data), "gact_engine_extend_submit")
        self._ext_n = getattr(self, "_ext_n", []) + [len(calls)]

    def extend_wait(self):
        n = self._ext_n.pop(0) if getattr(self, "_ext_n", None) else 0
        out = np.zeros(n, dtype=ALIGNMENT_DTYPE)
        self._ck(self.L.gact_engine_extend_wait(self.h, out.ctypes.data), "gact_engine_extend_wait")
        return out

    def set_chain_mode(self, mode):
        self._ck(self.L.gact_engine_set_chain_mode(self.h, mode), "gact_engine_set_chain_mode")

    def chain_info(self):
        m, c, l = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.L.gact_engine_chain_info(self.h, C.byref(m), C.byref(c), C.byref(l)), "gact_engine_chain_info")
        return {"mode": m.value, "ctas": c.value, "n_long": l.value}

    def stats(self):
        s = Stats()
        self._ck(self.L.gact_engine_stats(self.h, C.byref(s)), "gact_engine_stats")
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def reset_stats(self):
        self._ck(self.L.gact_engine_reset_stats(self.h), "gact_engine_reset_stats")


def align_with_bt(engine, ref_seq, query_seq, reverse, first):
    """One tile through the GPU engine, returned in AlignWithBT's queue layout
    (align.cpp:190-199,208): [score, (max_i, max_j if first), states...]."""
    return align_batch(engine, [ref_seq], [query_seq], [reverse], [first])[0]


def align_batch(engine, ref_seqs, query_seqs, reverses, firsts):
    """Batch of independent tiles given as byte strings (the argument shape of the
    reference's Align_Batch / Align_Batch_GPU, align.cpp:17-54, cuda_host.cu:23-34).
    Uploads them into the two AUX-style sets and returns one queue-layout list per tile."""
    r_starts = engine.upload(SET_REF, ref_seqs)
    q_starts = engine.upload(SET_AUX, query_seqs)
    n = len(ref_seqs)
    d = make_descs(n)
    d["ref_off"] = r_starts
    d["query_off"] = q_starts
    d["ref_len"] = [len(s) for s in ref_seqs]
    d["query_len"] = [len(s) for s in query_seqs]
    d["ref_set"] = SET_REF
    d["query_set"] = SET_AUX
    d["reverse"] = np.asarray(reverses, dtype=np.uint8)
    d["first"] = np.asarray(firsts, dtype=np.uint8)
    out = []
    for lo in range(0, n, engine.max_tiles):
        res, st = engine.align_tiles(d[lo:lo + engine.max_tiles])
        for k in range(len(res)):
            q = [int(res["score"][k])]
            if d["first"][lo + k]:
                q += [int(res["max_i"][k]), int(res["max_j"][k])]
            out.append(q + unpack_states(st[k], int(res["n_states"][k])))
    return out


class SeedTable:
    """Seed-position table built on the device from the concatenated, bin-padded reference string."""

    def __init__(self, engine, ref_bytes, kmer_size=14, seed_occurence_multiple=32, bin_size=64, window_size=4):
        self.eng = engine
        self.h = C.c_void_p()
        rc = engine.L.gact_seed_table_build(C.byref(self.h), engine.h, ref_bytes, len(ref_bytes), kmer_size,
                                            seed_occurence_multiple, bin_size, window_size)
        engine._ck(rc, "gact_seed_table_build")
        ie, npos, mo, ms = C.c_uint64(), C.c_uint32(), C.c_uint32(), C.c_double()
        engine._ck(engine.L.gact_seed_table_info(self.h, C.byref(ie), C.byref(npos), C.byref(mo), C.byref(ms)),
                   "gact_seed_table_info")
        self.index_entries, self.num_minimizers, self.max_occ, self.build_ms = ie.value, npos.value, mo.value, ms.value

    def download(self):
        index = np.zeros(self.index_entries, dtype=np.uint32)
        pos = np.zeros(max(self.num_minimizers, 1), dtype=np.uint32)
        self.eng._ck(self.eng.L.gact_seed_table_download(self.h, index.ctypes.data, pos.ctypes.data),
                     "gact_seed_table_download")
        return index, pos[:self.num_minimizers]

    def close(self):
        if self.h:
            self.eng.L.gact_seed_table_destroy(self.h)
            self.h = None


class Dsoft:
    """Device-side D-SOFT filter bound to an engine.  The seed table comes either from the host builder
    (index_ptr / pos_ptr host arrays) or from a SeedTable built on the device (table=...)."""

    def __init__(self, engine, index_ptr=None, index_entries=0, pos_ptr=None, n_pos=0, kmer_size=14, window_size=4,
                 bin_size=64, max_occ=32, num_seeds=800, threshold=21, max_candidates=1000000, table=None):
        self.eng = engine
        self.h = C.c_void_p()
        if table is not None:
            rc = engine.L.gact_dsoft_create_from_table(C.byref(self.h), engine.h, table.h, num_seeds, threshold,
                                                       max_candidates)
            engine._ck(rc, "gact_dsoft_create_from_table")
            return
        rc = engine.L.gact_dsoft_create(C.byref(self.h), engine.h, index_ptr, index_entries, pos_ptr, n_pos,
                                        kmer_size, window_size, bin_size, max_occ, num_seeds, threshold, max_candidates)
        engine._ck(rc, "gact_dsoft_create")

    def run(self, sets, seq_index, cap=1 << 16):
        sets = np.ascontiguousarray(sets, dtype=np.int32)
        seq_index = np.ascontiguousarray(seq_index, dtype=np.int64)
        while True:
            out = np.zeros(cap, dtype=DSOFT_CAND_DTYPE)
            n = C.c_int64(0)
            rc = self.eng.L.gact_dsoft_run(self.h, len(sets), sets.ctypes.data, seq_index.ctypes.data,
                                           out.ctypes.data, cap, C.byref(n))
            if rc == -3 and n.value > cap:
                cap = int(n.value)
                continue
            self.eng._ck(rc, "gact_dsoft_run")
            return out[:n.value]

    def submit(self, sets, seq_index, cap=1 << 16):
        sets = np.ascontiguousarray(sets, dtype=np.int32)
        seq_index = np.ascontiguousarray(seq_index, dtype=np.int64)
        self.eng._ck(self.eng.L.gact_dsoft_submit(self.h, len(sets), sets.ctypes.data, seq_index.ctypes.data, cap), "gact_dsoft_submit")
        self._caps = getattr(self, "_caps", []) + [cap]

    def wait(self):
        cap = self._caps.pop(0)
        out = np.zeros(cap, dtype=DSOFT_CAND_DTYPE)
        n = C.c_int64(0)
        self.eng._ck(self.eng.L.gact_dsoft_wait(self.h, out.ctypes.data, cap, C.byref(n)), "gact_dsoft_wait")
        return out[:n.value]

    def last_kernel_ms(self):
        return self.eng.L.gact_dsoft_last_kernel_ms(self.h)

    def close(self):
        if self.h:
            self.eng.L.gact_dsoft_destroy(self.h)
            self.h = None
