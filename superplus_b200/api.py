"""ctypes binding of libgcgpu.so (include/gcgpu.h) for tests, bench.py and the multi-GPU driver.

The product is the C-ABI library plus the C shims in superplus_b200/gap_closer/ (the reference
is a C program); this module is only the Python-side handle on the same entry points.  It
fails loudly when the CUDA library is missing or no B200 is visible — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgcgpu.so")

SOFTCLIP, LEADING_INDEL, INDEL, IGNORE = 0, 1, 2, 3
SW_ASIS, SW_FIXED = 0, 1

HIT_DTYPE = np.dtype([("read", "<i4"), ("pos", "<i4"), ("tid", "<i4"), ("cpos_flags", "<u4")])
SWRES_DTYPE = np.dtype([("score", "<i4"), ("alignment_offset", "<i4"), ("has_softclip", "<i4"),
                        ("bt_tidx", "<i4"), ("bt_qidx", "<i4"), ("n_cigar", "<i4"), ("cigar_off", "<i8")])
RUN_DTYPE = np.dtype([("tid", "<i4"), ("n_fwd", "<i4"), ("n_bwd", "<i4"), ("pad", "<i4"),
                      ("first_fwd", "<u8"), ("last_fwd", "<u8"), ("first_bwd", "<u8"), ("last_bwd", "<u8")])
KMER_DTYPE = np.dtype([("kseq", "<u8"), ("hs_id", "<i4"), ("tid", "<i4"), ("pos", "<i4"), ("flag", "<u2"), ("kmer_len", "<i2")])

EXPORTS = [
    "gcg_device_count", "gcg_init", "gcg_destroy", "gcg_last_error", "gcg_set_host_threads", "gcg_stream", "gcg_sync", "gcg_warmup",
    "gcg_prof_enable", "gcg_prof_reset", "gcg_prof_report", "gcg_launch_count", "gcg_ubench_int16", "gcg_ubench_hbm", "gcg_ubench_gather",
    "gcg_seqs_upload", "gcg_seqs_upload_concat", "gcg_ascii_upload_concat", "gcg_seqs_pack", "gcg_ascii_free",
    "gcg_seqs_free", "gcg_seqs_count", "gcg_seqs_bases", "gcg_seqs_kmers", "gcg_host_pack_2bit",
    "gcg_chop_contigs", "gcg_table_build_seqs", "gcg_table_build", "gcg_table_free", "gcg_table_stats",
    "gcg_table_size", "gcg_table_dump", "gcg_table_clone", "gcg_table_merge_ont",
    "gcg_search_seqs", "gcg_hits_count", "gcg_hits_download", "gcg_hits_free", "gcg_search", "gcg_free",
    "gcg_search_compact", "gcg_search_seqs_compact", "gcg_hits_download_compact", "gcg_selftest_workers", "gcg_selftest_workers_stress", "gcg_host_alloc", "gcg_search_runs", "gcg_search_compact_packed", "gcg_search_runs_packed",
    "gcg_table_create_shared", "gcg_table_reset_shared", "gcg_table_shared_info", "gcg_search_seqs_remote",
    "gcg_sw_batch", "gcg_sw_batch_multi", "gcg_swbatch_upload", "gcg_swbatch_align", "gcg_swbatch_download", "gcg_swbatch_cells",
    "gcg_swbatch_path_counts", "gcg_swbatch_free",
    "gcg_kmer_owner", "gcg_seqs_tiles", "gcg_route_plan", "gcg_route_kmers", "gcg_route_keys", "gcg_route_records",
    "gcg_route_collect", "gcg_route_free", "gcg_table_create", "gcg_table_insert_records", "gcg_table_lookup_keys",
    "gcg_window_create", "gcg_window_ptr", "gcg_window_bytes", "gcg_window_export", "gcg_window_open", "gcg_window_close",
    "gcg_window_free", "gcg_peer_enable", "gcg_route_keys_direct", "gcg_table_lookup_keys_direct",
    "gcg_route_positions", "gcg_filter_shape", "gcg_filter_add_table", "gcg_filter_or", "gcg_route_plan_filtered",
]
MAX_PART = 16


class GcgError(RuntimeError):
    pass


class SWParams(C.Structure):
    _fields_ = [("type_c", C.c_int32),
                ("del_o", C.c_int32), ("del_e", C.c_int32), ("ins_o", C.c_int32), ("ins_e", C.c_int32),
                ("strategy", C.c_int32),
                ("border_kind", C.c_int32),
                ("b_del_o", C.c_int32), ("b_del_e", C.c_int32), ("b_ins_o", C.c_int32), ("b_ins_e", C.c_int32),
                ("mat", C.c_int32 * 64)]


def default_mat(type_c=5, match=1, mismatch=-5):
    """gc_graph.c:74-77,87-107"""
    m = np.full((type_c, type_c), mismatch, dtype=np.int32)
    np.fill_diagonal(m, match)
    return m


def make_sw_params(mat=None, del_o=2, del_e=1, ins_o=2, ins_e=1, strategy=SOFTCLIP, border=None) -> SWParams:
    if mat is None:
        mat = default_mat()
    mat = np.ascontiguousarray(mat, dtype=np.int32)
    p = SWParams()
    p.type_c = mat.shape[0]
    p.del_o, p.del_e, p.ins_o, p.ins_e = del_o, del_e, ins_o, ins_e
    p.strategy = strategy
    if border is None:
        border = (0, 0, 0, 0, 0) if strategy == SOFTCLIP else (1, del_o, del_e, ins_o, ins_e)
    p.border_kind, p.b_del_o, p.b_del_e, p.b_ins_o, p.b_ins_e = border
    for i, v in enumerate(mat.reshape(-1)):
        p.mat[i] = int(v)
    return p


_lib = None


def host_pack_2bit(seq: bytes) -> np.ndarray:
    """The host-side 2-bit packing of the search gather (no device needed): (len+31)//32 words."""
    L = load_library()
    a = np.frombuffer(seq, dtype=np.uint8)
    out = np.zeros((len(a) + 31) // 32, dtype=np.uint64)
    rc = L.gcg_host_pack_2bit(a.ctypes.data if len(a) else None, len(a), out.ctypes.data if len(out) else None)
    if rc:
        raise GcgError(L.gcg_last_error().decode())
    return out


def load_library(path: str = LIB_PATH):
    """Load libgcgpu.so.  Raises when it has not been built — never falls back.
    (GCG_LIB names another build of the same library, for A/B timing of kernel variants.)"""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("GCG_LIB", path)
    if not os.path.exists(path):
        raise GcgError("libgcgpu.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `make -C superplus_b200/csrc`; there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.gcg_last_error.restype = C.c_char_p
    L.gcg_init.argtypes = [C.c_int, C.POINTER(vp)]
    L.gcg_destroy.argtypes = [vp]
    L.gcg_set_host_threads.argtypes = [vp, C.c_int]
    L.gcg_stream.restype = vp
    L.gcg_stream.argtypes = [vp]
    L.gcg_sync.argtypes = [vp]
    L.gcg_warmup.argtypes = [vp]
    L.gcg_prof_enable.argtypes = [vp, C.c_int]
    L.gcg_prof_reset.argtypes = [vp]
    L.gcg_prof_report.argtypes = [vp, C.c_char_p, i64]
    L.gcg_launch_count.restype = i64
    L.gcg_launch_count.argtypes = [vp]
    L.gcg_ubench_int16.argtypes = [vp, C.POINTER(C.c_double)]
    L.gcg_ubench_hbm.argtypes = [vp, C.POINTER(C.c_double)]
    L.gcg_ubench_gather.argtypes = [vp, i64, C.POINTER(C.c_double)]
    L.gcg_seqs_upload.argtypes = [vp, vp, vp, i64, C.POINTER(vp)]
    L.gcg_seqs_upload_concat.argtypes = [vp, vp, vp, i64, C.POINTER(vp)]
    L.gcg_ascii_upload_concat.argtypes = [vp, vp, vp, i64, C.POINTER(vp)]
    L.gcg_seqs_pack.argtypes = [vp, vp, C.POINTER(vp)]
    L.gcg_ascii_free.argtypes = [vp]
    L.gcg_seqs_free.argtypes = [vp]
    for f in (L.gcg_seqs_count, L.gcg_seqs_bases):
        f.restype = i64
        f.argtypes = [vp]
    L.gcg_seqs_kmers.restype = i64
    L.gcg_seqs_kmers.argtypes = [vp, C.c_int]
    L.gcg_host_pack_2bit.argtypes = [vp, i64, vp]
    L.gcg_chop_contigs.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    L.gcg_table_build_seqs.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
    L.gcg_table_build.argtypes = [vp, vp, vp, i32, C.c_int, C.POINTER(vp)]
    L.gcg_table_free.argtypes = [vp]
    L.gcg_table_stats.argtypes = [vp, vp, vp]
    L.gcg_table_size.restype = i64
    L.gcg_table_size.argtypes = [vp, vp]
    L.gcg_table_dump.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    L.gcg_table_clone.argtypes = [vp, vp, C.POINTER(vp)]
    L.gcg_table_merge_ont.argtypes = [vp, vp, vp]
    L.gcg_search_seqs.argtypes = [vp, vp, vp, C.c_int, C.POINTER(vp)]
    L.gcg_hits_count.restype = i64
    L.gcg_hits_count.argtypes = [vp]
    L.gcg_hits_download.argtypes = [vp, vp, vp, i64]
    L.gcg_hits_free.argtypes = [vp]
    L.gcg_search.argtypes = [vp, vp, vp, vp, i64, C.c_int, C.POINTER(vp), C.POINTER(i64)]
    L.gcg_free.argtypes = [vp]
    L.gcg_search_runs.argtypes = [vp, vp, vp, vp, i64, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(i64)]
    L.gcg_search_compact_packed.argtypes = [vp, vp, vp, vp, vp, i64, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
    L.gcg_search_runs_packed.argtypes = [vp, vp, vp, vp, vp, i64, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(i64)]
    L.gcg_table_create_shared.argtypes = [vp, i64, C.c_int, C.POINTER(vp)]
    L.gcg_table_reset_shared.argtypes = [vp, vp, i64]
    L.gcg_table_shared_info.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), C.POINTER(C.c_uint32), vp]
    L.gcg_search_seqs_remote.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, i64, C.c_int, C.POINTER(vp)]
    L.gcg_host_alloc.restype = vp
    L.gcg_host_alloc.argtypes = [i64]
    L.gcg_search_compact.argtypes = [vp, vp, vp, vp, i64, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
    L.gcg_search_seqs_compact.argtypes = [vp, vp, vp, C.c_int, C.POINTER(vp)]
    L.gcg_hits_download_compact.argtypes = [vp, vp, vp, i64, vp]
    L.gcg_sw_batch.argtypes = [vp, C.POINTER(SWParams), C.c_int, vp, vp, vp, vp, i64, vp, C.POINTER(vp), C.POINTER(i64)]
    L.gcg_sw_batch_multi.argtypes = [vp, C.c_int, C.POINTER(SWParams), C.c_int, vp, vp, vp, vp, i64, vp, C.POINTER(vp), C.POINTER(i64)]
    L.gcg_swbatch_upload.argtypes = [vp, vp, vp, vp, vp, i64, C.POINTER(vp)]
    L.gcg_swbatch_align.argtypes = [vp, vp, C.POINTER(SWParams), C.c_int]
    L.gcg_swbatch_download.argtypes = [vp, vp, vp, C.POINTER(vp), C.POINTER(i64)]
    L.gcg_swbatch_cells.restype = i64
    L.gcg_swbatch_cells.argtypes = [vp]
    L.gcg_swbatch_path_counts.argtypes = [vp, vp]
    L.gcg_swbatch_free.argtypes = [vp]
    L.gcg_kmer_owner.argtypes = [C.c_uint64, C.c_int]
    L.gcg_seqs_tiles.restype = i64
    L.gcg_seqs_tiles.argtypes = [vp]
    L.gcg_route_plan.argtypes = [vp, vp, C.c_int, C.c_int, i64, i64, C.POINTER(vp), vp]
    L.gcg_route_kmers.restype = i64
    L.gcg_route_kmers.argtypes = [vp]
    L.gcg_route_keys.argtypes = [vp, vp, vp]
    L.gcg_route_records.argtypes = [vp, vp, vp]
    L.gcg_route_collect.argtypes = [vp, vp, vp, C.POINTER(vp)]
    L.gcg_route_free.argtypes = [vp]
    L.gcg_table_create.argtypes = [vp, i64, C.c_int, C.POINTER(vp)]
    L.gcg_table_insert_records.argtypes = [vp, vp, vp, i64]
    L.gcg_table_lookup_keys.argtypes = [vp, vp, vp, i64, vp]
    L.gcg_window_create.argtypes = [vp, i64, C.POINTER(vp)]
    L.gcg_window_ptr.restype = vp
    L.gcg_window_ptr.argtypes = [vp]
    L.gcg_window_bytes.restype = i64
    L.gcg_window_bytes.argtypes = [vp]
    L.gcg_window_export.argtypes = [vp, vp]
    L.gcg_window_open.argtypes = [vp, vp, C.POINTER(vp)]
    L.gcg_window_close.argtypes = [vp, vp]
    L.gcg_window_free.argtypes = [vp]
    L.gcg_peer_enable.argtypes = [vp, C.c_int]
    L.gcg_route_positions.restype = i64
    L.gcg_route_positions.argtypes = [vp]
    L.gcg_filter_shape.argtypes = [i64, C.POINTER(i64), C.POINTER(C.c_int)]
    L.gcg_filter_add_table.argtypes = [vp, vp, vp, i64, C.c_int]
    L.gcg_filter_or.argtypes = [vp, vp, vp, i64]
    L.gcg_route_plan_filtered.argtypes = [vp, vp, C.c_int, C.c_int, i64, i64, vp, i64, C.c_int, C.POINTER(vp), vp]
    L.gcg_route_keys_direct.argtypes = [vp, vp, vp, vp]
    L.gcg_table_lookup_keys_direct.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp]
    _lib = L
    return L


def _concat(seqs):
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    if len(seqs) and off[-1] > 0:
        buf = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs]))
    else:
        buf = np.zeros(1, np.uint8)
    return buf, off


class Context:
    """One per process and GPU (gcg_ctx)."""

    def __init__(self, device: int = 0, host_threads: int | None = None):
        self.L = load_library()
        self.h = C.c_void_p()
        self._chk(self.L.gcg_init(device, C.byref(self.h)))
        if host_threads:
            self._chk(self.L.gcg_set_host_threads(self.h, host_threads))

    def _chk(self, rc):
        if rc != 0:
            raise GcgError("libgcgpu error %d: %s" % (rc, self.L.gcg_last_error().decode()))

    def close(self):
        if self.h:
            self.L.gcg_destroy(self.h)
            self.h = C.c_void_p()

    # ---- plumbing ------------------------------------------------------------------------
    def stream_ptr(self) -> int:
        return int(self.L.gcg_stream(self.h) or 0)

    def sync(self):
        self._chk(self.L.gcg_sync(self.h))

    def launches(self) -> int:
        return int(self.L.gcg_launch_count(self.h))

    def prof(self, on: bool):
        self._chk(self.L.gcg_prof_enable(self.h, 1 if on else 0))

    def prof_reset(self):
        self._chk(self.L.gcg_prof_reset(self.h))

    def prof_report(self) -> dict:
        buf = C.create_string_buffer(1 << 16)
        self._chk(self.L.gcg_prof_report(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n = line.split()
            out[name] = (float(ms), int(n))
        return out

    def ubench_int16(self) -> float:
        v = C.c_double()
        self._chk(self.L.gcg_ubench_int16(self.h, C.byref(v)))
        return v.value

    def ubench_hbm(self) -> float:
        v = C.c_double()
        self._chk(self.L.gcg_ubench_hbm(self.h, C.byref(v)))
        return v.value

    def ubench_gather(self, table_bytes: int) -> float:
        v = C.c_double()
        self._chk(self.L.gcg_ubench_gather(self.h, int(table_bytes), C.byref(v)))
        return v.value

    # ---- sequences -----------------------------------------------------------------------
    def upload(self, seqs) -> "Seqs":
        """list of uint8 ASCII arrays -> device 2-bit stream (K1), via the concatenated entry point"""
        buf, off = _concat(seqs)
        return self.upload_concat(buf, off)

    def upload_concat(self, buf: np.ndarray, off: np.ndarray) -> "Seqs":
        h = C.c_void_p()
        self._chk(self.L.gcg_seqs_upload_concat(self.h, buf.ctypes.data, off.ctypes.data, len(off) - 1, C.byref(h)))
        return Seqs(self, h)

    def upload_ptrs(self, seqs) -> "Seqs":
        """same through the pointer-array entry point the C shims use"""
        arrs = [np.ascontiguousarray(s, dtype=np.uint8) for s in seqs]
        n = len(arrs)
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
        lens = np.array([len(a) for a in arrs], dtype=np.int32)
        h = C.c_void_p()
        self._chk(self.L.gcg_seqs_upload(self.h, C.cast(ptrs, C.c_void_p), lens.ctypes.data, n, C.byref(h)))
        return Seqs(self, h)

    def stage_ascii(self, seqs) -> "Ascii":
        buf, off = _concat(seqs)
        h = C.c_void_p()
        self._chk(self.L.gcg_ascii_upload_concat(self.h, buf.ctypes.data, off.ctypes.data, len(off) - 1, C.byref(h)))
        return Ascii(self, h)

    def pack(self, a: "Ascii") -> "Seqs":
        h = C.c_void_p()
        self._chk(self.L.gcg_seqs_pack(self.h, a.h, C.byref(h)))
        return Seqs(self, h)

    # ---- k-mers --------------------------------------------------------------------------
    def chop_contigs(self, contigs: "Seqs", lens, k: int, n_thread: int = 1):
        """-> list of KMER_DTYPE arrays, one per contig (the kmer_t records of def.h:58-66)"""
        n = len(lens)
        outs = [np.zeros(max(0, int(l) - k + 1), dtype=KMER_DTYPE) for l in lens]
        keep = [o if len(o) else np.zeros(1, dtype=KMER_DTYPE) for o in outs]
        ptrs = (C.c_void_p * max(n, 1))(*[o.ctypes.data for o in keep])
        nk = np.zeros(max(n, 1), dtype=np.int32)
        self._chk(self.L.gcg_chop_contigs(self.h, contigs.h, k, n_thread, C.cast(ptrs, C.c_void_p), nk.ctypes.data))
        assert all(int(nk[i]) == len(outs[i]) for i in range(n))
        return outs

    def table_build(self, contigs: "Seqs", k: int) -> "KmerTable":
        h = C.c_void_p()
        self._chk(self.L.gcg_table_build_seqs(self.h, contigs.h, k, C.byref(h)))
        return KmerTable(self, h, k)

    def search(self, table: "KmerTable", reads: "Seqs") -> np.ndarray:
        h = C.c_void_p()
        self._chk(self.L.gcg_search_seqs(self.h, table.h, reads.h, table.k, C.byref(h)))
        try:
            n = int(self.L.gcg_hits_count(h))
            out = np.zeros(n, dtype=HIT_DTYPE)
            self._chk(self.L.gcg_hits_download(self.h, h, out.ctypes.data, n))
        finally:
            self.L.gcg_hits_free(h)
        return out

    def search_device(self, table: "KmerTable", reads: "Seqs") -> int:
        """search leaving the anchors on the device (bench `value`); returns the anchor count"""
        h = C.c_void_p()
        self._chk(self.L.gcg_search_seqs(self.h, table.h, reads.h, table.k, C.byref(h)))
        n = int(self.L.gcg_hits_count(h))
        self.L.gcg_hits_free(h)
        return n

    def search_compact(self, table: "KmerTable", reads: "Seqs"):
        """device-resident compact search -> (anchors uint64[n], read_off int64[n_read + 1])"""
        h = C.c_void_p()
        self._chk(self.L.gcg_search_seqs_compact(self.h, table.h, reads.h, table.k, C.byref(h)))
        try:
            n = int(self.L.gcg_hits_count(h))
            anchors = np.zeros(n, dtype=np.uint64)
            read_off = np.zeros(reads.n + 1, dtype=np.int64)
            self._chk(self.L.gcg_hits_download_compact(self.h, h, anchors.ctypes.data if n else None, n, read_off.ctypes.data))
        finally:
            self.L.gcg_hits_free(h)
        return anchors, read_off

    def search_device_compact(self, table: "KmerTable", reads: "Seqs") -> int:
        """compact search leaving the anchors on the device (bench `value`); returns the anchor count"""
        h = C.c_void_p()
        self._chk(self.L.gcg_search_seqs_compact(self.h, table.h, reads.h, table.k, C.byref(h)))
        n = int(self.L.gcg_hits_count(h))
        self.L.gcg_hits_free(h)
        return n

    def search_host_compact(self, table: "KmerTable", reads):
        """the call the shim makes: host pointers in, pinned compact anchors + per-read offsets out (gcg_search_compact)"""
        arrs = [np.ascontiguousarray(s, dtype=np.uint8) for s in reads]
        n = len(arrs)
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
        lens = np.array([len(a) for a in arrs], dtype=np.int32)
        ap, rp, na = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._chk(self.L.gcg_search_compact(self.h, table.h, C.cast(ptrs, C.c_void_p), lens.ctypes.data, n, table.k, C.byref(ap), C.byref(rp), C.byref(na)))
        try:
            anchors = np.frombuffer((C.c_char * (na.value * 8)).from_address(ap.value), dtype=np.uint64).copy() if na.value else np.zeros(0, np.uint64)
            read_off = np.frombuffer((C.c_char * ((n + 1) * 8)).from_address(rp.value), dtype=np.int64).copy()
        finally:
            self.L.gcg_free(ap)
            self.L.gcg_free(rp)
        return anchors, read_off

    def search_runs(self, table: "KmerTable", reads):
        """N3: search + anchor grouping on the device (gcg_search_runs) -> (runs RUN_DTYPE[n_run], run_off int64[n_read + 1], n_anchor)"""
        arrs = [np.ascontiguousarray(s, dtype=np.uint8) for s in reads]
        n = len(arrs)
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
        lens = np.array([len(a) for a in arrs], dtype=np.int32)
        rp, op, nr, na = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        self._chk(self.L.gcg_search_runs(self.h, table.h, C.cast(ptrs, C.c_void_p), lens.ctypes.data, n, table.k, C.byref(rp), C.byref(op), C.byref(nr), C.byref(na)))
        try:
            runs = np.frombuffer((C.c_char * (nr.value * RUN_DTYPE.itemsize)).from_address(rp.value), dtype=RUN_DTYPE).copy() if nr.value else np.zeros(0, RUN_DTYPE)
            run_off = np.frombuffer((C.c_char * ((n + 1) * 8)).from_address(op.value), dtype=np.int64).copy()
        finally:
            self.L.gcg_free(rp)
            self.L.gcg_free(op)
        return runs, run_off, int(na.value)

    @staticmethod
    def pack_reads(reads, pinned: bool = False, gap_words: int = 0):
        """what a loader that packs on ingest leaves behind: (words, woff int64[n + 1], lens int32[n], keepalive);
        `gap_words` unused words between consecutive reads (the reads are then not one run)"""
        lens = np.array([len(r) for r in reads], dtype=np.int32)
        nw = (lens.astype(np.int64) + 31) // 32
        woff = np.zeros(len(reads) + 1, dtype=np.int64)
        np.cumsum(nw + gap_words, out=woff[1:])
        total = int(woff[-1]) + 2
        keep = PinnedArray((total,), np.uint64) if pinned else None
        words = keep.array if pinned else np.zeros(total, dtype=np.uint64)
        for i, r in enumerate(reads):
            if len(r):
                words[int(woff[i]):int(woff[i]) + int(nw[i])] = host_pack_2bit(np.ascontiguousarray(r, dtype=np.uint8).tobytes())
        return words, woff, lens, keep

    def search_host_compact_packed(self, table: "KmerTable", words, woff, lens):
        """gcg_search_compact_packed: reads already 2-bit packed by the caller"""
        n = len(lens)
        ap, rp, na = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._chk(self.L.gcg_search_compact_packed(self.h, table.h, words.ctypes.data, woff.ctypes.data, lens.ctypes.data, n, table.k, C.byref(ap), C.byref(rp), C.byref(na)))
        try:
            anchors = np.frombuffer((C.c_char * (na.value * 8)).from_address(ap.value), dtype=np.uint64).copy() if na.value else np.zeros(0, np.uint64)
            read_off = np.frombuffer((C.c_char * ((n + 1) * 8)).from_address(rp.value), dtype=np.int64).copy()
        finally:
            self.L.gcg_free(ap)
            self.L.gcg_free(rp)
        return anchors, read_off

    def search_host(self, table: "KmerTable", reads) -> np.ndarray:
        """the shim-facing call: host pointers in, pinned host anchors out (gcg_search)"""
        arrs = [np.ascontiguousarray(s, dtype=np.uint8) for s in reads]
        n = len(arrs)
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
        lens = np.array([len(a) for a in arrs], dtype=np.int32)
        return self.search_host_ptrs(table, ptrs, lens, n)

    def search_host_ptrs(self, table, ptrs, lens, n) -> np.ndarray:
        hp = C.c_void_p()
        nh = C.c_int64()
        self._chk(self.L.gcg_search(self.h, table.h, C.cast(ptrs, C.c_void_p), lens.ctypes.data, n, table.k, C.byref(hp), C.byref(nh)))
        try:
            if nh.value:
                out = np.frombuffer((C.c_char * (nh.value * HIT_DTYPE.itemsize)).from_address(hp.value), dtype=HIT_DTYPE).copy()
            else:
                out = np.zeros(0, dtype=HIT_DTYPE)
        finally:
            self.L.gcg_free(hp)
        return out

    # ---- partitioned table (device pointers; the exchange between the calls is the caller's) ----
    def route_plan(self, seqs: "Seqs", k: int, n_part: int, tile_begin: int, tile_end: int, prefilter=None) -> "Route":
        """prefilter = (device pointer, n_words, k3): route only the positions whose k-mer passes it"""
        h = C.c_void_p()
        counts = np.zeros(MAX_PART, dtype=np.int64)
        if prefilter is None:
            self._chk(self.L.gcg_route_plan(self.h, seqs.h, k, n_part, tile_begin, tile_end, C.byref(h), counts.ctypes.data))
        else:
            ptr, n_words, k3 = prefilter
            self._chk(self.L.gcg_route_plan_filtered(self.h, seqs.h, k, n_part, tile_begin, tile_end, C.c_void_p(int(ptr)), int(n_words), int(k3),
                                                     C.byref(h), counts.ctypes.data))
        return Route(self, h, counts[:n_part].copy())

    def filter_shape(self, n_keys: int):
        nw, k3 = C.c_int64(), C.c_int()
        self._chk(self.L.gcg_filter_shape(int(n_keys), C.byref(nw), C.byref(k3)))
        return int(nw.value), int(k3.value)

    def filter_or(self, d_words: int, d_other: int, n_words: int):
        self._chk(self.L.gcg_filter_or(self.h, C.c_void_p(int(d_words)), C.c_void_p(int(d_other)), int(n_words)))

    def window(self, n_bytes: int) -> "Window":
        h = C.c_void_p()
        self._chk(self.L.gcg_window_create(self.h, int(n_bytes), C.byref(h)))
        return Window(self, h)

    def window_open(self, handle: bytes) -> int:
        """map another process's window (its 64-byte IPC handle) -> device pointer usable by this context"""
        buf = C.create_string_buffer(handle, 64)
        p = C.c_void_p()
        self._chk(self.L.gcg_window_open(self.h, buf, C.byref(p)))
        return int(p.value)

    def window_close(self, ptr: int):
        self._chk(self.L.gcg_window_close(self.h, C.c_void_p(ptr)))

    def table_create_shared(self, n_records: int, k: int) -> "KmerTable":
        """owner-side partition in one cudaMalloc block that peer GPUs map and probe over NVLink"""
        h = C.c_void_p()
        self._chk(self.L.gcg_table_create_shared(self.h, int(n_records), k, C.byref(h)))
        return KmerTable(self, h, k)

    def search_remote(self, reads: "Seqs", k: int, key_ptrs, val_ptrs, n_buckets, filter_ptr: int, filter_words: int, filter_k3: int) -> "Hits":
        """one search kernel that probes the partitions of a hash-partitioned table where they lie (gcg_search_seqs_remote)"""
        n = len(key_ptrs)
        kp = (C.c_void_p * MAX_PART)(*[C.c_void_p(int(p)) for p in key_ptrs])
        vp_ = (C.c_void_p * MAX_PART)(*[C.c_void_p(int(p)) for p in val_ptrs])
        nb = np.zeros(MAX_PART, dtype=np.uint32)
        nb[:n] = n_buckets
        h = C.c_void_p()
        self._chk(self.L.gcg_search_seqs_remote(self.h, reads.h, k, n, C.cast(kp, C.c_void_p), C.cast(vp_, C.c_void_p), nb.ctypes.data,
                                                C.c_void_p(int(filter_ptr)), int(filter_words), int(filter_k3), C.byref(h)))
        return Hits(self, h)

    def table_create(self, n_records: int, k: int) -> "KmerTable":
        h = C.c_void_p()
        self._chk(self.L.gcg_table_create(self.h, int(n_records), k, C.byref(h)))
        return KmerTable(self, h, k)

    # ---- SW ------------------------------------------------------------------------------
    def sw_batch(self, P: SWParams, qrys, tgts, mode: int = SW_ASIS):
        """host-buffer form.  -> (results SWRES_DTYPE[n], list of cigar uint32 arrays)"""
        qbuf, qoff = _concat(qrys)
        tbuf, toff = _concat(tgts)
        n = len(qrys)
        res = np.zeros(n, dtype=SWRES_DTYPE)
        pool = C.c_void_p()
        npool = C.c_int64()
        self._chk(self.L.gcg_sw_batch(self.h, C.byref(P), mode, qbuf.ctypes.data, qoff.ctypes.data, tbuf.ctypes.data,
                                      toff.ctypes.data, n, res.ctypes.data, C.byref(pool), C.byref(npool)))
        try:
            cp = np.frombuffer((C.c_char * (npool.value * 4)).from_address(pool.value), dtype=np.uint32).copy() if npool.value else np.zeros(0, np.uint32)
        finally:
            self.L.gcg_free(pool)
        cigs = [cp[int(r["cigar_off"]):int(r["cigar_off"]) + int(r["n_cigar"])] for r in res]
        return res, cigs

    def sw_batch_multi(self, others, P: SWParams, qrys, tgts, mode: int = SW_ASIS):
        """the same batch sharded over this context and `others` (one per GPU): contiguous pair ranges of
        about equal cells, one host thread per context (gcg_sw_batch_multi)"""
        qbuf, qoff = _concat(qrys)
        tbuf, toff = _concat(tgts)
        n = len(qrys)
        res = np.zeros(n, dtype=SWRES_DTYPE)
        pool = C.c_void_p()
        npool = C.c_int64()
        ctxs = [self] + list(others)
        arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        self._chk(self.L.gcg_sw_batch_multi(C.cast(arr, C.c_void_p), len(ctxs), C.byref(P), mode, qbuf.ctypes.data, qoff.ctypes.data,
                                            tbuf.ctypes.data, toff.ctypes.data, n, res.ctypes.data, C.byref(pool), C.byref(npool)))
        try:
            cp = np.frombuffer((C.c_char * (npool.value * 4)).from_address(pool.value), dtype=np.uint32).copy() if npool.value else np.zeros(0, np.uint32)
        finally:
            self.L.gcg_free(pool)
        cigs = [cp[int(r["cigar_off"]):int(r["cigar_off"]) + int(r["n_cigar"])] for r in res]
        return res, cigs

    def swbatch_upload(self, qry2d: np.ndarray, tgt2d: np.ndarray) -> "SWBatch":
        n = qry2d.shape[0]
        qoff = (np.arange(n + 1, dtype=np.int64) * qry2d.shape[1])
        toff = (np.arange(n + 1, dtype=np.int64) * tgt2d.shape[1])
        return self.swbatch_upload_concat(np.ascontiguousarray(qry2d, dtype=np.uint8).reshape(-1), qoff,
                                          np.ascontiguousarray(tgt2d, dtype=np.uint8).reshape(-1), toff)

    def swbatch_upload_concat(self, qbuf, qoff, tbuf, toff) -> "SWBatch":
        h = C.c_void_p()
        qb = qbuf if len(qbuf) else np.zeros(1, np.uint8)
        tb = tbuf if len(tbuf) else np.zeros(1, np.uint8)
        self._chk(self.L.gcg_swbatch_upload(self.h, qb.ctypes.data, qoff.ctypes.data, tb.ctypes.data, toff.ctypes.data, len(qoff) - 1, C.byref(h)))
        return SWBatch(self, h, len(qoff) - 1)


class _Handle:
    _free = None

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h

    def free(self):
        if self.h:
            getattr(self.ctx.L, self._free)(self.h)
            self.h = C.c_void_p()


class Ascii(_Handle):
    _free = "gcg_ascii_free"


class Seqs(_Handle):
    _free = "gcg_seqs_free"

    @property
    def n(self):
        return int(self.ctx.L.gcg_seqs_count(self.h))

    @property
    def bases(self):
        return int(self.ctx.L.gcg_seqs_bases(self.h))

    def kmers(self, k):
        return int(self.ctx.L.gcg_seqs_kmers(self.h, k))

    @property
    def tiles(self):
        return int(self.ctx.L.gcg_seqs_tiles(self.h))


class Hits(_Handle):
    """device-resident anchor list (gcg_hits)"""
    _free = "gcg_hits_free"

    @property
    def n(self):
        return int(self.ctx.L.gcg_hits_count(self.h))

    def download(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=HIT_DTYPE)
        self.ctx._chk(self.ctx.L.gcg_hits_download(self.ctx.h, self.h, out.ctypes.data, len(out)))
        return out


class Window(_Handle):
    """exchange window: a device allocation other ranks store into over NVLink (gcg_window)"""
    _free = "gcg_window_free"

    @property
    def ptr(self) -> int:
        return int(self.ctx.L.gcg_window_ptr(self.h) or 0)

    @property
    def nbytes(self) -> int:
        return int(self.ctx.L.gcg_window_bytes(self.h))

    def export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self.ctx._chk(self.ctx.L.gcg_window_export(self.h, buf))
        return buf.raw


class Route(_Handle):
    """stable partition plan of the k-mers of one tile range by owner (gcg_route)"""
    _free = "gcg_route_free"

    def __init__(self, ctx, h, counts):
        super().__init__(ctx, h)
        self.counts = counts

    @property
    def kmers(self):
        """k-mers the plan routes (elements of the send buffer)"""
        return int(self.ctx.L.gcg_route_kmers(self.h))

    @property
    def positions(self):
        """k-mer start positions of the range"""
        return int(self.ctx.L.gcg_route_positions(self.h))

    def keys(self, d_send: int):
        self.ctx._chk(self.ctx.L.gcg_route_keys(self.ctx.h, self.h, d_send))

    def keys_direct(self, owner_ptrs, owner_off):
        """owner d's run of keys goes to owner_ptrs[d] + 8 * owner_off[d] (a window, possibly on a peer GPU)"""
        n = len(owner_ptrs)
        ptrs = (C.c_void_p * MAX_PART)(*[C.c_void_p(int(p)) for p in owner_ptrs])
        offs = np.zeros(MAX_PART, dtype=np.int64)
        offs[:n] = owner_off
        self.ctx._chk(self.ctx.L.gcg_route_keys_direct(self.ctx.h, self.h, C.cast(ptrs, C.c_void_p), offs.ctypes.data))

    def records(self, d_send: int):
        self.ctx._chk(self.ctx.L.gcg_route_records(self.ctx.h, self.h, d_send))

    def collect(self, d_answers: int) -> Hits:
        h = C.c_void_p()
        self.ctx._chk(self.ctx.L.gcg_route_collect(self.ctx.h, self.h, d_answers, C.byref(h)))
        return Hits(self.ctx, h)


class KmerTable(_Handle):
    _free = "gcg_table_free"

    def __init__(self, ctx, h, k):
        super().__init__(ctx, h)
        self.k = k

    def stats(self):
        out = np.zeros(4, dtype=np.int64)
        self.ctx._chk(self.ctx.L.gcg_table_stats(self.ctx.h, self.h, out.ctypes.data))
        return tuple(int(x) for x in out)

    def clone(self, ctx: "Context") -> "KmerTable":
        """slot-for-slot replica on the device of `ctx` (read batches sharded over the GPUs of a box)"""
        h = C.c_void_p()
        ctx._chk(ctx.L.gcg_table_clone(ctx.h, self.h, C.byref(h)))
        return KmerTable(ctx, h, self.k)

    def merge_ont(self, replica: "KmerTable"):
        """fold the ONT-side multiplicity collected by a replica into this table (stats over all reads)"""
        self.ctx._chk(self.ctx.L.gcg_table_merge_ont(self.ctx.h, self.h, replica.h))

    def reset_shared(self, n_records: int) -> bool:
        """empty a shared partition for a rebuild; False when its block is too small (create a new one)"""
        rc = self.ctx.L.gcg_table_reset_shared(self.ctx.h, self.h, int(n_records))
        if rc == -4:            # GCG_ERANGE
            return False
        self.ctx._chk(rc)
        return True

    def shared_info(self):
        """-> (block pointer, byte offset of the value array, bucket count, 64-byte IPC handle)"""
        p, off, nb = C.c_void_p(), C.c_int64(), C.c_uint32()
        buf = C.create_string_buffer(64)
        self.ctx._chk(self.ctx.L.gcg_table_shared_info(self.h, C.byref(p), C.byref(off), C.byref(nb), buf))
        return int(p.value), int(off.value), int(nb.value), buf.raw

    def insert_records(self, d_records: int, n: int):
        self.ctx._chk(self.ctx.L.gcg_table_insert_records(self.ctx.h, self.h, d_records, int(n)))

    def lookup_keys(self, d_keys: int, n: int, d_answers: int):
        self.ctx._chk(self.ctx.L.gcg_table_lookup_keys(self.ctx.h, self.h, d_keys, int(n), d_answers))

    def filter_add(self, d_words: int, n_words: int, k3: int):
        """set the filter bits of this table's anchoring keys (present exactly once)"""
        self.ctx._chk(self.ctx.L.gcg_filter_add_table(self.ctx.h, self.h, C.c_void_p(int(d_words)), int(n_words), int(k3)))

    def lookup_keys_direct(self, d_keys: int, src_count, answer_ptrs, answer_off):
        """keys of requester r (src_count[r] of them, in rank order) -> answers at answer_ptrs[r] + 8 * answer_off[r]"""
        n = len(src_count)
        cnt = np.ascontiguousarray(src_count, dtype=np.int64)
        ptrs = (C.c_void_p * MAX_PART)(*[C.c_void_p(int(p)) for p in answer_ptrs])
        offs = np.ascontiguousarray(answer_off, dtype=np.int64)
        self.ctx._chk(self.ctx.L.gcg_table_lookup_keys_direct(self.ctx.h, self.h, d_keys, n, cnt.ctypes.data,
                                                               C.cast(ptrs, C.c_void_p), offs.ctypes.data))

    def dump(self):
        """-> key, multi(1|2), tid, pos, rev sorted by key"""
        n = int(self.ctx.L.gcg_table_size(self.ctx.h, self.h))
        key = np.zeros(max(n, 1), np.uint64); multi = np.zeros(max(n, 1), np.int32); tid = np.zeros(max(n, 1), np.int32)
        pos = np.zeros(max(n, 1), np.int32); rev = np.zeros(max(n, 1), np.uint8)
        self.ctx._chk(self.ctx.L.gcg_table_dump(self.ctx.h, self.h, n, key.ctypes.data, multi.ctypes.data, tid.ctypes.data,
                                                pos.ctypes.data, rev.ctypes.data))
        o = np.argsort(key[:n], kind="stable")
        return key[:n][o], multi[:n][o], tid[:n][o], pos[:n][o], rev[:n][o]


class SWBatch(_Handle):
    _free = "gcg_swbatch_free"

    def __init__(self, ctx, h, n):
        super().__init__(ctx, h)
        self.n = n

    def align(self, P: SWParams, mode: int = SW_ASIS):
        self.ctx._chk(self.ctx.L.gcg_swbatch_align(self.ctx.h, self.h, C.byref(P), mode))

    def cells(self) -> int:
        return int(self.ctx.L.gcg_swbatch_cells(self.h))

    def path_counts(self):
        c = np.zeros(2, dtype=np.int64)
        self.ctx._chk(self.ctx.L.gcg_swbatch_path_counts(self.h, c.ctypes.data))
        return int(c[0]), int(c[1])

    def download(self):
        res = np.zeros(self.n, dtype=SWRES_DTYPE)
        pool = C.c_void_p()
        npool = C.c_int64()
        self.ctx._chk(self.ctx.L.gcg_swbatch_download(self.ctx.h, self.h, res.ctypes.data, C.byref(pool), C.byref(npool)))
        try:
            cp = np.frombuffer((C.c_char * (npool.value * 4)).from_address(pool.value), dtype=np.uint32).copy() if npool.value else np.zeros(0, np.uint32)
        finally:
            self.ctx.L.gcg_free(pool)
        cigs = [cp[int(r["cigar_off"]):int(r["cigar_off"]) + int(r["n_cigar"])] for r in res]
        return res, cigs


def expand_compact(anchors: np.ndarray, read_off: np.ndarray, contig_lens) -> np.ndarray:
    """compact anchors (gcgpu.h) -> the same anchors as HIT_DTYPE records, for comparison with gcg_search"""
    out = np.zeros(len(anchors), dtype=HIT_DTYPE)
    n_read = len(read_off) - 1
    out["read"] = np.repeat(np.arange(n_read, dtype=np.int32), np.diff(read_off))
    out["pos"] = (anchors >> np.uint64(36)).astype(np.int32)
    gpos = ((anchors >> np.uint64(2)) & np.uint64(0x3FFFFFFFF)).astype(np.int64)
    cbase = np.concatenate([[0], np.cumsum(np.asarray(contig_lens, dtype=np.int64))])
    tid = np.searchsorted(cbase, gpos, side="right") - 1
    out["tid"] = tid.astype(np.int32)
    out["cpos_flags"] = (((gpos - cbase[tid]) << 2) | (anchors & np.uint64(3)).astype(np.int64)).astype(np.uint32)
    return out


class PinnedArray:
    """numpy view of page-locked host memory from gcg_host_alloc (inputs that cross PCIe without a staging copy)"""

    def __init__(self, shape, dtype=np.uint8):
        self.L = load_library()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.p = self.L.gcg_host_alloc(max(n, 16))
        if not self.p:
            raise GcgError("gcg_host_alloc(%d) failed: %s" % (n, self.L.gcg_last_error().decode()))
        self.array = np.frombuffer((C.c_char * max(n, 16)).from_address(self.p), dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.p:
            self.array = None
            self.L.gcg_free(self.p)
            self.p = None


def split_reads_by_bases(lens, n_share: int):
    """Contiguous read ranges [b_i, b_{i+1}) with about equal bases each (the shim's split, ont.c of this
    repo): share i ends at the first read where the running base count reaches (i + 1) / n of the total."""
    lens = np.asarray(lens, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(lens)])
    total = int(cum[-1])
    bounds = [0]
    for i in range(1, n_share):
        bounds.append(max(bounds[-1], int(np.searchsorted(cum, (total * i + n_share - 1) // n_share, side="left"))))
    bounds.append(len(lens))
    return [min(b, len(lens)) for b in bounds]


class ReplicatedSearch:
    """ONT reads sharded by batch over the GPUs of one box inside one process (replicated
    table): the table built on ctxs[0] is cloned slot for slot onto the other contexts, every context
    searches a contiguous share of the reads from its own host thread (the library call releases the
    GIL), the anchors are concatenated in read order and the replicas' ONT-side multiplicity is folded
    back into the primary table, whose statistics then cover all reads.  Mirrors what gc_b200 does with
    GC_DEVICES=0,1,... (superplus_b200/gap_closer/ont.c)."""

    def __init__(self, ctxs, table: KmerTable):
        assert table.ctx is ctxs[0]
        self.ctxs = list(ctxs)
        self.tables = [table] + [table.clone(c) for c in self.ctxs[1:]]

    def search_host(self, reads) -> np.ndarray:
        import threading
        bounds = split_reads_by_bases([len(r) for r in reads], len(self.ctxs))
        parts = [None] * len(self.ctxs)
        errs = []

        def work(i):
            try:
                parts[i] = self.ctxs[i].search_host(self.tables[i], reads[bounds[i]:bounds[i + 1]])
                parts[i]["read"] += bounds[i]
            except Exception as e:      # noqa: BLE001 - re-raised on the calling thread
                errs.append(e)
        th = [threading.Thread(target=work, args=(i,)) for i in range(len(self.ctxs))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        for rep in self.tables[1:]:
            self.tables[0].merge_ont(rep)
        return np.concatenate(parts) if parts else np.zeros(0, dtype=HIT_DTYPE)

    def free(self):
        for rep in self.tables[1:]:
            rep.free()
        self.tables = self.tables[:1]
