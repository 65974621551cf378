"""TEST INFRASTRUCTURE — ctypes bindings for the checker libraries.

* `Oracle`  : oracle/_build/libgcoracle.so  (the CPU restatement, gc_oracle.c)
* `RefSW`   : oracle/_ref/libref_sw_{asis,fixed}.so (the reference's own sw.c/cigar.c, compiled
              unmodified by oracle/build_ref.sh)
* `run_ref_kmer` / `run_ref_gc` : the reference harness / CLI binaries in oracle/_ref/

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libgcoracle.so")
REF_DIR = os.path.join(HERE, "_ref")

SOFTCLIP, LEADING_INDEL, INDEL, IGNORE = 0, 1, 2, 3
CIGAR_OPS = "MIDNSHP=XB"


def cigar_str(c) -> str:
    if len(c) == 0:
        return "*"
    return "".join("%d%s" % (int(x) >> 4, CIGAR_OPS[int(x) & 0xF]) for x in c)


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


def make_params(mat=None, del_o=2, del_e=1, ins_o=2, ins_e=1, strategy=SOFTCLIP, border=None) -> SWParams:
    """border=None reproduces a freshly initialised aligner followed by one sw_set_parameter
    call (zero borders for SOFTCLIP, affine borders otherwise, sw.c:61-110)."""
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
    flat = mat.reshape(-1)
    for i, v in enumerate(flat):
        p.mat[i] = int(v)
    return p


def _as_char_p(a: np.ndarray):
    return a.ctypes.data_as(C.c_char_p)


class Oracle:
    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            raise RuntimeError("oracle library missing: build with `make -C oracle` (%s)" % path)
        L = C.CDLL(path)
        self.L = L
        L.gco_chop.restype = C.c_int64
        L.gco_chop.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        L.gco_table_build.restype = C.c_void_p
        L.gco_table_build.argtypes = [C.c_char_p, C.c_void_p, C.c_int32, C.c_int]
        L.gco_table_free.argtypes = [C.c_void_p]
        L.gco_table_size.restype = C.c_int64
        L.gco_table_size.argtypes = [C.c_void_p]
        L.gco_table_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.gco_table_dump.restype = C.c_int64
        L.gco_table_dump.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.gco_search.restype = C.c_int64
        L.gco_search.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64] + [C.c_void_p] * 7
        L.gco_sw_align.restype = C.c_int
        L.gco_sw_align.argtypes = [C.POINTER(SWParams), C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.gco_sw_edges.argtypes = [C.POINTER(SWParams), C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_void_p, C.c_void_p]
        L.gco_blizzard.restype = C.c_uint64
        L.gco_blizzard.argtypes = [C.c_char_p, C.c_int32, C.c_int]
        L.gco_cigar2ref_len.restype = C.c_int32
        L.gco_cigar2ref_len.argtypes = [C.c_void_p, C.c_int32]
        L.gco_cigar2qry_len.restype = C.c_int32
        L.gco_cigar2qry_len.argtypes = [C.c_void_p, C.c_int32]
        L.gco_hs_id.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
        L.gco_kseq_crc32.restype = C.c_uint32
        L.gco_kseq_crc32.argtypes = [C.c_uint64]
        L.gco_unanchored_segs.restype = C.c_int64
        L.gco_unanchored_segs.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]

    # ---- k-mers -------------------------------------------------------------------------
    def chop(self, seq: np.ndarray, k: int):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        n = max(0, len(seq) - k + 1)
        ks = np.zeros(n, dtype=np.uint64)
        rv = np.zeros(n, dtype=np.uint8)
        got = self.L.gco_chop(_as_char_p(seq), len(seq), k, ks.ctypes.data, rv.ctypes.data)
        assert got == n
        return ks, rv

    def hs_id(self, kseq: np.ndarray, n_thread: int) -> np.ndarray:
        """kmer_t.hs_id = crc32 of the canonical k-mer's 8 bytes % n_thread (kmer.c:88, crc32.h:70-81)"""
        kseq = np.ascontiguousarray(kseq, dtype=np.uint64)
        out = np.zeros(len(kseq), dtype=np.int32)
        self.L.gco_hs_id(kseq.ctypes.data, len(kseq), n_thread, out.ctypes.data)
        return out

    def unanchored_segs(self, read_len: int, anchored_pos) -> np.ndarray:
        """okseq->segs of one read after find_unankor_segs (ont.c:264-309): int32 [n_seg, 2] of {beg, end}"""
        a = np.zeros(max(read_len, 1), dtype=np.uint8)
        if len(anchored_pos):
            a[np.asarray(anchored_pos, dtype=np.int64)] = 1
        out = np.zeros((read_len // 2 + 2, 2), dtype=np.int32)
        n = self.L.gco_unanchored_segs(a.ctypes.data, read_len, out.ctypes.data, len(out))
        return out[:n].copy()

    @staticmethod
    def concat(seqs):
        lens = np.array([len(s) for s in seqs], dtype=np.int64)
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        buf = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs]) if len(seqs) else np.zeros(0, np.uint8)
        if len(buf) == 0:
            buf = np.zeros(1, np.uint8)
        return np.ascontiguousarray(buf), off

    def table_build(self, contigs, k: int):
        buf, off = self.concat(contigs)
        h = self.L.gco_table_build(_as_char_p(buf), off.ctypes.data, len(contigs), k)
        return h

    def table_free(self, h):
        self.L.gco_table_free(h)

    def table_stats(self, h):
        out = np.zeros(2, dtype=np.int64)
        self.L.gco_table_stats(h, out.ctypes.data)
        return int(out[0]), int(out[1])

    def table_dump(self, h):
        n = self.L.gco_table_size(h)
        key = np.zeros(n, np.uint64); multi = np.zeros(n, np.int32); tid = np.zeros(n, np.int32)
        pos = np.zeros(n, np.int32); rev = np.zeros(n, np.uint8)
        self.L.gco_table_dump(h, key.ctypes.data, multi.ctypes.data, tid.ctypes.data, pos.ctypes.data, rev.ctypes.data)
        o = np.argsort(key, kind="stable")
        return key[o], multi[o], tid[o], pos[o], rev[o]

    def search(self, h, reads, k: int):
        """-> dict(read,pos,tid,cpos,krev,orev) in (read,pos) order, (ont_total, ont_unique)"""
        buf, off = self.concat(reads)
        st = np.zeros(2, dtype=np.int64)
        cap = int(max(1, sum(max(0, len(r) - k + 1) for r in reads)))
        a = dict(read=np.zeros(cap, np.int64), pos=np.zeros(cap, np.int32), tid=np.zeros(cap, np.int32),
                 cpos=np.zeros(cap, np.int32), krev=np.zeros(cap, np.uint8), orev=np.zeros(cap, np.uint8))
        n = self.L.gco_search(h, _as_char_p(buf), off.ctypes.data, len(reads), k, cap,
                              a["read"].ctypes.data, a["pos"].ctypes.data, a["tid"].ctypes.data,
                              a["cpos"].ctypes.data, a["krev"].ctypes.data, a["orev"].ctypes.data, st.ctypes.data)
        return {k_: v[:n] for k_, v in a.items()}, (int(st[0]), int(st[1]))

    # ---- SW -----------------------------------------------------------------------------
    def sw_align(self, P: SWParams, qry: np.ndarray, tgt: np.ndarray, mode: int = 0, want_trace=False):
        qry = np.ascontiguousarray(qry, dtype=np.uint8); tgt = np.ascontiguousarray(tgt, dtype=np.uint8)
        q = qry if len(qry) else np.zeros(1, np.uint8)
        t = tgt if len(tgt) else np.zeros(1, np.uint8)
        out = np.zeros(8, dtype=np.int64)
        cap = len(qry) + len(tgt) + 8
        cig = np.zeros(cap, dtype=np.uint32)
        trace = np.zeros((len(tgt) + 1, len(qry) + 1), dtype=np.uint8) if want_trace else None
        rc = self.L.gco_sw_align(C.byref(P), len(qry), _as_char_p(q), len(tgt), _as_char_p(t), mode,
                                 out.ctypes.data, cig.ctypes.data, cap,
                                 trace.ctypes.data if want_trace else None)
        assert rc == 0
        res = dict(score=int(out[0]), offset=int(out[1]), softclip=int(out[2]), cigar=cig[:int(out[3])].copy(),
                   bt_tidx=int(out[4]), bt_qidx=int(out[5]), seg_len=int(out[6]))
        if want_trace:
            res["trace"] = trace
        return res

    def sw_edges(self, P: SWParams, qry: np.ndarray, tgt: np.ndarray):
        qry = np.ascontiguousarray(qry, dtype=np.uint8); tgt = np.ascontiguousarray(tgt, dtype=np.uint8)
        lc = np.zeros(len(tgt) + 1, np.int32); lr = np.zeros(len(qry) + 1, np.int32)
        self.L.gco_sw_edges(C.byref(P), len(qry), _as_char_p(qry), len(tgt), _as_char_p(tgt), lc.ctypes.data, lr.ctypes.data)
        return lc, lr

    def blizzard(self, s: bytes, hash_type: int = 1) -> int:
        return int(self.L.gco_blizzard(s, len(s), hash_type))

    def cigar_lens(self, cig: np.ndarray):
        cig = np.ascontiguousarray(cig, dtype=np.uint32)
        return (int(self.L.gco_cigar2ref_len(cig.ctypes.data, len(cig))),
                int(self.L.gco_cigar2qry_len(cig.ctypes.data, len(cig))))


class RefSW:
    """The reference's own aligner (sw.c) — mode 'asis' or 'fixed' (one-line traceback patch)."""

    def __init__(self, mode: str = "asis"):
        path = os.path.join(REF_DIR, "libref_sw_%s.so" % mode)
        if not os.path.exists(path):
            raise RuntimeError("reference build missing: %s (run oracle/build_ref.sh where /root/reference exists)" % path)
        L = C.CDLL(path)
        self.L = L
        L.refsw_new.restype = C.c_void_p
        L.refsw_set.argtypes = [C.c_void_p, C.c_int, C.c_void_p] + [C.c_int32] * 4 + [C.c_int]
        L.refsw_align.restype = C.c_int
        L.refsw_align.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.refsw_cell.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.refsw_score_row.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.refsw_free.argtypes = [C.c_void_p]
        L.refsw_bench.restype = C.c_double
        L.refsw_bench.argtypes = [C.c_int, C.c_int64, C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_int, C.c_void_p] + [C.c_int32] * 4 + [C.c_int, C.c_void_p]
        L.blizzard_hash_func.restype = C.c_uint64
        L.blizzard_hash_func.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.hash_func_init.restype = None
        self.h = C.c_void_p(L.refsw_new())
        self._hash_ready = False

    def set(self, mat=None, del_o=2, del_e=1, ins_o=2, ins_e=1, strategy=SOFTCLIP):
        if mat is None:
            mat = default_mat()
        mat = np.ascontiguousarray(mat, dtype=np.int32)
        self.L.refsw_set(self.h, mat.shape[0], mat.ctypes.data, del_o, del_e, ins_o, ins_e, strategy)

    def align(self, qry: np.ndarray, tgt: np.ndarray):
        qry = np.ascontiguousarray(qry, dtype=np.uint8); tgt = np.ascontiguousarray(tgt, dtype=np.uint8)
        q = qry if len(qry) else np.zeros(1, np.uint8)
        t = tgt if len(tgt) else np.zeros(1, np.uint8)
        out = np.zeros(4, dtype=np.int64)
        cap = len(qry) + len(tgt) + 8
        cig = np.zeros(cap, dtype=np.uint32)
        self.L.refsw_align(self.h, len(qry), _as_char_p(q), len(tgt), _as_char_p(t), out.ctypes.data, cig.ctypes.data, cap)
        return dict(score=int(out[0]), offset=int(out[1]), softclip=int(out[2]), cigar=cig[:int(out[3])].copy())

    def cell(self, i, j):
        o = np.zeros(8, np.int32)
        self.L.refsw_cell(self.h, i, j, o.ctypes.data)
        return dict(zip(("ms", "is", "ds", "ml", "dl", "il", "score", "status"), (int(x) for x in o)))

    def score_row(self, i, n):
        o = np.zeros(n, np.int32)
        self.L.refsw_score_row(self.h, i, n, o.ctypes.data)
        return o

    def bench(self, n_thread, qry2d: np.ndarray, tgt2d: np.ndarray, mat=None, del_o=2, del_e=1, ins_o=2, ins_e=1, strategy=SOFTCLIP):
        if mat is None:
            mat = default_mat()
        mat = np.ascontiguousarray(mat, dtype=np.int32)
        qry2d = np.ascontiguousarray(qry2d, dtype=np.uint8); tgt2d = np.ascontiguousarray(tgt2d, dtype=np.uint8)
        n = qry2d.shape[0]
        scores = np.zeros(n, np.int32)
        sec = self.L.refsw_bench(n_thread, n, qry2d.shape[1], _as_char_p(qry2d), tgt2d.shape[1], _as_char_p(tgt2d),
                                 mat.shape[0], mat.ctypes.data, del_o, del_e, ins_o, ins_e, strategy, scores.ctypes.data)
        return float(sec), scores

    def blizzard(self, s: bytes, hash_type: int = 1) -> int:
        if not self._hash_ready:
            self.L.hash_func_init()
            self._hash_ready = True
        return int(self.L.blizzard_hash_func(s, len(s), hash_type))

    def close(self):
        if self.h:
            self.L.refsw_free(self.h)
            self.h = None


HIT_DTYPE = np.dtype([("read", "<i4"), ("pos", "<i4"), ("tid", "<i4"), ("cpos", "<i4"), ("kflag", "<u2"), ("oflag", "<u2")])
TABLE_DTYPE = np.dtype([("kseq", "<u8"), ("multi", "<i4"), ("tid", "<i4"), ("pos", "<i4"), ("flag", "<u2"), ("klen", "<i2")])
CTGK_DTYPE = np.dtype([("kseq", "<u8"), ("tid", "<i4"), ("pos", "<i4"), ("flag", "<u2"), ("klen", "<i2")])


def have_ref() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("gc", "ref_kmer", "libref_sw_asis.so", "libref_sw_fixed.so"))


def run_ref_kmer(fa: str, fq: str, k: int, prefix: str, n_thread: int = 4, dump: int = 1):
    """Runs the reference harness; returns (json dict, hits, table, ctgk) with arrays per dump level."""
    exe = os.path.join(REF_DIR, "ref_kmer")
    subprocess.run([exe, fa, fq, str(n_thread), str(k), prefix, str(dump)], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    with open(prefix + ".json") as f:
        info = json.load(f)
    hits = np.fromfile(prefix + ".hits.bin", dtype=HIT_DTYPE) if dump >= 1 else None
    table = np.fromfile(prefix + ".table.bin", dtype=TABLE_DTYPE) if dump >= 2 else None
    ctgk = np.fromfile(prefix + ".ctgk.bin", dtype=CTGK_DTYPE) if dump >= 2 else None
    return info, hits, table, ctgk


def runs_from_hits(hits, n_read: int, contig_lens):
    """Restates the anchor grouping of map_ont2contigs (ctg_graph.c:600-656) + what ont_node_init keeps of a run
    (ctg_graph.c:93-181) on the oracle's anchors (dict of arrays read, pos, tid, cpos, krev, orev in (read,pos)
    order): a run = maximal stretch of consecutive anchors of one read on one contig; FORW anchors are those whose
    ONT strand flag equals the contig occurrence's (ctg_graph.c:617-618).  Returns (runs, run_off) with runs a dict of
    arrays: tid, n_fwd, n_bwd and first_fwd / last_fwd / first_bwd / last_bwd as compact anchor words
    (pos << 36 | (contig_base[tid] + cpos) << 2 | orev << 1 | krev; 0 when the direction has no anchor)."""
    read = np.asarray(hits["read"], dtype=np.int64)
    tid = np.asarray(hits["tid"], dtype=np.int64)
    n = len(read)
    cbase = np.concatenate([[0], np.cumsum(np.asarray(contig_lens, dtype=np.int64))])
    word = ((np.asarray(hits["pos"], dtype=np.uint64) << np.uint64(36)) |
            ((cbase[tid] + np.asarray(hits["cpos"], dtype=np.int64)).astype(np.uint64) << np.uint64(2)) |
            (np.asarray(hits["orev"], dtype=np.uint64) << np.uint64(1)) | np.asarray(hits["krev"], dtype=np.uint64))
    fwd = np.asarray(hits["orev"]) == np.asarray(hits["krev"])
    start = np.ones(n, dtype=bool)
    if n > 1:
        start[1:] = (read[1:] != read[:-1]) | (tid[1:] != tid[:-1])
    rid = np.cumsum(start) - 1                      # run index of every anchor
    n_run = int(rid[-1]) + 1 if n else 0
    out = {k_: np.zeros(n_run, dtype=np.uint64) for k_ in ("first_fwd", "last_fwd", "first_bwd", "last_bwd")}
    out["tid"] = tid[start].astype(np.int32)
    out["n_fwd"] = np.bincount(rid[fwd], minlength=n_run).astype(np.int32) if n else np.zeros(0, np.int32)
    out["n_bwd"] = np.bincount(rid[~fwd], minlength=n_run).astype(np.int32) if n else np.zeros(0, np.int32)
    idx = np.arange(n)
    for sel, a, b in ((fwd, "first_fwd", "last_fwd"), (~fwd, "first_bwd", "last_bwd")):
        r_, i_ = rid[sel], idx[sel]
        if len(r_):
            first = np.full(n_run, n, dtype=np.int64); np.minimum.at(first, r_, i_)
            last = np.full(n_run, -1, dtype=np.int64); np.maximum.at(last, r_, i_)
            have = last >= 0
            out[a][have] = word[first[have]]
            out[b][have] = word[last[have]]
    run_read = read[start]
    run_off = np.zeros(n_read + 1, dtype=np.int64)
    np.cumsum(np.bincount(run_read, minlength=n_read), out=run_off[1:])
    return out, run_off


def read_hsid(prefix: str) -> np.ndarray:
    """<prefix>.hsid.bin of a dump level 2 harness run: kmer_t.hs_id per contig position"""
    return np.fromfile(prefix + ".hsid.bin", dtype=np.int32)


def read_segs(prefix: str):
    """<prefix>.segs.bin of a dump level >= 1 harness run -> list (per read) of int32 [n_seg, 2] arrays"""
    raw = np.fromfile(prefix + ".segs.bin", dtype=np.int32)
    out, i = [], 0
    while i < len(raw):
        n = int(raw[i])
        out.append(raw[i + 1:i + 1 + 2 * n].reshape(n, 2).copy())
        i += 1 + 2 * n
    return out


def run_ref_gc(fa: str, fq: str, workdir: str, n_thread: int = 4):
    """Runs the reference CLI in `workdir` (it writes gc_fix1.fa, ont_link.txt, valid_ont_link.txt
    into CWD).  Returns stdout."""
    exe = os.path.join(REF_DIR, "gc")
    os.makedirs(workdir, exist_ok=True)
    r = subprocess.run([exe, os.path.abspath(fa), os.path.abspath(fq), str(n_thread), "out"], cwd=workdir,
                       check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
    return r.stdout.decode()
