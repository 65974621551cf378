"""Helper of tests/test_shim_sw_gpu.py (run as a subprocess so that GC_DEVICES is seen by a fresh bridge):
aligns seeded pairs once through sw_align_batch (sw_batch.h) and once pair by pair through sw_align
(sw.h:70) on the same aligner, and prints both result lists as JSON."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_shim_sw_gpu import OPS, SHIM, SW, result  # noqa: E402  (the sw_t / cigar_t mirrors of the test module)


class Res(C.Structure):       # gcg_sw_result (include/gcgpu.h)
    _fields_ = [("score", C.c_int32), ("alignment_offset", C.c_int32), ("has_softclip", C.c_int32), ("bt_tidx", C.c_int32),
                ("bt_qidx", C.c_int32), ("n_cigar", C.c_int32), ("cigar_off", C.c_int64)]


def main():
    L = C.CDLL(SHIM)
    L.sw_init.restype = C.POINTER(SW)
    L.sw_set_parameter.argtypes = [C.POINTER(SW), C.c_int, C.c_void_p] + [C.c_int32] * 4 + [C.c_int]
    L.sw_align.argtypes = [C.POINTER(SW), C.c_int32, C.c_char_p, C.c_int32, C.c_char_p]
    L.sw_align_batch.argtypes = [C.POINTER(SW), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_int64)]
    L.gcg_free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    qs, ts = [], []
    for _ in range(37):
        t = rng.integers(0, 4, int(rng.integers(1, 400))).astype(np.uint8)
        q = np.concatenate([rng.integers(0, 4, int(rng.integers(0, 60))).astype(np.uint8), t[len(t) // 4:], rng.integers(0, 4, int(rng.integers(1, 60))).astype(np.uint8)])
        flip = rng.random(len(q)) < 0.1
        q[flip] = (q[flip] + 1) & 3
        qs.append(q); ts.append(t)
    mat = np.full((5, 5), -5, dtype=np.int32); np.fill_diagonal(mat, 1)
    sw = L.sw_init()
    L.sw_set_parameter(sw, 5, mat.ctypes.data, 2, 1, 2, 1, 0)
    # grow the matrix to its final size first: a batch shares one border state (sw.c:134-162)
    big_q, big_t = max(qs, key=len), max(ts, key=len)
    L.sw_align(sw, len(big_q), big_q.ctypes.data_as(C.c_char_p), len(big_t), big_t.ctypes.data_as(C.c_char_p))
    single = []
    for q, t in zip(qs, ts):
        assert L.sw_align(sw, len(q), q.ctypes.data_as(C.c_char_p), len(t), t.ctypes.data_as(C.c_char_p)) == 0
        single.append(result(sw))
    qbuf, tbuf = np.concatenate(qs), np.concatenate(ts)
    qoff = np.concatenate([[0], np.cumsum([len(q) for q in qs])]).astype(np.int64)
    toff = np.concatenate([[0], np.cumsum([len(t) for t in ts])]).astype(np.int64)
    res = (Res * len(qs))()
    pool, npool = C.POINTER(C.c_uint32)(), C.c_int64()
    assert L.sw_align_batch(sw, len(qs), qbuf.ctypes.data, qoff.ctypes.data, tbuf.ctypes.data, toff.ctypes.data, C.addressof(res),
                            C.byref(pool), C.byref(npool)) == 0
    batch = []
    for r in res:
        cig = "".join("%d%s" % (pool[r.cigar_off + i] >> 4, OPS[pool[r.cigar_off + i] & 15]) for i in range(r.n_cigar)) if r.n_cigar else "*"
        batch.append(dict(score=r.score, offset=r.alignment_offset, softclip=r.has_softclip, cigar=cig))
    L.gcg_free(pool)
    print(json.dumps({"single": single, "batch": batch}))


if __name__ == "__main__":
    main()
