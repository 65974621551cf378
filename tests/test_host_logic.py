"""CPU: host-side pieces that need no GPU — synthetic data, contig splitting, the host C
containers of the shim library (hash.c, cigar.c, hash_func.c replacement files)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from superplus_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "superplus_b200", "_build", "libgcshim.so")


def test_synth_is_deterministic():
    a, b = synth.make_config("tiny"), synth.make_config("tiny")
    assert np.array_equal(a.scaffold, b.scaffold) and len(a.reads) == len(b.reads)
    assert all(np.array_equal(x, y) for x, y in zip(a.reads, b.reads))


def test_split_contigs_matches_reference_state_machine():
    s = np.frombuffer(b"NNACGTNNNNACnGGTTN", dtype=np.uint8)
    parts = [bytes(x) for x in synth.split_contigs(s)]
    assert parts == [b"", b"ACGT", b"AC", b"GGTT"]
    assert [bytes(x) for x in synth.split_contigs(np.frombuffer(b"ACGT", dtype=np.uint8))] == [b"ACGT"]
    assert synth.split_contigs(np.zeros(0, np.uint8)) == []


def test_sw_pairs_shape():
    q, t = synth.make_sw_pairs(3, 500, 120, seed=1)
    assert q.shape == (3, 500) and t.shape == (3, 120) and q.max() <= 3 and t.max() <= 3


needs_shim = pytest.mark.skipif(not os.path.exists(SHIM), reason="libgcshim.so not built")


@needs_shim
def test_shim_cigar_container():
    L = C.CDLL(SHIM)

    class Cigar(C.Structure):
        _fields_ = [("c", C.POINTER(C.c_uint32)), ("n", C.c_int32), ("m", C.c_int32)]

    L.cigar_init.restype = C.POINTER(Cigar)
    L.cigar_add.argtypes = [C.POINTER(Cigar), C.c_uint32]
    L.cigar2ref_len.argtypes = [C.POINTER(Cigar)]
    L.cigar2qry_len.argtypes = [C.POINTER(Cigar)]
    c = L.cigar_init()
    ops = [(5, 4), (10, 0), (0, 2), (2, 1), (3, 2), (7, 0), (1, 5)]
    for l, o in ops:
        L.cigar_add(c, (l << 4) | o)
    assert c.contents.n == 7 and c.contents.m == 8
    assert L.cigar2ref_len(c) == 20 and L.cigar2qry_len(c) == 24
    assert L.cigar_has_zero_size_element(c) == 1
    L.cigar_reverse(c)
    assert [c.contents.c[i] for i in range(7)] == [(l << 4) | o for l, o in ops[::-1]]
    out = L.cigar_init()
    L.cigar_unclip(c, out)
    assert out.contents.n == 5
    d = L.cigar_init()
    for l, o in [(3, 2), (0, 0), (4, 0), (2, 2)]:
        L.cigar_add(d, (l << 4) | o)
    L.cigar_cleanup(d, out)
    assert [out.contents.c[i] for i in range(out.contents.n)] == [(4 << 4) | 0, (2 << 4) | 2]


@needs_shim
def test_shim_blizzard_hash_matches_fixture():
    L = C.CDLL(SHIM)
    L.blizzard_hash_func.restype = C.c_uint64
    L.blizzard_hash_func.argtypes = [C.c_char_p, C.c_int, C.c_int]
    L.hash_func_init()
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "sw_vectors.json")))["blizzard"]
    for w, vals in gold.items():
        assert [L.blizzard_hash_func(w.encode(), len(w), ht) for ht in (0, 1, 2)] == vals


@needs_shim
def test_shim_hash_set_semantics():
    """_xh_* contract of hash.h:60-110: insertion order pool, multiplicity, growth"""
    L = C.CDLL(SHIM)
    HF = C.CFUNCTYPE(C.c_uint64, C.c_void_p)
    EQ = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)
    hf = HF(lambda p: C.cast(p, C.POINTER(C.c_uint64))[0])
    eq = EQ(lambda a, b: int(C.cast(a, C.POINTER(C.c_uint64))[0] == C.cast(b, C.POINTER(C.c_uint64))[0]))

    class Item(C.Structure):
        pass
    Item._fields_ = [("key", C.c_void_p), ("val", C.c_void_p), ("hash_val", C.c_uint64), ("id", C.c_int64),
                     ("multi_deleted", C.c_uint32), ("next", C.POINTER(Item))]

    class XH(C.Structure):
        _fields_ = [("size", C.c_uint64), ("max", C.c_uint64), ("cnt", C.c_uint64), ("del_c", C.c_uint64), ("load_factor", C.c_double),
                    ("hash_func", C.c_void_p), ("is_equal_func", C.c_void_p), ("pool", C.POINTER(Item)), ("slots", C.c_void_p)]

    L._xh_init.restype = C.POINTER(XH)
    L._xh_init.argtypes = [C.c_int64, C.c_double, HF, EQ]
    L._xh_set_add.argtypes = [C.POINTER(XH), C.c_void_p]
    L._xh_set_search3.restype = C.POINTER(Item)
    L._xh_set_search3.argtypes = [C.POINTER(XH), C.c_void_p]
    h = L._xh_init(16, 0.75, hf, eq)
    assert h.contents.size == 256 and h.contents.max == 192
    keys = (C.c_uint64 * 1000)(*[(i * 7919) % 600 for i in range(1000)])
    rets = [L._xh_set_add(h, C.addressof(keys) + 8 * i) for i in range(1000)]
    assert rets.count(1) == 600 and rets.count(2) == 400 and h.contents.cnt == 600
    assert h.contents.size > 256                               # grew through next_prime
    probe = C.c_uint64(7919 % 600)
    it = L._xh_set_search3(h, C.addressof(probe))
    assert it and (it.contents.multi_deleted & 0x7FFFFFFF) == 2 and it.contents.id == 1
    missing = C.c_uint64(100000)
    assert not L._xh_set_search3(h, C.addressof(missing))
    assert [h.contents.pool[i].id for i in range(5)] == [0, 1, 2, 3, 4]


def test_host_pack_2bit_matches_the_oracle_packing():
    """gcg_host_pack_2bit (what the search gather does on the host) == (c >> 1) & 3 per base, 32
    bases per word, first base on top, zero tail — for every tail length, any byte value, and all
    four code paths (AVX-512 VBMI, AVX2, PEXT, portable multiply); the AVX-512 path also with every
    destination alignment inside a 64-byte line (head words, whole lines, pairs, single words, tail)."""
    import subprocess, sys
    rng = np.random.default_rng(3)

    def want(b):
        a = np.frombuffer(b, dtype=np.uint8)
        codes = ((a >> 1) & 3).astype(np.uint64)
        nw = (len(a) + 31) // 32
        pad = np.zeros(nw * 32, dtype=np.uint64)
        pad[:len(a)] = codes
        sh = np.uint64(62) - np.uint64(2) * np.arange(32, dtype=np.uint64)
        return (pad.reshape(nw, 32) << sh).sum(axis=1, dtype=np.uint64)

    cases = [b"", b"A", b"ACGT", b"ACGTNacgtn" * 7]
    cases += [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in (31, 32, 33, 63, 64, 65, 1000, 4097)]
    cases += [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n)) for n in (25, 96, 12345)]
    for c in cases:
        assert np.array_equal(api.host_pack_2bit(c), want(c)), len(c)
    # destinations at every 8-byte step of a 64-byte line, lengths around the 256-base blocks
    import ctypes as C
    L = api.load_library()
    L.gcg_host_pack_2bit.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    raw = np.zeros(4096 + 16, dtype=np.uint64)
    base = (-raw.ctypes.data // 8) % 8                                   # index of the first 64-byte aligned word
    for shift in range(8):
        for n in (0, 1, 32, 255, 256, 257, 511, 512, 700, 2048 + 31, 2048 + 96 + 5, 4000):
            c = bytes(rng.integers(0, 256, n, dtype=np.uint8))
            src = np.frombuffer(c, dtype=np.uint8) if n else np.zeros(1, np.uint8)
            raw[:] = np.uint64(0xDEADBEEFDEADBEEF)
            dst = raw[base + shift:]
            assert L.gcg_host_pack_2bit(src.ctypes.data, n, dst.ctypes.data) == 0
            nw = (n + 31) // 32
            assert np.array_equal(dst[:nw], want(c)), (shift, n)
            assert dst[nw] == np.uint64(0xDEADBEEFDEADBEEF) and (base + shift == 0 or raw[base + shift - 1] == np.uint64(0xDEADBEEFDEADBEEF))
    # every path in a fresh process (the choice is made once per process)
    code = ("import numpy as np, sys; sys.path.insert(0, %r); from superplus_b200 import api; "
            "b = bytes(range(256)) * 3 + b'ACGTTGCA'; print(','.join(str(int(x)) for x in api.host_pack_2bit(b)))" % ROOT)
    b = bytes(range(256)) * 3 + b"ACGTTGCA"
    for path in ("swar", "pext", "avx2", "avx512"):       # a path the CPU lacks falls back to the default one
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GCG_HOST_PACK=path), capture_output=True, text=True, check=True).stdout
        assert [int(x) for x in out.strip().split(",")] == [int(x) for x in want(b)], path


def test_split_reads_by_bases_covers_and_balances():
    """the share boundaries of the multi-device search (api.ReplicatedSearch, gap_closer/ont.c)"""
    from superplus_b200.api import split_reads_by_bases
    rng = np.random.default_rng(5)
    for n_reads in (0, 1, 2, 7, 100, 1000):
        lens = rng.integers(0, 60000, size=n_reads)
        for n in (1, 2, 3, 4, 8, 16):
            b = split_reads_by_bases(lens, n)
            assert len(b) == n + 1 and b[0] == 0 and b[-1] == n_reads
            assert all(x <= y for x, y in zip(b, b[1:]))
            if n_reads:
                shares = [int(lens[b[i]:b[i + 1]].sum()) for i in range(n)]
                assert sum(shares) == int(lens.sum())
                # no share exceeds its fair part by more than one read
                assert max(shares) <= -(-int(lens.sum()) // n) + int(lens.max())


def test_replicated_search_orchestration_with_a_host_double():
    """api.ReplicatedSearch (reads sharded over the GPUs of one process) driven with host doubles of the
    context and the table: shares are contiguous and searched on their own replica, anchors come back
    in read order with global read indices, and every replica is merged into the primary exactly once
    per search (the device side of clone / merge is covered by tests/test_gc_e2e_gpu.py)."""
    import threading
    log, lock = [], threading.Lock()

    class Tab:
        def __init__(self, ctx, name):
            self.ctx, self.name, self.k, self.seen, self.merged = ctx, name, 5, [], []

        def clone(self, ctx):
            return Tab(ctx, "replica@%s" % ctx.name)

        def merge_ont(self, rep):
            self.merged.append(rep.name)
            self.seen += rep.seen
            rep.seen = []                                  # the counts move

        def free(self):
            with lock:
                log.append(("free", self.name))

    class Ctx:
        def __init__(self, name):
            self.name = name

        def search_host(self, table, reads):
            # one "anchor" per read whose first byte is even: (local read index, pos 0, tid = length, 0)
            out = np.zeros(sum(1 for r in reads if len(r) and r[0] % 2 == 0), dtype=api.HIT_DTYPE)
            j = 0
            for i, r in enumerate(reads):
                if len(r) and r[0] % 2 == 0:
                    out[j] = (i, 0, len(r), 0); j += 1
            table.seen += [bytes(r[:4]) for r in reads]
            with lock:
                log.append(("search", self.name, table.name, len(reads)))
            return out

    rng = np.random.default_rng(12)
    reads = [rng.integers(0, 250, int(rng.integers(0, 400))).astype(np.uint8) for _ in range(57)]
    want = Ctx("solo").search_host(Tab(None, "t"), reads)
    for n in (1, 2, 3, 8):
        ctxs = [Ctx("c%d" % i) for i in range(n)]
        prim = Tab(ctxs[0], "primary")
        rep = api.ReplicatedSearch(ctxs, prim)
        assert [t.name for t in rep.tables] == ["primary"] + ["replica@c%d" % i for i in range(1, n)]
        del log[:]
        got = rep.search_host(reads)
        assert np.array_equal(got, want), n
        searched = sorted(e for e in log if e[0] == "search")
        assert [e[2] for e in searched] == sorted(t.name for t in rep.tables) and sum(e[3] for e in searched) == len(reads)
        assert prim.merged == [t.name for t in rep.tables[1:]]
        assert sorted(prim.seen) == sorted(bytes(r[:4]) for r in reads)          # every read reached exactly one replica
        rep.search_host(reads[:5])
        assert prim.merged == [t.name for t in rep.tables[1:]] * 2 and len(prim.seen) == len(reads) + 5
        rep.free()
        assert sorted(e[1] for e in log if e[0] == "free") == sorted(t for t in prim.merged[:n - 1])
        assert rep.tables == [prim]


def test_worker_pool_holds_one_job_at_a_time():
    """a synchronous job offered to the host worker pool while an asynchronous one
    (the gather of the next chunk) was still out overwrote its task counters and dropped the tasks nobody
    had taken yet.  Now the second job runs on the calling thread: every task of both jobs runs exactly once."""
    import ctypes as C
    from superplus_b200 import api
    L = api.load_library()
    L.gcg_selftest_workers.restype = C.c_int64
    L.gcg_selftest_workers.argtypes = [C.c_int, C.c_int, C.c_int]
    for nt, na, nb in ((4, 16, 4), (8, 64, 8), (2, 5, 3), (1, 7, 2)):
        assert L.gcg_selftest_workers(nt, na, nb) == na * 1000 + nb, (nt, na, nb)


def test_worker_pool_hands_every_task_out_once():
    """the pool's workers poll for the next job for a while before they sleep, and take tasks by compare-and-swap on a
    ticket that carries the job's generation: thousands of short jobs back to back, spinning, sleeping and mixed"""
    import ctypes as C
    from superplus_b200 import api
    L = api.load_library()
    L.gcg_selftest_workers_stress.restype = C.c_int64
    L.gcg_selftest_workers_stress.argtypes = [C.c_int, C.c_int, C.c_int]
    for nt, spin in ((2, -1), (4, 0), (8, 5), (16, 50), (3, 300)):
        assert L.gcg_selftest_workers_stress(nt, 6000, spin) == 0, (nt, spin)
