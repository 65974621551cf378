"""CPU: the parallel FASTQ loader of the B200 build (superplus_b200/gap_closer/rseq_fast.c,
SURVEY 8f row N1) against the reference's own sefq_load (rseq.c:307-374, compiled from the
reference tree under the name sefq_load_reference): same reads, lengths, capacities and bytes on
regular files and on the odd ones (CRLF, empty lines, missing final newline, partial record, a
0xFF byte, empty file, a file above the multi-thread threshold)."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "superplus_b200", "_build", "libgcshim.so")
pytestmark = pytest.mark.skipif(not os.path.exists(SHIM), reason="libgcshim.so not built")


class RSeq(C.Structure):
    _fields_ = [("b", C.c_void_p), ("q", C.c_void_p), ("l", C.c_int32), ("m", C.c_int32)]


class Pool(C.Structure):          # mp_t(rs), mp.h:18-28
    _fields_ = [("n", C.c_int64), ("mn", C.c_int64), ("m", C.c_int64), ("pool", C.POINTER(RSeq)),
                ("init_type_f", C.c_void_p), ("data4init_type", C.c_void_p)]


def load(fn_name, path):
    L = C.CDLL(SHIM)
    f = getattr(L, fn_name)
    f.restype = C.POINTER(Pool)
    f.argtypes = [C.c_char_p]
    p = f(path.encode())
    out = []
    for i in range(p.contents.n):
        r = p.contents.pool[i]
        out.append((r.l, r.m, C.string_at(r.b, r.l + 1), C.string_at(r.q, r.l + 1)))
    return out


def fastq(reads, nl=b"\n"):
    return b"".join(b"@r%d" % i + nl + b + nl + b"+" + nl + q + nl for i, (b, q) in enumerate(reads))


CASES = {
    "plain": fastq([(b"ACGTACGT", b"IIIIIIII"), (b"N", b"#"), (b"acgtn" * 40, b"I" * 200)]),
    "empty_file": b"",
    "crlf": fastq([(b"ACGT", b"IIII"), (b"GG", b"II")], nl=b"\r\n"),
    "empty_lines": b"@a\n\n+\n\n@b\nAC\n+\nII\n",
    "no_final_newline": fastq([(b"ACGT", b"IIII")]) + b"@x\nAC\n+\nII",
    "partial_record": fastq([(b"ACGT", b"IIII")]) + b"@x\nACGT\n+\n",
    "byte_ff": fastq([(b"ACGT", b"IIII")]) + b"@x\nAC\xffGT\n+\nIIIII\n" + fastq([(b"TT", b"II")]),
    "len_127_128_129": fastq([(b"A" * 127, b"I" * 127), (b"C" * 128, b"I" * 128), (b"G" * 129, b"I" * 129)]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_loader_matches_reference(tmp_path, name):
    path = str(tmp_path / (name + ".fq"))
    open(path, "wb").write(CASES[name])
    assert load("sefq_load", path) == load("sefq_load_reference", path)


def test_loader_matches_reference_large(tmp_path):
    """above the single-thread threshold: newline search and copies run on several threads"""
    rng = np.random.default_rng(3)
    reads = []
    for _ in range(600):
        n = int(rng.integers(1, 9000))
        reads.append((bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), n)), bytes(rng.integers(33, 74, n, dtype=np.uint8))))
    path = str(tmp_path / "big.fq")
    open(path, "wb").write(fastq(reads))
    assert os.path.getsize(path) > (1 << 20)
    got, want = load("sefq_load", path), load("sefq_load_reference", path)
    assert len(got) == 600 and got == want


def test_loader_packs_on_ingest(tmp_path):
    """pack on ingest (rest of SURVEY 8f row N1): the loader leaves, beside the rseq_t records, the 2-bit words of
    every read in the library's layout (what gcg_host_pack_2bit writes) — found through rseq_packed_lookup by the
    address of the read set; GC_NO_PACK_ON_INGEST switches it off"""
    from superplus_b200 import api
    rng = np.random.default_rng(5)
    reads = [(b"", b"")]
    for _ in range(300):
        n = int(rng.integers(1, 6000))
        reads.append((bytes(rng.choice(np.frombuffer(b"ACGTNacgtRY", np.uint8), n)), b"I" * n))
    reads += [(b"A" * 32, b"I" * 32), (b"C" * 33, b"I" * 33), (b"", b"")]
    path = str(tmp_path / "p.fq")
    open(path, "wb").write(fastq(reads))
    L = C.CDLL(SHIM)
    L.sefq_load.restype = C.POINTER(Pool)
    L.sefq_load.argtypes = [C.c_char_p]
    L.rseq_packed_lookup.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    p = L.sefq_load(path.encode())
    n = p.contents.n
    assert n == len(reads)
    words, woff = C.c_void_p(), C.c_void_p()
    assert L.rseq_packed_lookup(p, n, C.byref(words), C.byref(woff)) == 1
    assert L.rseq_packed_lookup(p, n + 1, C.byref(words), C.byref(woff)) == 0
    wo = np.frombuffer((C.c_char * ((n + 1) * 8)).from_address(woff.value), dtype=np.int64)
    wd = np.frombuffer((C.c_char * (int(wo[-1]) * 8 + 8)).from_address(words.value), dtype=np.uint64)
    for i, (b, _) in enumerate(reads):
        nw = (len(b) + 31) // 32
        assert int(wo[i + 1] - wo[i]) == nw
        assert np.array_equal(wd[int(wo[i]):int(wo[i]) + nw], api.host_pack_2bit(b)), i
    os.environ["GC_NO_PACK_ON_INGEST"] = "1"
    try:
        p2 = L.sefq_load(path.encode())
        assert L.rseq_packed_lookup(p2, n, C.byref(words), C.byref(woff)) == 0
    finally:
        os.environ.pop("GC_NO_PACK_ON_INGEST")
