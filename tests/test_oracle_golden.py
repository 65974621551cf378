"""CPU: the oracle (oracle/gc_oracle.c) against the committed golden fixtures, which were produced
by the reference itself (tests/golden/make_golden.py).  This is what pins the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from superplus_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("cfg,k", [("tiny", 25), ("repeats", 25), ("repeats", 17), ("tiny", 31)])
def test_kmer_oracle_matches_reference_fixture(oracle, cfg, k):
    g = np.load(os.path.join(GOLD, "kmer_%s_k%d.npz" % (cfg, k)))
    inp = synth.make_config(cfg)
    contigs = inp.contigs
    assert [len(contigs), len(inp.reads)] == list(g["counts"][:2])
    h = oracle.table_build(contigs, k)
    hits, ont = oracle.search(h, inp.reads, k)
    st = oracle.table_stats(h)
    oracle.table_free(h)
    assert list(st + ont) == list(g["stats"])
    assert np.array_equal(hits["read"], g["read"]) and np.array_equal(hits["pos"], g["pos"])
    assert np.array_equal(hits["tid"], g["tid"]) and np.array_equal(hits["cpos"], g["cpos"])
    assert np.array_equal(hits["krev"], g["kflag"]) and np.array_equal(hits["orev"], g["oflag"])


def test_hs_id_and_unanchored_segments_match_reference_fixture(oracle):
    """kmer_t.hs_id = crc32(canonical k-mer) % n_thread (kmer.c:88, crc32.h:70-81) and okseq->segs
    (find_unankor_segs, ont.c:264-309) as dumped by the reference harness"""
    g = np.load(os.path.join(GOLD, "kmer_aux_tiny_k25.npz"))
    inp = synth.make_config("tiny")
    ks = np.concatenate([oracle.chop(c, 25)[0] for c in inp.contigs])
    for nt in (3, 7):
        assert np.array_equal(oracle.hs_id(ks, nt), g["hs_id_nt%d" % nt]), nt
    h = oracle.table_build(inp.contigs, 25)
    hits, _ = oracle.search(h, inp.reads, 25)
    oracle.table_free(h)
    cnt, flat = g["seg_count"], g["seg_flat"]
    assert len(cnt) == len(inp.reads)
    at = 0
    for r, read in enumerate(inp.reads):
        segs = oracle.unanchored_segs(len(read), hits["pos"][hits["read"] == r])
        assert len(segs) == cnt[r] and np.array_equal(segs.reshape(-1), flat[at:at + 2 * cnt[r]]), r
        at += 2 * int(cnt[r])
    # known answers of the CRC itself: zlib's CRC-32 over the 8 little-endian bytes
    import zlib
    for v in (0, 1, 0x123456789ABCDEF, (1 << 50) - 1):
        assert int(oracle.L.gco_kseq_crc32(v)) & 0xffffffff == zlib.crc32(int(v).to_bytes(8, "little"))


def _vectors():
    return json.load(open(os.path.join(GOLD, "sw_vectors.json")))


def test_sw_oracle_matches_reference_fixture(oracle):
    for n, c in enumerate(_vectors()["cases"]):
        mat = np.array(c["mat"], dtype=np.int32).reshape(c["type_c"], c["type_c"])
        P = orc.make_params(mat, *c["pen"], strategy=c["strategy"])
        q, t = np.array(c["q"], np.uint8), np.array(c["t"], np.uint8)
        for mode, name in ((0, "asis"), (1, "fixed")):
            r = oracle.sw_align(P, q, t, mode)
            want = c[name]
            got = dict(score=r["score"], offset=r["offset"], softclip=r["softclip"], cigar=orc.cigar_str(r["cigar"]))
            assert got == want, (n, name)


def border_after(state, set_call, strategy, pen, qlen, tlen):
    """Host model of the reference aligner's border state (sw.c:61-110,134-162): returns the new
    state and the border tuple to align with.  state = dict(m_qry, m_tgt, border)"""
    if set_call:
        state["strategy"], state["pen"] = strategy, pen
        if strategy != orc.SOFTCLIP:
            state["border"] = (1, pen[0], pen[1], pen[2], pen[3])
    grown = False
    while state["m_qry"] < qlen + 2:
        state["m_qry"] *= 2; grown = True
    while state["m_tgt"] < tlen + 2:
        state["m_tgt"] *= 2; grown = True
    if grown:
        p = state["pen"]
        state["border"] = (0, 0, 0, 0, 0) if state["strategy"] == orc.SOFTCLIP else (1, p[0], p[1], p[2], p[3])
    return state


def test_sw_oracle_call_sequences_with_stale_borders(oracle):
    """one aligner, parameters changing between calls and the matrix growing: the border state
    model used by the sw.c shim must reproduce the reference results"""
    for steps in _vectors()["sequences"]:
        state = dict(m_qry=128, m_tgt=512, border=(0, 0, 0, 0, 0), strategy=orc.SOFTCLIP, pen=(0, 0, 0, 0))
        for st in steps:
            q, t = np.array(st["q"], np.uint8), np.array(st["t"], np.uint8)
            state = border_after(state, st["set"], st["strategy"], tuple(st["pen"]), len(q), len(t))
            pen = state["pen"]
            P = orc.make_params(orc.default_mat(), pen[0], pen[1], pen[2], pen[3], strategy=state["strategy"], border=state["border"])
            for mode, name in ((0, "asis"), (1, "fixed")):
                r = oracle.sw_align(P, q, t, mode)
                got = dict(score=r["score"], offset=r["offset"], softclip=r["softclip"], cigar=orc.cigar_str(r["cigar"]))
                assert got == st[name], name


def test_blizzard_hash_fixture(oracle):
    for w, vals in _vectors()["blizzard"].items():
        assert [oracle.blizzard(w.encode(), ht) for ht in (0, 1, 2)] == vals


def test_known_answer_vectors(oracle):
    """vectors measured on the reference at survey time"""
    code = lambda s: np.array([(c >> 1) & 3 for c in s.encode()], dtype=np.uint8)
    tgt = code("ACGTACGTTTGACCAGTAGGCATCGATCGGATTACAGATTACA")
    P = orc.make_params()
    for q, sc, oa, ca, of, cf in [("GACCAGTAGGCATCG", 15, 10, "15M", 10, "15M"), ("GACCAGTGGCATCG", 12, 11, "14M", 10, "7M1D7M"),
                                  ("GACCAGTAAGGCATCG", 13, 9, "16M", 10, "7M1I8M"), ("TTTTGACCAGTAGGCATCGTTTT", 11, 4, "23M", 7, "1I18M3I1D1M"),
                                  ("GACCAGTAGGCTTCGATCGGATT", 18, 0, "33M", 10, "11M1I1D11M")]:
        a, f = oracle.sw_align(P, code(q), tgt, 0), oracle.sw_align(P, code(q), tgt, 1)
        assert (a["score"], a["offset"], orc.cigar_str(a["cigar"])) == (sc, oa, ca)
        assert (f["score"], f["offset"], orc.cigar_str(f["cigar"])) == (sc, of, cf)


def test_cigar_lengths(oracle):
    cig = np.array([(5 << 4) | 4, (10 << 4) | 0, (2 << 4) | 1, (3 << 4) | 2, (7 << 4) | 0], dtype=np.uint32)
    assert oracle.cigar_lens(cig) == (20, 24)
