"""CPU: the oracle against the reference binaries/libraries in oracle/_ref, live.  Skipped when
oracle/_ref has not been built (it is built in the container that holds /root/reference and
travels with the snapshot)."""
import os
import tempfile

import numpy as np
import pytest

from oracle import oracle as orc
from superplus_b200 import synth

pytestmark = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")


def test_sw_random_against_reference_library(oracle):
    rng = np.random.default_rng(99)
    for it in range(120):
        ql, tl = int(rng.integers(1, 160)), int(rng.integers(1, 160))
        t = rng.integers(0, 5, tl).astype(np.uint8)
        q = (synth.mutate(t, 0.2, rng) % 5) if rng.random() < 0.5 else rng.integers(0, 5, ql).astype(np.uint8)
        if len(q) == 0:
            q = np.array([0], np.uint8)
        strategy = int(rng.integers(0, 4))
        mat = rng.integers(-6, 4, size=(5, 5)).astype(np.int32)
        np.fill_diagonal(mat, rng.integers(1, 6, size=5))
        pen = tuple(int(x) for x in rng.integers(1, 7, size=4))
        for mode, name in ((0, "asis"), (1, "fixed")):
            R = orc.RefSW(name)
            R.set(mat, pen[0], pen[1], pen[2], pen[3], strategy)
            r = R.align(q, t)
            R.close()
            g = oracle.sw_align(orc.make_params(mat, pen[0], pen[1], pen[2], pen[3], strategy), q, t, mode)
            assert (r["score"], r["offset"], r["softclip"], orc.cigar_str(r["cigar"])) == \
                   (g["score"], g["offset"], g["softclip"], orc.cigar_str(g["cigar"])), (it, name)


def test_sw_trace_bits_match_reference_cells(oracle):
    """the 4 trace bits per cell the CUDA path spills are exactly the reference's (status, dl>1, il>1)"""
    rng = np.random.default_rng(5)
    t = rng.integers(0, 4, 70).astype(np.uint8)
    q = synth.mutate(t, 0.2, rng) % 4
    R = orc.RefSW("asis"); R.set(); R.align(q, t)
    g = oracle.sw_align(orc.make_params(), q, t, 0, want_trace=True)
    for i in range(1, len(t) + 1, 7):
        for j in range(1, len(q) + 1, 5):
            c = R.cell(i, j)
            st = {1: 1, 2: 2, 4: 3}[c["status"]]
            assert int(g["trace"][i, j]) == st | ((c["dl"] > 1) << 2) | ((c["il"] > 1) << 3)
    R.close()


def test_kmer_pipeline_against_reference_harness(oracle):
    inp = synth.make_config("repeats")
    with tempfile.TemporaryDirectory() as tmp:
        fa, fq = os.path.join(tmp, "r.fa"), os.path.join(tmp, "r.fq")
        synth.write_fasta(fa, inp.scaffold); synth.write_fastq(fq, inp.reads)
        info, hits, table, ctgk = orc.run_ref_kmer(fa, fq, 21, os.path.join(tmp, "out"), n_thread=2, dump=2)
        hsid, segs = orc.read_hsid(os.path.join(tmp, "out")), orc.read_segs(os.path.join(tmp, "out"))
    contigs = inp.contigs
    h = oracle.table_build(contigs, 21)
    key, multi, tid, pos, rev = oracle.table_dump(h)
    T = table[np.argsort(table["kseq"], kind="stable")]
    assert np.array_equal(T["kseq"], key) and np.array_equal(T["multi"], multi)
    assert np.array_equal(T["tid"], tid) and np.array_equal(T["pos"], pos) and np.array_equal(T["flag"], rev)
    H, ont = oracle.search(h, inp.reads, 21)
    assert (oracle.table_stats(h) + ont) == (info["scaf_total"], info["scaf_unique"], info["ont_total"], info["ont_unique"])
    assert np.array_equal(H["read"], hits["read"]) and np.array_equal(H["pos"], hits["pos"]) and np.array_equal(H["cpos"], hits["cpos"])
    ks = np.concatenate([oracle.chop(c, 21)[0] for c in contigs])
    assert np.array_equal(ks, ctgk["kseq"])
    assert np.array_equal(oracle.hs_id(ks, 2), hsid)                          # kmer.c:88
    assert len(segs) == len(inp.reads)
    for r, read in enumerate(inp.reads):                                       # ont.c:264-309
        assert np.array_equal(oracle.unanchored_segs(len(read), H["pos"][H["read"] == r]), segs[r]), r
    oracle.table_free(h)


def test_blizzard_against_reference(oracle):
    R = orc.RefSW("asis")
    for w in (b"chr1", b"scaffold_9", b"", b"MixedCase_name"):
        for ht in (0, 1, 2, 3):
            assert oracle.blizzard(w, ht) == R.blizzard(w, ht)
    R.close()


def test_run_grouping_against_the_reference_link_file(oracle):
    """oracle.runs_from_hits restates map_ont2contigs' grouping (ctg_graph.c:600-656) and ont_node_init's 20-fold
    majority rule (ctg_graph.c:93-181).  The reference prints every node of a read with more than one node into
    ont_link.txt as "tid <tab> n_okmers <tab> F|B" (add_ont_link2graph, ctg_graph.c:246-257): compare those lines
    for every read whose runs all pass the majority rule (a deleted node prints stale fields of an earlier read)."""
    for name in ("repeats", "small"):
        inp = synth.make_config(name)
        with tempfile.TemporaryDirectory() as tmp:
            fa, fq = os.path.join(tmp, "r.fa"), os.path.join(tmp, "r.fq")
            synth.write_fasta(fa, inp.scaffold); synth.write_fastq(fq, inp.reads)
            orc.run_ref_gc(fa, fq, os.path.join(tmp, "wd"), n_thread=2)
            text = open(os.path.join(tmp, "wd", "ont_link.txt")).read()
        ref = {}
        cur = None
        for line in text.splitlines():
            if line.startswith("> ONT"):
                cur = int(line.split()[2]); ref[cur] = []
            elif line.strip():
                a, b, c = line.split("\t")
                ref[cur].append((int(a), int(b), c))
        h = oracle.table_build(inp.contigs, 25)
        hits, _ = oracle.search(h, inp.reads, 25)
        oracle.table_free(h)
        runs, off = orc.runs_from_hits(hits, len(inp.reads), [len(c) for c in inp.contigs])
        checked = 0
        for r in range(len(inp.reads)):
            q0, q1 = int(off[r]), int(off[r + 1])
            assert (q1 > q0) == (r in ref), r                    # "> ONT r" is printed for every read with an anchor
            if q1 - q0 < 2:
                assert r not in ref or ref[r] == []
                continue
            nf, nb = runs["n_fwd"][q0:q1].astype(int), runs["n_bwd"][q0:q1].astype(int)
            ok_f, ok_b = nf > 20 * nb, nb > 20 * nf
            if not np.all(ok_f | ok_b):
                continue
            want = [(int(runs["tid"][q]), int(nf[q - q0] if ok_f[q - q0] else nb[q - q0]), "F" if ok_f[q - q0] else "B") for q in range(q0, q1)]
            assert ref[r] == want, (name, r)
            checked += 1
        assert checked > 0


def test_run_grouping_against_a_literal_loop(oracle):
    """the vectorised restatement against a loop that follows ctg_graph.c:600-656 statement by statement"""
    inp = synth.make_config("repeats")
    h = oracle.table_build(inp.contigs, 17)
    hits, _ = oracle.search(h, inp.reads, 17)
    oracle.table_free(h)
    runs, off = orc.runs_from_hits(hits, len(inp.reads), [len(c) for c in inp.contigs])
    cb = np.concatenate([[0], np.cumsum([len(c) for c in inp.contigs])])
    res = []
    for r in range(len(inp.reads)):
        cur = None
        for i in np.flatnonzero(hits["read"] == r):
            w = (int(hits["pos"][i]) << 36) | ((int(cb[hits["tid"][i]]) + int(hits["cpos"][i])) << 2) | (int(hits["orev"][i]) << 1) | int(hits["krev"][i])
            if cur is None or cur[0] != int(hits["tid"][i]):
                if cur:
                    res.append(tuple(cur))
                cur = [int(hits["tid"][i]), 0, 0, 0, 0, 0, 0]
            if hits["orev"][i] == hits["krev"][i]:
                if cur[1] == 0:
                    cur[3] = w
                cur[4] = w; cur[1] += 1
            else:
                if cur[2] == 0:
                    cur[5] = w
                cur[6] = w; cur[2] += 1
        if cur:
            res.append(tuple(cur))
    got = [(int(runs["tid"][i]), int(runs["n_fwd"][i]), int(runs["n_bwd"][i]), int(runs["first_fwd"][i]), int(runs["last_fwd"][i]),
            int(runs["first_bwd"][i]), int(runs["last_bwd"][i])) for i in range(len(runs["tid"]))]
    assert got == res and len(res) > 100
