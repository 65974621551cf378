"""Generates the committed golden fixtures from the REFERENCE ITSELF (oracle/_ref, built by
oracle/build_ref.sh from /root/reference/gap_closer, unmodified).  Run in the build container:

    python tests/golden/make_golden.py

Outputs (small, committed):
  kmer_<cfg>_k<k>.npz   anchors (read,pos,tid,cpos,kflag,oflag) + the four stat integers, from ref_kmer
  gc_e2e.json           md5 of gc_fix1.fa / ont_link.txt / valid_ont_link.txt + stat lines, from gc
  kmer_aux_<cfg>_k<k>.npz  kmer_t.hs_id of every contig position for n_thread = 3 and 7 (crc32 % n_thread,
                        kmer.c:88) and okseq->segs of every read after find_unankor_segs (ont.c:264-309), from ref_kmer
  sw_vectors.json       sw_align results (score, offset, softclip, CIGAR) in as-is and fixed mode,
                        including call sequences that exercise the stale-border state
"""
import hashlib
import json
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc          # noqa: E402
from superplus_b200 import synth          # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def kmer_fixture(tmp, cfg, k):
    fa, fq, _ = synth.materialise(cfg, tmp)
    info, hits, _, _ = orc.run_ref_kmer(fa, fq, k, os.path.join(tmp, "%s_k%d" % (cfg, k)), n_thread=3, dump=1)
    np.savez_compressed(os.path.join(OUT, "kmer_%s_k%d.npz" % (cfg, k)),
                        read=hits["read"], pos=hits["pos"], tid=hits["tid"], cpos=hits["cpos"],
                        kflag=hits["kflag"].astype(np.uint8), oflag=hits["oflag"].astype(np.uint8),
                        stats=np.array([info["scaf_total"], info["scaf_unique"], info["ont_total"], info["ont_unique"]], dtype=np.int64),
                        counts=np.array([info["n_contigs"], info["n_reads"], info["n_ctg_kmers"], info["n_ont_kmers"]], dtype=np.int64))
    print("kmer", cfg, k, info["n_hits"], "anchors")


def aux_fixture(tmp, cfg, k):
    fa, fq, _ = synth.materialise(cfg, tmp)
    out = {}
    for nt in (3, 7):
        prefix = os.path.join(tmp, "%s_k%d_aux%d" % (cfg, k, nt))
        info, hits, _, ctgk = orc.run_ref_kmer(fa, fq, k, prefix, n_thread=nt, dump=2)
        out["hs_id_nt%d" % nt] = orc.read_hsid(prefix)
        assert len(out["hs_id_nt%d" % nt]) == len(ctgk)
        segs = orc.read_segs(prefix)
        flat = np.concatenate([s_.reshape(-1) for s_ in segs]) if segs else np.zeros(0, np.int32)
        cnt = np.array([len(s_) for s_ in segs], dtype=np.int32)
        if "seg_count" in out:
            assert np.array_equal(out["seg_count"], cnt) and np.array_equal(out["seg_flat"], flat), "segs depend on n_thread?"
        out["seg_count"], out["seg_flat"] = cnt, flat
    np.savez_compressed(os.path.join(OUT, "kmer_aux_%s_k%d.npz" % (cfg, k)), **out)
    print("aux", cfg, k, len(out["hs_id_nt3"]), "contig k-mers,", int(out["seg_count"].sum()), "segments")


def gc_fixture(tmp, cfg, nts=(1, 4)):
    fa, fq, _ = synth.materialise(cfg, tmp)
    res = {}
    for nt in nts:
        wd = os.path.join(tmp, "gc_%s_%d" % (cfg, nt))
        out = orc.run_ref_gc(fa, fq, wd, n_thread=nt)
        stat = [int(x) for x in re.findall(r"(?:total|unique) kmer count: (\d+)", out)]
        cur = dict(fa=md5(os.path.join(wd, "gc_fix1.fa")), link=md5(os.path.join(wd, "ont_link.txt")),
                   valid=md5(os.path.join(wd, "valid_ont_link.txt")), stats=stat,
                   n_left=int(open(os.path.join(wd, "gc_fix1.fa")).read().count("N")))
        if res:
            assert cur == res, "reference output depends on n_thread?"
        res = cur
    print("gc", cfg, res)
    return res


def sw_fixture():
    rng = np.random.default_rng(2024)
    cases = []
    refs = {m: orc.RefSW(m) for m in ("asis", "fixed")}
    # (a) independent aligners: fresh sw_init + one sw_set_parameter per case
    for n in range(160):
        tl = int(rng.integers(1, 90)); ql = int(rng.integers(1, 140))
        t = rng.integers(0, 4, tl).astype(np.uint8)
        if rng.random() < 0.65 and tl > 4:
            a = int(rng.integers(0, tl // 2))
            q = np.concatenate([rng.integers(0, 4, int(rng.integers(0, 30))).astype(np.uint8), synth.mutate(t[a:], 0.15, rng) % 4,
                                rng.integers(0, 4, int(rng.integers(0, 30))).astype(np.uint8)]).astype(np.uint8)
        else:
            q = rng.integers(0, 4, ql).astype(np.uint8)
        if len(q) == 0:
            q = np.array([2], np.uint8)
        strategy = int(rng.integers(0, 4))
        if n < 60:
            mat, pen = orc.default_mat(), (2, 1, 2, 1)
        else:
            mat = orc.default_mat(5, int(rng.integers(1, 5)), -int(rng.integers(1, 8)))
            pen = (int(rng.integers(1, 8)), int(rng.integers(1, 3)), int(rng.integers(1, 8)), int(rng.integers(1, 3)))
        case = dict(q=q.tolist(), t=t.tolist(), strategy=strategy, mat=mat.reshape(-1).tolist(), type_c=int(mat.shape[0]), pen=list(pen))
        for m in ("asis", "fixed"):
            R = orc.RefSW(m)
            R.set(mat, pen[0], pen[1], pen[2], pen[3], strategy)
            r = R.align(q, t)
            R.close()
            case[m] = dict(score=r["score"], offset=r["offset"], softclip=r["softclip"], cigar=orc.cigar_str(r["cigar"]))
        cases.append(case)
    # (b) call sequences on ONE aligner: parameters change between calls, the matrix grows
    seqs = []
    for s in range(6):
        steps = []
        Rs = {m: orc.RefSW(m) for m in ("asis", "fixed")}
        for st in range(7):
            strategy = int(rng.integers(0, 4))
            pen = (int(rng.integers(1, 6)), int(rng.integers(1, 3)), int(rng.integers(1, 6)), int(rng.integers(1, 3)))
            big = rng.random() < 0.3
            tl = int(rng.integers(400, 700)) if big else int(rng.integers(5, 80))
            ql = int(rng.integers(100, 300)) if big else int(rng.integers(5, 100))
            t = rng.integers(0, 4, tl).astype(np.uint8)
            q = np.resize(synth.mutate(t, 0.1, rng) % 4, ql).astype(np.uint8)
            step = dict(set=bool(st == 0 or rng.random() < 0.6), strategy=strategy, pen=list(pen), q=q.tolist(), t=t.tolist())
            for m in ("asis", "fixed"):
                if step["set"]:
                    Rs[m].set(orc.default_mat(), pen[0], pen[1], pen[2], pen[3], strategy)
                r = Rs[m].align(q, t)
                step[m] = dict(score=r["score"], offset=r["offset"], softclip=r["softclip"], cigar=orc.cigar_str(r["cigar"]))
            steps.append(step)
        for R in Rs.values():
            R.close()
        seqs.append(steps)
    # (c) Blizzard hash known answers (hash_func.c:52-67)
    bl = {w: [refs["asis"].blizzard(w.encode(), ht) for ht in (0, 1, 2)] for w in ("chr1", "scaffold_12", "ACGT", "", "Chr1")}
    json.dump(dict(cases=cases, sequences=seqs, blizzard=bl), open(os.path.join(OUT, "sw_vectors.json"), "w"))
    print("sw", len(cases), "cases,", len(seqs), "call sequences")


def main():
    assert orc.have_ref(), "build oracle/_ref first (bash oracle/build_ref.sh)"
    if len(sys.argv) > 2 and sys.argv[1] == "--gc":        # add / refresh the CLI fixtures of the named configs only
        path = os.path.join(OUT, "gc_e2e.json")
        gc = json.load(open(path))
        with tempfile.TemporaryDirectory() as tmp:
            for cfg in sys.argv[2:]:
                gc[cfg] = gc_fixture(tmp, cfg, nts=(3, 8))
        json.dump(gc, open(path, "w"), indent=1)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--aux":
        with tempfile.TemporaryDirectory() as tmp:
            aux_fixture(tmp, "tiny", 25)
        return
    with tempfile.TemporaryDirectory() as tmp:
        for cfg, k in (("tiny", 25), ("repeats", 25), ("repeats", 17), ("tiny", 31)):
            kmer_fixture(tmp, cfg, k)
        aux_fixture(tmp, "tiny", 25)
        gc = {cfg: gc_fixture(tmp, cfg) for cfg in ("tiny", "small", "repeats", "cfg1")}
        json.dump(gc, open(os.path.join(OUT, "gc_e2e.json"), "w"), indent=1)
    sw_fixture()


if __name__ == "__main__":
    main()
