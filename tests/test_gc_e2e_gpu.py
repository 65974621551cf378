"""GPU end-to-end: gc_b200 (the reference's unchanged callers linked against the replacement
kmer.c/hash.c/hash_func.c/sw.c/cigar.c/ont.c and libgcgpu.so) must write byte-identical
gc_fix1.fa / ont_link.txt / valid_ont_link.txt and print the same four k-mer statistics as the
reference gc (fixtures: tests/golden/gc_e2e.json, produced by the reference itself)."""
import hashlib
import json
import os
import re
import subprocess
import tempfile

import pytest

from superplus_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GC = os.path.join(ROOT, "superplus_b200", "_build", "gc_b200")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "gc_e2e.json")))


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


FULL = [("cfg5", 16)] if os.environ.get("GC_E2E_FULL") else []      # BASELINE configs[4] at full size: 100 GB of host memory, minutes


@pytest.mark.parametrize("cfg,n_thread", [("tiny", 1), ("small", 4), ("repeats", 3), ("cfg1", 8), ("cfg5s", 8), ("cfg2", 8)] + FULL)
def test_gap_filled_fasta_is_bit_exact(cfg, n_thread):
    # cfg5 (only with GC_E2E_FULL=1): BASELINE configs[4] at full size — 250 Mb, 20 000 gaps, 20x; the fixture is the
    #   reference's own run on the GPU box (profiles/r02_cli_cfg5_full.json: reference 206 s, gc_b200 43 s, files identical)
    # cfg2: BASELINE configs[1], the bench's headline config, whole CLI
    # cfg5s: cfg5's gap density at 20 Mb / 6x — the contig table is beyond the L2 (pre-filter path) and
    # the reads go through the host pipeline in 15 chunks
    assert os.path.exists(GC), "gc_b200 was not built (python -c 'import __graft_entry__ as g; g.build()')"
    with tempfile.TemporaryDirectory() as tmp:
        fa, fq, _ = synth.materialise(cfg, tmp)
        wd = os.path.join(tmp, "run")
        os.makedirs(wd)
        r = subprocess.run([GC, fa, fq, str(n_thread), "out"], cwd=wd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=1800)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        out = r.stdout.decode()
        stats = [int(x) for x in re.findall(r"(?:total|unique) kmer count: (\d+)", out)]
        g = GOLD[cfg]
        assert stats == g["stats"]
        assert md5(os.path.join(wd, "gc_fix1.fa")) == g["fa"]
        assert md5(os.path.join(wd, "ont_link.txt")) == g["link"]
        assert md5(os.path.join(wd, "valid_ont_link.txt")) == g["valid"]
        # the phase lines of the reference are still printed
        for line in ("chop kmers cost", "hash kmers cost", "chop and search ont kmers cost", "re-hash ont kmers cost",
                     "find un-ankored positions on onts costs", "search ONT kmers total cost", "Program Cost"):
            assert line in out, line


def _device_count():
    from superplus_b200 import api
    return int(api.load_library().gcg_device_count())


@pytest.mark.parametrize("cfg,devices", [("small", "0,0"), ("repeats", "0,0,0"), ("cfg1", "0,0"), ("cfg1", "all"), ("cfg5s", "0,1")])
def test_reads_sharded_over_devices_same_outputs(cfg, devices):
    """GC_DEVICES shards the ONT read batch over several GPUs (one table replica and one host thread per
    device): files and the four statistics must not change.  "0,0" runs two contexts on one
    GPU (the single-GPU box still covers clone / shares / merge); "0,1" needs a second GPU."""
    assert os.path.exists(GC), "gc_b200 was not built (python -c 'import __graft_entry__ as g; g.build()')"
    if devices == "0,1" and _device_count() < 2:
        pytest.skip("needs two GPUs")
    with tempfile.TemporaryDirectory() as tmp:
        fa, fq, _ = synth.materialise(cfg, tmp)
        wd = os.path.join(tmp, "run")
        os.makedirs(wd)
        env = dict(os.environ, GC_DEVICES=devices, GCG_TRACE="1")
        r = subprocess.run([GC, fa, fq, "8", "out"], cwd=wd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, env=env)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        stats = [int(x) for x in re.findall(r"(?:total|unique) kmer count: (\d+)", r.stdout.decode())]
        g = GOLD[cfg]
        assert stats == g["stats"]
        assert md5(os.path.join(wd, "gc_fix1.fa")) == g["fa"]
        assert md5(os.path.join(wd, "ont_link.txt")) == g["link"]
        assert md5(os.path.join(wd, "valid_ont_link.txt")) == g["valid"]
        if devices != "all" or _device_count() > 1:
            assert "reads [" in r.stderr.decode()          # the shares were really spread


def test_bad_device_list_fails_loudly():
    with tempfile.TemporaryDirectory() as tmp:
        fa, fq, _ = synth.materialise("tiny", tmp)
        r = subprocess.run([GC, fa, fq, "1", "out"], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120,
                           env=dict(os.environ, GC_DEVICES="0;1"))
        assert r.returncode != 0 and b"GC_DEVICES" in r.stderr
        r = subprocess.run([GC, fa, fq, "1", "out"], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120,
                           env=dict(os.environ, GC_DEVICES="0,99"))
        assert r.returncode != 0 and b"cannot open CUDA device 99" in r.stderr


def test_replicas_and_merged_statistics():
    """gcg_table_clone / gcg_table_merge_ont: reads split over 1..4 contexts (all devices of the box in
    turn) give the anchors and the statistics of one search; a second batch keeps adding up; a table
    that is not a clone is refused"""
    import numpy as np
    from superplus_b200 import api
    inp = synth.make_config("small")
    reads = inp.reads[:40] + [np.zeros(0, np.uint8), inp.reads[3][:24]] + inp.reads[40:]
    ndev = _device_count()
    ctx0 = api.Context(0)
    cs = ctx0.upload(inp.contigs)
    t = ctx0.table_build(cs, 25)
    want = ctx0.search_host(t, reads)
    st1 = t.stats()
    ctx0.search_host(t, reads[:30])
    st2 = t.stats()
    t.free()
    assert st2[2] == st1[2] and st2[3] < st1[3]                       # same k-mers anchored, fewer of them exactly once
    for n in (1, 2, 3, 4):
        ctxs = [ctx0] + [api.Context(i % ndev) for i in range(1, n)]
        t = ctx0.table_build(cs, 25)
        rep = api.ReplicatedSearch(ctxs, t)
        got = rep.search_host(reads)
        assert np.array_equal(got, want), n
        assert t.stats() == st1, n
        again = rep.search_host(reads[:30])                            # counts move to the primary and keep adding up
        assert np.array_equal(again, want[want["read"] < 30]), n
        assert t.stats() == st2, n
        if n > 1:
            late = t.clone(ctxs[1])                                    # a clone starts with nothing collected
            assert late.stats()[:2] == st2[:2] and late.stats()[2:] == (0, 0)
            late.free()
        rep.free(); t.free()
        for c in ctxs[1:]:
            c.close()
    t = ctx0.table_build(cs, 25)
    other = ctx0.table_build(ctx0.upload(inp.contigs[:1]), 25)
    with pytest.raises(api.GcgError):
        t.merge_ont(other)
    with pytest.raises(api.GcgError):
        t.merge_ont(t)
    t.free(); other.free(); cs.free(); ctx0.close()


def test_search_in_groups_matches_single_call():
    """gcg_search streams the reads through a three-slot pipeline in chunks of whole reads: the
    device-resident search, the default chunking and tiny chunks (many chunks, slot reuse, result
    buffer growth, a read larger than the chunk) must give the same anchors and statistics"""
    import numpy as np
    from superplus_b200 import api
    ctx = api.Context(0)
    inp = synth.make_config("small")
    reads = inp.reads[:60] + [inp.reads[0][:10], np.zeros(0, np.uint8), inp.reads[1][:24]] + inp.reads[60:]
    cs, rs = ctx.upload(inp.contigs), ctx.upload(reads)
    t = ctx.table_build(cs, 25)
    want = ctx.search(t, rs)
    st_want = t.stats()
    t.free()
    for chunk in (None, 1 << 20, 65536, 4096):
        t = ctx.table_build(cs, 25)
        if chunk is not None:
            os.environ["GCG_SEARCH_CHUNK_BYTES"] = str(chunk)
        try:
            got = ctx.search_host(t, reads)
            again = ctx.search_host(t, reads[:7])          # a second call on the same pipeline
        finally:
            os.environ.pop("GCG_SEARCH_CHUNK_BYTES", None)
        assert np.array_equal(got, want), chunk
        assert np.array_equal(again, want[want["read"] < 7]), chunk
        t.free()
        t = ctx.table_build(cs, 25)
        ctx.search_host(t, reads)
        assert t.stats() == st_want
        t.free()
    cs.free(); rs.free(); ctx.close()


@pytest.mark.parametrize("cfg,k,n_thread,env", [("tiny", 25, 3, {}), ("repeats", 17, 7, {}), ("small", 31, 2, {"GC_ANCHORS16": "1"}),
                                                ("cfg1", 25, 5, {"GC_DEVICES": "0,0"})])
def test_shim_harness_dumps_equal_the_reference_harness(cfg, k, n_thread, env):
    """oracle/ref_kmer_harness.c performs main.c's call sequence (main.c:147-187) and dumps the host
    structures the unchanged consumers read: kmers[] (with hs_id = crc32 % n_thread, kmer.c:88), every
    non-NULL okmers[] entry, and okseq->segs (find_unankor_segs, ont.c:264-309).  Built once against the
    reference's own objects (oracle/_ref/ref_kmer) and once against the replacement files + libgcgpu.so
    (superplus_b200/_build/shim_kmer): the dumps must be byte-identical, the printed statistics equal.
    GC_ANCHORS16 selects the 16-byte anchor download instead of the compact one."""
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_kmer")
    shim = os.path.join(ROOT, "superplus_b200", "_build", "shim_kmer")
    assert os.path.exists(ref) and os.path.exists(shim), "harness binaries were not built (python -c 'import __graft_entry__ as g; g.build()')"
    with tempfile.TemporaryDirectory() as tmp:
        fa, fq, _ = synth.materialise(cfg, tmp)
        outs = {}
        for name, exe, e in (("ref", ref, {}), ("shim", shim, env)):
            r = subprocess.run([exe, fa, fq, str(n_thread), str(k), os.path.join(tmp, name), "2"], cwd=tmp, stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, timeout=600, env=dict(os.environ, **e))
            assert r.returncode == 0, (name, r.stderr.decode()[-2000:])
            outs[name] = [int(x) for x in re.findall(r"(?:total|unique) kmer count: (\d+)", r.stdout.decode())]
        assert outs["ref"] == outs["shim"] and len(outs["ref"]) == 4
        for ext in ("hits.bin", "ctgk.bin", "hsid.bin", "segs.bin"):
            a, b = open(os.path.join(tmp, "ref." + ext), "rb").read(), open(os.path.join(tmp, "shim." + ext), "rb").read()
            assert len(a) > 0 and a == b, ext


@pytest.mark.parametrize("cfg,env", [("tiny", {}), ("small", {}), ("repeats", {}), ("cfg1", {}), ("cfg5s", {}), ("cfg2", {}),
                                     ("cfg1", {"GC_DEVICES": "0,0"}), ("repeats", {"GC_DEVICES": "0,0,0"})])
def test_runs_mode_writes_the_same_files(cfg, env):
    """GC_RUNS=1 (opt-in, SURVEY 8f rows N2 + N3): anchors are reduced to run records on the device, the dense
    okmers[] (16 bytes per ONT base) and ctg->kmers[] (24 per contig base) are never written, map_ont2contigs builds its
    nodes from the records and hands them to the reference's own add_ont_link2graph — files and statistics must be
    the reference's, byte for byte"""
    assert os.path.exists(GC), "gc_b200 was not built (python -c 'import __graft_entry__ as g; g.build()')"
    with tempfile.TemporaryDirectory() as tmp:
        fa, fq, _ = synth.materialise(cfg, tmp)
        wd = os.path.join(tmp, "run")
        os.makedirs(wd)
        r = subprocess.run([GC, fa, fq, "8", "out"], cwd=wd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900,
                           env=dict(os.environ, GC_RUNS="1", **env))
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        stats = [int(x) for x in re.findall(r"(?:total|unique) kmer count: (\d+)", r.stdout.decode())]
        g = GOLD[cfg]
        assert stats == g["stats"]
        assert md5(os.path.join(wd, "gc_fix1.fa")) == g["fa"]
        assert md5(os.path.join(wd, "ont_link.txt")) == g["link"]
        assert md5(os.path.join(wd, "valid_ont_link.txt")) == g["valid"]


@pytest.mark.parametrize("cfg,env", [("tiny", {}), ("small", {"GC_ANCHORS16": "1"}), ("repeats", {}), ("cfg1", {}), ("cfg5s", {}),
                                     ("cfg1", {"GC_DEVICES": "0,0"}), ("cfg2", {})])
def test_sparse_kmers_mode_writes_the_same_files(cfg, env):
    """GC_SPARSE_KMERS=1 (opt-in, SURVEY 8f row N2 with every file of the reference's consumer side untouched, default
    map_ont2contigs included): the chop does not download 24 bytes per contig base; the search's back-fill writes the
    kmer_t record of every contig position an anchor points at (same fields as chop_kmer_core), the rest of
    ctg->kmers[] stays the loader's untouched zero pages.  Files and statistics must be the reference's."""
    assert os.path.exists(GC), "gc_b200 was not built (python -c 'import __graft_entry__ as g; g.build()')"
    with tempfile.TemporaryDirectory() as tmp:
        fa, fq, _ = synth.materialise(cfg, tmp)
        wd = os.path.join(tmp, "run")
        os.makedirs(wd)
        r = subprocess.run([GC, fa, fq, "8", "out"], cwd=wd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900,
                           env=dict(os.environ, GC_SPARSE_KMERS="1", **env))
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        stats = [int(x) for x in re.findall(r"(?:total|unique) kmer count: (\d+)", r.stdout.decode())]
        g = GOLD[cfg]
        assert stats == g["stats"]
        assert md5(os.path.join(wd, "gc_fix1.fa")) == g["fa"]
        assert md5(os.path.join(wd, "ont_link.txt")) == g["link"]
        assert md5(os.path.join(wd, "valid_ont_link.txt")) == g["valid"]


# ----------------------------------------------------------------------------------------------------------------------
# GC_SW_FILL=1 (opt-in): gap junctions from batched alignments instead of the extrapolation
def _code(a):
    return ((a >> 1) & 3).astype("uint8")                                   # bio.h:22-24


def _fasta_bytes(scaffolds):
    import numpy as np
    out = []
    for i, seq in enumerate(scaffolds):
        out.append(b">%d\n" % i)
        n = len(seq)
        body = np.full(n + (n + 59) // 60, ord("\n"), np.uint8)
        idx = np.arange(n)
        body[idx + idx // 60] = seq
        out.append(body[: n + n // 60].tobytes() if n % 60 == 0 else body.tobytes())
    return b"".join(out)


def _read_sw_fill(path):
    rows = []
    for line in open(path):
        if line.startswith("#"):
            continue
        f = line.split()
        rows.append((f[0], [int(x) for x in f[1:]]))
    return rows


G_COLS = "gap from_ctg to_ctg ont fwd k pL cL pR cR read_len flank_a head_c jL_ref jR_ref jL jR wL0 wL1 wR0 wR1 score_a score_b".split()


@pytest.mark.parametrize("cfg,env", [("tiny", {}), ("small", {}), ("repeats", {"GC_RUNS": "1"}), ("cfg1", {}), ("cfg1", {"GC_RUNS": "1"})])
def test_sw_fill_junctions_equal_the_cpu_oracle(cfg, env):
    """Every gap's two alignments are recomputed on the CPU with the oracle's restatement of sw.c (pinned against the
    reference's own sw.c by tests/test_oracle_vs_ref.py) from the FASTA, the FASTQ and the anchors of sw_fill.tsv: the end
    cells, scores and therefore junctions must be the batch's, and gc_fix1.fa must be the scaffolds rebuilt from them.
    The same rebuild with the extrapolated junctions must give the REFERENCE's file (golden md5) — which pins the
    test's own reading of fix_ont1 / combine_ont_info (ctg_graph.c:455-531, 669-755).  The link files do not change."""
    import numpy as np
    from oracle import oracle as orc_mod
    orc = orc_mod.Oracle()
    P = orc_mod.make_params(strategy=2)                                       # SWOS_INDEL, +1/-5, open 2 extend 1 (gc_graph.c:74-107)
    with tempfile.TemporaryDirectory() as tmp:
        fa, fq, inp = synth.materialise(cfg, tmp)
        if inp is None:
            inp = synth.make_config(cfg)
        wd = os.path.join(tmp, "run")
        os.makedirs(wd)
        r = subprocess.run([GC, fa, fq, "8", "out"], cwd=wd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900,
                           env=dict(os.environ, GC_SW_FILL="1", **env))
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        g = GOLD[cfg]
        assert md5(os.path.join(wd, "ont_link.txt")) == g["link"]
        assert md5(os.path.join(wd, "valid_ont_link.txt")) == g["valid"]
        contigs = inp.contigs
        rows = _read_sw_fill(os.path.join(wd, "sw_fill.tsv"))
        ref_scafs, sw_scafs, fills = [], [], []
        n_gap = n_moved = 0
        for kind, f in rows:
            if kind == "S":
                ref_scafs.append([contigs[f[1]]]); sw_scafs.append([contigs[f[1]]])
            elif kind == "N":
                for s in (ref_scafs, sw_scafs):
                    s[-1] += [np.full(max(f[2], 0), ord("N"), np.uint8), contigs[f[1]]]
            else:
                d = dict(zip(G_COLS, f))
                n_gap += 1
                read = inp.reads[d["ont"]]
                L, k = len(read), d["k"]
                assert L == d["read_len"]
                rs = read if d["fwd"] else synth.revcomp(read)
                cf, ct = contigs[d["from_ctg"]], contigs[d["to_ctg"]]
                a, c = len(cf) - d["cL"], d["cR"] + k
                assert (a, c) == (d["flank_a"], d["head_c"])
                pLs = d["pL"] if d["fwd"] else L - d["pL"] - k
                pRs = d["pR"] if d["fwd"] else L - d["pR"] - k
                jL_ref, jR_ref = pLs + a, pRs - d["cR"]
                assert (jL_ref, jR_ref) == (d["jL_ref"], d["jR_ref"])
                jL, jR = jL_ref, jR_ref
                ok_l = 0 <= pLs and pLs + k <= L and np.array_equal(_code(rs[pLs:pLs + k]), _code(cf[d["cL"]:d["cL"] + k]))
                ok_r = 0 <= pRs and pRs + k <= L and np.array_equal(_code(rs[pRs:pRs + k]), _code(ct[d["cR"]:d["cR"] + k]))
                if 1 <= a <= 4000 and ok_l:
                    w0, w1 = pLs, min(L, pLs + a + a // 4 + 32)
                    assert (w0, w1) == (d["wL0"], d["wL1"])
                    res = orc.sw_align(P, _code(cf[d["cL"]:]), _code(rs[w0:w1]))
                    jL = w0 + res["bt_tidx"]
                    assert res["score"] == d["score_a"], d
                if 1 <= c <= 4000 and ok_r:
                    w1, w0 = pRs + k, max(0, pRs + k - (c + c // 4 + 32))
                    assert (w0, w1) == (d["wR0"], d["wR1"])
                    res = orc.sw_align(P, _code(ct[:c])[::-1].copy(), _code(rs[w0:w1])[::-1].copy())
                    jR = w1 - res["bt_tidx"]
                    assert res["score"] == d["score_b"], d
                assert (jL, jR) == (d["jL"], d["jR"]), d
                n_moved += (jL, jR) != (jL_ref, jR_ref)
                sw_scafs[-1] += [rs[jL:jR] if jR > jL else rs[:0], ct]
                fills.append((d["from_ctg"], d["to_ctg"], d["fwd"], sw_scafs[-1][-2]))
                # the reference's own fill: forward reads start at the extrapolated junction, reverse reads k - 1 bases
                # behind it (its reverse branch walks down from the k-mer's START in the stored read, ctg_graph.c:516-528)
                n_ref = jR_ref - jL_ref
                s0 = jL_ref if d["fwd"] else jL_ref + k - 1
                ref_scafs[-1] += [rs[s0:s0 + n_ref] if n_ref > 0 else rs[:0], ct]
                fills[-1] += (ref_scafs[-1][-2],)
        assert n_gap > 0
        ref_fa = _fasta_bytes([np.concatenate(s) for s in ref_scafs])
        assert hashlib.md5(ref_fa).hexdigest() == g["fa"]                   # the test reads fix_ont1 as the reference wrote it
        sw_fa = _fasta_bytes([np.concatenate(s) for s in sw_scafs])
        assert open(os.path.join(wd, "gc_fix1.fa"), "rb").read() == sw_fa
        assert n_moved > 0                                                   # 10 % read errors: the alignment must move some junction
        # what the junctions are for: the filled gap with 150 contig bases either side, aligned against the true genome
        # (the synthetic truth) — the aligned junctions must reproduce it at least as well as the extrapolated ones
        F, tot = 150, {"sw": 0, "ref": 0, "sw_rev": 0, "ref_rev": 0, "gaps": 0, "gaps_rev": 0}
        for (fc, tc, fwd, f_sw, f_ref) in fills:
            if tc != fc + 1 or fc >= len(inp.gaps) or len(contigs[fc]) < F or len(contigs[tc]) < F:
                continue
            gs, gl = inp.gaps[fc]
            truth = _code(inp.genome[gs - F: gs + gl + F])
            for name, fill in (("sw", f_sw), ("ref", f_ref)):
                sc = orc.sw_align(P, _code(np.concatenate([contigs[fc][-F:], fill, contigs[tc][:F]])), truth)["score"]
                tot[name] += sc
                if not fwd:
                    tot[name + "_rev"] += sc
            tot["gaps"] += 1
            tot["gaps_rev"] += not fwd
        tot.update(cfg=cfg, n_filled=n_gap, n_moved=int(n_moved))
        print("sw_fill quality:", json.dumps(tot))
        out_dir = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(out_dir):
            json.dump(tot, open(os.path.join(out_dir, "sw_fill_quality_%s_%s.json" % (cfg, "runs" if env else "dense")), "w"))
        assert tot["gaps"] > 0 and tot["sw"] >= tot["ref"], tot
