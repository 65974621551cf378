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
