"""GPU parity: CUDA k-mer path (through the C ABI) vs the CPU oracle, bit exact."""
import numpy as np
import pytest

from superplus_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["stream", "chunked"])
def host_pipeline(request, monkeypatch):
    """every test runs with both forms of the host-buffer search: one launch over reads that are still being uploaded
    (default) and one launch per chunk (GCG_SEARCH_STREAM=0); GCG_SEARCH_CHUNK_BYTES sizes the pieces / the chunks"""
    monkeypatch.setenv("GCG_SEARCH_STREAM", "1" if request.param == "stream" else "0")
    return request.param


def asc(s):
    return np.frombuffer(s.encode(), dtype=np.uint8).copy()


def check_case(ctx, oracle, contigs, reads, k, via_ptrs=False):
    # ---- oracle
    h = oracle.table_build(contigs, k)
    o_st = oracle.table_stats(h)
    o_key, o_multi, o_tid, o_pos, o_rev = oracle.table_dump(h)
    o_hits, o_ont = oracle.search(h, reads, k)
    oracle.table_free(h)
    # ---- CUDA
    up = ctx.upload_ptrs if via_ptrs else ctx.upload
    cs = up(contigs)
    rs = up(reads)
    assert cs.bases == sum(len(c) for c in contigs) and rs.kmers(k) == sum(max(0, len(r) - k + 1) for r in reads)
    t = ctx.table_build(cs, k)
    g_key, g_multi, g_tid, g_pos, g_rev = t.dump()
    assert np.array_equal(g_key, o_key)
    assert np.array_equal(g_multi, np.minimum(o_multi, 2))
    u = o_multi == 1
    assert np.array_equal(g_tid[u], o_tid[u]) and np.array_equal(g_pos[u], o_pos[u]) and np.array_equal(g_rev[u], o_rev[u])
    hits = ctx.search(t, rs)
    st = t.stats()
    assert st[:2] == o_st, (st, o_st)
    assert st[2:] == o_ont, (st, o_ont)
    assert len(hits) == len(o_hits["read"])
    assert np.array_equal(hits["read"], o_hits["read"].astype(np.int32))
    assert np.array_equal(hits["pos"], o_hits["pos"])
    assert np.array_equal(hits["tid"], o_hits["tid"])
    assert np.array_equal(hits["cpos_flags"] >> 2, o_hits["cpos"].astype(np.uint32))
    assert np.array_equal(hits["cpos_flags"] & 1, o_hits["krev"])
    assert np.array_equal((hits["cpos_flags"] >> 1) & 1, o_hits["orev"])
    # the same search as compact anchors (8 bytes per anchor + per-read offsets) on a fresh table:
    # same anchors, same order, same ONT-side multiplicity
    t2 = ctx.table_build(cs, k)
    anchors, read_off = ctx.search_compact(t2, rs)
    assert len(read_off) == len(reads) + 1 and read_off[0] == 0 and read_off[-1] == len(anchors) and np.all(np.diff(read_off) >= 0)
    assert np.array_equal(api.expand_compact(anchors, read_off, [len(c) for c in contigs]), hits)
    assert t2.stats() == st
    t2.free()
    # kmer_t records for the unchanged host consumers; hs_id = crc32(kseq) % n_thread (kmer.c:88)
    n_thread = 1 + (len(hits) + k) % 7
    recs = ctx.chop_contigs(cs, [len(c) for c in contigs], k, n_thread=n_thread)
    for i, c in enumerate(contigs):
        ks, rv = oracle.chop(c, k)
        r = recs[i]
        assert np.array_equal(r["kseq"], ks) and np.array_equal(r["flag"], rv)
        assert np.all(r["tid"] == i) and np.array_equal(r["pos"], np.arange(len(ks), dtype=np.int32)) and np.all(r["kmer_len"] == k)
        assert np.array_equal(r["hs_id"], oracle.hs_id(ks, n_thread))
    t.free(); cs.free(); rs.free()
    return len(hits)


@pytest.mark.parametrize("name,k", [("tiny", 25), ("repeats", 25), ("repeats", 17), ("small", 31), ("tiny", 5)])
def test_synthetic_parity(ctx, oracle, name, k):
    inp = synth.make_config(name)
    n = check_case(ctx, oracle, inp.contigs, inp.reads, k)
    assert n > 0 or k < 12      # short k: every k-mer repeats, nothing anchors


def test_pointer_entry_point(ctx, oracle):
    inp = synth.make_config("tiny")
    check_case(ctx, oracle, inp.contigs, inp.reads, 25, via_ptrs=True)


def test_edge_cases(ctx, oracle):
    rng = np.random.default_rng(3)
    g = synth.random_genome(5000, rng)
    contigs = [g[:40], g[40:40], g[100:124], g[200:225], g[300:1500], asc("acgtnACGTN" * 9), g[2000:5000], g[300:900]]
    reads = [g[0:0], g[5:20], g[200:225], g[190:260], asc("NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN"), g[310:1400],
             synth.revcomp(g[2100:4000]), np.char.lower(g[2500:2600].view("S1")).view(np.uint8).copy(),
             asc("ACGTNRYKM" * 20), g[299:325]]
    for k in (25, 24, 31, 1, 13):
        check_case(ctx, oracle, contigs, reads, k)


def test_empty_inputs(ctx, oracle):
    rng = np.random.default_rng(4)
    g = synth.random_genome(300, rng)
    check_case(ctx, oracle, [g], [], 25)
    check_case(ctx, oracle, [], [g], 25)
    check_case(ctx, oracle, [g[:10]], [g], 25)


def test_search_host_form(ctx, oracle):
    inp = synth.make_config("tiny")
    cs = ctx.upload(inp.contigs)
    t = ctx.table_build(cs, 25)
    a = ctx.search_host(t, inp.reads)
    rs = ctx.upload(inp.reads)
    b = ctx.search(t, rs)
    assert np.array_equal(a, b)
    t.free(); cs.free(); rs.free()


def test_error_paths(ctx):
    inp = synth.make_config("tiny")
    cs = ctx.upload(inp.contigs)
    with pytest.raises(api.GcgError):
        ctx.table_build(cs, 32)
    with pytest.raises(api.GcgError):
        ctx.table_build(cs, 0)
    t = ctx.table_build(cs, 25)
    t.k = 21
    rs = ctx.upload(inp.reads[:2])
    with pytest.raises(api.GcgError):
        ctx.search(t, rs)
    t.k = 25
    t.free(); cs.free(); rs.free()


@pytest.mark.parametrize("max_mb", ["64", "1"])
def test_prefilter_forced_on(ctx, oracle, max_mb, monkeypatch):
    """tables beyond the L2 are probed through a Bloom-style pre-filter of the anchoring keys;
    forced on here (also with a tiny, collision-heavy filter) the results must not change"""
    monkeypatch.setenv("GCG_FILTER", "1")
    monkeypatch.setenv("GCG_FILTER_MAX_MB", max_mb)
    for name, k in (("repeats", 17), ("small", 31), ("tiny", 5)):
        inp = synth.make_config(name)
        check_case(ctx, oracle, inp.contigs, inp.reads, k)
    inp = synth.make_config("small")
    cs = ctx.upload(inp.contigs)
    t = ctx.table_build(cs, 25)
    a = ctx.search_host(t, inp.reads)
    monkeypatch.setenv("GCG_FILTER", "0")
    t2 = ctx.table_build(cs, 25)
    b = ctx.search_host(t2, inp.reads)
    assert np.array_equal(a, b) and t.stats() == t2.stats()
    t.free(); t2.free(); cs.free()


def test_randomised_small_cases(ctx, oracle, monkeypatch):
    """many tiny inputs with every k, lengths around the 32-base word and 1024-base tile boundaries,
    empty and shorter-than-k sequences, repeats, N and lowercase: device search, host pipeline with
    tiny chunks, and the partition routing must all agree with the oracle"""
    from part_double import HostSeqs, NumpyOps
    rng = np.random.default_rng(2024)
    alphabet = np.frombuffer(b"ACGTACGTACGTacgtN", dtype=np.uint8)
    lens_pool = [0, 1, 2, 24, 25, 30, 31, 32, 33, 34, 63, 64, 65, 95, 96, 97, 127, 128, 129, 1023, 1024, 1025, 1056, 2047, 2049]
    monkeypatch.setenv("GCG_SEARCH_CHUNK_BYTES", "2048")
    dbl = NumpyOps(oracle)
    for case in range(60):
        k = int(rng.integers(1, 32))
        base = alphabet[rng.integers(0, len(alphabet), size=int(rng.integers(200, 3000)))]
        contigs = []
        for _ in range(int(rng.integers(1, 12))):
            l = int(rng.choice(lens_pool)) if rng.random() < 0.6 else int(rng.integers(0, 400))
            a = int(rng.integers(0, max(1, len(base) - l + 1)))
            contigs.append(base[a:a + l].copy())
        reads = []
        for _ in range(int(rng.integers(0, 40))):
            l = int(rng.choice(lens_pool)) if rng.random() < 0.5 else int(rng.integers(0, 300))
            a = int(rng.integers(0, max(1, len(base) - l + 1)))
            r = base[a:a + l].copy()
            if len(r) and rng.random() < 0.5:
                r[rng.integers(0, len(r), size=max(1, len(r) // 12))] = ord("A")
            if rng.random() < 0.3:
                r = synth.revcomp(np.char.upper(r.view("S1")).view(np.uint8)) if len(r) else r
            reads.append(r)
        n_hit = check_case(ctx, oracle, contigs, reads, k, via_ptrs=bool(case & 1))
        # host pipeline (chunks of 2 KiB: slot reuse, reads larger than a chunk)
        cs = ctx.upload(contigs)
        t = ctx.table_build(cs, k)
        host = ctx.search_host(t, reads)
        rs0 = ctx.upload(reads)
        assert len(host) == n_hit and np.array_equal(host, ctx.search(t, rs0))
        rs0.free(); t.free()
        # partition routing of the reads
        rs = ctx.upload(reads)
        n_part = int(rng.integers(1, 17))
        want = dbl.plan(HostSeqs(reads), k, n_part, 0, rs.tiles)
        got = ctx.route_plan(rs, k, n_part, 0, rs.tiles)
        assert np.array_equal(got.counts, want.counts), (case, k, n_part)
        got.free(); rs.free(); cs.free()


def test_search_host_compact_form(ctx, oracle, monkeypatch):
    """gcg_search_compact (what the shim calls): the anchors of gcg_search, 8 bytes each, grouped by
    read; default chunks and tiny chunks (read offsets stitched over many chunks, reads without
    k-mers, empty reads at either end)"""
    inp = synth.make_config("small")
    e = np.zeros(0, np.uint8)
    reads = [e, inp.reads[0][:10]] + inp.reads[:50] + [e, e, inp.reads[1][:24]] + inp.reads[50:] + [e]
    cs = ctx.upload(inp.contigs)
    t = ctx.table_build(cs, 25)
    want = ctx.search_host(t, reads)
    st = t.stats()
    t.free()
    clens = [len(c) for c in inp.contigs]
    for chunk in (None, 65536, 4096):
        if chunk is not None:
            monkeypatch.setenv("GCG_SEARCH_CHUNK_BYTES", str(chunk))
        t = ctx.table_build(cs, 25)
        anchors, read_off = ctx.search_host_compact(t, reads)
        assert len(read_off) == len(reads) + 1 and read_off[-1] == len(anchors)
        assert np.array_equal(api.expand_compact(anchors, read_off, clens), want), chunk
        assert t.stats() == st
        t.free()
    # nothing to search: offsets are still there
    t = ctx.table_build(cs, 25)
    anchors, read_off = ctx.search_host_compact(t, [e, inp.reads[0][:5]])
    assert len(anchors) == 0 and list(read_off) == [0, 0, 0]
    t.free(); cs.free()


@pytest.mark.parametrize("frac", ["0.0", "0.01", "0.05"])
def test_result_buffer_smaller_than_the_anchors(ctx, oracle, frac, monkeypatch):
    """the device-resident search sizes its anchor buffer from an estimate; GCG_SEARCH_CAP_FRAC forces
    estimates far below the truth: the second launch must finish the job without recording any
    anchor's ONT-side multiplicity twice (statistics unchanged)"""
    monkeypatch.setenv("GCG_SEARCH_CAP_FRAC", frac)
    for name, k in (("repeats", 17), ("small", 25)):
        inp = synth.make_config(name)
        check_case(ctx, oracle, inp.contigs, inp.reads, k)


def test_host_result_buffer_grows_while_the_pool_gathers(ctx, oracle, monkeypatch):
    """the pinned result of gcg_search starts from an estimate and is copied into a larger
    block when a chunk does not fit; with more than 4 MiB already placed that copy used the worker pool
    while the pool was still gathering the next chunk, which dropped gather tasks.  Force the path: a
    result buffer of one anchor, several host threads, chunks small enough that many follow."""
    from superplus_b200 import api as _api
    inp = synth.make_config("cfg1")
    c2 = _api.Context(0, host_threads=8)
    cs = c2.upload(inp.contigs)
    t = c2.table_build(cs, 25)
    want = c2.search_host(t, inp.reads)
    assert len(want) * 16 > (8 << 20)                    # well beyond the 4 MiB threshold of the threaded copy
    t.free()
    monkeypatch.setenv("GCG_SEARCH_RES_CAP", "1")
    monkeypatch.setenv("GCG_SEARCH_CHUNK_BYTES", str(1 << 20))
    for compact in (False, True):
        t = c2.table_build(cs, 25)
        if compact:
            anchors, read_off = c2.search_host_compact(t, inp.reads)
            got = _api.expand_compact(anchors, read_off, [len(c) for c in inp.contigs])
        else:
            got = c2.search_host(t, inp.reads)
        assert np.array_equal(got, want), compact
        t.free()
    cs.free(); c2.close()


def test_two_pass_search_still_agrees(ctx, oracle):
    """GCG_SEARCH_FUSED=0 keeps the round-1 form (probe -> masks, prefix sum, emit) for A/B timing; the
    switch is read once per process, so the comparison runs in a child process"""
    import subprocess, sys, os as _os
    code = ("import numpy as np, sys; sys.path.insert(0, %r)\n"
            "from superplus_b200 import api, synth\n"
            "inp = synth.make_config('small'); c = api.Context(0)\n"
            "cs, rs = c.upload(inp.contigs), c.upload(inp.reads); t = c.table_build(cs, 25)\n"
            "h = c.search(t, rs); print(len(h), int(h['pos'].astype(np.int64).sum()), int(h['cpos_flags'].astype(np.int64).sum()), t.stats())\n") % _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    outs = []
    for v in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(_os.environ, GCG_SEARCH_FUSED=v), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        outs.append(r.stdout.decode().strip())
    assert outs[0] == outs[1] and outs[0]


def test_synchronous_launches_keep_the_chunked_pipeline(host_pipeline):
    """the one-launch search needs the host to keep uploading AFTER the launch call; under ncu, compute-sanitizer or
    CUDA_LAUNCH_BLOCKING=1 a launch only returns when the kernel has finished, so the first search of a context probes
    for that and stays with one launch per chunk — same anchors, and it says so"""
    if host_pipeline != "stream":
        pytest.skip("only the default form probes")
    import subprocess, sys, os as _os
    code = ("import numpy as np, sys; sys.path.insert(0, %r)\n"
            "from superplus_b200 import api, synth\n"
            "inp = synth.make_config('small'); c = api.Context(0)\n"
            "cs = c.upload(inp.contigs); t = c.table_build(cs, 25)\n"
            "a, off = c.search_host_compact(t, inp.reads); print(len(a), int((a >> np.uint64(36)).sum()), int(off.sum()), t.stats())\n") % _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    outs = []
    for blocking in ("0", "1"):
        env = dict(_os.environ, GCG_SEARCH_STREAM="1", GCG_TRACE="1")
        env.pop("CUDA_LAUNCH_BLOCKING", None)
        if blocking == "1":
            env["CUDA_LAUNCH_BLOCKING"] = "1"
        r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        outs.append(r.stdout.decode().strip())
        err = r.stderr.decode()
        assert ("launches are SYNCHRONOUS" in err) == (blocking == "1"), err[-1500:]
        assert ("(streaming, one launch)" in err) == (blocking == "0"), err[-1500:]
    assert outs[0] == outs[1] and outs[0]


def _check_runs(ctx, oracle, contigs, reads, k):
    h = oracle.table_build(contigs, k)
    o_hits, o_ont = oracle.search(h, reads, k)
    o_st = oracle.table_stats(h)
    oracle.table_free(h)
    from oracle import oracle as orc
    want, want_off = orc.runs_from_hits(o_hits, len(reads), [len(c) for c in contigs])
    cs = ctx.upload(contigs)
    t = ctx.table_build(cs, k)
    runs, run_off, n_anchor = ctx.search_runs(t, reads)
    assert n_anchor == len(o_hits["read"]) and t.stats() == o_st + o_ont
    assert np.array_equal(run_off, want_off)
    for f in ("tid", "n_fwd", "n_bwd", "first_fwd", "last_fwd", "first_bwd", "last_bwd"):
        assert np.array_equal(runs[f], want[f]), f
    t.free(); cs.free()
    return len(runs)


def test_run_records_match_the_grouping_of_map_ont2contigs(ctx, oracle, monkeypatch):
    """N3 (gcg_search_runs): per read, maximal stretches of consecutive anchors on one contig with the counts of
    the two strand directions and the first / last anchor of each — against the oracle's restatement of
    ctg_graph.c:600-656 on the oracle's anchors.  Repeat-rich input (runs of one anchor, direction flips),
    reads without anchors, empty reads, many tiny chunks, an anchor estimate that forces the second pass."""
    e = np.zeros(0, np.uint8)
    for name, k in (("tiny", 25), ("repeats", 17), ("repeats", 25), ("small", 31)):
        inp = synth.make_config(name)
        reads = [e, inp.reads[0][:12]] + inp.reads + [e]
        assert _check_runs(ctx, oracle, inp.contigs, reads, k) > 0
    inp = synth.make_config("repeats")
    monkeypatch.setenv("GCG_SEARCH_CHUNK_BYTES", "4096")
    _check_runs(ctx, oracle, inp.contigs, inp.reads, 17)
    monkeypatch.setenv("GCG_SEARCH_RES_CAP", "1000")
    _check_runs(ctx, oracle, inp.contigs, inp.reads, 17)
    monkeypatch.delenv("GCG_SEARCH_CHUNK_BYTES")
    _check_runs(ctx, oracle, inp.contigs, inp.reads, 21)
    monkeypatch.delenv("GCG_SEARCH_RES_CAP")
    # chimeric reads: pieces of distant contigs glued together, some reverse-complemented -> several runs per read
    rng = np.random.default_rng(5)
    inp = synth.make_config("small")
    g = inp.genome
    chim = []
    for _ in range(40):
        parts = []
        for _p in range(int(rng.integers(2, 6))):
            a = int(rng.integers(0, len(g) - 3000)); l = int(rng.integers(200, 3000))
            seg = g[a:a + l]
            parts.append(synth.revcomp(seg) if rng.random() < 0.5 else seg)
        chim.append(np.concatenate(parts))
    assert _check_runs(ctx, oracle, inp.contigs, chim + [e], 25) > len(chim)
    _check_runs(ctx, oracle, inp.contigs, [e, e], 25)
    _check_runs(ctx, oracle, inp.contigs, [], 25)


@pytest.mark.parametrize("pinned,gap", [(False, 0), (True, 0), (True, 3), (False, 1)])
def test_search_of_reads_packed_by_the_caller(ctx, oracle, pinned, gap, monkeypatch):
    """gcg_search_compact_packed (pack on ingest): the caller hands over 2-bit words instead of
    ASCII.  Words in page-locked memory whose reads follow each other go over PCIe from where they lie; pageable
    words, or reads with gaps between them, are copied into the pinned ring first.  Same anchors either way."""
    inp = synth.make_config("small")
    e = np.zeros(0, np.uint8)
    reads = [e] + inp.reads[:70] + [inp.reads[0][:9], e] + inp.reads[70:]
    cs = ctx.upload(inp.contigs)
    t = ctx.table_build(cs, 25)
    want_a, want_off = ctx.search_host_compact(t, reads)
    st = t.stats()
    t.free()
    words, woff, lens, keep = api.Context.pack_reads(reads, pinned=pinned, gap_words=gap)
    for chunk in (None, 8192):
        if chunk:
            monkeypatch.setenv("GCG_SEARCH_CHUNK_BYTES", str(chunk))
        t = ctx.table_build(cs, 25)
        a, off = ctx.search_host_compact_packed(t, words, woff, lens)
        assert np.array_equal(a, want_a) and np.array_equal(off, want_off) and t.stats() == st, (pinned, gap, chunk)
        t.free()
    if keep is not None:
        keep.free()
    cs.free()


@pytest.mark.parametrize("threads", [1, 2, 5])
def test_streaming_search_piece_sizes_and_pool_sizes(oracle, threads, host_pipeline, monkeypatch):
    """the one-launch search with pieces from one read each up to the whole read set, with pools of one (no worker
    thread at all), two and five host threads (the host thread lends a hand below nine), ASCII and packed input
    (pinned words in one run: sent from where they lie; with gaps: through the slots), anchors, 16-byte hits and run
    records: always the anchors of the device-resident search"""
    inp = synth.make_config("small")
    e = np.zeros(0, np.uint8)
    reads = [e, inp.reads[0][:7]] + inp.reads[:60] + [e, e, inp.reads[2][:25]] + inp.reads[60:] + [inp.reads[1][:24], e]
    c = api.Context(0, host_threads=threads)
    try:
        cs, rs = c.upload(inp.contigs), c.upload(reads)
        t = c.table_build(cs, 25)
        want = c.search(t, rs)
        st = t.stats()
        t.free(); rs.free()
        clens = [len(x) for x in inp.contigs]
        packed = {gap: api.Context.pack_reads(reads, pinned=True, gap_words=gap) for gap in (0, 2)}
        for piece in (32, 96, 1000, 5000, 70000, 1 << 22, None):
            if piece is None:
                monkeypatch.delenv("GCG_SEARCH_CHUNK_BYTES", raising=False)
            else:
                monkeypatch.setenv("GCG_SEARCH_CHUNK_BYTES", str(piece))
            t = c.table_build(cs, 25)
            a, off = c.search_host_compact(t, reads)
            assert np.array_equal(api.expand_compact(a, off, clens), want) and t.stats() == st, (threads, piece, "ascii")
            t.free()
            for gap, (words, woff, lens, keep) in packed.items():
                t = c.table_build(cs, 25)
                a2, off2 = c.search_host_compact_packed(t, words, woff, lens)
                assert np.array_equal(a2, a) and np.array_equal(off2, off) and t.stats() == st, (threads, piece, "packed", gap)
                t.free()
            if piece in (96, 70000, None):
                t = c.table_build(cs, 25)
                assert np.array_equal(c.search_host(t, reads), want) and t.stats() == st, (threads, piece, "hits16")
                t.free()
                t = c.table_build(cs, 25)
                runs, run_off, n_anchor = c.search_runs(t, reads)
                assert n_anchor == len(want) and run_off[-1] == len(runs) and t.stats() == st, (threads, piece, "runs")
                t.free()
        for words, woff, lens, keep in packed.values():
            keep.free()
        cs.free()
    finally:
        c.close()


def test_search_right_behind_an_unfinished_table_build(oracle, host_pipeline):
    """gcg_table_build returns while its upload, pack and insert kernels are still queued, and releases the packed
    contigs into the context's block cache; a host-buffer search called right behind it takes blocks from that cache
    for what it writes on the context's stream; what its upload stream writes must NOT come from there (the ready word
    once landed on the released contig lengths and cleared the first contigs under the pending insert kernel): it has a
    block of the pipeline's own.  Many rounds, nothing between the two calls, contig sets of several sizes."""
    import ctypes as C
    inp = synth.make_config("cfg1")
    c = api.Context(0, host_threads=8)
    try:
        reads = [np.ascontiguousarray(r) for r in inp.reads[:400]]
        rptrs = (C.c_void_p * len(reads))(*[a.ctypes.data for a in reads])
        rlens = np.array([len(a) for a in reads], dtype=np.int32)
        for n_ctg in (len(inp.contigs), 3, 17):
            ctgs = [np.ascontiguousarray(x) for x in inp.contigs[:n_ctg]]
            cptrs = (C.c_void_p * len(ctgs))(*[a.ctypes.data for a in ctgs])
            clens = np.array([len(a) for a in ctgs], dtype=np.int32)
            cs = c.upload(ctgs)
            t0 = c.table_build(cs, 25)
            c.sync()
            want_a, want_off = c.search_host_compact(t0, reads)
            want_st = t0.stats()
            t0.free(); cs.free()
            for it in range(40):
                h = C.c_void_p()
                c._chk(c.L.gcg_table_build(c.h, C.cast(cptrs, C.c_void_p), clens.ctypes.data, len(ctgs), 25, C.byref(h)))
                ap, rp, na = C.c_void_p(), C.c_void_p(), C.c_int64()
                c._chk(c.L.gcg_search_compact(c.h, h, C.cast(rptrs, C.c_void_p), rlens.ctypes.data, len(reads), 25, C.byref(ap), C.byref(rp), C.byref(na)))
                tab = api.KmerTable(c, h, 25)
                try:
                    assert na.value == len(want_a), (n_ctg, it, na.value, len(want_a))
                    a = np.frombuffer((C.c_char * (na.value * 8)).from_address(ap.value), dtype=np.uint64)
                    assert np.array_equal(a, want_a) and tab.stats() == want_st, (n_ctg, it)
                    del a
                finally:
                    c.L.gcg_free(ap); c.L.gcg_free(rp); tab.free()
    finally:
        c.close()
