"""GPU parity of the hash-partitioned table (superplus_b200/csrc/part.cu through the C ABI):
route / insert / lookup / collect kernels against the numpy double + CPU oracle, the whole
PartitionedKmerIndex with several partitions on one GPU (ThreadComm), and — when the box has two
GPUs — one process per GPU over NCCL.  Bit exact."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from superplus_b200 import api, synth
from superplus_b200 import dist as gdist

from part_double import HostSeqs, NumpyOps, MISS, owner_np
from test_dist_cpu import oracle_answer, shard

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev_u64(n):
    return torch.empty(max(n, 1), dtype=torch.int64, device="cuda")


def as_u64(t, n):
    return t[:n].cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("name,k,n_part", [("tiny", 25, 1), ("tiny", 25, 2), ("repeats", 17, 3), ("tiny", 31, 8), ("tiny", 5, 16)])
def test_route_kernels_match_double(ctx, oracle, name, k, n_part):
    inp = synth.make_config(name)
    rng = np.random.default_rng(11)
    reads = inp.reads[:12] + [inp.reads[0][:k - 1], inp.reads[1][:k], np.zeros(0, np.uint8), inp.reads[2][:33], inp.reads[3][:1025]]
    dbl = NumpyOps(oracle)
    hs, ds = HostSeqs(reads), ctx.upload(reads)
    assert ds.tiles == hs.tiles
    T = hs.tiles
    cuts = sorted(set([0, T] + [int(x) for x in rng.integers(0, T + 1, size=3)]))
    for t0, t1 in list(zip(cuts, cuts[1:])) + [(0, T), (T, T)]:
        want = dbl.plan(hs, k, n_part, t0, t1)
        got = ctx.route_plan(ds, k, n_part, t0, t1)
        assert got.kmers == want.kmers and np.array_equal(got.counts, want.counts), (t0, t1)
        n = want.kmers
        keys, recs = dev_u64(n), dev_u64(2 * n)
        got.keys(keys.data_ptr()); got.records(recs.data_ptr())
        ctx.sync()
        wk, wr = torch.empty(max(n, 1), dtype=torch.int64), torch.empty(max(2 * n, 1), dtype=torch.int64)
        dbl.route_keys(want, wk); dbl.route_records(want, wr)
        assert np.array_equal(as_u64(keys, n), wk.numpy().view(np.uint64)[:n])
        assert np.array_equal(as_u64(recs, 2 * n), wr.numpy().view(np.uint64)[:2 * n])
        if n:
            # segment d really holds owner-d keys only
            seg = np.concatenate([[0], np.cumsum(want.counts)])
            k0 = as_u64(keys, n) - np.uint64(1)
            for d in range(n_part):
                assert np.all(owner_np(k0[seg[d]:seg[d + 1]], n_part) == d)
        got.free()
    ds.free()


def test_filtered_plan_routes_a_superset_of_the_anchors(ctx, oracle):
    """a filtered plan routes every anchoring position (no false negatives) and few others"""
    inp = synth.make_config("small")
    k = 25
    cs, rs = ctx.upload(inp.contigs), ctx.upload(inp.reads)
    plain = ctx.table_build(cs, k)
    want = ctx.search(plain, rs)
    n_words, k3 = ctx.filter_shape(cs.kmers(k))
    flt = torch.zeros(n_words, dtype=torch.int32, device="cuda")
    plain.filter_add(flt.data_ptr(), n_words, k3)
    ctx.sync()
    for n_part in (1, 3):
        r = ctx.route_plan(rs, k, n_part, 0, rs.tiles, prefilter=(flt.data_ptr(), n_words, k3))
        assert r.positions == rs.kmers(k) and len(want) <= r.kmers < r.positions // 4
        n = r.kmers
        keys, ans = dev_u64(n), dev_u64(n)
        r.keys(keys.data_ptr())
        # every key in one table here, whatever its owner segment: answer them all and collect
        plain2 = ctx.table_build(cs, k)
        plain2.lookup_keys(keys.data_ptr(), n, ans.data_ptr())
        h = r.collect(ans.data_ptr())
        assert np.array_equal(h.download(), want)
        for x in (h, r, plain2):
            x.free()
    for x in (plain, cs, rs):
        x.free()


def test_owner_side_insert_lookup_collect(ctx, oracle):
    """one partition holding everything == the plain table: dump, answers, anchors, statistics"""
    inp = synth.make_config("repeats")
    k = 17
    cs, rs = ctx.upload(inp.contigs), ctx.upload(inp.reads)
    plain = ctx.table_build(cs, k)
    want_hits = ctx.search(plain, rs)
    want_stats = plain.stats()
    r = ctx.route_plan(cs, k, 1, 0, cs.tiles)
    n = int(r.counts.sum())
    recs = dev_u64(2 * n)
    r.records(recs.data_ptr())
    t = ctx.table_create(n, k)
    t.insert_records(recs.data_ptr(), n)
    gk, gm, gt, gp, gr = t.dump()
    pk, pm, pt, pp, pr = plain.dump()
    assert np.array_equal(gk, pk) and np.array_equal(gm, pm)
    u = pm == 1
    assert np.array_equal(gt[u], pt[u]) and np.array_equal(gp[u], pp[u]) and np.array_equal(gr[u], pr[u])
    q = ctx.route_plan(rs, k, 1, 0, rs.tiles)
    nq = int(q.counts.sum())
    keys, ans = dev_u64(nq), dev_u64(nq)
    q.keys(keys.data_ptr())
    t.lookup_keys(keys.data_ptr(), nq, ans.data_ptr())
    h = q.collect(ans.data_ptr())
    got = h.download()
    assert np.array_equal(got, want_hits)
    assert int((as_u64(ans, nq) != MISS).sum()) == len(want_hits)
    assert t.stats() == want_stats
    for x in (h, q, r, t, plain, cs, rs):
        x.free()


@pytest.mark.parametrize("exchange,prefilter", [("all_to_all", False), ("all_to_all", True), ("direct", False), ("direct", True), ("remote", True)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("name,k,round_kmers", [("tiny", 25, 1 << 31), ("repeats", 17, 50_000), ("small", 31, 300_000)])
def test_partitioned_index_one_gpu(oracle, world, name, k, round_kmers, exchange, prefilter):
    inp = synth.make_config(name)
    batches = shard(inp.reads, world)
    want_hits, want_stats = oracle_answer(oracle, inp.contigs, batches, k)

    def make_ops(rank):
        return gdist.DeviceOps(api.Context(0, host_threads=2), 0)

    def body(rank, ops, comm):
        cs, rs = ops.ctx.upload(inp.contigs), ops.ctx.upload(batches[rank])
        idx = gdist.PartitionedKmerIndex(ops, comm, k, round_kmers=round_kmers, exchange=exchange, prefilter=prefilter).build(cs)
        if exchange == "remote":
            idx.search(rs)                  # ... and a rebuild into the same shared blocks (the peers keep their mappings)
            idx.build(cs)
        hits = idx.search(rs)
        st = idx.stats()
        if exchange != "remote":            # (remote probes route nothing: the search kernel reads the owners' partitions where they lie)
            assert idx.n_routed <= idx.n_positions and (prefilter or idx.n_routed == idx.n_positions)
            assert not prefilter or len(hits) <= idx.n_routed < max(idx.n_positions // 2, len(hits) + 1)
        n_dev = idx.search(rs, keep_on_device=True)
        out = (hits, st, idx.n_local_records, n_dev, ops.ctx.launches())
        idx.free(); cs.free(); rs.free()
        ops.ctx.close()
        return out

    out = gdist.run_threaded(world, body, torch.device("cuda", 0), make_ops)
    assert sum(o[2] for o in out) == sum(max(0, len(c) - k + 1) for c in inp.contigs)
    for r in range(world):
        assert np.array_equal(out[r][0], want_hits[r]), "rank %d anchors differ" % r
        assert out[r][1] == want_stats
        assert out[r][3] == len(want_hits[r]) and out[r][4] > 0


NCCL_SCRIPT = textwrap.dedent("""
    import os, sys, json
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %(root)r)
    from superplus_b200 import api, synth
    from superplus_b200 import dist as gdist
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    inp = synth.make_config("small")
    k = 31
    ctx = api.Context(local, host_threads=2)
    ops = gdist.DeviceOps(ctx, local)
    comm = gdist.TorchComm(ops.device)
    cs, rs = ctx.upload(inp.contigs), ctx.upload(inp.reads[rank::world])
    idx = gdist.PartitionedKmerIndex(ops, comm, k, round_kmers=400_000, exchange=%(exchange)r, prefilter=%(prefilter)r).build(cs)
    hits = idx.search(rs)
    st = idx.stats()
    np.save(os.path.join(%(out)r, "hits_%%d.npy" %% rank), hits)
    json.dump({"stats": st, "records": idx.n_local_records, "sent": comm.bytes_sent, "launches": ctx.launches()},
              open(os.path.join(%(out)r, "info_%%d.json" %% rank), "w"))
    idx.free()
    dist.destroy_process_group()
""")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("exchange,prefilter", [("all_to_all", False), ("direct", False), ("direct", True), ("remote", True)])
def test_partitioned_index_nccl_world2(oracle, tmp_path, exchange, prefilter):
    """one process per GPU: NCCL all-to-all, direct stores into CUDA-IPC windows over NVLink, and remote probes of
    the owners' partitions (CUDA-IPC mapped table blocks, loads and atomics over NVLink)"""
    script = tmp_path / "run.py"
    script.write_text(NCCL_SCRIPT % {"root": ROOT, "out": str(tmp_path), "exchange": exchange, "prefilter": prefilter})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    inp = synth.make_config("small")
    want_hits, want_stats = oracle_answer(oracle, inp.contigs, shard(inp.reads, 2), 31)
    for r in range(2):
        info = json.load(open(tmp_path / ("info_%d.json" % r)))
        assert tuple(info["stats"]) == want_stats and (info["sent"] > 0 or exchange == "remote") and info["launches"] > 0
        assert np.array_equal(np.load(tmp_path / ("hits_%d.npy" % r)), want_hits[r])
