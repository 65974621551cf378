"""CPU: the multi-rank orchestration of superplus_b200.dist (partitioned table, all-to-all
exchange, stats all-reduce) with the numpy test double in place of the device, checked against
the single-table CPU oracle.  World size 2 over gloo (real torch.distributed processes) and
1/2/3/5 partitions over threads."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from superplus_b200 import dist as gdist
from superplus_b200 import synth

from part_double import HostSeqs, NumpyOps, hits_from_oracle, owner_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_answer(oracle, contigs, read_batches, k):
    """single table, all batches searched in one call (the reference's semantics: one table, the
    ONT-side counters run over all reads); anchors split back per batch"""
    h = oracle.table_build(contigs, k)
    st = oracle.table_stats(h)
    o_hits, ont = oracle.search(h, [r for b in read_batches for r in b], k)
    oracle.table_free(h)
    allh = hits_from_oracle(o_hits)
    hits, first = [], 0
    for b in read_batches:
        sel = (allh["read"] >= first) & (allh["read"] < first + len(b))
        part = allh[sel].copy()
        part["read"] -= first
        hits.append(part)
        first += len(b)
    return hits, st + ont


def shard(reads, world):
    return [reads[r::world] for r in range(world)]


@pytest.mark.parametrize("exchange,prefilter", [("all_to_all", False), ("all_to_all", True), ("direct", False), ("direct", True)])
@pytest.mark.parametrize("world", [1, 2, 3, 5])
@pytest.mark.parametrize("name,k", [("tiny", 25), ("repeats", 17)])
def test_partitioned_index_threads(oracle, world, name, k, exchange, prefilter):
    inp = synth.make_config(name)
    batches = shard(inp.reads, world)
    want_hits, want_stats = oracle_answer(oracle, inp.contigs, batches, k)

    def body(rank, ops, comm):
        idx = gdist.PartitionedKmerIndex(ops, comm, k, round_kmers=40_000, exchange=exchange, prefilter=prefilter).build(HostSeqs(inp.contigs))
        hits = idx.search(HostSeqs(batches[rank]))
        assert idx.n_routed <= idx.n_positions and (prefilter or idx.n_routed == idx.n_positions)
        assert not prefilter or world == 1 or idx.n_routed < idx.n_positions // 2      # the filter really thins the exchange
        out = hits, idx.stats(), idx.n_local_records, comm.bytes_sent
        idx.free()
        return out

    out = gdist.run_threaded(world, body, torch.device("cpu"), lambda r: NumpyOps(oracle))
    n_ctg_kmers = sum(max(0, len(c) - k + 1) for c in inp.contigs)
    assert sum(o[2] for o in out) == n_ctg_kmers              # every contig k-mer landed on exactly one owner
    for r in range(world):
        assert np.array_equal(out[r][0], want_hits[r]), "rank %d anchors differ" % r
        assert out[r][1] == want_stats
    if world > 1:
        assert all(o[3] > 0 for o in out)


def test_owner_matches_library():
    from superplus_b200 import api
    L = api.load_library()
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 1 << 62, size=2000, dtype=np.uint64)
    for n_part in (1, 2, 3, 8, 16):
        got = np.array([L.gcg_kmer_owner(int(x), n_part) for x in keys])
        assert np.array_equal(got, owner_np(keys, n_part))
        assert got.min() >= 0 and got.max() < n_part
    assert L.gcg_kmer_owner(1, 0) == -1 and L.gcg_kmer_owner(1, 17) == -1
    # balanced enough to size the owners' tables from the mean
    cnt = np.bincount(owner_np(keys, 8), minlength=8)
    assert cnt.min() > 150


def test_round_and_slice_helpers():
    assert gdist.tile_slice(10, 0, 3) == (0, 3) and gdist.tile_slice(10, 2, 3) == (6, 10)
    assert [gdist.tile_slice(7, r, 8) for r in range(8)][-1] == (6, 7)
    assert gdist.search_rounds(0) == [(0, 0)]
    r = gdist.search_rounds(100, round_kmers=10 * 1024)
    assert r[0] == (0, 9) and r[-1][1] == 100 and all(a[1] == b[0] for a, b in zip(r, r[1:]))


GLOO_SCRIPT = textwrap.dedent("""
    import os, sys, json
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    from oracle.oracle import Oracle
    from superplus_b200 import dist as gdist, synth
    from part_double import HostSeqs, NumpyOps
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    inp = synth.make_config("tiny")
    k = 25
    comm = gdist.TorchComm(torch.device("cpu"))
    idx = gdist.PartitionedKmerIndex(NumpyOps(Oracle()), comm, k, round_kmers=60_000).build(HostSeqs(inp.contigs))
    hits = idx.search(HostSeqs(inp.reads[rank::world]))
    st = idx.stats()
    np.save(os.path.join(%(out)r, "hits_%%d.npy" %% rank), hits)
    json.dump({"stats": st, "records": idx.n_local_records, "sent": comm.bytes_sent}, open(os.path.join(%(out)r, "info_%%d.json" %% rank), "w"))
    dist.destroy_process_group()
""")


def test_partitioned_index_gloo_world2(oracle, tmp_path):
    import json
    script = tmp_path / "run.py"
    script.write_text(GLOO_SCRIPT % {"root": ROOT, "out": str(tmp_path)})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    inp = synth.make_config("tiny")
    want_hits, want_stats = oracle_answer(oracle, inp.contigs, shard(inp.reads, 2), 25)
    recs = 0
    for r in range(2):
        info = json.load(open(tmp_path / ("info_%d.json" % r)))
        assert tuple(info["stats"]) == want_stats
        assert info["sent"] > 0
        recs += info["records"]
        assert np.array_equal(np.load(tmp_path / ("hits_%d.npy" % r)), want_hits[r])
    assert recs == sum(max(0, len(c) - 25 + 1) for c in inp.contigs)
