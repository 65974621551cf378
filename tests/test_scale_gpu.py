"""GPU, BASELINE configs[1] at full size (cfg2: 4.6 Mb scaffolds, 500 gaps, 30x ONT, 138 M ONT
k-mers): the oracle needs minutes there, so parity is checked through size-independent properties
  * the three product paths agree bit for bit: device-resident search, the pipelined host-buffer
    search (many chunks, slot reuse) and the hash-partitioned table with 4 partitions;
  * table statistics equal an independent numpy count over the oracle's canonical contig k-mers;
  * every anchor is to a contig k-mer that occurs exactly once, and spells the same k-mer as the
    read position (reverse-complemented when the strand flags differ);
  * completeness on a sample of reads: their anchors are exactly the positions whose canonical
    k-mer is unique in the contigs (oracle chop + numpy set membership).
"""
import numpy as np
import pytest
import torch

from superplus_b200 import api, synth
from superplus_b200 import dist as gdist

pytestmark = pytest.mark.gpu
K = 25


def revcomp_ascii(a):
    comp = np.zeros(256, dtype=np.uint8)
    for x, y in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
        comp[x] = y
    return comp[a][::-1]


def test_cfg2_full_size_properties(ctx, oracle):
    inp = synth.make_config("cfg2")
    contigs, reads = inp.contigs, inp.reads
    n_kmers = sum(max(0, len(r) - K + 1) for r in reads)
    assert n_kmers > 130_000_000

    cs, rs = ctx.upload(contigs), ctx.upload(reads)
    t = ctx.table_build(cs, K)
    hits = ctx.search(t, rs)
    stats = t.stats()
    t.free()

    # ---- the host-buffer pipeline gives the same anchors and statistics
    t = ctx.table_build(cs, K)
    hits_host = ctx.search_host(t, reads)
    assert np.array_equal(hits_host, hits)
    assert t.stats() == stats
    t.free()

    # ---- independent statistics: distinct / unique canonical contig k-mers
    ck = np.concatenate([oracle.chop(c, K)[0] for c in contigs if len(c) >= K])
    uniq, cnt = np.unique(ck, return_counts=True)
    assert stats[0] == len(uniq) and stats[1] == int((cnt == 1).sum())
    once = uniq[cnt == 1]

    # ---- anchors are ordered, unique-in-contig, and spell the read's k-mer
    key = hits["read"].astype(np.int64) << 32 | hits["pos"].astype(np.int64)
    assert np.all(np.diff(key) > 0)
    rng = np.random.default_rng(7)
    for i in rng.choice(len(hits), size=3000, replace=False):
        h = hits[i]
        r = reads[h["read"]][h["pos"]: h["pos"] + K]
        cpos, krev, orev = int(h["cpos_flags"]) >> 2, int(h["cpos_flags"]) & 1, (int(h["cpos_flags"]) >> 1) & 1
        c = contigs[h["tid"]][cpos: cpos + K]
        assert len(r) == K and len(c) == K
        same = np.array_equal(np.char.upper(r.view("S1")), np.char.upper(c.view("S1")))
        rc = np.array_equal(np.char.upper(revcomp_ascii(r).view("S1")), np.char.upper(c.view("S1")))
        assert (same if krev == orev else rc), (i, h)
    # the ONT-side statistics follow from the anchors themselves
    slot = hits["tid"].astype(np.int64) << 32 | (hits["cpos_flags"] >> 2).astype(np.int64)
    u2, c2 = np.unique(slot, return_counts=True)
    assert stats[2] == len(u2) and stats[3] == int((c2 == 1).sum())

    # ---- completeness on a sample of reads
    for ri in rng.choice(len(reads), size=40, replace=False):
        ks, _ = oracle.chop(reads[ri], K)
        want = np.nonzero(np.isin(ks, once))[0].astype(np.int32)
        got = hits["pos"][hits["read"] == ri]
        assert np.array_equal(got, want), ri

    # ---- the hash-partitioned table (4 partitions on this GPU, direct exchange) agrees
    world = 4
    batches = [reads[r::world] for r in range(world)]

    def body(rank, ops, comm):
        c2_, r2_ = ops.ctx.upload(contigs), ops.ctx.upload(batches[rank])
        idx = gdist.PartitionedKmerIndex(ops, comm, K, exchange="direct").build(c2_)
        h = idx.search(r2_)
        st = idx.stats()
        idx.free(); c2_.free(); r2_.free(); ops.ctx.close()
        return h, st

    out = gdist.run_threaded(world, body, torch.device("cuda", 0), lambda r: gdist.DeviceOps(api.Context(0, host_threads=2), 0))
    for rank in range(world):
        sel = hits[hits["read"] % world == rank].copy()
        sel["read"] //= world
        assert np.array_equal(out[rank][0], sel), rank
        assert out[rank][1] == stats
    cs.free(); rs.free()
