"""BASELINE configs[3] at full size: k = 31 k-mer table of a 100 Mb synthetic genome, hash-partitioned
over the GPUs of one box, searched with 40x simulated ONT reads (each rank draws 40/N x of its own).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/cfg4_part.py [genome_mb] [coverage]

Every rank also builds the whole table on its own GPU (replicated layout) and searches its read batch
there: the partitioned search must find the same number of anchors, and the all-reduced statistics
of the partitions must equal the replicated table's contig-side counts — the parity check at a size
the CPU oracle does not reach.  Rank 0 prints one JSON line (device time, max over ranks)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from superplus_b200 import api, synth
from superplus_b200 import dist as gdist

K = 31
genome_mb = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
coverage = float(sys.argv[2]) if len(sys.argv) > 2 else 40.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))


def allred(x, op):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=op)
    return float(t.item())


t0 = time.time()
genome = synth.random_genome(int(genome_mb * 1e6), np.random.Generator(np.random.PCG64(44)))      # the same on every rank
reads = synth.make_reads(genome, coverage / world, seed=44 + 1000 * (rank + 1))
t_gen = time.time() - t0
ctx = api.Context(local_rank, host_threads=max(1, (os.cpu_count() or 8) // world))
ops = gdist.DeviceOps(ctx, local_rank)
comm = gdist.TorchComm(ops.device) if world > 1 else gdist.ThreadComm(gdist.ThreadGroup(1), 0, ops.device, ops.sync)
cs, rs = ctx.upload([genome]), ctx.upload(reads)
n_pos = rs.kmers(K)

# ---- replicated reference on this GPU
tab = ctx.table_build(cs, K)
n_rep = ctx.search_device(tab, rs)
st_rep = tab.stats()
tab.free()

res = {}
for exchange in ("direct", "all_to_all"):
    idx = gdist.PartitionedKmerIndex(ops, comm, K, exchange=exchange)
    idx.build(cs)
    n_part = idx.search(rs, keep_on_device=True)         # warm-up and parity
    st = idx.stats()
    assert n_part == n_rep, "rank %d: partitioned search found %d anchors, the replicated table %d" % (rank, n_part, n_rep)
    assert list(st[:2]) == list(st_rep[:2]), "contig-side statistics differ: %s vs %s" % (st, st_rep)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tb0 = time.perf_counter()
    idx.build(cs)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - tb0
    ctx.prof(True); ctx.prof_reset()
    sent0 = comm.bytes_sent
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ts0 = time.perf_counter()
    for _ in range(steps):
        idx.search(rs, keep_on_device=True)
    torch.cuda.synchronize()
    t_search = (time.perf_counter() - ts0) / steps
    t_search = allred(t_search, dist.ReduceOp.MAX) if world > 1 else t_search
    t_build = allred(t_build, dist.ReduceOp.MAX) if world > 1 else t_build
    prof = ctx.prof_report()
    tot_pos = allred(float(n_pos), dist.ReduceOp.SUM) if world > 1 else float(n_pos)
    res[exchange] = {"search_ms": 1e3 * t_search, "build_ms": 1e3 * t_build, "kmers_per_s": tot_pos / t_search,
                     "routed_fraction": (allred(float(idx.n_routed), dist.ReduceOp.SUM) if world > 1 else idx.n_routed) /
                                        max(1.0, allred(float(idx.n_positions), dist.ReduceOp.SUM) if world > 1 else idx.n_positions),
                     "nvlink_bytes_per_search": (allred(float(comm.bytes_sent - sent0), dist.ReduceOp.SUM) if world > 1 else 0.0) / steps,
                     "stats": [int(x) for x in st],
                     "kernel_ms_per_search": {k_: v[0] / steps for k_, v in sorted(prof.items()) if not k_.startswith(("part_insert", "route_records", "k1_", "filter"))}}
    idx.free()
if rank == 0:
    print(json.dumps({"workload": "configs[3]: k=31 table of a %.0f Mb genome hash-partitioned over %d GPU(s), %.0fx ONT (%.1fx per rank)" % (genome_mb, world, coverage, coverage / world),
                      "n_gpus": world, "contig_kmers": int(st_rep[0]), "ont_kmers_per_rank": int(n_pos), "anchors_rank0": int(n_rep),
                      "data_generation_s": t_gen, "parity": "anchor count per rank and contig-side statistics equal the replicated table's", **res}))
if world > 1:
    dist.destroy_process_group()
