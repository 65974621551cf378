#!/usr/bin/env python
"""bench.py — headline measurement of the gap-closing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path over one batch of synthetic input:
  k-mer part (the JSON line's `value`): BASELINE.json configs[1] — 4.6 Mb synthetic scaffold, 500
      N-gaps, 30x simulated ONT (10 % error), k = 25: 2-bit pack of contigs and reads, contig
      k-mer table build, ONT search with anchor emission in (read,pos) order, statistics.
      Unit = ONT k-mers looked up per second (whole job, all ranks).
  SW part (the `sw` object): BASELINE.json configs[2] shape — ONT-read(10 kb) x gap-flank(2 kb)
      pairs, default scoring of gc_graph.c:74-77, fill + trace spill + end cell + CIGAR, as
      many pairs per step as fill two whole waves of resident warps.  Unit = GCUPS.
`value` is measured with the inputs resident in HBM (CUDA events on the library's stream, max
over ranks); `e2e` is the same metric through the host-buffer C-ABI calls the C shims use
(gcg_table_build + gcg_search_compact, gcg_sw_batch), host<->device copies inside the timed region.
Anchors are the compact 8-byte records of include/gcgpu.h in both (SURVEY 8d's hit record).
The SW headline is the FIXED traceback (configs[2] is the fixed-mode config); the as-shipped
traceback is timed beside it (`sw.asis`).  `roofline_hbm_table` repeats the k-mer roofline on a
100 Mb / k = 31 table that does not fit the L2 (BASELINE configs[3]'s table on one GPU).
N > 1: one process per GPU; reads / pairs are sharded by batch, the contig table is replicated
(SURVEY §8e), no collective on the data path -> weak scaling.

--impl reference times the reference's own CPU code (oracle/_ref, built unmodified from
/root/reference/gap_closer) on the box's host cores, same metric and config, bounded sample.
"""
import argparse
import json
import os
import statistics
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from superplus_b200 import synth  # noqa: E402

K = 25
WORKLOAD = "cfg2"
SW_QLEN, SW_TLEN = 10_000, 2_000
BYTES_PER_LOOKUP, BYTES_PER_HIT = 16.25, 16.0        # SURVEY §8(d) algorithmic bytes
OPS_PER_CELL = 12.0                                  # SURVEY §8(d) integer ops per DP cell


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """per-launch DRAM bytes of `kernel` from the committed ncu capture, or None"""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons through NVML while the timed regions run.
    NVML queries are not free for the GPU under load: polling at 20 Hz slowed the SW step by 19 %
    (scripts/sampler_cost.py: 128 -> 153 ms), at 1 Hz by nothing measurable.  So the thread polls
    once a second and every timed region asks for one extra sample right after it started
    (`mark()`), which covers the k-mer region that only lasts milliseconds."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        self.period = 1.0
        self.wake = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def mark(self):
        self.wake.set()

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.wake.wait(self.period)
            self.wake.clear()

    def result(self):
        self.stop_flag = True
        self.wake.set()
        if self.nv is None or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "how": "NVML, one sample at the start of every timed region + 1 Hz"}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_workload(rank):
    """cfg2; rank r > 0 draws its own read batch (same scaffold) — reads are sharded by batch"""
    cfg = synth.CONFIGS[WORKLOAD]
    if rank == 0:
        inp = synth.make_config(WORKLOAD)
        return inp.contigs, inp.reads, inp
    base = synth.make_gap_closer_input(cfg["genome_len"], cfg["n_gaps"], 0.0, cfg["seed"])
    reads = synth.make_reads(base.genome, cfg["coverage"], cfg["seed"] + 1000 * rank)
    return base.contigs, reads, base


def reference_available():
    from oracle import oracle as orc
    return orc.have_ref()


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own code on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_kmer(contig_scaffold, reads, n_thread, tmp):
    """runs oracle/_ref/ref_kmer (reference contig.c/rseq.c/kmer.c/hash.c/ont.c) and returns
    (ONT k-mers per second over chop+put+search, info dict)"""
    from oracle import oracle as orc
    fa, fq = os.path.join(tmp, "b.fa"), os.path.join(tmp, "b.fq")
    synth.write_fasta(fa, contig_scaffold)
    synth.write_fastq(fq, reads)
    info, _, _, _ = orc.run_ref_kmer(fa, fq, K, os.path.join(tmp, "b"), n_thread=n_thread, dump=0)
    t = info["t_chop"] + info["t_put"] + info["t_search"]
    return info["n_ont_kmers"] / t, info


def cpu_sw(n_thread, pairs_per_thread, seed=46):
    from oracle import oracle as orc
    n = n_thread * pairs_per_thread
    q, t = synth.make_sw_pairs(n, SW_QLEN, SW_TLEN, seed=seed)
    R = orc.RefSW("asis")
    sec, scores = R.bench(n_thread, q, t)
    R.close()
    return n * SW_QLEN * SW_TLEN / sec / 1e9, n, scores


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if not reference_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built (needs /root/reference at build time)"}))
        return
    from oracle import oracle as orc
    inp = synth.make_config(WORKLOAD)
    vals, sw_vals, times = [], [], []
    budget_s = 150.0                      # the whole arm has to end within a few minutes
    with tempfile.TemporaryDirectory() as tmp:
        fa, fq = os.path.join(tmp, "b.fa"), os.path.join(tmp, "b.fq")
        synth.write_fasta(fa, inp.scaffold)
        synth.write_fastq(fq, inp.reads)                         # the FULL cfg2 read set: same config as our arm
        t_start = time.perf_counter()
        warm, steps = min(args.warmup, 1), args.steps
        it = 0
        while it < warm + steps:
            t0 = time.perf_counter()
            info, _, _, _ = orc.run_ref_kmer(fa, fq, K, os.path.join(tmp, "b"), n_thread=cores, dump=0)
            v = info["n_ont_kmers"] / (info["t_chop"] + info["t_put"] + info["t_search"])
            g, npairs, _ = cpu_sw(cores, 1)
            dt = time.perf_counter() - t0
            if it >= warm:
                vals.append(v); sw_vals.append(g); times.append(dt)
            elif warm:
                # a step reloads the FASTQ with the reference's own loader and re-zeroes 16 bytes per ONT base before
                # the timed phases: fit as many steps as the budget allows, at least one
                steps = max(1, min(args.steps, int((budget_s - dt) / dt)))
            it += 1
            if it >= warm + 1 and time.perf_counter() - t_start > budget_s:
                break
    value = statistics.mean(vals)
    sample = ("the full cfg2 workload per step (%d reads, %d ONT k-mers, 4.6 Mb scaffold); chop+put+search timed inside ref_kmer "
              "(loading and okseq set-up outside); %d of the %d requested steps fit the %.0f s budget of this arm" %
              (len(inp.reads), info["n_ont_kmers"], len(vals), args.steps, budget_s))
    line = {
        "impl": "reference", "metric": "kmers_per_s", "value": value, "unit": "k-mers/s", "n_gpus": args.gpus, "steps": len(vals),
        "steps_requested": args.steps, "warmup": warm, "ms_per_step": 1e3 * statistics.mean(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "cfg2: 4.6 Mb synthetic scaffolds, 500 N-gaps, 30x ONT (10% error), k=25; reference CPU code (kmer.c, hash.c, ont.c) on the host cores",
                   "k": K, "threads": cores, "reads": len(inp.reads), "ont_kmers": info["n_ont_kmers"]},
        "cpu_baseline": {"value": value, "unit": "k-mers/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sw": {"metric": "sw_gcups", "value": statistics.mean(sw_vals), "unit": "GCUPS", "impl": "reference",
               "cpu_baseline": {"value": statistics.mean(sw_vals), "unit": "GCUPS", "cores": cores, "kind": "reference",
                                "sample": "%d pairs of 10 kb x 2 kb per step, one reference sw_t per core (sw.c:400-414), as-shipped traceback" % cores},
               "e2e": {"value": statistics.mean(sw_vals), "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    from superplus_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    torch.cuda.set_device(local_rank)
    host_threads = min(16, max(1, (os.cpu_count() or 8) // max(world, 1)))
    ctx = api.Context(local_rank, host_threads=host_threads)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=torch.device("cuda", local_rank))

    contigs, reads, inp = make_workload(rank)
    n_ont_kmers = sum(max(0, len(r) - K + 1) for r in reads)
    n_ctg_kmers = sum(max(0, len(c) - K + 1) for c in contigs)
    read_bytes, ctg_bytes = sum(len(r) for r in reads), sum(len(c) for c in contigs)

    # inputs resident in HBM before the timed region (ASCII, as the reference's inputs are)
    a_ctg, a_reads = ctx.stage_ascii(contigs), ctx.stage_ascii(reads)

    def kmer_step():
        cs = ctx.pack(a_ctg)
        rs = ctx.pack(a_reads)
        t = ctx.table_build(cs, K)
        n_hit = ctx.search_device_compact(t, rs)
        st = t.stats()
        t.free(); cs.free(); rs.free()
        return n_hit, st

    # SW batch resident in HBM: two whole waves of the packed kernel (5920 pairs each on a 148-SM part)
    sw_pairs = args.sw_pairs
    bq, bt = synth.make_sw_pairs(min(sw_pairs, 512), SW_QLEN, SW_TLEN, seed=46 + rank)
    reps = (sw_pairs + len(bq) - 1) // len(bq)
    q2, t2 = np.tile(bq, (reps, 1))[:sw_pairs], np.tile(bt, (reps, 1))[:sw_pairs]
    swb = ctx.swbatch_upload(q2, t2)
    P = api.make_sw_params()
    sw_cells = swb.cells()

    def sw_step(mode=api.SW_FIXED):
        swb.align(P, mode)

    # ---- warm-up
    for _ in range(args.warmup):
        n_hit, st = kmer_step()
        sw_step(api.SW_FIXED)
        sw_step(api.SW_ASIS)
    r_int16 = ctx.ubench_int16()

    # ---- timed: k-mer steps
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.prof(True); ctx.prof_reset()
    launches0 = ctx.launches()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sampler.mark()
    for _ in range(args.steps):
        n_hit, st = kmer_step()
    e1.record(stream)
    barrier()
    kmer_ms = allmax(e0.elapsed_time(e1)) / args.steps
    kprof = ctx.prof_report()
    kmer_launches = ctx.launches() - launches0

    # ---- hash-partitioned table (BASELINE configs[3] / SURVEY 8e): same step, the table split over the
    #      ranks by hash of the canonical k-mer, records / keys / answers exchanged by all-to-all
    part = {}
    if world > 1 or args.partitioned:
        from superplus_b200 import dist as gdist
        ops = gdist.DeviceOps(ctx, local_rank)
        comm = gdist.TorchComm(ops.device) if dist is not None else gdist.ThreadComm(gdist.ThreadGroup(1), 0, ops.device, ops.sync)
        for exchange in ("direct", "remote", "all_to_all"):
            # one index per exchange mode; every step rebuilds its table from the packed contigs (as the
            # replicated leg does) while the exchange windows of the direct mode stay mapped
            idx = gdist.PartitionedKmerIndex(ops, comm, K, exchange=exchange)

            def part_step():
                cs = ctx.pack(a_ctg)
                rs = ctx.pack(a_reads)
                idx.build(cs)
                nh = idx.search(rs, keep_on_device=True)
                st4 = idx.stats()
                if idx.profile and rank == 0:
                    print("[dist profile %s, ms] " % exchange + "  ".join("%s %.2f" % kv for kv in sorted(idx.timers.items())), file=sys.stderr)
                    idx.timers.clear()
                cs.free(); rs.free()
                return nh, st4

            for _ in range(args.warmup):
                p_hit, p_st = part_step()
            ctx.prof_reset()
            sent0, pl0 = comm.bytes_sent, ctx.launches()
            barrier()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(stream)
            sampler.mark()
            for _ in range(args.steps):
                p_hit, p_st = part_step()
            p1.record(stream)
            barrier()
            part_ms = allmax(p0.elapsed_time(p1)) / args.steps
            pprof = ctx.prof_report()
            assert p_hit == n_hit, "partitioned and replicated tables disagree on this rank's anchors"
            part[exchange] = {"ms_per_step": part_ms, "launches": ctx.launches() - pl0, "stats": list(p_st),
                              "routed_fraction": allsum(float(idx.n_routed)) / max(1.0, allsum(float(idx.n_positions))),
                              "bytes_sent_per_step": allsum(float(comm.bytes_sent - sent0)) / args.steps,
                              "kernel_ms_per_step": {k_: v[0] / args.steps for k_, v in sorted(pprof.items())}}
            idx.free()

    # ---- BASELINE configs[3] at FULL size (default at N = 8, --cfg4 on|off elsewhere): k = 31 table of a 100 Mb genome
    #      hash-partitioned over the ranks, 40x ONT in total (40 / N per rank), direct exchange over NVLink peer memory.
    #      Checked against the CPU oracle, not against ourselves: the all-reduced scaffold statistics against the
    #      oracle's table over the whole genome, and the anchors of a sample of rank 0's reads against oracle.search.
    cfg4 = None
    if (world == 8 and args.cfg4 != "off") or args.cfg4 == "on":
        from superplus_b200 import dist as gdist
        K4 = 31
        c4 = synth.CONFIGS["cfg4"]
        G4 = int(args.cfg4_genome or c4["genome_len"])
        rng4 = np.random.Generator(np.random.PCG64(c4["seed"]))
        genome4 = synth.random_genome(G4, rng4)
        reads4 = synth.make_reads(genome4, c4["coverage"] / world, c4["seed"] + 1 + 1000 * rank)
        n4 = sum(max(0, len(r) - K4 + 1) for r in reads4)
        ops4 = gdist.DeviceOps(ctx, local_rank)
        comm4 = gdist.TorchComm(ops4.device) if dist is not None else gdist.ThreadComm(gdist.ThreadGroup(1), 0, ops4.device, ops4.sync)
        cs4, rs4 = ctx.upload([genome4]), ctx.upload(reads4)
        sample = reads4[:48]
        ss4 = ctx.upload(sample)
        cfg4_all = {}
        for x4 in [args.cfg4_exchange]:
            idx4 = gdist.PartitionedKmerIndex(ops4, comm4, K4, exchange=x4)
            idx4.build(cs4)
            got = idx4.search(ss4)                                  # the sample's anchors (collective: every rank searches its own sample)
            st_after_sample = idx4.stats()
            idx4.build(cs4)                                         # fresh ONT-side counts for the timed searches
            barrier()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record(stream)
            idx4.build(cs4)
            b1.record(stream)
            barrier()
            build4_ms = allmax(b0.elapsed_time(b1))
            for _ in range(2):
                h4 = idx4.search(rs4, keep_on_device=True)
            ctx.prof_reset()
            sent0 = comm4.bytes_sent
            barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record(stream)
            sampler.mark()
            c4_steps = 3
            for _ in range(c4_steps):
                h4 = idx4.search(rs4, keep_on_device=True)
            q1.record(stream)
            barrier()
            search4_ms = allmax(q0.elapsed_time(q1)) / c4_steps
            prof4 = ctx.prof_report()
            tot4, hits4 = allsum(float(n4)), allsum(float(h4))
            nv4 = allsum(float(comm4.bytes_sent - sent0)) / c4_steps
            routed4 = allsum(float(idx4.n_routed)) / max(1.0, allsum(float(idx4.n_positions)))
            check = {"kind": "not run (oracle library missing)"}
            if rank == 0:
                try:
                    from oracle import oracle as orc
                    O = orc.Oracle()
                    t_or = time.perf_counter()
                    h_or = O.table_build([genome4], K4)
                    want_hits, _ = O.search(h_or, sample, K4)
                    want_st = O.table_stats(h_or)
                    O.table_free(h_or)
                    ok_st = tuple(st_after_sample[:2]) == tuple(want_st)
                    ok_hits = (len(got) == len(want_hits["read"]) and np.array_equal(got["read"], want_hits["read"].astype(np.int32)) and
                               np.array_equal(got["pos"], want_hits["pos"]) and np.array_equal(got["tid"], want_hits["tid"]) and
                               np.array_equal(got["cpos_flags"] >> 2, want_hits["cpos"].astype(np.uint32)) and
                               np.array_equal(got["cpos_flags"] & 1, want_hits["krev"]) and np.array_equal((got["cpos_flags"] >> 1) & 1, want_hits["orev"]))
                    check = {"kind": "CPU oracle (oracle/gc_oracle.c) over the whole genome", "scaffold_stats_equal": bool(ok_st), "oracle_stats": list(want_st),
                             "sample_reads": len(sample), "sample_anchors": int(len(got)), "sample_anchors_equal": bool(ok_hits), "oracle_s": time.perf_counter() - t_or}
                    assert ok_st and ok_hits, "partitioned cfg4 table disagrees with the CPU oracle: %r" % (check,)
                except ImportError:
                    pass
            cfg4_all[x4] = {"metric": "kmers_per_s", "value": tot4 / (search4_ms * 1e-3), "unit": "k-mers/s", "ms_per_search": search4_ms, "build_ms": build4_ms,
                    "config": {"workload": "BASELINE configs[3]: %d Mb synthetic genome, k=31, %.0fx ONT in total (%.2fx = %d reads per GPU), contig table hash-partitioned over %d GPU(s), exchange '%s' (remote = one search kernel per rank probing the owners' partitions over NVLink peer memory)" %
                               (G4 // 1_000_000, c4["coverage"], c4["coverage"] / world, len(reads4), world, x4),
                               "ont_kmers_total": int(tot4), "anchors_total": int(hits4), "stats": list(idx4.stats())},
                    "nvlink_bytes_per_search": nv4, "routed_fraction": routed4, "check": check,
                    "kernel_ms_per_search": {k_: v[0] / c4_steps for k_, v in sorted(prof4.items())}}
            idx4.free()
        cfg4 = cfg4_all
        cs4.free(); rs4.free(); ss4.free()
        del genome4, reads4

    # ---- timed: SW steps, fixed traceback (headline) and the as-shipped traceback beside it
    sw_launches = 0
    sw_timed = {}
    for mode, name in ((api.SW_FIXED, "fixed"), (api.SW_ASIS, "asis")):
        ctx.prof_reset()
        launches1 = ctx.launches()
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        sampler.mark()
        for _ in range(args.steps):
            sw_step(mode)
        e3.record(stream)
        barrier()
        sw_timed[name] = (allmax(e2.elapsed_time(e3)) / args.steps, ctx.prof_report())
        sw_launches += ctx.launches() - launches1
    sw_ms, sprof = sw_timed["fixed"]
    sw_asis_ms, sprof_asis = sw_timed["asis"]
    ctx.prof(False)
    clocks = sampler.result()

    # ---- the k-mer roofline again on a table that does not fit the L2: 100 Mb genome, k = 31 (BASELINE configs[3]'s
    #      table, 1.6 GB of keys + 1.6 GB of values on this GPU), 2x ONT; device resident, per-kernel CUDA events
    hbm_leg = None
    if world == 1 and not args.no_hbm_table:
        K4 = 31
        rng4 = np.random.Generator(np.random.PCG64(synth.CONFIGS["cfg4"]["seed"]))
        genome4 = synth.random_genome(synth.CONFIGS["cfg4"]["genome_len"], rng4)
        reads4 = synth.make_reads(genome4, args.hbm_coverage, synth.CONFIGS["cfg4"]["seed"] + 1)
        n4 = sum(max(0, len(r) - K4 + 1) for r in reads4)
        c4, r4 = ctx.upload([genome4]), ctx.upload(reads4)
        t4 = ctx.table_build(c4, K4)
        for _ in range(2):
            h4 = ctx.search_device_compact(t4, r4)
        ctx.prof(True); ctx.prof_reset()
        h_steps = 3
        for _ in range(h_steps):
            h4 = ctx.search_device_compact(t4, r4)
        ctx.sync()
        p4 = ctx.prof_report()
        ctx.prof(False)
        st4 = t4.stats()
        t4.free(); c4.free(); r4.free()
        f_ms = p4.get("k45_fused", (0.0, 1))
        f_avg = f_ms[0] / max(1, f_ms[1])
        alg4 = BYTES_PER_LOOKUP * n4 + BYTES_PER_HIT * h4
        hbm_leg = {"kernel": "k45_fused_kernel", "avg_launch_ms": f_avg, "ont_kmers": n4, "anchors": int(h4), "stats": list(st4),
                   "algorithmic_bytes_per_launch": alg4, "kmers_per_s": n4 / (f_avg * 1e-3) if f_avg > 0 else 0.0,
                   "achieved": alg4 / (f_avg * 1e-3) / 1e9 if f_avg > 0 else 0.0,
                   "kernel_ms_per_launch": {k_: v[0] / max(1, v[1]) for k_, v in sorted(p4.items())},
                   "workload": "100 Mb random genome, k=31: 100 M contig k-mers, 3.2 GB of slots (beyond the 126 MB L2, Bloom pre-filter on), %.1fx ONT (%d reads)" % (args.hbm_coverage, len(reads4))}
        del genome4, reads4

    # ---- end to end through the host-buffer C ABI (what the C shims call)
    arrs = [np.ascontiguousarray(r) for r in reads]
    import ctypes as C
    rptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    rlens = np.array([len(a) for a in arrs], dtype=np.int32)
    carrs = [np.ascontiguousarray(c) for c in contigs]
    cptrs = (C.c_void_p * max(1, len(carrs)))(*[a.ctypes.data for a in carrs])
    clens = np.array([len(a) for a in carrs], dtype=np.int32)
    e2e_steps = max(2, min(args.steps, 5))
    e2e_hits = 0

    def e2e_kmer():
        h = C.c_void_p()
        ctx._chk(ctx.L.gcg_table_build(ctx.h, C.cast(cptrs, C.c_void_p), clens.ctypes.data, len(carrs), K, C.byref(h)))
        tab = api.KmerTable(ctx, h, K)
        hp, rp, nh = C.c_void_p(), C.c_void_p(), C.c_int64()
        ctx._chk(ctx.L.gcg_search_compact(ctx.h, tab.h, C.cast(rptrs, C.c_void_p), rlens.ctypes.data, len(arrs), K, C.byref(hp), C.byref(rp), C.byref(nh)))
        # the anchors and the per-read offsets are now in (library-owned, pinned) host memory: read the ends, release
        ro = np.frombuffer((C.c_char * ((len(arrs) + 1) * 8)).from_address(rp.value), dtype=np.int64)
        assert int(ro[-1]) == nh.value and int(ro[0]) == 0
        if nh.value:
            view = np.frombuffer((C.c_char * (nh.value * 8)).from_address(hp.value), dtype=np.uint64)
            assert (int(view[-1]) >> 36) < len(arrs[-1]) or int(ro[-2]) == int(ro[-1])
            del view
        del ro
        ctx.L.gcg_free(hp); ctx.L.gcg_free(rp)
        s4 = tab.stats()
        tab.free()
        return nh.value, s4

    for _ in range(3):          # untimed warm-up calls (pinned blocks parked, workers awake), as for the device-resident step
        e2e_kmer()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_hits, e2e_st = e2e_kmer()
    barrier()
    e2e_kmer_s = allmax(time.perf_counter() - t0) / e2e_steps
    assert e2e_hits == n_hit and e2e_st == st, "e2e and device-resident paths disagree"

    # the same step through the N3 entry point (GC_RUNS mode of the shim): the anchors are reduced on the device to the
    # run records map_ont2contigs needs, so the device -> host stream is a few hundred kilobytes instead of 8 bytes per anchor
    def e2e_runs():
        h = C.c_void_p()
        ctx._chk(ctx.L.gcg_table_build(ctx.h, C.cast(cptrs, C.c_void_p), clens.ctypes.data, len(carrs), K, C.byref(h)))
        tab = api.KmerTable(ctx, h, K)
        rp, op, nr, na = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        ctx._chk(ctx.L.gcg_search_runs(ctx.h, tab.h, C.cast(rptrs, C.c_void_p), rlens.ctypes.data, len(arrs), K, C.byref(rp), C.byref(op), C.byref(nr), C.byref(na)))
        ro = np.frombuffer((C.c_char * ((len(arrs) + 1) * 8)).from_address(op.value), dtype=np.int64)
        assert int(ro[-1]) == nr.value
        del ro
        ctx.L.gcg_free(rp); ctx.L.gcg_free(op)
        s4 = tab.stats()
        tab.free()
        return na.value, nr.value, s4

    # ... and with the reads as the FASTQ loader of the B200 build leaves them (pack on ingest): 2-bit
    # words in page-locked host memory, packed once while loading — outside the step, as loading is for every arm
    pk_words, pk_woff, pk_lens, pk_keep = api.Context.pack_reads(arrs, pinned=True)

    def e2e_packed():
        h = C.c_void_p()
        ctx._chk(ctx.L.gcg_table_build(ctx.h, C.cast(cptrs, C.c_void_p), clens.ctypes.data, len(carrs), K, C.byref(h)))
        tab = api.KmerTable(ctx, h, K)
        hp, rp, nh = C.c_void_p(), C.c_void_p(), C.c_int64()
        ctx._chk(ctx.L.gcg_search_compact_packed(ctx.h, tab.h, pk_words.ctypes.data, pk_woff.ctypes.data, pk_lens.ctypes.data, len(arrs), K, C.byref(hp), C.byref(rp), C.byref(nh)))
        ro = np.frombuffer((C.c_char * ((len(arrs) + 1) * 8)).from_address(rp.value), dtype=np.int64)
        assert int(ro[-1]) == nh.value
        del ro
        ctx.L.gcg_free(hp); ctx.L.gcg_free(rp)
        s4 = tab.stats()
        tab.free()
        return nh.value, s4

    for _ in range(3):          # untimed warm-up calls (pinned blocks parked, workers awake), as for the device-resident step
        e2e_packed()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        p_hits, p_st = e2e_packed()
    barrier()
    e2e_packed_s = allmax(time.perf_counter() - t0) / e2e_steps
    assert p_hits == n_hit and p_st == st, "packed-input and ASCII-input paths disagree"

    # both together — what gc_b200 does with GC_RUNS=1: packed reads up (8 bytes per 32 bases), run records down
    def e2e_runs_packed():
        h = C.c_void_p()
        ctx._chk(ctx.L.gcg_table_build(ctx.h, C.cast(cptrs, C.c_void_p), clens.ctypes.data, len(carrs), K, C.byref(h)))
        tab = api.KmerTable(ctx, h, K)
        rp, op, nr, na = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        ctx._chk(ctx.L.gcg_search_runs_packed(ctx.h, tab.h, pk_words.ctypes.data, pk_woff.ctypes.data, pk_lens.ctypes.data, len(arrs), K, C.byref(rp), C.byref(op), C.byref(nr), C.byref(na)))
        ctx.L.gcg_free(rp); ctx.L.gcg_free(op)
        s4 = tab.stats()
        tab.free()
        return na.value, nr.value, s4

    for _ in range(3):          # untimed warm-up calls (pinned blocks parked, workers awake), as for the device-resident step
        e2e_runs_packed()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rp_hits, rp_runs, rp_st = e2e_runs_packed()
    barrier()
    e2e_runs_packed_s = allmax(time.perf_counter() - t0) / e2e_steps
    assert rp_hits == n_hit and rp_st == st, "run-record (packed input) and anchor paths disagree"
    pk_keep.free()

    for _ in range(3):          # untimed warm-up calls (pinned blocks parked, workers awake), as for the device-resident step
        e2e_runs()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r_hits, r_runs, r_st = e2e_runs()
    barrier()
    e2e_runs_s = allmax(time.perf_counter() - t0) / e2e_steps
    assert r_hits == n_hit and r_st == st, "run-record and anchor paths disagree"

    sw_paths = list(swb.path_counts())
    swb.free()                                     # the resident batch holds up to 64 GB of trace scratch: give it back first
    sw_e2e_pairs = sw_pairs                         # the same batch as the device-resident leg, from host buffers
    # host buffers in page-locked memory (the contract's "inputs from pinned host memory"; gcg_host_alloc): no staging copy
    pin_q, pin_t = api.PinnedArray((sw_e2e_pairs * SW_QLEN,)), api.PinnedArray((sw_e2e_pairs * SW_TLEN,))
    pin_q.array[:] = q2[:sw_e2e_pairs].reshape(-1)
    pin_t.array[:] = t2[:sw_e2e_pairs].reshape(-1)
    e2e_sw_steps = 3
    barrier()
    res_e2e, cig_e2e = None, None
    for it in range(1 + e2e_sw_steps):                 # one untimed warm-up call of the same shape
        if it == 1:
            barrier()
            t0 = time.perf_counter()
        qb, qo = pin_q.array, np.arange(sw_e2e_pairs + 1, dtype=np.int64) * SW_QLEN
        tb, to = pin_t.array, np.arange(sw_e2e_pairs + 1, dtype=np.int64) * SW_TLEN
        res = np.zeros(sw_e2e_pairs, dtype=api.SWRES_DTYPE)
        pool, npool = C.c_void_p(), C.c_int64()
        ctx._chk(ctx.L.gcg_sw_batch(ctx.h, C.byref(P), api.SW_FIXED, qb.ctypes.data, qo.ctypes.data, tb.ctypes.data, to.ctypes.data,
                                    sw_e2e_pairs, res.ctypes.data, C.byref(pool), C.byref(npool)))
        n_ops = npool.value
        ctx.L.gcg_free(pool)
    barrier()
    e2e_sw_s = allmax(time.perf_counter() - t0) / e2e_sw_steps
    pin_q.free(); pin_t.free()

    # ---- aggregate over ranks
    tot_ont_kmers = allsum(float(n_ont_kmers))
    tot_cells = allsum(float(sw_cells))
    tot_e2e_cells = allsum(float(sw_e2e_pairs) * SW_QLEN * SW_TLEN)
    value = tot_ont_kmers / (kmer_ms * 1e-3)
    sw_value = tot_cells / (sw_ms * 1e-3) / 1e9

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak_gbs, peak_src = measured_peaks()
    # K4+K5 is ONE kernel now (probe + ordered anchor index + ONT-side multiplicity + anchor records): SURVEY 8d's
    # algorithmic bytes of the whole unit — 16.25 B per ONT k-mer looked up + 16 B per anchor (8-byte record + 8-byte
    # count RMW) — are charged to it
    kf = kprof.get("k45_fused", (0.0, 1))
    kf_avg = kf[0] / max(1, kf[1])
    kf_per_step = max(1.0, kf[1] / max(1, args.steps))
    alg_bytes = (BYTES_PER_LOOKUP * n_ont_kmers + BYTES_PER_HIT * n_hit) / kf_per_step
    achieved = alg_bytes / (kf_avg * 1e-3) / 1e9 if kf_avg > 0 else 0.0
    step_alg = BYTES_PER_LOOKUP * n_ont_kmers + BYTES_PER_HIT * n_hit + 32.25 * n_ctg_kmers

    def sw_leg(ms, prof, mode_name):
        fill = prof.get("k7_sw_fill_packed", (0.0, 1))
        g = sw_cells / (fill[0] / args.steps * 1e-3) / 1e9 if fill[0] > 0 else 0.0
        return fill, g
    fill, fill_gcups = sw_leg(sw_ms, sprof, "fixed")
    fill_a, fill_gcups_asis = sw_leg(sw_asis_ms, sprof_asis, "asis")
    sw_value_asis = tot_cells / (sw_asis_ms * 1e-3) / 1e9
    sw_peak = r_int16 * 2.0 / OPS_PER_CELL / 1e9          # 16-bit results per second / 12 ops per cell
    # DRAM bytes of one fill launch: the ncu capture is one launch over a known number of cfg3 pairs
    sw_traffic, cap_pairs = ncu_traffic("sw_fill_packed_kernel"), ncu_traffic("sw_fill_packed_kernel_pairs")
    if sw_traffic is not None and cap_pairs:
        sw_traffic = sw_traffic / cap_pairs * sw_pairs

    line = {
        "metric": "kmers_per_s", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": kmer_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "cfg2: 4.6 Mb synthetic scaffolds, 500 N-gaps, 30x ONT (10% error), k=25 — pack + contig table build + ONT search + ordered anchors + stats per step",
                   "k": K, "contigs": len(contigs), "reads_per_gpu": len(reads), "ont_kmers_per_gpu": n_ont_kmers, "contig_kmers": n_ctg_kmers,
                   "anchors_per_gpu": int(n_hit), "stats": list(st), "parallelism": "reads sharded by batch, table replicated (no data-path collective)",
                   "anchor_record": "8 bytes (pos << 36 | scaffold coordinate << 2 | flags) + one offset per read, include/gcgpu.h",
                   "l2": "inputs larger than L2 (138 MB ASCII reads + 106 MB table per step vs 126 MB L2); no explicit flush"},
        # reads cross PCIe 2-bit packed by the host gather (8 bytes per 32 bases, every read padded to whole words)
        # plus 12 bytes of layout per read and 4 per 1024 positions; contigs go up as ASCII
        "e2e": {"value": tot_ont_kmers / e2e_kmer_s, "unit": "k-mers/s",
                "h2d_bytes_per_step": int(sum((len(r) + 31) // 32 * 8 for r in reads) + 12 * len(reads) + read_bytes // 256 + ctg_bytes),
                "host_input_bytes_per_step": int(read_bytes + ctg_bytes),
                "d2h_bytes_per_step": int(n_hit * 8 + 8 * (len(reads) + 1) + 32), "ms_per_step": e2e_kmer_s * 1e3,
                "host_threads_per_rank": host_threads, "host_cores": os.cpu_count(),
                "api": "gcg_table_build + gcg_search_compact (host pointers in, pinned 8-byte anchors + per-read offsets out) + gcg_table_stats",
                "pipeline": "one search launch over reads that are still being gathered, packed and uploaded (GCG_SEARCH_STREAM=0: one launch per chunk)" if os.environ.get("GCG_SEARCH_STREAM", "1") != "0" else "one launch per chunk (GCG_SEARCH_STREAM=0)"},
        "e2e_packed": {"value": tot_ont_kmers / e2e_packed_s, "unit": "k-mers/s", "ms_per_step": e2e_packed_s * 1e3,
                       "h2d_bytes_per_step": int(sum((len(r) + 31) // 32 * 8 for r in reads) + 12 * len(reads) + read_bytes // 256 + ctg_bytes),
                       "d2h_bytes_per_step": int(n_hit * 8 + 8 * (len(reads) + 1) + 32),
                       "api": "gcg_table_build + gcg_search_compact_packed + gcg_table_stats: the reads as the B200 build's FASTQ loader leaves them (2-bit words in page-locked host memory, packed once on ingest — rseq_fast.c); no host pass over the bases inside the step"},
        "e2e_runs_packed": {"value": tot_ont_kmers / e2e_runs_packed_s, "unit": "k-mers/s", "ms_per_step": e2e_runs_packed_s * 1e3,
                            "h2d_bytes_per_step": int(sum((len(r) + 31) // 32 * 8 for r in reads) + 12 * len(reads) + read_bytes // 256 + ctg_bytes),
                            "d2h_bytes_per_step": int(rp_runs * 48 + 16 * (len(reads) + 1) + 32),
                            "api": "gcg_table_build + gcg_search_runs_packed + gcg_table_stats: what gc_b200 does with GC_RUNS=1 — the loader's 2-bit words up, run records down"},
        "e2e_runs": {"value": tot_ont_kmers / e2e_runs_s, "unit": "k-mers/s", "ms_per_step": e2e_runs_s * 1e3, "runs_per_gpu": int(r_runs),
                     "d2h_bytes_per_step": int(r_runs * 48 + 16 * (len(reads) + 1) + 32),
                     "api": "gcg_table_build + gcg_search_runs (N3, opt-in GC_RUNS mode of the shim: anchors reduced on the device to the run records of map_ont2contigs, ctg_graph.c:600-656) + gcg_table_stats"},
        "gpu_launches": int(kmer_launches + sw_launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k45_fused_kernel", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": ncu_traffic("k45_fused_kernel"), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": kf_avg,
                     "unit_of_work": "SURVEY K4+K5 as one kernel: probe, ordered anchor index (chained scan), ONT multiplicity, anchor records",
                     "whole_step": {"algorithmic_bytes": step_alg, "achieved": step_alg / (kmer_ms * 1e-3) / 1e9, "frac": step_alg / (kmer_ms * 1e-3) / 1e9 / peak_gbs,
                                    "note": "all kernels of the step (pack, build, search, stats) over the step time"},
                     "note": "the 106 MB cfg2 table is mostly L2 resident, so this is an algorithmic-bytes figure against the HBM copy peak, not DRAM traffic (`traffic`); the probe is bound by the L1 wavefront rate of one random 32-byte bucket load per k-mer (DESIGN.md 4); `roofline_hbm_table` is the same kernel on a table that does not fit the L2",
                     "secondary": {"bound": "l1_wavefront", "unit": "G wavefronts/s",
                                   "achieved": (n_ont_kmers + 3 * int(n_hit)) / kf_per_step / (kf_avg * 1e-3) / 1e9 if kf_avg > 0 else 0.0,
                                   "peak": torch.cuda.get_device_properties(local_rank).multi_processor_count * float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 0) * 1e-3,
                                   "note": "the bound that binds on an L2-sized table: a random 32-byte bucket load is one L1 wavefront per lane, "
                                           "one wavefront per clock and SM; counted: one per ONT k-mer + three per anchor (key bucket, value word, record)"},
                     "kernel_ms_per_step": {k_: v[0] / args.steps for k_, v in sorted(kprof.items())}},
        "sw": {"metric": "sw_gcups", "value": sw_value, "unit": "GCUPS", "ms_per_step": sw_ms, "dtype": "s16x2",
               "config": {"workload": "cfg3 shape: ONT-read(10 kb) x gap-flank(2 kb) pairs, default scoring (+1/-5/2/1, softclip), FIXED traceback (cell re-fetched every step, SURVEY H4); fill + trace spill + end cell + traceback walk + CIGAR",
                          "pairs_per_gpu_per_step": int(sw_pairs), "cells_per_gpu_per_step": int(sw_cells), "paths": sw_paths,
                          "l2": "10.4 MB of trace per pair streams through L2 (%.0f GB per step; every resident warp reuses one 20.8 MB slot)" % (sw_pairs * 10.4e-3)},
               "e2e": {"value": tot_e2e_cells / e2e_sw_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(sw_e2e_pairs * (SW_QLEN + SW_TLEN)),
                       "d2h_bytes_per_step": int(sw_e2e_pairs * 32 + n_ops * 4), "pairs_per_step": int(sw_e2e_pairs), "api": "gcg_sw_batch (pairs in page-locked host buffers from gcg_host_alloc), FIXED traceback"},
               "asis": {"value": sw_value_asis, "unit": "GCUPS", "ms_per_step": sw_asis_ms, "roofline_frac": fill_gcups_asis / sw_peak if sw_peak else None,
                        "note": "traceback exactly as shipped (sw.c:289-319 never re-fetches the cell): the CIGAR is a function of the end cell only"},
               "roofline": {"bound": "int_alu", "kernel": "sw_fill_packed_kernel", "achieved": fill_gcups, "peak": sw_peak, "unit": "GCUPS",
                            "frac": fill_gcups / sw_peak if sw_peak else None, "traffic": sw_traffic,
                            "peak_source": "measured in this run: VIADDMNMX.S16x2 issue rate %.2f T lane-ops/s x 2 halves / 12 integer ops per cell (SURVEY 8d)" % (r_int16 / 1e12),
                            "secondary": {"bound": "hbm", "achieved": 0.5 * sw_cells / (fill[0] / args.steps * 1e-3) / 1e9 if fill[0] > 0 else 0.0,
                                          "peak": peak_gbs, "unit": "GB/s", "note": "trace spill, 0.5 byte per cell"},
                            "kernel_ms_per_step": {k_: v[0] / args.steps for k_, v in sorted(sprof.items())}}},
    }
    sec_ = line["roofline"]["secondary"]
    sec_["frac"] = sec_["achieved"] / sec_["peak"] if sec_["peak"] else None
    if cfg4 is not None:
        for x4, leg in cfg4.items():
            line["partitioned_cfg4" if x4 == args.cfg4_exchange else "partitioned_cfg4_" + x4] = leg
    if hbm_leg is not None:
        hbm_leg.update({"bound": "hbm", "peak": peak_gbs, "unit": "GB/s", "frac": hbm_leg["achieved"] / peak_gbs, "peak_source": peak_src})
        line["roofline_hbm_table"] = hbm_leg

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference itself on the host cores
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        try:
            if reference_available():
                with tempfile.TemporaryDirectory() as tmp:
                    v, info = cpu_kmer(inp.scaffold, reads, cores, tmp)
                line["cpu_baseline"] = {"value": v, "unit": "k-mers/s", "cores": cores, "kind": "reference",
                                        "sample": "the full cfg2 workload once (%d ONT k-mers); reference chop %.2fs + hash %.2fs + search/rehash %.2fs on %d threads" %
                                                  (info["n_ont_kmers"], info["t_chop"], info["t_put"], info["t_search"], cores)}
                assert [info["scaf_total"], info["scaf_unique"], info["ont_total"], info["ont_unique"]] == list(st), "GPU and reference statistics differ"
                assert info["n_hits"] == n_hit, "GPU and reference anchor counts differ"
                g, npairs, scores = cpu_sw(cores, 2, seed=46)
                line["sw"]["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": cores, "kind": "reference",
                                              "sample": "%d pairs of 10 kb x 2 kb, one reference sw_t per core" % npairs}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "k-mers/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref not built"}
        except Exception as e:      # the baseline must never sink the bench line
            line["cpu_baseline"] = {"value": None, "unit": "k-mers/s", "cores": cores, "kind": "reference", "sample": "failed: %r" % (e,)}
    for exchange, pr in part.items():
        line[{"direct": "partitioned", "remote": "partitioned_remote", "all_to_all": "partitioned_all_to_all"}[exchange]] = {
            "metric": "kmers_per_s", "value": tot_ont_kmers / (pr["ms_per_step"] * 1e-3), "unit": "k-mers/s", "ms_per_step": pr["ms_per_step"],
            "config": {"workload": "same step with the contig table hash-partitioned over %d GPU(s)" % world,
                       "exchange": ({"remote": "remote probes: the partitions stay where their owners built them (CUDA-IPC mapped blocks) and ONE search kernel per rank reads them where they lie — bucket loads and the anchors' atomicOr over NVLink peer memory, after a replicated union filter; no routing, no exchange buffers, no collective on the data path",
                                     "direct": "route by owner, owner-side lookup, ordered collect: keys (8 B per ONT k-mer) stored by the routing kernel straight into the owner's window and answers (8 B) by the lookup kernel straight into the requester's window over NVLink peer memory (CUDA IPC); two stream-ordered barriers per round",
                                     "all_to_all": "route by owner, owner-side lookup, ordered collect through send / receive buffers and NCCL all_to_all_single (8 B key out, 8 B answer back per ONT k-mer)"}[exchange]) if world > 1 else "single partition, nothing crosses a link"},
            "nvlink_bytes_per_step": pr["bytes_sent_per_step"], "stats": pr["stats"], "kernel_ms_per_step": pr["kernel_ms_per_step"],
            "prefilter": "union of the partitions' anchoring keys (blocked Bloom filter, all-gathered after the build)" + ("" if exchange == "remote" else ": %.1f %% of the ONT k-mers are routed" % (100.0 * pr["routed_fraction"]))}
        line["gpu_launches"] += int(pr["launches"])
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sw-pairs", type=int, default=47360, help="pairs per GPU per step of the cfg3 leg (16 items per resident warp)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cfg4", default="auto", choices=["auto", "on", "off"], help="BASELINE configs[3] at full size on the partitioned table (auto: at 8 GPUs)")
    ap.add_argument("--cfg4-exchange", default="direct", choices=["remote", "direct", "all_to_all"],
                    help="exchange of the full-size cfg4 leg (remote probes collapse at 8 GPUs x 450 MB partitions: 1.04 s against 10.2 ms per search, profiles/r02_bench_n8_remote_ab.json)")
    ap.add_argument("--cfg4-genome", type=int, default=0, help="genome length of that leg (default: 100 Mb, the config's)")
    ap.add_argument("--no-hbm-table", action="store_true", help="skip the second k-mer roofline leg (100 Mb table beyond the L2, N=1 only)")
    ap.add_argument("--hbm-coverage", type=float, default=2.0, help="ONT coverage of the 100 Mb genome in that leg")
    ap.add_argument("--partitioned", action="store_true", help="also time the hash-partitioned table at N=1 (always timed at N>1)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                      # timing rule: at least three warm-up steps
    rank, local_rank, world = dist_env()
    # stdout carries the one JSON line and nothing else: native libraries that write to file
    # descriptor 1 (NCCL prints its version banner there) are pointed at stderr, Python's own
    # sys.stdout keeps the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
