"""Multi-GPU driver of the k-mer path: hash-partitioned contig table with an all-to-all exchange.

SURVEY §8e / BASELINE configs[3].  The reference already partitions its tables — table
`crc32(kseq) % n_thread` per thread (kmer.c:88,124-152; ont.c:169,193) — and lets every thread
scan all k-mers for its share.  Here a partition is one GPU (one process per GPU):

  build  : rank r chops its slice of the contig tiles, routes {k-mer, tid, pos, strand} records to
           their owners (all-to-all, 16 B per contig k-mer), the owner inserts them.
  search : rank r chops its own read batch, routes the canonical k-mers to their owners
           (all-to-all, 8 B per ONT k-mer), the owner answers (value word of a k-mer present
           exactly once, ont.c:171,195 — or a miss) and counts the ONT-side multiplicity
           (ont.c:245); answers return by the reverse all-to-all (8 B) and are turned into
           anchors in (read,pos) order.
  stats  : all-reduce of the four counters of kmer.c:265-312.

The device work is libgcgpu's (`gcg_route_*`, `gcg_table_insert_records`, `gcg_table_lookup_keys`
through `DeviceOps`); this module is only the orchestration and the collectives
(`torch.distributed`: NCCL over NVLink on the GPUs, gloo in the CPU tests of the host logic).
`ops` and `comm` are injected so that the same orchestration runs
  * one process per GPU under torchrun                     (`DeviceOps` + `TorchComm`, the product),
  * several partitions in one process on one GPU           (`DeviceOps` + `ThreadComm`, GPU tests),
  * the CPU tests of the exchange logic                    (a numpy double from tests/ + gloo).
There is no CPU implementation of the device work in this package.
"""
from __future__ import annotations

import contextlib
import os
import threading
import time

import numpy as np
import torch

from . import api

MISS = np.uint64(0xFFFFFFFFFFFFFFFF)
ROUND_KMERS = 1 << 31          # k-mer positions routed per exchange round (the plan's offsets are 32 bit)


# ------------------------------------------------------------------------------------------------
# collectives
# ------------------------------------------------------------------------------------------------
class TorchComm:
    """torch.distributed default group: NCCL with CUDA tensors, gloo with CPU tensors."""

    def __init__(self, device: torch.device):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device
        self.bytes_sent = 0           # payload handed to all_to_all by this rank, excluding the self segment

    def empty(self, n_words: int) -> torch.Tensor:
        return torch.empty(max(int(n_words), 1), dtype=torch.int64, device=self.device)

    def exchange_counts(self, counts: np.ndarray) -> np.ndarray:
        """counts[p] = elements this rank sends to p  ->  elements this rank receives from p"""
        send = torch.as_tensor(np.ascontiguousarray(counts, dtype=np.int64)).to(self.device)
        recv = torch.empty_like(send)
        self.dist.all_to_all_single(recv, send)
        return recv.cpu().numpy()

    def all_to_all(self, send: torch.Tensor, send_counts, recv_counts, width: int) -> torch.Tensor:
        """elements are `width` int64 words; segment p of `send` goes to rank p"""
        ssz = [int(c) * width for c in send_counts]
        rsz = [int(c) * width for c in recv_counts]
        recv = self.empty(sum(rsz))
        self.dist.all_to_all_single(recv[: sum(rsz)], send[: sum(ssz)], rsz, ssz)
        self.bytes_sent += 8 * (sum(ssz) - ssz[self.rank])
        return recv

    def all_reduce(self, values, op: str = "sum") -> np.ndarray:
        t = torch.as_tensor(np.asarray(values, dtype=np.int64)).to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MAX)
        return t.cpu().numpy()

    def gather_counts(self, counts: np.ndarray) -> np.ndarray:
        """-> M[world, world], M[r][d] = elements rank r sends to rank d"""
        send = torch.as_tensor(np.ascontiguousarray(counts, dtype=np.int64)).to(self.device)
        out = torch.empty(self.world * self.world, dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(out, send)
        return out.cpu().numpy().reshape(self.world, self.world)

    def barrier(self):
        self.dist.barrier()

    def gather_ints(self, values) -> np.ndarray:
        """-> [world, len(values)]: every rank's small integer vector"""
        send = torch.as_tensor(np.asarray(values, dtype=np.int64)).to(self.device)
        out = torch.empty(self.world * send.numel(), dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(out, send)
        return out.cpu().numpy().reshape(self.world, -1)

    def stream_barrier(self):
        """a barrier in STREAM order, without a host round trip: a one-element all-reduce enqueued on the current
        stream (the library's, see DeviceOps.stream).  Kernels queued behind it on any rank start only after every
        rank's kernels queued before it have finished — a kernel's stores into a peer's window are complete when
        the kernel is, and a rank joins the all-reduce only then."""
        if getattr(self, "_flag", None) is None:
            self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.dist.all_reduce(self._flag)

    def all_gather(self, t: torch.Tensor):
        """-> list of every rank's tensor (same shape everywhere)"""
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t)
        return list(out.view(self.world, -1))

    def share(self, ops, window):
        """every rank's window as a reference this rank's kernels can store through (CUDA IPC)"""
        if self.device.type != "cuda":
            raise RuntimeError("direct exchange needs device windows (NVLink peer memory); use the all-to-all exchange on CPU")
        handles = [None] * self.world
        self.dist.all_gather_object(handles, ops.window_export(window))
        return [ops.window_ref(window) if r == self.rank else ops.window_open(handles[r]) for r in range(self.world)]

    def unshare(self, ops, refs):
        for r, ref in enumerate(refs):
            if r != self.rank:
                ops.window_close(ref)

    def share_block(self, ops, ptr: int, handle: bytes):
        """every rank's cudaMalloc block (a shared table partition) as an address this rank's kernels can use"""
        handles = [None] * self.world
        self.dist.all_gather_object(handles, handle)
        return [ptr if r == self.rank else ops.window_open(handles[r]) for r in range(self.world)]


class ThreadGroup:
    def __init__(self, world: int):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world


class ThreadComm:
    """`world` partitions driven by `world` threads of one process (GPU tests on a single device;
    also the shape a single-process multi-GPU host would use).  `sync` must drain the caller's
    stream: tensors are handed to other threads' streams."""

    def __init__(self, group: ThreadGroup, rank: int, device: torch.device, sync=lambda: None):
        self.g, self.rank, self.world, self.device, self.sync = group, rank, group.world, device, sync
        self.bytes_sent = 0

    def empty(self, n_words: int) -> torch.Tensor:
        return torch.empty(max(int(n_words), 1), dtype=torch.int64, device=self.device)

    def _publish(self, obj):
        self.sync()
        self.g.slots[self.rank] = obj
        self.g.barrier.wait()
        return list(self.g.slots)

    def _done(self):
        self.sync()
        self.g.barrier.wait()

    def stream_barrier(self):
        self._done()

    def gather_ints(self, values) -> np.ndarray:
        allv = self._publish(np.asarray(values, dtype=np.int64))
        out = np.stack(allv)
        self._done()
        return out

    def exchange_counts(self, counts: np.ndarray) -> np.ndarray:
        allc = self._publish(np.array(counts, dtype=np.int64))
        out = np.array([allc[p][self.rank] for p in range(self.world)], dtype=np.int64)
        self._done()
        return out

    def all_to_all(self, send: torch.Tensor, send_counts, recv_counts, width: int) -> torch.Tensor:
        pubs = self._publish((send, [int(c) for c in send_counts]))
        pieces = []
        for p in range(self.world):
            t, cnt = pubs[p]
            off = sum(cnt[: self.rank]) * width
            assert cnt[self.rank] == int(recv_counts[p])
            pieces.append(t[off: off + cnt[self.rank] * width])
        recv = self.empty(sum(int(c) for c in recv_counts) * width)
        if pieces:
            cat = torch.cat(pieces)
            recv[: cat.numel()] = cat
        self.bytes_sent += 8 * width * (sum(int(c) for c in send_counts) - int(send_counts[self.rank]))
        self._done()
        return recv

    def all_reduce(self, values, op: str = "sum") -> np.ndarray:
        allv = self._publish(np.asarray(values, dtype=np.int64))
        out = np.sum(allv, axis=0) if op == "sum" else np.max(allv, axis=0)
        self._done()
        return out

    def gather_counts(self, counts: np.ndarray) -> np.ndarray:
        allc = self._publish(np.array(counts, dtype=np.int64))
        out = np.stack(allc)
        self._done()
        return out

    def barrier(self):
        self.sync()
        self.g.barrier.wait()

    def all_gather(self, t: torch.Tensor):
        out = [x.clone() if i != self.rank else x for i, x in enumerate(self._publish(t))]
        self._done()
        return out

    def share(self, ops, window):
        refs = self._publish(ops.window_ref(window))      # one address space: the reference is the window itself
        self._done()
        return refs

    def unshare(self, ops, refs):
        pass

    def share_block(self, ops, ptr: int, handle: bytes):
        refs = self._publish(ptr)                          # one address space
        self._done()
        return refs


# ------------------------------------------------------------------------------------------------
# device work (the C ABI)
# ------------------------------------------------------------------------------------------------
class DeviceOps:
    """libgcgpu through api.Context; tensors are CUDA int64 tensors whose data_ptr is handed to
    the library.  All torch work is issued on the library's stream so that kernels, allocator and
    collectives are ordered without extra events."""

    def __init__(self, ctx: api.Context, device_index: int):
        self.ctx = ctx
        self.device = torch.device("cuda", device_index)
        self._stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=self.device)

    def stream(self):
        return torch.cuda.stream(self._stream)

    def sync(self):
        self.ctx.sync()

    def tiles(self, seqs) -> int:
        return seqs.tiles

    def plan(self, seqs, k, n_part, t0, t1, prefilter=None):
        if prefilter is not None:
            t, n_words, k3 = prefilter
            prefilter = (t.data_ptr(), n_words, k3)
        return self.ctx.route_plan(seqs, k, n_part, t0, t1, prefilter)

    # pre-filter of the routed search: an int32 tensor of filter words
    def filter_new(self, n_keys: int):
        n_words, k3 = self.ctx.filter_shape(n_keys)
        return torch.zeros(n_words, dtype=torch.int32, device=self.device), n_words, k3

    def filter_add_table(self, table, flt):
        t, n_words, k3 = flt
        table.filter_add(t.data_ptr(), n_words, k3)

    def filter_or(self, flt, other: torch.Tensor):
        t, n_words, _ = flt
        self.ctx.filter_or(t.data_ptr(), other.data_ptr(), n_words)

    def route_keys(self, route, t: torch.Tensor):
        route.keys(t.data_ptr())

    def route_records(self, route, t: torch.Tensor):
        route.records(t.data_ptr())

    def table_create(self, n, k):
        return self.ctx.table_create(n, k)

    def table_create_shared(self, n, k):
        return self.ctx.table_create_shared(n, k)

    def search_remote(self, reads, k, key_ptrs, val_ptrs, n_buckets, flt):
        t, n_words, k3 = flt
        return self.ctx.search_remote(reads, k, key_ptrs, val_ptrs, n_buckets, t.data_ptr(), n_words, k3)

    def insert(self, table, t: torch.Tensor, n: int):
        table.insert_records(t.data_ptr(), n)

    def lookup(self, table, keys: torch.Tensor, n: int, answers: torch.Tensor):
        table.lookup_keys(keys.data_ptr(), n, answers.data_ptr())

    def collect(self, route, answers: torch.Tensor):
        return route.collect(answers.data_ptr())

    # direct exchange: windows other ranks store into over NVLink
    def window(self, n_words: int):
        return self.ctx.window(8 * max(int(n_words), 1))

    def window_ref(self, window) -> int:
        return window.ptr

    def window_export(self, window) -> bytes:
        return window.export()

    def window_open(self, handle: bytes) -> int:
        return self.ctx.window_open(handle)

    def window_close(self, ref: int):
        self.ctx.window_close(ref)

    def route_keys_direct(self, route, owner_refs, owner_off):
        route.keys_direct(owner_refs, owner_off)

    def lookup_direct(self, table, key_window, src_count, answer_refs, answer_off):
        table.lookup_keys_direct(key_window.ptr, src_count, answer_refs, answer_off)

    def collect_window(self, route, answer_window):
        return route.collect(answer_window.ptr)

    def stats(self, table):
        return table.stats()


# ------------------------------------------------------------------------------------------------
# the partitioned index
# ------------------------------------------------------------------------------------------------
def tile_slice(n_tiles: int, rank: int, world: int):
    """contiguous share of [0, n_tiles) owned by `rank` when `world` ranks split the chop work"""
    return n_tiles * rank // world, n_tiles * (rank + 1) // world


def search_rounds(n_tiles: int, round_kmers: int = ROUND_KMERS):
    """tile ranges of one rank's read batch, each holding fewer than `round_kmers` positions"""
    per = max(1, round_kmers // 1024 - 1)
    return [(a, min(n_tiles, a + per)) for a in range(0, n_tiles, per)] or [(0, 0)]


class PartitionedKmerIndex:
    """contig k-mer table partitioned over comm.world GPUs by hash of the canonical k-mer"""

    def __init__(self, ops, comm, k: int, round_kmers: int = ROUND_KMERS, exchange: str = "all_to_all", prefilter: bool = True):
        """exchange = "all_to_all": keys and answers travel through send / receive buffers and
                       `all_to_all_single` (NCCL; gloo in the CPU tests);
           exchange = "direct": the routing and lookup kernels store straight into the peers' windows
                       over NVLink / NVSwitch peer memory, the ranks only meet at two barriers per round.
           exchange = "remote": no routing at all — the partitions stay where their owners built them (one cudaMalloc
                       block each, mapped by every rank) and ONE search kernel per rank probes them where they lie:
                       a k-mer that passes the union filter loads its bucket from the owner's key array over NVLink,
                       an anchor's atomicOr travels to the owner's value array and returns (tid, pos, flag).
           prefilter: after the build every rank adds its anchoring keys (present exactly once) to a
                       Bloom-style filter, the partial filters are all-gathered and OR-ed, and the
                       search routes only the ONT k-mers that pass it — the anchors plus a few
                       percent of false positives instead of all of them (96 % are sequencing-error
                       k-mers that no table holds), which shrinks the exchange and the owner-side
                       lookups by an order of magnitude."""
        if not 1 <= comm.world <= api.MAX_PART:
            raise ValueError("world size %d outside [1,%d]" % (comm.world, api.MAX_PART))
        if exchange not in ("all_to_all", "direct", "remote"):
            raise ValueError("exchange must be 'all_to_all', 'direct' or 'remote'")
        if exchange == "remote" and not prefilter:
            raise ValueError("the remote-probe search needs the union pre-filter")
        self.ops, self.comm, self.k, self.round_kmers, self.exchange = ops, comm, k, round_kmers, exchange
        self.table = None
        self.n_local_records = 0
        self.use_prefilter, self.prefilter = prefilter, None
        self.n_routed = self.n_positions = 0      # of the searches so far: k-mers exchanged / k-mer positions
        self.qwin = self.awin = None              # direct exchange: my key / answer windows ...
        self.q_refs = self.a_refs = None          # ... and every rank's, as seen from this rank
        self.q_cap = self.a_cap = 0
        self.block_refs = None                    # remote probes: every rank's partition block as this rank addresses it
        self.part_shape = None                    # ... and (value offset, bucket count) of every partition
        # GCG_DIST_PROFILE=1: wall-clock per phase with a device sync on both sides (diagnosis only;
        # the syncs serialise what normally overlaps)
        self.profile = os.environ.get("GCG_DIST_PROFILE") == "1"
        self.timers = {}

    @contextlib.contextmanager
    def _phase(self, name):
        if not self.profile:
            yield
            return
        self.ops.sync()
        t0 = time.perf_counter()
        yield
        self.ops.sync()
        self.timers[name] = self.timers.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)

    def _stream(self):
        return self.ops.stream() if hasattr(self.ops, "stream") else contextlib.nullcontext()

    # ---- build: kmer.c:155-213 across ranks ------------------------------------------------------
    def build(self, contigs) -> "PartitionedKmerIndex":
        """`contigs`: the full contig set, uploaded on every rank (2 bits per base; the *table* is
        what is partitioned).  Each rank chops 1/world of the tiles."""
        ops, comm = self.ops, self.comm
        reuse = False
        if self.table is not None:              # rebuilding: the exchange windows of the direct mode are kept
            if self.exchange != "remote":
                self.table.free()
                self.table = None
        with self._stream():
            t0, t1 = tile_slice(ops.tiles(contigs), comm.rank, comm.world)
            with self._phase("build.plan"):
                route = ops.plan(contigs, self.k, comm.world, t0, t1)
            with self._phase("build.counts"):
                recv_counts = comm.exchange_counts(route.counts)
            with self._phase("build.route"):
                send = comm.empty(2 * int(route.counts.sum()))
                ops.route_records(route, send)
            with self._phase("build.all_to_all"):
                recv = comm.all_to_all(send, route.counts, recv_counts, width=2)
            n = int(recv_counts.sum())
            with self._phase("build.insert"):
                if self.exchange == "remote":
                    # the partition lives in a block the peers have mapped: keep the block (and the mappings) when the
                    # rebuild fits it.  Nobody may still be probing the old contents: the search ends with a barrier.
                    reuse = self.table is not None and self.table.reset_shared(n)
                    if not reuse:
                        self._drop_blocks()
                        self.table = ops.table_create_shared(n, self.k)
                else:
                    self.table = ops.table_create(n, self.k)
                ops.insert(self.table, recv, n)
                ops.sync()                      # recv/send are released after the inserts ran
            route.free()
            self.n_local_records = n
            self.prefilter = None
            if self.use_prefilter:
                with self._phase("build.prefilter"):
                    n_total = int(comm.all_reduce([n], "sum")[0])
                    flt = ops.filter_new(n_total)
                    ops.filter_add_table(self.table, flt)
                    ops.sync()
                    for r, part in enumerate(comm.all_gather(flt[0])):
                        if r != comm.rank:
                            ops.filter_or(flt, part)
                    ops.sync()
                    self.prefilter = flt
            if self.exchange == "remote":
                with self._phase("build.share"):
                    ptr, vals_off, nb, handle = self.table.shared_info()
                    shapes = comm.gather_ints([vals_off, nb, 0 if reuse else 1])
                    if self.block_refs is None or shapes[:, 2].any():
                        # some rank has a new block: everybody maps everybody again (rare: the first build, or a larger table)
                        if self.block_refs is not None:
                            self._close_block_refs()
                        self.block_refs = comm.share_block(ops, ptr, handle)
                    self.part_shape = [(int(shapes[r, 0]), int(shapes[r, 1])) for r in range(comm.world)]
        return self

    def _close_block_refs(self):
        for r, ref in enumerate(self.block_refs):
            if r != self.comm.rank and hasattr(self.comm, "dist"):
                self.ops.window_close(ref)
        self.block_refs = None

    def _drop_blocks(self):
        """collective: nobody keeps a mapping of a partition block that is about to be freed"""
        if self.table is None:
            return
        self.ops.sync()
        self.comm.barrier()
        if self.block_refs is not None:
            self._close_block_refs()
        self.comm.barrier()
        self.table.free()
        self.table = None

    # ---- direct exchange ------------------------------------------------------------------------
    def _drop_windows(self):
        if self.qwin is None:
            return
        comm, ops = self.comm, self.ops
        ops.sync()
        comm.barrier()                            # nobody stores into a window that is about to go
        comm.unshare(ops, self.q_refs)
        comm.unshare(ops, self.a_refs)
        comm.barrier()                            # every mapping is closed before the memory is freed
        self.qwin.free(); self.awin.free()
        self.qwin = self.awin = self.q_refs = self.a_refs = None
        self.q_cap = self.a_cap = 0

    def _ensure_windows(self, need_q: int, need_a: int):
        """need_* are maxima over ALL ranks (computed from the gathered counts), so every rank takes
        the same branch and the (re)creation is collective"""
        if self.qwin is not None and need_q <= self.q_cap and need_a <= self.a_cap:
            return
        self._drop_windows()
        self.q_cap, self.a_cap = need_q + need_q // 4 + 1024, need_a + need_a // 4 + 1024
        self.qwin, self.awin = self.ops.window(self.q_cap), self.ops.window(self.a_cap)
        self.q_refs = self.comm.share(self.ops, self.qwin)
        self.a_refs = self.comm.share(self.ops, self.awin)

    def _round_direct(self, reads, t0, t1):
        ops, comm, me, world = self.ops, self.comm, self.comm.rank, self.comm.world
        with self._phase("search.plan"):
            route = ops.plan(reads, self.k, world, t0, t1, self.prefilter)
            self.n_routed += route.kmers; self.n_positions += route.positions
        with self._phase("search.counts"):
            M = comm.gather_counts(route.counts)            # M[r][d]: keys of rank r owned by rank d
            self._ensure_windows(int(M.sum(axis=0).max()), int(M.sum(axis=1).max()))
        with self._phase("search.route+store"):
            # my run inside owner d's key window starts behind the runs of the lower ranks
            ops.route_keys_direct(route, self.q_refs, [int(M[:me, d].sum()) for d in range(world)])
            comm.stream_barrier()                           # every key window is complete (stream order: the host does not wait)
        with self._phase("search.lookup+store"):
            # the answers of requester r go where r's collect pass expects owner `me`: behind the lower owners
            ops.lookup_direct(self.table, self.qwin, M[:, me], self.a_refs, [int(M[r, :me].sum()) for r in range(world)])
            comm.stream_barrier()                           # every answer window is complete
        with self._phase("search.collect"):
            h = ops.collect_window(route, self.awin)        # (returns with the anchor count: the round's one wait besides the counts)
        comm.bytes_sent += 8 * (int(M[me].sum()) - int(M[me, me])) + 8 * (int(M[:, me].sum()) - int(M[me, me]))
        route.free()
        return h

    # ---- search: ont.c:141-254 across ranks ------------------------------------------------------
    def search(self, reads, keep_on_device: bool = False):
        """anchors of this rank's read batch in (read,pos) order (read = index in `reads`).
        Every rank must call it (the exchange is collective), with its own batch."""
        ops, comm = self.ops, self.comm
        parts, n_total = [], 0
        with self._stream():
            if self.exchange == "remote":
                with self._phase("search.remote"):
                    comm.stream_barrier()               # every partition (and the union filter) is built before anybody probes it
                    keys = [self.block_refs[r] for r in range(comm.world)]
                    vals = [self.block_refs[r] + self.part_shape[r][0] for r in range(comm.world)]
                    h = ops.search_remote(reads, self.k, keys, vals, [self.part_shape[r][1] for r in range(comm.world)], self.prefilter)
                    comm.stream_barrier()               # every rank has finished probing before statistics are read or a partition is rebuilt
                self.n_positions += reads.kmers(self.k) if hasattr(reads, "kmers") else 0
                if keep_on_device:
                    n = h.n
                    h.free()
                    return n
                out = h.download()
                h.free()
                return out
            rounds = search_rounds(ops.tiles(reads), self.round_kmers)
            n_rounds = int(comm.all_reduce([len(rounds)], "max")[0])
            for i in range(n_rounds):
                t0, t1 = rounds[i] if i < len(rounds) else (0, 0)
                if self.exchange == "direct":
                    h = self._round_direct(reads, t0, t1)
                    n_total += h.n
                    if not keep_on_device:
                        parts.append(h.download())
                    h.free()
                    continue
                with self._phase("search.plan"):
                    route = ops.plan(reads, self.k, comm.world, t0, t1, self.prefilter)
                    self.n_routed += route.kmers; self.n_positions += route.positions
                with self._phase("search.counts"):
                    recv_counts = comm.exchange_counts(route.counts)
                n_send, n_recv = int(route.counts.sum()), int(recv_counts.sum())
                with self._phase("search.route"):
                    keys = comm.empty(n_send)
                    ops.route_keys(route, keys)
                with self._phase("search.keys_all_to_all"):
                    q = comm.all_to_all(keys, route.counts, recv_counts, width=1)
                with self._phase("search.lookup"):
                    a = comm.empty(n_recv)
                    ops.lookup(self.table, q, n_recv, a)
                with self._phase("search.answers_all_to_all"):
                    answers = comm.all_to_all(a, recv_counts, route.counts, width=1)
                with self._phase("search.collect"):
                    h = ops.collect(route, answers)
                    ops.sync()
                route.free()
                n_total += h.n
                if keep_on_device:
                    h.free()
                else:
                    parts.append(h.download())
                    h.free()
        if keep_on_device:
            return n_total
        return np.concatenate(parts) if parts else np.zeros(0, dtype=api.HIT_DTYPE)

    # ---- stats: kmer.c:265-312 -------------------------------------------------------------------
    def stats(self):
        with self._stream():
            return tuple(int(x) for x in self.comm.all_reduce(self.ops.stats(self.table), "sum"))

    def free(self):
        """collective when the direct exchange was used (the windows are unmapped everywhere first)"""
        with self._stream():
            self._drop_windows()
            if self.exchange == "remote":
                self._drop_blocks()
        if self.table is not None:
            self.table.free()
            self.table = None


def run_threaded(world: int, fn, device: torch.device, make_ops):
    """run fn(rank, ops, comm) on `world` threads sharing one process (ThreadComm); returns the
    list of results.  Exceptions are re-raised in the caller."""
    group = ThreadGroup(world)
    out, err = [None] * world, [None] * world

    def body(r):
        try:
            ops = make_ops(r)
            comm = ThreadComm(group, r, device, sync=getattr(ops, "sync", lambda: None))
            out[r] = fn(r, ops, comm)
        except BaseException as e:          # noqa: BLE001 - reported below
            err[r] = e
            group.barrier.abort()

    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in err:
        if e is not None:
            raise e
    return out
