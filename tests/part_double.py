"""Test double for the device half of superplus_b200.dist (TEST INFRASTRUCTURE ONLY).

`NumpyOps` has the interface of dist.DeviceOps but does the partition arithmetic with numpy on
top of the CPU oracle (oracle/gc_oracle.c via oracle.oracle.Oracle).  It exists so that the
orchestration and the collectives of dist.PartitionedKmerIndex can be exercised on CPU with gloo
(world_size 2) and with threads; it is also the checker of the CUDA partition kernels in the
`-m gpu` tests.  It is never imported by the package."""
import numpy as np

from superplus_b200 import api

MISS = np.uint64(0xFFFFFFFFFFFFFFFF)


def owner_np(keys: np.ndarray, n_part: int) -> np.ndarray:
    """same mix as kmer_owner in superplus_b200/csrc/part.cu"""
    with np.errstate(over="ignore"):
        x = keys.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(29)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(32)
        hi = x >> np.uint64(32)
        return ((hi * np.uint64(n_part)) >> np.uint64(32)).astype(np.int64)


class HostSeqs:
    """word-aligned layout of a sequence set: 32 bases per word, 32 words per tile"""

    def __init__(self, seqs):
        self.seqs = [np.asarray(s, dtype=np.uint8) for s in seqs]
        lens = np.array([len(s) for s in self.seqs], dtype=np.int64)
        self.woff = np.zeros(len(self.seqs) + 1, dtype=np.int64)
        np.cumsum((lens + 31) // 32, out=self.woff[1:])
        self.n_words = int(self.woff[-1])

    @property
    def tiles(self):
        return (self.n_words + 31) // 32

    def free(self):
        pass


class _Route:
    def __init__(self, counts, keys, vals, owner, read, pos, orev, positions=None):
        self.counts, self.keys, self.vals, self.owner, self.read, self.pos, self.orev = counts, keys, vals, owner, read, pos, orev
        self.order = np.argsort(owner, kind="stable")
        self.positions = len(keys) if positions is None else positions

    @property
    def kmers(self):
        return len(self.keys)

    def free(self):
        pass


class _FreeableArray(np.ndarray):
    def free(self):
        pass


class _Hits:
    def __init__(self, arr):
        self.arr = arr

    @property
    def n(self):
        return len(self.arr)

    def download(self):
        return self.arr

    def free(self):
        pass


class _Table:
    def __init__(self, k):
        self.k, self.multi, self.val, self.ont = k, {}, {}, {}

    def free(self):
        pass


class NumpyOps:
    def __init__(self, oracle):
        self.oracle = oracle

    def sync(self):
        pass

    def tiles(self, seqs: HostSeqs) -> int:
        return seqs.tiles

    # pre-filter: the double keeps its own small Bloom filter (any deterministic hash will do: it is only
    # ever tested by the double itself); same shape on every rank so that the partial filters can be OR-ed
    FILTER_WORDS = 1 << 12

    @staticmethod
    def _filter_bits(keys):
        with np.errstate(over="ignore"):
            h = keys.astype(np.uint64) * np.uint64(0xD6E8FEB86659FD93)
        return ((h >> np.uint64(40)) % np.uint64(NumpyOps.FILTER_WORDS)).astype(np.int64), (np.uint32(1) << ((h >> np.uint64(33)) & np.uint64(31)).astype(np.uint32))

    def filter_new(self, n_keys):
        import torch
        return torch.zeros(self.FILTER_WORDS, dtype=torch.int32), self.FILTER_WORDS, 0

    def filter_add_table(self, table, flt):
        w = flt[0].numpy().view(np.uint32)
        uniq = np.array([key - 1 for key, m in table.multi.items() if m == 1], dtype=np.uint64)
        if len(uniq):
            idx, bit = self._filter_bits(uniq)
            np.bitwise_or.at(w, idx, bit)

    def filter_or(self, flt, other):
        w = flt[0].numpy().view(np.uint32)
        w |= other.numpy().view(np.uint32)

    def plan(self, seqs: HostSeqs, k, n_part, t0, t1, prefilter=None):
        K, V, R, P, O = [], [], [], [], []
        for i, s in enumerate(seqs.seqs):
            if len(s) < k:
                continue
            w = seqs.woff[i] + np.arange(len(s) - k + 1, dtype=np.int64) // 32
            sel = (w >= t0 * 32) & (w < t1 * 32)
            if not sel.any():
                continue
            ks, rv = self.oracle.chop(s, k)
            pos = np.nonzero(sel)[0]
            K.append(ks[sel]); O.append(rv[sel].astype(np.uint64))
            R.append(np.full(len(pos), i, dtype=np.int64)); P.append(pos.astype(np.int64))
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
        keys, orev, read, pos = cat(K, np.uint64), cat(O, np.uint64), cat(R, np.int64), cat(P, np.int64)
        positions = len(keys)
        if prefilter is not None and len(keys):
            idx, bit = self._filter_bits(keys)
            keep = (prefilter[0].numpy().view(np.uint32)[idx] & bit) != 0
            keys, orev, read, pos = keys[keep], orev[keep], read[keep], pos[keep]
        vals = (read.astype(np.uint64) << np.uint64(32)) | (pos.astype(np.uint64) << np.uint64(1)) | orev
        own = owner_np(keys, n_part) if len(keys) else np.zeros(0, np.int64)
        counts = np.bincount(own, minlength=n_part).astype(np.int64)
        return _Route(counts, keys, vals, own, read, pos, orev, positions)

    def route_keys(self, route, t):
        t.numpy().view(np.uint64)[: route.kmers] = route.keys[route.order] + np.uint64(1)

    def route_records(self, route, t):
        v = t.numpy().view(np.uint64)
        v[0: 2 * route.kmers: 2] = route.keys[route.order] + np.uint64(1)
        v[1: 2 * route.kmers: 2] = route.vals[route.order]

    def table_create(self, n, k):
        return _Table(k)

    def insert(self, table, t, n):
        v = t.numpy().view(np.uint64)[: 2 * n]
        for key, val in zip(v[0::2].tolist(), v[1::2].tolist()):
            table.multi[key] = table.multi.get(key, 0) + 1
            table.val.setdefault(key, val)

    def lookup(self, table, keys, n, answers):
        q = keys.numpy().view(np.uint64)[:n].tolist()
        a = answers.numpy().view(np.uint64)
        for i, key in enumerate(q):
            if table.multi.get(key, 0) == 1:
                a[i] = table.val[key]
                table.ont[key] = table.ont.get(key, 0) + 1
            else:
                a[i] = MISS

    def collect(self, route, answers):
        a = np.empty(route.kmers, dtype=np.uint64)
        a[route.order] = answers.numpy().view(np.uint64)[: route.kmers]
        hit = a != MISS
        out = np.zeros(int(hit.sum()), dtype=api.HIT_DTYPE)
        v = a[hit]
        out["read"] = route.read[hit]
        out["pos"] = route.pos[hit]
        out["tid"] = (v >> np.uint64(32)).astype(np.int32)
        cpos = (v >> np.uint64(1)) & np.uint64(0x3FFFFFFF)
        out["cpos_flags"] = ((cpos << np.uint64(2)) | (v & np.uint64(1)) | (route.orev[hit] << np.uint64(1))).astype(np.uint32)
        return _Hits(out)

    # direct exchange between threads of one process: a window is a numpy array, its reference is itself
    def window(self, n_words):
        return np.zeros(max(int(n_words), 1), dtype=np.uint64).view(_FreeableArray)

    def window_ref(self, window):
        return window

    def route_keys_direct(self, route, owner_refs, owner_off):
        keys = route.keys[route.order] + np.uint64(1)
        seg = np.concatenate([[0], np.cumsum(route.counts)])
        for d in range(len(route.counts)):
            owner_refs[d][owner_off[d]: owner_off[d] + route.counts[d]] = keys[seg[d]: seg[d + 1]]

    def lookup_direct(self, table, key_window, src_count, answer_refs, answer_off):
        first = 0
        for r, n in enumerate(src_count):
            n = int(n)
            for i, key in enumerate(key_window[first: first + n].tolist()):
                if table.multi.get(key, 0) == 1:
                    answer_refs[r][answer_off[r] + i] = table.val[key]
                    table.ont[key] = table.ont.get(key, 0) + 1
                else:
                    answer_refs[r][answer_off[r] + i] = MISS
            first += n

    def collect_window(self, route, answer_window):
        import torch
        return self.collect(route, torch.from_numpy(answer_window.view(np.int64)))

    def stats(self, table):
        total = len(table.multi)
        uniq = sum(1 for m in table.multi.values() if m == 1)
        return (total, uniq, len(table.ont), sum(1 for c in table.ont.values() if c == 1))


def hits_from_oracle(o_hits) -> np.ndarray:
    out = np.zeros(len(o_hits["read"]), dtype=api.HIT_DTYPE)
    out["read"], out["pos"], out["tid"] = o_hits["read"], o_hits["pos"], o_hits["tid"]
    out["cpos_flags"] = (o_hits["cpos"].astype(np.uint32) << 2) | o_hits["krev"].astype(np.uint32) | (o_hits["orev"].astype(np.uint32) << 1)
    return out
