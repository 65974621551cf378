"""Deterministic synthetic inputs for the gap-closing hot path.

genome   : i.i.d. uniform ACGT, PCG64(seed)
scaffold : the genome with `n_gaps` evenly spaced N-runs, lengths U[100, 5000)
ONT reads: length ~ Gamma(shape 4, mean 10 kb) clipped to [1 k, 60 k], uniform start on the
           *un-gapped* genome, 50 % reverse-complemented, per-base error `err` split equally
           into substitution / deletion / insertion, quality 'I'
SW pairs : target = random `tlen`-mer flank, query = `qlen` bases holding an `err`-error copy of
           the flank at a uniform offset, the rest random; symbols integer-coded 0..3 with the
           reference's code (bio.h:24: A=0 C=1 T=2 G=3), which is what sw_align indexes `mat` by
           (sw.c:216).

Everything is numpy so the same bytes come out here and on the GPU box (same image).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

# reference code order (bio.h:23): int2base = "ACTG"
_INT2BASE = np.frombuffer(b"ACTG", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


@dataclass
class GapCloserInput:
    scaffold: np.ndarray          # uint8 ASCII, with N runs
    genome: np.ndarray            # uint8 ASCII, truth
    gaps: list                    # (start, length)
    reads: list                   # list of uint8 ASCII arrays

    @property
    def contigs(self):
        """Contigs exactly as contig.c:142-196 splits them (maximal non-N runs; an N at
        position 0 yields an empty leading contig)."""
        return split_contigs(self.scaffold)


def split_contigs(s: np.ndarray) -> list:
    """Maximal non-N runs of one scaffold in order (contig.c:142-196); `is_N` is the low-nibble
    test of kmer.h:18.  A scaffold that starts with N yields an empty leading contig, as the
    reference's state machine (initial state NOT_N) does."""
    if len(s) == 0:
        return []
    is_n = ((s & 0xF) == 0xE).astype(np.int8)
    idx = np.flatnonzero(np.diff(is_n))
    bounds = np.concatenate(([0], idx + 1, [len(s)]))
    state = bool(is_n[0])
    out = []
    if state:
        out.append(s[0:0])
    for a, b in zip(bounds[:-1], bounds[1:]):
        if not state:
            out.append(s[a:b])
        state = not state
    return out


def random_genome(n: int, rng: np.random.Generator) -> np.ndarray:
    return _INT2BASE[rng.integers(0, 4, size=n, dtype=np.uint8)]


def revcomp(seq: np.ndarray) -> np.ndarray:
    return _COMP[seq[::-1]]


def mutate(seq: np.ndarray, err: float, rng: np.random.Generator) -> np.ndarray:
    """10 %-style ONT noise: each base independently substituted / deleted / followed by an insertion."""
    n = len(seq)
    u = rng.random(n)
    sub = u < err / 3
    dele = (u >= err / 3) & (u < 2 * err / 3)
    ins = (u >= 2 * err / 3) & (u < err)
    out = seq.copy()
    ns = int(sub.sum())
    if ns:
        # substitute with a *different* base
        cur = out[sub]
        code = np.searchsorted(np.sort(_INT2BASE), cur)  # index in sorted ACGT
        alt = (code + rng.integers(1, 4, size=ns)) % 4
        out[sub] = np.sort(_INT2BASE)[alt]
    counts = np.ones(n, dtype=np.int64)
    counts[dele] = 0
    counts[ins] = 2
    res = np.repeat(out, counts)
    ni = int(ins.sum())
    if ni:
        ends = np.cumsum(counts) - 1                # index of last copy of every base
        res[ends[ins]] = _INT2BASE[rng.integers(0, 4, size=ni)]
    return res


def make_gap_closer_input(genome_len: int, n_gaps: int, coverage: float, seed: int,
                          err: float = 0.10, mean_len: float = 10_000.0, n_repeats: int = 0) -> GapCloserInput:
    rng = np.random.Generator(np.random.PCG64(seed))
    genome = random_genome(genome_len, rng)
    # optional repeats (forward or reverse-complement copies) so that multi >= 2 k-mers exist
    for _ in range(n_repeats):
        L = int(rng.integers(200, 3000))
        a = int(rng.integers(0, genome_len - L))
        b = int(rng.integers(0, genome_len - L))
        seg = genome[a:a + L].copy()
        if rng.random() < 0.5:
            seg = revcomp(seg)
        genome[b:b + L] = seg
    scaffold = genome.copy()
    gaps = []
    if n_gaps > 0:
        starts = (np.arange(1, n_gaps + 1) * (genome_len // (n_gaps + 1))).astype(np.int64)
        lens = rng.integers(100, 5000, size=n_gaps)
        for s, l in zip(starts, lens):
            l = int(min(l, genome_len - s - 1))
            if l <= 0:
                continue
            scaffold[s:s + l] = ord("N")
            gaps.append((int(s), l))
    target_bases = int(coverage * genome_len)
    reads = []
    total = 0
    while total < target_bases:
        L = int(np.clip(rng.gamma(4.0, mean_len / 4.0), 1000, 60000))
        L = min(L, genome_len)
        st = int(rng.integers(0, genome_len - L + 1))
        frag = genome[st:st + L]
        if rng.random() < 0.5:
            frag = revcomp(frag)
        r = mutate(frag, err, rng)
        reads.append(r)
        total += len(r)
    return GapCloserInput(scaffold=scaffold, genome=genome, gaps=gaps, reads=reads)


def make_reads(genome: np.ndarray, coverage: float, seed: int, err: float = 0.10, mean_len: float = 10_000.0) -> list:
    """an independent batch of simulated ONT reads over `genome` (multi-GPU runs: one batch per rank)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    genome_len = len(genome)
    target_bases = int(coverage * genome_len)
    reads, total = [], 0
    while total < target_bases:
        L = int(np.clip(rng.gamma(4.0, mean_len / 4.0), 1000, 60000))
        L = min(L, genome_len)
        st = int(rng.integers(0, genome_len - L + 1))
        frag = genome[st:st + L]
        if rng.random() < 0.5:
            frag = revcomp(frag)
        r = mutate(frag, err, rng)
        reads.append(r)
        total += len(r)
    return reads


def write_fasta(path: str, seq: np.ndarray, name: str = "scaffold1", width: int = 60) -> None:
    n = len(seq)
    full = n // width
    body = seq[:full * width].reshape(full, width)
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        if full:
            lines = np.concatenate([body, np.full((full, 1), 10, dtype=np.uint8)], axis=1)
            f.write(lines.tobytes())
        if n % width:
            f.write(seq[full * width:].tobytes() + b"\n")


def write_fastq(path: str, reads: list) -> None:
    """Strict 4-line plain FASTQ ending in '\\n' (what rseq.c:307-374 requires)."""
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b"@ont%d\n" % i)
            f.write(r.tobytes())
            f.write(b"\n+\n")
            f.write(b"I" * len(r))
            f.write(b"\n")


# BASELINE.json configs
CONFIGS = {
    "cfg1": dict(genome_len=1_000_000, n_gaps=50, coverage=20.0, seed=42),
    "cfg2": dict(genome_len=4_600_000, n_gaps=500, coverage=30.0, seed=43),
    "cfg4": dict(genome_len=100_000_000, n_gaps=0, coverage=40.0, seed=44),
    "cfg5": dict(genome_len=250_000_000, n_gaps=20_000, coverage=20.0, seed=45),
    # cfg4 with a tenth of the reads: the 100 Mb table (HBM resident, 1.6 GB of keys) on one GPU
    "cfg4s": dict(genome_len=100_000_000, n_gaps=0, coverage=4.0, seed=44),
    # cfg5's gap density (80 per Mb) at 20 Mb / 6x: a table beyond the L2 (320 MB of keys: the pre-filter
    # path) and a read set of many pipeline chunks, small enough for the reference to make its fixture
    "cfg5s": dict(genome_len=20_000_000, n_gaps=1_600, coverage=6.0, seed=45),
    # small cases for unit tests
    "tiny": dict(genome_len=60_000, n_gaps=4, coverage=8.0, seed=7),
    "small": dict(genome_len=200_000, n_gaps=10, coverage=10.0, seed=11),
    "repeats": dict(genome_len=120_000, n_gaps=6, coverage=6.0, seed=13, n_repeats=25),
}


def make_config(name: str) -> GapCloserInput:
    return make_gap_closer_input(**CONFIGS[name])


def materialise(name: str, out_dir: str):
    """Write <out_dir>/<name>.fa and .fq if absent; return (fa, fq, GapCloserInput or None)."""
    os.makedirs(out_dir, exist_ok=True)
    fa = os.path.join(out_dir, name + ".fa")
    fq = os.path.join(out_dir, name + ".fq")
    inp = None
    if not (os.path.exists(fa) and os.path.exists(fq)):
        inp = make_config(name)
        write_fasta(fa, inp.scaffold)
        write_fastq(fq, inp.reads)
    return fa, fq, inp


def make_sw_pairs(n_pairs: int, qlen: int = 10_000, tlen: int = 2_000, seed: int = 46,
                  err: float = 0.10):
    """cfg3 pairs.  Returns (qry[n,qlen], tgt[n,tlen]) uint8 integer-coded 0..3."""
    rng = np.random.Generator(np.random.PCG64(seed))
    code = np.zeros(256, dtype=np.uint8)
    for ch in b"ACGT":
        code[ch] = (ch >> 1) & 3                   # bio.h:24
    qry = rng.integers(0, 4, size=(n_pairs, qlen), dtype=np.uint8)
    tgt_ascii = _INT2BASE[rng.integers(0, 4, size=(n_pairs, tlen), dtype=np.uint8)]
    tgt = code[tgt_ascii]
    for p in range(n_pairs):
        cp = code[mutate(tgt_ascii[p], err, rng)]
        if len(cp) > qlen:
            cp = cp[:qlen]
        off = int(rng.integers(0, qlen - len(cp) + 1))
        qry[p, off:off + len(cp)] = cp
    return qry, tgt
