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


@pytest.mark.parametrize("cfg,n_thread", [("tiny", 1), ("small", 4), ("repeats", 3), ("cfg1", 8)])
def test_gap_filled_fasta_is_bit_exact(cfg, n_thread):
    assert os.path.exists(GC), "gc_b200 was not built (python -c 'import __graft_entry__ as g; g.build()')"
    with tempfile.TemporaryDirectory() as tmp:
        fa, fq, _ = synth.materialise(cfg, tmp)
        wd = os.path.join(tmp, "run")
        os.makedirs(wd)
        r = subprocess.run([GC, fa, fq, str(n_thread), "out"], cwd=wd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
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


def test_search_in_groups_matches_single_call():
    """gcg_search cuts huge read sets into groups; force tiny groups and compare"""
    import numpy as np
    from superplus_b200 import api
    ctx = api.Context(0)
    inp = synth.make_config("tiny")
    cs = ctx.upload(inp.contigs)
    t = ctx.table_build(cs, 25)
    a = ctx.search_host(t, inp.reads)
    st_a = t.stats()
    t.free()
    t = ctx.table_build(cs, 25)
    os.environ["GCG_SEARCH_GROUP_KMERS"] = "40000"
    try:
        b = ctx.search_host(t, inp.reads)
    finally:
        del os.environ["GCG_SEARCH_GROUP_KMERS"]
    assert np.array_equal(a, b) and t.stats() == st_a
    t.free(); cs.free(); ctx.close()
