"""GPU parity: CUDA Smith-Waterman + CIGAR (through the C ABI) vs the CPU oracle, bit exact."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from superplus_b200 import api, synth

pytestmark = pytest.mark.gpu


def code(s):
    return np.array([(c >> 1) & 3 for c in s.encode()], dtype=np.uint8)


def oracle_params(P):
    o = orc.SWParams()
    for f, _ in api.SWParams._fields_:
        if f == "mat":
            for i in range(64):
                o.mat[i] = P.mat[i]
        else:
            setattr(o, f, getattr(P, f))
    return o


def check_batch(ctx, oracle, P, qrys, tgts, mode):
    res, cigs = ctx.sw_batch(P, qrys, tgts, mode)
    OP = oracle_params(P)
    for p, (q, t) in enumerate(zip(qrys, tgts)):
        g = oracle.sw_align(OP, q, t, mode)
        r = res[p]
        got = (int(r["score"]), int(r["alignment_offset"]), int(r["has_softclip"]), int(r["bt_tidx"]), int(r["bt_qidx"]), orc.cigar_str(cigs[p]))
        want = (g["score"], g["offset"], g["softclip"], g["bt_tidx"], g["bt_qidx"], orc.cigar_str(g["cigar"]))
        assert got == want, "pair %d (%dx%d) mode %d: got %s want %s" % (p, len(q), len(t), mode, got, want)
    return res


def random_pairs(rng, n, qmax, tmax, related=0.6, nsym=4):
    qs, ts = [], []
    for _ in range(n):
        tl = int(rng.integers(1, tmax + 1)); ql = int(rng.integers(1, qmax + 1))
        t = rng.integers(0, nsym, tl).astype(np.uint8)
        if rng.random() < related and tl > 4:
            a = int(rng.integers(0, tl // 2)); b = int(rng.integers(a + 1, tl + 1))
            core = synth.mutate(t[a:b], 0.15, rng) % nsym
            pre = rng.integers(0, nsym, int(rng.integers(0, max(1, qmax // 4)))).astype(np.uint8)
            post = rng.integers(0, nsym, int(rng.integers(0, max(1, qmax // 4)))).astype(np.uint8)
            q = np.concatenate([pre, core, post]).astype(np.uint8)
            if len(q) == 0:
                q = np.array([1], np.uint8)
        else:
            q = rng.integers(0, nsym, ql).astype(np.uint8)
        qs.append(q); ts.append(t)
    return qs, ts


KATS = [("GACCAGTAGGCATCG", 15, 10, "15M", 10, "15M"), ("GACCAGTGGCATCG", 12, 11, "14M", 10, "7M1D7M"),
        ("GACCAGTAAGGCATCG", 13, 9, "16M", 10, "7M1I8M"), ("TTTTGACCAGTAGGCATCGTTTT", 11, 4, "23M", 7, "1I18M3I1D1M"),
        ("GACCAGTAGGCTTCGATCGGATT", 18, 0, "33M", 10, "11M1I1D11M")]
KAT_TGT = "ACGTACGTTTGACCAGTAGGCATCGATCGGATTACAGATTACA"


def test_known_answers(ctx):
    """SURVEY.md §8c vectors measured on the reference itself (as-is and fixed traceback)."""
    P = api.make_sw_params()
    t = code(KAT_TGT)
    for mode, (oi, ci) in ((api.SW_ASIS, (2, 3)), (api.SW_FIXED, (4, 5))):
        res, cigs = ctx.sw_batch(P, [code(k[0]) for k in KATS], [t] * len(KATS), mode)
        for k, r, c in zip(KATS, res, cigs):
            assert (int(r["score"]), int(r["alignment_offset"]), orc.cigar_str(c)) == (k[1], k[oi], k[ci])


@pytest.mark.parametrize("mode", [api.SW_ASIS, api.SW_FIXED])
@pytest.mark.parametrize("strategy", [0, 1, 2, 3])
def test_random_small_default_scoring(ctx, oracle, mode, strategy):
    rng = np.random.default_rng(100 + strategy)
    qs, ts = random_pairs(rng, 96, 150, 120)
    P = api.make_sw_params(strategy=strategy)
    check_batch(ctx, oracle, P, qs, ts, mode)


@pytest.mark.parametrize("mode", [api.SW_ASIS, api.SW_FIXED])
def test_packed_path_is_used_and_matches_generic(ctx, oracle, mode):
    rng = np.random.default_rng(7)
    qs, ts = random_pairs(rng, 65, 700, 400)        # odd count: one unpaired packed item
    P = api.make_sw_params()
    qb, qo = api._concat(qs)
    tb, to = api._concat(ts)
    b = ctx.swbatch_upload_concat(qb, qo, tb, to)
    b.align(P, mode)
    assert b.path_counts() == (65, 0)
    r1, c1 = b.download()
    os.environ["GCG_SW_FORCE_GENERIC"] = "1"
    try:
        b.align(P, mode)
        assert b.path_counts() == (0, 65)
        r2, c2 = b.download()
    finally:
        del os.environ["GCG_SW_FORCE_GENERIC"]
    b.free()
    for f in ("score", "alignment_offset", "has_softclip", "bt_tidx", "bt_qidx", "n_cigar"):
        assert np.array_equal(r1[f], r2[f]), f
    for a, c in zip(c1, c2):
        assert np.array_equal(a, c)
    check_batch(ctx, oracle, P, qs, ts, mode)


@pytest.mark.parametrize("mode", [api.SW_ASIS, api.SW_FIXED])
def test_generic_scoring_and_alphabet(ctx, oracle, mode):
    rng = np.random.default_rng(11)
    # 5-letter alphabet with symbol 4 present, asymmetric gap costs, non-equality matrix
    qs, ts = random_pairs(rng, 48, 300, 260, nsym=5)
    mat = rng.integers(-9, 6, size=(5, 5)).astype(np.int32)
    np.fill_diagonal(mat, rng.integers(1, 12, size=5))
    for strategy in (0, 2):
        P = api.make_sw_params(mat, del_o=7, del_e=2, ins_o=4, ins_e=3, strategy=strategy)
        check_batch(ctx, oracle, P, qs, ts, mode)
    # big scores force the s32 kernel even on a 4-letter alphabet
    qs, ts = random_pairs(rng, 24, 200, 200)
    P = api.make_sw_params(api.default_mat(4, 50, -25), del_o=110, del_e=6, ins_o=110, ins_e=6)
    check_batch(ctx, oracle, P, qs, ts, mode)


def test_stale_border_state(ctx, oracle):
    """sw_set_parameter(SOFTCLIP) after a non-softclip call keeps the affine borders (sw.c:93-94)."""
    rng = np.random.default_rng(5)
    qs, ts = random_pairs(rng, 32, 90, 90)
    P = api.make_sw_params(strategy=api.SOFTCLIP, border=(1, 3, 1, 2, 2))
    check_batch(ctx, oracle, P, qs, ts, api.SW_FIXED)
    check_batch(ctx, oracle, P, qs, ts, api.SW_ASIS)


@pytest.mark.parametrize("mode", [api.SW_ASIS, api.SW_FIXED])
def test_band_and_row_boundaries(ctx, oracle, mode):
    """lengths around the 256-column band, the 8-column lane and the 32-lane skew"""
    rng = np.random.default_rng(21)
    qs, ts = [], []
    for ql in (1, 7, 8, 9, 255, 256, 257, 511, 513, 1030):
        for tl in (1, 2, 31, 32, 33, 300):
            t = rng.integers(0, 4, tl).astype(np.uint8)
            q = np.resize(np.concatenate([rng.integers(0, 4, ql // 3).astype(np.uint8), synth.mutate(t, 0.1, rng) % 4]), ql).astype(np.uint8)
            qs.append(q); ts.append(t)
    check_batch(ctx, oracle, api.make_sw_params(), qs, ts, mode)


@pytest.mark.parametrize("mode", [api.SW_ASIS, api.SW_FIXED])
def test_packed_pairs_near_the_16_bit_range(ctx, oracle, mode):
    """Affine borders and queries close to the packed kernel's value range.  Each alignment fits
    the 16-bit halves on its own, but the band-padded rectangle two of them would share does not
    (border of column 2048 = -2049): such alignments must not share a warp, because the packed
    kernel adds both halves with one 32-bit add (a borrow out of the low half would reach the
    other alignment).  Shorter ones in the same batch still pair up."""
    rng = np.random.default_rng(77)
    qs, ts = [], []
    for ql in (1700, 1790, 1795, 300, 1100, 1279, 1281, 1500, 1793, 40):
        t = rng.integers(0, 4, 60).astype(np.uint8)
        q = rng.integers(0, 4, ql).astype(np.uint8)
        a = int(rng.integers(0, ql - 30))
        q[a:a + 30] = t[10:40]
        qs.append(q); ts.append(t)
    for strategy in (api.LEADING_INDEL, api.INDEL, api.IGNORE):
        P = api.make_sw_params(strategy=strategy)
        qb, qo = api._concat(qs)
        tb, to = api._concat(ts)
        b = ctx.swbatch_upload_concat(qb, qo, tb, to)
        b.align(P, mode)
        assert b.path_counts() == (len(qs), 0)              # all of them on the packed kernel
        b.free()
        check_batch(ctx, oracle, P, qs, ts, mode)


def test_ties_heavy(ctx, oracle):
    """low-complexity sequences: many equal scores exercise every tie rule"""
    rng = np.random.default_rng(33)
    qs, ts = [], []
    for _ in range(64):
        tl = int(rng.integers(5, 200)); ql = int(rng.integers(5, 400))
        unit = rng.integers(0, 2, int(rng.integers(1, 4))).astype(np.uint8)
        t = np.resize(unit, tl); q = np.resize(unit, ql).copy()
        flip = rng.random(ql) < 0.05
        q[flip] ^= 1
        qs.append(q); ts.append(t)
    for mode in (api.SW_ASIS, api.SW_FIXED):
        for strategy in (0, 3):
            check_batch(ctx, oracle, api.make_sw_params(strategy=strategy), qs, ts, mode)
        check_batch(ctx, oracle, api.make_sw_params(api.default_mat(5, 1, -1), 1, 1, 1, 1), qs, ts, mode)


def test_empty_sides(ctx, oracle):
    e = np.zeros(0, np.uint8)
    a = np.array([0, 1, 2, 3, 1, 1], np.uint8)
    for strategy in (0, 1, 2, 3):
        P = api.make_sw_params(strategy=strategy)
        for mode in (api.SW_ASIS, api.SW_FIXED):
            check_batch(ctx, oracle, P, [e, a, e], [a, e, e], mode)


def test_cfg3_shape_pairs(ctx, oracle):
    """a few full-size cfg3 pairs (10 kb read x 2 kb flank), both modes, packed kernel"""
    q, t = synth.make_sw_pairs(6, 10_000, 2_000, seed=46)
    P = api.make_sw_params()
    for mode in (api.SW_ASIS, api.SW_FIXED):
        res = check_batch(ctx, oracle, P, list(q), list(t), mode)
        assert int(res["score"].min()) > 500            # the embedded flank copy aligns


def test_largest_sizes_of_each_kernel(ctx, oracle):
    """The largest alignments each kernel takes, both traceback modes: target lengths around the
    packed kernel's provable 16-bit range (default scoring: 2 020 rows fit, 2 030 do not and go to the
    s32 kernel), a 20 kb query on the packed kernel, and a 20 kb x 4 kb pair on the s32 kernel."""
    rng = np.random.default_rng(91)

    def pair(ql, tl):
        t = rng.integers(0, 4, tl).astype(np.uint8)
        q = rng.integers(0, 4, ql).astype(np.uint8)
        cp = synth.mutate(t, 0.1, rng) % 4
        cp = cp[:ql]
        a = int(rng.integers(0, ql - len(cp) + 1))
        q[a:a + len(cp)] = cp
        return q, t

    P = api.make_sw_params()
    for (ql, tl), want_paths in (((20_000, 2_020), (1, 0)), ((12_000, 2_030), (0, 1)), ((20_000, 4_000), (0, 1))):
        q, t = pair(ql, tl)
        qb, qo = api._concat([q])
        tb, to = api._concat([t])
        b = ctx.swbatch_upload_concat(qb, qo, tb, to)
        b.align(P, api.SW_ASIS)
        assert b.path_counts() == want_paths, (ql, tl)
        b.free()
        for mode in (api.SW_ASIS, api.SW_FIXED):
            res = check_batch(ctx, oracle, P, [q], [t], mode)
            assert int(res["score"][0]) > tl // 4          # the embedded copy aligns


def test_multi_wave_and_pool_growth(ctx, oracle):
    rng = np.random.default_rng(2)
    qs, ts = random_pairs(rng, 200, 600, 500)
    os.environ["GCG_SW_TRACE_BUDGET_MB"] = "2"
    os.environ["GCG_SW_POOL_INIT"] = "64"
    try:
        check_batch(ctx, oracle, api.make_sw_params(), qs, ts, api.SW_FIXED)
    finally:
        del os.environ["GCG_SW_TRACE_BUDGET_MB"]
        del os.environ["GCG_SW_POOL_INIT"]


def test_symbol_out_of_alphabet_is_rejected(ctx):
    P = api.make_sw_params(api.default_mat(4))
    with pytest.raises(api.GcgError):
        ctx.sw_batch(P, [np.array([0, 4, 1], np.uint8)], [np.array([0, 1], np.uint8)])


@pytest.mark.parametrize("n_ctx", [2, 3, 5])
def test_batch_sharded_over_contexts(ctx, n_ctx):
    """gcg_sw_batch_multi: contiguous pair ranges of about equal cells on several contexts (every GPU of
    the box in turn; on a one-GPU box several contexts of that GPU) give exactly the results and CIGARs
    of one gcg_sw_batch call; empty shares (fewer pairs than contexts, a batch of nothing) included"""
    ndev = int(api.load_library().gcg_device_count())
    others = [api.Context(i % ndev) for i in range(1, n_ctx)]
    try:
        rng = np.random.default_rng(77 + n_ctx)
        qs, ts = random_pairs(rng, 41, 300, 200)
        qs[5], ts[5] = np.zeros(0, np.uint8), ts[5]                      # an empty query inside a share
        q3, t3 = synth.make_sw_pairs(3, 3000, 700, seed=5)               # a few heavy pairs shift the boundaries
        qs += list(q3); ts += list(t3)
        for P in (api.make_sw_params(), api.make_sw_params(strategy=1), api.make_sw_params(api.default_mat(5, 2, -3), 4, 2, 3, 1)):
            for mode in (api.SW_ASIS, api.SW_FIXED):
                want_res, want_cig = ctx.sw_batch(P, qs, ts, mode)
                got_res, got_cig = ctx.sw_batch_multi(others, P, qs, ts, mode)
                for f in ("score", "alignment_offset", "has_softclip", "bt_tidx", "bt_qidx", "n_cigar"):
                    assert np.array_equal(got_res[f], want_res[f]), f
                assert all(np.array_equal(a, b) for a, b in zip(got_cig, want_cig))
        P = api.make_sw_params()
        for m in (0, 1, n_ctx - 1):
            want_res, want_cig = ctx.sw_batch(P, qs[:m], ts[:m])
            got_res, got_cig = ctx.sw_batch_multi(others, P, qs[:m], ts[:m])
            assert np.array_equal(got_res["score"], want_res["score"]) and all(np.array_equal(a, b) for a, b in zip(got_cig, want_cig))
    finally:
        for c in others:
            c.close()
