"""GPU: the reference-facing functions of the replacement sw.c / cigar.c (sw_init,
sw_set_parameter, sw_align — sw.h:66-71) driven through ctypes exactly like the reference's own
library, against fixtures produced by the reference (tests/golden/sw_vectors.json), including call
sequences in which the aligner's border state goes stale (sw.c:93-94) and the matrix grows."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "superplus_b200", "_build", "libgcshim.so")
OPS = "MIDNSHP=XB"


class Cigar(C.Structure):
    _fields_ = [("c", C.POINTER(C.c_uint32)), ("n", C.c_int32), ("m", C.c_int32)]


class SW(C.Structure):          # sw_t, sw.h:38-60
    _fields_ = [("sm", C.c_void_p), ("sm_vol", C.c_int32), ("m_qry", C.c_int32), ("m_tgt", C.c_int32),
                ("score", C.c_int32), ("alignment_offset", C.c_int32), ("status", C.c_int8),
                ("type_c", C.c_int8), ("qry_nbits", C.c_int8), ("overhang_strategy", C.c_int8),
                ("has_softclip", C.c_int64), ("mat", C.POINTER(C.c_int32)),
                ("del_o", C.c_int32), ("del_e", C.c_int32), ("ins_o", C.c_int32), ("ins_e", C.c_int32),
                ("cigar", C.POINTER(Cigar))]


@pytest.fixture(scope="module")
def shim():
    assert os.path.exists(SHIM), "libgcshim.so was not built"
    L = C.CDLL(SHIM)
    L.sw_init.restype = C.POINTER(SW)
    L.sw_set_parameter.argtypes = [C.POINTER(SW), C.c_int, C.c_void_p] + [C.c_int32] * 4 + [C.c_int]
    L.sw_align.argtypes = [C.POINTER(SW), C.c_int32, C.c_char_p, C.c_int32, C.c_char_p]
    L.sw_free.argtypes = [C.POINTER(SW)]
    return L


def result(sw):
    c = sw.contents.cigar.contents
    cig = "".join("%d%s" % (c.c[i] >> 4, OPS[c.c[i] & 15]) for i in range(c.n)) if c.n else "*"
    return dict(score=sw.contents.score, offset=sw.contents.alignment_offset, softclip=int(sw.contents.has_softclip), cigar=cig)


def align(L, sw, q, t):
    qa = np.array(q, np.uint8); ta = np.array(t, np.uint8)
    assert L.sw_align(sw, len(qa), qa.ctypes.data_as(C.c_char_p), len(ta), ta.ctypes.data_as(C.c_char_p)) == 0
    return result(sw)


def test_struct_layout(shim):
    assert C.sizeof(SW) == 72 and C.sizeof(Cigar) == 16


@pytest.mark.parametrize("mode", ["asis", "fixed"])
def test_independent_cases(shim, mode):
    shim.sw_set_traceback_mode(1 if mode == "fixed" else 0)
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "sw_vectors.json")))["cases"]
    for n, c in enumerate(cases[::2]):
        sw = shim.sw_init()
        mat = np.array(c["mat"], dtype=np.int32)
        shim.sw_set_parameter(sw, c["type_c"], mat.ctypes.data, *c["pen"], c["strategy"])
        assert align(shim, sw, c["q"], c["t"]) == c[mode], n
        shim.sw_free(sw)
    shim.sw_set_traceback_mode(0)


@pytest.mark.parametrize("mode", ["asis", "fixed"])
def test_call_sequences_keep_reference_border_state(shim, mode):
    shim.sw_set_traceback_mode(1 if mode == "fixed" else 0)
    mat = np.full((5, 5), -5, dtype=np.int32); np.fill_diagonal(mat, 1)
    for steps in json.load(open(os.path.join(ROOT, "tests", "golden", "sw_vectors.json")))["sequences"]:
        sw = shim.sw_init()
        for st in steps:
            if st["set"]:
                shim.sw_set_parameter(sw, 5, mat.ctypes.data, *st["pen"], st["strategy"])
            assert align(shim, sw, st["q"], st["t"]) == st[mode]
        shim.sw_free(sw)
    shim.sw_set_traceback_mode(0)


@pytest.mark.parametrize("devices", [None, "0,0", "0,0,0", "all"])
def test_batch_entry_point_matches_single_calls(devices):
    """sw_align_batch (sw_batch.h) against sw_align pair by pair on the same aligner; with GC_DEVICES the
    pairs are sharded over several contexts (gcg_sw_batch_multi) — a fresh process per setting, the
    bridge reads the variable once"""
    import subprocess
    import sys
    env = dict(os.environ)
    env.pop("GC_DEVICES", None)
    if devices:
        env["GC_DEVICES"] = devices
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "shim_batch_driver.py"), "3"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=env, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    out = json.loads(r.stdout.decode().strip().splitlines()[-1])
    assert len(out["single"]) == 37 and out["batch"] == out["single"]
