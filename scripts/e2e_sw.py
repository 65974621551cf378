"""host-buffer SW call timing with the library's phase trace (GCG_TRACE=1)"""
import os, sys, time
os.environ.setdefault("GCG_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from superplus_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5920
bq, bt = synth.make_sw_pairs(256, 10000, 2000, seed=46)
reps = (n + 255) // 256
q = np.ascontiguousarray(np.tile(bq, (reps, 1))[:n]).reshape(-1); t = np.ascontiguousarray(np.tile(bt, (reps, 1))[:n]).reshape(-1)
qo = np.arange(n + 1, dtype=np.int64) * 10000; to = np.arange(n + 1, dtype=np.int64) * 2000
ctx = api.Context(0, host_threads=16)
P = api.make_sw_params()
for it in range(4):
    res = np.zeros(n, dtype=api.SWRES_DTYPE); pool, npool = C.c_void_p(), C.c_int64()
    t0 = time.perf_counter()
    ctx._chk(ctx.L.gcg_sw_batch(ctx.h, C.byref(P), api.SW_ASIS, q.ctypes.data, qo.ctypes.data, t.ctypes.data, to.ctypes.data, n, res.ctypes.data, C.byref(pool), C.byref(npool)))
    dt = time.perf_counter() - t0
    ctx.L.gcg_free(pool)
    print("call %d: %.1f ms  %.1f GCUPS" % (it, dt * 1e3, n * 2e7 / dt / 1e9), file=sys.stderr)
