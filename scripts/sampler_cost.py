"""does the NVML clock sampler of bench.py slow the SW step down?  (diagnosis)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from superplus_b200 import api, synth
ctx = api.Context(0)
bq, bt = synth.make_sw_pairs(256, 10000, 2000, seed=46)
n = 11840
q = np.tile(bq, (n // 256 + 1, 1))[:n]; t = np.tile(bt, (n // 256 + 1, 1))[:n]
b = ctx.swbatch_upload(q, t)
P = api.make_sw_params()
for _ in range(2):
    b.align(P, api.SW_ASIS)
for period in (None, 0.05, 0.25, 1.0):      # 0.05 s was bench.py's first setting
    s = None
    if period is not None:
        s = bench.ClockSampler(0)
        s.period = period
        s.start()
    t0 = time.perf_counter()
    for _ in range(5):
        b.align(P, api.SW_ASIS)
    dt = (time.perf_counter() - t0) / 5
    if s is not None:
        r = s.result()
    print("sampler period %s: %.2f ms per align" % (period, dt * 1e3), flush=True)
