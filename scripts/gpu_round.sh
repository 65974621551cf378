#!/usr/bin/env bash
# One GPU-box visit: parity tests, the bench line, the gather ceiling, then the ncu evidence
# (launch list of the bench command + one full capture of each dominant kernel).
# usage: scripts/gpu_round.sh [tag]      outputs land in gpurun_out/
set -u
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$TAG.log
GCG_TRACE=1 python bench.py --steps 5 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_$TAG.json").read().strip().splitlines()[-1])
    print("value %.3e k-mers/s  ms/step %.3f  e2e %.3e (%.1f ms)  frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
    print(d["roofline"]["kernel_ms_per_step"])
    print("sw %.1f GCUPS e2e %.1f frac %.3f" % (d["sw"]["value"], d["sw"]["e2e"]["value"], d["sw"]["roofline"]["frac"]))
    print("cpu", d.get("cpu_baseline"), d["sw"].get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
grep "\[gcg\]" $O/bench_$TAG.err | tail -24
python - <<PY
from superplus_b200 import api
c = api.Context(0)
for mb in (53, 424, 1600, 8000):
    print("gather %5d MB: %.1f G lookups/s" % (mb, c.ubench_gather(mb << 20) / 1e9))
print("hbm copy %.0f GB/s  int16 %.2f T lane-ops/s" % (c.ubench_hbm() / 1e9, c.ubench_int16() / 1e12))
PY
# ---- ncu: launch list of the bench command (shares of the step), then one full capture per kernel
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sw-pairs 11840"
$BCMD > $O/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv $BCMD > $O/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
KCMD="python scripts/perf_kmer.py cfg2 25 3"
$KCMD > $O/plain_kmer.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k45_search -s 2 -c 1 -o $O/prof_k45_$TAG -f $KCMD > $O/ncu_k45.log 2>&1
echo "ncu k45 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:hits_emit -s 2 -c 1 -o $O/prof_emit_$TAG -f $KCMD > $O/ncu_emit.log 2>&1
echo "ncu emit rc=$?"
# the counting kernel: atomicCAS claims + atomicOr flags into the HBM table (atomic throughput, north star)
ncu --set full --clock-control none --import-source on -k regex:k23_build -s 2 -c 1 -o $O/prof_k23_$TAG -f $KCMD > $O/ncu_k23.log 2>&1
echo "ncu k23 rc=$?"
SCMD="python scripts/perf_sw.py 5920 0 1"
$SCMD > $O/plain_sw.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sw_fill_packed -c 1 -o $O/prof_sw_packed_$TAG -f $SCMD > $O/ncu_sw.log 2>&1
echo "ncu sw rc=$?"
tail -4 $O/plain_kmer.log
