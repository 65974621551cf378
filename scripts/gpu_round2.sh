#!/usr/bin/env bash
# One GPU-box visit of round 2: parity tests (also against the bounds-checked build), the bench line, then the ncu
# evidence (launch list of the bench command + one full capture of each dominant kernel).
# usage: scripts/gpu_round2.sh [tag]      outputs land in gpurun_out/
set -u
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$TAG.log
if [ -f superplus_b200/_build/variants/libgcgpu_checked.so ]; then
  GCG_LIB=$PWD/superplus_b200/_build/variants/libgcgpu_checked.so python -m pytest tests/test_kmer_gpu.py tests/test_part_gpu.py tests/test_scale_gpu.py -x -q -m gpu > $O/pytest_checked_$TAG.log 2>&1
  echo "pytest (GCG_CHECKED build) rc=$?"; tail -2 $O/pytest_checked_$TAG.log; grep -c "GCG_CHECKED violation" $O/pytest_checked_$TAG.log
fi
GCG_TRACE=1 python bench.py --steps 10 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_$TAG.json").read().strip().splitlines()[-1])
    print("value %.3e k-mers/s  ms/step %.3f  e2e %.3e (%.2f ms)  e2e_runs %.3e (%.2f ms)  frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_runs"]["value"], d["e2e_runs"]["ms_per_step"], d["roofline"]["frac"]))
    print(d["roofline"]["kernel_ms_per_step"])
    print("sw fixed %.1f GCUPS e2e %.1f frac %.3f ; asis %.1f frac %.3f" % (d["sw"]["value"], d["sw"]["e2e"]["value"], d["sw"]["roofline"]["frac"], d["sw"]["asis"]["value"], d["sw"]["asis"]["roofline_frac"]))
    h = d.get("roofline_hbm_table"); print("hbm table: %.3e k-mers/s frac %.3f" % (h["kmers_per_s"], h["frac"]) if h else None)
    print("cpu", d.get("cpu_baseline"), d["sw"].get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
# ---- ncu: launch list of the bench command (shares of the step), then one full capture per kernel
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hbm-table --sw-pairs 11840"
$BCMD > $O/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv $BCMD > $O/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
KCMD="python scripts/perf_kmer.py cfg2 25 3"
$KCMD > $O/plain_kmer.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k45_fused -s 2 -c 1 -o $O/prof_k45f_$TAG -f $KCMD > $O/ncu_k45.log 2>&1
echo "ncu k45_fused rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k23_build -s 2 -c 1 -o $O/prof_k23_$TAG -f $KCMD > $O/ncu_k23.log 2>&1
echo "ncu k23 rc=$?"
SCMD="python scripts/perf_sw.py 5920 1 1"
$SCMD > $O/plain_sw.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sw_fill_packed -c 1 -o $O/prof_sw_packed_fixed_$TAG -f $SCMD > $O/ncu_sw.log 2>&1
echo "ncu sw rc=$?"
tail -8 $O/plain_kmer.log
