set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_kmer_gpu.py tests/test_gc_e2e_gpu.py -x -q -m gpu > $O/pytest_r02d.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_r02d.log
for mode in "GCG_SEARCH_DIRECT=1" "GCG_SEARCH_DIRECT=0" "GCG_SEARCH_ZEROCOPY=1"; do
  env $mode timeout 600 python -m pytest tests/test_kmer_gpu.py -x -q -m gpu > $O/pytest_r02d_mode.log 2>&1; echo "pytest $mode rc=$?"; tail -1 $O/pytest_r02d_mode.log
  env $mode GCG_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hbm-table --sw-pairs 5920 > $O/bench_r02d_$mode.json 2> $O/bench_r02d_$mode.err; echo "bench rc=$?"
  grep "search pipeline\|search: reads" $O/bench_r02d_$mode.err | tail -2
  python - <<PY
import json
d = json.loads(open("$O/bench_r02d_$mode.json").read().strip().splitlines()[-1])
print("$mode", "value %.3e ms %.3f e2e %.3e (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
done
