set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_kmer_gpu.py tests/test_gc_e2e_gpu.py -x -q -m gpu > $O/pytest_r02c.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_r02c.log
GCG_SEARCH_ZEROCOPY=0 timeout 600 python -m pytest tests/test_kmer_gpu.py -x -q -m gpu > $O/pytest_r02c_staged.log 2>&1; echo "pytest staged rc=$?"; tail -3 $O/pytest_r02c_staged.log
GCG_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hbm-table --sw-pairs 5920 > $O/bench_r02c.json 2> $O/bench_r02c.err; echo "bench rc=$?"
grep "search pipeline\|search: reads" $O/bench_r02c.err | tail -4
GCG_SEARCH_ZEROCOPY=0 GCG_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hbm-table --sw-pairs 5920 > $O/bench_r02c_staged.json 2> $O/bench_r02c_staged.err; echo "bench staged rc=$?"
grep "search pipeline\|search: reads" $O/bench_r02c_staged.err | tail -4
python - <<PY
import json
for n in ("bench_r02c", "bench_r02c_staged"):
    d = json.loads(open("$O/%s.json" % n).read().strip().splitlines()[-1])
    print(n, "value %.3e ms %.3f e2e %.3e (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
