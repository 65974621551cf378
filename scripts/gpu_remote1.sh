set -u
O=gpurun_out; mkdir -p $O
GCG_LIB=$PWD/superplus_b200/_build/variants/libgcgpu_checked.so timeout 900 python -m pytest tests/test_part_gpu.py -x -q -m gpu -k "remote" > $O/pytest_remote_checked.log 2>&1; echo "pytest remote (checked build) rc=$?"; tail -3 $O/pytest_remote_checked.log; grep -c "violation\|never arrived" $O/pytest_remote_checked.log
timeout 900 python -m pytest tests/test_part_gpu.py tests/test_kmer_gpu.py -x -q -m gpu > $O/pytest_remote.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_remote.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-hbm-table --sw-pairs 5920 --partitioned --cfg4 on --cfg4-genome 20000000 > $O/bench_remote1.json 2> $O/bench_remote1.err; echo "bench rc=$?"; tail -3 $O/bench_remote1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_remote1.json").read().strip().splitlines()[-1])
for k in ("partitioned", "partitioned_direct", "partitioned_all_to_all", "partitioned_cfg4"):
    if k in d:
        p = d[k]; print(k, "%.3e" % p["value"], p.get("ms_per_step", p.get("ms_per_search")), p.get("check"), {a: round(b, 3) for a, b in p.get("kernel_ms_per_step", p.get("kernel_ms_per_search")).items()})
PY
