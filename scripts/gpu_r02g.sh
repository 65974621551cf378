set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_part_gpu.py -x -q -m gpu > $O/pytest_r02g_part.log 2>&1; echo "pytest part rc=$?"; tail -3 $O/pytest_r02g_part.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-hbm-table --sw-pairs 5920 --partitioned --cfg4 on --cfg4-genome 20000000 > $O/bench_r02g.json 2> $O/bench_r02g.err; echo "bench rc=$?"; tail -5 $O/bench_r02g.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02g.json").read().strip().splitlines()[-1])
print("value %.3e e2e %.3e" % (d["value"], d["e2e"]["value"]))
for k in ("partitioned", "partitioned_all_to_all", "partitioned_cfg4"):
    if k in d:
        p = d[k]; print(k, "%.3e" % p["value"], p.get("ms_per_step", p.get("ms_per_search")), p.get("check"), p.get("kernel_ms_per_step", p.get("kernel_ms_per_search")))
PY
