set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_kmer_gpu.py tests/test_gc_e2e_gpu.py -x -q -m gpu > $O/pytest_r02k.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_r02k.log
GCG_TRACE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hbm-table --sw-pairs 5920 > $O/bench_r02k.json 2> $O/bench_r02k.err; echo "bench rc=$?"; tail -3 $O/bench_r02k.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02k.json").read().strip().splitlines()[-1])
print("value %.3e (%.3f ms) e2e %.3e (%.2f ms) packed %.3e (%.2f ms) runs %.3e (%.2f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_packed"]["value"], d["e2e_packed"]["ms_per_step"], d["e2e_runs"]["value"], d["e2e_runs"]["ms_per_step"]))
print(d["roofline"]["kernel_ms_per_step"])
PY
grep "search pipeline" $O/bench_r02k.err | tail -12 | cut -c1-220
python scripts/cli_modes_time.py cfg2 > $O/cli_modes_cfg2_k.log 2>&1; grep "wall\|pack on ingest" $O/cli_modes_cfg2_k.log | cut -c1-160
