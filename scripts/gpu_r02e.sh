set -u
O=gpurun_out; mkdir -p $O
for cb in 8388608 4194304 2097152; do
  GCG_SEARCH_CHUNK_BYTES=$cb GCG_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hbm-table --sw-pairs 5920 > $O/bench_r02e_$cb.json 2> $O/bench_r02e_$cb.err; echo "bench rc=$?"
  grep "search pipeline\|search: reads" $O/bench_r02e_$cb.err | tail -2
  python - <<PY
import json
d = json.loads(open("$O/bench_r02e_$cb.json").read().strip().splitlines()[-1])
print("chunk $cb", "value %.3e ms %.3f e2e %.3e (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
done
