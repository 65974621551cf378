set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_part_gpu.py -x -q -m gpu -k "nccl_world2" > $O/pytest_remote_n2.log 2>&1; echo "pytest n2 rc=$?"; tail -3 $O/pytest_remote_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-hbm-table --sw-pairs 5920 --cfg4 on --cfg4-genome 20000000 > $O/bench_remote_n2.json 2> $O/bench_remote_n2.err; echo "bench n2 rc=$?"; tail -3 $O/bench_remote_n2.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_remote_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.3e (%.3f ms) e2e %.3e packed %.3e runs_packed %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_packed"]["value"], d["e2e_runs_packed"]["value"]))
for k in ("partitioned", "partitioned_direct", "partitioned_all_to_all", "partitioned_cfg4"):
    if k in d:
        p = d[k]; print(k, "%.3e" % p["value"], p.get("ms_per_step", p.get("ms_per_search")), p.get("check"), {a: round(b, 3) for a, b in p.get("kernel_ms_per_step", p.get("kernel_ms_per_search")).items()})
PY
