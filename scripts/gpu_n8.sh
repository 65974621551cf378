set -u
O=gpurun_out; mkdir -p $O
nvidia-smi -L | wc -l; nproc; free -g | head -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 ${N8_EXTRA:-} > $O/bench_r02_n8b.json 2> $O/bench_r02_n8b.err; echo "bench n8 rc=$?"; tail -4 $O/bench_r02_n8b.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02_n8b.json").read().strip().splitlines()[-1])
print("N=8 value %.3e (%.3f ms) e2e %.3e (%.2f ms) packed %.3e (%.2f ms) runs %.3e (%.2f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_packed"]["value"], d["e2e_packed"]["ms_per_step"], d["e2e_runs"]["value"], d["e2e_runs"]["ms_per_step"]))
print("sw fixed %.0f e2e %.0f asis %.0f" % (d["sw"]["value"], d["sw"]["e2e"]["value"], d["sw"]["asis"]["value"]))
for k in ("partitioned", "partitioned_remote", "partitioned_all_to_all", "partitioned_cfg4"):
    if k in d:
        p = d[k]; print(k, "%.3e" % p["value"], p.get("ms_per_step", p.get("ms_per_search")), p.get("check"), {a: round(b, 3) for a, b in p.get("kernel_ms_per_step", p.get("kernel_ms_per_search")).items()})
PY
