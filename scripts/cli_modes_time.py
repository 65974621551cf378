"""wall time of gc_b200 on one config in its modes (default / GC_RUNS=1 / + GC_SW_FILL=1), two runs each"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from superplus_b200 import synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
tmp = tempfile.mkdtemp()
fa, fq, _ = synth.materialise(cfg, tmp)
exe = os.path.join(ROOT, "superplus_b200", "_build", "gc_b200")
for env in ({}, {"GC_RUNS": "1"}, {"GC_RUNS": "1", "GC_SW_FILL": "1"}):
    for it in range(2):
        wd = os.path.join(tmp, "w%d%s" % (it, "".join(sorted(env)) or "d")); os.makedirs(wd)
        t0 = time.time()
        r = subprocess.run([exe, fa, fq, str(os.cpu_count()), "out"], cwd=wd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, GCG_TRACE="1", **env))
        print(cfg, env or "default", "rc", r.returncode, "wall %.2f s" % (time.time() - t0), [l.strip() for l in r.stdout.decode().splitlines() if "cost" in l.lower()], flush=True)
        if it == 1:
            print("\n".join(l for l in r.stderr.decode().splitlines() if l.startswith("[gcg]"))[-1800:], flush=True)
