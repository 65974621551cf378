set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_kmer_gpu.py -x -q -m gpu > $O/pytest_r02f_kmer.log 2>&1; echo "pytest kmer rc=$?"; tail -4 $O/pytest_r02f_kmer.log
timeout 1500 python -m pytest tests/test_gc_e2e_gpu.py -x -q -m gpu > $O/pytest_r02f_e2e.log 2>&1; echo "pytest e2e rc=$?"; tail -4 $O/pytest_r02f_e2e.log
# runs mode at cfg2: timing of the CLI phases against the default mode
python - <<'PY'
import os, subprocess, tempfile, time, sys
sys.path.insert(0, os.getcwd())
from superplus_b200 import synth
tmp = tempfile.mkdtemp()
fa, fq, _ = synth.materialise("cfg2", tmp)
for env in ({}, {"GC_RUNS": "1"}):
    for it in range(2):
        wd = os.path.join(tmp, "w%d%s" % (it, "r" if env else "d")); os.makedirs(wd)
        t0 = time.time()
        r = subprocess.run([os.path.join("superplus_b200", "_build", "gc_b200"), fa, fq, "16", "out"], cwd=wd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, GCG_TRACE="1", **env))
        print("cfg2", env or "default", "rc", r.returncode, "wall %.2f s" % (time.time() - t0), [l.strip() for l in r.stdout.decode().splitlines() if "Program Cost" in l])
PY
