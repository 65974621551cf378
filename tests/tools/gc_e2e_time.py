"""Wall-clock of the whole gap_closer CLI on one config: reference gc (oracle/_ref, host cores) vs
gc_b200 (same callers + the CUDA path), outputs compared by md5, stdout lines time-stamped through
a pty so that the phases the CLI does not time itself show up.  usage: tests/tools/gc_e2e_time.py cfg2 [threads] [devices]
(devices, e.g. 0,1: one more gc_b200 run with GC_DEVICES set — the reads sharded over those GPUs)"""
import hashlib, os, pty, select, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from superplus_b200 import synth
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
nt = sys.argv[2] if len(sys.argv) > 2 else str(os.cpu_count())
md5 = lambda p: hashlib.md5(open(p, "rb").read()).hexdigest()


def run(exe, args, cwd, extra_env=None):
    m, s = pty.openpty()
    t0 = time.perf_counter()
    p = subprocess.Popen([exe] + args, cwd=cwd, stdout=s, stderr=s, env=dict(os.environ, GCG_TRACE='1', **(extra_env or {})))
    os.close(s)
    buf, lines = b"", []
    while True:
        r, _, _ = select.select([m], [], [], 0.2)
        if r:
            try:
                d = os.read(m, 65536)
            except OSError:
                d = b""
            if not d:
                break
            buf += d
            while b"\n" in buf:
                l, buf = buf.split(b"\n", 1)
                lines.append((time.perf_counter() - t0, l.decode(errors="replace").strip()))
        elif p.poll() is not None:
            break
    p.wait()
    return p.returncode, time.perf_counter() - t0, lines


with tempfile.TemporaryDirectory() as tmp:
    fa, fq, _ = synth.materialise(cfg, tmp)
    res = {}
    runs = [("reference_gc", os.path.join(ROOT, "oracle", "_ref", "gc"), None), ("gc_b200", os.path.join(ROOT, "superplus_b200", "_build", "gc_b200"), None),
            ("gc_b200_again", os.path.join(ROOT, "superplus_b200", "_build", "gc_b200"), None)]
    if len(sys.argv) > 3:
        runs.append(("gc_b200_devices_" + sys.argv[3], os.path.join(ROOT, "superplus_b200", "_build", "gc_b200"), {"GC_DEVICES": sys.argv[3]}))
    for name, exe, extra in runs:
        wd = os.path.join(tmp, name)
        os.makedirs(wd)
        rc, dt, lines = run(exe, [fa, fq, nt, "out"], wd, extra)
        print("== %s: rc %d, %.2f s wall, %s threads" % (name, rc, dt, nt))
        for t, l in lines:
            if l and ("cost" in l.lower() or "kmer count" in l or "gcg" in l):
                print("   %7.2fs  %s" % (t, l))
        res[name] = tuple(md5(os.path.join(wd, f)) for f in ("gc_fix1.fa", "ont_link.txt", "valid_ont_link.txt"))
    print("outputs identical:", len(set(res.values())) == 1)
