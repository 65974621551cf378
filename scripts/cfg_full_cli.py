"""Whole CLI at a BASELINE config's full size: reference gc (oracle/_ref/gc, the reference's unmodified sources)
against gc_b200 on the same synthetic input, on this box.  Prints and stores host facts, both wall times, both
phase splits, the md5 of the three output files and the four statistics of each, and whether they agree.

    python scripts/cfg_full_cli.py cfg5 [--threads N] [--devices all|0,1,...] [--workdir DIR] [--skip-ref]

cfg5 (BASELINE configs[4]): 250 Mb, 20 000 gaps, 20x ONT.  Host memory of either program: 16 bytes per ONT base
(okmers, ont.c:483-512) + 24 per contig base (contig.c:203-208) + the reads = about 95 GB.
"""
import argparse, hashlib, json, os, re, subprocess, sys, time, gc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from superplus_b200 import synth  # noqa: E402


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def run(exe, fa, fq, threads, wd, env=None, as_limit_gb=None):
    os.makedirs(wd, exist_ok=True)
    t0 = time.time()

    def limit():
        # the reference must fail with an allocation error, not take the box down, if it needs more than the host has
        import resource
        resource.setrlimit(resource.RLIMIT_AS, (int(as_limit_gb * 2**30), int(as_limit_gb * 2**30)))
    import resource
    ru0 = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss
    r = subprocess.run([exe, fa, fq, str(threads), "out"], cwd=wd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env,
                       preexec_fn=limit if as_limit_gb else None)
    dt = time.time() - t0
    ru1 = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss
    out = r.stdout.decode(errors="replace")
    res = {"rc": r.returncode, "wall_s": dt, "max_rss_gb_of_children_so_far": round(ru1 / 2**20, 1), "max_rss_gb_before": round(ru0 / 2**20, 1), "stats": [int(x) for x in re.findall(r"(?:total|unique) kmer count: (\d+)", out)],
           "phase_lines": [l.strip() for l in out.splitlines() if "cost" in l.lower()],
           "stderr_tail": r.stderr.decode(errors="replace")[-3000:]}
    if r.returncode == 0:
        for name, f in (("fa", "gc_fix1.fa"), ("link", "ont_link.txt"), ("valid", "valid_ont_link.txt")):
            res[name] = md5(os.path.join(wd, f))
        res["n_left"] = sum(l.count(b"N") for l in open(os.path.join(wd, "gc_fix1.fa"), "rb") if not l.startswith(b">"))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cfg")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--devices", default=None)
    ap.add_argument("--workdir", default="/tmp/gc_full")
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--skip-b200", action="store_true")
    ap.add_argument("--runs-first", action="store_true", help="run gc_b200 with GC_RUNS=1 first (its peak memory is then the first child's)")
    ap.add_argument("--golden", default=None, help="compare with this entry of tests/golden/gc_e2e.json instead of running the reference")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    info = {"config": a.cfg, "params": synth.CONFIGS[a.cfg], "threads": a.threads, "host_cores": os.cpu_count(),
            "host_mem_gb": round(os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2**30, 1)}
    st = os.statvfs(os.path.dirname(a.workdir) or "/")
    info["disk_free_gb"] = round(st.f_bavail * st.f_frsize / 2**30, 1)
    print(info, flush=True)
    cfg = synth.CONFIGS[a.cfg]
    need_gb = (cfg["genome_len"] * cfg["coverage"] * 17.5 + cfg["genome_len"] * 60.0) / 2**30      # okmers + reads, kmers + reference tables
    need_disk = cfg["genome_len"] * cfg["coverage"] * 2.1 / 2**30 + cfg["genome_len"] * 3.5 / 2**30
    info["estimated_host_gb"], info["estimated_disk_gb"] = round(need_gb, 1), round(need_disk, 1)
    if need_gb > 0.8 * info["host_mem_gb"] or need_disk > 0.9 * info["disk_free_gb"]:
        print("NOT RUN: needs about %.0f GB of host memory and %.0f GB of disk; this box has %.0f / %.0f" % (need_gb, need_disk, info["host_mem_gb"], info["disk_free_gb"]))
        json.dump(info, open(a.out or os.path.join(ROOT, "gpurun_out", "cli_%s.json" % a.cfg), "w"), indent=1)
        return
    t0 = time.time()
    fa, fq, inp = synth.materialise(a.cfg, a.workdir)
    if inp is not None:
        info["reads"] = len(inp.reads); info["read_bases"] = int(sum(len(r) for r in inp.reads)); info["contigs"] = len(inp.contigs)
    info["generate_s"] = time.time() - t0
    info["fastq_gb"] = round(os.path.getsize(fq) / 2**30, 2)
    del inp
    gc.collect()
    print("generated in %.0f s" % info["generate_s"], flush=True)
    if a.golden:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "gc_e2e.json")))[a.golden]
        info["reference"] = dict(g, rc=0, wall_s=None, source="tests/golden/gc_e2e.json (the reference's own run, profiles/r02_cli_cfg5_full.json)")
    if a.runs_first and not a.skip_b200:
        env = dict(os.environ, GCG_TRACE="1", GC_RUNS="1")
        info["gc_b200_runs"] = run(os.path.join(ROOT, "superplus_b200", "_build", "gc_b200"), fa, fq, a.threads, os.path.join(a.workdir, "b200r"), env)
        print("gc_b200 GC_RUNS=1", {k: v for k, v in info["gc_b200_runs"].items() if k != "stderr_tail"}, flush=True)
        print(info["gc_b200_runs"]["stderr_tail"][-1200:])
        r, g = info.get("reference"), info["gc_b200_runs"]
        if r:
            info["runs_identical"] = bool(g["rc"] == 0 and all(r.get(k) == g.get(k) for k in ("fa", "link", "valid", "stats")))
            print("GC_RUNS=1:", "IDENTICAL" if info["runs_identical"] else "DIFFERENT", flush=True)
    if not a.skip_ref and not a.golden:
        info["reference"] = run(os.path.join(ROOT, "oracle", "_ref", "gc"), fa, fq, a.threads, os.path.join(a.workdir, "ref"), as_limit_gb=0.85 * info["host_mem_gb"])
        print("reference", {k: v for k, v in info["reference"].items() if k != "stderr_tail"}, flush=True)
    if not a.skip_b200:
        env = dict(os.environ, GCG_TRACE="1")
        if a.devices:
            env["GC_DEVICES"] = a.devices
        info["gc_b200"] = run(os.path.join(ROOT, "superplus_b200", "_build", "gc_b200"), fa, fq, a.threads, os.path.join(a.workdir, "b200"), env)
        info["gc_b200"]["devices"] = a.devices or "0"
        print("gc_b200", {k: v for k, v in info["gc_b200"].items() if k != "stderr_tail"}, flush=True)
        print(info["gc_b200"]["stderr_tail"][-1500:])
    if "reference" in info and "gc_b200" in info:
        r, g = info["reference"], info["gc_b200"]
        info["identical"] = bool(r["rc"] == 0 and g["rc"] == 0 and all(r.get(k) == g.get(k) for k in ("fa", "link", "valid", "stats")))
        info["speedup_wall"] = r["wall_s"] / g["wall_s"] if (g["wall_s"] and r.get("wall_s")) else None
        print("IDENTICAL" if info["identical"] else "DIFFERENT", "speed-up %.2fx" % (info["speedup_wall"] or 0), flush=True)
    out = a.out or os.path.join(ROOT, "gpurun_out", "cli_%s.json" % a.cfg)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(info, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
