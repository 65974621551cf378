"""Summarise the SASS page of an ncu report: instruction mix and where the warps stall.
usage: python scripts/ncu_sass_summary.py gpurun_out/prof.ncu-rep [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot_inst = sum(int(r["Instructions Executed"]) for r in rows)
tot_thr = sum(int(r["Thread Instructions Executed"]) for r in rows)
tot_samp = sum(int(r["# Samples"]) for r in rows)
print("kernel:", lines[0][:120])
print("SASS instructions %d ; warp-instr executed %d ; thread-instr %d ; avg active %.1f ; samples %d" % (len(rows), tot_inst, tot_thr, tot_thr / max(tot_inst, 1), tot_samp))
mix = collections.Counter(); mixs = collections.Counter()
for r in rows:
    op = r["Source"].split()[0] if not r["Source"].strip().startswith("@") else r["Source"].split()[1]
    op = op.split(".")[0]
    mix[op] += int(r["Instructions Executed"]); mixs[op] += int(r["# Samples"])
print("\nopcode mix (warp-instr share, stall-sample share):")
for op, n in mix.most_common(18):
    print("  %-10s %6.2f%%  %6.2f%%" % (op, 100 * n / tot_inst, 100 * mixs[op] / max(tot_samp, 1)))
stalls = [k for k in rows[0].keys() if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[k] or 0) for r in rows) for k in stalls}
print("\nstall reasons (all samples):")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print("  %-24s %6.2f%%" % (k, 100 * v / max(tot_samp, 1)))
print("\ntop %d instructions by samples:" % top)
for i, r in sorted(enumerate(rows), key=lambda ir: -int(ir[1]["# Samples"]))[:top]:
    print("  #%-4d %5.2f%%  exec %9s thr %4s  %s" % (i, 100 * int(r["# Samples"]) / max(tot_samp, 1), r["Instructions Executed"], r["Avg. Threads Executed"], r["Source"].strip()[:90]))
