"""Turn the scratch ncu outputs of one GPU visit (gpurun_out/) into the committed evidence under
profiles/:  launch list summary (per-kernel share of the step), one key-metric sheet + SASS
opcode/stall summary per `--set full` capture, and roofline_traffic.json (per-launch DRAM bytes
that bench.py reports as roofline.traffic).

usage: python scripts/ncu_profiles.py <tag> [round-label]     e.g.  r01b r01
"""
import collections
import re
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEY_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    # atomics (table build: atomicCAS claims, atomicOr flags; emit: the ONT-count atomicOr)
    "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_atom.sum",
    "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_atom.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
    "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum",
    "lts__t_requests_srcunit_tex_op_atom_dot_alu.sum", "lts__t_requests_srcunit_tex_op_atom_dot_cas.sum", "lts__t_requests_srcunit_tex_op_red.sum",
    "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_atom.sum.per_second",
    "lts__t_sectors_srcunit_tex_op_atom.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_atom_dot_cas.sum", "lts__t_sectors_srcunit_tex_op_atom_dot_cas.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum.per_second",
]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict((h, (v, u)) for h, u, v in zip(hdr, units, r)) for r in rows[2:]]


def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


def sheet(rep, title, fh):
    traffic = {}
    for k in raw_page(rep):
        name = k["Kernel Name"][0]
        short = name.split("(")[0]
        fh.write("== %s ==\nkernel: %s\nreport: %s (ncu --set full --clock-control none --import-source on)\n" % (title, name, os.path.basename(rep)))
        for m in KEY_METRICS:
            if m in k:
                fh.write("  %-72s %16s %s\n" % (m, k[m][0], k[m][1]))
        rd, wr = to_bytes(*k["dram__bytes_read.sum"]), to_bytes(*k["dram__bytes_write.sum"])
        dur_us = float(k["gpu__time_duration.sum"][0].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(k["gpu__time_duration.sum"][1], 1)
        fh.write("  -> DRAM traffic per launch %.1f MB (read %.1f + write %.1f) ; %.1f GB/s under the profiler's cold-cache replay\n" %
                 ((rd + wr) / 1e6, rd / 1e6, wr / 1e6, (rd + wr) / dur_us / 1e3))
        atom = sum(float(k[m][0].replace(",", "")) for m in ("l1tex__m_l1tex2xbar_write_sectors_mem_global_op_atom.sum",
                                                              "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum") if m in k)
        if atom:
            fh.write("  -> global atomic / reduction sectors sent to the L2: %.2f M per launch, %.1f G sectors/s\n" % (atom / 1e6, atom / dur_us / 1e3))
        traffic[short] = rd + wr
    return traffic


def sass_summary(rep, fh, top=16):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
    lines = out.splitlines()
    try:
        start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    except StopIteration:
        return
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    tot_inst = sum(int(r["Instructions Executed"]) for r in rows) or 1
    tot_samp = sum(int(r["# Samples"]) for r in rows) or 1
    mix, mixs = collections.Counter(), collections.Counter()
    for r in rows:
        toks = r["Source"].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        mix[op] += int(r["Instructions Executed"]); mixs[op] += int(r["# Samples"])
    fh.write("SASS: %d instructions, %d warp-instructions executed, %d stall samples\n" % (len(rows), tot_inst, tot_samp))
    fh.write("opcode mix (share of executed warp-instructions / share of stall samples):\n")
    for op, n in mix.most_common(14):
        fh.write("  %-10s %6.2f%%  %6.2f%%\n" % (op, 100 * n / tot_inst, 100 * mixs[op] / tot_samp))
    stalls = [k for k in rows[0].keys() if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[k] or 0) for r in rows) for k in stalls}
    fh.write("stall reasons (share of samples):\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        fh.write("  %-26s %6.2f%%\n" % (k, 100 * v / tot_samp))
    fh.write("hottest instructions:\n")
    for i, r in sorted(enumerate(rows), key=lambda ir: -int(ir[1]["# Samples"]))[:top]:
        fh.write("  #%-4d %5.2f%%  %s\n" % (i, 100 * int(r["# Samples"]) / tot_samp, r["Source"].strip()[:100]))
    fh.write("\n")


def launch_list(path, fh):
    txt = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(txt))))
    agg, n = collections.OrderedDict(), collections.Counter()
    for r in rows:
        k = r["Kernel Name"].split("(")[0].replace("void ", "")
        agg[k] = agg.get(k, 0.0) + float(r["Metric Value"]) / 1e3
        n[k] += 1
    tot = sum(agg.values()) or 1
    fh.write("launch list: %s  (%d launches; gpu__time_duration.sum, --clock-control none; serialised, cold caches:\n"
             "             shares are meaningful, absolute times are not bench values)\n" % (os.path.basename(path), len(rows)))
    fh.write("  %-28s %8s %12s %10s %8s\n" % ("kernel", "launches", "total us", "avg us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        fh.write("  %-28s %8d %12.1f %10.1f %7.1f%%\n" % (k, n[k], v, v / n[k], 100 * v / tot))
    # (launch_probe_kernel: the streaming search's one-off test for synchronous launches — under ncu it runs into its 20 ms
    #  timeout, which is how the library knows to keep one launch per chunk; post_count_kernel: the chunked pipeline's
    #  completion post; both belong to the end-to-end legs, not to the device-resident step)
    kmer = {k: v for k, v in agg.items() if not k.startswith("sw_") and "ubench" not in k and "launch_probe" not in k and "post_count" not in k}
    kt = sum(kmer.values()) or 1
    fh.write("k-mer step only (k1 / k23 / k45 / k6; runs_* belong to the e2e_runs leg; under ncu the end-to-end legs run the chunked\n"
             "pipeline, because launches are synchronous there — launch_probe_kernel finds that out):\n")
    for k, v in sorted(kmer.items(), key=lambda kv: -kv[1]):
        fh.write("  %-28s %7.1f%%   avg %8.1f us\n" % (k, 100 * v / kt, v / n[k]))
    fh.write("\n")


SW_CAPTURE_PAIRS = 5920


def main():
    tag = sys.argv[1]
    label = sys.argv[2] if len(sys.argv) > 2 else tag
    os.makedirs(PROF, exist_ok=True)
    traffic = {}
    tpath = os.path.join(PROF, "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = {k: v for k, v in json.load(open(tpath)).items() if "<" not in k}
    with open(os.path.join(PROF, "%s_ncu_summary.txt" % label), "w") as fh:
        fh.write("ncu evidence of round %s (captures tagged %s); regenerate with scripts/ncu_profiles.py\n\n" % (label, tag))
        ll = os.path.join(OUT, "launches_%s.csv" % tag)
        if os.path.exists(ll):
            launch_list(ll, fh)
        for f in sorted(os.listdir(OUT)):
            if f.endswith("_%s.ncu-rep" % tag):
                rep = os.path.join(OUT, f)
                t = sheet(rep, f[: -len(".ncu-rep")], fh)
                # plain kernel names ("void k<0>(...)" -> "k"); the SW capture is one launch over
                # SW_CAPTURE_PAIRS pairs (gpu_round.sh), bench.py scales it to the pairs of its own launch
                for name, v in t.items():
                    name = re.sub(r"<.*", "", name.replace("void ", "")).strip()
                    traffic[name] = v
                    if name == "sw_fill_packed_kernel":
                        traffic["sw_fill_packed_kernel_pairs"] = SW_CAPTURE_PAIRS
                sass_summary(rep, fh)
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
    print(open(os.path.join(PROF, "%s_ncu_summary.txt" % label)).read())


if __name__ == "__main__":
    main()
