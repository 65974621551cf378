#!/usr/bin/env bash
# one wave of the packed SW kernel under ncu (after the same command ran clean)
set -e
python scripts/perf_sw.py 5920 0 1 > gpurun_out/sw_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sw_fill_packed -c 1 -o gpurun_out/prof_sw_packed \
    python scripts/perf_sw.py 5920 0 1 > gpurun_out/sw_ncu.log 2>&1
tail -3 gpurun_out/sw_ncu.log
