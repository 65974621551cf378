#!/usr/bin/env bash
# round 2, visit A: host facts, the new tests first (fail fast), whole GPU suite, perf A/B of the search forms
set -u
O=gpurun_out; mkdir -p $O
nproc > $O/host_r02.txt; free -g >> $O/host_r02.txt; nvidia-smi -L >> $O/host_r02.txt; lscpu | head -20 >> $O/host_r02.txt
timeout 900 python -m pytest tests/test_kmer_gpu.py -x -q -m gpu > $O/pytest_kmer_r02a.log 2>&1; echo "pytest kmer rc=$?"; tail -5 $O/pytest_kmer_r02a.log
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_kmer_gpu.py > $O/pytest_gpu_r02a.log 2>&1; echo "pytest rest rc=$?"; tail -5 $O/pytest_gpu_r02a.log
timeout 300 python scripts/perf_kmer.py cfg2 25 4 > $O/perf_kmer_fused_r02a.log 2>&1; echo "perf fused rc=$?"; tail -22 $O/perf_kmer_fused_r02a.log
GCG_SEARCH_FUSED=0 timeout 300 python scripts/perf_kmer.py cfg2 25 3 > $O/perf_kmer_twopass_r02a.log 2>&1; echo "perf twopass rc=$?"; grep -A12 "iter 2" $O/perf_kmer_twopass_r02a.log
