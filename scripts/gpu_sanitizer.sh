#!/usr/bin/env bash
# compute-sanitizer evidence: memcheck and racecheck over the smoke path (table build with lock-free CAS
# inserts, the one-pass search with its chained scan and shared-memory dealing, SW fill + traceback that re-reads
# trace words the same warp wrote, CIGAR pool reservations) and over the partitioned index with two contexts
# (route / insert / lookup / collect kernels, windows written by a peer context).  Logs -> gpurun_out/sanitizer_*.log
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-r02}
S="compute-sanitizer --error-exitcode 9 --print-limit 20"
for tool in memcheck racecheck; do
  timeout 1500 $S --tool $tool python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_${tool}_smoke_$TAG.log 2>&1
  echo "$tool smoke rc=$? : $(grep -c 'ERROR SUMMARY' $O/sanitizer_${tool}_smoke_$TAG.log) summaries: $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $O/sanitizer_${tool}_smoke_$TAG.log | tail -1)"
  timeout 1500 $S --tool $tool python -m pytest tests/test_kmer_gpu.py -x -q -m gpu -k "run_records or compact_form or edge_cases or result_buffer" > $O/sanitizer_${tool}_kmer_$TAG.log 2>&1
  echo "$tool kmer tests rc=$? : $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed' $O/sanitizer_${tool}_kmer_$TAG.log | tail -2 | tr '\n' ' ')"
  timeout 1500 $S --tool $tool python -m pytest tests/test_part_gpu.py -x -q -m gpu -k "owner_side or tiny-25-2147483648-2-" > $O/sanitizer_${tool}_part_$TAG.log 2>&1
  echo "$tool partitioned rc=$? : $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed' $O/sanitizer_${tool}_part_$TAG.log | tail -2 | tr '\n' ' ')"
done
