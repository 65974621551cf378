#!/usr/bin/env bash
# end-to-end k-mer step (scripts/perf_e2e_trace.py) under several chunk schedules of the host pipeline
O=gpurun_out; mkdir -p $O
for spec in "${@:-4:32:8}"; do
  IFS=: read first max last <<< "$spec"
  GCG_TRACE=0 GCG_SEARCH_CHUNK_FIRST_MB=$first GCG_SEARCH_CHUNK_MAX_MB=$max GCG_SEARCH_CHUNK_LAST_MB=$last python scripts/perf_e2e_trace.py cfg2 > $O/sweep_$first-$max-$last.log 2>/dev/null
  echo "first $first max $max last $last: ascii $(grep '^ascii' $O/sweep_$first-$max-$last.log | tail -3 | awk '{print $15}' | tr '\n' ' ') packed $(grep '^packed' $O/sweep_$first-$max-$last.log | tail -3 | awk '{print $15}' | tr '\n' ' ')"
done
