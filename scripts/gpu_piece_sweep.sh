#!/usr/bin/env bash
# end-to-end k-mer step (scripts/perf_e2e_trace.py) under several piece schedules of the streaming search: first:max:last MiB
O=gpurun_out; mkdir -p $O
for spec in "${@:-2:8:2}"; do
  IFS=: read first max last <<< "$spec"
  GCG_SEARCH_PIECE_FIRST_MB=$first GCG_SEARCH_PIECE_MB=$max GCG_SEARCH_PIECE_LAST_MB=$last python scripts/perf_e2e_trace.py cfg2 > $O/psweep_$first-$max-$last.log 2>/dev/null
  echo "first $first max $max last $last: ascii $(grep '^ascii' $O/psweep_$first-$max-$last.log | tail -3 | awk '{print $15}' | tr '\n' ' ') packed $(grep '^packed' $O/psweep_$first-$max-$last.log | tail -3 | awk '{print $15}' | tr '\n' ' ')"
done
