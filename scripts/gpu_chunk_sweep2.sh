#!/usr/bin/env bash
# end-to-end k-mer step under fixed chunk sizes (bytes of bases): whole waves of the search kernel against powers of two
O=gpurun_out; mkdir -p $O
for b in "$@"; do
  GCG_TRACE=0 GCG_SEARCH_CHUNK_BYTES=$b python scripts/perf_e2e_trace.py cfg2 > $O/sweepb_$b.log 2>/dev/null
  echo "chunk $b: ascii $(grep '^ascii' $O/sweepb_$b.log | tail -3 | awk '{print $15}' | tr '\n' ' ') packed $(grep '^packed' $O/sweepb_$b.log | tail -3 | awk '{print $15}' | tr '\n' ' ')"
done
