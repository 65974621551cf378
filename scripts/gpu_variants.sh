#!/usr/bin/env bash
# time every variant library in superplus_b200/_build/variants on the cfg2 k-mer step (perf_kmer.py, last iteration)
set -u
O=gpurun_out; mkdir -p $O
CFG=${1:-cfg2}; K=${2:-25}
for lib in superplus_b200/_build/variants/libgcgpu_*.so; do
  n=$(basename $lib .so); n=${n#libgcgpu_}
  for load in ${LOADS:-50}; do
    GCG_TABLE_LOAD=$load GCG_LIB=$PWD/$lib timeout 300 python scripts/perf_kmer.py $CFG $K 4 > $O/var_${n}_$load.log 2>&1
    echo "== $n load=$load rc=$? $(grep -A6 'iter 3' $O/var_${n}_$load.log | grep 'k45_fused\|k23_build' | tr -s ' ' | tr '\n' ' ') $(grep 'compact' $O/var_${n}_$load.log | tail -1)"
  done
done
