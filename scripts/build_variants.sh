#!/usr/bin/env bash
# A/B builds of libgcgpu.so with the build-time switches of the search kernel (kmer.cu / kmer_dev.cuh):
#   scripts/build_variants.sh name "-DK45F_MINB=8 -DK45F_PREFETCH=0" [name2 "flags2" ...]
# -> superplus_b200/_build/variants/libgcgpu_<name>.so ; select at run time with GCG_LIB=<path>.
set -euo pipefail
cd "$(dirname "$0")/../superplus_b200/csrc"
make -s
OUT=../_build/variants; mkdir -p $OUT
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  for f in kmer part runs sw; do
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v $flags -c $f.cu -o $OUT/${f}_$name.o 2> $OUT/${f}_$name.ptxas.log
  done
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libgcgpu_$name.so _obj/ctx.o $OUT/kmer_$name.o $OUT/part_$name.o $OUT/runs_$name.o $OUT/sw_$name.o _obj/ubench.o _obj/host_par.o -cudart static
  echo "$name: $(grep -A2 'k45_fused_kernelILb0ELi25ELi1' $OUT/kmer_$name.ptxas.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')"
  rm -f $OUT/kmer_$name.o $OUT/part_$name.o $OUT/runs_$name.o $OUT/sw_$name.o
done
