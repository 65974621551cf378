set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_kmer_gpu.py -x -q -m gpu > $O/pytest_r02i.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_r02i.log
GCG_LIB=$PWD/superplus_b200/_build/variants/libgcgpu_checked.so timeout 600 python -m pytest tests/test_kmer_gpu.py -x -q -m gpu > $O/pytest_r02i_checked.log 2>&1; echo "pytest checked rc=$?"; tail -2 $O/pytest_r02i_checked.log
for v in b1 p2 p2m5 p2m4; do
  for cfg in "cfg2 25" "cfg4s 31"; do
    GCG_LIB=$PWD/superplus_b200/_build/variants/libgcgpu_$v.so timeout 300 python scripts/perf_kmer.py $cfg 4 > $O/var_${v}.log 2>&1
    echo "== $v $cfg rc=$? $(grep -A6 'iter 3' $O/var_${v}.log | grep 'k45_fused' | tr -s ' ')"
  done
done
