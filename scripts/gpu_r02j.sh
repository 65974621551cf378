set -u
O=gpurun_out; mkdir -p $O
GCG_LIB=$PWD/superplus_b200/_build/variants/libgcgpu_checked.so timeout 900 python -m pytest tests/test_kmer_gpu.py tests/test_part_gpu.py -x -q -m gpu > $O/pytest_r02j_checked.log 2>&1; echo "pytest checked rc=$?"; tail -2 $O/pytest_r02j_checked.log; grep -c violation $O/pytest_r02j_checked.log
timeout 900 python -m pytest tests/test_kmer_gpu.py tests/test_gc_e2e_gpu.py tests/test_scale_gpu.py -x -q -m gpu > $O/pytest_r02j.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_r02j.log
for v in p2 p2m4; do
  for cfg in "cfg2 25" "cfg4s 31"; do
    GCG_LIB=$PWD/superplus_b200/_build/variants/libgcgpu_$v.so timeout 300 python scripts/perf_kmer.py $cfg 4 > $O/var_${v}.log 2>&1
    echo "== $v $cfg rc=$? $(grep -A6 'iter 3' $O/var_${v}.log | grep 'k45_fused' | tr -s ' ')"
  done
done
KCMD="python scripts/perf_kmer.py cfg2 25 3"
ncu --set full --clock-control none --import-source on -k regex:k45_fused -s 2 -c 1 -o $O/prof_k45f_r02j -f $KCMD > $O/ncu_k45j.log 2>&1; echo "ncu rc=$?"
