# First GPU call of the next round: measures the experiments prepared (compile-checked, never run) at the end of round 1.
#   here:    make -C hkd-mpc_b200 experiments && (cd tools/microbench && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gj_variants gj_variants.cu)
#   then:    gpurun --timeout 900 -- 'bash tools/dev_round2_first.sh 2>&1 | tee gpurun_out/round2_first.log | tail -60'
P=$PWD/hkd-mpc_b200
echo "#### Gauss-Jordan variants (cycles per elimination; variant 6 = 4x4 pivots, must agree with variant 0)"
timeout 60 tools/microbench/gj_variants
for v in qxxp3 b3 b4 lrsw; do
  echo "#### parity tests with variant $v"
  HSDDP_LIB=$P/libhsddp_b200_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size" 2>&1 | tail -3
done
echo "#### throughput / latency A/B (- = default build)"
bash tools/dev_ab_lat.sh - qxxp3 b3 b4 lrsw sw6 sw4 2>&1 | grep -v "^$"
echo "#### 16 groups"
HSDDP_PHASED_GROUPS=16 python tools/profile_case.py 16384 config3 2 | tail -1
