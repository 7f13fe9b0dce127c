#!/bin/bash
# GPU session: cohort lock-step variants A/B (+ warp-state captures of two of them)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
python tools/variant_bench.py 131072 pitchy_pl 2
for v in w16 c16x8 c16x4 c16x2 c20x10 c20x5; do
  RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_$v.so timeout 300 python tools/variant_bench.py 131072 pitchy_pl 2 || echo "variant $v failed rc=$?"
done
} > $O/s3_variants.log 2>&1
for v in c16x8 c20x5; do
RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_$v.so timeout 600 ncu --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 -o $O/s3_${v}_warpstate -f python tools/profile_small.py 8192 0xFF > $O/s3_ncu_$v.log 2>&1
done
timeout 600 python -m pytest tests/test_crank_out.py tests/test_examples.py -m gpu -x -q > $O/s3_pytest_small.log 2>&1
cat $O/s3_variants.log
