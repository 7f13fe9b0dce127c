#!/bin/bash
# GPU session: lock-step variants A/B, then the GPU test-suite and the fixture check of the head build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python tools/variant_bench.py 131072 pitchy_pl 2
for v in w20 ls20 ls20g ls16 ls12 ls12g ls8x2; do
  RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_$v.so timeout 300 python tools/variant_bench.py 131072 pitchy_pl 2 || echo "variant $v failed rc=$?"
done
} > gpurun_out/s1_variants.log 2>&1
timeout 900 python tools/fast_check.py 0xFF 131072 > gpurun_out/s1_fast_check.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1
tail -5 gpurun_out/s1_pytest.log
cat gpurun_out/s1_variants.log
