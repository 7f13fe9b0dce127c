set -x
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu5.log
python tools/fast_check.py 0xFF 131072 > gpurun_out/fast6.log 2>&1
python bench.py --points 131072 --steps 2 --warmup 3 > gpurun_out/bench4.log 2>&1
tail -6 gpurun_out/pytest_gpu5.log | cut -c1-250; grep throughput gpurun_out/fast6.log; cut -c1-2500 gpurun_out/bench4.log
