set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu7.log
python bench.py > gpurun_out/bench7.log 2>&1
FAST_CHECK_ONLY_THROUGHPUT=1 python tools/fast_check.py 0xFF 131072 > gpurun_out/fast7.log 2>&1
python tools/fast_check.py 0xC0 1024 > gpurun_out/fast7_hey.log 2>&1
tail -5 gpurun_out/pytest_gpu7.log | cut -c1-250; cut -c1-2600 gpurun_out/bench7.log; grep throughput gpurun_out/fast7.log | cut -c1-200; grep -A2 "pitchy_pl_4k\|juettner" gpurun_out/fast7_hey.log | cut -c1-260
