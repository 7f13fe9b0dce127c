set -x
export FAST_CHECK_ONLY_THROUGHPUT=1
python tools/fast_check.py 0xFF 131072 > gpurun_out/ab3.log 2>&1
python tools/fast_check.py 0xFF 131072 > gpurun_out/ab3b.log 2>&1
grep throughput gpurun_out/ab3.log gpurun_out/ab3b.log | sed 's/.*throughput pitchy_pl n=131072 //' | cut -c1-110
