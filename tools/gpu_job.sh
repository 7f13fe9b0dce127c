set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python tools/profile_small.py 4096 0xFF > gpurun_out/plain_fast.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 -o gpurun_out/prof_fast -f python tools/profile_small.py 4096 0xFF > gpurun_out/ncu_fast.log 2>&1
python bench.py --points 32768 --steps 2 --warmup 3 > gpurun_out/bench3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --points 32768 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/plain_fast.log; cut -c1-1500 gpurun_out/bench3.log
