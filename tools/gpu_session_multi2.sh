#!/bin/bash
# Multi-GPU session of the final round-2 build (run with gpurun --gpus 8): the real sharding path and weak scaling
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/m2_gpus.log 2>&1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_entry or current_device or async_calls" > $O/m2_pytest.log 2>&1; echo "rc=$?" >> $O/m2_pytest.log
timeout 200 python bench.py --single-process --gpus 8 --points 10000000 --steps 1 --warmup 1 --no-parity > $O/m2_bench_sp8_10M.json 2> $O/m2_bench_sp8_10M.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --no-parity > $O/m2_bench_weak8.json 2> $O/m2_bench_weak8.err
tail -3 $O/m2_pytest.log; for f in weak8 sp8_10M; do python -c "
import json,sys
try:
    d=json.load(open('$O/m2_bench_$f.json')); print('$f', d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['scaling'])
except Exception as e: print('$f failed', e)
"; done
