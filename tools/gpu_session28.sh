#!/bin/bash
# faithful continuation: one warp per (handed point, accumulator)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/kappa_check.py > $O/s28_kappa.log 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "kappa or reroute or handed or guard or fast_mode or faithful" > $O/s28_pytest.log 2>&1; echo "rc=$?" >> $O/s28_pytest.log
timeout 400 python bench.py --config pitchy_kappa --points 32768 --steps 1 --warmup 1 > $O/s28_bench_c4.json 2> $O/s28_bench_c4.err
cat $O/s28_kappa.log; tail -3 $O/s28_pytest.log; python -c "
import json; d=json.load(open('$O/s28_bench_c4.json')); print('c4', round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['parity']['within_1e-3'], d['parity']['nan_mismatch'])"
