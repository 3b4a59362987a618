#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TAILN=15 run r2m_tests 1200 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_model.py tests/test_gpu_dropin.py -q -m gpu --timeout 600 -p no:cacheprovider
for l in 2 1; do
  DYK_TRAIN_LANES=$l timeout 600 python bench.py --mode train --steps 20 --warmup 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes $l ms_per_step', d['ms_per_step'], d['cuda_graphs'])"
done
