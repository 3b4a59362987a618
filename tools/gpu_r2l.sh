#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TAILN=4 run r2l_tests 1200 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_model.py -q -m gpu --timeout 600 -p no:cacheprovider
for sp in 512 256 128; do
  DYK_SLAB_PIX=$sp timeout 600 python bench.py --mode train --steps 20 --warmup 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('slab_pix $sp ms_per_step', d['ms_per_step'])"
done
