#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TAILN=6 run r2i_tests 1200 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_model.py tests/test_gpu_dropin.py -q -m gpu --timeout 600 -p no:cacheprovider
for v in "2 0 8" "4 0 8" "2 1 8" "4 1 8" "4 0 16" "4 1 16" "4 1 32"; do
  set -- $v
  echo "=== BN variant inflight=$1 contig=$2 blocks/SM=$3"
  DYK_BN_INFLIGHT=$1 DYK_BN_CONTIG=$2 DYK_BN_BLOCKS=$3 python tools/bn_bench.py mish 2>&1 | grep -v Summary
done
echo "=== train: fused finalize on / off"
for f in 1 0; do
  DYK_BN_FUSED_FINALIZE=$f timeout 600 python bench.py --mode train --steps 20 --warmup 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fused_finalize $f ms_per_step', d['ms_per_step'], d['gpu_launches'])"
done
DYK_TORCH_OPT=1 timeout 600 python bench.py --mode train --steps 20 --warmup 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('torch SGD ms_per_step', d['ms_per_step'])"
