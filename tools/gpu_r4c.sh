#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -x 2>&1 | tail -8
export DYK_CHAIN_ONLY=0,1,2,6,12,13,14
echo "--- multicast on"; python tools/chain_bench.py 2>&1 | grep -v Summary
echo "--- multicast off"; DYK_TC_MULTICAST=0 python tools/chain_bench.py 2>&1 | grep -v Summary
unset DYK_CHAIN_ONLY
for mc in 1 0; do
echo "--- conv_bench multicast=$mc"; DYK_TC_MULTICAST=$mc python tools/conv_bench.py --only 6,7,10,13 --iters 10 2>&1 | grep -v Summary
done
