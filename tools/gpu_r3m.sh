#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -x 2>&1 | tail -2
python tools/chain_bench.py 2>&1 | grep -v Summary | tee gpurun_out/r3m_chain.txt
