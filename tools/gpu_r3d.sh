#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -x 2>&1 | tail -2
echo "--- default"; python tools/chain_bench.py 2>&1 | grep -v Summary | tee gpurun_out/r3d_chain.txt
echo "--- no resident"; DYK_HALO2_RES=0 python tools/chain_bench.py 2>&1 | grep -v Summary | tail -2
echo "--- no resident, PF=1"; DYK_RES_PF=1 DYK_HALO2_RES=0 python tools/chain_bench.py 2>&1 | grep -v Summary | tail -6
echo "--- resident, PF=1"; DYK_RES_PF=1 python tools/chain_bench.py 2>&1 | grep -v Summary | tail -2
