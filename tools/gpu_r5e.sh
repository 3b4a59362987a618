#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r5e_sweep.jsonl
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "depthwise" 2>&1 | tail -5 | tee gpurun_out/r5e_tests.log
DYK_DW_TILE_DEBUG=1 timeout 300 python tools/dw_bench.py --modes 1 2>&1 | grep -v Summary | tee gpurun_out/r5e_dw_bench.txt
timeout 600 python tools/dw_bench.py --sweep --dump gpurun_out/r5e_sweep.jsonl 2>&1 | grep -v Summary | cut -c1-400 | tee gpurun_out/r5e_dw_sweep.txt
