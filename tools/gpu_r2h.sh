#!/bin/bash
mkdir -p gpurun_out
python tools/bn_bench.py mish 2>&1 | tee gpurun_out/r2h_bn_bench_mish.txt
python tools/bn_bench.py leaky 2>&1 | tee gpurun_out/r2h_bn_bench_leaky.txt
