#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:halo2 -c 1 -s 2 -f -o gpurun_out/r3s_halo2_big python tools/conv_bench.py --only 2 --iters 2 > gpurun_out/r3s_ncu.log 2>&1; tail -2 gpurun_out/r3s_ncu.log
