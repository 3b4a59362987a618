#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2s_prof_1x1mish python tools/conv_bench.py --only 13 --iters 1 > gpurun_out/r2s_ncu.log 2>&1; tail -2 gpurun_out/r2s_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2s_prof_1x1_128 python tools/conv_bench.py --only 0 --iters 1 > gpurun_out/r2s_ncu2.log 2>&1; tail -2 gpurun_out/r2s_ncu2.log
