#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 1 -s 3 -f -o gpurun_out/r3p_mish1x1 python tools/conv_bench.py --only 13 --iters 2 > gpurun_out/r3p_ncu.log 2>&1; tail -2 gpurun_out/r3p_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 1 -s 3 -f -o gpurun_out/r3p_1x1_512 python tools/conv_bench.py --only 7 --iters 2 > gpurun_out/r3p_ncu2.log 2>&1; tail -2 gpurun_out/r3p_ncu2.log
