#!/bin/bash
mkdir -p gpurun_out
python tools/conv_bench.py --only 13,9,8,0,7 --iters 10 2>&1 | grep -v Summary | tee gpurun_out/r2o_convbench.txt
DYK_B200_LIB=$PWD/double-yolo-kaist_b200/libdyk_b200_prof.so python tools/conv_bench.py --only 13,9,8,0,7 --iters 3 2>&1 | grep -v Summary | tee gpurun_out/r2o_convprof.txt
for sp in 2 4; do echo "DYK_EPI_SPLIT=$sp"; DYK_EPI_SPLIT=$sp python tools/conv_bench.py --only 13,9 --iters 10 2>&1 | grep -v Summary; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2o_prof_1x1mish python tools/conv_bench.py --only 13 --iters 1 > gpurun_out/r2o_ncu.log 2>&1; tail -3 gpurun_out/r2o_ncu.log
