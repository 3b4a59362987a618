#!/bin/bash
# r02 evidence: racecheck / memcheck after the __syncthreads fix, launch lists + DRAM traffic of the bench step, ncu full of the dominant kernels
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
K="conv_bn_act and (2x64x16x20 or 2x128x16x20 or 2x256x8x10 or 2x512x4x5 or 1x32x32x40 or 2x40x10 or 4x32x64x80 or 3x48x70)"
TAILN=12 run r3f_racecheck 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 1100 -p no:cacheprovider -k "$K or fused_into_consumer or squeeze or batched or weighted_fusion_and"
TAILN=8 run r3f_memcheck 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 800 -p no:cacheprovider
run r3f_ncu_dram 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r3f_dram_v3.csv python tools/one_forward.py kaist_dyolov3_add_sl.cfg 16 2
python tools/dram_summary.py gpurun_out/r3f_dram_v3.csv 11.04 > gpurun_out/r3f_dram_v3_summary.txt 2>&1; cat gpurun_out/r3f_dram_v3_summary.txt | head -30
run r3f_ncu_full 600 ncu --set full --clock-control none --import-source on -k regex:"halo2|conv_tc" -c 6 -f -o gpurun_out/r3f_prof python tools/prof_kernels.py
