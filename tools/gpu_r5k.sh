#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-4} gpurun_out/$name.log; }
run r5k_tests_k 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider -x
MN=kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg
TAILN=1 run r5k_mnv3 600 python bench.py --cfg $MN --batch 64 --no-cpu-baseline --no-train-leg --steps 200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dwconv_tile|conv1x1_mma" -c 8 -f -o gpurun_out/r5k_mnv3_kernels python tools/prof_mnv3_kernels.py > gpurun_out/r5k_ncu.log 2>&1; tail -2 gpurun_out/r5k_ncu.log
python tools/ncu_summary.py gpurun_out/r5k_mnv3_kernels.ncu-rep > gpurun_out/r5k_ncu_summary.txt 2>&1
timeout 300 python tools/dw_bench.py 2>&1 | grep -v Summary | tee gpurun_out/r5k_dw_bench.txt
