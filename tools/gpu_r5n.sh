#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-4} gpurun_out/$name.log; }
run r5n_tests_k 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -x -k "conv_bn_act or se_gate"
MN=kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg
python tools/layer_times.py $MN 64 > gpurun_out/r5n_layer_times_mnv3.txt 2>&1; grep -E "total|'conv', (16|24|40|72|120), " gpurun_out/r5n_layer_times_mnv3.txt
TAILN=1 run r5n_mnv3 600 python bench.py --cfg $MN --batch 64 --no-cpu-baseline --no-train-leg --steps 200
