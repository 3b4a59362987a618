#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-4} gpurun_out/$name.log; }
run r5o_tests 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider
run r5o_smoke 300 python __graft_entry__.py smoke
timeout 300 python tools/dw_bench.py 2>&1 | grep -v Summary | tee gpurun_out/r5o_dw_bench.txt
MN=kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg
TAILN=1 run r5o_mnv3 600 python bench.py --cfg $MN --batch 64 --no-cpu-baseline --no-train-leg --steps 200
TAILN=1 run r5o_v4 600 python bench.py --cfg kaist_dyolov4_fshare_global_concat_se3.cfg --batch 16 --no-cpu-baseline --no-train-leg --steps 100
python tools/layer_times.py $MN 64 > gpurun_out/r5o_layer_times_mnv3.txt 2>&1
