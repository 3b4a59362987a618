#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TAILN=12 run r2f_optim 600 python -m pytest tests/test_gpu_optim.py tests/test_gpu_dropin.py -q -m gpu --timeout 600 -p no:cacheprovider
DYK_TRAIN_TIMES_JSON=gpurun_out/r2f_train_times.json TAILN=30 run r2f_train_times 600 python tools/train_times.py
DYK_TRAIN_GRAPH=0 TAILN=2 run r2f_ncu_train 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2f_launches_train.csv python tools/train_once.py
python tools/launch_summary.py gpurun_out/r2f_launches_train.csv | head -24
