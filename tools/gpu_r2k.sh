#!/bin/bash
mkdir -p gpurun_out
DYK_TRAIN_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2k_launches_train.csv python tools/train_once.py > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2k_launches_train.csv 2>/dev/null | head -30
DYK_WG_SIDE=0 DYK_TRAIN_TIMES_JSON=gpurun_out/r2k_train_times.json timeout 600 python tools/train_times.py 2>&1 | grep -E "conv blocks|  (fwd|bwd) +[0-9]+x|forward "
