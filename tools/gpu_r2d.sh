#!/bin/bash
# Round-2 session D: fp32 mode with negated twin accumulators; wgrad kernel times with / without the in-cluster reduction
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run r2d_f32 1200 python -m pytest tests/test_gpu_f32_mode.py -q -m gpu --timeout 600 -p no:cacheprovider
grep -n "native'\|AssertionError: (" gpurun_out/r2d_f32.log | cut -c1-420
TAILN=40 run r2d_drift 1200 python tools/drift_table.py gpurun_out/r2d_drift.json 2
DYK_TRAIN_GRAPH=0 TAILN=2 run r2d_ncu_wg8 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wgrad -c 2000 --csv --log-file gpurun_out/r2d_wg_cluster8.csv python tools/train_once.py
DYK_WG_CLUSTER=1 DYK_TRAIN_GRAPH=0 TAILN=2 run r2d_ncu_wg1 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wgrad -c 2000 --csv --log-file gpurun_out/r2d_wg_cluster1.csv python tools/train_once.py
python tools/launch_summary.py gpurun_out/r2d_wg_cluster8.csv | head -5
python tools/launch_summary.py gpurun_out/r2d_wg_cluster1.csv | head -5
