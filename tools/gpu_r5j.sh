#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-4} gpurun_out/$name.log; }
run r5j_tests 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider
run r5j_smoke 300 python __graft_entry__.py smoke
MN=kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg
TAILN=1 run r5j_bench 900 python bench.py
TAILN=1 run r5j_mnv3 600 python bench.py --cfg $MN --batch 64 --no-cpu-baseline --no-train-leg --steps 200
TAILN=3 run r5j_mnv3_ncu 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r5j_mnv3_launches.csv python bench.py --cfg $MN --batch 64 --no-cpu-baseline --no-train-leg --steps 2 --warmup 3 --sustain-s 0
python tools/launch_summary.py gpurun_out/r5j_mnv3_launches.csv > gpurun_out/r5j_mnv3_launches_summary.txt 2>&1; head -14 gpurun_out/r5j_mnv3_launches_summary.txt
