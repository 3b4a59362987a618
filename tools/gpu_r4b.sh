#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-4} gpurun_out/$name.log; }
run r4b_tests 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider
run r4b_smoke 300 python __graft_entry__.py smoke
run r4b_ref 600 python bench.py --impl reference --steps 3 --warmup 1
TAILN=3 run r4b_smoke_ncu 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r4b_smoke_launches.csv python __graft_entry__.py smoke
python tools/launch_summary.py gpurun_out/r4b_smoke_launches.csv > gpurun_out/r4b_smoke_launches_summary.txt 2>&1; head -12 gpurun_out/r4b_smoke_launches_summary.txt
