#!/bin/bash
# Round-2 session B: fp32-accurate mode tests, drift table (native vs reference autocast / TF32)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run r2b_f32 1200 python -m pytest tests/test_gpu_f32_mode.py -q -m gpu --timeout 600 -p no:cacheprovider
TAILN=60 run r2b_drift 1200 python tools/drift_table.py gpurun_out/r2b_drift.json 2
