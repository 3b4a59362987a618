#!/bin/bash
# Round-2 session C: fp32 mode (chunked accumulation), drift-vs-autocast test, wgrad cluster reduction (tests + train bench)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run r2c_f32 1200 python -m pytest tests/test_gpu_f32_mode.py -q -m gpu --timeout 600 -p no:cacheprovider
run r2c_drift_test 1200 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider -k reference_autocast
run r2c_train_kernels 1200 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_model.py -q -m gpu --timeout 600 -p no:cacheprovider
TAILN=3 run r2c_train 600 python bench.py --mode train --steps 20 --warmup 4
DYK_WG_CLUSTER=1 TAILN=3 run r2c_train_nocluster 600 python bench.py --mode train --steps 20 --warmup 4
DYK_WG_WAVES=1 TAILN=3 run r2c_train_1wave 600 python bench.py --mode train --steps 20 --warmup 4
TAILN=60 run r2c_drift 1200 python tools/drift_table.py gpurun_out/r2c_drift.json 2
