#!/bin/bash
# Round-2 session E: fp32 mode (main / correction accumulators, 8-MMA chunks), fused optimizer tests, wgrad cluster / wave sweep
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TAILN=12 run r2e_f32 1200 python -m pytest tests/test_gpu_f32_mode.py -q -m gpu --timeout 600 -p no:cacheprovider
grep -n "native'\|AssertionError: (" gpurun_out/r2e_f32.log | cut -c1-420
TAILN=12 run r2e_optim 600 python -m pytest tests/test_gpu_optim.py -q -m gpu --timeout 600 -p no:cacheprovider
TAILN=40 run r2e_drift 1200 python tools/drift_table.py gpurun_out/r2e_drift.json 2
for cfg in "1 2" "1 1" "2 2" "2 1" "4 2" "4 1" "8 2"; do
  set -- $cfg
  echo "=== wgrad cluster $1 waves $2"
  DYK_WG_CLUSTER=$1 DYK_WG_WAVES=$2 timeout 600 python bench.py --mode train --steps 20 --warmup 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'])"
done
