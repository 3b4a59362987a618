#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider -x 2>&1 | tail -2
bench() { timeout 600 python bench.py "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), d.get('clocks'))"; }
for split in 1 0 1 0; do
echo "dyolov3 split=$split"; DYK_SM_SPLIT=$split bench --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
done
for split in 1 0; do
echo "dyolov4 split=$split"; DYK_SM_SPLIT=$split bench --cfg kaist_dyolov4_fshare_global_concat_se3.cfg --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
echo "mnv3 split=$split"; DYK_SM_SPLIT=$split bench --cfg kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg --batch 64 --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
done
