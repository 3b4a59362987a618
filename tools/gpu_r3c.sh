#!/bin/bash
mkdir -p gpurun_out
bench() { timeout 600 python bench.py "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), 'roofline', d.get('roofline',{}).get('frac'), d.get('clocks'))"; }
echo "dyolov3 fp16 bs16 steps 100"; bench --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
echo "dyolov3 fp16 bs16 steps 20"; bench --steps 20 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
echo "dyolov3 fp16 bs16 steps 20 sustain"; bench --steps 20 --warmup 5 --no-cpu-baseline --no-train-leg
echo "dyolov4 fp16 bs16"; bench --cfg kaist_dyolov4_fshare_global_concat_se3.cfg --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
python tools/layer_times.py kaist_dyolov3_add_sl.cfg 16 gpurun_out/r3c_layers_v3.json 2>&1 | grep -v Summary | head -24
