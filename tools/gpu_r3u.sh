#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -x 2>&1 | tail -2
python tools/chain_bench.py 2>&1 | grep -v Summary | head -3
python tools/conv_bench.py --only 13,9,8,0,7,10 --iters 10 2>&1 | grep -v Summary
bench() { timeout 600 python bench.py "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), 'roofline', d.get('roofline',{}).get('frac'), d.get('clocks'))"; }
echo "dyolov4 fp16 bs16"; bench --cfg kaist_dyolov4_fshare_global_concat_se3.cfg --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
echo "dyolov3 fp16 bs16"; bench --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
