#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-4} gpurun_out/$name.log; }
run r3o_kernels 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -x
python tools/chain_bench.py 2>&1 | grep -v Summary | tee gpurun_out/r3o_chain.txt
python tools/conv_bench.py --only 13,12,9,8 --iters 10 2>&1 | grep -v Summary
bench() { timeout 600 python bench.py "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), 'roofline', d.get('roofline',{}).get('frac'), d.get('clocks'))"; }
echo "dyolov3 fp16 bs16"; bench --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
echo "dyolov4 fp16 bs16"; bench --cfg kaist_dyolov4_fshare_global_concat_se3.cfg --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
echo "mnv3 fp16 bs64"; bench --cfg kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg --batch 64 --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
echo "train"; bench --mode train --steps 20 --warmup 4
