#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=15 run r2x_kernels 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -x
TAILN=15 run r2x_model 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider -x
run r2x_smoke 300 python __graft_entry__.py smoke
python tools/chain_bench.py 2>&1 | grep -v Summary | tee gpurun_out/r2x_chain.txt
bench() { timeout 600 python bench.py "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), 'roofline', d.get('roofline',{}).get('frac'))"; }
echo "dyolov3 fp16 bs16"; bench --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
echo "dyolov3 fp16 bs16 no fused add"; DYK_FUSE_ADD=0 bench --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
