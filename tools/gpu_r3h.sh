#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=25 run r3h_kernels 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -x
echo "--- K-split on"; python tools/chain_bench.py 2>&1 | grep -v Summary | sed -n 4,6p | tee gpurun_out/r3h_chain.txt
echo "--- K-split off"; DYK_CHAIN_WS=0 python tools/chain_bench.py 2>&1 | grep -v Summary | sed -n 4,6p
TAILN=15 run r3h_model 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_dropin.py -q -m gpu --timeout 600 -p no:cacheprovider -x
run r3h_smoke 300 python __graft_entry__.py smoke
bench() { timeout 600 python bench.py "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), 'roofline', d.get('roofline',{}).get('frac'), d.get('clocks'))"; }
echo "dyolov3 fp16 bs16"; bench --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
echo "dyolov3 fp16 bs16 no ksplit"; DYK_H2_KSPLIT=0 bench --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
echo "dyolov4 fp16 bs16"; bench --cfg kaist_dyolov4_fshare_global_concat_se3.cfg --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
