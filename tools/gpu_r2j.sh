#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TAILN=6 run r2j_tests 1800 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x
python tools/bn_bench.py mish 2>&1 | grep -v Summary
python tools/bn_bench.py leaky 2>&1 | grep -v Summary
for f in 1 0; do
  DYK_BN_FUSED_FINALIZE=$f timeout 600 python bench.py --mode train --steps 20 --warmup 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fused_finalize $f ms_per_step', d['ms_per_step'], d['gpu_launches'])"
done
echo "=== MNv3 bs64 inference"
timeout 600 python bench.py --cfg kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg --batch 64 --dtype fp16 --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MNv3 bs64 fps', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])"
echo "=== dyolov4 bs16 inference"
timeout 600 python bench.py --cfg kaist_dyolov4_fshare_global_concat_se3.cfg --batch 16 --dtype fp16 --steps 50 --warmup 5 --no-cpu-baseline --no-train-leg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dyolov4 bs16 fps', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])"
