#!/bin/bash
mkdir -p gpurun_out
python tools/prof_stem.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_resize.py -q -m gpu --timeout 300 -p no:cacheprovider -x -k "stem or resize or input_size" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -p no:cacheprovider -x 2>&1 | tail -2
bench() { timeout 600 python bench.py "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), 'roofline', d.get('roofline',{}).get('frac'), d.get('clocks'))"; }
echo "dyolov3 fp16 bs16"; bench --steps 100 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0
