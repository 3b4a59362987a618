#!/bin/bash
# One GPU-box session: per-kernel parity, NMS, whole model, smoke, short bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-15} gpurun_out/$name.log; }
run t_kernels 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 120 -p no:cacheprovider
run t_nms 300 python -m pytest tests/test_gpu_nms.py -q -m gpu --timeout 120 -p no:cacheprovider
run t_model 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 -p no:cacheprovider
run smoke 300 python __graft_entry__.py smoke
run bench 600 python bench.py --steps 10 --warmup 3
