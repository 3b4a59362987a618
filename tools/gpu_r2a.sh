#!/bin/bash
# Round-2 session A: parity suite with the graph-replayed training step, train bench (graphs on / off), default bench line,
# per-op train timing, ncu launch list of one training step, compute-sanitizer memcheck over the kernel tests.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-8} gpurun_out/$name.log; }
run r2a_tests 1200 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider
run r2a_train 600 python bench.py --mode train --steps 20 --warmup 4
DYK_TRAIN_GRAPH=0 run r2a_train_eager 600 python bench.py --mode train --steps 20 --warmup 4
run r2a_bench 900 python bench.py --steps 20 --warmup 5
run r2a_train_times 600 python tools/train_times.py
DYK_TRAIN_GRAPH=0 run r2a_ncu_train 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2a_launches_train.csv python tools/train_once.py
run r2a_memcheck 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 900 -p no:cacheprovider
