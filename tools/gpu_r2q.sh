#!/bin/bash
mkdir -p gpurun_out
T='tests/test_gpu_dropin.py::test_reference_training_loop_over_native_modules'
for i in 1 2 3; do timeout 300 python -m pytest "$T" -q -m gpu -p no:cacheprovider -x -k "fused_sgd and dyolov3" 2>&1 | tail -1; done
echo "--- graphs off"; DYK_TRAIN_GRAPH=0 timeout 300 python -m pytest "$T" -q -m gpu -p no:cacheprovider -x -k "fused_sgd and dyolov3" 2>&1 | tail -1
echo "--- lanes 1"; DYK_TRAIN_LANES=1 timeout 300 python -m pytest "$T" -q -m gpu -p no:cacheprovider -x -k "fused_sgd and dyolov3" 2>&1 | tail -1
echo "--- wg side 0"; DYK_WG_SIDE=0 timeout 300 python -m pytest "$T" -q -m gpu -p no:cacheprovider -x -k "fused_sgd and dyolov3" 2>&1 | tail -1
echo "--- all dropin"; timeout 600 python -m pytest tests/test_gpu_dropin.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
echo "--- memcheck"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "$T" -q -m gpu -p no:cacheprovider -x -k "fused_sgd and dyolov3" > gpurun_out/r2q_memcheck.log 2>&1; echo "exit $?"; grep -m1 -A25 "Invalid\|========= Error\|=========     at" gpurun_out/r2q_memcheck.log | head -60; tail -5 gpurun_out/r2q_memcheck.log
