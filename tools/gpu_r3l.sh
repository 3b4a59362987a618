#!/bin/bash
mkdir -p gpurun_out
DYK_B200_LIB=$PWD/double-yolo-kaist_b200/libdyk_b200_prof.so python tools/halo2_prof.py 2>&1 | grep -v Summary | tee gpurun_out/r3l_halo2_prof.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -x 2>&1 | tail -2
