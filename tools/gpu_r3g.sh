#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resize.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -30
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train_model.py -q -m gpu --timeout 600 -p no:cacheprovider -x 2>&1 | tail -3
