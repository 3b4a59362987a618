#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider -x -k "bench_shape" 2>&1 | tail -15
