#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_train_model.py -q -m gpu --timeout 600 -p no:cacheprovider -x 2>&1 | tail -25
