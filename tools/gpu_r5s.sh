#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
python tools/layer_times.py kaist_dyolov3_add_sl.cfg 16 2>&1 | grep -E "total|maxpool"
python tools/layer_times.py kaist_dyolov4_mobilenetv3_fshare_global_cse3.cfg 64 2>&1 | grep -E "total|maxpool"
python bench.py --steps 20 --no-cpu-baseline --no-train-leg --sustain-s 0 2>&1 | tail -1 > gpurun_out/r5s_v3.log
python -c "
import json; d=json.loads(open('gpurun_out/r5s_v3.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'])"
