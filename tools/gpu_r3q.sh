#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3q_bench.log 2>&1; echo "exit $?"
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r3q_bench.log') if l.startswith('{\"metric\"')][-1]
r=d['roofline']
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'frac',r['frac'],'kernel_ms',r['kernel_ms_per_step'],'all',r['all_dense_convs']['ms_per_step'],r['all_dense_convs']['frac'],'other',r['other_kernels_ms_per_step'],'train',d['train']['ms_per_step'],d['clocks'])
"
