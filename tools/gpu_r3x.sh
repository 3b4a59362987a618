#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r3x_bench_n8.log 2> gpurun_out/r3x_bench_n8.err; echo "exit $?"
grep '^{"metric"' gpurun_out/r3x_bench_n8.log > gpurun_out/r3x_bench_n8.json
python -c "
import json
d=json.load(open('gpurun_out/r3x_bench_n8.json'))
print('n_gpus',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],3))
t=d['train']; print('train',round(t['value'],1),'ms',round(t['ms_per_step'],2),t['allreduce']['ms_alone'], t['allreduce']['buckets_per_step'])
"
