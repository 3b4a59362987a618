#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3n_bench_n2.log 2> gpurun_out/r3n_bench_n2.err; echo "exit $?"
grep '^{"metric"' gpurun_out/r3n_bench_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],3))
t=d['train']; print('train',round(t['value'],1),'ms',round(t['ms_per_step'],2),t['allreduce'])
print('cpu_baseline' in d)
"
tail -3 gpurun_out/r3n_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
