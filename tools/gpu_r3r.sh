#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
DYK_BENCH_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-leg --sustain-s 0 2> gpurun_out/r3r_dbg_$i.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('run', $i, ' fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],3), 'graph', d['cuda_graph'])"
grep DEBUG gpurun_out/r3r_dbg_$i.err | head -1 | cut -c1-700
done
