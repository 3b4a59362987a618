#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=12 run r3i_tests 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider
run r3i_smoke 300 python __graft_entry__.py smoke
run r3i_bench 900 python bench.py --steps 20 --warmup 5
python -c "
import json
d=json.loads(open('gpurun_out/r3i_bench.log').read().strip().splitlines()[-2])
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'roofline',d['roofline']['frac'],'train',d.get('train',{}).get('ms_per_step'),'sustained',d.get('sustained',{}).get('value'),d['clocks'])
"
