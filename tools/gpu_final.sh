#!/bin/bash
# Final check of a build on one B200: GPU test suite, smoke(), the default bench line and the reference arm.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400; }
run final_tests 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider
run final_smoke 300 python __graft_entry__.py smoke
TAILN=1 run final_bench 900 python bench.py
TAILN=1 run final_ref 600 python bench.py --impl reference --steps 3 --warmup 1
timeout 200 python tools/dw_bench.py --modes 1 --only 0,2,5 2>&1 | grep -v Summary
