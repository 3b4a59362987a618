#!/bin/bash
# Round-2 session V: parity suite after the batched native weight bind, smoke() under ncu (the driver's view of the first
# 1000 launches), racecheck / synccheck over the small kernel tests, per-layer times, launch list + DRAM traffic of the bench step.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run r2v_tests 1200 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider
run r2v_smoke_ncu 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r2v_smoke_launches.csv python __graft_entry__.py smoke
python tools/launch_summary.py gpurun_out/r2v_smoke_launches.csv > gpurun_out/r2v_smoke_launches_summary.txt 2>&1; head -30 gpurun_out/r2v_smoke_launches_summary.txt
run r2v_layers_v3 300 python tools/layer_times.py kaist_dyolov3_add_sl.cfg 16 gpurun_out/r2v_layers_v3.json
TAILN=40 run r2v_racecheck 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 900 -p no:cacheprovider -k "conv_bn_act or fusion or squeeze or batched"
TAILN=12 run r2v_synccheck 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 600 -p no:cacheprovider -k "conv_bn_act"
run r2v_ncu_dram 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2v_dram_v3.csv python tools/one_forward.py kaist_dyolov3_add_sl.cfg 16 2
python tools/dram_summary.py gpurun_out/r2v_dram_v3.csv > gpurun_out/r2v_dram_v3_summary.txt 2>&1; cat gpurun_out/r2v_dram_v3_summary.txt | head -40
