#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, per-layer times, ncu launch list + full capture.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
if [ -z "$SKIP_TESTS" ]; then
run t_gpu 1200 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x
run smoke 300 python __graft_entry__.py smoke
fi
run bench 600 python bench.py --steps 20 --warmup 3
run layers_v3 300 python tools/layer_times.py kaist_dyolov3_add_sl.cfg 16 gpurun_out/layers_v3.json
run layers_v4 300 python tools/layer_times.py kaist_dyolov4_fshare_global_concat_se3.cfg 16 gpurun_out/layers_v4.json
run convbench 300 python tools/conv_bench.py
DYK_B200_LIB=$PWD/double-yolo-kaist_b200/libdyk_b200_prof.so run convprof 300 python tools/conv_bench.py --iters 3
if [ -z "$SKIP_NCU" ]; then
run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python tools/one_forward.py kaist_dyolov3_add_sl.cfg 16 2
run ncu_full 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 9 -f -o gpurun_out/prof_conv python tools/conv_bench.py --only 1,2,3 --iters 1
fi
