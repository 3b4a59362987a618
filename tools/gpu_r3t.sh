#!/bin/bash
# final r02 captures of the bench step: launch list (durations) + DRAM bytes per kernel
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r3t_dram_v3.csv python tools/one_forward.py kaist_dyolov3_add_sl.cfg 16 2 > gpurun_out/r3t_ncu.log 2>&1; tail -1 gpurun_out/r3t_ncu.log
python tools/dram_summary.py gpurun_out/r3t_dram_v3.csv 11.04 > gpurun_out/r3t_dram_v3_summary.txt 2>&1; cat gpurun_out/r3t_dram_v3_summary.txt
python tools/launch_summary.py gpurun_out/r3t_dram_v3.csv > gpurun_out/r3t_launch_summary.txt 2>&1; head -20 gpurun_out/r3t_launch_summary.txt
python tools/layer_times.py kaist_dyolov3_add_sl.cfg 16 gpurun_out/r3t_layers_v3.json 2>&1 | grep -v Summary > gpurun_out/r3t_layers_v3.txt; head -8 gpurun_out/r3t_layers_v3.txt
python tools/layer_times.py kaist_dyolov4_fshare_global_concat_se3.cfg 16 gpurun_out/r3t_layers_v4.json 2>&1 | grep -v Summary > gpurun_out/r3t_layers_v4.txt; head -8 gpurun_out/r3t_layers_v4.txt
python tools/chain_bench.py 2>&1 | grep -v Summary > gpurun_out/r3t_chain.txt; cat gpurun_out/r3t_chain.txt
