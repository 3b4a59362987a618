#!/bin/bash
mkdir -p gpurun_out
python tools/chain_bench.py 2>&1 | grep -v Summary | tee gpurun_out/r2w_chain.txt
DYK_B200_LIB=$PWD/double-yolo-kaist_b200/libdyk_b200_prof.so python tools/chain_bench.py 2>&1 | grep -v Summary | tee gpurun_out/r2w_chain_prof.txt
DYK_PDL=0 python tools/chain_bench.py 2>&1 | grep -v Summary | tee gpurun_out/r2w_chain_nopdl.txt
