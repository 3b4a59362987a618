#!/bin/bash
mkdir -p gpurun_out
export DYK_B200_LIB=$PWD/double-yolo-kaist_b200/libdyk_b200_prof.so
for pf in 0 1 2; do echo "--- DYK_RES_PF=$pf"; DYK_RES_PF=$pf python tools/halo2_prof.py 2>&1 | grep -v Summary | head -8; done | tee gpurun_out/r2z_halo2_prof.txt
unset DYK_B200_LIB
for pf in 0 1; do echo "--- chain DYK_RES_PF=$pf"; DYK_RES_PF=$pf python tools/chain_bench.py 2>&1 | grep -v Summary | tail -3; done | tee gpurun_out/r2z_chain.txt
echo "--- chain DYK_RES_PF=1 no resident"; DYK_HALO2_RES=0 DYK_RES_PF=1 python tools/chain_bench.py 2>&1 | grep -v Summary | tail -3 | tee -a gpurun_out/r2z_chain.txt
