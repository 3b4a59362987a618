#!/bin/bash
export DYK_CHAIN_ONLY=15,16,17
echo "--- plain"; python tools/chain_bench.py 2>&1 | grep -v Summary
echo "--- dual, 8 combiner warps"; DYK_CHAIN_DUAL=1 python tools/chain_bench.py 2>&1 | grep -v Summary
echo "--- dual, 4 combiner warps"; DYK_B200_LIB=$PWD/double-yolo-kaist_b200/build/alt/libdyk_comb4.so DYK_CHAIN_DUAL=1 python tools/chain_bench.py 2>&1 | grep -v Summary
