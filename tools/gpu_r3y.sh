#!/bin/bash
export DYK_CHAIN_ONLY=9,10,11
echo "--- default"; python tools/chain_bench.py 2>&1 | grep -v Summary
echo "--- DYK_HALO2=0 (1-CTA halo, auto BN)"; DYK_HALO2=0 python tools/chain_bench.py 2>&1 | grep -v Summary
echo "--- DYK_HALO2=0 DYK_HALO_BN=128"; DYK_HALO2=0 DYK_HALO_BN=128 python tools/chain_bench.py 2>&1 | grep -v Summary
echo "--- DYK_HALO2=0 DYK_HALO=all (first shape via halo)"; DYK_HALO2=0 DYK_HALO=all python tools/chain_bench.py 2>&1 | grep -v Summary
echo "--- DYK_NO_HALO=1 (generic)"; DYK_NO_HALO=1 python tools/chain_bench.py 2>&1 | grep -v Summary
