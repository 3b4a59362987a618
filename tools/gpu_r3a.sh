#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:halo2 -c 1 -s 2 -f -o gpurun_out/r3a_resident python tools/prof_resident.py > gpurun_out/r3a_ncu.log 2>&1; tail -3 gpurun_out/r3a_ncu.log
