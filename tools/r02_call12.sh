#!/bin/bash
# Last GPU minutes of the round: the push kernel's copy loop on one GPU, timed and under ncu.
mkdir -p gpurun_out
timeout 100 python tools/push_local_profile.py 2>&1 | tee gpurun_out/r02l_push_local.txt
timeout 150 ncu -k regex:gdv_sel_push --set full --clock-control none --import-source on -c 4 -o gpurun_out/r02_sel_push_local python tools/push_local_profile.py > gpurun_out/r02l_ncu.log 2>&1; tail -3 gpurun_out/r02l_ncu.log
