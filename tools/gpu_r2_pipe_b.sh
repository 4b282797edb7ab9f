#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"mind_pipe_kernel" -s 2 -c 1 -f -o gpurun_out/pipe_b_noise \
    python tools/prof_mind.py mind_noise > gpurun_out/pipe_b_ncu.log 2>&1
tail -3 gpurun_out/pipe_b_ncu.log
ls -la gpurun_out/*.ncu-rep
