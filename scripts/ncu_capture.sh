#!/bin/bash
# ncu --set full capture of kernels of one timestep (32 cells, N = Nv = 32, eager launches) -> gpurun_out/$1.ncu-rep
# usage (on the GPU box): bash scripts/ncu_capture.sh <name> <kernel-regex> [launches to skip] [launches to capture]
name=$1; regex=$2; skip=${3:-4}; count=${4:-6}
ncu --set full --import-source on --clock-control none -k "regex:$regex" -s $skip -c $count -o gpurun_out/$name python scripts/dev_one_collide.py > gpurun_out/$name.log 2>&1
tail -2 gpurun_out/$name.log
