#!/bin/bash
# ncu --set full capture of the secondary kernels of the ComputeQ chain (one launch each, warm) -> gpurun_out/$1.ncu-rep
# usage (on the GPU box): bash scripts/ncu_capture.sh <name> <kernel-regex> [script args]
name=$1; regex=$2; shift 2
ncu --set full --import-source on --clock-control none -k "regex:$regex" -s 4 -c 6 -o gpurun_out/$name python scripts/dev_f2_one.py "$@" > gpurun_out/$name.log 2>&1
tail -2 gpurun_out/$name.log
