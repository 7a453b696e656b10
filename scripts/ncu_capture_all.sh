#!/bin/bash
# Profile set of a build (run on the GPU box): ncu --set full of the kernels of one 32-cell timestep, and the launch lists
# of bench.py itself at the headline size (Nx = 512 on one GPU) and at the weak-scaling shard (32 cells).
# Outputs under gpurun_out/ (summaries are made from the .ncu-rep files with scripts/ncu_summary.py).
tag=$1
ncu --set full --import-source on --clock-control none -k "regex:k_fc3_f2|k_fc3_f1|k_fc3_f3" -s 6 -c 3 -o gpurun_out/${tag}_fc3 python scripts/dev_one_collide.py > gpurun_out/${tag}_fc3.log 2>&1
ncu --set full --clock-control none -k "regex:k_dg_stage|k_project|k_tf_|k_field|k_sample|k_conserve" -s 30 -c 16 -o gpurun_out/${tag}_secondary python scripts/dev_one_collide.py > gpurun_out/${tag}_secondary.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 260 --csv --log-file gpurun_out/${tag}_bench512_launch_list.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${tag}_bench512_under_ncu.json 2> gpurun_out/${tag}_bench512_under_ncu.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/${tag}_bench32_launch_list.csv python bench.py --steps 2 --warmup 3 --scaling weak --no-extras > gpurun_out/${tag}_bench32_under_ncu.json 2> gpurun_out/${tag}_bench32_under_ncu.err
for f in fc3 secondary; do python scripts/ncu_summary.py gpurun_out/${tag}_$f.ncu-rep > gpurun_out/${tag}_${f}_ncu_summary.txt; done
python scripts/launch_summary.py gpurun_out/${tag}_bench512_launch_list.csv 1 > gpurun_out/${tag}_bench512_launch_summary.txt
ls -la gpurun_out/${tag}_*
