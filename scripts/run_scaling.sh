#!/bin/bash
# bench.py on N GPUs of this box (strong = BASELINE config 5, weak = config 4 shard); usage: bash scripts/run_scaling.sh N tag
n=$1; tag=$2
for sc in strong weak; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 40 --scaling $sc --no-extras > gpurun_out/${tag}_bench_${n}gpu_$sc.json 2> gpurun_out/${tag}_bench_${n}gpu_$sc.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $n --steps 40 --scaling $sc --no-extras > gpurun_out/${tag}_bench_${n}gpu_$sc.json 2> gpurun_out/${tag}_bench_${n}gpu_$sc.err
  fi
  tail -c 300 gpurun_out/${tag}_bench_${n}gpu_$sc.json | head -c 300; echo
done
