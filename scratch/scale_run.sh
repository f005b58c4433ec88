#!/bin/bash
# 1/2/4/8-GPU bench lines of c3 (weak scaling, DP) and c5 (strong scaling, spatial slabs); run on an 8-GPU box
mkdir -p gpurun_out
for w in c3 c5; do
  for n in 1 2 4 8; do
    out=gpurun_out/r02_scale_${w}_${n}gpu.json
    if [ $n = 1 ]; then timeout 300 python bench.py --workload $w --no-cpu-baseline > $out 2> ${out%.json}.err
    else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --workload $w --no-cpu-baseline > $out 2> ${out%.json}.err; fi
    python scratch/show_bench.py $out
  done
done
# strong scaling of c3: fixed global batch 256
for n in 2 4 8; do
  out=gpurun_out/r02_strong_c3_${n}gpu.json
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --workload c3 --molecules $((256 / n)) --no-cpu-baseline > $out 2> ${out%.json}.err
  python scratch/show_bench.py $out
done
