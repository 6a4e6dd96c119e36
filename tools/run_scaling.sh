#!/bin/bash
# Multi-GPU runs of BASELINE configs 2, 3 and 4 (one rank per GPU, torchrun on one node); JSON lines land in gpurun_out/scale_r2/.
# usage: gpurun --gpus 8 -- bash tools/run_scaling.sh "2 4 8"
out=gpurun_out/scale_r2
mkdir -p $out
port=29500
for n in ${1:-2 4 8}; do
  for cfg in 3 4 2; do
    steps=2; [ $cfg = 2 ] && steps=4
    port=$((port+1))
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --config $cfg --steps $steps --warmup 3 --no-cpu-baseline > $out/c${cfg}_n${n}.json 2> $out/c${cfg}_n${n}.err
    echo "config $cfg N=$n rc=$? $(tail -c 300 $out/c${cfg}_n${n}.json | head -c 300)"
  done
done
nvidia-smi topo -m > $out/topo.txt 2>&1
