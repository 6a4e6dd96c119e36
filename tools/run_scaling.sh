#!/bin/bash
# Multi-GPU runs of BASELINE configs 2, 3 and 4 (one rank per GPU, torchrun on one node); JSON lines land in gpurun_out/scale_r2/.
# usage: gpurun --gpus 8 -- bash tools/run_scaling.sh "2 4 8" "3 4" ; third argument: configs to run only at the largest N
out=gpurun_out/scale_r2
mkdir -p $out
port=29500
ns=${1:-2 4 8}
cfgs=${2:-3 4}
last=$(echo $ns | awk '{print $NF}')
for n in $ns; do
  list="$cfgs"
  [ "$n" = "$last" ] && list="$cfgs ${3:-}"
  for cfg in $list; do
    steps=2; [ $cfg = 2 ] && steps=4
    port=$((port+1))
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --config $cfg --steps $steps --warmup 3 --no-cpu-baseline > $out/c${cfg}_n${n}.json 2> $out/c${cfg}_n${n}.err
    echo "config $cfg N=$n rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('$out/c${cfg}_n${n}.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'])
except Exception as e: print('ERR', e)")"
  done
done
nvidia-smi topo -m > $out/topo.txt 2>&1
