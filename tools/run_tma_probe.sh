#!/bin/bash
# Runs the tiled-TMA probes (tools/tma_probe*.cu, prebuilt here with nvcc for sm_100a) and records their stdout.
out=gpurun_out/tma_probe_r2.log
: > $out
nvidia-smi --query-gpu=name,driver_version,compute_cap,mig.mode.current --format=csv >> $out 2>&1
for mode in 0 4 1 2 3; do
  for var in 0 1 2 4; do
    echo "== tma_probe mode=$mode var=$var" >> $out
    timeout 60 tools/tma_probe $mode $var >> $out 2>&1
    echo "exit=$?" >> $out
  done
done
for x in 40 32; do echo "== tma_probe mode=1 var=0 x=$x" >> $out; timeout 60 tools/tma_probe 1 0 $x >> $out 2>&1; echo "exit=$?" >> $out; done
echo "== tma_probe2 (CUTLASS SM90_TMA_LOAD_2D)" >> $out
timeout 60 tools/tma_probe2 >> $out 2>&1
echo "exit=$?" >> $out
if [ -x tools/tma_probe3 ]; then
  echo "== tma_probe3 (cuda::ptx / driver-API launch)" >> $out
  for x in 40 37 32 8 1; do for t in 128 32; do
    timeout 60 tools/tma_probe3 $x $t >> $out 2>&1
    echo "exit=$?" >> $out
  done; done
fi
tail -100 $out
