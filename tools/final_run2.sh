#!/bin/bash
# Second one-GPU measurement pass of round 2 (after the CABAC / SAO work): tests, smoke, every bench configuration, micro-benchmarks,
# launch list, isolated k_entropy timings.  Outputs: gpurun_out/final_r2b/
o=gpurun_out/final_r2b
mkdir -p $o
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $o/pytest_gpu.txt; cat $o/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $o/smoke.txt 2>&1; tail -1 $o/smoke.txt
python bench.py --gpus 1 --steps 10 --warmup 3 > $o/bench_c2_n1.json 2> $o/bench_c2_n1.err; tail -c 300 $o/bench_c2_n1.err
python bench.py --steps 6 --no-cpu-baseline > $o/bench_c2_n1_repeat1.json 2>/dev/null
python bench.py --steps 6 --no-cpu-baseline > $o/bench_c2_n1_repeat2.json 2>/dev/null
python bench.py --config 1 --steps 10 > $o/bench_c1_n1.json 2> $o/bench_c1_n1.err
python bench.py --config 3 --steps 2 > $o/bench_c3_n1.json 2> $o/bench_c3_n1.err
python bench.py --config 4 --steps 2 > $o/bench_c4_n1.json 2> $o/bench_c4_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/bench_reference_arm.json 2> $o/bench_reference_arm.err
python tools/pixel_bench.py > $o/microbench.jsonl 2> $o/microbench.err
python tools/dropin_bench.py 4k60_hdr 1080p_sdr > $o/dropin.jsonl 2> $o/dropin.err; cat $o/dropin.jsonl
ncu -k regex:^k_ --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $o/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/ncu_launches.log 2>&1
python tools/launch_summary.py $o/launches.csv > $o/launch_summary.md; cat $o/launch_summary.md
for run in 1 2; do
ncu -k regex:^k_entropy --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file $o/entropy_isolated_run$run.csv python bench.py --steps 1 --warmup 3 --frames 24 --batch 24 --streams 1 --no-cpu-baseline > /dev/null 2>&1
echo "k_entropy isolated, run $run"; python tools/launch_summary.py $o/entropy_isolated_run$run.csv
done
ncu --set full --clock-control none --import-source on -k regex:^k_entropy --launch-skip 9 -c 1 -f -o $o/ncu_entropy python bench.py --steps 1 --warmup 3 --frames 24 --batch 24 --streams 1 --no-cpu-baseline > $o/ncu_entropy.log 2>&1
ncu -i $o/ncu_entropy.ncu-rep --page raw --csv > $o/ncu_r2b_entropy_raw.csv 2>/dev/null; rm -f $o/ncu_entropy.ncu-rep
for f in c2_n1 c2_n1_repeat1 c2_n1_repeat2 c1_n1 c3_n1 c4_n1 reference_arm; do python - <<PY
import json
try:
    d=json.loads(open('$o/bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['e2e']['value'], d.get('ms_per_step_median'), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print('$f ERR', e)
PY
done
