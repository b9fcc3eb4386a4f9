#!/bin/bash
# Round-2 validation visit (1 GPU), the driver's own sequence: full -m gpu suite, smoke(), bench c3 as the driver calls it,
# c4 / c5 lines, reference arm (short).
TAG=r02m
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$OUT/bench.json')); print('c3 n1', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity']['ok_all_ranks'], 'cpu', d['cpu_baseline'], d['extras'].get('reference_on_b200_ms'), d['roofline']['step_frac_of_burst_peak_per_gpu'], d['roofline']['step_hbm'])"
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 > $OUT/bench_c5.json 2> $OUT/bench_c5.err; echo "bench c5 rc=$?"
timeout 300 python bench.py --workload c4 --steps 10 --warmup 5 --no-cpu-baseline > $OUT/bench_c4_n1.json 2> $OUT/bench_c4_n1.err; echo "bench c4 rc=$?"
timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench c2 rc=$?"
python -c "
import json
for n in ['bench_c5','bench_c4_n1','bench_c2']:
    d=json.load(open('$OUT/'+n+'.json')); print(n, d['ms_per_step'], d['value'], d['unit'], d['parity']['ok_all_ranks'])"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"; cat $OUT/bench_reference.json | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > $OUT/launches_bench.log 2>&1; echo "ncu launches rc=$?"
