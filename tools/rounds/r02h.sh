#!/bin/bash
# ncu full captures of the HBM-bound SIMT kernels (gather / scatter / sampling / fused SGD) at c4's shard shape, the launch
# list of one c4 step, compute-sanitizer racecheck on the small parity tests.
TAG=r02h
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'move_rows2|sample_|sgd_rows|remap_labels|normalize_rows|cast_rows' -s 40 -c 24 \
  -o $OUT/prof_rows python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $OUT/prof_rows.log 2>&1; echo "ncu rows rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 160 --csv --log-file $OUT/launches_c4.csv \
  python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $OUT/launches_c4.log 2>&1; echo "ncu launches rc=$?"
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_single_rank and w1_sr01" > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 $OUT/sanitizer_racecheck.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "rows_vs_bf16 or fedavg or margin_callables" > $OUT/sanitizer_memcheck2.log 2>&1; echo "memcheck2 rc=$?"
tail -4 $OUT/sanitizer_memcheck2.log
ls -la $OUT
