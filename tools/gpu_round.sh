#!/bin/bash
# One GPU-box visit: parity tests, bench lines (c3 + c5 + reference arm), per-rank shape bench, ncu launch list,
# ncu full captures of the dominant kernels.      usage: tools/gpu_round.sh <tag> [skip-tests|-] [skip-ncu|-] [sanitize]
TAG=${1:-rXX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
FEDFR_BWD_MODE=recompute timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_recompute.json 2> $OUT/bench_recompute.err; echo "bench recompute rc=$?"
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 > $OUT/bench_c5.json 2> $OUT/bench_c5.err; echo "bench c5 rc=$?"
cat $OUT/bench_c5.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench ref rc=$?"
timeout 300 python tools/shape_bench.py 1 2 4 8 > $OUT/shape_bench.jsonl 2> $OUT/shape_bench.err
cat $OUT/shape_bench.jsonl
timeout 120 python tools/roc_bench.py > $OUT/roc_bench.jsonl 2> $OUT/roc_bench.err
timeout 120 python tools/roc_bench.py --mode 1 >> $OUT/roc_bench.jsonl 2>> $OUT/roc_bench.err
cat $OUT/roc_bench.jsonl
if [ "$3" != "skip-ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'logits2?_kernel|dx2?_kernel|dw_kernel|normalize_rows' -s 8 -c 7 \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/prof_bench.log 2>&1
  timeout 600 ncu --set full --clock-control none -k regex:'fedavg' -s 3 -c 1 \
    -o $OUT/prof_c5 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/prof_c5_bench.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:'roc_hist' -s 1 -c 1 \
    -o $OUT/prof_roc python tools/roc_bench.py > $OUT/prof_roc_bench.log 2>&1
fi
if [ "$4" == "sanitize" ]; then   # memcheck + racecheck over the small-shape parity tests (slow: minutes)
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_roc.py tests/test_gpu_parity.py -m gpu -x -q \
    -k "golden or roc or oracle" > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
  timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_roc.py -m gpu -x -q \
    > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
  tail -3 $OUT/sanitizer_memcheck.log $OUT/sanitizer_racecheck.log
fi
ls -la $OUT
