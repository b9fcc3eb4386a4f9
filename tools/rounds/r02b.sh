#!/bin/bash
# Round-2 second GPU visit (1 GPU): new parity tests, bench lines with the post-timing self-check and the unmodified
# reference arm, c4 / c5 workloads, dw-kernel SM-count experiment (per-SM vs chip-wide limit).
TAG=r02b
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x > $OUT/pytest_round2.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_round2.log
tail -30 $OUT/pytest_round2.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json; tail -3 $OUT/bench.err
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 > $OUT/bench_c5.json 2> $OUT/bench_c5.err; echo "bench c5 rc=$?"
cat $OUT/bench_c5.json; tail -3 $OUT/bench_c5.err
timeout 300 python bench.py --workload c4 --steps 10 --warmup 5 --no-cpu-baseline > $OUT/bench_c4_n1.json 2> $OUT/bench_c4_n1.err; echo "bench c4 rc=$?"
cat $OUT/bench_c4_n1.json; tail -3 $OUT/bench_c4_n1.err
for cfg in "44 52" "88 30" "118 15"; do
  set -- $cfg
  echo "dw alone, dx_sms=$1 clusters=$2" >> $OUT/dw_probe.log
  FEDFR_EXP=3 PROBE_DX_SMS=$1 PROBE_CLUSTERS=$2 timeout 120 python tools/dw_probe.py >> $OUT/dw_probe.log 2>&1
  FEDFR_EXP=3 FEDFR_DW_EXP=15 PROBE_DX_SMS=$1 PROBE_CLUSTERS=$2 timeout 120 python tools/dw_probe.py >> $OUT/dw_probe.log 2>&1
done
grep -E "dw alone|per item" $OUT/dw_probe.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench ref rc=$?"
cat $OUT/bench_reference.json; tail -3 $OUT/bench_reference.err
