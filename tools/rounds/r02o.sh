#!/bin/bash
TAG=r02o
OUT=gpurun_out/$TAG
mkdir -p $OUT
export FEDFR_DW_P2=1
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "rows_vs_bf16 or job_shapes" > $OUT/pytest_rows.log 2>&1; rc=$?; echo "pytest rows rc=$rc" >> $OUT/pytest_rows.log
tail -5 $OUT/pytest_rows.log
if [ $rc -ne 0 ]; then exit 0; fi
PROBE_CLUSTERS=23 timeout 120 python tools/dw_probe.py > $OUT/dw_probe.log 2>&1; grep "per item" $OUT/dw_probe.log
timeout 300 python tools/shape_bench.py 1 2 > $OUT/shape_bench_p2.jsonl 2> $OUT/shape_bench_p2.err; cat $OUT/shape_bench_p2.jsonl
for dx in 0 44 52 60; do
  timeout 200 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-parity --prob-split $dx,0,0 > $OUT/bench_dx$dx.json 2> $OUT/bench_dx$dx.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_dx$dx.json"))
    print("dx_sms $dx ms/step", round(d["ms_per_step"],4), {k[:10]: round(v,4) for k,v in d["roofline"]["phase_ms_per_step"].items()})
except Exception as e:
    print("dx_sms $dx failed", e)
PY
done
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
