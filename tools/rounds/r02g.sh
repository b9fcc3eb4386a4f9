#!/bin/bash
TAG=r02g
OUT=gpurun_out/$TAG
mkdir -p $OUT
export FEDFR_DW4=1
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "rows_vs_bf16 or job_shapes" > $OUT/pytest_rows.log 2>&1; rc=$?; echo "pytest rows rc=$rc" >> $OUT/pytest_rows.log
tail -5 $OUT/pytest_rows.log
for e in 0 2; do
  echo "FEDFR_DW_EXP=$e" >> $OUT/dw_probe.log
  FEDFR_DW_EXP=$e PROBE_DX_SMS=44 PROBE_CLUSTERS=26 timeout 120 python tools/dw_probe.py >> $OUT/dw_probe.log 2>&1
done
grep -E "FEDFR_DW_EXP|per item" $OUT/dw_probe.log
for dx in 0 52 60; do
timeout 200 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-parity --prob-split $dx,0,0 > $OUT/bench$dx.json 2> $OUT/bench$dx.err
python - <<PY
import json
d=json.load(open("$OUT/bench$dx.json"))
print("dx $dx ms/step", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["phase_ms_per_step"].items()})
PY
done
unset FEDFR_DW4
# c5 after the staging ring; forward graph DRAM traffic reference
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_c5.json 2> $OUT/bench_c5.err; python -c "
import json; d=json.load(open('$OUT/bench_c5.json')); print('c5', d['ms_per_step'], d['value'], d['extras'], d['parity'])"
timeout 120 python tools/roc_bench.py --mode 1 > $OUT/roc.jsonl 2>&1; cat $OUT/roc.jsonl
