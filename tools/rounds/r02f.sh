#!/bin/bash
TAG=r02f
OUT=gpurun_out/$TAG
mkdir -p $OUT
export FEDFR_DW4=1
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "rows_vs_bf16" > $OUT/pytest_rows.log 2>&1; rc=$?; echo "pytest rows rc=$rc" >> $OUT/pytest_rows.log
tail -5 $OUT/pytest_rows.log
for e in 0 1 2 4 3 7; do
  echo "FEDFR_DW_EXP=$e" >> $OUT/dw_probe.log
  FEDFR_DW_EXP=$e PROBE_DX_SMS=44 PROBE_CLUSTERS=26 timeout 120 python tools/dw_probe.py >> $OUT/dw_probe.log 2>&1
  echo "FEDFR_DW_EXP=$e dw alone" >> $OUT/dw_probe.log
  FEDFR_EXP=3 FEDFR_DW_EXP=$e PROBE_DX_SMS=44 PROBE_CLUSTERS=26 timeout 120 python tools/dw_probe.py >> $OUT/dw_probe.log 2>&1
done
grep -E "FEDFR_DW_EXP|per item" $OUT/dw_probe.log
timeout 200 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-parity > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("ms/step", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["phase_ms_per_step"].items()})
PY
timeout 120 python tools/host_profile.py 0.1 4096 > $OUT/host_profile_c4.log 2>&1; head -60 $OUT/host_profile_c4.log
