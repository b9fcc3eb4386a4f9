#!/bin/bash
# Sweep the pipelined backward: SM split between the G / dx / dw chains, ring depth, chunk size.
OUT=gpurun_out/${1:-sweep_pipe}
mkdir -p $OUT
i=0
while read -r pipe chunk; do
  [ -z "$pipe" ] && continue
  i=$((i+1))
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --pipe $pipe --chunk-mb $chunk > $OUT/run_$i.json 2> $OUT/run_$i.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/run_$i.json").read().strip().splitlines()[-1])
    print("pipe=$pipe chunk_mb=$chunk ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {k: round(v, 3) for k, v in d["roofline"]["phase_ms_per_step"].items()})
except Exception as e:
    print("pipe=$pipe chunk_mb=$chunk failed", e, open("$OUT/run_$i.err").read()[-300:])
PY
done <<CFG
${SWEEP_CFG}
CFG
