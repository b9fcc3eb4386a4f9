#!/bin/bash
# Sweep the backward G-chunk size (L2 residency of G / w_hat between the G, dx and dw kernels), graph replay on/off.
OUT=gpurun_out/${1:-sweep}
mkdir -p $OUT
for cfg in "0 256" "1 256" "1 128" "1 64" "1 48" "1 32" "1 24" "1 16"; do
  set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --graph $1 --chunk-mb $2 > $OUT/g$1_chunk_$2.json 2> $OUT/g$1_chunk_$2.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/g$1_chunk_$2.json").read().strip().splitlines()[-1])
    print("graph=$1 chunk_mb=$2 ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {k: round(v, 3) for k, v in d["roofline"]["phase_ms_per_step"].items()})
except Exception as e:
    print("graph=$1 chunk_mb=$2 failed", e)
PY
done
