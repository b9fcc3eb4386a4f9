#!/bin/bash
# usage: SWEEP_CFG="<bench args line>\n..." tools/sweep_args.sh <tag>   -- one bench.py run per line, prints ms/step + phases
OUT=gpurun_out/${1:-sweep_args}
mkdir -p $OUT
i=0
while IFS= read -r line; do
  [ -z "$line" ] && continue
  i=$((i+1))
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $line > $OUT/run_$i.json 2> $OUT/run_$i.err
  python - "$line" <<PY
import json, sys
try:
    d = json.loads(open("$OUT/run_$i.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "| ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), {k[:8]: round(v, 3) for k, v in d["roofline"]["phase_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "| failed", e, open("$OUT/run_$i.err").read()[-400:])
PY
done <<CFG
${SWEEP_CFG}
CFG
