#!/bin/bash
# knob sweeps at N=1 (forward chunks, dx/dw SM split) with the final kernels
TAG=r02k
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {
  name=$1; shift
  timeout 200 python bench.py --steps 12 --warmup 5 --no-cpu-baseline --no-parity "$@" > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$name.json"))
    print("$name", "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k[:12]: round(v,4) for k,v in d["roofline"]["phase_ms_per_step"].items()})
except Exception as e:
    print("$name failed", e)
PY
}
run base
for c in 2 3 6 8; do run chunks$c --fwd-overlap $c,2; done
for b in 1 3 4; do run normblk$b --fwd-overlap 4,$b; done
for dx in 36 40 48 52; do run dx$dx --prob-split $dx,0,0; done
