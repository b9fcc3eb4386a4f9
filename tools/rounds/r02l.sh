#!/bin/bash
TAG=r02l
OUT=gpurun_out/$TAG
mkdir -p $OUT
for e in 0 1 2 3; do
  echo "FEDFR_FWD_EXP=$e"
  FEDFR_FWD_EXP=$e timeout 300 python tools/shape_bench.py 1 4 8 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d=json.loads(l); print(' W', d['W'], 'step', d['ms_per_step'], 'fwd', d['phase_ms']['fwd'], 'norm', d['phase_ms']['normalize'])
    except Exception: pass"
done
