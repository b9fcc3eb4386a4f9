#!/bin/bash
TAG=r02i
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sampling or golden_single_rank" > $OUT/pytest_sampling.log 2>&1; echo "pytest sampling rc=$?" >> $OUT/pytest_sampling.log
tail -4 $OUT/pytest_sampling.log
timeout 300 python bench.py --workload c4 --steps 10 --warmup 5 --no-cpu-baseline > $OUT/bench_c4_n1.json 2> $OUT/bench_c4_n1.err; echo "bench c4 rc=$?"
python -c "
import json; d=json.load(open('$OUT/bench_c4_n1.json')); print('c4 n1', d['ms_per_step'], d['parity']['ok_all_ranks'])"
export FEDFR_DW4=2
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "rows_vs_bf16 or job_shapes" > $OUT/pytest_rows_lite.log 2>&1; rc=$?; echo "pytest rows lite rc=$rc" >> $OUT/pytest_rows_lite.log
tail -5 $OUT/pytest_rows_lite.log
if [ $rc -eq 0 ]; then
  timeout 300 python tools/shape_bench.py 1 2 4 8 > $OUT/shape_bench_lite.jsonl 2> $OUT/shape_bench_lite.err; cat $OUT/shape_bench_lite.jsonl
  PROBE_DX_SMS=44 PROBE_CLUSTERS=26 timeout 120 python tools/dw_probe.py > $OUT/dw_probe_lite.log 2>&1; grep "per item" $OUT/dw_probe_lite.log
fi
unset FEDFR_DW4
timeout 300 python tools/shape_bench.py 1 2 4 8 > $OUT/shape_bench_base.jsonl 2> $OUT/shape_bench_base.err; cat $OUT/shape_bench_base.jsonl
