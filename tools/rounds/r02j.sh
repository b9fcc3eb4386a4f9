#!/bin/bash
TAG=r02j
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -x > $OUT/pytest_round2.log 2>&1; echo "pytest round2 rc=$?" >> $OUT/pytest_round2.log
tail -5 $OUT/pytest_round2.log
timeout 300 python tools/shape_bench.py 1 2 4 8 > $OUT/shape_bench_auto.jsonl 2> $OUT/shape_bench_auto.err; cat $OUT/shape_bench_auto.jsonl
SHAPE_CLASSES=2000000 timeout 300 python tools/shape_bench.py 8 > $OUT/shape_bench_c4shape.jsonl 2>&1; cat $OUT/shape_bench_c4shape.jsonl
