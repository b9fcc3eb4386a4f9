#!/bin/bash
# Round-2 first GPU visit: whole -m gpu suite (incl. files never run on a B200), micro-benchmarks, dw epilogue breakdown,
# dx||dw pacing sweep (+ DRAM bytes of the backward graph), ROC bench + ncu, compute-sanitizer (time-boxed).
TAG=r02a
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_bw tools/micro/tmem_bw.cu && timeout 60 /tmp/tmem_bw > $OUT/tmem_bw.jsonl 2>&1
cat $OUT/tmem_bw.jsonl
for e in 0 1 2 4 8 6 7 15; do
  echo "FEDFR_DW_EXP=$e" >> $OUT/dw_probe.log
  FEDFR_DW_EXP=$e PROBE_CLUSTERS=52 timeout 120 python tools/dw_probe.py >> $OUT/dw_probe.log 2>&1
done
grep -E "FEDFR_DW_EXP|per item" $OUT/dw_probe.log
for lead in 0 2048 8192 32768; do
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prob-split 0,0,$lead > $OUT/bench_lead$lead.json 2> $OUT/bench_lead$lead.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_lead$lead.json"))
print("lead $lead ms/step", d["ms_per_step"], d["roofline"]["phase_ms_per_step"])
PY
done
for lead in 0 8192; do
  timeout 300 ncu --graph-profiling graph --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 30 -c 30 --csv \
    --log-file $OUT/graph_traffic_lead$lead.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --prob-split 0,0,$lead > $OUT/graph_traffic_lead$lead.log 2>&1
done
timeout 120 python tools/roc_bench.py > $OUT/roc_bench.jsonl 2> $OUT/roc_bench.err
timeout 120 python tools/roc_bench.py --mode 1 >> $OUT/roc_bench.jsonl 2>> $OUT/roc_bench.err
cat $OUT/roc_bench.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'roc_hist' -s 1 -c 1 -o $OUT/prof_roc python tools/roc_bench.py > $OUT/prof_roc_bench.log 2>&1
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden" > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 $OUT/sanitizer_memcheck.log
ls -la $OUT
