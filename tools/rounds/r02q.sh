#!/bin/bash
# Final round-2 multi-GPU visit (8 GPUs, final code: dw4 LITE auto, sampling fixes): the 2-rank NCCL parity tests, then c3 / c4 / c5 at N = 8 (and c3 at N = 2) with the
# post-timing parity self-check on every rank; logs are copied to profiles/ by hand afterwards.
TAG=r02q
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q > $OUT/pytest_multirank.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multirank.log
tail -8 $OUT/pytest_multirank.log
run() {   # name, nproc, args...
  name=$1; n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n "$@" \
    > $OUT/$name.json 2> $OUT/$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$name.json"))
    print("$name", "ms/step", round(d["ms_per_step"],4), "value", round(d["value"],1), d["unit"], "e2e", d["e2e"] and round(d["e2e"]["value"],1), "parity", d.get("parity",{}).get("ok_all_ranks"))
    r=d.get("roofline",{})
    print("   phases", {k: round(v,4) for k,v in r.get("phase_ms_per_step",{}).items()})
except Exception as e:
    print("$name failed", e)
PY
  tail -2 $OUT/$name.err
}
run bench_c3_n8 8 --steps 20 --warmup 5
run bench_c4_n8 8 --workload c4 --steps 20 --warmup 5
run bench_c5_n8 8 --workload c5 --steps 10 --warmup 3
run bench_c3_n2 2 --steps 20 --warmup 5
run bench_c3_n4 4 --steps 20 --warmup 5
run bench_c3_n8_dw4off 8 --steps 20 --warmup 5 --dw4 0
run bench_ref_c3_n8 8 --impl reference --steps 2 --warmup 1
