#!/bin/bash
# Development aid: one GPU-box session = parity tests + A/B timings of library variants (TROYB_LIB) + bench line.
# Usage (from the repo root on the GPU box):  bash tools/gpu_ab.sh [tag]
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_PYTEST" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
fi
SIZES=${NTT_SIZES:-4096,16384,65536}
for v in $VARIANTS; do
  name=${v%%:*}; rest=${v#*:}; lib=${rest%%:*}; envs=${rest#*:}
  [ "$envs" = "$rest" ] && envs=""
  echo "== $name lib=$lib env=$envs"
  ( [ "$lib" != "-" ] && export TROYB_LIB=$PWD/$lib; for e in ${envs//,/ }; do export $e; done
    NTT_SIZES=$SIZES timeout 300 python tests/microbench/ntt_bench.py > $OUT/sweep_$name.jsonl 2>&1
    cat $OUT/sweep_$name.jsonl
    if [ -n "$WITH_BENCH" ]; then timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_$name.json 2>$OUT/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$name.json").read().strip().splitlines()[-1])
    print("bench", "$name", round(d["value"]), "ops/s  rotate", round(d["rotate"]["value"]), " e2e", round(d["e2e"]["value"]))
    for k in d["kernels"]: print("   ", k["kernel"], round(k["ms"],3), "ms", round(k["share"],3))
except Exception as e: print("bench parse failed", e)
PY
    fi )
done
