#!/bin/bash
# Round-2 GPU session 1: pipe microbench, parity subset, A/B of solo forward tiles (TROYB_SOLO), ncu of the mod-up kernel.
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 300 tests/microbench/pipes2 > $OUT/pipes2.jsonl 2>&1; echo "pipes2 rc=$?"; cat $OUT/pipes2.jsonl
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "ntt_sweep or evaluator_flow or batched" > $OUT/pytest_subset.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_subset.log
for solo in 0 1; do
  echo "== solo=$solo"
  TROYB_SOLO=$solo NTT_SIZES=4096,8192,16384 timeout 300 python tests/microbench/ntt_bench.py > $OUT/sweep_solo$solo.jsonl 2>&1; cat $OUT/sweep_solo$solo.jsonl
  TROYB_SOLO=$solo timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_solo$solo.json 2>$OUT/bench_solo$solo.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_solo$solo.json").read().strip().splitlines()[-1])
    print("bench solo=$solo", round(d["value"]), "ops/s  rotate", round(d["rotate"]["value"]), " e2e", round(d["e2e"]["value"]), "ntt", d["ntt"])
    for k in d["kernels"]: print("   ", k["kernel"], round(k["ms"],3), "ms", round(k["share"],3))
except Exception as e: print("bench parse failed", e)
PY
done
PROFILE_BATCH=256 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:NttFwdTile -s 1 -c 1 \
  -o $OUT/modup -f python tests/microbench/profile_target.py > $OUT/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la $OUT
