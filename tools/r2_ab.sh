#!/bin/bash
# Round-2 development session: parity subset + A/B of env knobs (VARIANTS="name:ENV=val,ENV2=val ...") on the NTT sweep and the bench step.
TAG=${1:-r2ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_PYTEST" ]; then
  timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "${PYTEST_K:-ntt_sweep or evaluator_flow or batched}" > $OUT/pytest_subset.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_subset.log
fi
for v in ${VARIANTS:-base:}; do
  name=${v%%:*}; envs=${v#*:}
  echo "== $name env=$envs"
  ( for e in ${envs//,/ }; do export $e; done
    NTT_SIZES=${NTT_SIZES:-4096,8192,16384} timeout 300 python tests/microbench/ntt_bench.py > $OUT/sweep_$name.jsonl 2>&1; cat $OUT/sweep_$name.jsonl
    timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_$name.json 2>$OUT/bench_$name.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$name.json").read().strip().splitlines()[-1])
    print("bench $name", round(d["value"]), "ops/s  rotate", round(d["rotate"]["value"]), " rescale", round(d["rescale"]["value"]), " e2e", round(d["e2e"]["value"]), "ntt fwd/inv frac", round(d["ntt"]["forward_frac_of_hbm_roofline"],3), round(d["ntt"]["inverse_frac_of_hbm_roofline"],3))
    for k in d["kernels"]: print("   ", k["kernel"], round(k["ms"],3), "ms", round(k["share"],3))
except Exception as e: print("bench parse failed", e)
PY
  )
done
if [ -n "$NCU" ]; then
  PROFILE_BATCH=256 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:${NCU_K:-NttFwdTile} -s ${NCU_S:-1} -c 1 \
    -o $OUT/prof -f python tests/microbench/profile_target.py > $OUT/ncu_full.log 2>&1
  echo "full capture rc=$?"
fi
