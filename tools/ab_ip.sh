#!/bin/bash
# A/B of the key inner product's 40-bit form (SmallAcc): general form only, 40-bit form at 3 CTAs per SM, at 2 CTAs per SM.
OUT=gpurun_out/ab_ip; mkdir -p $OUT
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k=[(x["kernel"],x["launches"],round(x["ms"],4)) for x in d.get("kernels",[])]
print(sys.argv[1], round(d["value"],1), d["ms_per_step"], k)
PY
}
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "extreme or evaluator_flow or many_digits or batched" > $OUT/pytest.log 2>&1; echo "parity rc=$?"; tail -2 $OUT/pytest.log
TROYB_IP_SMALL=0 timeout 300 python bench.py --no-cpu-baseline --no-sweep > $OUT/general.json 2> $OUT/general.err; pick $OUT/general.json
timeout 300 python bench.py --no-cpu-baseline --no-sweep > $OUT/small_c3.json 2> $OUT/small_c3.err; pick $OUT/small_c3.json
cp build/ab/libtroy_b200_c2.so troy_b200/libtroy_b200.so
timeout 300 python bench.py --no-cpu-baseline --no-sweep > $OUT/small_c2.json 2> $OUT/small_c2.err; pick $OUT/small_c2.json
