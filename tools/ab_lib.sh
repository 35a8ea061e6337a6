#!/bin/bash
# A/B of two builds of the library on one box: the in-tree one, then build/ab/<name>.so copied over it (the box copy is scratch).
# Usage: bash tools/ab_lib.sh <name> [pytest -k expression]
NAME=$1; KEXPR=${2:-"extreme or evaluator_flow or many_digits or batched"}
OUT=gpurun_out/ab_$NAME; mkdir -p $OUT
pick() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k=[(x["kernel"],x["launches"],round(x["ms"],4)) for x in d.get("kernels",[])]
print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],4), k, d.get("matmul",{}).get("ms"), d.get("rescale",{}).get("value"))
PY
}
timeout 300 python bench.py --no-cpu-baseline --no-sweep > $OUT/base.json 2> $OUT/base.err; pick $OUT/base.json
cp build/ab/libtroy_b200_$NAME.so troy_b200/libtroy_b200.so
timeout 400 python -m pytest tests/test_parity_gpu.py -x -q -k "$KEXPR" > $OUT/pytest.log 2>&1; echo "parity rc=$?"; tail -2 $OUT/pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-sweep > $OUT/variant.json 2> $OUT/variant.err; pick $OUT/variant.json
