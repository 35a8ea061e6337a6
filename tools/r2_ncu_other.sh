#!/bin/bash
# ncu --set full of one launch each of the key inner product, the tensor product and the first inverse transform at the bench batch
OUT=${1:-gpurun_out/r2other}
mkdir -p $OUT
for spec in "KsInnerProductStaged 1 inner_product" "EltTensor 1 tensor_product" "NttInvTile2 1 ntt_inv_first"; do
  set -- $spec
  PROFILE_BATCH=256 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$1 -s $2 -c 1 \
    -o $OUT/$3 -f python tests/microbench/profile_target.py > $OUT/ncu_$3.log 2>&1
  echo "$3 rc=$?"
  ncu -i $OUT/$3.ncu-rep --page raw --csv > $OUT/$3_raw.csv 2>/dev/null
  ncu -i $OUT/$3.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_trim.py > $OUT/$3_sass.csv
  rm -f $OUT/$3.ncu-rep
done
