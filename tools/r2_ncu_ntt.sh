#!/bin/bash
# ncu --set full capture of one standalone transform launch per kernel (N, BITS from the environment); the report is reduced to
# its raw-metrics and SASS-source CSV pages on the box (reports are too large to bring back several at a time).
TAG=${1:-r2ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for k in ${KERNELS:-NttFwdTile2 NttInvTile2}; do
  name=${k}_n${N:-16384}_b${BITS:-40}
  PROFILE_N=${N:-16384} PROFILE_BITS=${BITS:-40} timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:$k -s 2 -c 1 -o $OUT/$name -f python tests/microbench/profile_ntt.py > $OUT/ncu_$name.log 2>&1
  echo "$k rc=$?"
  ncu -i $OUT/$name.ncu-rep --page raw --csv > $OUT/$name.raw.csv 2>/dev/null
  ncu -i $OUT/$name.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_trim.py > $OUT/$name.sass.csv
  [ -z "$KEEP_REP" ] && rm -f $OUT/$name.ncu-rep
done
