#!/bin/bash
# ncu --set full capture of one standalone forward and one inverse transform launch (N, BITS from the environment).
TAG=${1:-r2ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for k in ${KERNELS:-NttFwdTile2 NttInvTile2}; do
  PROFILE_N=${N:-16384} PROFILE_BITS=${BITS:-40} timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:$k -s 2 -c 1 -o $OUT/$k -f python tests/microbench/profile_ntt.py > $OUT/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
