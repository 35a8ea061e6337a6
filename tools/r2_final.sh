#!/bin/bash
# Final evidence session of round 2 (one GPU box): tools/r2_session.sh (parity tests, both bench arms, ncu launch list, full
# capture of the mod-up transform, compute-sanitizer) plus a full capture of the mod-down transform and the encryption timings.
TAG=${1:-r2z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
bash tools/r2_session.sh $TAG
# mod-down transform of the relinearisation (launch 8 of NttFwdTile in the profile target: six per iteration - plain, mod-up, mod-down, mod-up, mod-down, rescale)
PROFILE_BATCH=256 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:NttFwdTile -s 8 -c 1 \
  -o $OUT/moddown -f python tests/microbench/profile_target.py > $OUT/ncu_moddown.log 2>&1
echo "moddown capture rc=$?"
ncu -i $OUT/moddown.ncu-rep --page raw --csv > $OUT/moddown_raw.csv 2>/dev/null
ncu -i $OUT/moddown.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_trim.py > $OUT/moddown_sass.csv
rm -f $OUT/moddown.ncu-rep
timeout 600 python tests/microbench/encrypt_bench.py > $OUT/encrypt_bench.jsonl 2> $OUT/encrypt_bench.err; echo "encrypt bench rc=$?"
