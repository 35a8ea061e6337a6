#!/bin/bash
# Round-2 evidence session (one GPU box): tools/gpu_session.sh plus compute-sanitizer over every kernel family.
TAG=${1:-r2final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
bash tools/gpu_session.sh $TAG
rm -f $OUT/modup.ncu-rep.tmp
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tests/microbench/sanitizer_target.py > $OUT/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(grep -h 'ERROR SUMMARY\|RACECHECK SUMMARY' $OUT/sanitizer_$tool.log | tail -1)"
done
# keep the transfer small: the full report is reduced to CSV pages
ncu -i $OUT/modup.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_trim.py > $OUT/modup_sass.csv
rm -f $OUT/modup.ncu-rep
