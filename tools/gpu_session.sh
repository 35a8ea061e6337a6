#!/bin/bash
# One GPU-box session that produces the evidence kept under profiles/: parity tests, both bench arms, the ncu launch
# list of the bench command (with per-launch DRAM bytes) and one `--set full` capture of the dominant kernel.
# Usage (repo root on the GPU box):  bash tools/gpu_session.sh <tag>
TAG=${1:-sess}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_PYTEST" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
fi
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 600 $OUT/bench.json
if [ -z "$SKIP_REF" ]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"; cat $OUT/bench_reference.json
fi
# launch list of the bench command: duration and DRAM bytes of every launch (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $OUT/launches.csv)"
# full capture of the mod-up NTT at the bench batch (profile_target launch order: standalone forward NTT, then mod-up)
PROFILE_BATCH=256 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:NttFwdTile -s 1 -c 1 \
  -o $OUT/modup -f python tests/microbench/profile_target.py > $OUT/ncu_full.log 2>&1
echo "full capture rc=$?"
ncu -i $OUT/modup.ncu-rep --page raw --csv > $OUT/modup_raw.csv 2>/dev/null
# one tiny hot-path call checked against the oracle, as the driver runs it
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
