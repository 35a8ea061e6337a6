#!/bin/bash
# The reference's own timing program (test/timetest.cu, unchanged) on the same GPU with both backends: the reference's CUDA
# library (oracle/_ref/timetest_reference) and this library through the drop-in header (tests/cpp/_build/timetest_dropin).
# Its clock is the host's (gettimeofday around each call, no device synchronisation): it measures what a caller of the
# blocking-looking API waits for; calls that return before the device is done (both backends enqueue asynchronously) are
# paid by the next call that synchronises.
OUT=${1:-gpurun_out/timetest}
mkdir -p $OUT
timeout 900 ./oracle/_ref/timetest_reference > $OUT/timetest_reference.txt 2>&1; echo "reference rc=$?"
timeout 900 ./tests/cpp/_build/timetest_dropin > $OUT/timetest_dropin.txt 2>&1; echo "drop-in rc=$?"
paste $OUT/timetest_reference.txt $OUT/timetest_dropin.txt
