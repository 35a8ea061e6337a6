mkdir -p gpurun_out/$1
for x in 0 40000 80000; do
TROYB_SMEM_EXTRA=$x TROYB_TIMELINE=gpurun_out/$1/timeline_40_x$x.txt PROFILE_N=16384 PROFILE_BITS=40 python tests/microbench/profile_ntt.py
done
