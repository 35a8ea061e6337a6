mkdir -p gpurun_out/$1
for x in 0 80000; do
AB_LIB=troy_b200/libtroy_b200_dev.so TROYB_SMEM_EXTRA=$x TROYB_TIMELINE=gpurun_out/$1/timeline_40_x$x.txt PROFILE_N=16384 PROFILE_BITS=40 python tests/microbench/profile_ntt.py
done
N=16384 BITS=40 KERNELS="NttFwdTile2" bash tools/r2_ncu_ntt.sh $1
