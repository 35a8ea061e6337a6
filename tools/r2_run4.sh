N=16384 BITS=60 KERNELS="NttFwdTile2" bash tools/r2_ncu_ntt.sh $1
N=16384 BITS=40 KERNELS="NttFwdTile2 NttInvTile2" bash tools/r2_ncu_ntt.sh $1
