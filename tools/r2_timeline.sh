mkdir -p gpurun_out/r2g
TROYB_TIMELINE=gpurun_out/r2g/timeline_fwd_n16384_40bit.txt PROFILE_N=16384 PROFILE_BITS=40 python tests/microbench/profile_ntt.py
TROYB_TIMELINE=gpurun_out/r2g/timeline_fwd_n16384_60bit.txt PROFILE_N=16384 PROFILE_BITS=60 python tests/microbench/profile_ntt.py
ls -la gpurun_out/r2g
