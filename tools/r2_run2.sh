mkdir -p gpurun_out/$1
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "${PYTEST_K:-ntt_sweep or evaluator_flow or batched or many_digits}" 2>&1 | tail -3
NTT_SIZES=${NTT_SIZES:-4096,8192,16384,32768,65536} python tests/microbench/ntt_bench.py | cut -d, -f1,2,3,4
TROYB_TIMELINE=gpurun_out/$1/timeline_40.txt PROFILE_N=16384 PROFILE_BITS=40 python tests/microbench/profile_ntt.py
