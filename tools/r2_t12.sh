for t in 0 1; do echo "T12=$t"; TROYB_T12=$t NTT_SIZES=8192,16384 python tests/microbench/ntt_bench.py | cut -d, -f1,2,3,5; done
TROYB_T12=1 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "ntt_sweep" 2>&1 | tail -2
