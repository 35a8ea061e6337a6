#!/bin/sh
# TEST INFRASTRUCTURE ONLY: builds tests/emul/libtroy_b200_emul.so — the library's kernel functors re-targeted to
# sequential CPU loops (TROYB_HOST_EMUL) — for debugging index math in `-m "not gpu"` tests. Never shipped.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/../../troy_b200/csrc
g++ -std=c++17 -O2 -ffp-contract=off -fPIC -shared -DTROYB_HOST_EMUL -D__CUDACC__ -I"$HERE" -x c++ "$SRC/api.cu" -x c++ "$SRC/ntt_fwd.cu" -x c++ "$SRC/ntt_inv.cu" -x c++ "$SRC/context.cpp" \
    -o "$HERE/libtroy_b200_emul.so"
