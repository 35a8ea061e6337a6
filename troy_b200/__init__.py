"""troy_b200 — B200-native RNS polynomial engine behind the troy (``troyn::``) API.

The product is ``troy_b200/libtroy_b200.so`` (hand-written sm_100a CUDA behind the C ABI in ``include/troy_b200.h``).
This package is the Python mirror of the reference's ``troyn::`` host classes for that path (same names, argument
meaning and error behaviour); it binds the C ABI with ctypes and owns no arithmetic. There is no CPU fallback: if the
CUDA library is missing the import of :mod:`troy_b200.lib` fails loudly.
"""
from .host import (  # noqa: F401
    BFV, BGV, CKKS,
    CiphertextCuda, CoeffModulus, DecryptorCuda, EncryptionParameters, EncryptorCuda, EvaluatorCuda, GaloisKeysCuda, KSwitchKeysCuda, KernelProvider,
    PlainModulus, PlaintextCuda, RelinKeysCuda, SEALContextCuda, SchemeType,
)
from .lib import TroyError, library  # noqa: F401
from .linear import Cipher2d, MatmulHelper, Plain2d  # noqa: F401
from . import serialize  # noqa: F401
