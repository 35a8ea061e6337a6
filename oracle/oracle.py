"""TEST INFRASTRUCTURE ONLY. ctypes binding of oracle/liboracle.so (the plain-C restatement in troy_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
BFV, CKKS, BGV = 1, 2, 3
_lib = None
_u64p = C.POINTER(C.c_uint64)


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_LIB_PATH)
        sz = C.c_size_t
        sig = {
            "orc_is_prime": (C.c_int, [C.c_uint64]),
            "orc_get_primes": (C.c_int, [C.c_uint64, C.c_int, C.c_int, _u64p]),
            "orc_coeff_modulus_create": (C.c_int, [C.c_uint64, C.POINTER(C.c_int), C.c_int, _u64p]),
            "orc_minimal_primitive_root": (C.c_uint64, [C.c_uint64, C.c_uint64]),
            "orc_ntt_tables": (None, [C.c_uint64, C.c_uint64, _u64p, _u64p]),
            "orc_ntt_forward": (None, [C.c_uint64, C.c_uint64, _u64p]),
            "orc_ntt_inverse": (None, [C.c_uint64, C.c_uint64, _u64p]),
            "orc_dyadic_tensor": (None, [C.c_uint64, C.c_uint64, _u64p, C.c_int, _u64p, C.c_int, _u64p, sz, sz, sz]),
            "orc_galois_coeff": (None, [C.c_uint64, C.c_uint64, _u64p, C.c_uint32, _u64p]),
            "orc_galois_ntt": (None, [C.c_uint64, _u64p, C.c_uint32, _u64p]),
            "orc_galois_elt_from_step": (C.c_uint32, [C.c_uint64, C.c_int]),
            "orc_drop_last": (None, [C.c_int, C.c_uint64, C.c_int, _u64p, C.c_uint64, _u64p]),
            "orc_switch_key": (None, [C.c_int, C.c_uint64, C.c_int, C.c_int, _u64p, C.c_uint64, _u64p, _u64p, _u64p]),
            "orc_behz_base": (C.c_int, [C.c_uint64, C.c_int, _u64p, C.c_uint64, _u64p]),
            "orc_bfv_multiply": (C.c_int, [C.c_uint64, C.c_int, _u64p, C.c_uint64, _u64p, C.c_int, _u64p, C.c_int, _u64p]),
            "orc_multiply_ntt_schemes": (None, [C.c_int, C.c_uint64, C.c_int, _u64p, _u64p, C.c_int, _u64p, C.c_int, _u64p]),
            "orc_apply_galois": (None, [C.c_int, C.c_uint64, C.c_int, C.c_int, _u64p, C.c_uint64, _u64p, C.c_uint32, _u64p]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_u64p)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def coeff_modulus_create(n, bits):
    b = (C.c_int * len(bits))(*bits)
    out = np.zeros(len(bits), dtype=np.uint64)
    if lib().orc_coeff_modulus_create(n, b, len(bits), _p(out)) != 0:
        raise RuntimeError("not enough primes")
    return [int(x) for x in out]


def get_primes(factor, bits, count):
    out = np.zeros(count, dtype=np.uint64)
    if lib().orc_get_primes(factor, bits, count, _p(out)) != 0:
        raise RuntimeError("not enough primes")
    return [int(x) for x in out]


def minimal_primitive_root(degree, q):
    return int(lib().orc_minimal_primitive_root(degree, q))


def ntt_tables(n, q):
    r, i = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
    lib().orc_ntt_tables(n, q, _p(r), _p(i))
    return r, i


def ntt_forward(q, a):
    d = _c(a).copy()
    lib().orc_ntt_forward(d.shape[-1], q, _p(d))
    return d


def ntt_inverse(q, a):
    d = _c(a).copy()
    lib().orc_ntt_inverse(d.shape[-1], q, _p(d))
    return d


def ntt_poly(moduli, poly, inverse=False):
    """poly [..., L, N] with limb l under moduli[l]"""
    d = _c(poly).copy()
    flat = d.reshape(-1, len(moduli), d.shape[-1])
    for x in flat:
        for l, q in enumerate(moduli):
            (lib().orc_ntt_inverse if inverse else lib().orc_ntt_forward)(d.shape[-1], q, _p(x[l]))
    return d


def galois_poly(moduli, poly, elt, ntt_form):
    a = _c(poly)
    out = np.zeros_like(a)
    n = a.shape[-1]
    for l, q in enumerate(moduli):
        if ntt_form:
            lib().orc_galois_ntt(n, _p(a[l]), elt, _p(out[l]))
        else:
            lib().orc_galois_coeff(n, q, _p(a[l]), elt, _p(out[l]))
    return out


def galois_elt_from_step(n, step):
    return int(lib().orc_galois_elt_from_step(n, step))


def drop_last(scheme, moduli, t, ct):
    """ct [size][L][N] -> [size][L-1][N] (rescale for CKKS, mod-switch-to-next for BFV/BGV)"""
    a = _c(ct).copy()
    q = np.array(moduli, dtype=np.uint64)
    for p in a:
        lib().orc_drop_last(scheme, a.shape[-1], len(moduli), _p(q), t, _p(p))
    return np.ascontiguousarray(a[:, :-1, :])


def switch_key(scheme, key_moduli, t, ct, target, key):
    """ct [2][L][N] (copy returned), target [L][N], key [Ltop][2][Ltop+1][N]"""
    c = _c(ct).copy()
    tg, k = _c(target), _c(key)
    kq = np.array(key_moduli, dtype=np.uint64)
    lib().orc_switch_key(scheme, c.shape[-1], c.shape[1], len(key_moduli) - 1, _p(kq), t, _p(c), _p(tg), _p(k))
    return c


def apply_galois(scheme, key_moduli, t, ct, elt, key):
    c = _c(ct).copy()
    k = _c(key)
    kq = np.array(key_moduli, dtype=np.uint64)
    lib().orc_apply_galois(scheme, c.shape[-1], c.shape[1], len(key_moduli) - 1, _p(kq), t, _p(c), elt, _p(k))
    return c


def multiply(scheme, moduli, t, a, b):
    a, b = _c(a), _c(b)
    q = np.array(moduli, dtype=np.uint64)
    out = np.zeros((a.shape[0] + b.shape[0] - 1,) + a.shape[1:], dtype=np.uint64)
    if scheme == BFV:
        if lib().orc_bfv_multiply(a.shape[-1], len(moduli), _p(q), t, _p(a), a.shape[0], _p(b), b.shape[0], _p(out)) != 0:
            raise RuntimeError("BEHZ base")
    else:
        lib().orc_multiply_ntt_schemes(scheme, a.shape[-1], len(moduli), _p(q), _p(a), a.shape[0], _p(b), b.shape[0], _p(out))
    return out


def behz_base(n, moduli, t):
    q = np.array(moduli, dtype=np.uint64)
    out = np.zeros(80, dtype=np.uint64)
    k = lib().orc_behz_base(n, len(moduli), _p(q), t, _p(out))
    return [int(x) for x in out[:k]]
