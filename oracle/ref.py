"""TEST INFRASTRUCTURE ONLY. ctypes binding of oracle/_ref/libtroyref.so — the UNMODIFIED reference
troy:: CPU library (built by oracle/Makefile from /root/reference/src) behind oracle/ref_shim.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this.
The product package troy_b200 never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libtroyref.so")

BFV, CKKS, BGV = 1, 2, 3

(OP_MULTIPLY, OP_SQUARE, OP_RELIN, OP_RESCALE_NEXT, OP_MODSWITCH_NEXT, OP_ROTATE_ROWS, OP_ROTATE_COLUMNS,
 OP_ROTATE_VECTOR, OP_CONJUGATE, OP_APPLY_GALOIS, OP_TO_NTT, OP_FROM_NTT, OP_ADD, OP_SUB, OP_NEGATE) = range(1, 16)

RNS_FASTBCONV_MTILDE, RNS_SM_MRQ, RNS_FAST_FLOOR, RNS_FASTBCONV_SK, RNS_DIVROUND_LAST, RNS_DIVROUND_LAST_NTT, \
    RNS_MODT_DIV_LAST = range(7)

_lib = None


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_LIB_PATH)
        vp, u64p, i32p, dp = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_double)
        sig = {
            "ref_last_error": (C.c_char_p, []),
            "ref_coeff_modulus_create": (C.c_int, [C.c_uint64, i32p, C.c_int, u64p]),
            "ref_ctx_create": (vp, [C.c_int, C.c_uint64, u64p, C.c_int, C.c_uint64, u64p]),
            "ref_ctx_destroy": (None, [vp]),
            "ref_ctx_num_levels": (C.c_int, [vp]),
            "ref_ctx_level_moduli": (C.c_int, [vp, C.c_int, u64p]),
            "ref_ctx_using_keyswitching": (C.c_int, [vp]),
            "ref_ntt_tables": (C.c_int, [vp, C.c_int, C.c_int, u64p, u64p, u64p]),
            "ref_ntt": (C.c_int, [vp, C.c_int, C.c_int, u64p, C.c_int]),
            "ref_ntt_prime": (C.c_int, [C.c_uint64, C.c_uint64, u64p, C.c_int, C.c_int]),
            "ref_negacyclic_shift": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p]),
            "ref_multiply_inv_degree": (C.c_int, [C.c_uint64, C.c_uint64, u64p, u64p]),
            "ref_rns_info": (C.c_int, [vp, C.c_int, u64p, u64p]),
            "ref_rns_op": (C.c_int, [vp, C.c_int, C.c_int, u64p, u64p]),
            "ref_galois_elt_from_step": (C.c_uint32, [vp, C.c_int]),
            "ref_apply_galois_poly": (C.c_int, [vp, C.c_int, u64p, C.c_int, C.c_uint32, u64p, C.c_int]),
            "ref_pt_new": (vp, [vp, u64p, C.c_uint64, C.c_int, C.c_double]),
            "ref_pt_destroy": (None, [vp]),
            "ref_pt_count": (C.c_uint64, [vp]),
            "ref_pt_words": (C.c_int, [vp, u64p]),
            "ref_pt_to_ntt": (C.c_int, [vp, vp, C.c_int]),
            "ref_eval_plain": (C.c_int, [vp, C.c_int, vp, vp]),
            "ref_secret_key_words": (C.c_int, [vp, u64p]),
            "ref_decrypt_plain": (vp, [vp, vp]),
            "ref_parms_id": (C.c_int, [vp, C.c_int, u64p]),
            "ref_keygen": (C.c_int, [vp]),
            "ref_relin_keys": (vp, [vp]),
            "ref_galois_keys_elts": (vp, [vp, C.POINTER(C.c_uint32), C.c_int]),
            "ref_galois_keys_steps": (vp, [vp, i32p, C.c_int]),
            "ref_galois_keys_all": (vp, [vp]),
            "ref_keys_synthetic": (vp, [vp, C.c_int, C.c_int]),
            "ref_kswitch_key_set_words": (C.c_int, [vp, vp, C.c_int, u64p]),
            "ref_keys_destroy": (None, [vp]),
            "ref_keys_num_slots": (C.c_int, [vp]),
            "ref_kswitch_key_digits": (C.c_int, [vp, C.c_int]),
            "ref_kswitch_key_words": (C.c_int, [vp, vp, C.c_int, u64p]),
            "ref_ct_new": (vp, [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64, u64p]),
            "ref_ct_clone": (vp, [vp]),
            "ref_ct_destroy": (None, [vp]),
            "ref_ct_meta": (C.c_int, [vp, vp, u64p, dp]),
            "ref_ct_words": (C.c_int, [vp, vp, u64p]),
            "ref_eval": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_int64]),
            "ref_encrypt_batch": (vp, [vp, u64p, C.c_int]),
            "ref_decrypt_batch": (C.c_int, [vp, vp, u64p]),
            "ref_encrypt_ckks": (vp, [vp, dp, C.c_int, C.c_double]),
            "ref_decrypt_ckks": (C.c_int, [vp, vp, dp]),
            "ref_prng_bytes": (C.c_int, [u64p, C.c_uint64, C.POINTER(C.c_uint8)]),
            "ref_sample_poly": (C.c_int, [vp, C.c_int, C.c_int, u64p, C.c_uint64, u64p, u64p]),
            "ref_sample_poly_replay": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_uint64, C.c_uint64, u64p, u64p]),
            "ref_public_key_words": (C.c_int64, [vp, u64p]),
            "ref_encrypt_zero": (vp, [vp, C.c_int, C.c_int]),
            "ref_encrypt_plain": (vp, [vp, vp, C.c_int]),
            "ref_time_ops": (C.c_double, [vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int]),
            "ref_time_ntt": (C.c_double, [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64)) if a is not None else None


class RefError(RuntimeError):
    pass


def _err():
    return lib().ref_last_error().decode()


def coeff_modulus_create(n, bits):
    b = (C.c_int * len(bits))(*bits)
    out = np.zeros(len(bits), dtype=np.uint64)
    if lib().ref_coeff_modulus_create(n, b, len(bits), _p(out)) != 0:
        raise RefError(_err())
    return [int(x) for x in out]


class RefCiphertext:
    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    def __del__(self):
        if self.h:
            lib().ref_ct_destroy(self.h)
            self.h = None

    def meta(self):
        m = np.zeros(5, dtype=np.uint64)
        s = C.c_double()
        lib().ref_ct_meta(self.ctx.h, self.h, _p(m), C.byref(s))
        return dict(level=int(np.int64(m[0])), size=int(m[1]), limbs=int(m[2]), is_ntt=bool(m[3]),
                    correction_factor=int(m[4]), scale=s.value)

    def words(self):
        m = self.meta()
        out = np.zeros((m["size"], m["limbs"], self.ctx.n), dtype=np.uint64)
        lib().ref_ct_words(self.ctx.h, self.h, _p(out))
        return out

    def clone(self):
        return RefCiphertext(self.ctx, lib().ref_ct_clone(self.h))


class RefKeys:
    def __init__(self, ctx, handle):
        if not handle:
            raise RefError(_err())
        self.ctx, self.h = ctx, handle

    def __del__(self):
        if self.h:
            lib().ref_keys_destroy(self.h)
            self.h = None

    def num_slots(self):
        return lib().ref_keys_num_slots(self.h)

    def digits(self, index):
        return lib().ref_kswitch_key_digits(self.h, index)

    def words(self, index):
        """[digit][2][key_limbs][N] uint64, or None when the slot is empty."""
        d = self.digits(index)
        if d == 0:
            return None
        kl = self.ctx.level_moduli(-1)
        out = np.zeros((d, 2, len(kl), self.ctx.n), dtype=np.uint64)
        if lib().ref_kswitch_key_words(self.ctx.h, self.h, index, _p(out)) != 0:
            raise RefError(_err())
        return out

    def set_words(self, index, words):
        w = np.ascontiguousarray(words, dtype=np.uint64)
        if lib().ref_kswitch_key_set_words(self.ctx.h, self.h, index, _p(w)) != 0:
            raise RefError(_err())


PLAIN_MULTIPLY, PLAIN_ADD, PLAIN_SUB = 1, 2, 3


class RefPlaintext:
    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_pt_destroy(self.h)
            self.h = None

    def words(self):
        out = np.zeros(int(lib().ref_pt_count(self.h)), dtype=np.uint64)
        lib().ref_pt_words(self.h, _p(out))
        return out

    def to_ntt(self, level):
        """Evaluator::transformToNttInplace(Plaintext&, ParmsID)"""
        if lib().ref_pt_to_ntt(self.ctx.h, self.h, level) != 0:
            raise RefError(_err())
        return self


class RefContext:
    """The reference SEALContext + Evaluator (+ KeyGenerator/Encryptor/Decryptor/encoders after keygen())."""

    def __init__(self, scheme, n, moduli, plain_modulus=0, seed=(1, 2, 3, 4, 5, 6, 7, 8)):
        self.scheme, self.n, self.moduli, self.plain_modulus = scheme, n, list(moduli), plain_modulus
        q = np.array(moduli, dtype=np.uint64)
        s = np.array(seed, dtype=np.uint64) if seed is not None else None
        self.h = lib().ref_ctx_create(scheme, n, _p(q), len(moduli), plain_modulus, _p(s))
        if not self.h:
            raise RefError(_err())

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_ctx_destroy(self.h)
            self.h = None

    def num_levels(self):
        return lib().ref_ctx_num_levels(self.h)

    def level_moduli(self, level):
        cnt = lib().ref_ctx_level_moduli(self.h, level, None)
        out = np.zeros(cnt, dtype=np.uint64)
        lib().ref_ctx_level_moduli(self.h, level, _p(out))
        return [int(x) for x in out]

    def ntt_tables(self, level, limb):
        root = np.zeros((self.n, 2), dtype=np.uint64)
        inv = np.zeros((self.n, 2), dtype=np.uint64)
        sc = np.zeros(4, dtype=np.uint64)
        if lib().ref_ntt_tables(self.h, level, limb, _p(root), _p(inv), _p(sc)) != 0:
            raise RefError(_err())
        return root, inv, sc

    def ntt(self, level, limb, data, inverse=False):
        d = np.ascontiguousarray(data, dtype=np.uint64).copy()
        if lib().ref_ntt(self.h, level, limb, _p(d), int(inverse)) != 0:
            raise RefError(_err())
        return d

    def rns_info(self, level):
        info = np.zeros(8, dtype=np.uint64)
        bsk = np.zeros(80, dtype=np.uint64)
        if lib().ref_rns_info(self.h, level, _p(info), _p(bsk)) != 0:
            raise RefError(_err())
        keys = ["q_size", "B_size", "Bsk_size", "Bsk_mtilde_size", "m_sk", "gamma", "m_tilde", "t"]
        d = {k: int(v) for k, v in zip(keys, info)}
        d["Bsk"] = [int(x) for x in bsk[: d["Bsk_size"]]]
        return d

    def rns_op(self, level, which, inp, out_limbs=None):
        a = np.ascontiguousarray(inp, dtype=np.uint64).copy()
        out = np.zeros((out_limbs, self.n), dtype=np.uint64) if out_limbs else None
        if lib().ref_rns_op(self.h, level, which, _p(a), _p(out)) != 0:
            raise RefError(_err())
        return out if out_limbs else a

    def galois_elt_from_step(self, step):
        return int(lib().ref_galois_elt_from_step(self.h, step))

    def apply_galois_poly(self, level, poly, elt, ntt_form):
        a = np.ascontiguousarray(poly, dtype=np.uint64)
        out = np.zeros_like(a)
        if lib().ref_apply_galois_poly(self.h, level, _p(a), a.shape[0], elt, _p(out), int(ntt_form)) != 0:
            raise RefError(_err())
        return out

    # -- keys
    def keygen(self):
        if lib().ref_keygen(self.h) != 0:
            raise RefError(_err())

    def relin_keys(self):
        return RefKeys(self, lib().ref_relin_keys(self.h))

    def galois_keys(self, elts=None, steps=None):
        if elts is not None:
            a = (C.c_uint32 * len(elts))(*elts)
            return RefKeys(self, lib().ref_galois_keys_elts(self.h, a, len(elts)))
        if steps is not None:
            a = (C.c_int * len(steps))(*steps)
            return RefKeys(self, lib().ref_galois_keys_steps(self.h, a, len(steps)))
        return RefKeys(self, lib().ref_galois_keys_all(self.h))

    def synthetic_keys(self, galois, n_slots):
        return RefKeys(self, lib().ref_keys_synthetic(self.h, int(galois), n_slots))

    # -- ciphertexts
    def ciphertext(self, words, level=0, is_ntt=False, scale=1.0, correction_factor=1):
        w = np.ascontiguousarray(words, dtype=np.uint64)
        h = lib().ref_ct_new(self.h, level, w.shape[0], int(is_ntt), scale, correction_factor, _p(w))
        if not h:
            raise RefError(_err())
        return RefCiphertext(self, h)

    # -- plaintexts (ntt_level None: coefficient form modulo t; else L*N words in NTT form at that level)
    def plaintext(self, words, ntt_level=None, scale=1.0):
        w = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1)
        h = lib().ref_pt_new(self.h, _p(w), w.size, -1 if ntt_level is None else ntt_level, scale)
        if not h:
            raise RefError(_err())
        return RefPlaintext(self, h)

    def eval_plain(self, op, a, pt):
        if lib().ref_eval_plain(self.h, op, a.h, pt.h) != 0:
            raise RefError(_err())
        return a

    def parms_id(self, level):
        """The reference's ParmsID of a level (-1: key level) as 32 bytes."""
        out = np.zeros(4, dtype=np.uint64)
        if lib().ref_parms_id(self.h, level, _p(out)) != 0:
            raise RefError(_err())
        return out.tobytes()

    def secret_key_words(self):
        """[key_limbs][N] NTT-form words of the secret key (after keygen())."""
        kl = len(self.level_moduli(-1))
        out = np.zeros((kl, self.n), dtype=np.uint64)
        if lib().ref_secret_key_words(self.h, _p(out)) != kl * self.n:
            raise RefError(_err())
        return out

    def decrypt_plain(self, ct):
        """Decryptor::decrypt -> RefPlaintext (coefficient form modulo t for BFV / BGV, NTT form for CKKS)."""
        h = lib().ref_decrypt_plain(self.h, ct.h)
        if not h:
            raise RefError(_err())
        return RefPlaintext(self, h)

    def eval(self, op, a, b=None, keys=None, arg=0):
        r = lib().ref_eval(self.h, op, a.h, b.h if b is not None else None, keys.h if keys is not None else None, arg)
        if r != 0:
            raise RefError(_err())
        return a

    def public_key_words(self):
        """[2][key_limbs][N] NTT-form words of the public key (after keygen())."""
        kl = len(self.level_moduli(-1))
        out = np.zeros((2, kl, self.n), dtype=np.uint64)
        if lib().ref_public_key_words(self.h, _p(out)) != out.size:
            raise RefError(_err())
        return out

    def sample_poly(self, level, which, seed, skip_bytes=0):
        """util::samplePolyTernary (0) / samplePolyCbd (1) / samplePolyUniform (2) from a fresh Blake2xbPRNG(seed) after
        skip_bytes of its stream: ([L][N] words, stream position afterwards)."""
        L = len(self.level_moduli(level))
        out = np.zeros((L, self.n), dtype=np.uint64)
        sd = np.ascontiguousarray(seed, dtype=np.uint64)
        pos = np.zeros(1, dtype=np.uint64)
        if lib().ref_sample_poly(self.h, level, which, _p(sd), skip_bytes, _p(out), _p(pos)) != 0:
            raise RefError(_err())
        return out, int(pos[0])

    def sample_poly_replay(self, level, which, stream, skip_bytes=0):
        """The same samplers drawing from the given byte stream (uint8 array; zeros past its end)."""
        L = len(self.level_moduli(level))
        out = np.zeros((L, self.n), dtype=np.uint64)
        st = np.ascontiguousarray(stream, dtype=np.uint8)
        pos = np.zeros(1, dtype=np.uint64)
        if lib().ref_sample_poly_replay(self.h, level, which, st.ctypes.data_as(C.POINTER(C.c_uint8)), st.size, skip_bytes,
                                        _p(out), _p(pos)) != 0:
            raise RefError(_err())
        return out, int(pos[0])

    def encrypt_zero(self, level=0, asymmetric=True):
        """Encryptor::encryptZero / encryptZeroSymmetric at a level (-1: key level), generator seeded as the context was."""
        h = lib().ref_encrypt_zero(self.h, level, int(asymmetric))
        if not h:
            raise RefError(_err())
        return RefCiphertext(self, h)

    def encrypt_plain(self, pt, asymmetric=True):
        h = lib().ref_encrypt_plain(self.h, pt.h, int(asymmetric))
        if not h:
            raise RefError(_err())
        return RefCiphertext(self, h)

    def encrypt_batch(self, values):
        v = np.ascontiguousarray(values, dtype=np.uint64)
        h = lib().ref_encrypt_batch(self.h, _p(v), len(v))
        if not h:
            raise RefError(_err())
        return RefCiphertext(self, h)

    def decrypt_batch(self, ct):
        out = np.zeros(self.n, dtype=np.uint64)
        r = lib().ref_decrypt_batch(self.h, ct.h, _p(out))
        if r < 0:
            raise RefError(_err())
        return out[:r]

    def encrypt_ckks(self, values, scale):
        v = np.ascontiguousarray(np.asarray(values, dtype=np.complex128))
        h = lib().ref_encrypt_ckks(self.h, v.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)), len(v), scale)
        if not h:
            raise RefError(_err())
        return RefCiphertext(self, h)

    def decrypt_ckks(self, ct):
        out = np.zeros(self.n, dtype=np.float64)
        r = lib().ref_decrypt_ckks(self.h, ct.h, out.ctypes.data_as(C.POINTER(C.c_double)))
        if r < 0:
            raise RefError(_err())
        return out.view(np.complex128)[:r]

    def time_ops(self, seq, a, b, rlk, gk, arg, reps, threads):
        t = lib().ref_time_ops(self.h, seq, a.h, b.h if b else None, rlk.h if rlk else None, gk.h if gk else None,
                               arg, reps, threads)
        if t < 0:
            raise RefError(_err())
        return t


def ntt_prime(n, prime, data, inverse=False, reps=1):
    d = np.ascontiguousarray(data, dtype=np.uint64).copy()
    if lib().ref_ntt_prime(n, prime, _p(d), int(inverse), reps) != 0:
        raise RefError(_err())
    return d


def negacyclic_shift(n, prime, shift, data):
    """data * X^shift modulo X^n + 1 and ``prime`` (src/utils/polyarithsmallmod.cpp:128-152)"""
    d = np.ascontiguousarray(data, dtype=np.uint64)
    out = np.empty_like(d)
    if lib().ref_negacyclic_shift(n, prime, shift, _p(d), _p(out)) != 0:
        raise RefError(_err())
    return out


def multiply_inv_degree(n, prime, data):
    """data * n^-1 modulo ``prime`` (EvaluatorCuda::divideByPolyModulusDegreeInplace per limb)"""
    d = np.ascontiguousarray(data, dtype=np.uint64)
    out = np.empty_like(d)
    if lib().ref_multiply_inv_degree(n, prime, _p(d), _p(out)) != 0:
        raise RefError(_err())
    return out


def time_ntt(n, prime, inverse, reps, threads):
    return lib().ref_time_ntt(n, prime, int(inverse), reps, threads)


def prng_bytes(seed, nbytes):
    """The first nbytes of Blake2xbPRNG(seed)'s stream (src/randomgen.cpp:164-194)."""
    sd = np.ascontiguousarray(seed, dtype=np.uint64)
    out = np.zeros(nbytes, dtype=np.uint8)
    if lib().ref_prng_bytes(_p(sd), nbytes, out.ctypes.data_as(C.POINTER(C.c_uint8))) != 0:
        raise RefError(_err())
    return out
