"""The reference's wire format for the carrier types of the path: byte-compatible save / load of
ciphertexts, plaintexts and key-switching keys, so existing troy users can exchange data with this library.

Formats (all little-endian raw members, no compression; `savet` = raw bytes of the member, src/serialize.h):
  CiphertextCuda::save    src/ciphertext_cuda.cu:16-43   parms_id[32] is_ntt_form[1] size[8] N[8] limbs[8] scale[f64]
                                                          correction_factor[8] seed[8] terms[1] data_size[8] data[8*data_size]
  PlaintextCuda::save     src/plaintext_cuda.cu:7-15     parms_id[32] coeff_count[8] scale[f64] data_size[8] data
  KSwitchKeysCuda::save   src/kswitchkeys_cuda.cuh:330-339 parms_id[32] n[8] { m[8] { CiphertextCuda::save x m } } x n
A parms_id is the BLAKE2b-256 hash of {scheme, N, coeff moduli of the level, plain modulus} as 64-bit words
(EncryptionParameters::computeParmsID, src/encryptionparams.cpp:118-156); this package names levels by index, so the hash
is computed on save and looked up on load. A seeded ciphertext (seed != 0: a fresh symmetric encryption) travels as c0 + seed;
its c1 is re-expanded on load with the reference's cuRAND sampler semantics (CiphertextCuda::load(stream, context),
src/ciphertext_cuda.cu:141-192; troyb_expand_seeded_poly). Termed ciphertexts are rejected here as CiphertextCuda::load does.
"""
import hashlib
import struct

import numpy as np

from .host import CKKS, CiphertextCuda, KSwitchKeysCuda, PlaintextCuda, SEALContextCuda
from .lib import TroyInvalidArgument

PARMS_ID_ZERO = bytes(32)


def parms_id(context, level):
    """The reference's ParmsID (4 x uint64) of a level of the chain (-1: the key level) as 32 bytes."""
    t = 0 if context.scheme == CKKS else int(context.parms.plain_modulus)
    words = [int(context.scheme), int(context.n)] + [int(m) for m in context.level_moduli(level)] + [t]
    return hashlib.blake2b(np.array(words, dtype=np.uint64).tobytes(), digest_size=32).digest()


def _level_of(context, pid):
    for level in [SEALContextCuda.KEY_LEVEL] + list(range(context.num_levels)):
        if parms_id(context, level) == pid:
            return level
    raise TroyInvalidArgument("parms_id is not valid for the current context")


def _read(stream, fmt):
    size = struct.calcsize(fmt)
    raw = stream.read(size)
    if len(raw) != size:
        raise TroyInvalidArgument("unexpected end of stream")
    return struct.unpack(fmt, raw)


def _save_ct_words(stream, pid, is_ntt, size, n, limbs, scale, correction_factor, words, seed=0):
    stream.write(pid)
    stream.write(struct.pack("<?QQQdQQ?Q", bool(is_ntt), size, n, limbs, float(scale), int(correction_factor), int(seed), False,
                             words.size))
    stream.write(np.ascontiguousarray(words, dtype=np.uint64).tobytes())


def _load_ct_words(stream):
    pid = stream.read(32)
    is_ntt, size, n, limbs, scale, cf, seed, terms, data_size = _read(stream, "<?QQQdQQ?Q")
    if terms:
        raise TroyInvalidArgument("Trying to load a termed ciphertext, but indices is not specified")
    if seed != 0 and size > 2:
        raise TroyInvalidArgument("Seed exists but size is not 2.")
    stored = (1 if seed != 0 else size) * n * limbs
    words = np.frombuffer(stream.read(8 * data_size), dtype=np.uint64)
    if words.size != data_size or data_size < stored:
        raise TroyInvalidArgument("unexpected end of stream")
    return pid, is_ntt, size, n, limbs, scale, cf, words[:stored], seed


def save_ciphertext(ct, stream, batch_index=0):
    """CiphertextCuda::save of one ciphertext of a batch."""
    words = ct.to_host()[batch_index]
    seed = 0 if ct.seeds is None else int(ct.seeds[batch_index])
    if seed != 0 and ct.size > 2:
        raise TroyInvalidArgument("Seed exists but size is not 2.")
    if seed != 0:
        words = words[:1]  # c0 only; c1 is the expansion of the seed
    _save_ct_words(stream, parms_id(ct.context, ct.parms_id), ct.is_ntt_form, ct.size, ct.context.n, ct.limbs, ct.scale,
                   ct.correction_factor, words.reshape(-1), seed)


def load_ciphertext(context, stream):
    """CiphertextCuda::load(stream, context)  src/ciphertext_cuda.cu:141-192."""
    import ctypes as C
    from .host import _L, _vp
    pid, is_ntt, size, n, limbs, scale, cf, words, seed = _load_ct_words(stream)
    level = _level_of(context, pid)
    if n != context.n or limbs != context.level_limbs(level):
        raise TroyInvalidArgument("encrypted is not valid for encryption parameters")
    if seed == 0:
        return CiphertextCuda.from_host(context, words.reshape(size, limbs, n), level, is_ntt, scale, cf)
    ct = CiphertextCuda(context, level, 2, is_ntt, scale, cf)
    ct.data.upload(words)
    d_seed = context.alloc(1)
    d_seed.upload(np.array([seed], dtype=np.uint64))
    _L().call("troyb_expand_seeded_poly", _vp(context.h), level, _vp(d_seed.ptr), _vp(ct.ptr(1)), ct.poly_words, 1, _vp(None))
    if context.scheme != CKKS:  # coefficient-form schemes (the reference's loader inverts for BFV only, src/ciphertext_cuda.cu:187-189,
        # which leaves its own BGV seeded ciphertexts undecryptable; that slip is not reproduced)
        name = "troyb_ntt_inverse_key" if level == SEALContextCuda.KEY_LEVEL else "troyb_ntt_inverse"
        args = (_vp(context.h),) + (() if level == SEALContextCuda.KEY_LEVEL else (level,))
        _L().call(name, *args, _vp(ct.ptr(1)), 1, _vp(None))
    _L().call("troyb_stream_synchronize", _vp(None))
    return ct


def save_plaintext(pt, stream, batch_index=0):
    """PlaintextCuda::save."""
    words = pt.to_host()[batch_index]
    pid = PARMS_ID_ZERO if pt.parms_id is None else parms_id(pt.context, pt.parms_id)
    stream.write(pid)
    stream.write(struct.pack("<QdQ", pt.coeff_count, float(pt.scale), words.size))
    stream.write(words.tobytes())


def load_plaintext(context, stream):
    """PlaintextCuda::load."""
    pid = stream.read(32)
    coeff_count, scale, data_size = _read(stream, "<QdQ")
    words = np.frombuffer(stream.read(8 * data_size), dtype=np.uint64)
    if words.size != data_size or data_size < coeff_count:
        raise TroyInvalidArgument("unexpected end of stream")
    level = None if pid == PARMS_ID_ZERO else _level_of(context, pid)
    w = words[:coeff_count]
    return PlaintextCuda.from_host(context, w if level is None else w.reshape(context.level_limbs(level), context.n), level, scale)


def save_kswitch_keys(keys, stream):
    """KSwitchKeysCuda::save (RelinKeys / GaloisKeys): every digit of every present key is one size-2 ciphertext at the key
    level, NTT form."""
    ctx = keys.context
    pid = parms_id(ctx, SEALContextCuda.KEY_LEVEL)
    kl, n = ctx.level_limbs(SEALContextCuda.KEY_LEVEL), ctx.n
    stream.write(pid)
    stream.write(struct.pack("<Q", len(keys.keys)))
    for k in keys.keys:
        if k is None:
            stream.write(struct.pack("<Q", 0))
            continue
        per = 2 * kl * n
        digits = k.words // per
        stream.write(struct.pack("<Q", digits))
        w = k.download(digits * per).reshape(digits, per)
        for j in range(digits):
            _save_ct_words(stream, pid, True, 2, n, kl, 1.0, 1, w[j])


def load_kswitch_keys(context, stream, cls=KSwitchKeysCuda):
    """KSwitchKeysCuda::load into ``cls`` (RelinKeysCuda / GaloisKeysCuda)."""
    pid = stream.read(32)
    if _level_of(context, pid) != SEALContextCuda.KEY_LEVEL:
        raise TroyInvalidArgument("kswitch_keys is not valid for encryption parameters")
    (count,) = _read(stream, "<Q")
    by_index = {}
    for i in range(count):
        (digits,) = _read(stream, "<Q")
        if digits == 0:
            continue
        rows = []
        for _ in range(digits):
            dpid, is_ntt, size, n, limbs, _, _, words, _ = _load_ct_words(stream)
            if dpid != pid or size != 2 or not is_ntt:
                raise TroyInvalidArgument("kswitch_keys is not valid for encryption parameters")
            rows.append(words.reshape(2, limbs, n))
        by_index[i] = np.stack(rows)
    if not by_index:
        return cls(context)
    k = cls.from_host(context, by_index)
    if len(k.keys) < count:
        k.keys += [None] * (count - len(k.keys))
    return k
