"""Host-side mirror of the reference's ``troyn::`` classes for the RNS polynomial path.

Names, argument meaning and error behaviour follow the reference (file:line cited per class); the arithmetic is all in
the CUDA library behind the C ABI. Differences that are deliberate and documented in DESIGN.md:

* a ``ParmsID`` is the chain index (``int``; ``-1`` is the key level) instead of a Blake2 hash of the parameters;
* every ciphertext object may carry a leading ``batch`` dimension (``[batch][poly][limb][coeff]``); ``batch == 1`` is
  exactly the reference object, larger batches are how one GPU is kept busy with independent ciphertexts;
* key generation and encoding stay on the host in the reference (``KeyGeneratorCuda`` wraps the host ``KeyGenerator``,
  src/keygenerator_cuda.cuh:36-60), so keys enter through ``from_host`` uploads; ``EncryptorCuda`` encrypts on the device
  with the words the reference's CPU encryptor gives for the same seed.
"""
import ctypes as C
import math
import os

import numpy as np

from . import lib as _lib
from .lib import TroyInvalidArgument, TroyLogicError, TroyOutOfRange

BFV, CKKS, BGV = 1, 2, 3


class SchemeType:
    """src/encryptionparams.h SchemeType."""
    none, bfv, ckks, bgv = 0, 1, 2, 3


def _L():
    return _lib.library()


def _vp(x):
    return C.c_void_p(x)


class KernelProvider:
    """src/kernelprovider.cuh:18-85."""

    @staticmethod
    def initialize(device=0):
        _L().call("troyb_initialize", device)

    @staticmethod
    def initialized():
        return bool(_L().cdll.troyb_initialized())


class CoeffModulus:
    """src/modulus.h:412-505."""

    @staticmethod
    def Create(poly_modulus_degree, bit_sizes):
        b = (C.c_int * len(bit_sizes))(*bit_sizes)
        out = np.zeros(len(bit_sizes), dtype=np.uint64)
        _L().call("troyb_coeff_modulus_create", int(poly_modulus_degree), b, len(bit_sizes), out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return [int(x) for x in out]

    create = Create


class PlainModulus:
    """src/modulus.h:511-552."""

    @staticmethod
    def Batching(poly_modulus_degree, bit_size):
        return CoeffModulus.Create(poly_modulus_degree, [bit_size])[0]

    batching = Batching


class DeviceArray:
    """Owning device buffer of uint64 words (DeviceArray<uint64_t>, src/utils/devicearray.cuh:90-200).

    Stream discipline: the buffer is allocated on ``stream`` (stream-ordered allocator) and released on the stream that used
    it LAST (``self.stream``; ``EvaluatorCuda`` re-tags every buffer it reads or writes with its own stream), so that the free
    is ordered behind the kernels still reading it. ``device`` is the CUDA device the buffer lives on (None: the calling
    thread's current device)."""

    def __init__(self, words, stream=None, device=None):
        self.words = int(words)
        self.stream = stream
        self.device = device
        if device is not None:
            _L().call("troyb_set_device", int(device))
        p = C.c_void_p()
        _L().call("troyb_malloc", C.byref(p), self.words * 8, _vp(stream))
        self.ptr = p.value

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                if self.device is not None:
                    _L().cdll.troyb_set_device(int(self.device))
                _L().cdll.troyb_free(_vp(self.ptr), _vp(self.stream))
                self.ptr = None
        except Exception:
            pass

    def at(self, word_offset):
        return self.ptr + 8 * int(word_offset)

    def _use_device(self):
        if self.device is not None:
            _L().call("troyb_set_device", int(self.device))

    def upload(self, arr, word_offset=0):
        a = np.ascontiguousarray(arr, dtype=np.uint64)
        if a.size + word_offset > self.words:
            raise TroyInvalidArgument("upload larger than buffer")
        self._use_device()
        _L().call("troyb_copy_to_device", _vp(self.at(word_offset)), a.ctypes.data_as(C.c_void_p), a.size * 8,
                  _vp(self.stream))
        _L().call("troyb_stream_synchronize", _vp(self.stream))

    def download(self, words=None, word_offset=0):
        n = self.words - word_offset if words is None else int(words)
        out = np.empty(n, dtype=np.uint64)
        self._use_device()
        _L().call("troyb_copy_to_host", out.ctypes.data_as(C.c_void_p), _vp(self.at(word_offset)), n * 8, _vp(self.stream))
        return out


class ExternalArray:
    """Non-owning view of device memory owned by someone else (DevicePointer<uint64_t>, src/utils/devicearray.cuh:16-55);
    used to run the evaluator directly on torch tensors."""

    def __init__(self, ptr, words):
        self.ptr, self.words, self.stream, self.device = int(ptr), int(words), None, None

    def at(self, word_offset):
        return self.ptr + 8 * int(word_offset)


class EncryptionParameters:
    """src/encryptionparams.h:90-260 (only what the path needs: scheme, N, coeff_modulus, plain_modulus)."""

    def __init__(self, scheme=SchemeType.none):
        self.scheme = int(scheme)
        self.poly_modulus_degree = 0
        self.coeff_modulus = []
        self.plain_modulus = 0

    def set_poly_modulus_degree(self, n):
        self.poly_modulus_degree = int(n)

    def set_coeff_modulus(self, moduli):
        self.coeff_modulus = [int(m) for m in moduli]

    def set_plain_modulus(self, t):
        if self.scheme == SchemeType.ckks and int(t) != 0:
            raise TroyLogicError("unsupported scheme")
        self.plain_modulus = int(t)

    setPolyModulusDegree, setCoeffModulus, setPlainModulus = set_poly_modulus_degree, set_coeff_modulus, set_plain_modulus


class SEALContextCuda:
    """src/context_cuda.cuh:146-186 + src/context_cuda.cu:5-62. ``parms_id`` values are chain indices."""

    KEY_LEVEL = -1

    def __init__(self, parms, expand_mod_chain=True):
        self.parms = parms
        q = np.array(parms.coeff_modulus, dtype=np.uint64)
        h = C.c_void_p()
        _L().call("troyb_context_create", parms.scheme, parms.poly_modulus_degree,
                  q.ctypes.data_as(C.POINTER(C.c_uint64)), len(parms.coeff_modulus), parms.plain_modulus,
                  int(bool(expand_mod_chain)), C.byref(h))
        self.h = h.value
        self.scheme = parms.scheme
        self.n = parms.poly_modulus_degree
        self.num_levels = _L().cdll.troyb_context_num_levels(_vp(self.h))
        self.device = _L().cdll.troyb_context_device(_vp(self.h))

    def alloc(self, words, stream=None):
        """Device words on this context's device, allocated in the order of ``stream``."""
        return DeviceArray(words, stream, self.device)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _L().cdll.troyb_context_destroy(_vp(self.h))
                self.h = None
        except Exception:
            pass

    def level_limbs(self, level):
        r = _L().cdll.troyb_context_level_limbs(_vp(self.h), level)
        if r < 0:
            raise TroyInvalidArgument(_L().last_error())
        return r

    def level_moduli(self, level):
        n = self.level_limbs(level)
        out = np.zeros(n, dtype=np.uint64)
        _L().call("troyb_context_level_moduli", _vp(self.h), level, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return [int(x) for x in out]

    def total_coeff_modulus_bit_count(self, level):
        return _L().cdll.troyb_context_total_coeff_bits(_vp(self.h), level)

    def using_keyswitching(self):
        return bool(_L().cdll.troyb_context_using_keyswitching(_vp(self.h)))

    def first_parms_id(self):
        return 0

    def last_parms_id(self):
        return self.num_levels - 1

    def key_parms_id(self):
        return self.KEY_LEVEL if self.using_keyswitching() else 0

    def bsk(self, level):
        out = np.zeros(80, dtype=np.uint64)
        n = _L().cdll.troyb_context_bsk(_vp(self.h), level, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return [int(x) for x in out[:max(n, 0)]]

    def root_powers(self, key_limb):
        out = np.zeros((self.n, 2), dtype=np.uint64)
        _L().call("troyb_context_root_powers", _vp(self.h), key_limb, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out

    def galois_elt_from_step(self, step):
        e = C.c_uint32()
        _L().call("troyb_galois_elt_from_step", _vp(self.h), int(step), C.byref(e))
        return e.value


class CiphertextCuda:
    """src/ciphertext_cuda.cuh:12-268: metadata + device words ``[batch][poly][limb][coeff]``."""

    def __init__(self, context, level=0, size=2, is_ntt_form=False, scale=1.0, correction_factor=1, batch=1,
                 capacity=None, data=None, stream=None):
        self.context = context
        self.parms_id = int(level)
        self.size = int(size)
        self.is_ntt_form = bool(is_ntt_form)
        self.scale = float(scale)
        self.correction_factor = int(correction_factor)
        self.batch = int(batch)
        self.limbs = context.level_limbs(level)
        self.capacity = max(int(capacity or 0), self.size)
        self.data = data if data is not None else context.alloc(self.batch * self.capacity * self.limbs * context.n, stream)
        # CiphertextCuda::seed  src/ciphertext_cuda.cuh:189: [batch] 64-bit seeds of a FRESH symmetric encryption whose c1 is the
        # cuRAND expansion of the seed (serialize.save_ciphertext then stores c0 only); None otherwise. Every evaluator call
        # that takes the ciphertext clears it (the reference's `seed() = 0`).
        self.seeds = None

    # -- layout
    @property
    def poly_words(self):
        return self.limbs * self.context.n

    @property
    def stride(self):
        """words between consecutive batch elements"""
        return self.capacity * self.poly_words

    def poly_modulus_degree(self):
        return self.context.n

    def coeff_modulus_size(self):
        return self.limbs

    def ptr(self, poly=0, batch_index=0):
        return self.data.at(batch_index * self.stride + poly * self.poly_words)

    # -- host <-> device (CiphertextCuda(const Ciphertext&) / toHost, src/ciphertext_cuda.cuh:20-51)
    @classmethod
    def from_host(cls, context, words, level=0, is_ntt_form=False, scale=1.0, correction_factor=1, capacity=None):
        w = np.ascontiguousarray(words, dtype=np.uint64)
        if w.ndim == 3:
            w = w[None]
        if w.ndim != 4 or w.shape[2] != context.level_limbs(level) or w.shape[3] != context.n:
            raise TroyInvalidArgument("encrypted is not valid for encryption parameters")
        ct = cls(context, level, w.shape[1], is_ntt_form, scale, correction_factor, batch=w.shape[0], capacity=capacity)
        for b in range(ct.batch):
            ct.data.upload(w[b], b * ct.stride)
        return ct

    @classmethod
    def from_device(cls, context, ptr, batch, size, level=0, is_ntt_form=False, scale=1.0, correction_factor=1,
                    capacity=None):
        cap = max(int(capacity or 0), size)
        words = batch * cap * context.level_limbs(level) * context.n
        return cls(context, level, size, is_ntt_form, scale, correction_factor, batch=batch, capacity=cap,
                   data=ExternalArray(ptr, words))

    @property
    def stream(self):
        """the stream that last produced or consumed the words (see DeviceArray)"""
        return self.data.stream

    def to_host(self):
        """Copies on the stream that last touched the words and waits for it."""
        out = np.empty((self.batch, self.size, self.limbs, self.context.n), dtype=np.uint64)
        _L().call("troyb_set_device", int(self.context.device))
        for b in range(self.batch):
            buf = np.empty(self.size * self.poly_words, dtype=np.uint64)
            _L().call("troyb_copy_to_host", buf.ctypes.data_as(C.c_void_p), _vp(self.ptr(0, b)), buf.size * 8, _vp(self.stream))
            out[b] = buf.reshape(self.size, self.limbs, self.context.n)
        return out

    def clone(self, stream=None):
        """A copy made in the order of ``stream`` (default: the stream that last touched this ciphertext)."""
        s = self.stream if stream is None else stream
        c = CiphertextCuda(self.context, self.parms_id, self.size, self.is_ntt_form, self.scale, self.correction_factor,
                           self.batch, self.capacity, stream=s)
        _L().call("troyb_copy_on_device", _vp(c.data.ptr), _vp(self.data.ptr), self.batch * self.stride * 8, _vp(s))
        return c

    def reserve(self, capacity, stream=None):
        """src/ciphertext_cuda.cuh:68-152: grow the per-ciphertext capacity, preserving contents (in the order of ``stream``)."""
        if capacity <= self.capacity:
            return
        s = self.stream if stream is None else stream
        new = self.context.alloc(self.batch * capacity * self.poly_words, s)
        for b in range(self.batch):
            _L().call("troyb_copy_on_device", _vp(new.at(b * capacity * self.poly_words)), _vp(self.ptr(0, b)),
                      self.size * self.poly_words * 8, _vp(s))
        if isinstance(self.data, DeviceArray):
            self.data.stream = s  # the old words are released behind the copies that read them
        self.data, self.capacity = new, capacity

    def resize(self, size, stream=None):
        if size < 2 and size != 0:
            raise TroyInvalidArgument("invalid size")
        self.reserve(size, stream)
        self.size = size


class PlaintextCuda:
    """src/plaintext_cuda.cuh: device plaintext. Coefficient form (``parms_id is None``, the reference's parmsIDZero):
    ``coeff_count`` words modulo t. NTT form: ``[L][N]`` words over the limbs of level ``parms_id`` (CKKS encodings, or a
    BFV / BGV plaintext after ``EvaluatorCuda.transform_plain_to_ntt_inplace``). ``batch`` > 1 holds that many plaintexts
    back to back (one per ciphertext of a batched call); batch = 1 is shared by every ciphertext of a batch."""

    def __init__(self, context, coeff_count, parms_id=None, scale=1.0, batch=1, data=None):
        self.context, self.coeff_count, self.parms_id, self.scale, self.batch = context, int(coeff_count), parms_id, scale, batch
        self.data = data if data is not None else context.alloc(batch * self.coeff_count)

    @property
    def is_ntt_form(self):
        return self.parms_id is not None

    @property
    def stride(self):
        """words between the plaintexts of a batch as the C ABI wants it (0: one plaintext for all)"""
        return self.coeff_count if self.batch > 1 else 0

    @classmethod
    def from_host(cls, context, words, parms_id=None, scale=1.0):
        """words: [coeff_count] / [batch][coeff_count] (coefficient form) or [L][N] / [batch][L][N] (NTT form)."""
        w = np.ascontiguousarray(words, dtype=np.uint64)
        unit_dims = 1 if parms_id is None else 2
        batch = w.shape[0] if w.ndim == unit_dims + 1 else 1
        count = w.size // batch
        if parms_id is not None and count != context.level_limbs(parms_id) * context.n:
            raise TroyInvalidArgument("plain is not valid for encryption parameters")
        if parms_id is None and count > context.n:
            raise TroyInvalidArgument("plain is not valid for encryption parameters")
        if count & 1 and batch > 1:
            raise TroyInvalidArgument("batched plaintexts need an even coefficient count")
        pt = cls(context, count, parms_id, scale, batch)
        pt.data.upload(w.reshape(-1))
        return pt

    @classmethod
    def from_device(cls, context, ptr, coeff_count, parms_id=None, scale=1.0, batch=1):
        return cls(context, coeff_count, parms_id, scale, batch, data=ExternalArray(ptr, batch * coeff_count))

    def to_host(self):
        return self.data.download(self.batch * self.coeff_count).reshape(self.batch, -1)

    isNttForm = is_ntt_form


class KSwitchKeysCuda:
    """src/kswitchkeys_cuda.cuh:31-366: ``keys[index]`` = one key-switching key, device words
    ``[digit][2][key_limbs][N]`` in NTT form at the key level."""

    def __init__(self, context):
        self.context = context
        self.parms_id = context.key_parms_id()
        self.keys = []

    @classmethod
    def from_host(cls, context, words_by_index):
        """words_by_index: {index: ndarray [digits][2][key_limbs][N]} (RelinKeysCuda(const RelinKeys&) upload)."""
        k = cls(context)
        n_key_limbs = context.level_limbs(SEALContextCuda.KEY_LEVEL)
        digits = context.level_limbs(0)
        size = max(words_by_index) + 1 if words_by_index else 0
        k.keys = [None] * size
        for idx, w in words_by_index.items():
            a = np.ascontiguousarray(w, dtype=np.uint64)
            if a.shape != (digits, 2, n_key_limbs, context.n):
                raise TroyInvalidArgument("kswitch_keys is not valid for encryption parameters")
            buf = context.alloc(a.size)
            buf.upload(a)
            k.keys[idx] = buf
        return k

    def data(self, index):
        if index >= len(self.keys) or self.keys[index] is None:
            raise TroyOutOfRange("kswitch_keys_index")
        return self.keys[index]


class RelinKeysCuda(KSwitchKeysCuda):
    """src/relinkeys_cuda.cuh."""

    @staticmethod
    def get_index(key_power):
        if key_power < 2:
            raise TroyInvalidArgument("key_power cannot be less than 2")
        return key_power - 2

    def has_key(self, key_power):
        i = self.get_index(key_power)
        return i < len(self.keys) and self.keys[i] is not None

    def key(self, key_power):
        return self.data(self.get_index(key_power))


class GaloisKeysCuda(KSwitchKeysCuda):
    """src/galoiskeys_cuda.cuh."""

    @staticmethod
    def get_index(galois_elt):
        if not (galois_elt & 1) or galois_elt < 3:
            raise TroyInvalidArgument("galois_elt is not valid")
        return (galois_elt - 1) >> 1

    def has_key(self, galois_elt):
        i = self.get_index(galois_elt)
        return i < len(self.keys) and self.keys[i] is not None

    def key(self, galois_elt):
        return self.data(self.get_index(galois_elt))


class DecryptorCuda:
    """src/decryptor_cuda.cuh / .cu: decryption on the device for the three schemes. ``secret_key_words``: ``[key_limbs][N]`` NTT-form words of the secret key at the key
    level, as ``KeyGenerator`` stores it. A batched ciphertext decrypts to a batched plaintext."""

    def __init__(self, context, secret_key_words, stream=None):
        self.context, self.stream = context, stream
        w = np.ascontiguousarray(secret_key_words, dtype=np.uint64)
        self.key_words = context.level_limbs(SEALContextCuda.KEY_LEVEL) * context.n
        if w.size != self.key_words:
            raise TroyInvalidArgument("secret key is not valid for encryption parameters")
        self.powers = context.alloc(self.key_words, stream)
        self.powers.upload(w.reshape(-1))
        self.n_powers = 1

    def _ensure_powers(self, max_power):
        """computeSecretKeyArray (src/decryptor_cuda.cu:188-235)"""
        if max_power <= self.n_powers:
            return
        new = self.context.alloc(max_power * self.key_words, self.stream)
        _L().call("troyb_secret_key_powers", _vp(self.context.h), _vp(self.powers.ptr), max_power, _vp(new.ptr), _vp(self.stream))
        self.powers.stream = self.stream  # released behind the kernel that read it
        self.powers, self.n_powers = new, max_power

    def decrypt(self, ct):
        scheme = self.context.scheme
        if ct.size < 2:
            raise TroyInvalidArgument("encrypted is empty")
        if scheme == CKKS and not ct.is_ntt_form:
            raise TroyInvalidArgument("encrypted must be in NTT form")
        if scheme != CKKS and ct.is_ntt_form:
            raise TroyInvalidArgument("encrypted cannot be in NTT form")
        self._ensure_powers(ct.size - 1)
        words = ct.limbs * self.context.n
        if isinstance(ct.data, DeviceArray) and ct.data.stream is not self.stream:
            # the ciphertext was last touched on another stream: wait for it, then it belongs to this one
            _L().call("troyb_stream_synchronize", _vp(ct.data.stream))
            ct.data.stream = self.stream
        phase = self.context.alloc(ct.batch * words, self.stream)
        _L().call("troyb_decrypt_dot_product", _vp(self.context.h), ct.parms_id, _vp(ct.ptr()), ct.size, ct.stride,
                  int(ct.is_ntt_form), _vp(self.powers.ptr), _vp(phase.ptr), ct.batch, _vp(self.stream))
        if scheme == CKKS:
            return PlaintextCuda(self.context, words, ct.parms_id, ct.scale, ct.batch, data=phase)
        out = PlaintextCuda(self.context, self.context.n, None, 1.0, ct.batch,
                            data=self.context.alloc(ct.batch * self.context.n, self.stream))
        if scheme == BGV:
            t = self.context.parms.plain_modulus
            try:
                fix = pow(ct.correction_factor, -1, t) if ct.correction_factor != 1 else 1
            except ValueError:
                raise TroyLogicError("invalid correction factor")
            _L().call("troyb_decrypt_mod_t", _vp(self.context.h), ct.parms_id, _vp(phase.ptr), _vp(out.data.ptr), C.c_uint64(fix),
                      ct.batch, _vp(self.stream))
            return out
        _L().call("troyb_decrypt_scale_and_round", _vp(self.context.h), ct.parms_id, _vp(phase.ptr), _vp(out.data.ptr), ct.batch,
                  _vp(self.stream))
        return out


class EncryptorCuda:
    """src/encryptor_cuda.cuh / src/encryptor.cpp: encryption on the device with the CPU path's generator semantics.

    ``public_key_words``: ``[2][key_limbs][N]``, ``secret_key_words``: ``[key_limbs][N]`` (NTT form at the key level, as
    ``KeyGenerator`` stores them); either may be None. Every ciphertext draws from its own Blake2xb generator, as
    ``parms.randomGenerator()->create()`` gives the reference (src/utils/rlwe.cpp:183, 284): ``seeds`` is ``[batch][8]`` words
    (64 bytes each); None takes them from ``os.urandom`` like the reference's default factory (src/randomgen.cpp:20-36). For a
    given seed the words equal those of the reference's CPU encryptor."""

    def __init__(self, context, public_key_words=None, secret_key_words=None, stream=None):
        self.context, self.stream = context, stream
        kl = context.level_limbs(SEALContextCuda.KEY_LEVEL)
        self.pk = self.sk = None
        if public_key_words is not None:
            w = np.ascontiguousarray(public_key_words, dtype=np.uint64)
            if w.size != 2 * kl * context.n:
                raise TroyInvalidArgument("public key is not valid for encryption parameters")
            self.pk = context.alloc(w.size, stream)
            self.pk.upload(w.reshape(-1))
        if secret_key_words is not None:
            w = np.ascontiguousarray(secret_key_words, dtype=np.uint64)
            if w.size != kl * context.n:
                raise TroyInvalidArgument("secret key is not valid for encryption parameters")
            self.sk = context.alloc(w.size, stream)
            self.sk.upload(w.reshape(-1))

    def _seeds(self, seeds, batch):
        if seeds is None:
            w = np.frombuffer(os.urandom(64 * batch), dtype=np.uint64)
        else:
            w = np.ascontiguousarray(seeds, dtype=np.uint64).reshape(-1)
        if w.size != 8 * batch:
            raise TroyInvalidArgument("one 64-byte seed per ciphertext")
        d = self.context.alloc(w.size, self.stream)
        d.upload(w)
        return d

    def _encrypt_zero(self, level, asymmetric, batch, seeds, save_seed=False):
        """Encryptor::encryptZeroInternal  src/encryptor.cpp:88-164. ``save_seed`` (symmetric only): the seeded form of the
        reference's CUDA encryptor (src/utils/rlwe_cuda.cu:261-333) - c1 is the cuRAND expansion of a 64-bit seed kept in
        ``ct.seeds``, so that the ciphertext travels as c0 + seed."""
        ctx = self.context
        key = self.pk if asymmetric else self.sk
        if key is None:
            raise TroyLogicError("public key is not set" if asymmetric else "secret key is not set")
        if level != SEALContextCuda.KEY_LEVEL and not 0 <= level < ctx.num_levels:
            raise TroyInvalidArgument("parms_id is not valid for encryption parameters")
        d_seeds = self._seeds(seeds, batch)
        ct = CiphertextCuda(ctx, level, 2, ctx.scheme == CKKS, 1.0, 1, batch=batch, stream=self.stream)
        if save_seed and not asymmetric:
            d_ct_seeds = ctx.alloc(batch, self.stream)
            _L().call("troyb_encrypt_zero_symmetric_seeded", _vp(ctx.h), level, _vp(key.ptr), _vp(d_seeds.ptr), _vp(ct.data.ptr),
                      ct.stride, _vp(d_ct_seeds.ptr), batch, _vp(self.stream))
            ct.seeds = d_ct_seeds.download()
            return ct
        name = "troyb_encrypt_zero_asymmetric" if asymmetric else "troyb_encrypt_zero_symmetric"
        _L().call(name, _vp(ctx.h), level, _vp(key.ptr), _vp(d_seeds.ptr), _vp(ct.data.ptr), ct.stride, batch, _vp(self.stream))
        return ct

    def encrypt_zero(self, parms_id=0, batch=1, seeds=None):
        return self._encrypt_zero(parms_id, True, batch, seeds)

    def encrypt_zero_symmetric(self, parms_id=0, batch=1, seeds=None, save_seed=False):
        return self._encrypt_zero(parms_id, False, batch, seeds, save_seed)

    def _encrypt(self, plain, asymmetric, seeds, save_seed=False):
        """Encryptor::encryptInternal  src/encryptor.cpp:166-253: encryption of zero, then the plaintext goes into c0."""
        ctx, scheme = self.context, self.context.scheme
        batch = plain.batch
        if scheme == CKKS:
            if not plain.is_ntt_form:
                raise TroyInvalidArgument("plain must be in NTT form")
            ct = self._encrypt_zero(plain.parms_id, asymmetric, batch, seeds, save_seed)
            for b in range(batch):  # addPolyCoeffmod on c0
                _L().call("troyb_add", _vp(ctx.h), ct.parms_id, _vp(ct.ptr(0, b)), _vp(plain.data.at(b * plain.coeff_count)),
                          _vp(ct.ptr(0, b)), 1, _vp(self.stream))
            ct.scale = plain.scale
            return ct
        if plain.is_ntt_form:
            raise TroyInvalidArgument("plain cannot be in NTT form")
        ct = self._encrypt_zero(0, asymmetric, batch, seeds, save_seed)
        # BFV: multiplyAddPlainWithScalingVariant; BGV: addPlainWithoutScalingVariant (correction factor 1)
        _L().call("troyb_add_plain", _vp(ctx.h), 0, _vp(ct.ptr()), ct.stride, _vp(plain.data.ptr), plain.coeff_count,
                  plain.stride, C.c_uint64(1), 0, batch, _vp(self.stream))
        return ct

    def encrypt(self, plain, seeds=None):
        return self._encrypt(plain, True, seeds)

    def encrypt_symmetric(self, plain, seeds=None, save_seed=False):
        return self._encrypt(plain, False, seeds, save_seed)


def _balance_correction_factors(f1, f2, t):
    """balanceCorrectionFactors, src/evaluator_cuda.cu:53-114: the pair (e1, e2) with e1 f1 = e2 f2 (mod t) of least
    |e1| + |e2| among the remainders of the extended Euclid run on (t, f2 / f1). Returns (new factor, e1, e2)."""
    half = t // 2

    def sum_abs(x, y):
        return abs(x - t if x > half else x) + abs(y - t if y > half else y)

    ratio = pow(f1, -1, t) * f2 % t
    e1, e2, best = ratio, 1, sum_abs(ratio, 1)
    prev_a, prev_b, a, b = t, 0, ratio, 1
    while a != 0:
        q = prev_a // a
        prev_a, a = a, prev_a % a
        prev_b, b = b, prev_b - b * q
        a_mod, b_mod = a % t, (abs(b) % t if b >= 0 else (t - abs(b) % t) % t)
        if a_mod != 0 and math.gcd(a_mod, t) == 1:
            s = sum_abs(a_mod, b_mod)
            if s < best:
                best, e1, e2 = s, a_mod, b_mod
    return e1 * f1 % t, e1, e2


def _naf(value):
    """src/utils/numth.h:16-36."""
    res = []
    sign = value < 0
    value = abs(value)
    i = 0
    while value:
        zi = (2 - (value & 3)) if (value & 1) else 0
        value = (value - zi) >> 1
        if zi:
            res.append((-zi if sign else zi) * (1 << i))
        i += 1
    return res


class EvaluatorCuda:
    """src/evaluator_cuda.cuh:17-440 — the methods on the RNS polynomial hot path. ``stream`` is a raw cudaStream_t
    handle (e.g. ``torch.cuda.current_stream().cuda_stream``); ``None`` is the default stream like the reference."""

    def __init__(self, context, stream=None):
        self.context = context
        self.stream = stream

    # -- helpers
    def _h(self):
        return _vp(self.context.h)

    def _s(self):
        return _vp(self.stream)

    def _own(self, *objs):
        """Stream hand-over: every buffer this call reads or writes joins the evaluator's stream. A buffer last touched on a
        different stream is waited for first (an event-free, conservative hand-over: streams that stay with one evaluator
        never synchronise), and from here on it is copied, grown and released in the order of this stream."""
        for o in objs:
            if isinstance(o, CiphertextCuda):
                o.seeds = None
            d = getattr(o, "data", None)
            if isinstance(d, DeviceArray) and d.stream is not self.stream:
                if d.stream != self.stream:
                    _L().call("troyb_stream_synchronize", _vp(d.stream))
                d.stream = self.stream

    def _check_scale(self, scale, level):
        """isScaleWithinBounds, src/evaluator_cuda.cu:32-51."""
        if self.context.scheme == CKKS:
            bound = self.context.total_coeff_modulus_bit_count(level)
        else:
            bound = int(self.context.parms.plain_modulus).bit_length()
        return not (scale <= 0 or int(math.log2(scale)) >= bound)

    def _same_parms(self, a, b):
        if a.parms_id != b.parms_id or a.batch != b.batch:
            raise TroyInvalidArgument("encrypted1 and encrypted2 parameter mismatch")
        if a.is_ntt_form != b.is_ntt_form:
            raise TroyInvalidArgument("NTT form mismatch")

    # -- add / sub / negate (src/evaluator_cuda.cu:143-260)
    @staticmethod
    def _dense(ct):
        """the batch is one contiguous run of polys (no spare capacity between ciphertexts): one launch covers it"""
        return ct.capacity == ct.size

    def negate_inplace(self, ct):
        self._own(ct)
        if self._dense(ct):
            _L().call("troyb_negate", self._h(), ct.parms_id, _vp(ct.ptr()), _vp(ct.ptr()), ct.batch * ct.size, self._s())
            return
        for b in range(ct.batch):
            _L().call("troyb_negate", self._h(), ct.parms_id, _vp(ct.ptr(0, b)), _vp(ct.ptr(0, b)), ct.size, self._s())

    def _addsub(self, name, a, b):
        self._own(a, b)
        self._same_parms(a, b)
        if self.context.scheme == BGV and a.correction_factor != b.correction_factor:
            # balance the correction factors with scalar multiplications first (src/evaluator_cuda.cu:171-190)
            f, e1, e2 = _balance_correction_factors(a.correction_factor, b.correction_factor, self.context.parms.plain_modulus)
            b = b.clone(self.stream)
            for ct, e in ((a, e1), (b, e2)):
                for i in range(ct.batch):
                    _L().call("troyb_multiply_scalar", self._h(), ct.parms_id, _vp(ct.ptr(0, i)), e, _vp(ct.ptr(0, i)), ct.size,
                              self._s())
                ct.correction_factor = f
        if self.context.scheme == CKKS and not math.isclose(a.scale, b.scale, rel_tol=1e-15):
            raise TroyInvalidArgument("scale mismatch")
        mn, mx = min(a.size, b.size), max(a.size, b.size)
        a_old = a.size
        if a.size == b.size and self._dense(a) and self._dense(b):
            _L().call(name, self._h(), a.parms_id, _vp(a.ptr()), _vp(b.ptr()), _vp(a.ptr()), a.batch * a.size, self._s())
            return
        a.resize(mx, self.stream)
        for i in range(a.batch):
            _L().call(name, self._h(), a.parms_id, _vp(a.ptr(0, i)), _vp(b.ptr(0, i)), _vp(a.ptr(0, i)), mn, self._s())
            if a_old < b.size:
                words = (b.size - a_old) * a.poly_words
                if name == "troyb_add":
                    _L().call("troyb_copy_on_device", _vp(a.ptr(a_old, i)), _vp(b.ptr(a_old, i)), words * 8, self._s())
                else:
                    _L().call("troyb_negate", self._h(), a.parms_id, _vp(b.ptr(a_old, i)), _vp(a.ptr(a_old, i)),
                              b.size - a_old, self._s())

    def add_inplace(self, a, b):
        self._addsub("troyb_add", a, b)

    def sub_inplace(self, a, b):
        self._addsub("troyb_sub", a, b)

    # -- multiply / square (src/evaluator_cuda.cu:262-701)
    def multiply_inplace(self, a, b):
        self._own(a, b)
        self._same_parms(a, b)
        scheme = self.context.scheme
        if scheme == CKKS and not (a.is_ntt_form and b.is_ntt_form):
            raise TroyInvalidArgument("encrypted1 or encrypted2 must be in NTT form")
        if scheme != CKKS and (a.is_ntt_form or b.is_ntt_form):
            raise TroyInvalidArgument("encrypted1 or encrypted2 cannot be in NTT form")
        out_size = a.size + b.size - 1
        new_scale = a.scale * b.scale
        if scheme == CKKS and not self._check_scale(new_scale, a.parms_id):
            raise TroyInvalidArgument("scale out of bounds")
        a_size = a.size
        b_ptr, b_stride = b.ptr(), b.stride  # taken before a possible reallocation when b is a
        a.reserve(out_size, self.stream)
        if b is a:
            b_ptr, b_stride = a.ptr(), a.stride
        _L().call("troyb_multiply", self._h(), a.parms_id, _vp(a.ptr()), a_size, a.stride, _vp(b_ptr), b.size, b_stride,
                  _vp(a.ptr()), a.stride, a.batch, self._s())
        a.size = out_size
        if scheme == CKKS:
            a.scale = new_scale
        if scheme == BGV:
            a.correction_factor = (a.correction_factor * b.correction_factor) % self.context.parms.plain_modulus

    def multiply(self, a, b, destination=None):
        """multiply(encrypted1, encrypted2, destination) (src/evaluator_cuda.cuh:85-97). With a ``destination`` of the same
        batch and level and capacity for the product, the product is written there directly (the reference copies first)."""
        self._own(a, b, destination)
        out_size = a.size + b.size - 1
        if destination is None or destination is a or destination is b or destination.batch != a.batch \
                or destination.limbs != a.limbs or destination.capacity < out_size:
            d = a.clone(self.stream)
            self.multiply_inplace(d, b)
            if destination is not None:
                self._assign(destination, d)
                return destination
            return d
        self._same_parms(a, b)
        scheme = self.context.scheme
        if scheme == CKKS and not (a.is_ntt_form and b.is_ntt_form):
            raise TroyInvalidArgument("encrypted1 or encrypted2 must be in NTT form")
        if scheme != CKKS and (a.is_ntt_form or b.is_ntt_form):
            raise TroyInvalidArgument("encrypted1 or encrypted2 cannot be in NTT form")
        new_scale = a.scale * b.scale
        if scheme == CKKS and not self._check_scale(new_scale, a.parms_id):
            raise TroyInvalidArgument("scale out of bounds")
        d = destination
        _L().call("troyb_multiply", self._h(), a.parms_id, _vp(a.ptr()), a.size, a.stride, _vp(b.ptr()), b.size, b.stride,
                  _vp(d.ptr()), d.stride, a.batch, self._s())
        d.parms_id, d.size, d.is_ntt_form = a.parms_id, out_size, a.is_ntt_form
        d.scale = new_scale if scheme == CKKS else a.scale
        d.correction_factor = a.correction_factor
        if scheme == BGV:
            d.correction_factor = (a.correction_factor * b.correction_factor) % self.context.parms.plain_modulus
        return d

    def square_inplace(self, a):
        self.multiply_inplace(a, a)

    # -- key switching (src/evaluator_cuda.cu:703-747, 1163-1378)
    def _check_keyswitch(self, ct, keys):
        scheme = self.context.scheme
        if not self.context.using_keyswitching():
            raise TroyLogicError("keyswitching is not supported by the context")
        if keys.parms_id != self.context.key_parms_id():
            raise TroyInvalidArgument("parameter mismatch")
        if scheme == BFV and ct.is_ntt_form:
            raise TroyInvalidArgument("BFV encrypted cannot be in NTT form")
        if scheme == CKKS and not ct.is_ntt_form:
            raise TroyInvalidArgument("CKKS encrypted must be in NTT form")
        if scheme == BGV and ct.is_ntt_form:
            raise TroyInvalidArgument("BGV encrypted cannot be in NTT form")

    def relinearize_inplace(self, ct, relin_keys):
        """relinearizeInternal(ct, keys, 2); only 3 -> 2 is on the ported path (§A.9)."""
        self._own(ct)
        if ct.size == 2:
            return
        if ct.size != 3:
            raise TroyLogicError("relinearization of size > 3 is outside the ported path (SURVEY.md §A.9)")
        self._check_keyswitch(ct, relin_keys)
        if not relin_keys.has_key(2):
            raise TroyInvalidArgument("not enough relinearization keys")
        _L().call("troyb_relinearize", self._h(), ct.parms_id, _vp(ct.ptr()), ct.stride, _vp(relin_keys.key(2).ptr), ct.batch,
                  self._s())
        ct.size = 2

    def switch_key_inplace(self, ct, target, keys, index):
        """switchKeyInplace(ct, target, keys, index): target is a DeviceArray/[batch][L][N] buffer."""
        self._own(ct, target)
        self._check_keyswitch(ct, keys)
        _L().call("troyb_switch_key", self._h(), ct.parms_id, _vp(ct.ptr()), ct.stride, _vp(target.ptr), ct.poly_words,
                  _vp(keys.data(index).ptr), ct.batch, self._s())

    # -- Galois / rotations (src/evaluator_cuda.cu:2024-2176)
    def apply_galois_inplace(self, ct, galois_elt, galois_keys):
        self._own(ct)
        if galois_keys.parms_id != self.context.key_parms_id():
            raise TroyInvalidArgument("galois_keys is not valid for encryption parameters")
        if not (galois_elt & 1) or galois_elt >= 2 * self.context.n:
            raise TroyInvalidArgument("Galois element is not valid")
        if not galois_keys.has_key(galois_elt):
            raise TroyInvalidArgument("Galois key not present")
        if ct.size > 2:
            raise TroyInvalidArgument("encrypted size must be 2")
        self._check_keyswitch(ct, galois_keys)
        _L().call("troyb_apply_galois", self._h(), ct.parms_id, _vp(ct.ptr()), ct.stride, galois_elt,
                  _vp(galois_keys.key(galois_elt).ptr), ct.batch, self._s())

    def apply_galois(self, ct, galois_elt, galois_keys, destination):
        """applyGalois(encrypted, elt, keys, destination) (src/evaluator_cuda.cuh:275): ``destination`` (same batch, level and
        size-2 capacity) receives the result; ``ct`` is only read."""
        if destination is ct:
            return self.apply_galois_inplace(ct, galois_elt, galois_keys)
        self._own(ct, destination)
        if galois_keys.parms_id != self.context.key_parms_id():
            raise TroyInvalidArgument("galois_keys is not valid for encryption parameters")
        if not (galois_elt & 1) or galois_elt >= 2 * self.context.n:
            raise TroyInvalidArgument("Galois element is not valid")
        if not galois_keys.has_key(galois_elt):
            raise TroyInvalidArgument("Galois key not present")
        if ct.size > 2:
            raise TroyInvalidArgument("encrypted size must be 2")
        if destination.batch != ct.batch or destination.limbs != ct.limbs or destination.capacity < 2:
            raise TroyInvalidArgument("destination is not valid for encryption parameters")
        self._check_keyswitch(ct, galois_keys)
        d = destination
        _L().call("troyb_apply_galois_to", self._h(), ct.parms_id, _vp(ct.ptr()), ct.stride, _vp(d.ptr()), d.stride, galois_elt,
                  _vp(galois_keys.key(galois_elt).ptr), ct.batch, self._s())
        d.parms_id, d.size, d.is_ntt_form, d.scale, d.correction_factor = ct.parms_id, 2, ct.is_ntt_form, ct.scale, ct.correction_factor
        return d

    def rotate_vector(self, ct, steps, galois_keys, destination):
        """rotateVector(encrypted, steps, keys, destination): one key switch when the key for ``steps`` exists."""
        if self.context.scheme != CKKS:
            raise TroyLogicError("unsupported scheme")
        return self._rotate_to(ct, steps, galois_keys, destination)

    def rotate_rows(self, ct, steps, galois_keys, destination):
        if self.context.scheme not in (BFV, BGV):
            raise TroyLogicError("unsupported scheme")
        return self._rotate_to(ct, steps, galois_keys, destination)

    def _rotate_to(self, ct, steps, galois_keys, destination):
        elt = self.context.galois_elt_from_step(steps) if steps != 0 else None
        if elt is not None and galois_keys.has_key(elt):
            return self.apply_galois(ct, elt, galois_keys, destination)
        self._own(ct, destination)
        self._assign(destination, ct.clone(self.stream))
        self._rotate_internal(destination, steps, galois_keys)
        return destination

    def _rotate_internal(self, ct, steps, galois_keys):
        if galois_keys.parms_id != self.context.key_parms_id():
            raise TroyInvalidArgument("galois_keys is not valid for encryption parameters")
        if steps == 0:
            return
        elt = self.context.galois_elt_from_step(steps)
        if galois_keys.has_key(elt):
            self.apply_galois_inplace(ct, elt, galois_keys)
            return
        naf_steps = _naf(steps)
        if len(naf_steps) == 1:
            raise TroyInvalidArgument("Galois key not present")
        for s in naf_steps:
            if abs(s) != (self.context.n >> 1):
                self._rotate_internal(ct, s, galois_keys)

    def rotate_rows_inplace(self, ct, steps, galois_keys):
        if self.context.scheme not in (BFV, BGV):
            raise TroyLogicError("unsupported scheme")
        self._rotate_internal(ct, steps, galois_keys)

    def rotate_columns_inplace(self, ct, galois_keys):
        if self.context.scheme not in (BFV, BGV):
            raise TroyLogicError("unsupported scheme")
        self.apply_galois_inplace(ct, self.context.galois_elt_from_step(0), galois_keys)

    def rotate_vector_inplace(self, ct, steps, galois_keys):
        if self.context.scheme != CKKS:
            raise TroyLogicError("unsupported scheme")
        self._rotate_internal(ct, steps, galois_keys)

    def complex_conjugate_inplace(self, ct, galois_keys):
        if self.context.scheme != CKKS:
            raise TroyLogicError("unsupported scheme")
        self.apply_galois_inplace(ct, self.context.galois_elt_from_step(0), galois_keys)

    # -- rescale / mod switch (src/evaluator_cuda.cu:749-1010, 1380)
    def _scale_to_next(self, ct):
        self._own(ct)
        scheme = self.context.scheme
        if ct.parms_id + 1 >= self.context.num_levels:
            raise TroyInvalidArgument("end of modulus switching chain reached")
        if scheme == CKKS and not ct.is_ntt_form:
            raise TroyInvalidArgument("CKKS encrypted must be in NTT form")
        if scheme != CKKS and ct.is_ntt_form:
            raise TroyInvalidArgument("BFV encrypted cannot be in NTT form" if scheme == BFV
                                      else "BGV encrypted cannot be in NTT form")
        q_last = self.context.level_moduli(ct.parms_id)[-1]
        out = CiphertextCuda(self.context, ct.parms_id + 1, ct.size, ct.is_ntt_form, ct.scale, ct.correction_factor, ct.batch,
                             stream=self.stream)
        _L().call("troyb_mod_switch_scale_to_next", self._h(), ct.parms_id, _vp(ct.ptr()), ct.stride, _vp(out.ptr()),
                  out.stride, ct.size, ct.batch, self._s())
        if scheme == CKKS:
            out.scale = ct.scale / float(q_last)
        elif scheme == BGV:
            t = self.context.parms.plain_modulus
            out.correction_factor = (ct.correction_factor * pow(q_last % t, -1, t)) % t
        return out

    def _assign(self, dst, src):
        if isinstance(dst.data, DeviceArray) and dst.data is not src.data:
            dst.data.stream = self.stream  # the words being replaced are released behind this stream's kernels
        dst.parms_id, dst.size, dst.is_ntt_form, dst.scale = src.parms_id, src.size, src.is_ntt_form, src.scale
        dst.correction_factor, dst.batch, dst.limbs, dst.capacity, dst.data = \
            src.correction_factor, src.batch, src.limbs, src.capacity, src.data

    def rescale_to_next(self, ct):
        if self.context.scheme != CKKS:
            raise TroyInvalidArgument("unsupported operation for scheme type")
        return self._scale_to_next(ct)

    def rescale_to_next_inplace(self, ct):
        self._assign(ct, self.rescale_to_next(ct))

    def mod_switch_to_next(self, ct):
        """modSwitchToNext: BFV/BGV scale-and-round; CKKS just drops the last limb (modSwitchDropToNext,
        src/evaluator_cuda.cu:826-880)."""
        if self.context.scheme != CKKS:
            return self._scale_to_next(ct)
        if ct.parms_id + 1 >= self.context.num_levels:
            raise TroyInvalidArgument("end of modulus switching chain reached")
        self._own(ct)
        out = CiphertextCuda(self.context, ct.parms_id + 1, ct.size, ct.is_ntt_form, ct.scale, ct.correction_factor, ct.batch,
                             stream=self.stream)
        n = self.context.n
        for b in range(ct.batch):
            for p in range(ct.size):
                _L().call("troyb_copy_on_device", _vp(out.ptr(p, b)), _vp(ct.ptr(p, b)), out.limbs * n * 8, self._s())
        return out

    def mod_switch_to_next_inplace(self, ct):
        self._assign(ct, self.mod_switch_to_next(ct))

    # -- ciphertext (x) plaintext (src/evaluator_cuda.cu:1555-1935)
    def _check_plain_batch(self, ct, plain):
        if plain.batch not in (1, ct.batch):
            raise TroyInvalidArgument("plain batch does not match the ciphertext batch")

    def _addsub_plain(self, ct, plain, subtract):
        self._own(ct, plain)
        scheme = self.context.scheme
        if scheme == BFV and ct.is_ntt_form:
            raise TroyInvalidArgument("BFV encrypted cannot be in NTT form")
        if scheme == CKKS and not ct.is_ntt_form:
            raise TroyInvalidArgument("CKKS encrypted must be in NTT form")
        if scheme == BGV and ct.is_ntt_form:
            raise TroyInvalidArgument("BGV encrypted cannot be in NTT form")
        if plain.is_ntt_form != ct.is_ntt_form:
            raise TroyInvalidArgument("NTT form mismatch")
        if ct.is_ntt_form and ct.parms_id != plain.parms_id:
            raise TroyInvalidArgument("encrypted and plain parameter mismatch")
        if scheme == CKKS and not math.isclose(ct.scale, plain.scale, rel_tol=1e-15):
            raise TroyInvalidArgument("scale mismatch")
        self._check_plain_batch(ct, plain)
        if scheme == CKKS:
            # kAddPolyCoeffmod / kSubPolyCoeffmod on poly 0 (src/evaluator_cuda.cu:1607-1612)
            name = "troyb_sub" if subtract else "troyb_add"
            for b in range(ct.batch):
                pp = plain.data.at((b if plain.batch > 1 else 0) * plain.coeff_count)
                _L().call(name, self._h(), ct.parms_id, _vp(ct.ptr(0, b)), _vp(pp), _vp(ct.ptr(0, b)), 1, self._s())
        else:
            _L().call("troyb_add_plain", self._h(), ct.parms_id, _vp(ct.ptr()), ct.stride, _vp(plain.data.ptr),
                      plain.coeff_count, plain.stride, C.c_uint64(ct.correction_factor), int(subtract), ct.batch, self._s())

    def add_plain_inplace(self, ct, plain):
        self._addsub_plain(ct, plain, False)

    def sub_plain_inplace(self, ct, plain):
        self._addsub_plain(ct, plain, True)

    def multiply_plain_inplace(self, ct, plain):
        """multiplyPlainInplace → multiplyPlainNtt / multiplyPlainNormal (src/evaluator_cuda.cu:1704-1862)."""
        self._own(ct, plain)
        if ct.is_ntt_form != plain.is_ntt_form:
            raise TroyInvalidArgument("NTT form mismatch")
        self._check_plain_batch(ct, plain)
        new_scale = ct.scale * plain.scale
        if ct.is_ntt_form:
            if ct.parms_id != plain.parms_id:
                raise TroyInvalidArgument("encrypted_ntt and plain_ntt parameter mismatch")
            _L().call("troyb_multiply_plain_ntt", self._h(), ct.parms_id, _vp(ct.ptr()), ct.size, ct.stride,
                      _vp(plain.data.ptr), plain.stride, _vp(ct.ptr()), ct.stride, ct.batch, 0, self._s())
            ct.scale = new_scale
            if not self._check_scale(ct.scale, ct.parms_id):
                raise TroyInvalidArgument("scale out of bounds")
        else:
            if plain.batch != 1:
                raise TroyInvalidArgument("multiply_plain in coefficient form takes one plaintext per call")
            _L().call("troyb_multiply_plain_normal", self._h(), ct.parms_id, _vp(ct.ptr()), ct.size, ct.stride,
                      _vp(plain.data.ptr), plain.coeff_count, ct.batch, self._s())
            if self.context.scheme == CKKS:
                ct.scale = new_scale
                if not self._check_scale(ct.scale, ct.parms_id):
                    raise TroyInvalidArgument("scale out of bounds")

    def multiply_plain(self, ct, plain):
        self._own(ct)
        d = ct.clone(self.stream)
        self.multiply_plain_inplace(d, plain)
        return d

    def multiply_plain_accumulate(self, ct, plain, destination):
        """destination += ct * plain (NTT form) in one pass: the multiplyPlain + addInplace pair of
        MatmulHelper::matmul (app/LinearHelperCKKS.cuh:225-249). The scales must agree like in addInplace."""
        self._own(ct, plain, destination)
        if not (ct.is_ntt_form and plain.is_ntt_form and destination.is_ntt_form):
            raise TroyInvalidArgument("NTT form mismatch")
        if ct.parms_id != plain.parms_id or ct.parms_id != destination.parms_id or ct.size != destination.size \
                or ct.batch != destination.batch:
            raise TroyInvalidArgument("encrypted_ntt and plain_ntt parameter mismatch")
        self._check_plain_batch(ct, plain)
        if self.context.scheme == CKKS and not math.isclose(ct.scale * plain.scale, destination.scale, rel_tol=1e-15):
            raise TroyInvalidArgument("scale mismatch")
        _L().call("troyb_multiply_plain_ntt", self._h(), ct.parms_id, _vp(ct.ptr()), ct.size, ct.stride, _vp(plain.data.ptr),
                  plain.stride, _vp(destination.ptr()), destination.stride, ct.batch, 1, self._s())

    def transform_plain_to_ntt_inplace(self, plain, parms_id):
        """transformToNttInplace(PlaintextCuda&, ParmsID) (src/evaluator_cuda.cu:1865-1935): lift + forward NTT."""
        if plain.is_ntt_form:
            raise TroyInvalidArgument("plain is already in NTT form")
        if not (0 <= parms_id < self.context.num_levels):
            raise TroyInvalidArgument("parms_id is not valid for the current context")
        self._own(plain)
        words = self.context.level_limbs(parms_id) * self.context.n
        out = self.context.alloc(plain.batch * words, self.stream)
        _L().call("troyb_plain_lift", self._h(), parms_id, _vp(plain.data.ptr), plain.coeff_count, plain.coeff_count,
                  _vp(out.ptr), plain.batch, 1, self._s())
        plain.data, plain.coeff_count, plain.parms_id = out, words, parms_id

    # -- NTT transforms (src/evaluator_cuda.cu:1950-2022)
    def transform_to_ntt_inplace(self, ct):
        if ct.is_ntt_form:
            raise TroyInvalidArgument("encrypted is already in NTT form")
        self._own(ct)
        if self._dense(ct):
            _L().call("troyb_ntt_forward", self._h(), ct.parms_id, _vp(ct.ptr()), ct.batch * ct.size, self._s())
        else:
            for b in range(ct.batch):
                _L().call("troyb_ntt_forward", self._h(), ct.parms_id, _vp(ct.ptr(0, b)), ct.size, self._s())
        ct.is_ntt_form = True

    def transform_from_ntt_inplace(self, ct):
        if not ct.is_ntt_form:
            raise TroyInvalidArgument("encrypted_ntt is not in NTT form")
        self._own(ct)
        if self._dense(ct):
            _L().call("troyb_ntt_inverse", self._h(), ct.parms_id, _vp(ct.ptr()), ct.batch * ct.size, self._s())
        else:
            for b in range(ct.batch):
                _L().call("troyb_ntt_inverse", self._h(), ct.parms_id, _vp(ct.ptr(0, b)), ct.size, self._s())
        ct.is_ntt_form = False

    def negacyclic_shift(self, ct, shift):
        """src/evaluator_cuda.cu:2185-2205: ct * X^shift modulo X^N + 1 (coefficient form), a new ciphertext; shift in [0, 2N)."""
        if ct.is_ntt_form:
            raise TroyInvalidArgument("encrypted is in NTT form")
        self._own(ct)
        out = CiphertextCuda(self.context, ct.parms_id, ct.size, False, ct.scale, ct.correction_factor, ct.batch, ct.size,
                             stream=self.stream)
        for b in range(ct.batch):
            _L().call("troyb_negacyclic_shift", self._h(), ct.parms_id, _vp(ct.ptr(0, b)), _vp(out.ptr(0, b)), ct.size,
                      C.c_uint64(int(shift)), self._s())
        return out

    def divide_by_poly_modulus_degree_inplace(self, ct, mul=1):
        """src/evaluator_cuda.cu:2262-2275: every word times N^-1 (then times ``mul``) modulo its limb's prime."""
        self._own(ct)
        for b in range(ct.batch):
            _L().call("troyb_multiply_inv_degree", self._h(), ct.parms_id, _vp(ct.ptr(0, b)), _vp(ct.ptr(0, b)), ct.size, self._s())
            if mul > 1:
                _L().call("troyb_multiply_scalar", self._h(), ct.parms_id, _vp(ct.ptr(0, b)), C.c_uint64(int(mul)), _vp(ct.ptr(0, b)),
                          ct.size, self._s())

    def synchronize(self):
        _L().call("troyb_stream_synchronize", self._s())

    # camelCase aliases matching the reference's spelling
    negacyclicShift, divideByPolyModulusDegreeInplace = negacyclic_shift, divide_by_poly_modulus_degree_inplace
    negateInplace, addInplace, subInplace = negate_inplace, add_inplace, sub_inplace
    multiplyInplace, squareInplace, relinearizeInplace = multiply_inplace, square_inplace, relinearize_inplace
    applyGaloisInplace, rotateRowsInplace, rotateColumnsInplace = apply_galois_inplace, rotate_rows_inplace, rotate_columns_inplace
    rotateVectorInplace, complexConjugateInplace = rotate_vector_inplace, complex_conjugate_inplace
    rescaleToNext, rescaleToNextInplace = rescale_to_next, rescale_to_next_inplace
    modSwitchToNext, modSwitchToNextInplace = mod_switch_to_next, mod_switch_to_next_inplace
    transformToNttInplace, transformFromNttInplace = transform_to_ntt_inplace, transform_from_ntt_inplace
    addPlainInplace, subPlainInplace, multiplyPlainInplace = add_plain_inplace, sub_plain_inplace, multiply_plain_inplace
    multiplyPlain = multiply_plain
    applyGalois, rotateVector, rotateRows = apply_galois, rotate_vector, rotate_rows
