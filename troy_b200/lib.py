"""ctypes binding of include/troy_b200.h. Loads the in-tree CUDA library; never falls back to anything else."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtroy_b200.so")

TROYB_OK = 0
E_INVALID_ARGUMENT, E_LOGIC, E_OUT_OF_RANGE, E_CUDA, E_NOT_INITIALIZED = 1, 2, 3, 4, 5


class TroyError(RuntimeError):
    """Base of the errors raised by the C ABI (std::runtime_error("CUDA error.") in the reference)."""


class TroyInvalidArgument(TroyError, ValueError):
    """std::invalid_argument in the reference."""


class TroyLogicError(TroyError):
    """std::logic_error in the reference."""


class TroyOutOfRange(TroyError, IndexError):
    """std::out_of_range in the reference."""


_ERRORS = {E_INVALID_ARGUMENT: TroyInvalidArgument, E_LOGIC: TroyLogicError, E_OUT_OF_RANGE: TroyOutOfRange,
           E_CUDA: TroyError, E_NOT_INITIALIZED: TroyInvalidArgument}

_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p
_sz = C.c_size_t

_SIGNATURES = {
    "troyb_last_error": (C.c_char_p, []),
    "troyb_initialize": (C.c_int, [C.c_int]),
    "troyb_initialized": (C.c_int, []),
    "troyb_set_device": (C.c_int, [C.c_int]),
    "troyb_context_device": (C.c_int, [_vp]),
    "troyb_launch_count": (C.c_uint64, []),
    "troyb_profile_enable": (C.c_int, [C.c_int]),
    "troyb_profile_report": (C.c_int, [C.c_char_p, _sz]),
    "troyb_coeff_modulus_create": (C.c_int, [C.c_uint64, C.POINTER(C.c_int), C.c_int, _u64p]),
    "troyb_malloc": (C.c_int, [C.POINTER(_vp), _sz, _vp]),
    "troyb_free": (C.c_int, [_vp, _vp]),
    "troyb_copy_to_device": (C.c_int, [_vp, _vp, _sz, _vp]),
    "troyb_copy_to_host": (C.c_int, [_vp, _vp, _sz, _vp]),
    "troyb_copy_on_device": (C.c_int, [_vp, _vp, _sz, _vp]),
    "troyb_memset_zero": (C.c_int, [_vp, _sz, _vp]),
    "troyb_stream_synchronize": (C.c_int, [_vp]),
    "troyb_context_create": (C.c_int, [C.c_int, C.c_uint64, _u64p, C.c_int, C.c_uint64, C.c_int, C.POINTER(_vp)]),
    "troyb_context_destroy": (None, [_vp]),
    "troyb_context_scheme": (C.c_int, [_vp]),
    "troyb_context_poly_degree": (C.c_uint64, [_vp]),
    "troyb_context_num_levels": (C.c_int, [_vp]),
    "troyb_context_level_limbs": (C.c_int, [_vp, C.c_int]),
    "troyb_context_level_moduli": (C.c_int, [_vp, C.c_int, _u64p]),
    "troyb_context_using_keyswitching": (C.c_int, [_vp]),
    "troyb_context_total_coeff_bits": (C.c_int, [_vp, C.c_int]),
    "troyb_context_plain_modulus": (C.c_uint64, [_vp]),
    "troyb_context_bsk": (C.c_int, [_vp, C.c_int, _u64p]),
    "troyb_context_root_powers": (C.c_int, [_vp, C.c_int, _u64p]),
    "troyb_ntt_forward": (C.c_int, [_vp, C.c_int, _vp, _sz, _vp]),
    "troyb_ntt_inverse": (C.c_int, [_vp, C.c_int, _vp, _sz, _vp]),
    "troyb_ntt_forward_key": (C.c_int, [_vp, _vp, _sz, _vp]),
    "troyb_ntt_inverse_key": (C.c_int, [_vp, _vp, _sz, _vp]),
    "troyb_add": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _sz, _vp]),
    "troyb_sub": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _sz, _vp]),
    "troyb_negate": (C.c_int, [_vp, C.c_int, _vp, _vp, _sz, _vp]),
    "troyb_multiply_scalar": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64, _vp, _sz, _vp]),
    "troyb_multiply_inv_degree": (C.c_int, [_vp, C.c_int, _vp, _vp, _sz, _vp]),
    "troyb_negacyclic_shift": (C.c_int, [_vp, C.c_int, _vp, _vp, _sz, C.c_uint64, _vp]),
    "troyb_multiply": (C.c_int, [_vp, C.c_int, _vp, _sz, _sz, _vp, _sz, _sz, _vp, _sz, _sz, _vp]),
    "troyb_switch_key": (C.c_int, [_vp, C.c_int, _vp, _sz, _vp, _sz, _vp, _sz, _vp]),
    "troyb_switch_key_partial": (C.c_int, [_vp, C.c_int, _vp, _sz, _vp, C.c_int, C.c_int, _vp, _sz, _vp]),
    "troyb_switch_key_finish": (C.c_int, [_vp, C.c_int, _vp, _sz, _vp, C.c_int, _sz, _vp]),
    "troyb_switch_key_partial_rows": (C.c_int, [_vp, C.c_int, _vp, _sz, _vp, C.c_int, C.c_int, _vp, _sz, _vp]),
    "troyb_switch_key_reduce_rows": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, _sz, _vp]),
    "troyb_switch_key_special_row": (C.c_int, [_vp, C.c_int, _vp, _vp, _sz, _vp]),
    "troyb_switch_key_moddown_rows": (C.c_int, [_vp, C.c_int, _vp, _sz, _vp, _vp, C.c_int, C.c_int, _sz, _vp]),
    "troyb_prng_fill": (C.c_int, [_vp, _sz, _vp, _sz, _sz, _sz, _vp]),
    "troyb_sample_poly_ternary": (C.c_int, [_vp, C.c_int, _vp, _sz, _sz, _vp, _sz, C.c_uint64, _vp, _sz, _vp, _sz, _vp]),
    "troyb_sample_poly_cbd": (C.c_int, [_vp, C.c_int, _vp, _sz, _sz, _vp, _sz, C.c_uint64, _vp, _sz, _sz, _vp]),
    "troyb_sample_poly_uniform": (C.c_int, [_vp, C.c_int, _vp, _sz, _sz, _vp, _sz, _vp, _sz, _sz, _vp]),
    "troyb_encrypt_zero_asymmetric": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _sz, _sz, _vp]),
    "troyb_encrypt_zero_symmetric": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _sz, _sz, _vp]),
    "troyb_encrypt_zero_symmetric_seeded": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    "troyb_expand_seeded_poly": (C.c_int, [_vp, C.c_int, _vp, _vp, _sz, _sz, _vp]),
    "troyb_relinearize": (C.c_int, [_vp, C.c_int, _vp, _sz, _vp, _sz, _vp]),
    "troyb_galois_permute": (C.c_int, [_vp, C.c_int, _vp, _vp, _sz, C.c_uint32, C.c_int, _vp]),
    "troyb_galois_elt_from_step": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_uint32)]),
    "troyb_apply_galois": (C.c_int, [_vp, C.c_int, _vp, _sz, C.c_uint32, _vp, _sz, _vp]),
    "troyb_apply_galois_to": (C.c_int, [_vp, C.c_int, _vp, _sz, _vp, _sz, C.c_uint32, _vp, _sz, _vp]),
    "troyb_mod_switch_scale_to_next": (C.c_int, [_vp, C.c_int, _vp, _sz, _vp, _sz, _sz, _sz, _vp]),
    "troyb_multiply_plain_ntt": (C.c_int, [_vp, C.c_int, _vp, _sz, _sz, _vp, _sz, _vp, _sz, _sz, C.c_int, _vp]),
    "troyb_matmul_plain_ntt": (C.c_int, [_vp, C.c_int, _vp, _sz, _sz, _vp, _vp, _sz, _sz, _sz, _sz, _vp]),
    "troyb_plain_lift": (C.c_int, [_vp, C.c_int, _vp, _sz, _sz, _vp, _sz, C.c_int, _vp]),
    "troyb_multiply_plain_normal": (C.c_int, [_vp, C.c_int, _vp, _sz, _sz, _vp, _sz, _sz, _vp]),
    "troyb_secret_key_powers": (C.c_int, [_vp, _vp, _sz, _vp, _vp]),
    "troyb_decrypt_dot_product": (C.c_int, [_vp, C.c_int, _vp, _sz, _sz, C.c_int, _vp, _vp, _sz, _vp]),
    "troyb_decrypt_scale_and_round": (C.c_int, [_vp, C.c_int, _vp, _vp, _sz, _vp]),
    "troyb_decrypt_mod_t": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_uint64, _sz, _vp]),
    "troyb_add_plain": (C.c_int, [_vp, C.c_int, _vp, _sz, _vp, _sz, _sz, C.c_uint64, C.c_int, _sz, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class Library:
    """A loaded C-ABI library with checked calls."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). troy_b200 has no CPU fallback.")
        self.path = path
        self.cdll = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            f = getattr(self.cdll, name)  # AttributeError if the symbol is not exported
            f.restype, f.argtypes = res, args

    def last_error(self):
        return self.cdll.troyb_last_error().decode()

    def call(self, name, *args):
        rc = getattr(self.cdll, name)(*args)
        if rc != TROYB_OK:
            raise _ERRORS.get(rc, TroyError)(self.last_error())
        return rc


_library = None


def library():
    """The product library (CUDA): always troy_b200/libtroy_b200.so. Raises ImportError when it has not been built."""
    global _library
    if _library is None:
        _library = Library(LIB_PATH)
    return _library
