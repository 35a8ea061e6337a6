"""`-m "not gpu"`: the C-ABI library loads without a GPU and exports every symbol include/troy_b200.h declares;
host-side logic that needs no device (argument checks, NAF, key indices)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "troy_b200.h")).read()
    return sorted(set(re.findall(r"\b(troyb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from troy_b200 import lib as tlib
    L = tlib.Library(tlib.LIB_PATH)  # the CUDA build; dlopen only, no compute calls
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L.cdll, n), n
    assert set(names) == set(tlib.EXPORTED_SYMBOLS)


def test_calls_before_initialize_fail_loudly():
    from troy_b200 import lib as tlib
    import ctypes as C
    L = tlib.Library(tlib.LIB_PATH)
    if L.cdll.troyb_initialized():
        pytest.skip("already initialised in this process")
    p = C.c_void_p()
    with pytest.raises(tlib.TroyInvalidArgument, match="KernelProvider not initialized"):
        L.call("troyb_malloc", C.byref(p), 64, None)


def test_missing_library_is_an_import_error(tmp_path):
    from troy_b200 import lib as tlib
    with pytest.raises(ImportError, match="no CPU fallback"):
        tlib.Library(str(tmp_path / "libtroy_b200.so"))


def test_host_helpers():
    from troy_b200.host import GaloisKeysCuda, RelinKeysCuda, _naf
    assert _naf(7) == [-1, 8] and _naf(-7) == [1, -8] and _naf(12) == [-4, 16] and _naf(1) == [1]
    assert GaloisKeysCuda.get_index(3) == 1 and GaloisKeysCuda.get_index(15) == 7
    assert RelinKeysCuda.get_index(2) == 0
    with pytest.raises(ValueError):
        RelinKeysCuda.get_index(1)


def test_pytroy_module_imports_and_names_follow_the_reference_binder():
    """binder/pytroy*.so (when built): importable without a GPU, and the classes / methods a script written for the reference's
    pytroy module uses on the path are there under the same names (binder/binder.cu)."""
    import glob
    import sys
    if not glob.glob(os.path.join(ROOT, "binder", "pytroy*.so")):
        pytest.skip("binder/pytroy*.so not built")
    sys.path.insert(0, os.path.join(ROOT, "binder"))
    import pytroy
    for cls, methods in {
        "EncryptionParameters": ["set_poly_modulus_degree", "set_coeff_modulus", "set_plain_modulus"],
        "CoeffModulus": ["create"], "PlainModulus": ["batching"], "SEALContext": ["first_parms_id", "key_parms_id"],
        "KeyGenerator": ["secret_key", "create_public_key", "create_relin_keys", "create_galois_keys", "create_automorphism_keys"],
        "LWECiphertext": ["copy"],
        "Encryptor": ["encrypt", "encrypt_symmetric", "encrypt_zero"], "Decryptor": ["decrypt"],
        "CKKSEncoder": ["encode", "decode", "slot_count"], "BatchEncoder": ["encode", "decode"],
        "Ciphertext": ["save", "load", "save_terms", "load_terms", "scale", "parms_id", "size"], "Plaintext": ["save", "load"],
        "RelinKeys": ["save", "load"], "GaloisKeys": ["save", "load"], "PublicKey": ["save", "load"], "SecretKey": ["save", "load"],
        "Evaluator": ["multiply_inplace", "relinearize_inplace", "rescale_to_next_inplace", "rotate_vector_inplace", "rotate_rows_inplace",
                      "mod_switch_to_next_inplace", "add_inplace", "sub_inplace", "negate_inplace", "square_inplace", "multiply_plain_inplace",
                      "add_plain_inplace", "transform_to_ntt_inplace", "transform_from_ntt_inplace", "apply_galois_inplace",
                      "complex_conjugate_inplace", "rotate_columns_inplace", "extract_lwe", "assemble_lwe", "field_trace_inplace",
                      "divide_by_poly_modulus_degree_inplace", "pack_lwe_ciphertexts", "negacyclic_shift", "negacyclic_shift_inplace"],
    }.items():
        c = getattr(pytroy, cls)
        for m in methods:
            assert hasattr(c, m), f"pytroy.{cls}.{m}"
    parms = pytroy.EncryptionParameters(pytroy.SchemeType.ckks)
    parms.set_poly_modulus_degree(8192)
    parms.set_coeff_modulus(pytroy.CoeffModulus.create(8192, [40, 40, 40]))
    assert parms.poly_modulus_degree() == 8192 and len(parms.coeff_modulus()) == 3
