"""`-m gpu`: the C++ drop-in layer (include/troyn/troy_cuda.cuh over libtroy_b200.so).

tests/cpp/_build/troyn_dropin_tests is a gtest binary built by tests/cpp/build_dropin_tests.py from (1) the reference's own
CUDA tests for the hot path (test/evaluator_cuda.cu, compiled where they lie, only the `troy_cuda.cuh` they include is
ours) and (2) tests/cpp/dropin_parity.cpp, which compares every output word with the reference's CPU evaluator."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "troyn_dropin_tests")


def _run(filt):
    if not os.path.exists(BIN):
        pytest.skip("tests/cpp/_build/troyn_dropin_tests not built (run __graft_entry__.build() where /root/reference exists)")
    r = subprocess.run([BIN, f"--gtest_filter={filt}"], capture_output=True, text=True, timeout=1500)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-40:])
    assert r.returncode == 0, tail
    assert "[  PASSED  ]" in r.stdout, tail
    return r.stdout


def test_word_for_word_parity_through_the_cpp_layer():
    out = _run("DropInParity.*")
    assert "6 tests" in out  # 3 evaluator flows + 3 encryption flows (device encryptor, seeded transport form)


def test_reference_cuda_tests_pass_against_the_drop_in():
    """All 52 tests of the reference's test/evaluator_cuda.cu, unchanged (the last one, BFVAssembleLWETest, covers LWE extraction,
    assembly, the field trace and packLWECiphertexts)."""
    out = _run("EvaluatorCudaTest.*")
    assert "[  FAILED  ]" not in out
    assert "52 tests" in out


def test_reference_encryptor_and_encoder_test_files_pass_against_the_drop_in():
    """The reference's test/encryptor_cuda.cu (6 tests: encryption and decryption now run on the device) and test/ckks_cuda.cu
    (2 tests), whole files, unchanged but for the include of the CUDA API."""
    out = _run("EncryptorCudaTest.*:CKKSEncoderCudaTest.*")
    assert "[  FAILED  ]" not in out
    assert "8 tests" in out


def test_reference_app_header_matmul_runs_on_the_drop_in():
    """app/LinearHelperCKKS.cuh (MatmulHelper) compiled unchanged against the drop-in header: encode, encrypt, matmul,
    serialise by terms, decrypt on the device; the decoded product equals x @ W."""
    out = _run("DropInLinear.*")
    assert "2 tests" in out
