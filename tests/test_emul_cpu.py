"""`-m "not gpu"`: the library's own kernel sources, compiled for sequential CPU execution (tests/emul), against the
reference CPU library. This checks host orchestration, table generation and kernel index arithmetic without a GPU;
it is a debugging aid, not the product path (the GPU parity tests in test_parity_gpu.py are the parity tests proper)."""
import numpy as np
import pytest

from common import run_decrypt_flow, run_evaluator_flow, run_matmul_flow, run_plain_flow


@pytest.mark.parametrize("scheme", [1, 2, 3])
@pytest.mark.parametrize("n,bits", [(64, [40, 30, 30, 40]), (4096, [60, 40, 40, 60])])
def test_evaluator_flow_emulated(ref, emul_lib, scheme, n, bits):
    run_evaluator_flow(ref, emul_lib, scheme, n, bits, np.random.default_rng(100 + scheme))


@pytest.mark.parametrize("scheme,n,bits", [(2, 16384, [60, 40, 60]), (1, 32768, [50, 40, 60]), (3, 65536, [50, 40, 60])])
def test_cluster_transforms_emulated(ref, emul_lib, scheme, n, bits):
    """N >= 16384: a limb is a cluster of 2 / 4 / 8 tiles exchanging data through distributed shared memory."""
    run_evaluator_flow(ref, emul_lib, scheme, n, bits, np.random.default_rng(200 + scheme))


def test_batch_emulated(ref, emul_lib):
    run_evaluator_flow(ref, emul_lib, ref.CKKS, 256, [40, 30, 30, 30, 40], np.random.default_rng(5), batch=3)


def test_many_digits_emulated(ref, emul_lib):
    """More than 8 digits: the key inner product runs over digit groups that are added up in the product buffer."""
    run_evaluator_flow(ref, emul_lib, ref.CKKS, 1024, [40] * 10 + [50], np.random.default_rng(11), batch=2)
    run_evaluator_flow(ref, emul_lib, ref.BFV, 512, [36] * 9 + [40], np.random.default_rng(12))


@pytest.mark.parametrize("scheme,n,bits", [(2, 64, [40, 30, 30, 40]), (1, 128, [40, 30, 30, 40]), (3, 4096, [60, 40, 40, 60]),
                                           (2, 16384, [60, 40, 60]), (1, 4096, [60, 40, 40, 60])])
def test_plain_ops_emulated(ref, emul_lib, scheme, n, bits):
    """multiplyPlain / addPlain / subPlain and the fused multiply-accumulate."""
    run_plain_flow(ref, emul_lib, scheme, n, bits, np.random.default_rng(300 + scheme + n))


def test_matmul_emulated(ref, emul_lib):
    """The one-launch HE matmul against the reference's multiplyPlain + addInplace loop; odd grids hit the edge tiles."""
    run_matmul_flow(ref, emul_lib, 256, [40, 30, 30, 40], np.random.default_rng(41), B=3, I=3, J=3)
    run_matmul_flow(ref, emul_lib, 4096, [60, 40, 60], np.random.default_rng(42), B=2, I=40, J=1, level=1)
    run_matmul_flow(ref, emul_lib, 1024, [40, 40, 36, 60], np.random.default_rng(43), B=3, I=21, J=2)
    run_matmul_flow(ref, emul_lib, 512, [60, 40, 40, 60], np.random.default_rng(44), B=2, I=19, J=3, extreme=True)


def test_matmul_block_layout():
    """determineBlock and the slot layout of MatmulHelper (app/LinearHelperCKKS.cuh:113-167,291-311): multiplying the
    coefficient vectors as polynomials modulo X^N + 1 and reading the outputs back gives x @ W."""
    from troy_b200.linear import MatmulHelper
    rng = np.random.default_rng(3)
    for (bs, din, dout, slots) in [(2, 5, 7, 16), (3, 40, 9, 32), (1, 64, 64, 64)]:
        h = MatmulHelper(bs, din, dout, slots)
        n = 2 * slots
        x = rng.integers(-5, 6, (bs, din)).astype(np.float64)
        w = rng.integers(-5, 6, (din, dout)).astype(np.float64)
        xv, wv = h.input_vectors(x), h.weight_vectors(w)
        I, J = h.grid
        out = np.zeros((bs, J, n))
        for b in range(bs):
            for j in range(J):
                for i in range(I):
                    full = np.convolve(xv[b, i], wv[i, j])
                    out[b, j] += full[:n] - np.concatenate([full[n:], np.zeros(2 * n - len(full))])[:n]
        assert np.array_equal(h.read_outputs(out), x @ w)


@pytest.mark.parametrize("scheme,n,bits", [(2, 64, [40, 30, 30, 40]), (1, 128, [40, 30, 30, 40]), (1, 4096, [60, 40, 40, 60]),
                                           (2, 16384, [60, 40, 60]), (3, 256, [40, 30, 30, 40]), (3, 4096, [60, 40, 40, 60])])
def test_decrypt_emulated(ref, emul_lib, scheme, n, bits):
    """CKKS / BFV / BGV decryption on the device against the reference CPU decryptor."""
    run_decrypt_flow(ref, emul_lib, scheme, n, bits, np.random.default_rng(500 + scheme + n))


@pytest.mark.parametrize("scheme,n,bits", [(2, 4096, [50, 50, 49, 48, 47, 50]), (1, 8192, [50, 50, 50, 50]), (3, 16384, [50, 49, 50]),
                                           (2, 65536, [50, 50, 50])])
def test_fp64_butterflies_for_50_bit_primes_emulated(ref, emul_lib, scheme, n, bits):
    """46..50-bit primes run the FP64 butterflies with per-butterfly re-centring (ArithF64L): no head-room to spare, so this
    is checked on parameter sets made only of such primes, at every transform size class."""
    run_evaluator_flow(ref, emul_lib, scheme, n, bits, np.random.default_rng(600 + scheme + n), batch=2)


def test_prng_stream_emulated(ref, emul_lib):
    from common import run_prng_flow
    run_prng_flow(ref, emul_lib, np.random.default_rng(700))


@pytest.mark.parametrize("scheme,n,bits", [(2, 1024, [60, 40, 40, 60]), (1, 256, [40, 30, 30, 40]), (3, 4096, [60, 50, 60])])
def test_samplers_emulated(ref, emul_lib, scheme, n, bits):
    """The device samplers against the reference's, including heavy rejection."""
    from common import run_sampler_flow
    run_sampler_flow(ref, emul_lib, scheme, n, bits, np.random.default_rng(710 + scheme))


@pytest.mark.parametrize("scheme,n,bits", [(2, 1024, [60, 40, 40, 60]), (1, 256, [40, 30, 30, 40]), (3, 4096, [60, 40, 60]),
                                           (2, 256, [40]), (1, 16384, [60, 40, 60]), (2, 64, [40, 30, 30, 40]), (3, 32, [40, 30, 40])])
def test_encrypt_emulated(ref, emul_lib, scheme, n, bits):
    """Encryption on the device: same seed, same words as the reference CPU encryptor."""
    from common import run_encrypt_flow
    run_encrypt_flow(ref, emul_lib, scheme, n, bits, np.random.default_rng(720 + scheme + n))


def test_encryptor_error_behaviour_emulated(ref, emul_lib):
    """Encryptor's argument checks (src/encryptor.cpp:88-253): missing keys, levels outside the chain, plaintext form."""
    from common import make_pair
    from troy_b200.lib import TroyInvalidArgument, TroyLogicError
    tb = emul_lib
    rc, ctx = make_pair(ref, tb, ref.CKKS, 256, [40, 30, 30, 40])
    rc.keygen()
    only_pk = tb.EncryptorCuda(ctx, rc.public_key_words(), None)
    with pytest.raises(TroyLogicError):
        only_pk.encrypt_zero_symmetric(0)
    only_sk = tb.EncryptorCuda(ctx, None, rc.secret_key_words())
    with pytest.raises(TroyLogicError):
        only_sk.encrypt_zero(0)
    with pytest.raises(TroyInvalidArgument):
        only_pk.encrypt_zero(ctx.num_levels)          # parms_id is not valid for encryption parameters
    with pytest.raises(TroyInvalidArgument):
        only_pk.encrypt_zero(0, batch=2, seeds=np.zeros(8, dtype=np.uint64))  # one seed per ciphertext
    with pytest.raises(TroyInvalidArgument):
        tb.EncryptorCuda(ctx, np.zeros(5, dtype=np.uint64), None)             # key of the wrong size
    coeff = tb.PlaintextCuda.from_host(ctx, np.arange(16, dtype=np.uint64))   # coefficient form: CKKS wants NTT form
    with pytest.raises(TroyInvalidArgument):
        only_pk.encrypt(coeff)
    # every level of the chain, asymmetric: the limb count follows the level
    for level in range(ctx.num_levels):
        assert only_pk.encrypt_zero(level).limbs == ctx.level_limbs(level)


@pytest.mark.parametrize("scheme,n,bits", [(1, 256, [40, 30, 30, 40]), (2, 64, [40, 30, 30, 40])])
def test_negacyclic_shift_and_divide_by_degree_emulated(ref, emul_lib, scheme, n, bits):
    from common import run_lwe_kernels_flow
    run_lwe_kernels_flow(ref, emul_lib, scheme, n, bits)


@pytest.mark.parametrize("scheme,n,bits", [(2, 1024, [40] * 8 + [60]), (1, 512, [60] + [40] * 7 + [60]), (2, 2048, [30, 40, 36, 50]), (3, 512, [40] * 9)])
def test_keyswitch_with_extreme_operands_emulated(ref, emul_lib, scheme, n, bits):
    from common import run_extreme_keyswitch_flow
    run_extreme_keyswitch_flow(ref, emul_lib, scheme, n, bits, np.random.default_rng(7))
