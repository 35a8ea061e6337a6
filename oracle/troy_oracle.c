/* TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's RNS polynomial hot path in plain C.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product (troy_b200/) never does.
 * Each function restates the *integer function* computed by the reference code it cites (file:line under
 * /root/reference), written for obviousness, not speed: every modular product is `(unsigned __int128) a * b % q`,
 * every output is a canonical residue. Parity pinning: tests/test_oracle_cpu.py checks every function here against the
 * reference's own known-answer vectors (tests/golden/kat_*.json, copied values from test/utils/ntt.cpp:52-99 and
 * test/utils/galois.cpp) and word-for-word against outputs of the reference library itself (oracle/_ref, and the
 * committed fixtures tests/golden/*.npz generated from it by tests/golden/make_golden.py).
 */
#include "troy_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

static uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)((u128)a * b % q); }
static uint64_t addmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a + b) % q); }
static uint64_t submod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a + q - b % q) % q); }

static uint64_t powmod(uint64_t a, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = mulmod(r, a, q);
        a = mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}

/* util::tryInvertUintMod (src/utils/numth.h): modular inverse, also for the composite m_tilde = 2^32 */
static uint64_t invmod(uint64_t a, uint64_t q) {
    __int128 t = 0, nt = 1, r = q, nr = a % q;
    while (nr != 0) {
        __int128 qt = r / nr, tmp = t - qt * nt;
        t = nt; nt = tmp;
        tmp = r - qt * nr; r = nr; nr = tmp;
    }
    if (t < 0) t += q;
    return (uint64_t)t;
}

static uint32_t bitrev(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

static int ilog2(uint64_t n) {
    int l = 0;
    while ((1ull << l) < n) l++;
    return l;
}

/* util::isPrime (src/utils/numth.cpp:215-259): any correct primality test gives the same answers */
int orc_is_prime(uint64_t n) {
    static const uint64_t bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return 0;
    for (int i = 0; i < 12; i++)
        if (n % bases[i] == 0) return n == bases[i];
    uint64_t d = n - 1;
    int r = 0;
    while (!(d & 1)) { d >>= 1; r++; }
    for (int i = 0; i < 12; i++) {
        uint64_t x = powmod(bases[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int j = 1; j < r; j++) {
            x = mulmod(x, x, n);
            if (x == n - 1) { comp = 0; break; }
        }
        if (comp) return 0;
    }
    return 1;
}

/* util::getPrimes (src/utils/numth.cpp:261-284): primes = 1 mod factor, descending from 2^bit_size */
int orc_get_primes(uint64_t factor, int bit_size, int count, uint64_t *out) {
    uint64_t value = ((1ull << bit_size) - 1) / factor * factor + 1;
    uint64_t lower = 1ull << (bit_size - 1);
    int found = 0;
    while (found < count && value > lower) {
        if (orc_is_prime(value)) out[found++] = value;
        value -= factor;
    }
    return found == count ? 0 : -1;
}

/* CoeffModulus::Create (src/modulus.cpp:80-122): per distinct bit size take the needed number of primes with
 * getPrimes(2N, bits, count) and hand them out from the BACK of each list in the order the sizes are requested. */
int orc_coeff_modulus_create(uint64_t n, const int *bits, int count, uint64_t *out) {
    for (int i = 0; i < count; i++) out[i] = 0;
    for (int i = 0; i < count; i++) {
        if (out[i]) continue;
        int same = 0;
        for (int j = 0; j < count; j++) same += bits[j] == bits[i];
        uint64_t *list = (uint64_t *)malloc(sizeof(uint64_t) * same);
        if (orc_get_primes(2 * n, bits[i], same, list) != 0) { free(list); return -1; }
        int back = same;
        for (int j = 0; j < count; j++)
            if (bits[j] == bits[i]) out[j] = list[--back];
        free(list);
    }
    return 0;
}

/* util::tryMinimalPrimitiveRoot (src/utils/numth.cpp:335-360): the smallest primitive degree-th root of unity */
uint64_t orc_minimal_primitive_root(uint64_t degree, uint64_t q) {
    uint64_t root = 0;
    for (uint64_t g = 2; g < 4096 && !root; g++) {
        uint64_t c = powmod(g, (q - 1) / degree, q);
        if (powmod(c, degree / 2, q) == q - 1) root = c;
    }
    uint64_t sq = mulmod(root, root, q), cur = root, best = root;
    for (uint64_t i = 0; i < degree; i += 2) {
        if (cur < best) best = cur;
        cur = mulmod(cur, sq, q);
    }
    return best;
}

/* NTTTables::initialize (src/utils/ntt.cpp:17-66): root_powers[bitrev(i)] = psi^i ; inv_root_powers[bitrev(i-1)+1] = psi^-i */
void orc_ntt_tables(uint64_t n, uint64_t q, uint64_t *root_powers, uint64_t *inv_root_powers) {
    int logn = ilog2(n);
    uint64_t psi = orc_minimal_primitive_root(2 * n, q), ipsi = invmod(psi, q);
    uint64_t p = psi, ip = ipsi;
    if (root_powers) root_powers[0] = 1;
    if (inv_root_powers) inv_root_powers[0] = 1;
    for (uint64_t i = 1; i < n; i++) {
        if (root_powers) root_powers[bitrev((uint32_t)i, logn)] = p;
        if (inv_root_powers) inv_root_powers[bitrev((uint32_t)(i - 1), logn) + 1] = ip;
        p = mulmod(p, psi, q);
        ip = mulmod(ip, ipsi, q);
    }
}

/* nttNegacyclicHarvey (src/utils/ntt.cpp:158-187) = DWTHandler::transformToRev (src/utils/dwthandler.h:88-204):
 * Cooley-Tukey layers, natural order in, bit-reversed order out. Lazy ranges are irrelevant here: exact arithmetic. */
void orc_ntt_forward(uint64_t n, uint64_t q, uint64_t *a) {
    uint64_t *rp = (uint64_t *)malloc(sizeof(uint64_t) * n);
    orc_ntt_tables(n, q, rp, NULL);
    size_t root_index = 1;
    for (size_t gap = n >> 1, m = 1; m < n; m <<= 1, gap >>= 1) {
        size_t offset = 0;
        for (size_t i = 0; i < m; i++) {
            uint64_t r = rp[root_index++];
            for (size_t j = 0; j < gap; j++) {
                uint64_t u = a[offset + j] % q, v = mulmod(a[offset + j + gap] % q, r, q);
                a[offset + j] = addmod(u, v, q);
                a[offset + j + gap] = submod(u, v, q);
            }
            offset += gap << 1;
        }
    }
    free(rp);
}

/* inverseNttNegacyclicHarvey (src/utils/ntt.cpp:189-213) = transformFromRev (src/utils/dwthandler.h:215-372):
 * Gentleman-Sande layers with the scrambled inverse table, then x N^-1. */
void orc_ntt_inverse(uint64_t n, uint64_t q, uint64_t *a) {
    uint64_t *irp = (uint64_t *)malloc(sizeof(uint64_t) * n);
    orc_ntt_tables(n, q, NULL, irp);
    size_t root_index = 1;
    for (size_t gap = 1, m = n >> 1; m >= 1; m >>= 1, gap <<= 1) {
        size_t offset = 0;
        for (size_t i = 0; i < m; i++) {
            uint64_t r = irp[root_index++];
            for (size_t j = 0; j < gap; j++) {
                uint64_t u = a[offset + j] % q, v = a[offset + j + gap] % q;
                a[offset + j] = addmod(u, v, q);
                a[offset + j + gap] = mulmod(submod(u, v, q), r, q);
            }
            offset += gap << 1;
        }
        if (m == 1) break;
    }
    uint64_t ninv = invmod(n % q, q);
    for (size_t i = 0; i < n; i++) a[i] = mulmod(a[i], ninv, q);
    free(irp);
}

/* dyadic tensor product of ciphertexts in NTT form, one limb (gDyadicConvolutionCoeffmod src/kernelutils.cu:89-130;
 * ckksMultiply src/evaluator.cpp:626-702): out[i] = sum_{j+k=i} a[j]*b[k] mod q */
void orc_dyadic_tensor(uint64_t n, uint64_t q, const uint64_t *a, int a_size, const uint64_t *b, int b_size, uint64_t *out,
                       size_t poly_stride_a, size_t poly_stride_b, size_t poly_stride_out) {
    for (int i = 0; i < a_size + b_size - 1; i++)
        for (size_t c = 0; c < n; c++) {
            uint64_t acc = 0;
            for (int j = 0; j < a_size; j++) {
                int k = i - j;
                if (k < 0 || k >= b_size) continue;
                acc = addmod(acc, mulmod(a[j * poly_stride_a + c] % q, b[k * poly_stride_b + c] % q, q), q);
            }
            out[i * poly_stride_out + c] = acc;
        }
}

/* GaloisTool::applyGalois, coefficient form (src/utils/galois.cpp:141-175): out[(i*elt) mod N] = +-in[i] */
void orc_galois_coeff(uint64_t n, uint64_t q, const uint64_t *in, uint32_t elt, uint64_t *out) {
    uint64_t m = 2 * n;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t raw = (i * elt) & (m - 1);
        uint64_t v = in[i];
        if (raw >= n) v = v ? q - v : 0;
        out[raw & (n - 1)] = v;
    }
}

/* GaloisTool::applyGaloisNtt + generateTableNtt (src/utils/galois.cpp:20-40,177-196):
 * table[i] = bitrev_logN( ((elt * bitrev_{logN+1}(i + N)) >> 1) & (N-1) ) ; out[i] = in[table[i]] */
void orc_galois_ntt(uint64_t n, const uint64_t *in, uint32_t elt, uint64_t *out) {
    int logn = ilog2(n);
    for (uint64_t i = 0; i < n; i++) {
        uint32_t reversed = bitrev((uint32_t)(i + n), logn + 1);
        uint64_t index_raw = ((uint64_t)elt * reversed) >> 1;
        index_raw &= n - 1;
        out[i] = in[bitrev((uint32_t)index_raw, logn)];
    }
}

/* GaloisTool::getEltFromStep (src/utils/galois.cpp:42-92) */
uint32_t orc_galois_elt_from_step(uint64_t n, int step) {
    uint32_t m = (uint32_t)(2 * n);
    if (step == 0) return m - 1;
    int neg = step < 0;
    uint32_t pos = (uint32_t)(neg ? -step : step);
    if (neg) pos = (uint32_t)(n >> 1) - pos;
    uint64_t g = 1;
    while (pos--) g = (g * 3) & (m - 1);
    return (uint32_t)g;
}

/* Divide-and-round by the last prime of the base {q_0..q_{L-1}} on one polynomial [L][N]; results in limbs 0..L-2.
 *   mode ORC_DROP_BFV  : RNSTool::divideAndRoundqLastInplace      src/utils/rns.cpp:805-829 (coefficient form)
 *   mode ORC_DROP_CKKS : RNSTool::divideAndRoundqLastNttInplace   src/utils/rns.cpp:831-876 (NTT form)
 *   mode ORC_DROP_BGV  : RNSTool::modTAndDivideqLastInplace       src/utils/rns.cpp:1097-1140 */
void orc_drop_last(int mode, uint64_t n, int L, const uint64_t *q, uint64_t t, uint64_t *poly) {
    uint64_t ql = q[L - 1], half = ql >> 1;
    uint64_t *last = poly + (size_t)(L - 1) * n;
    uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * n);
    if (mode == ORC_DROP_CKKS) orc_ntt_inverse(n, ql, last);
    if (mode != ORC_DROP_BGV)
        for (size_t c = 0; c < n; c++) last[c] = addmod(last[c], half, ql);
    for (int i = 0; i < L - 1; i++) {
        uint64_t qi = q[i], inv = invmod(ql % qi, qi);
        uint64_t *x = poly + (size_t)i * n;
        if (mode == ORC_DROP_BGV) {
            uint64_t inv_t = invmod(ql % t, t);
            for (size_t c = 0; c < n; c++) {
                uint64_t neg = mulmod(submod(0, last[c] % t, t), inv_t, t);  /* -c_last * q_last^-1 mod t */
                uint64_t delta = mulmod(neg % qi, ql % qi, qi);
                x[c] = mulmod(submod(submod(x[c], last[c] % qi, qi), delta, qi), inv, qi);
            }
            continue;
        }
        for (size_t c = 0; c < n; c++) tmp[c] = submod(last[c] % qi, half % qi, qi);
        if (mode == ORC_DROP_CKKS) orc_ntt_forward(n, qi, tmp);
        for (size_t c = 0; c < n; c++) x[c] = mulmod(submod(x[c], tmp[c], qi), inv, qi);
    }
    free(tmp);
}

/* Evaluator::switchKeyInplace (src/evaluator.cpp:2310-2653; CUDA twin src/evaluator_cuda.cu:1163-1378).
 *   scheme 1 BFV, 2 CKKS, 3 BGV; L data limbs in play; key_q = key-level moduli (Ltop data primes + special prime);
 *   ct [2][L][N] (updated in place: the switched value is added), target [L][N],
 *   key [Ltop digits][2][Ltop+1][N] in NTT form. */
void orc_switch_key(int scheme, uint64_t n, int L, int Ltop, const uint64_t *key_q, uint64_t t, uint64_t *ct,
                    const uint64_t *target, const uint64_t *key) {
    const int KL = Ltop + 1;
    const uint64_t qsp = key_q[Ltop];
    size_t N = n;
    uint64_t *tt = (uint64_t *)malloc(sizeof(uint64_t) * L * N);       /* t_target, coefficient form */
    uint64_t *prod = (uint64_t *)calloc((size_t)2 * (L + 1) * N, sizeof(uint64_t)); /* t_poly_prod [k][L+1][N] */
    uint64_t *op = (uint64_t *)malloc(sizeof(uint64_t) * N);
    memcpy(tt, target, sizeof(uint64_t) * L * N);
    if (scheme == 2)
        for (int j = 0; j < L; j++) orc_ntt_inverse(n, key_q[j], tt + j * N);
    for (int i = 0; i <= L; i++) {
        int key_index = (i == L) ? Ltop : i;
        uint64_t qi = key_q[key_index];
        for (int j = 0; j < L; j++) {
            if (scheme == 2 && i == j) {
                memcpy(op, target + j * N, sizeof(uint64_t) * N);
            } else {
                for (size_t c = 0; c < N; c++) op[c] = tt[j * N + c] % qi;
                orc_ntt_forward(n, qi, op);
            }
            for (int k = 0; k < 2; k++) {
                const uint64_t *kp = key + (((size_t)j * 2 + k) * KL + key_index) * N;
                uint64_t *acc = prod + ((size_t)k * (L + 1) + i) * N;
                for (size_t c = 0; c < N; c++) acc[c] = addmod(acc[c], mulmod(op[c], kp[c], qi), qi);
            }
        }
    }
    /* mod-down by the special prime */
    for (int k = 0; k < 2; k++) {
        uint64_t *last = prod + ((size_t)k * (L + 1) + L) * N;
        orc_ntt_inverse(n, qsp, last);
        uint64_t half = qsp >> 1;
        if (scheme != 3)
            for (size_t c = 0; c < N; c++) last[c] = addmod(last[c], half, qsp);
        for (int j = 0; j < L; j++) {
            uint64_t qj = key_q[j], inv = invmod(qsp % qj, qj);
            uint64_t *pj = prod + ((size_t)k * (L + 1) + j) * N;
            uint64_t *cj = ct + ((size_t)k * L + j) * N;
            if (scheme == 3) {
                uint64_t inv_t = invmod(qsp % t, t);
                orc_ntt_inverse(n, qj, pj);
                for (size_t c = 0; c < N; c++) {
                    uint64_t kk = mulmod(submod(0, last[c] % t, t), inv_t, t);
                    uint64_t delta = mulmod(kk % qj, qsp % qj, qj);
                    uint64_t d = submod(submod(pj[c], delta, qj), last[c] % qj, qj);
                    cj[c] = addmod(cj[c], mulmod(d, inv, qj), qj);
                }
                continue;
            }
            for (size_t c = 0; c < N; c++) op[c] = submod(last[c] % qj, half % qj, qj);
            if (scheme == 2) orc_ntt_forward(n, qj, op);
            else orc_ntt_inverse(n, qj, pj);
            for (size_t c = 0; c < N; c++) cj[c] = addmod(cj[c], mulmod(submod(pj[c], op[c], qj), inv, qj), qj);
        }
    }
    free(tt); free(prod); free(op);
}

/* ---- BEHZ multiply (Evaluator::bfvMultiply src/evaluator.cpp:385-624; RNSTool src/utils/rns.cpp:581-1095) ---- */

static uint64_t punct(const uint64_t *base, int n, int skip, uint64_t m) { /* prod_{k != skip} base[k] mod m */
    uint64_t r = 1 % m;
    for (int k = 0; k < n; k++)
        if (k != skip) r = mulmod(r, base[k] % m, m);
    return r;
}

/* BaseConverter::fastConvertArray (src/utils/rns.cpp:415-452) for one coefficient: in[ni] (base ib) -> out[no] (base ob) */
static void fast_convert(const uint64_t *ib, int ni, const uint64_t *ob, int no, const uint64_t *in, uint64_t *out) {
    uint64_t y[ORC_MAX_LIMBS];
    for (int i = 0; i < ni; i++) y[i] = mulmod(in[i] % ib[i], invmod(punct(ib, ni, i, ib[i]), ib[i]), ib[i]);
    for (int o = 0; o < no; o++) {
        uint64_t acc = 0;
        for (int i = 0; i < ni; i++) acc = addmod(acc, mulmod(y[i], punct(ib, ni, i, ob[o]), ob[o]), ob[o]);
        out[o] = acc;
    }
}

static int total_bits(const uint64_t *q, int L) {
    uint64_t acc[ORC_MAX_LIMBS + 1];
    int len = 1;
    acc[0] = 1;
    for (int i = 0; i < L; i++) {
        uint64_t carry = 0;
        for (int w = 0; w < len; w++) {
            u128 t = (u128)acc[w] * q[i] + carry;
            acc[w] = (uint64_t)t;
            carry = (uint64_t)(t >> 64);
        }
        if (carry) acc[len++] = carry;
    }
    int b = 0;
    while (acc[len - 1] >> b) b++;
    return (len - 1) * 64 + b;
}

int orc_behz_base(uint64_t n, int L, const uint64_t *q, uint64_t t, uint64_t *bsk_out) {
    /* RNSTool::initialize src/utils/rns.cpp:604-650 */
    int tb = 0;
    while (t >> tb) tb++;
    int nB = L;
    if (32 + tb + total_bits(q, L) >= 61 * L + 61) nB++;
    uint64_t aux[ORC_MAX_LIMBS + 3];
    if (orc_get_primes(2 * n, 61, nB + 2, aux) != 0) return -1;
    /* aux = m_sk, gamma, B_0.. ; Bsk = B ∪ {m_sk} */
    for (int b = 0; b < nB; b++) bsk_out[b] = aux[2 + b];
    bsk_out[nB] = aux[0];
    return nB + 1;
}

int orc_bfv_multiply(uint64_t n, int L, const uint64_t *q, uint64_t t, const uint64_t *a, int a_size, const uint64_t *b,
                     int b_size, uint64_t *out) {
    size_t N = n;
    uint64_t bsk[ORC_MAX_LIMBS];
    int nBsk = orc_behz_base(n, L, q, t, bsk);
    if (nBsk < 0) return -1;
    int nB = nBsk - 1, E = L + nBsk;
    uint64_t msk = bsk[nB];
    const uint64_t mt = 1ull << 32;
    uint64_t ext[ORC_MAX_LIMBS * 2];
    for (int i = 0; i < L; i++) ext[i] = q[i];
    for (int o = 0; o < nBsk; o++) ext[L + o] = bsk[o];
    int out_size = a_size + b_size - 1;
    uint64_t *ea = (uint64_t *)malloc(sizeof(uint64_t) * a_size * E * N);
    uint64_t *eb = (uint64_t *)malloc(sizeof(uint64_t) * b_size * E * N);
    uint64_t *pr = (uint64_t *)malloc(sizeof(uint64_t) * out_size * E * N);
    for (int which = 0; which < 2; which++) {
        const uint64_t *src = which ? b : a;
        uint64_t *dst = which ? eb : ea;
        int size = which ? b_size : a_size;
        for (int p = 0; p < size; p++)
            for (size_t c = 0; c < N; c++) {
                uint64_t in[ORC_MAX_LIMBS], tmp[ORC_MAX_LIMBS], conv[ORC_MAX_LIMBS], mtv;
                for (int i = 0; i < L; i++) {
                    in[i] = src[((size_t)p * L + i) * N + c];
                    dst[((size_t)p * E + i) * N + c] = in[i];
                    tmp[i] = mulmod(in[i], mt % q[i], q[i]); /* fastbconvmTilde: x * m_tilde first (rns.cpp:1012-1040) */
                }
                fast_convert(q, L, bsk, nBsk, tmp, conv);
                fast_convert(q, L, &mt, 1, tmp, &mtv);
                /* smMrq (rns.cpp:943-983) */
                uint64_t r = mulmod(mtv, submod(0, invmod(punct(q, L, -1, mt), mt), mt), mt);
                for (int o = 0; o < nBsk; o++) {
                    uint64_t m = bsk[o], tv = r;
                    if (tv >= (mt >> 1)) tv += m - mt;
                    uint64_t s = addmod(mulmod(tv % m, punct(q, L, -1, m), m), conv[o], m);
                    dst[((size_t)p * E + L + o) * N + c] = mulmod(s, invmod(mt % m, m), m);
                }
            }
        for (int p = 0; p < size; p++)
            for (int l = 0; l < E; l++) orc_ntt_forward(n, ext[l], dst + ((size_t)p * E + l) * N);
    }
    for (int l = 0; l < E; l++)
        orc_dyadic_tensor(n, ext[l], ea + (size_t)l * N, a_size, eb + (size_t)l * N, b_size, pr + (size_t)l * N, (size_t)E * N,
                          (size_t)E * N, (size_t)E * N);
    for (int p = 0; p < out_size; p++)
        for (int l = 0; l < E; l++) orc_ntt_inverse(n, ext[l], pr + ((size_t)p * E + l) * N);
    for (int p = 0; p < out_size; p++)
        for (size_t c = 0; c < N; c++) {
            uint64_t xq[ORC_MAX_LIMBS], conv[ORC_MAX_LIMBS], w[ORC_MAX_LIMBS], dq[ORC_MAX_LIMBS], tsk;
            /* multiply by t (evaluator.cpp:596-607), fastFloor (rns.cpp:985-1010) */
            for (int i = 0; i < L; i++) xq[i] = mulmod(pr[((size_t)p * E + i) * N + c], t % q[i], q[i]);
            fast_convert(q, L, bsk, nBsk, xq, conv);
            for (int o = 0; o < nBsk; o++) {
                uint64_t m = bsk[o];
                uint64_t z = mulmod(pr[((size_t)p * E + L + o) * N + c], t % m, m);
                w[o] = mulmod(submod(z, conv[o], m), invmod(punct(q, L, -1, m), m), m);
            }
            /* fastbconvSk (rns.cpp:879-941) */
            fast_convert(bsk, nB, q, L, w, dq);
            fast_convert(bsk, nB, &msk, 1, w, &tsk);
            uint64_t alpha = mulmod(submod(tsk, w[nB], msk), invmod(punct(bsk, nB, -1, msk), msk), msk);
            for (int i = 0; i < L; i++) {
                uint64_t pb = punct(bsk, nB, -1, q[i]);
                uint64_t r;
                if (alpha > (msk >> 1)) r = addmod(mulmod((msk - alpha) % q[i], pb, q[i]), dq[i], q[i]);
                else r = addmod(mulmod(alpha % q[i], q[i] - pb, q[i]), dq[i], q[i]);
                out[((size_t)p * L + i) * N + c] = r;
            }
        }
    free(ea); free(eb); free(pr);
    return 0;
}

/* ckksMultiply / bgvMultiply (src/evaluator.cpp:626-822): CKKS operands are already in NTT form; BGV operands are
 * transformed, multiplied and transformed back. a, b, out: [size][L][N]. */
void orc_multiply_ntt_schemes(int scheme, uint64_t n, int L, const uint64_t *q, const uint64_t *a, int a_size, const uint64_t *b,
                              int b_size, uint64_t *out) {
    size_t N = n;
    int out_size = a_size + b_size - 1;
    uint64_t *ta = (uint64_t *)malloc(sizeof(uint64_t) * a_size * L * N);
    uint64_t *tb = (uint64_t *)malloc(sizeof(uint64_t) * b_size * L * N);
    memcpy(ta, a, sizeof(uint64_t) * a_size * L * N);
    memcpy(tb, b, sizeof(uint64_t) * b_size * L * N);
    if (scheme == 3) {
        for (int p = 0; p < a_size; p++)
            for (int l = 0; l < L; l++) orc_ntt_forward(n, q[l], ta + ((size_t)p * L + l) * N);
        for (int p = 0; p < b_size; p++)
            for (int l = 0; l < L; l++) orc_ntt_forward(n, q[l], tb + ((size_t)p * L + l) * N);
    }
    for (int l = 0; l < L; l++)
        orc_dyadic_tensor(n, q[l], ta + (size_t)l * N, a_size, tb + (size_t)l * N, b_size, out + (size_t)l * N, (size_t)L * N,
                          (size_t)L * N, (size_t)L * N);
    if (scheme == 3)
        for (int p = 0; p < out_size; p++)
            for (int l = 0; l < L; l++) orc_ntt_inverse(n, q[l], out + ((size_t)p * L + l) * N);
    free(ta); free(tb);
}

/* Evaluator::applyGaloisInplace (src/evaluator.cpp:2153-2249): permute c0 and c1, zero c1, key-switch the permuted c1 */
void orc_apply_galois(int scheme, uint64_t n, int L, int Ltop, const uint64_t *key_q, uint64_t t, uint64_t *ct, uint32_t elt,
                      const uint64_t *key) {
    size_t N = n;
    uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * 2 * L * N);
    for (int p = 0; p < 2; p++)
        for (int l = 0; l < L; l++) {
            const uint64_t *in = ct + ((size_t)p * L + l) * N;
            uint64_t *o = tmp + ((size_t)p * L + l) * N;
            if (scheme == 2) orc_galois_ntt(n, in, elt, o);
            else orc_galois_coeff(n, key_q[l], in, elt, o);
        }
    memcpy(ct, tmp, sizeof(uint64_t) * L * N);
    memset(ct + (size_t)L * N, 0, sizeof(uint64_t) * L * N);
    orc_switch_key(scheme, n, L, Ltop, key_q, t, ct, tmp + (size_t)L * N, key);
    free(tmp);
}
