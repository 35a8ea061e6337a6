// Encrypt-zero on the device with the HOST generator's semantics.
//
// The reference's CPU encryptor (src/utils/rlwe.cpp:160-333) draws everything from one sequential byte stream, the
// Blake2xb generator of src/randomgen.cpp:184-194: 4096-byte buffers, buffer c = BLAKE2Xb(out 4096 B, input = c as 8
// little-endian bytes, key = the 64-byte seed). The stream is therefore RANDOM ACCESS by 64-byte block (one BLAKE2b
// compression of the buffer's root hash per block), and the only sequential part of the samplers is where rejections
// shift the positions of later draws. The kernels below restate that: the stream is expanded in parallel into a scratch
// buffer, rejections are resolved by counting + prefix sums, and every coefficient is then produced independently.
// (The reference's own CUDA encryptor, src/utils/rlwe_cuda.cu:193-333, draws from cuRAND and does not reproduce its CPU
// path; parity here is against the CPU path.)
//
// BLAKE2b / BLAKE2X are restated from their specifications (RFC 7693; "BLAKE2X", Aumasson et al. 2016), the construction
// src/utils/blake2b.c / blake2xb.c implement.
#pragma once
#include "launch.cuh"
#if !defined(TROYB_HOST_EMUL)
#include <curand_kernel.h>
#endif

namespace troyb {

// ---- BLAKE2b compression, BLAKE2Xb root / output blocks -------------------------------------------------------------------
struct Blake2 {
    static TB_HD u64 iv(int i) {
        constexpr u64 k[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                              0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
        return k[i];
    }
    static TB_HD u64 rotr(u64 x, int n) { return (x >> n) | (x << (64 - n)); }
    static TB_HD void g(u64 &a, u64 &b, u64 &c, u64 &d, u64 x, u64 y) {
        a = a + b + x; d = rotr(d ^ a, 32);
        c = c + d;     b = rotr(b ^ c, 24);
        a = a + b + y; d = rotr(d ^ a, 16);
        c = c + d;     b = rotr(b ^ c, 63);
    }
    // One compression of the 128-byte block m with byte counter t (< 2^64) and the last-block flag.
    static TB_HD void compress(u64 h[8], const u64 m[16], u64 t, bool last) {
        constexpr unsigned char sigma[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        u64 v[16];
#pragma unroll
        for (int i = 0; i < 8; i++) { v[i] = h[i]; v[8 + i] = iv(i); }
        v[12] ^= t;
        if (last) v[14] = ~v[14];
#pragma unroll
        for (int r = 0; r < 12; r++) {
            g(v[0], v[4], v[8], v[12], m[sigma[r][0]], m[sigma[r][1]]);
            g(v[1], v[5], v[9], v[13], m[sigma[r][2]], m[sigma[r][3]]);
            g(v[2], v[6], v[10], v[14], m[sigma[r][4]], m[sigma[r][5]]);
            g(v[3], v[7], v[11], v[15], m[sigma[r][6]], m[sigma[r][7]]);
            g(v[0], v[5], v[10], v[15], m[sigma[r][8]], m[sigma[r][9]]);
            g(v[1], v[6], v[11], v[12], m[sigma[r][10]], m[sigma[r][11]]);
            g(v[2], v[7], v[8], v[13], m[sigma[r][12]], m[sigma[r][13]]);
            g(v[3], v[4], v[9], v[14], m[sigma[r][14]], m[sigma[r][15]]);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[8 + i];
    }
    static constexpr u64 kBufferBytes = 4096; // UniformRandomGenerator::buffer_size_  src/randomgen.h:376

    // Root hash of buffer `counter`: keyed BLAKE2b (digest 64, key 64, fanout 1, depth 1, xof_length 4096) over the counter.
    static TB_HD void root(const u64 seed[8], u64 counter, u64 out[8]) {
        u64 h[8], m[16];
#pragma unroll
        for (int i = 0; i < 8; i++) h[i] = iv(i);
        h[0] ^= 0x01014040ull;            // digest_length 64 | key_length 64 << 8 | fanout 1 << 16 | depth 1 << 24, leaf_length 0
        h[1] ^= kBufferBytes << 32;       // node_offset 0, xof_length 4096
#pragma unroll
        for (int i = 0; i < 16; i++) m[i] = i < 8 ? seed[i] : 0;
        compress(h, m, 128, false);       // the key, padded to one block
#pragma unroll
        for (int i = 0; i < 16; i++) m[i] = 0;
        m[0] = counter;
        compress(h, m, 136, true);        // the 8 input bytes
#pragma unroll
        for (int i = 0; i < 8; i++) out[i] = h[i];
    }
    // Output block i (64 bytes) of the buffer with root hash `rt`: BLAKE2b (digest 64, fanout 0, depth 0, leaf_length 64,
    // node_offset i, xof_length 4096, inner_length 64) over the root hash.
    static TB_HD void block(const u64 rt[8], u32 i, u64 out[8]) {
        u64 h[8], m[16];
#pragma unroll
        for (int k = 0; k < 8; k++) h[k] = iv(k);
        h[0] ^= 0x40ull | (64ull << 32);
        h[1] ^= (u64)i | (kBufferBytes << 32);
        h[2] ^= 64ull << 8;
#pragma unroll
        for (int k = 0; k < 16; k++) m[k] = k < 8 ? rt[k] : 0;
        compress(h, m, 64, true);
#pragma unroll
        for (int k = 0; k < 8; k++) out[k] = h[k];
    }
};

// Read access to one generator's byte stream: the expanded prefix `buf` (`bytes` long, a multiple of 64) and, past it, blocks
// computed on demand from the seed (only reached when rejections exceed the slack the caller expanded; seed == null: zeros).
#if defined(TROYB_HOST_EMUL)
#define TB_NOINLINE
#else
#define TB_NOINLINE __noinline__
#endif
// 8 stream bytes at byte offset 8*w of the generator with this seed, computed from scratch (three compressions).
__device__ TB_NOINLINE inline u64 prng_word_on_demand(const u64 *seed, u64 w) {
    u64 sd[8], rt[8], out[8];
    for (int i = 0; i < 8; i++) sd[i] = seed[i];
    const u64 blk = w >> 3;
    Blake2::root(sd, blk >> 6, rt);
    Blake2::block(rt, (u32)(blk & 63), out);
    u64 r = 0;
    for (int i = 0; i < 8; i++) r = (w & 7) == (u64)i ? out[i] : r;
    return r;
}

struct PrngView {
    const u64 *buf;
    u64 bytes;
    const u64 *seed;
    TB_D u64 word(u64 w) const { // 8 stream bytes at byte offset 8*w, little endian
        if (w * 8 + 8 <= bytes) return buf[w];
        return seed ? prng_word_on_demand(seed, w) : 0;
    }
    TB_D u32 word32(u64 k) const { return (u32)(word(k >> 1) >> ((k & 1) * 32)); } // 4 bytes at byte offset 4*k
    TB_D u64 bytes6(u64 pos) const { // 6 bytes at any even byte offset, in the low 48 bits
        const u64 w = pos >> 3;
        const u32 sh = (u32)(pos & 7) * 8;
        u64 lo = word(w) >> sh;
        if (sh > 16) lo |= word(w + 1) << (64 - sh);
        return lo & 0xFFFFFFFFFFFFull;
    }
};

TB_HD int popc8(u32 x) {
#if defined(__CUDA_ARCH__)
    return __popc(x & 0xFFu);
#else
    return __builtin_popcount(x & 0xFFu);
#endif
}

// ---- stream expansion -------------------------------------------------------------------------------------------------------
// Blake2xbPRNG::refill_buffer  src/randomgen.cpp:184-194, for `buffers` consecutive buffers (from counter 0) of `batch`
// generators. Seeds: 8 words each, seed_stride words apart (they may be the first 64 bytes of another expanded stream).
struct PrngExpand {
    static constexpr int kThreads = 256;
    static constexpr u32 kBuffersPerCta = 16; // root hashes by 16 threads, then 4 rounds of one 64-byte block per thread
    const char *name() const { return "prng_blake2xb_expand"; }
    const u64 *seeds;
    u64 seed_stride;
    u64 *out;
    u64 out_stride; // words per generator
    u32 buffers;    // per generator
    u32 ctas_per_stream;
    TB_D void operator()(u32 bid, u64 *smem) const {
        const u32 b = bid / ctas_per_stream, c0 = (bid % ctas_per_stream) * kBuffersPerCta;
        TB_PHASE(kThreads, t, {
            if (t < kBuffersPerCta && c0 + t < buffers) {
                u64 sd[8], rt[8];
                for (int i = 0; i < 8; i++) sd[i] = seeds[b * seed_stride + i];
                Blake2::root(sd, c0 + t, rt);
                for (int i = 0; i < 8; i++) smem[t * 8 + i] = rt[i];
            }
        })
        TB_PHASE(kThreads, t, {
            for (u32 r = 0; r < kBuffersPerCta / 4; r++) {
                const u32 k = r * 4 + (t >> 6), c = c0 + k;
                if (c < buffers) {
                    u64 rt[8], o[8];
                    for (int i = 0; i < 8; i++) rt[i] = smem[k * 8 + i];
                    Blake2::block(rt, t & 63, o);
                    u64 *dst = out + b * out_stride + (u64)c * 512 + (t & 63) * 8;
                    for (int i = 0; i < 8; i++) dst[i] = o[i];
                }
            }
        })
    }
};

// ---- ternary sampler, pass 1 -----------------------------------------------------------------------------------------------
// samplePolyTernary  src/utils/rlwe.cpp:21-41: std::uniform_int_distribution<uint64_t>(0, 2) over 32-bit draws. With
// libstdc++ (GCC >= 11; the toolchain oracle/_ref is built with) that is Lemire's method: value = (3 * w) >> 32, and the
// draw w is REJECTED iff the low 32 bits of 3 * w are below 2^32 mod 3 = 1, i.e. iff w == 0. Coefficient i therefore takes
// the i-th NONZERO 32-bit word of the stream. Pass 1 counts the zero words of every 256-word segment of the first N words
// (seg[s] = zeros before segment s, seg[nseg] = total) and finds where the N-th nonzero word ends (seg[nseg + 1] = the
// stream position in bytes after the sampler, where the noise polynomials start).
constexpr u32 kSampleSeg = 256;

struct SampleTernaryScan {
    static constexpr int kThreads = 256;
    const char *name() const { return "sample_ternary_scan"; }
    PrngView view;      // of generator 0
    u64 view_stride;    // words between the generators' buffers
    u64 seed_stride;
    u64 start_bytes;    // stream position of the first draw (multiple of 4)
    u32 N, nseg;
    u64 *seg;           // [batch][nseg + 2]
    TB_D PrngView gen(u32 b) const {
        PrngView v = view;
        v.buf += b * view_stride;
        if (v.seed) v.seed += b * seed_stride;
        return v;
    }
    TB_D void operator()(u32 bid, u64 *) const {
        const PrngView v = gen(bid);
        u64 *sg = seg + (u64)bid * (nseg + 2);
        const u64 k0 = start_bytes >> 2;
        TB_PHASE(kThreads, t, {
            for (u32 s = t; s < nseg; s += kThreads) {
                u32 z = 0;
                const u32 len = N - s * kSampleSeg < kSampleSeg ? N - s * kSampleSeg : kSampleSeg;
                for (u32 k = 0; k < len; k++) z += v.word32(k0 + (u64)s * kSampleSeg + k) == 0 ? 1u : 0u;
                sg[s] = z;
            }
        })
        TB_PHASE(kThreads, t, {
            if (t == 0) {
                u64 run = 0;
                for (u32 s = 0; s < nseg; s++) { const u64 z = sg[s]; sg[s] = run; run += z; }
                sg[nseg] = run;
                u64 k = k0 + N;
                for (u64 need = run; need; k++) need -= v.word32(k) != 0 ? 1u : 0u; // one more nonzero word per rejected draw
                sg[nseg + 1] = k << 2;
            }
        })
    }
};

// ---- ternary + centred-binomial samplers, pass 2 ------------------------------------------------------------------------------
// samplePolyTernary (above) and samplePolyCbd  src/utils/rlwe.cpp:71-105 (6 bytes per coefficient, no rejection:
// popcount(x0) + popcount(x1) + popcount(x2 & 31) - popcount(x3) - popcount(x4) - popcount(x5 & 31)), written to every limb as
// the residue of the signed value. out: [batch][n_tern + n_cbd][n_limbs][N]. CTA (b, s) produces coefficients of segment s.
struct SampleExpand {
    static constexpr int kThreads = kSampleSeg;
    const char *name() const { return "sample_ternary_cbd"; }
    PrngView view;
    u64 view_stride, seed_stride;
    u64 start_bytes;   // as in pass 1
    u64 cbd_bytes;     // n_tern == 0: stream position of the first noise polynomial; else it comes from pass 1
    u32 N, nseg, n_tern, n_cbd, n_limbs;
    const u64 *seg;    // pass-1 output (n_tern == 1)
    const ModDev *mods;
    LimbMap map;
    u64 *out;
    u64 out_stride;
    TB_D void put(u64 *dst, u32 i, int value) const { // dst: [n_limbs][N]
        for (u32 l = 0; l < n_limbs; l++) dst[(u64)l * N + i] = value < 0 ? mods[map.idx[l]].p - (u64)(-value) : (u64)value;
    }
    TB_D void operator()(u32 bid, u64 *) const {
        const u32 b = bid / nseg, s = bid % nseg;
        PrngView v = view;
        v.buf += b * view_stride;
        if (v.seed) v.seed += b * seed_stride;
        const u64 *sg = seg ? seg + (u64)b * (nseg + 2) : nullptr;
        u64 *o = out + b * out_stride;
        const u64 k0 = start_bytes >> 2;
        TB_PHASE(kThreads, t, {
            const u32 i = s * kSampleSeg + t;
            if (i < N && n_tern) {
                const u64 zeros = sg[nseg];
                if (zeros == 0) {
                    const u32 w = v.word32(k0 + i);
                    put(o, i, (int)(((u64)w * 3) >> 32) - 1);
                } else if (t == 0) {
                    // a rejection somewhere in this polynomial (probability N * 2^-32): walk the segment in stream order
                    u64 z = sg[s];
                    for (u32 k = 0; k < kSampleSeg && s * kSampleSeg + k < N; k++) {
                        const u64 pos = (u64)s * kSampleSeg + k;
                        const u32 w = v.word32(k0 + pos);
                        if (w == 0) z++;
                        else put(o, (u32)(pos - z), (int)(((u64)w * 3) >> 32) - 1);
                    }
                    if (s == 0) { // the last `zeros` coefficients come from the words after the first N
                        u64 k = k0 + N;
                        for (u64 c = N - zeros; c < N; k++) {
                            const u32 w = v.word32(k);
                            if (w != 0) put(o, (u32)c++, (int)(((u64)w * 3) >> 32) - 1);
                        }
                    }
                }
            }
            const u64 base = n_tern ? sg[nseg + 1] : cbd_bytes;
            for (u32 j = 0; i < N && j < n_cbd; j++) {
                const u64 x = v.bytes6(base + 6ull * ((u64)j * N + i));
                const int noise = popc8((u32)x) + popc8((u32)(x >> 8)) + popc8((u32)(x >> 16) & 0x1F) - popc8((u32)(x >> 24)) -
                                  popc8((u32)(x >> 32)) - popc8((u32)(x >> 40) & 0x1F);
                put(o + (u64)(n_tern + j) * n_limbs * N, i, noise);
            }
        })
    }
};

// ---- uniform sampler -----------------------------------------------------------------------------------------------------------
// samplePolyUniform  src/utils/rlwe.cpp:123-158: the polynomial is first filled with n_limbs * N raw 64-bit words; then, limb by
// limb and coefficient by coefficient, every word >= max_multiple(q) = 2^64 - 1 - ((2^64 - 1) mod q) - 1 is redrawn from the
// words that FOLLOW in the stream until one is accepted, and reduced modulo q. Within one limb the criterion is the same for
// all its coefficients, so the rejected coefficients of a limb, in order, take the accepted words of the stream from the
// limb's cursor on, in order; only the cursor is carried from limb to limb.
TB_HD u64 uniform_max_multiple(const ModDev &m) { return ~0ull - barrett64(~0ull, m.p, m.ratio1) - 1; }

// pass 1: out = raw word mod q for every coefficient (accepted or not), grid-wide and coalesced.
struct SampleUniformFill {
    static constexpr int kThreads = 256;
    const char *name() const { return "sample_uniform_fill"; }
    PrngView view;
    u64 view_stride;
    u32 logN, n_limbs;
    const ModDev *mods;
    LimbMap map;
    u64 *out;
    u64 out_stride;
    u64 total; // batch * n_limbs * N
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            const u64 idx = (u64)bid * kThreads + t;
            if (idx < total) {
                const u64 per = (u64)n_limbs << logN;
                const u64 b = idx / per, r = idx % per;
                const ModDev &m = mods[map.idx[(u32)(r >> logN)]];
                out[b * out_stride + r] = barrett64(view.buf[b * view_stride + r], m.p, m.ratio1);
            }
        })
    }
};

// pass 2: one CTA per polynomial walks the limbs in order and replaces the rejected coefficients.
struct SampleUniformResolve {
    static constexpr int kThreads = 1024;
    const char *name() const { return "sample_uniform_resolve"; }
    PrngView view;
    u64 view_stride, seed_stride;
    u32 N, n_limbs;
    const ModDev *mods;
    LimbMap map;
    u64 *out;
    u64 out_stride;
    u64 *repl; // scratch [batch][N]: the accepted replacement words of the current limb, already reduced
    // shared memory (words): cnt[T] | off[T] | flg[T] | part[32] | ctl[4] = {rejected in limb, -, cursor, accepted in chunk}
    static constexpr u32 kSmemWords = 3 * kThreads + 32 + 4;
    // Exclusive prefix sums of in[0..kThreads) into out_pfx (a different array) and their total, in three phases.
    TB_D void scan(const u64 *in, u64 *out_pfx, u64 *part, u64 *total) const {
        TB_PHASE(kThreads, t, {
            if (t < 32) {
                u64 s = 0;
                for (u32 k = 0; k < 32; k++) s += in[t * 32 + k];
                part[t] = s;
            }
        })
        TB_PHASE(kThreads, t, {
            if (t == 0) {
                u64 run = 0;
                for (u32 k = 0; k < 32; k++) { const u64 s = part[k]; part[k] = run; run += s; }
                *total = run;
            }
        })
        TB_PHASE(kThreads, t, {
            u64 s = part[t >> 5];
            for (u32 k = t & ~31u; k < t; k++) s += in[k];
            out_pfx[t] = s;
        })
    }
    TB_D void operator()(u32 bid, u64 *smem) const {
        PrngView v = view;
        v.buf += bid * view_stride;
        if (v.seed) v.seed += bid * seed_stride;
        u64 *o = out + bid * out_stride;
        u64 *rp = repl + (u64)bid * N;
        u64 *cnt = smem, *off = smem + kThreads, *flg = smem + 2 * kThreads, *part = smem + 3 * kThreads, *ctl = part + 32;
        const u32 per = N >= (u32)kThreads ? N / kThreads : 1; // coefficients per thread (a contiguous range)
        TB_PHASE(kThreads, t, {
            if (t == 0) ctl[2] = (u64)n_limbs * N; // the first replacement word follows the raw fill
        })
        for (u32 l = 0; l < n_limbs; l++) {
            const ModDev &m = mods[map.idx[l]];
            const u64 maxm = uniform_max_multiple(m);
            const u64 raw0 = (u64)l * N;
            TB_PHASE(kThreads, t, {
                u64 c = 0;
                if ((u64)t * per < N)
                    for (u32 k = 0; k < per; k++) c += v.word(raw0 + (u64)t * per + k) >= maxm ? 1u : 0u;
                cnt[t] = c;
            })
            scan(cnt, off, part, &ctl[0]);
            // Values that steer the loops are read by every thread between two barriers and rewritten at least one barrier later.
            const u64 rejected = ctl[0];
            if (rejected == 0) continue;
            u64 got = 0, cursor = ctl[2];
            while (got < rejected) { // the next kThreads words of the stream: which are accepted, and as the how-manieth
                TB_PHASE(kThreads, t, {
                    flg[t] = v.word(cursor + t) < maxm ? 1u : 0u;
                })
                scan(flg, cnt, part, &ctl[3]); // prefixes into cnt (free once `off` holds the limb's ranks)
                TB_PHASE(kThreads, t, {
                    const u64 r = got + cnt[t];
                    if (flg[t] && r < rejected) {
                        rp[r] = barrett64(v.word(cursor + t), m.p, m.ratio1);
                        if (r + 1 == rejected) ctl[2] = cursor + t + 1; // the next limb's replacements start after this word
                    }
                })
                got += ctl[3];
                cursor += kThreads;
            }
            TB_PHASE(kThreads, t, {
                u64 r = off[t];
                if ((u64)t * per < N)
                    for (u32 k = 0; k < per; k++) {
                        const u64 i = (u64)t * per + k;
                        if (v.word(raw0 + i) >= maxm) o[raw0 + i] = rp[r++];
                    }
            })
        }
    }
};

// ---- seeded ciphertexts (the reference's transport form of fresh symmetric encryptions) --------------------------------------
// CiphertextCuda::load(stream, context)  src/ciphertext_cuda.cu:158-175 and encryptZeroSymmetric  src/utils/rlwe_cuda.cu:290-305:
// c1 is not transmitted; it is the expansion of a 64-bit seed by the reference's cuRAND sampler, coefficient index g owning the
// XORWOW generator curand_init(seed + g, 0, 0) and limb j taking its j-th curand_uniform_double draw times q_j, truncated
// (gInitCurandStates / gSamplePolyUniform, src/utils/rlwe_cuda.cu:25-28, 152-163). cuRAND's device generator is NVIDIA's header
// library; the arithmetic around it is restated. The result is in the "sampled" domain (NTT form; the caller inverts it for BFV
// as the reference's loader does).
struct SeededUniformExpand {
    static constexpr int kThreads = 256;
    const char *name() const { return "seeded_c1_expand"; }
    const u64 *seeds; // [batch]
    u32 N, n_limbs;
    const ModDev *mods;
    LimbMap map;
    u64 *out;
    u64 out_stride;
    u64 total; // batch * N
    TB_D void operator()(u32 bid, u64 *) const {
#if !defined(TROYB_HOST_EMUL)
        const u64 idx = (u64)bid * kThreads + threadIdx.x;
        if (idx < total) {
            const u64 b = idx / N, g = idx % N;
            curandState st;
            curand_init(seeds[b] + g, 0, 0, &st);
            for (u32 j = 0; j < n_limbs; j++)
                out[b * out_stride + (u64)j * N + g] = (u64)(curand_uniform_double(&st) * (double)mods[map.idx[j]].p);
        }
#else
        (void)bid;
#endif
    }
};

// The 64-bit seed the reference's symmetric encryptor takes from the ciphertext generator (the first two 32-bit draws with a
// nonzero combination, src/utils/rlwe_cuda.cu:292-297); one thread per generator.
struct SeedFromStream {
    static constexpr int kThreads = 32;
    const char *name() const { return "seed_from_stream"; }
    PrngView view;
    u64 view_stride, seed_stride;
    u32 batch;
    u64 *out; // [batch]
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            const u32 b = bid * kThreads + t;
            if (b < batch) {
                PrngView v = view;
                v.buf += b * view_stride;
                if (v.seed) v.seed += b * seed_stride;
                u64 seed = 0;
                for (u64 k = 0; seed == 0; k += 2) seed = ((u64)v.word32(k) << 32) + v.word32(k + 1);
                out[b] = seed;
            }
        })
    }
};

// ---- encrypt-zero arithmetic ------------------------------------------------------------------------------------------------
// The coefficient-wise steps of encryptZeroAsymmetric / encryptZeroSymmetric (src/utils/rlwe.cpp:199-242, 305-342): each is
// dyadicProductCoeffmod / multiplyPolyScalarCoeffmod / addPolyCoeffmod / negatePolyCoeffmod in one pass.
//   op 0: out = x * y                   (u * pk[j];  c1 * s)
//   op 1: out = x * y + z               (u * pk[j] + e_j, NTT form)
//   op 2: out = out + z * f             (+ e_j, or + t * e_j in BGV, coefficient form)
//   op 3: out = -(out + z * f)          (symmetric, coefficient form)
//   op 4: out = -(x * y + z * f)        (symmetric, NTT form)
// Elements are [batch][polys][n_limbs][N]; every operand has its own batch and poly strides (0 = shared).
struct EncCombine {
    static constexpr int kThreads = kEltThreads;
    const char *name() const { return "encrypt_zero_combine"; }
    int op, logN;
    u32 n_limbs, polys;
    u64 total; // batch * polys * n_limbs * N / 2 (two adjacent coefficients per thread, 128-bit accesses)
    const ModDev *mods;
    LimbMap map;
    u64 f;     // plain modulus (BGV) or 1
    const u64 *x, *y, *z;
    u64 *out;
    u64 x_batch, x_poly, y_batch, y_poly, z_batch, z_poly, out_batch, out_poly;
    static TB_D ulonglong2 ld2(const u64 *p) { return *reinterpret_cast<const ulonglong2 *>(p); }
    TB_D void operator()(u32 bid, u64 *) const {
        TB_PHASE(kThreads, t, {
            const u64 idx = ((u64)bid * kThreads + t) * 2;
            if (idx < total * 2) {
                const u32 ul = (u32)(idx >> logN), unit = ul / n_limbs;
                const u64 r = ((u64)(ul - unit * n_limbs) << logN) | (idx & ((u64(1) << logN) - 1));
                const u64 b = unit / polys, j = unit - (u32)b * polys;
                const ModDev &m = mods[map.idx[(u32)(r >> logN)]];
                const u64 p = m.p;
                u64 *po = out + b * out_batch + j * out_poly + r;
                ulonglong2 acc;
                if (op == 0 || op == 1 || op == 4) {
                    const ulonglong2 xv = ld2(x + b * x_batch + j * x_poly + r), yv = ld2(y + b * y_batch + j * y_poly + r);
                    acc.x = mulmod_barrett(xv.x, yv.x, p, m.ratio0, m.ratio1);
                    acc.y = mulmod_barrett(xv.y, yv.y, p, m.ratio0, m.ratio1);
                } else {
                    acc = ld2(po);
                }
                if (op != 0) {
                    ulonglong2 e = ld2(z + b * z_batch + j * z_poly + r);
                    if (f != 1) {
                        const u64 fm = barrett64(f, p, m.ratio1);
                        e.x = mulmod_barrett(e.x, fm, p, m.ratio0, m.ratio1);
                        e.y = mulmod_barrett(e.y, fm, p, m.ratio0, m.ratio1);
                    }
                    acc.x = addmod(acc.x, e.x, p);
                    acc.y = addmod(acc.y, e.y, p);
                }
                if (op == 3 || op == 4) { acc.x = negmod(acc.x, p); acc.y = negmod(acc.y, p); }
                *reinterpret_cast<ulonglong2 *>(po) = acc;
            }
        })
    }
};

} // namespace troyb
