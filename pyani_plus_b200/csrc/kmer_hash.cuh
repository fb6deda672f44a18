// kmer_hash.cuh -- per-thread core of the sketch kernel (K1): canonical k-mers of a 2-bit packed
// base stream -> MurmurHash3_x64_128(seed).h1 over the ASCII bytes of the canonical k-mer, exactly
// as `sourmash scripts singlesketch` computes it (reference call site pyani_plus/methods/sourmash.py:67-83;
// conventions SURVEY.md 8c; Appendix B for the 31-byte specialisation).
//
// The code is __host__ __device__ so that tests can run the identical logic on the CPU
// (csrc/hostemu.cpp) and compare it with the oracle before any GPU time is spent.
//
// Thread geometry ("stride-4" scheme).  A CTA tile is 256*KPT k-mer start positions handled by 256
// threads (KPT = kKmersPerThread = 16 in the text below).  Thread (u = tid>>2, a = tid&3) owns the
// KPT k-mers starting at tile positions 4*KPT*u + a + 4j, j = 0..KPT-1.  Because the starts are 4 bases = 4 ASCII bytes apart, the ASCII form of the thread's
// SPAN = 60+K bases, expanded ONCE into NA 32-bit registers, contains every one of the 16 k-mers as
// a register-aligned window G[j .. j+NWD): no per-k-mer byte shifting is needed.  The same holds for
// the reverse strand: the reverse complement of the whole span, expanded once into H[], contains the
// reverse complement of k-mer j as H[15-j .. 15-j+NWD) (SPAN-K = 60 is a multiple of 4).
// The canonical choice is made on the packed 2-bit form: with bases packed LSB-first (base i at
// bits 2i), fwd < revcomp lexicographically  <=>  F_lsb < R_lsb as integers, where F_lsb / R_lsb are
// the 2K-bit windows of the packed span and of its packed reverse complement (the MSB-first code
// of one strand is the bitwise complement of the LSB-first code of the other).
#pragma once
#include <stdint.h>

#include <type_traits>

#if defined(__CUDACC__)
#define PANIB_HD __host__ __device__ __forceinline__
#else
#define PANIB_HD inline
#endif

namespace panib {

#ifndef PANIB_KPT
#define PANIB_KPT 16
#endif
constexpr int kTileBases = 4096;     // == PANIB_TILE_BASES: alignment unit of genomes in the base stream
#ifndef PANIB_K1_THREADS
#define PANIB_K1_THREADS 256
#endif
constexpr int kThreadsK1 = PANIB_K1_THREADS;
constexpr int kKmersPerThread = PANIB_KPT;                  // k-mers per thread (multiple of 4)
constexpr int kCtaTile = kThreadsK1 * kKmersPerThread;      // k-mer starts per CTA pass (divides kTileBases)
constexpr int kTileWords = kCtaTile / 16 + 8;               // packed words staged per CTA tile (tile + halo)
constexpr int kTileMaskWords = kCtaTile / 32 + 4;
static_assert(kKmersPerThread % 4 == 0 && kTileBases % kCtaTile == 0, "tile geometry");

// ---- small intrinsics with host emulation ---------------------------------------------------
PANIB_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, s);
#else
    uint64_t pool = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(pool >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0;
        r |= byte << (8 * i);
    }
    return r;
#endif
}
// funnel shift right: low 32 bits of ((hi:lo) >> s), 0 <= s <= 31
PANIB_HD uint32_t shf_r(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}
PANIB_HD uint32_t brev32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __brev(v);
#else
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
#endif
}
PANIB_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
PANIB_HD uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

// ---- 64-bit arithmetic on explicit 32-bit halves ------------------------------------------------
// nvcc lowers `x * C` (64-bit, constant C) to 4 IMADs and turns `rotl(x * C, r)` into extra
// multiplies by C << r; spelled out on halves with mad.lo / funnel shifts every 64-bit multiply is
// exactly 3 IMAD-class instructions and every rotate 2 SHF (measured with ncu: 171 -> ~120
// instructions per k-mer).  The host branch states the same arithmetic in plain C.
struct U64 {
    uint32_t lo, hi;
};

PANIB_HD U64 make_u64(uint32_t lo, uint32_t hi) { return U64{lo, hi}; }
PANIB_HD uint64_t to_u64(U64 x) { return ((uint64_t)x.hi << 32) | x.lo; }

template <uint64_t C>
PANIB_HD U64 mul_const(U64 a) {
    constexpr uint32_t clo = (uint32_t)C, chi = (uint32_t)(C >> 32);
#if defined(__CUDA_ARCH__)
    uint32_t plo, phi, t, rhi;
    asm("{\n\t.reg .b64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}"
        : "=r"(plo), "=r"(phi) : "r"(a.lo), "r"(clo));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(a.hi), "r"(clo), "r"(phi));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(rhi) : "r"(a.lo), "r"(chi), "r"(t));
    return U64{plo, rhi};
#else
    const uint64_t p = (uint64_t)a.lo * clo;
    const uint32_t t = a.hi * clo + (uint32_t)(p >> 32);
    return U64{(uint32_t)p, a.lo * chi + t};
#endif
}

// x * 5 + c (c < 2^32)
template <uint32_t ADD>
PANIB_HD U64 mul5_add(U64 a) {
#if defined(__CUDA_ARCH__)
    uint32_t plo, phi, rhi;
    asm("{\n\t.reg .b64 p;\n\tmad.wide.u32 p, %2, 5, %3;\n\tmov.b64 {%0, %1}, p;\n\t}"
        : "=r"(plo), "=r"(phi) : "r"(a.lo), "l"((uint64_t)ADD));
    asm("mad.lo.u32 %0, %1, 5, %2;" : "=r"(rhi) : "r"(a.hi), "r"(phi));
    return U64{plo, rhi};
#else
    const uint64_t p = (uint64_t)a.lo * 5u + ADD;
    return U64{(uint32_t)p, a.hi * 5u + (uint32_t)(p >> 32)};
#endif
}

PANIB_HD uint32_t shf_l(uint32_t lo, uint32_t hi, uint32_t s) {  // high 32 bits of ((hi:lo) << s), 0 <= s <= 31
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}

template <int R>
PANIB_HD U64 rotl(U64 x) {
    static_assert(R > 0 && R < 64 && R != 32, "rotation amount");
    if constexpr (R < 32) {
        return U64{shf_l(x.hi, x.lo, R), shf_l(x.lo, x.hi, R)};
    } else {  // swap halves, then rotate by R-32
        return U64{shf_l(x.lo, x.hi, R - 32), shf_l(x.hi, x.lo, R - 32)};
    }
}

// (Measured and rejected: doing the 64-bit adds and the `>> 33` of fmix on the FMA pipe through
// IMAD.WIDE with multipliers ptxas cannot fold -- 5-15 % SLOWER; IMAD.WIDE/IMAD.HI are not cheap.)
PANIB_HD U64 add64(U64 a, U64 b) {
#if defined(__CUDA_ARCH__)
    U64 r;  // spelled as a carry chain: nvcc's own lowering of chained 64-bit adds wastes an IADD3
    asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;" : "=r"(r.lo), "=r"(r.hi) : "r"(a.lo), "r"(a.hi), "r"(b.lo), "r"(b.hi));
    return r;
#else
    const uint64_t r = to_u64(a) + to_u64(b);
    return U64{(uint32_t)r, (uint32_t)(r >> 32)};
#endif
}
PANIB_HD U64 xor64(U64 a, U64 b) { return U64{a.lo ^ b.lo, a.hi ^ b.hi}; }
PANIB_HD U64 xorshift33(U64 x) { return U64{x.lo ^ (x.hi >> 1), x.hi}; }  // x ^= x >> 33
PANIB_HD U64 fmix(U64 k) {
    k = xorshift33(k);
    k = mul_const<0xff51afd7ed558ccdULL>(k);
    k = xorshift33(k);
    k = mul_const<0xc4ceb9fe1a85ec53ULL>(k);
    return xorshift33(k);
}
PANIB_HD U64 mix_k1(U64 k1) {  // k1 *= c1; k1 = rotl(k1, 31); k1 *= c2
    return mul_const<0x4cf5ad432745937fULL>(rotl<31>(mul_const<0x87c37b91114253d5ULL>(k1)));
}
PANIB_HD U64 mix_k2(U64 k2) {  // k2 *= c2; k2 = rotl(k2, 33); k2 *= c1
    return mul_const<0x87c37b91114253d5ULL>(rotl<33>(mul_const<0x4cf5ad432745937fULL>(k2)));
}

// ---- MurmurHash3_x64_128 h1 of a K-byte key held as little-endian 32-bit words W[0..ceil(K/4)) ----
// Bytes of W beyond K are ignored (masked here), so W may be a window of a longer byte string.
//
// The hash is returned UNFINISHED: h = xorshift33(m1) + xorshift33(m2) (the last step of both fmix64
// and the final h1 += h2).  Only ~1/scaled of the k-mers survive `h <= max_hash`, so the hot loop
// tests a one-sided filter on the high words (Partial::prefilter) and finishes the hash on the rare
// path only: 2 instructions instead of 8 per k-mer.
struct HashConsts {
    uint32_t seed;
    // Both "h = (rotl(h) + other) * 5 + c" steps are computed as 5 * (rotl(h) + other + c/5), with c/5 taken
    // modulo 2^64 (5 is odd): one 64-bit 3-input add and one shift-add (LEA / LEA.HI.X), all on the ALU
    // pipe, instead of a wide multiply-add chain on the FMA-heavy pipe that the 12 real multiplies saturate.
    U64 a1;        // seed + 0x52dce729 / 5: block 0 adds h2 == seed, folded into the constant
    U64 a2;        // 0x38495ab5 / 5
    uint32_t thr;  // min(hi32(max_hash) + 1, 2^32 - 1)
    uint64_t max_hash;
};
constexpr uint64_t kInv5 = 0xCCCCCCCCCCCCCCCDull;  // 5 * kInv5 == 1 (mod 2^64)
static_assert(5ull * kInv5 == 1ull, "inverse of 5");
PANIB_HD HashConsts make_hash_consts(uint32_t seed, uint64_t max_hash) {
    const uint64_t a1 = (uint64_t)seed + 0x52dce729ull * kInv5;
    const uint64_t a2 = 0x38495ab5ull * kInv5;
    const uint32_t mh = (uint32_t)(max_hash >> 32);
    return HashConsts{seed, U64{(uint32_t)a1, (uint32_t)(a1 >> 32)}, U64{(uint32_t)a2, (uint32_t)(a2 >> 32)},
                      mh == 0xFFFFFFFFu ? mh : mh + 1u, max_hash};
}

struct Partial {
    U64 m1, m2;
    // Necessary condition for hash() <= max_hash.  hi32(hash) = m1.hi + m2.hi + carry, carry in {0,1},
    // so hi32(hash) <= hi32(max_hash) implies (m1.hi + m2.hi + 1 mod 2^32) <= hi32(max_hash) + 1 whatever
    // the carry is (with no carry the sum itself is <= hi32(max_hash) and the +1 cannot wrap).
    PANIB_HD bool prefilter(const HashConsts &hc) const { return m1.hi + m2.hi + 1u <= hc.thr; }
    PANIB_HD uint64_t hash() const { return to_u64(add64(xorshift33(m1), xorshift33(m2))); }
};

// 5 * (a + b + c): the 64-bit form lets nvcc emit IADD3 / IADD3.X and LEA / LEA.HI.X
PANIB_HD U64 times5_sum(U64 a, U64 b, U64 c) {
    const uint64_t t = to_u64(a) + to_u64(b) + to_u64(c);
    const uint64_t r = (t << 2) + t;
    return U64{(uint32_t)r, (uint32_t)(r >> 32)};
}
PANIB_HD U64 times5_sum(U64 a, U64 c) {
    const uint64_t t = to_u64(a) + to_u64(c);
    const uint64_t r = (t << 2) + t;
    return U64{(uint32_t)r, (uint32_t)(r >> 32)};
}
PANIB_HD U64 fmix_head(U64 k) {  // fmix64 without its last xorshift
    k = xorshift33(k);
    k = mul_const<0xff51afd7ed558ccdULL>(k);
    k = xorshift33(k);
    return mul_const<0xc4ceb9fe1a85ec53ULL>(k);
}

// S42 = the seed is sourmash's 42 (the only one the reference path ever uses): the seed-dependent constants are
// then immediates of the adds instead of uniform registers that ptxas re-loads from the constant bank for every
// k-mer (one LDCU.128 per k-mer less; measured +1.2 % at scaled=1000, +1.0 % at scaled=100).  hc.seed must be 42.
constexpr uint32_t kSourmashSeed = 42u;
template <int K, bool S42 = false>
PANIB_HD Partial murmur_words(const uint32_t *W, const HashConsts &hc) {
    constexpr uint64_t kA1 = (uint64_t)kSourmashSeed + 0x52dce729ull * kInv5, kA2 = 0x38495ab5ull * kInv5;
    const U64 a1 = S42 ? U64{(uint32_t)kA1, (uint32_t)(kA1 >> 32)} : hc.a1;
    const U64 a2 = S42 ? U64{(uint32_t)kA2, (uint32_t)(kA2 >> 32)} : hc.a2;
    const uint32_t seed = S42 ? kSourmashSeed : hc.seed;
    U64 h1{seed, 0u}, h2{seed, 0u};
    constexpr int nblocks = K / 16;
#pragma unroll
    for (int i = 0; i < nblocks; i++) {
        h1 = xor64(h1, mix_k1(U64{W[4 * i], W[4 * i + 1]}));
        if (i == 0) h1 = times5_sum(rotl<27>(h1), a1);  // h2 is still the seed
        else h1 = mul5_add<0x52dce729u>(add64(rotl<27>(h1), h2));
        h2 = xor64(h2, mix_k2(U64{W[4 * i + 2], W[4 * i + 3]}));
        h2 = times5_sum(rotl<31>(h2), h1, a2);
    }
    constexpr int tail = K & 15;
    constexpr int tb = 4 * nblocks;  // first tail word
    if (tail > 8) {
        constexpr int nb2 = tail - 8;  // 1..7 bytes in k2
        uint32_t lo = W[tb + 2];
        uint32_t hi = nb2 > 4 ? W[tb + 3] : 0u;
        if (nb2 < 4) lo &= (1u << (8 * (nb2 & 3))) - 1u;
        if (nb2 > 4 && nb2 < 8) hi &= (1u << (8 * (nb2 & 3))) - 1u;
        h2 = xor64(h2, mix_k2(U64{lo, hi}));
    }
    if (tail > 0) {
        constexpr int nb1 = tail > 8 ? 8 : tail;  // 1..8 bytes in k1
        uint32_t lo = W[tb];
        uint32_t hi = nb1 > 4 ? W[tb + 1] : 0u;
        if (nb1 < 4) lo &= (1u << (8 * (nb1 & 3))) - 1u;
        if (nb1 > 4 && nb1 < 8) hi &= (1u << (8 * (nb1 & 3))) - 1u;
        h1 = xor64(h1, mix_k1(U64{lo, hi}));
    }
    h1.lo ^= (uint32_t)K;
    h2.lo ^= (uint32_t)K;
    h1 = add64(h1, h2);
    h2 = add64(h2, h1);
    return Partial{fmix_head(h1), fmix_head(h2)};
}

// ---- geometry of one thread's span for k-mer size K ------------------------------------------
template <int K>
struct Geom {
    static_assert(K >= 1 && K <= 32, "register-resident kernel handles 1 <= K <= 32");
    static_assert(kKmersPerThread == 16, "the tile expansion below assumes 16 k-mers (64 bases) per thread");
    static constexpr int SPAN = 4 * (kKmersPerThread - 1) + K;  // bases touched by one thread
    static constexpr int NX = (2 * SPAN + 6 + 31) / 32;         // packed words (incl. 0..6 bits of alignment)
    static constexpr int NA = (SPAN + 3) / 4;                   // ASCII words per strand
    static constexpr int NWD = (K + 3) / 4;                     // ASCII words per k-mer
    static constexpr int NU = kThreadsK1 / 4;                   // 64-base blocks per CTA tile
    static constexpr int NI = kCtaTile / 16;                    // packed words per CTA tile
    static constexpr int ITEMS = NI + 2;                        // ... expanded per tile (incl. the K-1 halo)
    // reverse strand: the tile (kCtaTile + K bases) is read back to front as the stream
    //   RCm[x] = comp(T[kCtaTile + K - 2 - x]),
    // in which the reverse complement of the k-mer at tile position p = 64u + a + 4j starts at
    // 4 * (kCtaTile/4 - 1 - 16u - j) + (3 - a): word-aligned in the copy shifted by 3 - a bytes, and
    // 4-byte steps apart for consecutive j, exactly like the forward strand.  Packed word v of RCm is the
    // reverse complement of the 16 tile bases starting at 16 * (NI - v) + K - 17.
    static constexpr int RC_D = K - 17;  // first base of the 16-base window of item v: 16 * (NI - v) + RC_D
    static constexpr int RC_DQ = RC_D >= 0 ? RC_D / 16 : -((15 - RC_D) / 16);  // floor(RC_D / 16)
    static constexpr int RC_DR = RC_D - 16 * RC_DQ;                            // 0..15
};

// 16 bases (one packed word) -> 4 ASCII words.  LUT byte c = ASCII of code c.
PANIB_HD void expand16(uint32_t x, uint32_t *out4) {
    const uint32_t lut = 0x54474341u;  // 'A','C','G','T' for codes 0..3
    // spread each byte (4 bases) to 4 nibbles: PRMT puts bytes b0,b1 (b2,b3) into even byte lanes,
    // two shift-or-mask steps finish the 2-bit -> 4-bit spread; the nibbles are PRMT selectors.
    uint32_t lo = prmt(x, 0u, 0x4140u);  // [b0, 0, b1, 0]
    uint32_t hi = prmt(x, 0u, 0x4342u);  // [b2, 0, b3, 0]
    lo = (lo | (lo << 4)) & 0x0F0F0F0Fu;
    hi = (hi | (hi << 4)) & 0x0F0F0F0Fu;
    lo = (lo | (lo << 2)) & 0x33333333u;
    hi = (hi | (hi << 2)) & 0x33333333u;
    out4[0] = prmt(lut, 0u, lo);
    out4[1] = prmt(lut, 0u, lo >> 16);
    out4[2] = prmt(lut, 0u, hi);
    out4[3] = prmt(lut, 0u, hi >> 16);
}
// 4 bases (the low 8 bits of x; the rest must be zero) -> 1 ASCII word
PANIB_HD uint32_t expand4(uint32_t x) {
    x = (x | (x << 4)) & 0x0F0Fu;
    x = (x | (x << 2)) & 0x3333u;
    return prmt(0x54474341u, 0u, x);
}

// reverse the order of the 16 2-bit fields of a word and complement them
PANIB_HD uint32_t revcomp16(uint32_t x) {
    uint32_t r = brev32(x);
    return ~(((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1));
}

// 2K-bit window of a packed multiword value starting at (compile-time) bit offset `off`, as u64.
template <int K, int NXW>
PANIB_HD uint64_t window(const uint32_t *X, int off) {
    const int w = off >> 5, s = off & 31;
    uint32_t lo, hi;
    if (s == 0) {
        lo = X[w];
        hi = (2 * K > 32) ? X[w + 1] : 0u;
    } else {
        lo = shf_r(X[w], (w + 1 < NXW) ? X[w + 1] : 0u, s);
        hi = (2 * K > 32) ? shf_r((w + 1 < NXW) ? X[w + 1] : 0u, (w + 2 < NXW) ? X[w + 2] : 0u, s) : 0u;
    }
    if (2 * K < 32) lo &= (1u << ((2 * K) & 31)) - 1u;
    if (2 * K > 32 && 2 * K < 64) hi &= (1u << ((2 * K) & 31)) - 1u;
    return ((uint64_t)hi << 32) | lo;
}

// ================================================================================================
// K1 works on a CTA tile in two phases.
//
// Phase A (tile_expand_item, cooperative): the ASCII form of both strands of the tile is produced ONCE
// per CTA -- item t = packed word t of the tile (16 bases) is expanded by one thread, shifted by 0..3
// bytes with PRMT, and stored straight into the private scratch blocks of the threads that will hash
// it -- instead of every thread expanding its own 91-base span (each base 4x per strand).
//
// Phase B (hash_thread_kmers): thread (u, a) hashes its 16 k-mers.  Its scratch block holds the ASCII
// words of its span for both strands, so k-mer j reads its NWD words from
// forward or its reverse block (one address select; see the layout below): the canonical choice costs one select and the
// words arrive through the otherwise idle load/store pipe.
//
// Scratch layout: per strand and thread kBlkPos = 24 word positions, stored as 12 PAIRS; pair r of thread
// tid lives at scratch[((strand * 12 + r) * nthreads + tid) * 2 .. +2).  A warp's 64-bit access to one pair
// is 256 consecutive bytes whatever strand each lane picks (two wavefronts, the minimum), phase A stores
// pairs with one STS.64 and phase B loads a k-mer's 8 words with 4 (even j) or 5 (odd j) loads instead of 8.
//   forward position p of thread (u, a) = tile bytes [64u + a + 4p, +4)       = word 16u + p of the copy shifted by a
//   reverse position p of thread (u, a) = RCm bytes [64(NU-1-u) + (3-a) + 4(p-1), +4)
//                                       = word 16(NU-1-u) + p of the copy shifted by 3 - a of the stream
//                                         RCn = RCm delayed by one word (4 bases)
// The one-word delay makes k-mer j start at position j on the forward and 16 - j on the reverse strand: the
// same parity, so both strands of a k-mer are loaded with the same instruction shapes.  Word w of a shifted
// copy goes to block w >> 4 at position w & 15 and, because spans overlap, also to the neighbouring block
// at position (w & 15) + 16 when w & 15 < 8.  (Reverse position 0 and forward position 23 are never read.)
// ================================================================================================
constexpr int kBlkPos = 24;    // word positions per strand of a span (ceil((60 + 32) / 4) = 23 used, + the delay)
constexpr int kBlkPairs = kBlkPos / 2;
constexpr int kSpLead = 4;     // readable words in front of the staged packed tile (content irrelevant)

// index of word position p of `strand` in thread tid's scratch
PANIB_HD int scr_index(int strand, int p, int tid, int nthreads) {
    return ((strand * kBlkPairs + (p >> 1)) * nthreads + tid) * 2 + (p & 1);
}
PANIB_HD void load_pair(const uint32_t *at, uint32_t &v0, uint32_t &v1) {
#if defined(__CUDA_ARCH__)
    const uint2 v = *reinterpret_cast<const uint2 *>(at);
    v0 = v.x;
    v1 = v.y;
#else
    v0 = at[0];
    v1 = at[1];
#endif
}
PANIB_HD void store_pair(uint32_t *at, uint32_t v0, uint32_t v1) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint2 *>(at) = make_uint2(v0, v1);
#else
    at[0] = v0;
    at[1] = v1;
#endif
}

// packed word v of the reverse stream RCm (reverse complement of the 16 tile bases from 16(NI - v) + K - 17)
// and, in g8, the 4 tile bases BEHIND them (their reverse complement is the last 4 bases of word v - 1)
template <int K>
PANIB_HD uint32_t rc_packed_word(const uint32_t *sp, int v, uint32_t &g8) {
    using G_ = Geom<K>;
    const uint32_t *s = sp + (G_::NI - v + G_::RC_DQ);
    constexpr int sh = 2 * G_::RC_DR;  // bit offset of the 16-base window in s[0]
    uint32_t g32;
    if constexpr (sh == 0) {
        g32 = s[0];
        g8 = s[1] & 0xFFu;
    } else {
        g32 = shf_r(s[0], s[1], sh);
        if constexpr (sh + 8 <= 32) g8 = (s[1] >> sh) & 0xFFu;
        else g8 = shf_r(s[1], s[2], sh) & 0xFFu;
    }
    return revcomp16(g32);
}

// Phase A for item t (0 <= t < Geom<K>::NI; the halo items are tile_expand_halo's).  sp[-1] .. sp[NI + 2] must
// be readable; rcp receives packed word t of the reverse stream (phase B compares on it).
// In round r the lane stores the copy shifted by (t + r) & 3 bytes: the four lanes that share a block then
// write four different threads' slots, which keeps a warp's store on 32 different 8-byte columns.
template <int K>
PANIB_HD void tile_expand_item(const uint32_t *sp, uint32_t *rcp, uint32_t *scratch, int nthreads, int t) {
    using G_ = Geom<K>;
    static_assert(G_::NA <= kBlkPos - 1, "scratch block too small");
    const int up = t >> 2, m = t & 3;
    uint32_t sel[4];
    int ar[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        ar[r] = (m + r) & 3;
        sel[r] = 0x3210u + 0x1111u * (uint32_t)ar[r];
    }
    const bool sec = up >= 1 && m < 2;  // positions 4m .. 4m+3 also feed the neighbouring block at +16
    uint32_t E[5];
    // ---- forward strand: tile bases [16t, 16t + 20) -> words 4t .. 4t+3 of the four shifted copies
    expand16(sp[t], E);
    E[4] = expand4(sp[t + 1] & 0xFFu);
    {
        uint32_t *p0 = scratch + scr_index(0, 4 * m, 4 * up, nthreads);
#pragma unroll
        for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t v0 = prmt(E[2 * h], E[2 * h + 1], sel[r]);
                const uint32_t v1 = prmt(E[2 * h + 1], E[2 * h + 2], sel[r]);
                uint32_t *p = p0 + h * (2 * nthreads) + 2 * ar[r];
                store_pair(p, v0, v1);
                if (sec) store_pair(p + 8 * (2 * nthreads) - 8, v0, v1);
            }
        }
    }
    // ---- reverse strand: words 4t .. 4t+3 of the copies of RCn = RCm words 4t-1 .. 4t+2: the last word of
    //      RCm packed word t-1 (reverse complement of the 4 bases behind the window) and packed word t itself
    {
        uint32_t g8;
        const uint32_t rcw = rc_packed_word<K>(sp, t, g8);
        rcp[t] = rcw;
        E[0] = expand4(revcomp16(g8 << 24) & 0xFFu);
        expand16(rcw, E + 1);
        uint32_t *p0 = scratch + scr_index(1, 4 * m, 4 * (G_::NU - 1 - up) + 3, nthreads);
#pragma unroll
        for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t v0 = prmt(E[2 * h], E[2 * h + 1], sel[r]);
                const uint32_t v1 = prmt(E[2 * h + 1], E[2 * h + 2], sel[r]);
                uint32_t *p = p0 + h * (2 * nthreads) - 2 * ar[r];
                store_pair(p, v0, v1);
                if (sec) store_pair(p + 8 * (2 * nthreads) + 8, v0, v1);
            }
        }
    }
}

// Phase A, halo: the K-1 bases behind the tile only feed positions 16.. of the last forward block (u = NU-1)
// and of the last reverse block (u = 0).  ONE warp does this, one (strand, word, pair of shifts) per lane, so
// that no warp is a whole item behind the others at the barrier.  The forward lanes of word 7 (position 23,
// never read) store the two packed reverse words the last block compares instead.
template <int K>
PANIB_HD void tile_expand_halo(const uint32_t *sp, uint32_t *rcp, uint32_t *scratch, int nthreads, int lane) {
    using G_ = Geom<K>;
    const int strand = lane >> 4, hw = (lane >> 1) & 7, ap = lane & 1;
    if (strand == 0 && hw == 7) {
        uint32_t g8;
        rcp[G_::NI + ap] = rc_packed_word<K>(sp, G_::NI + ap, g8);
        return;
    }
    uint32_t g16;
    if (strand == 0) {  // forward word 4NI + hw: tile bases [kCtaTile + 4hw, +8)
        const uint32_t *s = sp + G_::NI + (hw >> 2);
        g16 = shf_r(s[0], s[1], 8u * (uint32_t)(hw & 3)) & 0xFFFFu;
    } else {  // RCn word 4NI + hw = RCm bases [kCtaTile - 4 + 4hw, +8) = reverse complement of tile bases [K-5-4hw, +8)
        const int tb = K - 5 - 4 * hw;
        const uint32_t *s = sp + (tb >> 4);  // arithmetic shift: the lead-in words cover tb < 0
        g16 = revcomp16(shf_r(s[0], s[1], 2u * (uint32_t)(tb & 15)) << 16) & 0xFFFFu;
    }
    const uint32_t e0 = expand4(g16 & 0xFFu), e1 = expand4(g16 >> 8);
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int sft = 2 * ap + q;  // copy shifted by sft bytes
        const uint32_t v = prmt(e0, e1, 0x3210u + 0x1111u * (uint32_t)sft);
        if (strand == 0) scratch[scr_index(0, 16 + hw, 4 * (G_::NU - 1) + sft, nthreads)] = v;
        else scratch[scr_index(1, 16 + hw, 3 - sft, nthreads)] = v;
    }
}

// validity of the 16 k-mers of thread (u, a): bit j set = no invalid base in k-mer j.  sm = validity
// mask words of the tile (bit set = invalid).  Only called for tiles that hold an invalid base.
template <int K>
PANIB_HD uint32_t thread_valid_mask(const uint32_t *sm, int u, int a) {
    uint32_t vm = 0;
#pragma unroll 1
    for (int j = 0; j < kKmersPerThread; j++) {
        const int pos = 4 * kKmersPerThread * u + a + 4 * j;
        uint32_t mw = shf_r(sm[pos >> 5], sm[(pos >> 5) + 1], pos & 31);
        if (K < 32) mw &= (1u << (K & 31)) - 1u;
        vm |= (mw == 0u ? 1u : 0u) << j;
    }
    return vm;
}

// ---- phase B loads: the thread's scratch through a 32-bit shared-memory address and explicit ld.shared with
// compile-time offsets (left to itself nvcc re-derives the shared window base for every k-mer); plain
// pointers in the host emulation
#if defined(__CUDA_ARCH__)
using ScrBase = uint32_t;
PANIB_HD ScrBase scr_base(const uint32_t *blk) { return (uint32_t)__cvta_generic_to_shared(blk); }
PANIB_HD ScrBase scr_advance(ScrBase b, int bytes) { return b + (uint32_t)bytes; }
template <int OFF>
PANIB_HD void scr_load2(ScrBase b, uint32_t &v0, uint32_t &v1) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v0), "=r"(v1) : "r"(b), "n"(OFF));
}
template <int OFF>
PANIB_HD uint32_t scr_load1(ScrBase b) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(b), "n"(OFF));
    return v;
}
#else
using ScrBase = const uint8_t *;
PANIB_HD ScrBase scr_base(const uint32_t *blk) { return reinterpret_cast<const uint8_t *>(blk); }
PANIB_HD ScrBase scr_advance(ScrBase b, int bytes) { return b + bytes; }
template <int OFF>
PANIB_HD void scr_load2(ScrBase b, uint32_t &v0, uint32_t &v1) {
    const uint32_t *p = reinterpret_cast<const uint32_t *>(b + OFF);
    v0 = p[0];
    v1 = p[1];
}
template <int OFF>
PANIB_HD uint32_t scr_load1(ScrBase b) { return *reinterpret_cast<const uint32_t *>(b + OFF); }
#endif

// 64 bits of a packed multiword value starting at (compile-time) bit offset `off`, no masking
template <int NXW>
PANIB_HD void window64(const uint32_t *X, int off, uint32_t &lo, uint32_t &hi) {
    const int w = off >> 5, s = off & 31;
    if (s == 0) {
        lo = X[w];
        hi = (w + 1 < NXW) ? X[w + 1] : 0u;
    } else {
        lo = shf_r(X[w], (w + 1 < NXW) ? X[w + 1] : 0u, s);
        hi = shf_r((w + 1 < NXW) ? X[w + 1] : 0u, (w + 2 < NXW) ? X[w + 2] : 0u, s);
    }
}

#ifndef PANIB_K1_GROUP
#define PANIB_K1_GROUP 1
#endif
// k-mer J of a thread (compile-time J: every scratch offset is an immediate); recursion = the unrolled loop
template <int K, int J, class Emit, bool S42 = false>
struct KmerStep {
    using G_ = Geom<K>;
    static constexpr int NX = G_::NX, NWD = G_::NWD;
    static constexpr bool WHOLE = (K == 31 || K == 32);
    static constexpr int ROW = 2 * kThreadsK1 * 4;  // bytes between pair rows

    template <int Q>
    static PANIB_HD void pairs_even(ScrBase b, uint32_t *W) {
        if constexpr (Q < (NWD + 1) / 2) {
            scr_load2<((J >> 1) + Q) * ROW>(b, W[2 * Q], W[2 * Q + 1]);
            pairs_even<Q + 1>(b, W);
        }
    }
    template <int Q>
    static PANIB_HD void pairs_odd(ScrBase b, uint32_t *W) {
        if constexpr (Q < NWD / 2) {
            scr_load2<((J >> 1) + 1 + Q) * ROW>(b, W[2 * Q + 1], W[2 * Q + 2]);
            pairs_odd<Q + 1>(b, W);
        }
    }

    // canonical choice + loads + hash of k-mer J (no side effects: callers can interleave several)
    static PANIB_HD Partial hash_one(const uint32_t *X, const uint32_t *Xr, ScrBase fw, const HashConsts &hc) {
        bool lt;
        if constexpr (WHOLE) {
            uint32_t flo, fhi, rlo, rhi;
            window64<NX>(X, 8 * J, flo, fhi);
            window64<NX>(Xr, 8 * (kKmersPerThread - 1 - J), rlo, rhi);
            lt = (((uint64_t)fhi << 32) | flo) < (((uint64_t)rhi << 32) | rlo);
        } else {
            lt = window<K, NX>(X, 8 * J) < window<K, NX>(Xr, 8 * (kKmersPerThread - 1 - J));
        }
        // forward words at positions J + i, reverse words at 16 - J + i (same parity): the same pair rows
        // of the forward block, or kBlkPairs + 8 - J rows further in the reverse block: one select, one add
        const ScrBase b = scr_advance(fw, lt ? 0 : (kBlkPairs + 8 - J) * ROW);
        uint32_t W[NWD + 1];
        if constexpr ((J & 1) == 0) {
            pairs_even<0>(b, W);
        } else {  // one word, then pairs (the last pair may bring one word too many)
            W[0] = scr_load1<(J >> 1) * ROW + 4>(b);
            pairs_odd<0>(b, W);
        }
        return murmur_words<K, S42>(W, hc);
    }

    // PANIB_K1_GROUP k-mers are hashed back to back before any of them is offered to the table: the survivor
    // test ends a basic block, and inside one block ptxas interleaves the independent MurmurHash3 chains.
    static PANIB_HD void run(const uint32_t *X, const uint32_t *Xr, ScrBase fw, uint32_t vmask,
                             const HashConsts &hc, Emit &emit) {
        if constexpr (J < kKmersPerThread) {
            const Partial p0 = hash_one(X, Xr, fw, hc);
#if PANIB_K1_GROUP == 2
            const Partial p1 = KmerStep<K, J + 1, Emit, S42>::hash_one(X, Xr, fw, hc);
#endif
#if defined(PANIB_K1_VOTE) && defined(__CUDA_ARCH__)
            // experiment: one warp-uniform branch (vote) around the rare path instead of a divergent one
#if PANIB_K1_GROUP == 2
            const bool f0 = p0.prefilter(hc), f1 = p1.prefilter(hc);
            if (__any_sync(0xFFFFFFFFu, f0 | f1)) {
                if (f0 && (vmask & (1u << J))) emit(p0);
                if (f1 && (vmask & (2u << J))) emit(p1);
            }
#else
            const bool f0 = p0.prefilter(hc);
            if (__any_sync(0xFFFFFFFFu, f0)) {
                if (f0 && (vmask & (1u << J))) emit(p0);
            }
#endif
#else
            if (p0.prefilter(hc)) {  // ~1/scaled of the k-mers; validity is tested on this rare path only
#if defined(__CUDA_ARCH__)
                // Left in the source order, ptxas merges the validity test into the hot predicate (one LOP3 per
                // k-mer).  The thread's mask therefore waits in its own scratch (reverse position 0, which no
                // k-mer reads; stored by hash_thread_kmers) behind a volatile load that cannot be hoisted.
                uint32_t vm;
                asm volatile("ld.volatile.shared.u32 %0, [%1+%2];" : "=r"(vm) : "r"(fw), "n"(kBlkPairs * ROW));
#else
                const uint32_t vm = vmask;
#endif
                if (vm & (1u << J)) emit(p0);
            }
#if PANIB_K1_GROUP == 2
            if (p1.prefilter(hc)) {
                if (vmask & (2u << J)) emit(p1);
            }
#endif
#endif
            KmerStep<K, J + PANIB_K1_GROUP, Emit, S42>::run(X, Xr, fw, vmask, hc, emit);
        }
    }
};

// Phase B: hash the 16 k-mers of thread (u, a) of a tile.
//   sp    : packed words of the tile (word 0 bit 0 = tile position 0); sp[-1] readable
//   rcp   : packed words of the reverse stream RCm (written by phase A); rcp[-1] readable
//   blk   : this thread's scratch (scratch + 2 * tid; the CTA's scratch has kThreadsK1 threads per pair row)
//   vmask : bit j set = k-mer j is valid (0xFFFF for a tile without invalid bases)
//   emit  : callable(const Partial &) invoked for every valid k-mer that passes the prefilter
//
// Canonical choice (fwd < revcomp) on the packed form: with LSB-first packing it is F < R for the 2K-bit
// windows of the packed span and of the packed reverse stream.  For K = 31 / 32 the spans are loaded
// 32 - K bases early, so that k-mer j's window sits in bits [8j, 8j + 64) and is compared whole: the low
// 2 bits (K = 31) hold a neighbouring base on both sides, which could only decide a comparison whose 31
// real bases tie -- impossible, an odd-length k-mer is never its own reverse complement.  Windows of
// every fourth k-mer are register-aligned.  Other K use masked windows.
template <int K, bool S42 = false, class Emit>
PANIB_HD void hash_thread_kmers(const uint32_t *sp, const uint32_t *rcp, const uint32_t *blk, int u, int a,
                                uint32_t vmask, const HashConsts &hc, Emit &&emit) {
    using G_ = Geom<K>;
    constexpr int NX = G_::NX;
    constexpr bool WHOLE = (K == 31 || K == 32);  // 64-bit windows without masks
    constexpr int EARLY = WHOLE ? 32 - K : 0;     // bases the spans start early
    uint32_t X[NX], Xr[NX];
    {
        // forward span from tile position 64u + a - EARLY, reverse span from RCm position
        // 64(NU-1-u) + 3 - a - EARLY (word index and shift computed per lane)
        const int pf = 4 * kKmersPerThread * u + a - EARLY;
        const int pr = 4 * kKmersPerThread * (G_::NU - 1 - u) + 3 - a - EARLY;
        const uint32_t *src = sp + (pf >> 4);  // arithmetic shift: -1 >> 4 = -1
        const uint32_t *srcr = rcp + (pr >> 4);
        const uint32_t sf = 2u * (uint32_t)(pf & 15), sr = 2u * (uint32_t)(pr & 15);
#pragma unroll
        for (int w = 0; w < NX; w++) {
            X[w] = shf_r(src[w], src[w + 1], sf);
            Xr[w] = shf_r(srcr[w], srcr[w + 1], sr);
        }
    }
#if defined(__CUDA_ARCH__)
    // the validity mask for the rare path (KmerStep::run): this thread's reverse position 0, free during phase B
    asm volatile("st.volatile.shared.u32 [%0+%1], %2;" ::"r"(scr_base(blk)), "n"(kBlkPairs * 2 * kThreadsK1 * 4), "r"(vmask));
#endif
    KmerStep<K, 0, typename std::remove_reference<Emit>::type, S42>::run(X, Xr, scr_base(blk), vmask, hc, emit);
}

}  // namespace panib
