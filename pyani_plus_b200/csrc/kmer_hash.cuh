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
    const uint64_t r = to_u64(a) + to_u64(b);
    return U64{(uint32_t)r, (uint32_t)(r >> 32)};
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
template <int K>
PANIB_HD uint64_t murmur_words(const uint32_t *W, uint32_t seed) {
    U64 h1{seed, 0u}, h2{seed, 0u};
    constexpr int nblocks = K / 16;
#pragma unroll
    for (int i = 0; i < nblocks; i++) {
        h1 = xor64(h1, mix_k1(U64{W[4 * i], W[4 * i + 1]}));
        h1 = mul5_add<0x52dce729u>(add64(rotl<27>(h1), h2));
        h2 = xor64(h2, mix_k2(U64{W[4 * i + 2], W[4 * i + 3]}));
        h2 = mul5_add<0x38495ab5u>(add64(rotl<31>(h2), h1));
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
    h1 = fmix(h1);
    h2 = fmix(h2);
    return to_u64(add64(h1, h2));
}

// ---- geometry of one thread's span for k-mer size K ------------------------------------------
template <int K>
struct Geom {
    static_assert(K >= 1 && K <= 32, "register-resident kernel handles 1 <= K <= 32");
    static constexpr int SPAN = 4 * (kKmersPerThread - 1) + K;  // bases touched by one thread
    static constexpr int NX = (2 * SPAN + 6 + 31) / 32;         // packed words (incl. 0..6 bits of alignment)
    static constexpr int NA = (SPAN + 3) / 4;                   // ASCII words per strand
    static constexpr int NWD = (K + 3) / 4;                     // ASCII words per k-mer
    static constexpr int RCSHIFT = 2 * (16 * NX - SPAN);        // bits to drop after pair-reversal
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

// Hash the 16 k-mers of thread (u, a) of a tile.
//   sp   : packed words of the tile (word 0 bit 0 = tile position 0), kTileWords valid words
//   sm   : validity-mask words of the tile (bit set = invalid), kTileMaskWords valid words
//   blk  : this thread's private scratch of 2*kBlkWords words (shared memory on the GPU, words of
//          consecutive threads kBlkWords apart: odd stride = bank-conflict free)
//   emit : callable(uint64_t h) invoked for every VALID k-mer (caller applies the max_hash test)
// DIRTY=false skips the per-k-mer validity test (caller guarantees the tile has no invalid base).
//
// The ASCII form of the thread's span (both strands) is written ONCE to the thread's scratch block;
// k-mer j then reads its NWD words from `fwd ? blk + j : blk + kBlkWords + 15 - j`: the canonical
// choice costs one address select and the words arrive through the (otherwise idle) load/store pipe,
// instead of NWD SEL instructions on the ALU pipe, which is the pipe that limits K1.
constexpr int kBlkWords = 23;  // ASCII words per strand of a span: ceil((60 + 32) / 4), odd on purpose

template <int K, bool DIRTY, class Emit>
PANIB_HD void hash_thread_kmers(const uint32_t *sp, const uint32_t *sm, uint32_t *blk, int blk_stride, int u, int a,
                                uint32_t seed, Emit &&emit) {
    using G_ = Geom<K>;
    constexpr int NX = G_::NX, NA = G_::NA, NWD = G_::NWD;
    static_assert(NA <= kBlkWords, "scratch block too small");
    const uint32_t *src = sp + (kKmersPerThread / 4) * u;  // span starts at tile position 4*KPT*u + a
    uint32_t X[NX];
#pragma unroll
    for (int w = 0; w < NX - 1; w++) X[w] = shf_r(src[w], src[w + 1], 2 * a);
    X[NX - 1] = src[NX - 1] >> (2 * a);

    // packed reverse complement of the span, LSB-first: Xr base p = comp(X base SPAN-1-p)
    uint32_t Z[NX + 2];
#pragma unroll
    for (int w = 0; w < NX; w++) Z[w] = revcomp16(X[NX - 1 - w]);
    Z[NX] = 0u;
    Z[NX + 1] = 0u;
    uint32_t Xr[NX];
    constexpr int wo = G_::RCSHIFT >> 5, sh = G_::RCSHIFT & 31;
#pragma unroll
    for (int w = 0; w < NX; w++) {
        uint32_t l = (w + wo < NX) ? Z[w + wo] : 0u;
        uint32_t h = (w + wo + 1 < NX) ? Z[w + wo + 1] : 0u;
        Xr[w] = sh ? shf_r(l, h, sh) : l;
    }

    // ASCII expansion of both strands, each base once per thread, straight into the scratch block
    // (element e of the block lives at blk[e * blk_stride]: stride 1 on the host, blockDim on the GPU
    // so that a warp's accesses to the same element are consecutive words)
    uint32_t *fw = blk, *rv = blk + kBlkWords * blk_stride;
#pragma unroll
    for (int w = 0; w < NX; w++) {
        if (4 * w < NA) {
            uint32_t e[4];
            expand16(X[w], e);
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (4 * w + i < NA) fw[(4 * w + i) * blk_stride] = e[i];
            expand16(Xr[w], e);
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (4 * w + i < NA) rv[(4 * w + i) * blk_stride] = e[i];
        }
    }

    // PANIB_K1_GROUP k-mers are hashed back to back before any of them is offered to the table: the
    // conditional insert (a call) ends a basic block, and inside one block the compiler can interleave
    // the independent MurmurHash3 chains of the group.  Measured neutral on B200 (groups of 1/2/4/8:
    // 2.68 / 2.68 / 2.71 / 2.68 ms at config 2 -- the kernel is bound by pipe throughput, not by
    // dependency stalls), so the default stays 1.
#ifndef PANIB_K1_GROUP
#define PANIB_K1_GROUP 1
#endif
    constexpr int GRP = PANIB_K1_GROUP;
    static_assert(kKmersPerThread % GRP == 0, "group size must divide the k-mers per thread");
#pragma unroll
    for (int j0 = 0; j0 < kKmersPerThread; j0 += GRP) {
        uint64_t h[GRP];
        bool valid[GRP];
#pragma unroll
        for (int jj = 0; jj < GRP; jj++) {
            const int j = j0 + jj;
            valid[jj] = true;
            if (DIRTY) {
                const int pos = 4 * kKmersPerThread * u + a + 4 * j;
                uint32_t mw = shf_r(sm[pos >> 5], sm[(pos >> 5) + 1], pos & 31);
                if (K < 32) mw &= (1u << (K & 31)) - 1u;
                valid[jj] = (mw == 0u);
            }
            const uint64_t F = window<K, NX>(X, 8 * j);
            const uint64_t R = window<K, NX>(Xr, 8 * (kKmersPerThread - 1 - j));
            const uint32_t *words = F < R ? fw + j * blk_stride : rv + (kKmersPerThread - 1 - j) * blk_stride;
            uint32_t W[NWD];
#pragma unroll
            for (int i = 0; i < NWD; i++) W[i] = words[i * blk_stride];
            h[jj] = murmur_words<K>(W, seed);
        }
#pragma unroll
        for (int jj = 0; jj < GRP; jj++)
            if (!DIRTY || valid[jj]) emit(h[jj]);
    }
}

}  // namespace panib
