/*
 * panib_oracle.c -- CPU ORACLE for the pyani-plus `sourmash` method hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (pyani_plus_b200/) never links, imports or falls back to this file.
 *
 * What it restates.  pyani-plus does no arithmetic itself on this path: it shells out to
 *   sourmash scripts singlesketch -I DNA -p k=K,scaled=S      (pyani_plus/methods/sourmash.py:67-83)
 *   sourmash scripts manysearch -m DNA -t 0                   (pyani_plus/methods/sourmash.py:184-200)
 * provided by the un-vendored Rust crates sourmash (sourmash-minimal>=4.8.11) and
 * sourmash_plugin_branchwater (>=0.9.11) (requirements-thirdparty-linux.txt:8-9).  Neither is
 * installable here, so this file restates their published algorithm (FracMinHash over canonical
 * DNA k-mers hashed with MurmurHash3_x64_128 seed 42; containment -> ANI) deliberately naively:
 * byte strings, memcmp, qsort.  PARITY IS PINNED, not assumed: tests/test_oracle_golden.py checks
 * this code against every sourmash golden the reference's own tests hold for the path
 * (9 .sig files, 3 manysearch.csv = 27 rows, tests/test_coverage.py:162-174 values); see SURVEY.md 8c.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------------------------
 * max_hash for a FracMinHash `scaled` value.
 * sourmash (Rust core, max_hash_for_scaled): 0 -> 0, 1 -> u64::MAX, else (u64::MAX as f64 / scaled as f64) as u64.
 * `u64::MAX as f64` rounds to 2^64.  Pinned by the fixture .sig files: scaled=300 -> 61489146912365176,
 * scaled=1000 -> 18446744073709552 (tests/fixtures/.../intermediates/sourmash/ *.sig "max_hash").
 * ------------------------------------------------------------------------------------------- */
ORACLE_API uint64_t oracle_max_hash(uint64_t scaled) {
    if (scaled == 0) return 0;
    if (scaled == 1) return UINT64_MAX;
    double v = 18446744073709551616.0 / (double)scaled;
    return (uint64_t)v;
}

/* ---------------------------------------------------------------------------------------------
 * MurmurHash3_x64_128 (public domain, Austin Appleby), generic length; returns the first 64-bit
 * word (h1) which is what sourmash's "0.murmur64" hash_function keeps.
 * ------------------------------------------------------------------------------------------- */
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33; return k;
}
ORACLE_API uint64_t oracle_murmur64(const uint8_t *key, int len, uint32_t seed) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    const int nblocks = len / 16;
    for (int i = 0; i < nblocks; i++) {
        uint64_t k1, k2;
        memcpy(&k1, key + 16 * i, 8);      /* little-endian host assumed (x86-64 / aarch64) */
        memcpy(&k2, key + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = key + nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    switch (len & 15) {
    case 15: k2 ^= ((uint64_t)tail[14]) << 48; /* fallthrough */
    case 14: k2 ^= ((uint64_t)tail[13]) << 40; /* fallthrough */
    case 13: k2 ^= ((uint64_t)tail[12]) << 32; /* fallthrough */
    case 12: k2 ^= ((uint64_t)tail[11]) << 24; /* fallthrough */
    case 11: k2 ^= ((uint64_t)tail[10]) << 16; /* fallthrough */
    case 10: k2 ^= ((uint64_t)tail[9]) << 8;   /* fallthrough */
    case 9:  k2 ^= ((uint64_t)tail[8]) << 0;
             k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; /* fallthrough */
    case 8:  k1 ^= ((uint64_t)tail[7]) << 56; /* fallthrough */
    case 7:  k1 ^= ((uint64_t)tail[6]) << 48; /* fallthrough */
    case 6:  k1 ^= ((uint64_t)tail[5]) << 40; /* fallthrough */
    case 5:  k1 ^= ((uint64_t)tail[4]) << 32; /* fallthrough */
    case 4:  k1 ^= ((uint64_t)tail[3]) << 24; /* fallthrough */
    case 3:  k1 ^= ((uint64_t)tail[2]) << 16; /* fallthrough */
    case 2:  k1 ^= ((uint64_t)tail[1]) << 8;  /* fallthrough */
    case 1:  k1 ^= ((uint64_t)tail[0]) << 0;
             k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

/* ---------------------------------------------------------------------------------------------
 * FracMinHash sketch of one genome given as concatenated record sequences.
 *   seq          : all records' sequence bytes back to back (whitespace already removed, as
 *                  pyani_plus/utils.py:40-90 fasta_bytes_iterator yields them); any case.
 *   rec_offsets  : n_records+1 offsets into seq; k-mers never span records (SURVEY.md 8c conv. 1).
 * Conventions (each pinned by a fixture unless flagged there):
 *   upper-case; skip any window containing a byte outside ACGT; canonical = lexicographically
 *   smaller of the k-mer and its reverse complement; h = murmur64(canonical, seed 42);
 *   keep iff h <= max_hash and h != 0; result is the sorted set.
 * Returns the number of distinct kept hashes; writes min(count, cap) of them to out (sorted).
 * ------------------------------------------------------------------------------------------- */
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

static inline int base_upper_valid(uint8_t c, uint8_t *up) {
    if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
    *up = c;
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}
static inline uint8_t comp_base(uint8_t c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; default: return 'A'; }
}

ORACLE_API int64_t oracle_sketch(const uint8_t *seq, const int64_t *rec_offsets, int64_t n_records,
                                 int k, uint64_t max_hash, uint32_t seed,
                                 uint64_t *out, int64_t cap) {
    if (k < 1 || k > 255) return -1;
    int64_t n = 0, alloc = 1024;
    uint64_t *buf = (uint64_t *)malloc((size_t)alloc * 8);
    uint8_t fwd[256], rc[256];
    for (int64_t r = 0; r < n_records; r++) {
        const uint8_t *s = seq + rec_offsets[r];
        int64_t len = rec_offsets[r + 1] - rec_offsets[r];
        for (int64_t i = 0; i + k <= len; i++) {
            int ok = 1;
            for (int j = 0; j < k; j++) {
                if (!base_upper_valid(s[i + j], &fwd[j])) { ok = 0; break; }
            }
            if (!ok) continue;
            for (int j = 0; j < k; j++) rc[j] = comp_base(fwd[k - 1 - j]);
            const uint8_t *canon = memcmp(fwd, rc, (size_t)k) <= 0 ? fwd : rc;
            uint64_t h = oracle_murmur64(canon, k, seed);
            if (h == 0 || h > max_hash) continue;
            if (n == alloc) { alloc *= 2; buf = (uint64_t *)realloc(buf, (size_t)alloc * 8); }
            buf[n++] = h;
        }
    }
    qsort(buf, (size_t)n, 8, cmp_u64);
    int64_t u = 0;
    for (int64_t i = 0; i < n; i++) {
        if (i == 0 || buf[i] != buf[i - 1]) {
            if (u < cap) out[u] = buf[i];
            u++;
        }
    }
    free(buf);
    return u;
}

/* ---------------------------------------------------------------------------------------------
 * Same sketch, engineered the way a production CPU implementation (sourmash's Rust core) does it:
 * the record is upper-cased and reverse-complemented ONCE, the number of invalid bytes in the
 * current window is kept rolling, and each k-mer costs one memcmp + one murmur.  This is the
 * function the CPU *baseline* timings use (oracle_sketch above stays the naive checker; the two
 * are compared on every fixture by tests/test_oracle_golden.py).
 * ------------------------------------------------------------------------------------------- */
ORACLE_API int64_t oracle_sketch_fast(const uint8_t *seq, const int64_t *rec_offsets, int64_t n_records,
                                      int k, uint64_t max_hash, uint32_t seed,
                                      uint64_t *out, int64_t cap) {
    if (k < 1 || k > 255) return -1;
    int64_t n = 0, alloc = 1024, maxlen = 0;
    for (int64_t r = 0; r < n_records; r++) {
        int64_t len = rec_offsets[r + 1] - rec_offsets[r];
        if (len > maxlen) maxlen = len;
    }
    uint64_t *buf = (uint64_t *)malloc((size_t)alloc * 8);
    uint8_t *up = (uint8_t *)malloc((size_t)maxlen + 1);
    uint8_t *rc = (uint8_t *)malloc((size_t)maxlen + 1);
    uint8_t *bad = (uint8_t *)malloc((size_t)maxlen + 1);
    for (int64_t r = 0; r < n_records; r++) {
        const uint8_t *s = seq + rec_offsets[r];
        const int64_t len = rec_offsets[r + 1] - rec_offsets[r];
        if (len < k) continue;
        for (int64_t i = 0; i < len; i++) {
            bad[i] = (uint8_t)!base_upper_valid(s[i], &up[i]);
            rc[len - 1 - i] = bad[i] ? 'N' : comp_base(up[i]);
        }
        int64_t nbad = 0;
        for (int j = 0; j < k - 1; j++) nbad += bad[j];
        for (int64_t i = 0; i + k <= len; i++) {
            nbad += bad[i + k - 1];
            if (nbad == 0) {
                const uint8_t *f = up + i, *c = rc + (len - i - k);
                const uint8_t *canon = memcmp(f, c, (size_t)k) <= 0 ? f : c;
                const uint64_t h = oracle_murmur64(canon, k, seed);
                if (h != 0 && h <= max_hash) {
                    if (n == alloc) { alloc *= 2; buf = (uint64_t *)realloc(buf, (size_t)alloc * 8); }
                    buf[n++] = h;
                }
            }
            nbad -= bad[i];
        }
    }
    qsort(buf, (size_t)n, 8, cmp_u64);
    int64_t u = 0;
    for (int64_t i = 0; i < n; i++) {
        if (i == 0 || buf[i] != buf[i - 1]) {
            if (u < cap) out[u] = buf[i];
            u++;
        }
    }
    free(buf); free(up); free(rc); free(bad);
    return u;
}

/* Exact |A n B| of two sorted, duplicate-free u64 lists (what manysearch calls intersect_hashes). */
ORACLE_API int64_t oracle_intersect(const uint64_t *a, int64_t na, const uint64_t *b, int64_t nb) {
    int64_t i = 0, j = 0, c = 0;
    while (i < na && j < nb) {
        if (a[i] < b[j]) i++;
        else if (a[i] > b[j]) j++;
        else { c++; i++; j++; }
    }
    return c;
}

/* sourmash ani_from_containment: 0 -> 0, 1 -> 1, else 1 - (1 - c^(1/k))  (SURVEY.md 8c conv. 11). */
ORACLE_API double oracle_ani_from_containment(double c, int k) {
    if (c == 0.0) return 0.0;
    if (c == 1.0) return 1.0;
    return 1.0 - (1.0 - pow(c, 1.0 / (double)k));
}

/* One manysearch row (branchwater): returns 0 and leaves outputs untouched when there is no row
 * (ov == 0 or an empty sketch, SURVEY.md 8c conv. 13-14), else 1 with
 *   out[0]=containment(query in subject) out[1]=max_containment out[2]=jaccard
 *   out[3]=query_containment_ani out[4]=match_containment_ani out[5]=average_containment_ani
 *   out[6]=max_containment_ani.
 * pyani-plus maps identity := out[6], cov_query := out[3] (private_cli.py:1875-1887). */
ORACLE_API int oracle_pair_row(int64_t ov, int64_t nq, int64_t ns, int k, double *out) {
    if (ov <= 0 || nq <= 0 || ns <= 0) return 0;
    double cq = (double)ov / (double)nq, cs = (double)ov / (double)ns;
    double qani = oracle_ani_from_containment(cq, k), mani = oracle_ani_from_containment(cs, k);
    out[0] = cq;
    out[1] = cq > cs ? cq : cs;
    out[2] = (double)ov / ((double)nq + (double)ns - (double)ov);
    out[3] = qani;
    out[4] = mani;
    out[5] = (qani + mani) / 2.0;
    out[6] = qani > mani ? qani : mani;
    return 1;
}

/* ---------------------------------------------------------------------------------------------
 * Synthetic genome family used by BASELINE.json configs 2-5: one random ancestor,
 * every genome an independent substitution-only mutant at identity p_g in [0.80, 0.99).
 * Counter-based so any base of any genome can be regenerated independently; the CUDA generator
 * in the product (csrc/synth.cu) states the same integer recipe and is tested against this one.
 * ------------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
#define SYNTH_ANCESTOR_ID 0xFFFFFFFFFFFFULL

/* threshold (out of 2^53) below which a base of genome g is substituted: (1 - p_g) * 2^53 */
ORACLE_API uint64_t oracle_synth_threshold(uint64_t seed, uint64_t g) {
    uint64_t u = splitmix64(splitmix64(seed ^ 0x5851F42D4C957F2DULL) + g) >> 11; /* 53 bits */
    /* p_g = 0.80 + 0.19*u/2^53 ; threshold = (0.20 - 0.19*u/2^53) * 2^53, in integers */
    uint64_t two53 = 1ULL << 53;
    uint64_t thr = two53 / 5 - (uint64_t)(((__uint128_t)u * 19) / 100);
    return thr;
}
static inline uint8_t synth_code(uint64_t ka, uint64_t kg, uint64_t thr, uint64_t pos) {
    uint8_t anc = (uint8_t)(splitmix64(ka ^ pos) & 3);
    uint64_t r = splitmix64(kg ^ pos);
    if ((r >> 11) < thr) {
        uint8_t step = (uint8_t)(1 + ((r & 0x7FF) % 3));
        return (uint8_t)((anc + step) & 3);
    }
    return anc;
}
ORACLE_API void oracle_synth_genome(uint64_t seed, uint64_t g, int64_t length, uint8_t *out_ascii) {
    static const char LUT[4] = {'A', 'C', 'G', 'T'};
    uint64_t thr = oracle_synth_threshold(seed, g);
    uint64_t ka = splitmix64(seed + SYNTH_ANCESTOR_ID);   /* ancestor stream key */
    uint64_t kg = splitmix64(seed + g);                   /* genome g's mutation stream key */
    for (int64_t p = 0; p < length; p++) out_ascii[p] = (uint8_t)LUT[synth_code(ka, kg, thr, (uint64_t)p)];
}

/* ---------------------------------------------------------------------------------------------
 * Batched, OpenMP-parallel forms: the timed CPU baseline (bench.py cpu_baseline / --impl reference).
 * Same arithmetic as above, threads over genomes / over pairs.
 * ------------------------------------------------------------------------------------------- */
ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Sketch n_genomes single-record synthetic genomes g0..g0+n-1 of `length` bases.
 * out is n_genomes*cap u64, counts[n] receives the sketch sizes.  Returns total bases hashed. */
ORACLE_API int64_t oracle_synth_sketch_batch(uint64_t seed, int64_t g0, int64_t n_genomes, int64_t length,
                                             int k, uint64_t max_hash, uint64_t *out, int64_t cap,
                                             int64_t *counts) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t g = 0; g < n_genomes; g++) {
        uint8_t *seq = (uint8_t *)malloc((size_t)length);
        oracle_synth_genome(seed, (uint64_t)(g0 + g), length, seq);
        int64_t offs[2] = {0, length};
        counts[g] = oracle_sketch_fast(seq, offs, 1, k, max_hash, 42, out + g * cap, cap);
        free(seq);
    }
    return n_genomes * length;
}

/* Sketch genomes already in memory (ASCII, single record each, fixed stride). */
ORACLE_API void oracle_sketch_batch(const uint8_t *seqs, int64_t n_genomes, int64_t length,
                                    int k, uint64_t max_hash, uint64_t *out, int64_t cap, int64_t *counts) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t g = 0; g < n_genomes; g++) {
        int64_t offs[2] = {0, length};
        counts[g] = oracle_sketch_fast(seqs + g * length, offs, 1, k, max_hash, 42, out + g * cap, cap);
    }
}

/* All-vs-all intersection counts for n sketches stored with fixed stride `cap`;
 * ov is n*n int64 (symmetric, diagonal = sizes). */
ORACLE_API void oracle_intersect_all(const uint64_t *hashes, int64_t cap, const int64_t *counts,
                                     int64_t n, int64_t *ov) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; i++) {
        ov[i * n + i] = counts[i];
        for (int64_t j = i + 1; j < n; j++) {
            int64_t c = oracle_intersect(hashes + i * cap, counts[i], hashes + j * cap, counts[j]);
            ov[i * n + j] = c;
            ov[j * n + i] = c;
        }
    }
}
