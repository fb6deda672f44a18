// hostemu.cpp -- CPU emulation of the K1 thread logic, for tests only.
// Compiles the SAME headers the CUDA kernels use (kmer_hash.cuh, pack.cuh) with g++ and walks the
// kernel's (tile, thread) geometry sequentially, so that bit-exactness of the register-level
// algorithm can be checked against the oracle on a machine without a GPU
// (tests/test_hostemu.py).  Never loaded by the product: there is no CPU compute path.
#include <stdint.h>
#include <string.h>

#include "kmer_hash.cuh"
#include "pack.cuh"

using namespace panib;

extern "C" {

__attribute__((visibility("default"))) void emu_pack_ascii(const uint8_t *ascii, int64_t n_bases,
                                                            uint32_t *packed, uint32_t *mask) {
    for (int64_t g = 0; g < n_bases / 32; g++) {
        uint32_t wa[4], wb[4], ia, ib;
        memcpy(wa, ascii + 32 * g, 16);
        memcpy(wb, ascii + 32 * g + 16, 16);
        packed[2 * g] = pack16(wa, ia);
        packed[2 * g + 1] = pack16(wb, ib);
        mask[g] = ia | (ib << 16);
    }
}

}  // extern "C"

struct Collect {
    uint64_t max_hash;
    uint64_t *out;
    int64_t cap;
    int64_t n;
    void operator()(uint64_t h) {
        if (h != 0ull && h <= max_hash) {
            if (n < cap) out[n] = h;
            n++;
        }
    }
};

template <int K>
static int64_t run(const uint32_t *packed, const uint32_t *mask, int64_t t0, int64_t t1, uint32_t seed,
                   uint64_t max_hash, uint64_t *out, int64_t cap) {
    Collect c{max_hash, out, cap, 0};
    constexpr int S = kTileBases / kCtaTile;  // CTA tiles per stream tile
    for (int64_t tile = t0 * S; tile < t1 * S; tile++) {
        const uint32_t *sp = packed + tile * (kCtaTile / 16);
        const uint32_t *sm = mask + tile * (kCtaTile / 32);
        uint32_t any = 0;
        for (int i = 0; i < kTileMaskWords; i++) any |= sm[i];
        for (int tid = 0; tid < kThreadsK1; tid++) {
            const int u = tid >> 2, a = tid & 3;
            uint32_t blk[2 * kBlkWords];  // the thread's scratch block (shared memory on the GPU)
            if (any) hash_thread_kmers<K, true>(sp, sm, blk, 1, u, a, seed, c);
            else hash_thread_kmers<K, false>(sp, sm, blk, 1, u, a, seed, c);
        }
    }
    return c.n;
}

extern "C" {

// Emulates sketch_hash_kernel<K> over tiles [t0, t1); returns the number of surviving hashes
// (unsorted, duplicates included) and writes up to cap of them.  -1 for a k that has no fast kernel.
__attribute__((visibility("default"))) int64_t emu_sketch_tiles(const uint32_t *packed, const uint32_t *mask,
                                                                 int64_t t0, int64_t t1, int k, uint32_t seed,
                                                                 uint64_t max_hash, uint64_t *out, int64_t cap) {
    switch (k) {
    case 21: return run<21>(packed, mask, t0, t1, seed, max_hash, out, cap);
    case 31: return run<31>(packed, mask, t0, t1, seed, max_hash, out, cap);
    case 32: return run<32>(packed, mask, t0, t1, seed, max_hash, out, cap);
    case 15: return run<15>(packed, mask, t0, t1, seed, max_hash, out, cap);
    case 16: return run<16>(packed, mask, t0, t1, seed, max_hash, out, cap);
    case 7: return run<7>(packed, mask, t0, t1, seed, max_hash, out, cap);
    default: return -1;
    }
}
}
