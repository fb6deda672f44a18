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
    void operator()(const Partial &p) {
        const uint64_t h = p.hash();
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
    const HashConsts hc = make_hash_consts(seed, max_hash);
    constexpr int S = kTileBases / kCtaTile;  // CTA tiles per stream tile
    static uint32_t stage[kSpLead + kTileWords];          // the kernel's staged tile with its lead-in words
    static uint32_t rcp_[kSpLead + kTileWords];           // packed reverse stream (phase A -> phase B)
    static uint32_t scratch[2 * kBlkPos * kThreadsK1];  // all threads' scratch blocks (shared memory on the GPU)
    for (int64_t tile = t0 * S; tile < t1 * S; tile++) {
        const uint32_t *sm = mask + tile * (kCtaTile / 32);
        for (int i = 0; i < kSpLead; i++) stage[i] = 0xDEADBEEFu;  // content must not matter
        memcpy(stage + kSpLead, packed + tile * (kCtaTile / 16), sizeof(uint32_t) * kTileWords);
        const uint32_t *sp = stage + kSpLead;
        uint32_t *rcp = rcp_ + kSpLead;
        for (int i = 0; i < kSpLead; i++) rcp_[i] = 0x13572468u;
        uint32_t any = 0;
        for (int i = 0; i < kTileMaskWords; i++) any |= sm[i];
        for (size_t i = 0; i < sizeof(scratch) / sizeof(scratch[0]); i++) scratch[i] = 0xA5A5A5A5u;
        for (int t = 0; t < Geom<K>::NI; t++) tile_expand_item<K>(sp, rcp, scratch, kThreadsK1, t);
        for (int lane = 0; lane < 32; lane++) tile_expand_halo<K>(sp, rcp, scratch, kThreadsK1, lane);
        for (int tid = 0; tid < kThreadsK1; tid++) {
            const int u = tid >> 2, a = tid & 3;
            const uint32_t vmask = any ? thread_valid_mask<K>(sm, u, a) : 0xFFFFu;
            hash_thread_kmers<K>(sp, rcp, scratch + 2 * tid, u, a, vmask, hc, c);
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
