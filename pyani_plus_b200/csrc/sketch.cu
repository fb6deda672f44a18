// sketch.cu -- stage 0 (pack) and stage 1 (kernel K1: FracMinHash sketching) of libpanib200.so.
//
// Replaces the arithmetic of `sourmash scripts singlesketch -I DNA -p k=K,scaled=S`
// (reference call site: pyani_plus/methods/sourmash.py:67-83, prepare_genomes :34-84).
//
// K1 is integer-ALU bound by construction (about 12 64-bit multiplies + rotates per k-mer, see
// DESIGN.md), so the design goal is instructions per k-mer, not bytes: the 2-bit stream is staged
// once per tile in shared memory (tile + K-1 halo), every base is expanded to ASCII once per thread
// (kmer_hash.cuh), and the only global writes are the ~1/scaled surviving hashes, inserted into
// value-range buckets of the genome's row so that a per-bucket sort yields the globally sorted,
// duplicate-free sketch.
#include <chrono>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"
#include "kmer_hash.cuh"
#include "pack.cuh"

namespace panib {

// ------------------------------------------------------------------------------------------------
// stage 0: ASCII stream -> 2-bit packed + validity mask.  One thread per 32 bases (two 16-byte loads).
// HBM-bound: 1 B/base read, 0.375 B/base written.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_ascii_kernel(const uint4 *__restrict__ ascii, int64_t n_groups,
                                                         uint32_t *__restrict__ packed,
                                                         uint32_t *__restrict__ mask) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        const uint4 a = __ldg(&ascii[2 * g]);
        const uint4 b = __ldg(&ascii[2 * g + 1]);
        const uint32_t wa[4] = {a.x, a.y, a.z, a.w};
        const uint32_t wb[4] = {b.x, b.y, b.z, b.w};
        uint32_t ia, ib;
        const uint32_t pa = pack16(wa, ia);
        const uint32_t pb = pack16(wb, ib);
        reinterpret_cast<uint2 *>(packed)[g] = make_uint2(pa, pb);
        mask[g] = ia | (ib << 16);
    }
}

// ------------------------------------------------------------------------------------------------
// insertion of a surviving hash into its genome's bucketed table
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ void table_insert(uint64_t *__restrict__ row, int nb, uint64_t bmul, uint64_t h,
                                             int32_t *flag, int32_t *status) {
    if (h == 0ull) return;  // sourmash skips hash 0 (tested here, on the rare path only)
    if (h == kEmpty) {  // only reachable when max_hash == 2^64-1 (scaled == 1); kept as a flag
        atomicOr(flag, 1);
        return;
    }
    uint32_t b = 0;
    if (nb > 1) {
        b = (uint32_t)__umul64hi(h, bmul);
        if (b >= (uint32_t)nb) b = (uint32_t)nb - 1u;
    }
    unsigned long long *bucket = reinterpret_cast<unsigned long long *>(row) + (size_t)b * kBucketSlots;
    uint32_t s = (uint32_t)h & (kBucketSlots - 1);
    for (int probe = 0; probe < kBucketSlots; ++probe) {
        const unsigned long long prev = atomicCAS(&bucket[s], (unsigned long long)kEmpty, (unsigned long long)h);
        if (prev == kEmpty || prev == h) return;  // inserted, or already present (set semantics)
        s = (s + 1) & (kBucketSlots - 1);
    }
    atomicOr(status, PANIB_ST_BUCKET_OVERFLOW);
}

// what phase B needs to insert a survivor: per-tile genome data, looked up by ONE thread per tile and
// parked in shared memory (the other 255 threads read it on the rare survivor path only)
struct GenomeSlot {
    uint64_t *row;
    uint64_t bmul;
    int32_t *flag;
    int nb;
    int genome;
};

// Survivors of one K1 launch, parked instead of inserted: CTA b appends (hash, genome) to its own region
// [b * cap, (b+1) * cap) of a caller-provided workspace (panib_set_workspace), the position coming from a
// SHARED-memory counter, and sketch_scatter_kernel inserts them afterwards with every thread busy.  The hot
// loop then never waits ~1 us for a global atomicCAS in a mostly idle warp (which is also what lets the
// warps of a CTA drift apart between barriers).  A full region falls back to the direct insert.
struct SurvivorBuf {
    uint64_t *hash;     // [ctas * cap]
    uint32_t *genome;   // [ctas * cap]
    uint32_t *count;    // [ctas]
    uint32_t cap;       // entries per CTA; 0 = no workspace: insert directly
};

struct EmitToTable {
    const GenomeSlot *slot;  // shared memory
    uint64_t max_hash;
    int32_t *status;
    SurvivorBuf buf;
    uint32_t *s_count;  // shared: survivors parked by this CTA
    // survivor path of the fast kernel: finish the hash, exact test, park (or insert)
    __device__ __forceinline__ void operator()(const Partial &p) const {
        const uint64_t h = p.hash();
        if (h > max_hash) return;
        if (buf.cap) {
            const uint32_t at = atomicAdd(s_count, 1u);
            if (at < buf.cap) {
                const size_t e = (size_t)blockIdx.x * buf.cap + at;
                buf.hash[e] = h;
                buf.genome[e] = (uint32_t)slot->genome;
                return;
            }
        }
        table_insert(slot->row, slot->nb, slot->bmul, h, slot->flag, status);
    }
};

// generic kernel: complete hash, genome data in registers
struct EmitHash {
    uint64_t *row;
    int nb;
    uint64_t bmul;
    uint64_t max_hash;
    int32_t *flag;
    int32_t *status;
    __device__ __forceinline__ void operator()(uint64_t h) const {
        if (h <= max_hash) table_insert(row, nb, bmul, h, flag, status);  // keep iff 0 < h <= max_hash
    }
};

// genome owning `tile`: largest g with tile_off[g] <= tile.  `g` is the previous answer, tried first.
__device__ __forceinline__ int find_genome(const int64_t *__restrict__ tile_off, int n_genomes, int64_t tile,
                                           int g) {
    if (tile >= __ldg(tile_off + g) && tile < __ldg(tile_off + g + 1)) return g;
    int lo = 0, hi = n_genomes - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(tile_off + mid) <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------
// K1, fast form for compile-time K <= 32 (kmer_hash.cuh describes the two phases and the thread
// geometry).  Persistent CTAs (grid = SMs x resident CTAs) draw CTA tiles from a ticket counter; the
// packed tile after the current one is prefetched with cp.async into the other half of a double buffer.
// Per tile: barrier (tile landed, scratch free) -> phase A, cooperative ASCII expansion of both strands
// into the threads' scratch blocks -> barrier -> phase B, 16 k-mers per thread.
// __launch_bounds__(256, 4): 64 registers, 32 resident warps per SM (K1 is bound by issue slots and the
// two integer pipes; 32 warps measured 8 % faster than 16).
// ------------------------------------------------------------------------------------------------
#ifndef PANIB_K1_MINBLOCKS
#define PANIB_K1_MINBLOCKS 4
#endif
constexpr int kSpStage = kSpLead + kTileWords;  // lead-in + tile + halo words of one staging buffer

#ifndef PANIB_K1_WARP
#define PANIB_K1_WARP 0
#endif
#if PANIB_K1_WARP
// ------------------------------------------------------------------------------------------------
// Warp-private form (PANIB_K1_WARP, geometry compiled with PANIB_K1_THREADS = 32): every WARP is an
// independent worker with its own 512-k-mer tiles, staging buffers, scratch and genome slot; the only CTA
// barrier is the one at the end of the kernel.  The CTA form above synchronises its 8 warps twice per tile,
// and ncu attributes 19 % of all warp samples to those two barriers (12.6 % top of the loop, 6.2 % between the
// phases): the warps of a CTA do not finish a phase together -- the scheduler is not fair, survivors take the
// rare path, one lane waits for the ticket atomic -- and every one of them waits for the slowest.
// Work is handed out in units of kUnitTiles warp tiles (one ticket each).
// ------------------------------------------------------------------------------------------------
static_assert(kThreadsK1 == 32, "the warp-private kernel needs the 32-thread tile geometry");
constexpr int kLaunchThreadsK1 = 256;
constexpr int kWarpsK1 = kLaunchThreadsK1 / 32;
constexpr int kUnitTiles = 2;  // warp tiles per ticket (divides kTileBases / kCtaTile: a unit lies in one genome)
static_assert((kTileBases / kCtaTile) % kUnitTiles == 0, "unit geometry");

template <int K, bool S42>
__global__ void __launch_bounds__(kLaunchThreadsK1, PANIB_K1_MINBLOCKS)
sketch_hash_kernel(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ mask,
                   const int64_t *__restrict__ tile_off, int n_genomes, int64_t tile_begin, int64_t n_tiles,
                   const HashConsts hc, const int32_t *__restrict__ nb,
                   const uint64_t *__restrict__ bmul, uint64_t *__restrict__ table, int64_t row_stride,
                   int32_t *flags, int32_t *status, uint32_t *ticket, const SurvivorBuf surv) {
    const uint64_t max_hash = hc.max_hash;
    __shared__ __align__(16) uint32_t sp_all[kWarpsK1][2][kSpStage];
    __shared__ __align__(16) uint32_t sm_all[kWarpsK1][2][kTileMaskWords];
    __shared__ __align__(16) uint32_t rcp_all[kWarpsK1][kSpStage];
    __shared__ GenomeSlot s_slot[kWarpsK1];
    __shared__ int64_t s_range[kWarpsK1][2];  // stream tiles [lo, hi) of the genome in s_slot
    __shared__ uint32_t s_nsurv;
    extern __shared__ __align__(16) uint32_t scratch_all[];  // kK1DynSmem bytes: one block of 2 * kBlkPos * 32 words per warp
    constexpr int S = kTileBases / kCtaTile;  // warp tiles per stream tile
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t *const scratch = scratch_all + w * (2 * kBlkPos * 32);
    uint32_t(*const sp_)[kSpStage] = sp_all[w];
    uint32_t(*const sm)[kTileMaskWords] = sm_all[w];
    uint32_t *const rcp = rcp_all[w] + kSpLead;
    if (threadIdx.x == 0) s_nsurv = 0u;
    if (lane < 2 * kSpLead) sp_[lane / kSpLead][lane % kSpLead] = 0u;  // lead-in words: read, never used
    else if (lane < 3 * kSpLead) rcp_all[w][lane - 2 * kSpLead] = 0u;
    else if (lane < 3 * kSpLead + 2) s_range[w][lane - 3 * kSpLead] = 0;  // empty range: the first unit looks its genome up
    __syncthreads();  // s_nsurv
    const int64_t wt0 = tile_begin * S;                              // first warp tile of the launch
    const int64_t units = (n_tiles - tile_begin) * (S / kUnitTiles);  // tickets in the launch
    const int64_t n_workers = (int64_t)gridDim.x * kWarpsK1;
    auto prefetch = [&](int64_t t, int b) {
        const uint32_t *gp = packed + t * (kCtaTile / 16);
        const uint32_t *gm = mask + t * (kCtaTile / 32);
        if (lane < kTileWords / 4) cp_async16(sp_[b] + kSpLead + 4 * lane, gp + 4 * lane);
        else if (lane < kTileWords / 4 + kTileMaskWords / 4)
            cp_async16(sm[b] + 4 * (lane - kTileWords / 4), gm + 4 * (lane - kTileWords / 4));
        cp_async_commit();
    };
    static_assert(kTileWords / 4 + kTileMaskWords / 4 <= 32, "one warp stages a tile with one cp.async per lane");
    int64_t unit = (int64_t)blockIdx.x * kWarpsK1 + w;  // first unit: static; later ones from the ticket counter
    int64_t next_unit = units;                          // known once sub-tile 0 of `unit` is done
    int64_t wt = wt0 + unit * kUnitTiles;
    int sub = 0;
    if (unit < units) prefetch(wt, 0);
    const int u = lane >> 2, a = lane & 3;
    const EmitToTable emit{&s_slot[w], max_hash, status, surv, &s_nsurv};
    int g = 0;
    for (int it = 0; unit < units; ++it) {
        const int cur = it & 1;
        cp_async_wait<0>();
        __syncwarp();  // tile `cur` has landed; every lane is done with the previous tile's scratch
        const bool last_sub = sub == kUnitTiles - 1;
        const int64_t following = last_sub ? wt0 + next_unit * kUnitTiles : wt + 1;
        if (!last_sub || next_unit < units) prefetch(following, cur ^ 1);
        const uint32_t *sp = sp_[cur] + kSpLead;
        uint32_t drawn = 0;
        if (sub == 0 && lane == 0) {  // once per unit: the next ticket, and this unit's genome
            if (ticket) drawn = atomicAdd(ticket, 1u);
            const int64_t st_tile = wt / S;
            if (st_tile < s_range[w][0] || st_tile >= s_range[w][1]) {
                g = find_genome(tile_off, n_genomes, st_tile, g);
                s_slot[w] = GenomeSlot{table + (size_t)g * row_stride, __ldg(bmul + g), flags + g, __ldg(nb + g), g};
                s_range[w][0] = __ldg(tile_off + g);
                s_range[w][1] = __ldg(tile_off + g + 1);
            }
        }
        // phase A: one item per lane, then the halo
        tile_expand_item<K>(sp, rcp, scratch, 32, lane);
        tile_expand_halo<K>(sp, rcp, scratch, 32, lane);
        const uint32_t mw = lane < kTileMaskWords ? sm[cur][lane] : 0u;
        const bool dirty = __any_sync(0xFFFFFFFFu, mw != 0u) != 0;
        __syncwarp();
        // phase B
        const uint32_t vmask = dirty ? thread_valid_mask<K>(sm[cur], u, a) : 0xFFFFu;
        hash_thread_kmers<K, S42>(sp, rcp, scratch + 2 * lane, u, a, vmask, hc, emit);
        if (sub == 0) {
            const uint32_t d0 = __shfl_sync(0xFFFFFFFFu, drawn, 0);
            next_unit = ticket ? n_workers + (int64_t)d0 : unit + n_workers;
        }
        if (last_sub) {
            unit = next_unit;
            wt = wt0 + unit * kUnitTiles;
            sub = 0;
        } else {
            sub++;
            wt++;
        }
    }
    if (surv.cap) {
        __syncthreads();
        if (threadIdx.x == 0) surv.count[blockIdx.x] = s_nsurv < surv.cap ? s_nsurv : surv.cap;
    }
}
#else
constexpr int kLaunchThreadsK1 = kThreadsK1;
template <int K, bool S42>
__global__ void __launch_bounds__(kThreadsK1, PANIB_K1_MINBLOCKS)
sketch_hash_kernel(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ mask,
                   const int64_t *__restrict__ tile_off, int n_genomes, int64_t tile_begin, int64_t n_tiles,
                   const HashConsts hc, const int32_t *__restrict__ nb,
                   const uint64_t *__restrict__ bmul, uint64_t *__restrict__ table, int64_t row_stride,
                   int32_t *flags, int32_t *status, uint32_t *ticket, const SurvivorBuf surv) {
    // hc arrives as a kernel parameter (computed on the host): its 64-bit constants then sit in uniform
    // registers and feed the 3-input adds directly, instead of being folded into immediates one at a time
    using G_ = Geom<K>;
    const uint64_t max_hash = hc.max_hash;
    __shared__ __align__(16) uint32_t sp_[2][kSpStage];
    __shared__ __align__(16) uint32_t sm[2][kTileMaskWords];
    __shared__ __align__(16) uint32_t rcp_[kSpStage];
    __shared__ GenomeSlot s_slot;
    __shared__ int64_t s_range[2];  // stream tiles [lo, hi) of the genome in s_slot
    __shared__ int64_t s_next;
    __shared__ uint32_t s_nsurv;
    // per-thread ASCII scratch as pairs of words (kmer_hash.cuh: scr_index)
    extern __shared__ __align__(16) uint32_t scratch[];  // kK1DynSmem bytes
    constexpr int S = kTileBases / kCtaTile;
    n_tiles *= S;
    int64_t tile = tile_begin * S + blockIdx.x;
    if (tile >= n_tiles) {
        if (surv.cap && threadIdx.x == 0) surv.count[blockIdx.x] = 0u;
        return;
    }
    const int tid = threadIdx.x;
    if (tid == 0) s_nsurv = 0u;
    auto prefetch = [&](int64_t t, int b) {
        const uint32_t *gp = packed + t * (kCtaTile / 16);
        const uint32_t *gm = mask + t * (kCtaTile / 32);
        if (tid < kTileWords / 4) cp_async16(sp_[b] + kSpLead + 4 * tid, gp + 4 * tid);
        else if (tid < kTileWords / 4 + kTileMaskWords / 4)
            cp_async16(sm[b] + 4 * (tid - kTileWords / 4), gm + 4 * (tid - kTileWords / 4));
        cp_async_commit();
    };
    prefetch(tile, 0);
    if (tid < 2 * kSpLead) sp_[tid / kSpLead][tid % kSpLead] = 0u;  // lead-in words: read, never used
    else if (tid < 3 * kSpLead) rcp_[tid - 2 * kSpLead] = 0u;
    else if (tid < 3 * kSpLead + 2) s_range[tid - 3 * kSpLead] = 0;  // empty: the first tile looks its genome up
    uint32_t *const rcp = rcp_ + kSpLead;
    // Tiles after a CTA's first come from a ticket counter (`ticket`, zeroed by the host before the
    // launch) rather than from a fixed stride: the warp scheduler favours some resident CTAs, so with
    // equal fixed shares the favoured CTAs leave early and the SM ends the launch under-occupied.
    // Thread 0 draws the ticket for the tile after next while the current tile is expanded.
    const int64_t dyn_base = tile_begin * S + gridDim.x;
    if (ticket && tid == 0) s_next = dyn_base + atomicAdd(ticket, 1u);
    __syncthreads();
    int64_t next = ticket ? s_next : tile + gridDim.x;
    const int u = tid >> 2, a = tid & 3;
    const EmitToTable emit{&s_slot, max_hash, status, surv, &s_nsurv};
    // The prefilter threshold as a COMPUTED uniform value (blockIdx.y is always 0: the grid is 1-D): ptxas keeps
    // that in a uniform register, whereas it re-loads a plain kernel parameter from the constant bank for every
    // k-mer (one LDCU per k-mer less; with the validity test on the rare path 97 -> 95 instructions per k-mer in
    // the hot loop, measured +1.0 % at configs[1] and configs[2])
    HashConsts hc2 = hc;
    hc2.thr ^= blockIdx.y;
    int g = 0;
    for (int it = 0; tile < n_tiles; ++it) {
        const int cur = it & 1;
        cp_async_wait<0>();
        __syncthreads();  // tile `cur` has landed; every thread is done with the previous tile's scratch
        if (it) next = s_next;  // drawn during the previous tile, stored after its phase B
        if (next < n_tiles) prefetch(next, cur ^ 1);
        const uint32_t *sp = sp_[cur] + kSpLead;
        // One thread per job, and neither may hold its warp back before the barrier between the phases (the
        // other seven would wait for it): the ticket of the tile after next is only ISSUED here, its value is
        // stored after phase B; the genome slot changes once in ~20,000 CTA tiles, so the common case is a
        // range test on two shared words instead of dependent global loads.
        uint32_t drawn = 0;
        if (tid == 0) {
            if (ticket && next < n_tiles) drawn = atomicAdd(ticket, 1u);
        } else if (tid == 32) {
            const int64_t st_tile = tile / S;
            if (st_tile < s_range[0] || st_tile >= s_range[1]) {
                g = find_genome(tile_off, n_genomes, st_tile, g);
                s_slot = GenomeSlot{table + (size_t)g * row_stride, __ldg(bmul + g), flags + g, __ldg(nb + g), g};
                s_range[0] = __ldg(tile_off + g);
                s_range[1] = __ldg(tile_off + g + 1);
            }
        }
        // phase A: items 0..255 one per thread, the halo by warp 2
        tile_expand_item<K>(sp, rcp, scratch, kThreadsK1, tid);
        if ((tid >> 5) == 2) tile_expand_halo<K>(sp, rcp, scratch, kThreadsK1, tid & 31);
        const uint32_t mw = tid < kTileMaskWords ? sm[cur][tid] : 0u;
        const bool dirty = __syncthreads_or(mw != 0u) != 0;
        // phase B
        const uint32_t vmask = dirty ? thread_valid_mask<K>(sm[cur], u, a) : 0xFFFFu;
        hash_thread_kmers<K, S42>(sp, rcp, scratch + 2 * tid, u, a, vmask, hc2, emit);
        // read by everyone after the next barrier; the last read of the old value was before the barrier above
        if (tid == 0) s_next = ticket ? (next < n_tiles ? dyn_base + drawn : n_tiles) : next + gridDim.x;
        tile = next;
    }
    if (surv.cap) {
        __syncthreads();
        if (tid == 0) surv.count[blockIdx.x] = s_nsurv < surv.cap ? s_nsurv : surv.cap;
    }
}

#endif  // PANIB_K1_WARP

// the parked survivors of a K1 launch -> their genomes' bucketed rows.  grid = (chunks of a region, CTAs of K1)
__global__ void __launch_bounds__(256)
sketch_scatter_kernel(const SurvivorBuf surv, const int32_t *__restrict__ nb, const uint64_t *__restrict__ bmul,
                      uint64_t *__restrict__ table, int64_t row_stride, int32_t *flags, int32_t *status) {
    const uint32_t n = surv.count[blockIdx.y];
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const size_t e = (size_t)blockIdx.y * surv.cap + i;
        const uint32_t g = surv.genome[e];
        table_insert(table + (size_t)g * row_stride, __ldg(nb + g), __ldg(bmul + g), surv.hash[e], flags + g, status);
    }
}

// synchronous staging used by the generic kernel
template <int NWORDS, int NMASK>
__device__ __forceinline__ void stage_tile(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ mask,
                                           int64_t tile, uint32_t *sp, uint32_t *sm) {
    const int tid = threadIdx.x;
    const uint32_t *gp = packed + tile * (kTileBases / 16);
    const uint32_t *gm = mask + tile * (kTileBases / 32);
    for (int i = tid; i < NWORDS; i += blockDim.x) sp[i] = __ldg(gp + i);
    for (int i = tid; i < NMASK; i += blockDim.x) sm[i] = __ldg(gm + i);
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// K1, generic form for any 1 <= k <= 255 (run-time k): one k-mer per thread-iteration, bases read
// from the staged 2-bit tile.  Correctness path for unusual k; not tuned.
// ------------------------------------------------------------------------------------------------
constexpr int kGenWords = (kTileBases + 256) / 16 + 1;
constexpr int kGenMask = (kTileBases + 256) / 32 + 1;

__device__ __forceinline__ uint32_t base_at(const uint32_t *sp, int pos) {
    return (sp[pos >> 4] >> (2 * (pos & 15))) & 3u;
}

__global__ void __launch_bounds__(256)
sketch_hash_generic_kernel(const uint32_t *__restrict__ packed, const uint32_t *__restrict__ mask,
                           const int64_t *__restrict__ tile_off, int n_genomes, int64_t tile_begin,
                           int64_t n_tiles, int k, uint32_t seed, uint64_t max_hash,
                           const int32_t *__restrict__ nb,
                           const uint64_t *__restrict__ bmul, uint64_t *__restrict__ table, int64_t row_stride,
                           int32_t *flags, int32_t *status) {
    __shared__ uint32_t sp[kGenWords];
    __shared__ uint32_t sm[kGenMask];
    const int64_t tile = tile_begin + (int64_t)blockIdx.x + (int64_t)blockIdx.y * gridDim.x;
    if (tile >= n_tiles) return;
    stage_tile<kGenWords, kGenMask>(packed, mask, tile, sp, sm);
    const int g = find_genome(tile_off, n_genomes, tile, 0);
    EmitHash emit{table + (size_t)g * row_stride, __ldg(nb + g), __ldg(bmul + g), max_hash, flags + g, status};
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    const uint32_t lut = 0x54474341u;
    for (int pos = threadIdx.x; pos < kTileBases; pos += blockDim.x) {
        bool valid = true;
        for (int i = 0; i < k; i++) valid &= ((sm[(pos + i) >> 5] >> ((pos + i) & 31)) & 1u) == 0u;
        if (!valid) continue;
        bool fwd = true;  // lexicographic compare of the k-mer with its reverse complement
        for (int i = 0; i < k; i++) {
            const uint32_t f = base_at(sp, pos + i), r = 3u - base_at(sp, pos + k - 1 - i);
            if (f != r) { fwd = f < r; break; }
        }
        auto key_byte = [&](int i) -> uint64_t {
            const uint32_t c = fwd ? base_at(sp, pos + i) : 3u - base_at(sp, pos + k - 1 - i);
            return (lut >> (8 * c)) & 0xFFu;
        };
        uint64_t h1 = seed, h2 = seed;
        const int nblocks = k / 16;
        for (int b = 0; b < nblocks; b++) {
            uint64_t k1 = 0, k2 = 0;
            for (int t = 0; t < 8; t++) {
                k1 |= key_byte(16 * b + t) << (8 * t);
                k2 |= key_byte(16 * b + 8 + t) << (8 * t);
            }
            k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
            h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
            k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
            h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
        }
        const int tail = k & 15, tb = 16 * nblocks;
        if (tail > 8) {
            uint64_t k2 = 0;
            for (int t = 8; t < tail; t++) k2 |= key_byte(tb + t) << (8 * (t - 8));
            k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        }
        if (tail > 0) {
            uint64_t k1 = 0;
            for (int t = 0; t < (tail < 8 ? tail : 8); t++) k1 |= key_byte(tb + t) << (8 * t);
            k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        }
        h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
        h1 += h2; h2 += h1;
        h1 = fmix64(h1); h2 = fmix64(h2);
        h1 += h2;
        emit(h1);
    }
}

// ------------------------------------------------------------------------------------------------
// finalize, step 1: sort every 1024-slot bucket of every genome in place (one CTA per bucket, bitonic
// sort in 8 KB of shared memory).  Empties (all-ones) sort to the end; the number of hashes is left
// in the bucket's LAST slot (a bucket planned at half load is never full; if it is, the overflow bit
// is raised and the host re-plans).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sketch_sort_buckets_kernel(uint64_t *__restrict__ table, int64_t row_stride, int n_genomes,
                           const int32_t *__restrict__ nb, int32_t *status) {
    __shared__ uint64_t s[kBucketSlots];
    __shared__ int s_cnt;
    const int tid = threadIdx.x, b = blockIdx.x;
    for (int g = blockIdx.y; g < n_genomes; g += gridDim.y) {
        if (b >= nb[g]) continue;  // uniform per CTA
        uint64_t *bucket = table + (size_t)g * row_stride + (size_t)b * kBucketSlots;
        for (int i = tid; i < kBucketSlots; i += 256) s[i] = bucket[i];
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        for (int k = 2; k <= kBucketSlots; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < kBucketSlots / 2; t += 256) {
                    const int i = 2 * t - (t & (j - 1));  // index with bit j clear
                    const int p = i | j;
                    const bool up = (i & k) == 0;
                    const uint64_t x = s[i], y = s[p];
                    if ((x > y) == up) { s[i] = y; s[p] = x; }
                }
                __syncthreads();
            }
        }
        for (int i = tid; i < kBucketSlots; i += 256) {
            if (s[i] != kEmpty && (i == kBucketSlots - 1 || s[i + 1] == kEmpty)) s_cnt = i + 1;
        }
        __syncthreads();
        const int cnt = s_cnt;
        if (cnt >= kBucketSlots && tid == 0 && status) atomicOr(status, PANIB_ST_BUCKET_OVERFLOW);
        for (int i = tid; i < kBucketSlots - 1; i += 256) bucket[i] = s[i];
        if (tid == 0) bucket[kBucketSlots - 1] = (uint64_t)(cnt < kBucketSlots ? cnt : kBucketSlots - 1);
        __syncthreads();
    }
}

// finalize, step 2: one CTA per genome moves the sorted buckets together (in place, left to right),
// which yields the globally sorted sketch because buckets are value ranges.
__global__ void __launch_bounds__(256)
sketch_compact_kernel(uint64_t *__restrict__ table, int64_t row_stride, const int32_t *__restrict__ nb,
                      int32_t *__restrict__ counts, const int32_t *__restrict__ flags, int32_t *status) {
    const int g = blockIdx.x, tid = threadIdx.x;
    uint64_t *row = table + (size_t)g * row_stride;
    const int n = nb[g];
    int off = 0;
    for (int b = 0; b < n; b++) {
        const uint64_t *bucket = row + (size_t)b * kBucketSlots;
        const int cnt = (int)bucket[kBucketSlots - 1];
        uint64_t v[kBucketSlots / 256];
#pragma unroll
        for (int r = 0; r < kBucketSlots / 256; r++) {
            const int i = tid + 256 * r;
            v[r] = i < cnt ? bucket[i] : 0ull;
        }
        __syncthreads();  // everything is read before anything moves (source and target may overlap)
#pragma unroll
        for (int r = 0; r < kBucketSlots / 256; r++) {
            const int i = tid + 256 * r;
            if (i < cnt) row[off + i] = v[r];
        }
        off += cnt;
        __syncthreads();
    }
    if (tid == 0) {
        if ((flags[g] & 1) && off < row_stride) row[off++] = kEmpty;  // the all-ones hash (scaled == 1 only)
        counts[g] = off;
        // the size also rides in the row's last slot, so that ONE all-gather of the rows moves
        // sketches and sizes together
        if (off < row_stride) row[row_stride - 1] = (uint64_t)off;
        else if (status) atomicOr(status, PANIB_ST_BUCKET_OVERFLOW);  // a completely full row: re-plan
        if (status) atomicMax(status + 1, off);  // largest sketch so far: sizes K2's shared memory
    }
}

// finalize, step 2 fused with the all-gather of the multi-GPU path: instead of compacting a row in
// place and handing the table to NCCL, the CTA writes the genome's sorted sketch straight into the
// gathered table of EVERY rank (its own included) through peer-mapped pointers -- NVLink stores,
// posted and coalesced, 8 bytes x 256 threads per instruction.  Row (rank*per_rank + g) of every
// peer's table receives the hashes, with the size in the row's last slot; a symmetric-memory barrier
// after the kernel (multi_gpu.SymmetricGather) publishes them.  This removes the separate NCCL
// all-gather kernel and its extra pass over the sketches.
constexpr int kMaxPeers = 16;
struct PeerTables {
    uint64_t *ptr[kMaxPeers];
};

__global__ void __launch_bounds__(256)
sketch_compact_scatter_kernel(const uint64_t *__restrict__ table, int64_t row_stride,
                              const int32_t *__restrict__ nb, int32_t *__restrict__ counts,
                              const int32_t *__restrict__ flags, int32_t *status, PeerTables peers, int world,
                              int64_t first_row) {
    // grid = (genome, peer): many CTAs keep enough NVLink stores in flight (the local re-reads of the
    // sorted buckets hit L2)
    const int g = blockIdx.x, r = blockIdx.y, tid = threadIdx.x;
    const uint64_t *row = table + (size_t)g * row_stride;
    uint64_t *dst = peers.ptr[r] + (size_t)(first_row + g) * row_stride;
    const int n = nb[g];
    int off = 0;
    for (int b = 0; b < n; b++) {
        const uint64_t *bucket = row + (size_t)b * kBucketSlots;
        const int cnt = (int)bucket[kBucketSlots - 1];
#pragma unroll 4
        for (int i = tid; i < cnt; i += 256) dst[off + i] = bucket[i];
        off += cnt;
    }
    if (tid == 0) {
        if ((flags[g] & 1) && off < row_stride - 1) dst[off++] = kEmpty;  // the all-ones hash (scaled == 1 only)
        if (off < row_stride) dst[row_stride - 1] = (uint64_t)off;
        if (r == 0) {
            counts[g] = off;
            if (off >= row_stride && status) atomicOr(status, PANIB_ST_BUCKET_OVERFLOW);
            if (status) atomicMax(status + 1, off);
        }
    }
    __threadfence_system();  // order the peer stores before the kernel-end / barrier signal
}

}  // namespace panib

// ================================================================================================
// C ABI
// ================================================================================================
using namespace panib;

static dim3 tile_grid(int64_t n_tiles) {  // one CTA per tile (generic kernel); surplus CTAs exit
    const int64_t gx = n_tiles < (1 << 30) ? n_tiles : (1 << 30);
    const int64_t gy = (n_tiles + gx - 1) / gx;
    return dim3((unsigned)gx, (unsigned)gy, 1);
}

constexpr size_t kK1DynSmem = 2 * kBlkPos * kLaunchThreadsK1 * sizeof(uint32_t);  // per-thread ASCII scratch

// k-mer starts one CTA takes per round of tickets (CTA form: one CTA tile; warp form: one unit per warp)
#if PANIB_K1_WARP
constexpr int kLaunchTileK1 = kWarpsK1 * kUnitTiles * kCtaTile;
#else
constexpr int kLaunchTileK1 = kCtaTile;
#endif
// persistent grid of the fast kernel: SMs x resident CTAs per SM
template <int K, bool S42>
static int persistent_grid(int64_t n_tiles) {
    static int cached = 0;
    if (!cached) {
        int dev = 0, sms = 148, per_sm = 2;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(sketch_hash_kernel<K, S42>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kK1DynSmem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sketch_hash_kernel<K, S42>, kLaunchThreadsK1, kK1DynSmem) !=
                cudaSuccess || per_sm < 1)
            per_sm = 2;
        cached = sms * per_sm;
    }
    return (int)(n_tiles < cached ? n_tiles : cached);
}

extern "C" int panib_pack_ascii(const uint8_t *d_ascii, int64_t n_bases, uint32_t *d_packed, uint32_t *d_mask,
                                void *stream) {
    if (n_bases < 0 || (n_bases & 31)) {
        set_error("panib_pack_ascii: n_bases=%lld must be a non-negative multiple of 32", (long long)n_bases);
        return PANIB_E_ARG;
    }
    if (n_bases == 0) return PANIB_OK;
    const int64_t groups = n_bases / 32;
    const int64_t want = (groups + 255) / 256;
    const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
    pack_ascii_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4 *>(d_ascii), groups,
                                                                 d_packed, d_mask);
    return check_launch("pack_ascii_kernel");
}

// ---- survivor workspace: registered per DEVICE by the caller (panib_set_workspace), owned by the caller.
// Per device, not per stream: a CUDA graph capture runs on a stream of its own, and the captured kernels must
// find the same scratch as the eager ones.  Sketch calls of one device are serialised by their callers.
namespace {
struct Workspace {
    void *ptr;
    int64_t bytes;
};
std::mutex g_ws_mutex;
std::map<int, Workspace> g_ws;  // key: device ordinal
}  // namespace

extern "C" int panib_set_workspace(void *d_ptr, int64_t bytes) {
    int dev = 0;
    PANIB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_ws_mutex);
    if (!d_ptr || bytes <= 0) g_ws.erase(dev);
    else g_ws[dev] = Workspace{d_ptr, bytes};
    return PANIB_OK;
}

// regions of `ctas` CTAs in the stream's workspace, sized for `bases` hashed at 1/scaled survival with 2x
// head room; cap = 0 (direct inserts) when there is no workspace or it is too small to be worth it
static SurvivorBuf survivor_regions(int ctas, int64_t bases, uint64_t max_hash) {
    SurvivorBuf b{nullptr, nullptr, nullptr, 0u};
    Workspace w{nullptr, 0};
    {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(g_ws_mutex);
        auto it = g_ws.find(dev);
        if (it != g_ws.end()) w = it->second;
    }
    if (!w.ptr || ctas <= 0) return b;
    const double rate = ((double)max_hash + 1.0) / 18446744073709551616.0;  // survival probability of a k-mer
    int64_t want = (int64_t)(2.0 * rate * (double)bases / ctas) + 1024;
    const int64_t room = (w.bytes - 256 - (int64_t)ctas * 4) / ((int64_t)ctas * 12) / 2 * 2;  // even: 8-byte alignment
    if (want > room) want = room;
    if (want < 256 || want > 0x7FFFFFFF) return b;
    char *base = static_cast<char *>(w.ptr);
    b.hash = reinterpret_cast<uint64_t *>(base);
    b.genome = reinterpret_cast<uint32_t *>(base + (size_t)ctas * want * 8);
    b.count = reinterpret_cast<uint32_t *>(base + (size_t)ctas * want * 12);
    b.cap = (uint32_t)want;
    return b;
}

// launch K1 over tiles [tile_begin, tile_end) of the stream (table / flags already initialised)
static int launch_hash_range(const uint32_t *d_packed, const uint32_t *d_mask, const int64_t *d_tile_off,
                             int64_t n_genomes, int64_t tile_begin, int64_t tile_end, int k, uint32_t seed,
                             uint64_t max_hash, const int32_t *d_nb, const uint64_t *d_bmul, uint64_t *d_table,
                             int64_t row_stride, int32_t *d_flags, int32_t *d_status, cudaStream_t st) {
    const int64_t n = tile_end - tile_begin;
    if (n <= 0) return PANIB_OK;
    // ticket counter of the dynamic tile schedule: word 2 of the caller's status block (one engine =
    // one stream), zeroed in stream order before every launch; without a status block, fixed strides
    uint32_t *ticket = nullptr;
#ifndef PANIB_K1_STATIC_TILES
    if (d_status && (k == 21 || k == 31) && n * (kTileBases / kCtaTile) < 0xF0000000LL) {
        ticket = reinterpret_cast<uint32_t *>(d_status) + 2;
        PANIB_CUDA(cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st));
    }
#endif
    const HashConsts hc = make_hash_consts(seed, max_hash);
    SurvivorBuf surv{nullptr, nullptr, nullptr, 0u};
    int ctas = 0;
#define PANIB_LAUNCH_KS(KK, SS)                                                                              \
    ctas = persistent_grid<KK, SS>((n * (int64_t)kTileBases + kLaunchTileK1 - 1) / kLaunchTileK1);           \
    surv = survivor_regions(ctas, n * (int64_t)kTileBases, max_hash);                                        \
    sketch_hash_kernel<KK, SS><<<ctas, kLaunchThreadsK1, kK1DynSmem, st>>>(                                  \
        d_packed, d_mask, d_tile_off, (int)n_genomes, tile_begin, tile_end, hc, d_nb, d_bmul,                \
        d_table, row_stride, d_flags, d_status, ticket, surv)
    // sourmash's seed (42) gets the instantiation with the seed-dependent constants as immediates
#define PANIB_LAUNCH_K(KK)                                                                                   \
    if (seed == kSourmashSeed) { PANIB_LAUNCH_KS(KK, true); } else { PANIB_LAUNCH_KS(KK, false); }
    switch (k) {
    case 21: PANIB_LAUNCH_K(21); break;
    case 31: PANIB_LAUNCH_K(31); break;
    default:
        sketch_hash_generic_kernel<<<tile_grid(n), 256, 0, st>>>(d_packed, d_mask, d_tile_off, (int)n_genomes,
                                                                  tile_begin, tile_end, k, seed, max_hash, d_nb,
                                                                  d_bmul, d_table, row_stride, d_flags, d_status);
    }
#undef PANIB_LAUNCH_KS
#undef PANIB_LAUNCH_K
    int rc = check_launch("sketch_hash_kernel");
    if (rc || !surv.cap) return rc;
    const unsigned gx = (unsigned)((surv.cap + 255) / 256 < 64 ? (surv.cap + 255) / 256 : 64);
    sketch_scatter_kernel<<<dim3(gx, (unsigned)ctas), 256, 0, st>>>(surv, d_nb, d_bmul, d_table, row_stride, d_flags,
                                                                    d_status);
    return check_launch("sketch_scatter_kernel");
}

static int check_sketch_args(int k, int64_t row_stride) {
    if (k < 1 || k > PANIB_MAX_K) {
        set_error("k-mer size %d outside 1..%d", k, PANIB_MAX_K);
        return PANIB_E_ARG;
    }
    if (row_stride <= 0 || row_stride % kBucketSlots) {
        set_error("row_stride=%lld must be a positive multiple of %d", (long long)row_stride, kBucketSlots);
        return PANIB_E_ARG;
    }
    return PANIB_OK;
}

extern "C" int panib_sketch_hash_only(const uint32_t *d_packed, const uint32_t *d_mask, const int64_t *d_tile_off,
                                      int64_t n_genomes, int64_t n_tiles, int k, uint32_t seed, uint64_t max_hash,
                                      const int32_t *d_nb, const uint64_t *d_bmul, uint64_t *d_table,
                                      int64_t row_stride, int32_t *d_flags, int32_t *d_status, void *stream) {
    int rc = check_sketch_args(k, row_stride);
    if (rc) return rc;
    if (n_genomes <= 0 || n_tiles <= 0) return PANIB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PANIB_CUDA(cudaMemsetAsync(d_table, 0xFF, (size_t)n_genomes * row_stride * sizeof(uint64_t), st));
    PANIB_CUDA(cudaMemsetAsync(d_flags, 0, (size_t)n_genomes * sizeof(int32_t), st));
    return launch_hash_range(d_packed, d_mask, d_tile_off, n_genomes, 0, n_tiles, k, seed, max_hash, d_nb, d_bmul,
                             d_table, row_stride, d_flags, d_status, st);
}

extern "C" int panib_sketch_finalize(uint64_t *d_table, int64_t row_stride, int64_t n_genomes, const int32_t *d_nb,
                                     int32_t *d_counts, const int32_t *d_flags, int32_t *d_status,
                                     void *stream) {
    if (n_genomes <= 0) return PANIB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned nbmax = (unsigned)(row_stride / kBucketSlots);
    const unsigned gy = (unsigned)(n_genomes < 65535 ? n_genomes : 65535);
    sketch_sort_buckets_kernel<<<dim3(nbmax, gy), 256, 0, st>>>(d_table, row_stride, (int)n_genomes, d_nb,
                                                                 d_status ? d_status : nullptr);
    int rc = check_launch("sketch_sort_buckets_kernel");
    if (rc) return rc;
    sketch_compact_kernel<<<(unsigned)n_genomes, 256, 0, st>>>(d_table, row_stride, d_nb, d_counts, d_flags,
                                                               d_status);
    return check_launch("sketch_compact_kernel");
}

extern "C" int panib_sketch_finalize_gather(uint64_t *d_table, int64_t row_stride, int64_t n_genomes,
                                            const int32_t *d_nb, int32_t *d_counts, const int32_t *d_flags,
                                            int32_t *d_status, const uint64_t *const *h_peer_tables, int world,
                                            int rank, int64_t per_rank, void *stream) {
    if (n_genomes <= 0) return PANIB_OK;
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !h_peer_tables || n_genomes > per_rank) {
        set_error("panib_sketch_finalize_gather: bad world/rank/per_rank (%d/%d/%lld)", world, rank,
                  (long long)per_rank);
        return PANIB_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned nbmax = (unsigned)(row_stride / kBucketSlots);
    const unsigned gy = (unsigned)(n_genomes < 65535 ? n_genomes : 65535);
    sketch_sort_buckets_kernel<<<dim3(nbmax, gy), 256, 0, st>>>(d_table, row_stride, (int)n_genomes, d_nb, d_status);
    int rc = check_launch("sketch_sort_buckets_kernel");
    if (rc) return rc;
    PeerTables peers;
    for (int r = 0; r < kMaxPeers; r++) peers.ptr[r] = r < world ? const_cast<uint64_t *>(h_peer_tables[r]) : nullptr;
    for (int64_t g0 = 0; g0 < n_genomes; g0 += 65535) {  // grid.x limit is fine, keep y = peers
        const int64_t ng = n_genomes - g0 < 65535 ? n_genomes - g0 : 65535;
        sketch_compact_scatter_kernel<<<dim3((unsigned)ng, (unsigned)world), 256, 0, st>>>(
            d_table + (size_t)g0 * row_stride, row_stride, d_nb + g0, d_counts + g0, d_flags + g0, d_status, peers,
            world, (int64_t)rank * per_rank + g0);
        rc = check_launch("sketch_compact_scatter_kernel");
        if (rc) return rc;
    }
    return PANIB_OK;
}

extern "C" int panib_sketch_stream(const uint32_t *d_packed, const uint32_t *d_mask, const int64_t *d_tile_off,
                                   int64_t n_genomes, int64_t n_tiles, int k, uint32_t seed, uint64_t max_hash,
                                   const int32_t *d_nb, const uint64_t *d_bmul, uint64_t *d_table,
                                   int64_t row_stride, int32_t *d_counts, int32_t *d_flags, int32_t *d_status,
                                   void *stream) {
    int rc = panib_sketch_hash_only(d_packed, d_mask, d_tile_off, n_genomes, n_tiles, k, seed, max_hash, d_nb,
                                    d_bmul, d_table, row_stride, d_flags, d_status, stream);
    if (rc) return rc;
    return panib_sketch_finalize(d_table, row_stride, n_genomes, d_nb, d_counts, d_flags, d_status, stream);
}

// Host-buffer form.  The stream is cut into chunks; chunk c+1 is copied host->device on an internal
// copy stream while chunk c is packed and hashed on the caller's stream, so that the PCIe transfer
// (1 byte per base) and K1 overlap.  K1 of a chunk stops one tile short of the chunk's end, because
// the last tile's k-1 halo lives in the next chunk; that tile is hashed with the next chunk.
static int sketch_ascii_host_impl(const uint8_t *h_ascii, uint8_t *d_ascii, int64_t n_bases, uint32_t *d_packed,
                                  uint32_t *d_mask, const int64_t *d_tile_off, int64_t n_genomes, int64_t n_tiles,
                                  int k, uint32_t seed, uint64_t max_hash, const int32_t *d_nb,
                                  const uint64_t *d_bmul, uint64_t *d_table, int64_t row_stride, int32_t *d_counts,
                                  int32_t *d_flags, int32_t *d_status, void *stream, bool finalize) {
    if (n_bases != (n_tiles + 1) * (int64_t)kTileBases) {
        set_error("n_bases=%lld must equal (n_tiles+1)*%d", (long long)n_bases, kTileBases);
        return PANIB_E_ARG;
    }
    int rc = check_sketch_args(k, row_stride);
    if (rc) return rc;
    if (n_genomes <= 0 || n_tiles <= 0) return PANIB_OK;
    cudaStream_t st = (cudaStream_t)stream;

    constexpr int kMaxChunks = 32;
    static thread_local cudaStream_t copy_stream = nullptr;
    static thread_local cudaEvent_t ev_copied[kMaxChunks], ev_ready = nullptr;
    static thread_local int ev_device = -1;
    int dev = 0;
    PANIB_CUDA(cudaGetDevice(&dev));
    if (!copy_stream || ev_device != dev) {
        PANIB_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < kMaxChunks; i++)
            PANIB_CUDA(cudaEventCreateWithFlags(&ev_copied[i], cudaEventDisableTiming));
        PANIB_CUDA(cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming));
        ev_device = dev;
    }
    // chunks of whole tiles, >= 4 MiB each, at most kMaxChunks; the trailing pad tile rides with the last
    const int64_t total_tiles = n_tiles + 1;
    int64_t per = (total_tiles + kMaxChunks - 1) / kMaxChunks;
    if (per < 1024) per = 1024;
    const int n_chunks = (int)((total_tiles + per - 1) / per);

    PANIB_CUDA(cudaMemsetAsync(d_table, 0xFF, (size_t)n_genomes * row_stride * sizeof(uint64_t), st));
    PANIB_CUDA(cudaMemsetAsync(d_flags, 0, (size_t)n_genomes * sizeof(int32_t), st));
    // the copy stream must not overwrite d_ascii while earlier work on `st` may still read it
    PANIB_CUDA(cudaEventRecord(ev_ready, st));
    PANIB_CUDA(cudaStreamWaitEvent(copy_stream, ev_ready, 0));
    for (int c = 0; c < n_chunks; c++) {
        const int64_t t0 = c * per, t1 = (c + 1) * per < total_tiles ? (c + 1) * per : total_tiles;
        PANIB_CUDA(cudaMemcpyAsync(d_ascii + t0 * kTileBases, h_ascii + t0 * kTileBases,
                                   (size_t)(t1 - t0) * kTileBases, cudaMemcpyHostToDevice, copy_stream));
        PANIB_CUDA(cudaEventRecord(ev_copied[c], copy_stream));
    }
    int64_t hashed = 0;  // tiles [0, hashed) are done
    for (int c = 0; c < n_chunks; c++) {
        const int64_t t0 = c * per, t1 = (c + 1) * per < total_tiles ? (c + 1) * per : total_tiles;
        PANIB_CUDA(cudaStreamWaitEvent(st, ev_copied[c], 0));
        rc = panib_pack_ascii(d_ascii + t0 * kTileBases, (t1 - t0) * kTileBases, d_packed + t0 * (kTileBases / 16),
                              d_mask + t0 * (kTileBases / 32), stream);
        if (rc) return rc;
        const int64_t upto = c == n_chunks - 1 ? n_tiles : t1 - 1;  // halo of tile t1-1 is in the next chunk
        rc = launch_hash_range(d_packed, d_mask, d_tile_off, n_genomes, hashed, upto, k, seed, max_hash, d_nb,
                               d_bmul, d_table, row_stride, d_flags, d_status, st);
        if (rc) return rc;
        if (upto > hashed) hashed = upto;
    }
    if (!finalize) return PANIB_OK;
    return panib_sketch_finalize(d_table, row_stride, n_genomes, d_nb, d_counts, d_flags, d_status, stream);
}

extern "C" int panib_sketch_ascii_host(const uint8_t *h_ascii, uint8_t *d_ascii, int64_t n_bases,
                                       uint32_t *d_packed, uint32_t *d_mask, const int64_t *d_tile_off,
                                       int64_t n_genomes, int64_t n_tiles, int k, uint32_t seed, uint64_t max_hash,
                                       const int32_t *d_nb, const uint64_t *d_bmul, uint64_t *d_table,
                                       int64_t row_stride, int32_t *d_counts, int32_t *d_flags, int32_t *d_status,
                                       void *stream) {
    return sketch_ascii_host_impl(h_ascii, d_ascii, n_bases, d_packed, d_mask, d_tile_off, n_genomes, n_tiles, k,
                                  seed, max_hash, d_nb, d_bmul, d_table, row_stride, d_counts, d_flags, d_status,
                                  stream, true);
}

// same, but the rows are left as sorted-per-bucket hash sets: the caller finalizes them itself
// (panib_sketch_finalize or, on several GPUs, panib_sketch_finalize_gather)
extern "C" int panib_sketch_ascii_host_hash_only(const uint8_t *h_ascii, uint8_t *d_ascii, int64_t n_bases,
                                                 uint32_t *d_packed, uint32_t *d_mask, const int64_t *d_tile_off,
                                                 int64_t n_genomes, int64_t n_tiles, int k, uint32_t seed,
                                                 uint64_t max_hash, const int32_t *d_nb, const uint64_t *d_bmul,
                                                 uint64_t *d_table, int64_t row_stride, int32_t *d_flags,
                                                 int32_t *d_status, void *stream) {
    return sketch_ascii_host_impl(h_ascii, d_ascii, n_bases, d_packed, d_mask, d_tile_off, n_genomes, n_tiles, k,
                                  seed, max_hash, d_nb, d_bmul, d_table, row_stride, nullptr, d_flags, d_status,
                                  stream, false);
}

// ------------------------------------------------------------------------------------------------
// Packed host-buffer form: the ingest pipeline of the drop-in path.  The genomes arrive as an ASCII base
// stream in host memory (h_ascii) or already packed (h_ascii == NULL).  The stream is cut into chunks of
// whole tiles and worked on from BOTH ends:
//   * the pool of host threads (hostpack.cpp) packs chunks from the front into the pinned h_packed; a packed
//     chunk crosses PCIe as 0.25 byte per base.  The validity mask is not sent densely: the pool flags the
//     tiles that hold an invalid base ("dirty": genome ends, record separators, N runs), only their 512-byte
//     masks are copied (compacted in place in h_mask, ids behind them) and a small kernel scatters them over
//     the zeroed mask of the chunk;
//   * whenever the link would otherwise idle (the next packed chunk is not ready and at most one copy is in
//     flight) and h_ascii is page-locked, the submitting thread takes the LAST chunk no pool thread has touched
//     and sends it as ASCII (1 byte per base) into a slot of d_scratch, where the GPU packs it (K0).
// So the host cores and the link are both kept busy whatever their relative speeds (1 or 8 GPUs per host, 4 or
// 32 cores per GPU), and K1 runs on every chunk as soon as it is on the device, in arrival order: the tile at
// a chunk boundary (its k-1 halo lives in the next chunk) is hashed with whichever neighbour arrives second.
// d_counts == NULL leaves the rows as bucketed hash sets (the caller finalizes: panib_sketch_finalize or, on
// several GPUs, panib_sketch_finalize_gather).
// ------------------------------------------------------------------------------------------------
namespace panib {
struct HostPackJob;
HostPackJob *host_pack_start(const uint8_t *h_ascii, int64_t n_bases, uint32_t *h_packed, uint32_t *h_mask,
                             int64_t bases_per_block, int threads, uint8_t *tile_dirty, int64_t blocks_per_chunk,
                             int64_t ring_blocks);
bool host_pack_chunk_ready(HostPackJob *hp, int64_t c);
bool host_pack_blocks_ready(HostPackJob *hp, int64_t b0, int64_t b1);
void host_pack_release(HostPackJob *hp, int64_t blocks);
bool host_pack_claim_raw(HostPackJob *hp, int64_t c);
bool host_pack_help(HostPackJob *hp);
void host_pack_finish(HostPackJob *hp);
}  // namespace panib

constexpr int kIngestMaxChunks = 64;
constexpr int64_t kIngestBlockTiles = 64;  // packing block: 64 tiles = 256 Ki bases
constexpr int kIngestMaskSlots = 3, kIngestRawSlots = 3;
constexpr int kIngestRingEvents = 64;  // copies of ring groups in flight
constexpr int kMaskWords = kTileBases / 32;  // mask words per tile

static int64_t ingest_chunk_tiles(int64_t total_tiles) {
    // chunks of whole packing blocks, >= 512 tiles (2 Mi bases) each, at most kIngestMaxChunks
    int64_t per = (total_tiles + kIngestMaxChunks - 1) / kIngestMaxChunks;
    if (per < 512) per = 512;
    return (per + kIngestBlockTiles - 1) / kIngestBlockTiles * kIngestBlockTiles;
}

extern "C" int64_t panib_ingest_scratch_bytes(int64_t n_bases) {
    if (n_bases <= 0) return 0;
    const int64_t per = ingest_chunk_tiles(n_bases / kTileBases);
    return per * ((int64_t)kIngestMaskSlots * kMaskWords * 4 + (int64_t)kIngestRawSlots * kTileBases);
}

// what the last panib_sketch_packed_host call of this thread moved: {bytes host->device, chunks, chunks sent as
// ASCII, dirty tiles, bytes of the packed-word ring (0 = whole-stream buffer), 1 if the call was a measuring one}
static thread_local int64_t g_ingest_last[6] = {0, 0, 0, 0, 0, 0};
extern "C" int panib_ingest_last(int64_t *out6) {
    if (!out6) return PANIB_E_ARG;
    for (int i = 0; i < 6; i++) out6[i] = g_ingest_last[i];
    return PANIB_OK;
}

// staging = [nd x 128 mask words][nd tile ids]: one warp moves one dirty tile's mask to its place
__global__ void __launch_bounds__(256)
mask_scatter_kernel(const uint32_t *__restrict__ staging, int nd, uint32_t *__restrict__ mask) {
    const int w = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= nd) return;
    const uint32_t tile = staging[(size_t)nd * kMaskWords + w];
    const uint4 v = reinterpret_cast<const uint4 *>(staging + (size_t)w * kMaskWords)[lane];
    reinterpret_cast<uint4 *>(mask + (size_t)tile * kMaskWords)[lane] = v;
}

extern "C" int panib_sketch_packed_host(const uint8_t *h_ascii, uint32_t *h_packed, uint32_t *h_mask,
                                        int64_t n_bases, uint8_t *d_scratch, int64_t scratch_bytes,
                                        uint32_t *d_packed, uint32_t *d_mask,
                                        const int64_t *d_tile_off, int64_t n_genomes, int64_t n_tiles, int k,
                                        uint32_t seed, uint64_t max_hash, const int32_t *d_nb,
                                        const uint64_t *d_bmul, uint64_t *d_table, int64_t row_stride,
                                        int32_t *d_counts, int32_t *d_flags, int32_t *d_status, int host_threads,
                                        void *stream) {
    if (n_bases != (n_tiles + 1) * (int64_t)kTileBases) {
        set_error("n_bases=%lld must equal (n_tiles+1)*%d", (long long)n_bases, kTileBases);
        return PANIB_E_ARG;
    }
    if (!h_packed || !h_mask) {
        set_error("panib_sketch_packed_host: h_packed / h_mask must be (pinned) host buffers");
        return PANIB_E_ARG;
    }
    int rc = check_sketch_args(k, row_stride);
    if (rc) return rc;
    if (n_genomes <= 0 || n_tiles <= 0) return PANIB_OK;
    cudaStream_t st = (cudaStream_t)stream;

    struct Res {
        cudaStream_t copy_stream = nullptr;
        cudaEvent_t copied[kIngestMaxChunks], ready = nullptr, mask_free[kIngestMaskSlots], raw_free[kIngestRawSlots];
        cudaEvent_t ring_ev[kIngestRingEvents];
        int device = -1;
        std::vector<uint8_t> dirty;
        std::vector<uint32_t> ids;
    };
    static thread_local Res R;
    int dev = 0;
    PANIB_CUDA(cudaGetDevice(&dev));
    if (!R.copy_stream || R.device != dev) {
        PANIB_CUDA(cudaStreamCreateWithFlags(&R.copy_stream, cudaStreamNonBlocking));
        for (auto &e : R.ring_ev) PANIB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : R.copied) PANIB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : R.mask_free) PANIB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : R.raw_free) PANIB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        PANIB_CUDA(cudaEventCreateWithFlags(&R.ready, cudaEventDisableTiming));
        R.device = dev;
    }
    cudaStream_t cs = R.copy_stream;
    const int64_t total_tiles = n_tiles + 1;  // the trailing pad tile rides with the last chunk
    const int64_t per = ingest_chunk_tiles(total_tiles);
    const int n_chunks = (int)((total_tiles + per - 1) / per);

    // ---- what the caller's scratch allows
    const int64_t mask_slot = per * kMaskWords * 4, raw_slot = per * (int64_t)kTileBases;
    auto env_off = [](const char *name) {
        const char *e = getenv(name);
        return e && e[0] == '0';
    };
    bool sparse = h_ascii && d_scratch && scratch_bytes >= kIngestMaskSlots * mask_slot && !env_off("PANIB_INGEST_SPARSE");
    int raw_slots = 0;
    const char *raw_env = getenv("PANIB_INGEST_RAW");  // 0 = never, 2 = whenever a chunk is free (tests); default: idle link
    const bool raw_eager = raw_env && raw_env[0] == '2';
    if (sparse && !env_off("PANIB_INGEST_RAW")) {
        const int64_t room = (scratch_bytes - kIngestMaskSlots * mask_slot) / raw_slot;
        raw_slots = (int)(room < kIngestRawSlots ? room : kIngestRawSlots);
        if (raw_slots > 0) {  // an asynchronous copy needs page-locked memory: pageable ASCII stays with the pool
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, h_ascii) != cudaSuccess || attr.type != cudaMemoryTypeHost) raw_slots = 0;
            cudaGetLastError();
        }
    }
    uint8_t *d_mask_stage = d_scratch, *d_raw = d_scratch ? d_scratch + kIngestMaskSlots * mask_slot : nullptr;

    // Packed output as a small RING inside h_packed (16 MB) instead of the whole-stream buffer: a slot is re-written
    // while it is still in the last-level cache and the DMA engine reads it from there, so packing costs the
    // host's DRAM one read of the ASCII and nothing else (the whole-stream buffer with its streaming stores costs a
    // write and a read of the packed form on top: 1.5 instead of 1.0 byte of DRAM traffic per base, on hosts whose
    // memory system is what bounds the pipeline).  Whether the DMA engine reads cache-resident lines fast depends
    // on where the host put the process relative to the GPU's root port: measured 39.1 -> 32.3 ms per 5 Gbp on one
    // 16-core box and 40.4 -> 48 ms on another.  So the library MEASURES: per thread and stream size the first
    // call warms up, the second times the whole-stream form, the third the ring, and the faster one is kept
    // (the sketches do not depend on it).  PANIB_INGEST_RING_MB pins it: 0 = never, N = always, N MB.
    const int64_t block_words = kIngestBlockTiles * (kTileBases / 16);
    const int64_t total_blocks = (total_tiles + kIngestBlockTiles - 1) / kIngestBlockTiles;
    struct Tune {
        int64_t key = -1;
        int calls = 0;
        double seconds[2] = {0.0, 0.0};  // [0] whole-stream form, [1] ring
    };
    static thread_local Tune tune;
    int64_t ring_blocks = 0, group_blocks = 1;
    int tune_slot = -1;  // >= 0: this call is a measurement of that form
    // in ring mode the submitting thread frees the slots and never packs: it needs at least one pool worker
    const int pool_threads = panib_host_threads();
    const bool has_workers = (host_threads > 0 && host_threads < pool_threads ? host_threads : pool_threads) >= 2;
    if (h_ascii && has_workers) {
        const char *e = getenv("PANIB_INGEST_RING_MB");
        int64_t mb = e ? atoll(e) : 16;
        if (!e) {  // no wish: measure, then keep the faster form (streams below 32 rings: not worth measuring)
            if (total_blocks * block_words * 4 < 32 * (mb << 20)) mb = 0;
            else {
                if (tune.key != total_blocks) tune = Tune{total_blocks, 0, {0.0, 0.0}};
                const int call = tune.calls++;
                if (call == 0) mb = 0;
                else if (call == 1) { mb = 0; tune_slot = 0; }
                else if (call == 2) tune_slot = 1;
                else if (!(tune.seconds[1] < 0.97 * tune.seconds[0])) mb = 0;
            }
        }
        ring_blocks = (mb << 20) / (block_words * 4);
        group_blocks = ring_blocks / 4 < 16 ? ring_blocks / 4 : 16;
        if (group_blocks < 1) group_blocks = 1;
        ring_blocks = ring_blocks / group_blocks * group_blocks;
        if (ring_blocks < 2 * group_blocks || ring_blocks >= total_blocks) ring_blocks = 0;  // not worth it / fits anyway
    }
    const auto tune_t0 = std::chrono::steady_clock::now();
    HostPackJob *job = nullptr;
    if (h_ascii) {
        if (sparse) R.dirty.assign((size_t)total_tiles, 0);
        job = host_pack_start(h_ascii, n_bases, h_packed, h_mask, kIngestBlockTiles * kTileBases, host_threads,
                              sparse ? R.dirty.data() : nullptr, per / kIngestBlockTiles, ring_blocks);
    }
    auto fail = [&](int code) {
        if (job) host_pack_finish(job);
        return code;
    };
#define PANIB_CUDA_JOB(call)                                                      \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            panib::set_error("%s failed: %s", #call, cudaGetErrorString(e__));   \
            return fail(PANIB_E_CUDA);                                            \
        }                                                                         \
    } while (0)
    PANIB_CUDA_JOB(cudaMemsetAsync(d_table, 0xFF, (size_t)n_genomes * row_stride * sizeof(uint64_t), st));
    PANIB_CUDA_JOB(cudaMemsetAsync(d_flags, 0, (size_t)n_genomes * sizeof(int32_t), st));
    // the copy stream must not overwrite d_packed / d_mask while earlier work on `st` may still read them
    PANIB_CUDA_JOB(cudaEventRecord(R.ready, st));
    PANIB_CUDA_JOB(cudaStreamWaitEvent(cs, R.ready, 0));

    int64_t moved[4] = {0, n_chunks, 0, 0};
    uint8_t arrived[kIngestMaxChunks] = {0};
    int inflight[kIngestMaxChunks], n_inflight = 0;  // chunks whose copies are enqueued and not yet seen complete
    int mask_used = 0, raw_used = 0;                 // slots handed out so far (ring position = count % slots)
    // K1 over the tiles of chunk c that have everything they need on the device (stream order on `st`)
    auto hash_chunk = [&](int c) -> int {
        const int64_t t0 = c * per, t1 = (c + 1) * per < total_tiles ? (c + 1) * per : total_tiles;
        arrived[c] = 1;
        const int64_t from = t0 - ((c > 0 && arrived[c - 1]) ? 1 : 0);
        int64_t upto = (c == n_chunks - 1 || arrived[c + 1]) ? t1 : t1 - 1;  // halo of tile t1-1 is in the next chunk
        if (upto > n_tiles) upto = n_tiles;
        if (upto <= from) return PANIB_OK;
        return launch_hash_range(d_packed, d_mask, d_tile_off, n_genomes, from, upto, k, seed, max_hash, d_nb, d_bmul,
                                 d_table, row_stride, d_flags, d_status, st);
    };
    auto send_packed = [&](int c, bool words_sent = false) -> int {
        const int64_t t0 = c * per, t1 = (c + 1) * per < total_tiles ? (c + 1) * per : total_tiles, n = t1 - t0;
        if (!words_sent) {  // (ring mode has sent them group by group)
            PANIB_CUDA_JOB(cudaMemcpyAsync(d_packed + t0 * (kTileBases / 16), h_packed + t0 * (kTileBases / 16),
                                           (size_t)n * (kTileBases / 4), cudaMemcpyHostToDevice, cs));
            moved[0] += n * (kTileBases / 4);
        }
        uint32_t *hm = h_mask + t0 * kMaskWords, *dm = d_mask + t0 * kMaskWords;
        int nd = 0, slot = -1;
        bool scatter = false;
        if (!job || !sparse) {  // dense mask (already packed input, or no scratch for the sparse form)
            PANIB_CUDA_JOB(cudaMemcpyAsync(dm, hm, (size_t)n * kMaskWords * 4, cudaMemcpyHostToDevice, cs));
            moved[0] += n * kMaskWords * 4;
        } else {
            const uint8_t *dirty = R.dirty.data() + t0;
            for (int64_t t = 0; t < n; t++) nd += dirty[t];
            if ((int64_t)nd * (kMaskWords + 1) > n * kMaskWords) {  // (nearly) every tile: the region IS the dense mask
                for (int64_t t = 0; t < n; t++)
                    if (!dirty[t]) memset(hm + t * kMaskWords, 0, kMaskWords * 4);
                PANIB_CUDA_JOB(cudaMemcpyAsync(dm, hm, (size_t)n * kMaskWords * 4, cudaMemcpyHostToDevice, cs));
            } else {
                PANIB_CUDA_JOB(cudaMemsetAsync(dm, 0, (size_t)n * kMaskWords * 4, cs));
                if (nd) {  // compact the dirty tiles' masks to the front of the region, their ids behind them
                    R.ids.clear();
                    int64_t at = 0;
                    for (int64_t t = 0; t < n; t++) {
                        if (!dirty[t]) continue;
                        if (t != at) memmove(hm + at * kMaskWords, hm + t * kMaskWords, kMaskWords * 4);
                        R.ids.push_back((uint32_t)(t0 + t));
                        at++;
                    }
                    memcpy(hm + (int64_t)nd * kMaskWords, R.ids.data(), (size_t)nd * 4);
                    slot = mask_used % kIngestMaskSlots;
                    if (mask_used >= kIngestMaskSlots) PANIB_CUDA_JOB(cudaStreamWaitEvent(cs, R.mask_free[slot], 0));
                    mask_used++;
                    PANIB_CUDA_JOB(cudaMemcpyAsync(d_mask_stage + slot * mask_slot, hm, (size_t)nd * (kMaskWords + 1) * 4,
                                                   cudaMemcpyHostToDevice, cs));
                    moved[0] += (int64_t)nd * (kMaskWords + 1) * 4;
                    moved[3] += nd;
                    scatter = true;
                }
            }
        }
        PANIB_CUDA_JOB(cudaEventRecord(R.copied[c], cs));
        PANIB_CUDA_JOB(cudaStreamWaitEvent(st, R.copied[c], 0));
        if (scatter) {
            mask_scatter_kernel<<<(unsigned)((nd + 7) / 8), 256, 0, st>>>(
                reinterpret_cast<const uint32_t *>(d_mask_stage + slot * mask_slot), nd, d_mask);
            const int rc2 = check_launch("mask_scatter_kernel");
            if (rc2) return fail(rc2);
            PANIB_CUDA_JOB(cudaEventRecord(R.mask_free[slot], st));
        }
        inflight[n_inflight++] = c;
        const int rc2 = hash_chunk(c);
        return rc2 ? fail(rc2) : PANIB_OK;
    };
    auto send_raw = [&](int c) -> int {
        const int64_t t0 = c * per, t1 = (c + 1) * per < total_tiles ? (c + 1) * per : total_tiles, n = t1 - t0;
        const int slot = raw_used % raw_slots;
        if (raw_used >= raw_slots) PANIB_CUDA_JOB(cudaStreamWaitEvent(cs, R.raw_free[slot], 0));
        raw_used++;
        uint8_t *dst = d_raw + slot * raw_slot;
        PANIB_CUDA_JOB(cudaMemcpyAsync(dst, h_ascii + t0 * kTileBases, (size_t)n * kTileBases, cudaMemcpyHostToDevice, cs));
        moved[0] += n * kTileBases;
        moved[2] += 1;
        PANIB_CUDA_JOB(cudaEventRecord(R.copied[c], cs));
        PANIB_CUDA_JOB(cudaStreamWaitEvent(st, R.copied[c], 0));
        int rc2 = panib_pack_ascii(dst, n * kTileBases, d_packed + t0 * (kTileBases / 16), d_mask + t0 * kMaskWords, stream);
        if (rc2) return fail(rc2);
        PANIB_CUDA_JOB(cudaEventRecord(R.raw_free[slot], st));
        inflight[n_inflight++] = c;
        rc2 = hash_chunk(c);
        return rc2 ? fail(rc2) : PANIB_OK;
    };

    int head = 0, tail = n_chunks;  // chunks [head, tail) have not been sent yet
    if (job && ring_blocks > 0) {
        // ---- ring mode: packed words leave group by group as soon as their blocks are packed; a group's ring slots
        // are handed back to the pool when its copy has completed
        const int64_t bpc = per / kIngestBlockTiles;
        struct Pending { int ev; int64_t end_block; } pend[kIngestRingEvents];
        int pend_at = 0, pend_n = 0, ev_next = 0;
        int64_t sent_block = 0;
        auto drain = [&](bool wait_oldest) {  // completed copies -> free ring slots
            bool any = false;
            while (pend_n > 0) {
                cudaError_t q = cudaEventQuery(R.ring_ev[pend[pend_at].ev]);
                if (q == cudaErrorNotReady) {
                    if (!wait_oldest) break;
                    q = cudaEventSynchronize(R.ring_ev[pend[pend_at].ev]);
                }
                if (q != cudaSuccess) break;
                wait_oldest = false;
                if (pend[pend_at].end_block >= 0) host_pack_release(job, pend[pend_at].end_block);
                pend_at = (pend_at + 1) % kIngestRingEvents;
                pend_n--;
                any = true;
            }
            cudaGetLastError();  // cudaErrorNotReady is not an error
            return any;
        };
        constexpr int64_t kRawPiece = 4 << 20;
        int raw_c = -1, raw_slot_at = 0;
        int64_t raw_off = 0;
        auto raw_piece = [&]() -> int {  // next piece of raw chunk raw_c; the last one hands the chunk to K0 + K1
            const int64_t t0 = raw_c * per, t1 = (raw_c + 1) * per < total_tiles ? (raw_c + 1) * per : total_tiles;
            const int64_t bytes = (t1 - t0) * kTileBases;
            uint8_t *dst = d_raw + raw_slot_at * raw_slot;
            const int64_t piece = bytes - raw_off < kRawPiece ? bytes - raw_off : kRawPiece;
            if (pend_n == kIngestRingEvents) drain(true);
            PANIB_CUDA_JOB(cudaMemcpyAsync(dst + raw_off, h_ascii + t0 * kTileBases + raw_off, (size_t)piece,
                                           cudaMemcpyHostToDevice, cs));
            moved[0] += piece;
            raw_off += piece;
            PANIB_CUDA_JOB(cudaEventRecord(R.ring_ev[ev_next], cs));  // keeps pend_n honest about the link; frees nothing
            pend[(pend_at + pend_n) % kIngestRingEvents] = Pending{ev_next, -1};
            pend_n++;
            ev_next = (ev_next + 1) % kIngestRingEvents;
            if (raw_off < bytes) return PANIB_OK;
            PANIB_CUDA_JOB(cudaEventRecord(R.copied[raw_c], cs));
            PANIB_CUDA_JOB(cudaStreamWaitEvent(st, R.copied[raw_c], 0));
            int rc2 = panib_pack_ascii(dst, bytes, d_packed + t0 * (kTileBases / 16), d_mask + t0 * kMaskWords, stream);
            if (rc2) return fail(rc2);
            PANIB_CUDA_JOB(cudaEventRecord(R.raw_free[raw_slot_at], st));
            rc2 = hash_chunk(raw_c);
            raw_c = -1;
            return rc2 ? fail(rc2) : PANIB_OK;
        };
        while (head < tail) {
            bool progress = drain(false);
            const int64_t cb1 = (head + 1) * bpc < total_blocks ? (head + 1) * bpc : total_blocks;
            while (sent_block < cb1) {
                int64_t g1 = (sent_block / group_blocks + 1) * group_blocks;  // groups never wrap around the ring
                if (g1 > cb1) g1 = cb1;
                if (!host_pack_blocks_ready(job, sent_block, g1)) break;
                if (pend_n == kIngestRingEvents) drain(true);
                const int64_t w0 = sent_block * block_words;
                int64_t w1 = g1 * block_words;
                if (w1 > total_tiles * (kTileBases / 16)) w1 = total_tiles * (kTileBases / 16);
                PANIB_CUDA_JOB(cudaMemcpyAsync(d_packed + w0, h_packed + (sent_block % ring_blocks) * block_words,
                                               (size_t)(w1 - w0) * 4, cudaMemcpyHostToDevice, cs));
                moved[0] += (w1 - w0) * 4;
                PANIB_CUDA_JOB(cudaEventRecord(R.ring_ev[ev_next], cs));
                pend[(pend_at + pend_n) % kIngestRingEvents] = Pending{ev_next, g1};
                pend_n++;
                ev_next = (ev_next + 1) % kIngestRingEvents;
                sent_block = g1;
                progress = true;
            }
            if (sent_block >= cb1) {  // every word of chunk `head` is on its way: mask, then K1
                rc = send_packed(head++, true);
                if (rc) return rc;
                continue;
            }
            if (raw_slots > 0 && (pend_n == 0 || raw_eager)) {
                // The link is idle: give it a PIECE of a raw chunk (4 MB of ASCII, on the same stream, so that it
                // never competes with a group of the ring for longer than that; whole raw chunks on a stream of
                // their own halved the ring's share of the link and stalled the packers).
                if (raw_c < 0) {
                    if (tail - 1 > head && host_pack_claim_raw(job, tail - 1)) {
                        raw_c = --tail;
                        raw_off = 0;
                        raw_slot_at = raw_used % raw_slots;
                        if (raw_used >= raw_slots) PANIB_CUDA_JOB(cudaStreamWaitEvent(cs, R.raw_free[raw_slot_at], 0));
                        raw_used++;
                        moved[2] += 1;
                    } else {
                        raw_slots = 0;  // the pool has reached the tail: everything left is being packed
                    }
                }
                if (raw_c >= 0) {
                    rc = raw_piece();
                    if (rc) return rc;
                    progress = true;
                }
            }
            if (!progress) std::this_thread::yield();
        }
        while (raw_c >= 0) {  // the rest of a raw chunk that was under way when the packed chunks ran out
            rc = raw_piece();
            if (rc) return rc;
        }
        host_pack_release(job, total_blocks + ring_blocks);
    }
    while (head < tail) {
        if (!job || host_pack_chunk_ready(job, head)) {
            rc = send_packed(head++);
            if (rc) return rc;
            continue;
        }
        if (raw_slots > 0) {  // the next packed chunk is not ready: is the link about to idle?
            while (n_inflight > 0 && cudaEventQuery(R.copied[inflight[0]]) == cudaSuccess) {
                for (int i = 1; i < n_inflight; i++) inflight[i - 1] = inflight[i];
                n_inflight--;
            }
            cudaGetLastError();  // cudaErrorNotReady is not an error
            if (n_inflight <= 1 || raw_eager) {
                if (tail - 1 > head && host_pack_claim_raw(job, tail - 1)) {
                    rc = send_raw(--tail);
                    if (rc) return rc;
                    continue;
                }
                raw_slots = 0;  // the pool has reached the tail: everything left is being packed
            }
        }
        if (!host_pack_help(job)) std::this_thread::yield();
    }
    if (tune_slot >= 0) {  // a measuring call: until the last byte is on the device
        PANIB_CUDA_JOB(cudaEventRecord(R.ready, cs));
        PANIB_CUDA_JOB(cudaEventSynchronize(R.ready));
        tune.seconds[tune_slot] = std::chrono::duration<double>(std::chrono::steady_clock::now() - tune_t0).count();
    }
#undef PANIB_CUDA_JOB
    if (job) host_pack_finish(job);
    for (int i = 0; i < 4; i++) g_ingest_last[i] = moved[i];
    g_ingest_last[4] = ring_blocks * block_words * 4;
    g_ingest_last[5] = tune_slot >= 0 ? 1 : 0;
    if (!d_counts) return PANIB_OK;
    return panib_sketch_finalize(d_table, row_stride, n_genomes, d_nb, d_counts, d_flags, d_status, stream);
}
