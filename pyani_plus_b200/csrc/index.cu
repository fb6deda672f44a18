// index.cu -- kernel K2, inverted-index form ("which genomes hold this hash?").
//
// The probing kernel in pairwise.cu touches every element of both sketches for every pair: its cost is
// N^2 * |sketch| whatever the genomes share.  This form costs what the genomes SHARE.  A sketch is a
// duplicate-free set, so "how many genomes hold hash h" is a group-by over all (hash, genome) entries, and
// every group of m genomes adds +1 to ov[i][j] for each of its m(m-1)/2 pairs.  The group-by is a hash table
// in global memory, hand-written (no sort, no library call):
//
//   1. insert   every entry looks its hash up in an open-addressing table (atomicCAS on the key) and takes a
//               ticket from the slot's counter (atomicAdd): it now knows its group (slot) and its arrival
//               index in the group.  Sketch hashes are uniform, so the slot is an ORDER-PRESERVING function
//               of the hash, and the grid walks "same chunk of every genome" before the next chunk: blocks in
//               flight touch one window of the table, which stays in L2.
//   2. classify the first arrival of each group reads the final count m: m >= tau -> the hash becomes a column
//               of a genome x column BIT MATRIX; 2 <= m < tau -> it gets room for its member list; m == 1
//               contributes nothing.  Columns and list space come from warp-aggregated atomic counters.
//   3. emit     every entry sets its bit (frequent) or writes its genome into its group's list (rare);
//   4. count    rare groups are expanded pair by pair with atomicAdd into ov (<= tau - 1 adds per entry);
//               frequent hashes: ov[i][j] += popcount(row_i AND row_j) for all pairs by a tiled AND+POPC
//               kernel (64 x 64 genome tiles, 4 x 4 pairs per thread, 32-word chunks in shared memory);
//   5. mirror   lower triangle := upper, diagonal := sketch sizes.
//
// Multi-GPU: the HASH RANGE is sharded -- rank r inserts only the entries whose hash falls into its 1/world
// slice of [0, max_hash], so table, lists, bit-matrix columns and pair expansion all shrink with the number
// of ranks; every rank runs all genome tiles over its own columns and the ranks' matrices sum to the result.
//
// Entries are indexed exactly (offsets = prefix sums of the sketch sizes, computed by the caller) or padded
// (`cap` slots per genome: no host knowledge of the sizes, the form a captured CUDA graph uses).
//
// Results are identical to the probing kernel (exact integer counts; additions commute); which form is
// cheaper depends on the data, so the host chooses (engine.py: Engine.intersect(method="auto")) from the
// statistics panib_index_build returns.  Replaces the same reference step as pairwise.cu: the external
// `sourmash scripts manysearch` call in pyani_plus/methods/sourmash.py:184-200.
#include "common.cuh"

namespace panib {

constexpr int kIdxThreads = 256;
constexpr int kDenseTile = 64;    // genomes per tile side of the AND+POPC kernel
constexpr int kDenseChunk = 32;   // bit-matrix words staged per step
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
constexpr uint32_t kFrequent = 0x80000000u;  // taux flag: the low bits are a bit-matrix column

struct IndexShard {
    uint64_t wmul;   // owner(h) = min(world-1, mulhi(h, wmul))
    uint64_t smul;   // slot(h)  = mulhi(h, smul) & mask: order-preserving inside a rank's hash range
    uint32_t mask;   // table slots - 1
    int rank, world;
};

__device__ __forceinline__ int64_t entry_index(int g, int i, int cap, const int64_t *__restrict__ offsets) {
    return offsets ? offsets[g] + i : (int64_t)g * cap + i;
}

// step 1.  grid = (genomes, chunks of 256 sketch positions): x (fast) walks the genomes of one chunk.
__global__ void __launch_bounds__(kIdxThreads)
index_insert_kernel(const uint64_t *__restrict__ rows, const int32_t *__restrict__ counts, int64_t stride, int n,
                    int cap, const int64_t *__restrict__ offsets, int64_t total, uint64_t max_hash, IndexShard sh,
                    unsigned long long *__restrict__ tkeys, uint32_t *__restrict__ tcount,
                    uint32_t *__restrict__ eslot, uint32_t *__restrict__ eidx, int32_t *status) {
    for (int g = blockIdx.x; g < n; g += gridDim.x) {
        const int c = counts[g];
        if (blockIdx.y == 0 && threadIdx.x == 0 && status) {
            const int64_t at = entry_index(g, 0, cap, offsets);
            if ((!offsets && c > cap) || at < 0 || at + (offsets ? c : 0) > total)
                atomicOr(status, PANIB_ST_SEGMENT_OVERFLOW);  // cap too small / offsets do not match the sizes
        }
        for (int i = blockIdx.y * kIdxThreads + threadIdx.x; i < c; i += gridDim.y * kIdxThreads) {
            if (!offsets && i >= cap) break;
            const int64_t e = entry_index(g, i, cap, offsets);
            if (e < 0 || e >= total) continue;
            const uint64_t h = rows[(size_t)g * stride + i];
            uint32_t slot = kNoSlot, arrival = 0;
            int owner = 0;
            if (sh.world > 1) {
                owner = (int)__umul64hi(h, sh.wmul);
                if (owner >= sh.world) owner = sh.world - 1;
            }
            if (h <= max_hash && owner == sh.rank) {
                uint32_t s = (uint32_t)__umul64hi(h, sh.smul) & sh.mask;
                for (uint32_t probe = 0; probe <= sh.mask; probe++) {
                    const unsigned long long prev = atomicCAS(&tkeys[s], (unsigned long long)kEmpty, (unsigned long long)h);
                    if (prev == kEmpty || prev == h) { slot = s; break; }
                    s = (s + 1) & sh.mask;
                }
                if (slot == kNoSlot) { if (status) atomicOr(status, PANIB_ST_INDEX_OVERFLOW); }
                else arrival = atomicAdd(&tcount[slot], 1u);
            }
            eslot[e] = slot;
            eidx[e] = arrival;
        }
    }
}

// step 2: one thread per entry; the group's first arrival classifies the group.
// stats: [0] columns, [1] pairs the rare groups expand to, [2] distinct hashes, [3] list words handed out.
// Columns and list space are handed out per CTA round (warp scans, then one scan over the 8 warp totals and
// ONE atomic per counter), the two pure statistics once per CTA: a single address takes an atomic every
// nanosecond or two, and there are millions of entries.
__global__ void __launch_bounds__(kIdxThreads)
index_classify_kernel(const int32_t *__restrict__ counts, int n, int cap, const int64_t *__restrict__ offsets,
                      int64_t total, int tau, const uint32_t *__restrict__ tcount, uint32_t *__restrict__ taux,
                      const uint32_t *__restrict__ eslot, const uint32_t *__restrict__ eidx,
                      unsigned long long *__restrict__ stats) {
    constexpr int kWarps = kIdxThreads / 32;
    __shared__ uint32_t s_freq[kWarps], s_rare[kWarps];
    __shared__ unsigned long long s_cbase, s_lbase;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long pairs = 0ull, distinct = 0ull;  // per thread, reduced once at the end
    for (int g = blockIdx.x; g < n; g += gridDim.x) {
        const int c = offsets ? counts[g] : min(counts[g], cap);
        const int rounds = (c + (int)(gridDim.y * kIdxThreads) - 1) / (int)(gridDim.y * kIdxThreads);
        for (int r = 0; r < rounds; r++) {  // whole CTAs stay in the loop: the barriers below need them
            const int i = (r * gridDim.y + blockIdx.y) * kIdxThreads + threadIdx.x;
            uint32_t slot = kNoSlot, m = 0;
            if (i < c) {
                const int64_t e = entry_index(g, i, cap, offsets);
                if (e >= 0 && e < total && eidx[e] == 0) slot = eslot[e];
                if (slot != kNoSlot) m = tcount[slot];
            }
            const bool frequent = m >= (uint32_t)tau, rare = m >= 2 && !frequent;
            const unsigned fmask = __ballot_sync(0xFFFFFFFFu, frequent);
            uint32_t incl = rare ? m : 0u;  // warp-wide inclusive scan of the rare group sizes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= (unsigned)o) incl += v;
            }
            if (lane == 31) {
                s_freq[warp] = (uint32_t)__popc(fmask);
                s_rare[warp] = incl;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t f = 0, w = 0;
#pragma unroll
                for (int k = 0; k < kWarps; k++) {  // exclusive scan of the warp totals, in place
                    const uint32_t fk = s_freq[k], wk = s_rare[k];
                    s_freq[k] = f;
                    s_rare[k] = w;
                    f += fk;
                    w += wk;
                }
                s_cbase = f ? atomicAdd(stats + 0, (unsigned long long)f) : 0ull;
                s_lbase = w ? atomicAdd(stats + 3, (unsigned long long)w) : 0ull;
            }
            __syncthreads();
            if (frequent)
                taux[slot] = kFrequent | (uint32_t)(s_cbase + s_freq[warp] + __popc(fmask & ((1u << lane) - 1u)));
            else if (rare) taux[slot] = (uint32_t)(s_lbase + s_rare[warp] + incl - m);
            if (rare) pairs += (unsigned long long)m * (m - 1) / 2ull;
            if (slot != kNoSlot) distinct += 1ull;
            __syncthreads();  // s_* are rewritten in the next round
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        pairs += __shfl_xor_sync(0xFFFFFFFFu, pairs, o);
        distinct += __shfl_xor_sync(0xFFFFFFFFu, distinct, o);
    }
    if (lane == 0) {
        if (pairs) atomicAdd(stats + 1, pairs);
        if (distinct) atomicAdd(stats + 2, distinct);
    }
}

// step 3: bits of the frequent hashes, member lists of the rare ones
__global__ void __launch_bounds__(kIdxThreads)
index_emit_kernel(const int32_t *__restrict__ counts, int n, int cap, const int64_t *__restrict__ offsets,
                  int64_t total, int tau, const uint32_t *__restrict__ tcount, const uint32_t *__restrict__ taux,
                  const uint32_t *__restrict__ eslot, const uint32_t *__restrict__ eidx,
                  uint32_t *__restrict__ lists, int64_t list_cap, uint32_t *__restrict__ bits, int64_t wcap,
                  int32_t *status) {
    for (int g = blockIdx.x; g < n; g += gridDim.x) {
        const int c = offsets ? counts[g] : min(counts[g], cap);
        for (int i = blockIdx.y * kIdxThreads + threadIdx.x; i < c; i += gridDim.y * kIdxThreads) {
            const int64_t e = entry_index(g, i, cap, offsets);
            if (e < 0 || e >= total) continue;
            const uint32_t slot = eslot[e];
            if (slot == kNoSlot) continue;
            const uint32_t m = tcount[slot];
            if (m < 2) continue;
            const uint32_t aux = taux[slot];
            if (m >= (uint32_t)tau) {
                const uint32_t col = aux & ~kFrequent;
                if ((int64_t)(col >> 5) >= wcap) { if (status) atomicOr(status, PANIB_ST_SEGMENT_OVERFLOW); continue; }
                atomicOr(bits + (size_t)g * wcap + (col >> 5), 1u << (col & 31));
            } else {
                const int64_t at = (int64_t)aux + eidx[e];
                if (at < list_cap) lists[at] = (uint32_t)g;
                else if (status) atomicOr(status, PANIB_ST_SEGMENT_OVERFLOW);
            }
        }
    }
}

// step 4a: rare hashes -- every entry pairs with the LATER arrivals of its group, so each unordered pair is
// written once; (min, max) puts it in the upper triangle whatever the arrival order was
__global__ void __launch_bounds__(kIdxThreads)
index_sparse_kernel(const int32_t *__restrict__ counts, int n, int cap, const int64_t *__restrict__ offsets,
                    int64_t total, int tau, const uint32_t *__restrict__ tcount, const uint32_t *__restrict__ taux,
                    const uint32_t *__restrict__ eslot, const uint32_t *__restrict__ eidx,
                    const uint32_t *__restrict__ lists, int64_t list_cap, uint32_t *__restrict__ ov, int64_t ld) {
    for (int g = blockIdx.x; g < n; g += gridDim.x) {
        const int c = offsets ? counts[g] : min(counts[g], cap);
        for (int i = blockIdx.y * kIdxThreads + threadIdx.x; i < c; i += gridDim.y * kIdxThreads) {
            const int64_t e = entry_index(g, i, cap, offsets);
            if (e < 0 || e >= total) continue;
            const uint32_t slot = eslot[e];
            if (slot == kNoSlot) continue;
            const uint32_t m = tcount[slot];
            if (m < 2 || m >= (uint32_t)tau) continue;
            const int64_t base = (int64_t)taux[slot];
            for (uint32_t b = eidx[e] + 1; b < m; b++) {
                if (base + b >= list_cap) break;
                const uint32_t other = lists[base + b];
                if (other >= (uint32_t)n || other == (uint32_t)g) continue;
                const uint32_t lo = other < (uint32_t)g ? other : (uint32_t)g;
                const uint32_t hi = other < (uint32_t)g ? (uint32_t)g : other;
                atomicAdd(ov + (size_t)lo * ld + hi, 1u);
            }
        }
    }
}

// frequent hashes: ov[i][j] += popcount(bits[i] & bits[j]) over the used words, for i < j.
// One CTA per 64 x 64 tile on or above the diagonal; thread (ty, tx) owns pairs (ty + 16p, tx + 16q).
__global__ void __launch_bounds__(256)
index_dense_kernel(const uint32_t *__restrict__ bits, int64_t wcap, const unsigned long long *__restrict__ stats,
                   int n, uint32_t *__restrict__ ov, int64_t ld, int rank, int world) {
    const int I = blockIdx.y, J = blockIdx.x;
    if (I > J) return;
    (void)rank; (void)world;  // hash-range sharding: every rank runs all tiles over its OWN columns
    int64_t W = (int64_t)((stats[0] + 31ull) >> 5);
    if (W > wcap) W = wcap;
    __shared__ uint32_t sa[kDenseChunk][kDenseTile + 1];
    __shared__ uint32_t sb[kDenseChunk][kDenseTile + 1];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int lw = t & 31, lr = t >> 5;  // loader: word lw of rows lr + 8q
    uint32_t acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int q = 0; q < 4; q++) acc[p][q] = 0u;
    for (int64_t w0 = 0; w0 < W; w0 += kDenseChunk) {
        const int64_t w = w0 + lw;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int r = lr + 8 * q;
            const int gi = I * kDenseTile + r, gj = J * kDenseTile + r;
            sa[lw][r] = (gi < n && w < W) ? bits[(size_t)gi * wcap + w] : 0u;
            sb[lw][r] = (gj < n && w < W) ? bits[(size_t)gj * wcap + w] : 0u;
        }
        __syncthreads();
#pragma unroll 8
        for (int x = 0; x < kDenseChunk; x++) {
            uint32_t a[4], b[4];
#pragma unroll
            for (int p = 0; p < 4; p++) {
                a[p] = sa[x][ty + 16 * p];
                b[p] = sb[x][tx + 16 * p];
            }
#pragma unroll
            for (int p = 0; p < 4; p++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[p][q] += (uint32_t)__popc(a[p] & b[q]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = I * kDenseTile + ty + 16 * p, j = J * kDenseTile + tx + 16 * q;
            if (i < n && j < n && i < j && acc[p][q]) ov[(size_t)i * ld + j] += acc[p][q];
        }
}

// lower triangle := upper triangle (32 x 32 tiles through shared memory), diagonal := sketch sizes
__global__ void __launch_bounds__(256)
index_mirror_kernel(uint32_t *__restrict__ ov, int64_t ld, int n, const int32_t *__restrict__ counts,
                    int write_diag) {
    const int bx = blockIdx.x, by = blockIdx.y;  // tile (rows by, cols bx) of the upper triangle
    if (bx < by) return;
    __shared__ uint32_t tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
    for (int r = ty; r < 32; r += 8) {
        const int i = by * 32 + r, j = bx * 32 + tx;
        tile[r][tx] = (i < n && j < n && i < j) ? ov[(size_t)i * ld + j] : 0u;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = bx * 32 + r, i = by * 32 + tx;  // writes ov[j][i] = upper[i][j]
        if (i < n && j < n) {
            if (i < j) ov[(size_t)j * ld + i] = tile[tx][r];
            else if (i == j && write_diag) ov[(size_t)i * ld + i] = (uint32_t)counts[i];
        }
    }
}

struct IndexLayout {
    int64_t total, wcap, slots;
    size_t off_keys, off_count, off_aux, off_eslot, off_eidx, off_lists, off_bits, bytes;
};

static int index_layout(int64_t n, int64_t total, int tau, int world, IndexLayout *L) {
    if (n <= 0 || total <= 0 || tau < 2 || total >= 0x7FFFFF00LL || world < 1) {
        set_error("panib_index: n=%lld entries=%lld tau=%d world=%d out of range (entries must be < 2^31)",
                  (long long)n, (long long)total, tau, world);
        return PANIB_E_ARG;
    }
    L->total = total;
    L->wcap = (L->total / tau + 31) / 32 + 1;
    // table: twice the entries a rank can own; a rank owns 1/world of the hash range, sized for twice its share
    int64_t own = world == 1 ? total : 2 * (total / world) + 4096;
    if (own > total) own = total;
    int64_t slots = 1024;
    while (slots < 2 * own) slots <<= 1;
    L->slots = slots;
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t at = 0;
    L->off_keys = at;  at = align(at + (size_t)slots * 8);
    L->off_count = at; at = align(at + (size_t)slots * 4);
    L->off_aux = at;   at = align(at + (size_t)slots * 4);
    L->off_eslot = at; at = align(at + (size_t)total * 4);
    L->off_eidx = at;  at = align(at + (size_t)total * 4);
    L->off_lists = at; at = align(at + (size_t)total * 4);
    L->off_bits = at;  at = align(at + (size_t)n * L->wcap * 4);
    L->bytes = at;
    return PANIB_OK;
}

static dim3 entry_grid(int64_t n, int64_t cap) {  // x = genomes (fast), y = chunks of the longest sketch
    int64_t gy = (cap + kIdxThreads - 1) / kIdxThreads;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    const int64_t gx = n < (1 << 30) ? n : (1 << 30);
    return dim3((unsigned)gx, (unsigned)gy, 1);
}

}  // namespace panib

using namespace panib;

extern "C" int panib_index_workspace_bytes(int64_t n, int64_t total, int tau, int world, int64_t *bytes) {
    IndexLayout L;
    int rc = index_layout(n, total, tau, world, &L);
    if (rc) return rc;
    *bytes = (int64_t)L.bytes;
    return PANIB_OK;
}

// Phase 1: insert every entry this rank owns, classify the groups.  d_stats (uint64[4]) receives
// [0] frequent hashes (bit-matrix columns), [1] pairs the rare hashes expand to, [2] distinct hashes,
// [3] list words -- all for THIS rank's slice of the hash range.
extern "C" int panib_index_build(const uint64_t *d_rows, const int32_t *d_counts, int64_t stride, int64_t n,
                                 uint64_t max_hash, int64_t cap, int64_t total, const int64_t *d_offsets, int tau,
                                 int rank, int world, void *d_work, int64_t work_bytes, uint64_t *d_stats,
                                 int32_t *d_status, void *stream) {
    IndexLayout L;
    int rc = index_layout(n, total, tau, world, &L);
    if (rc) return rc;
    if (cap <= 0 || cap > 0x7FFFFFFFLL || (!d_offsets && total != n * cap) || rank < 0 || rank >= world) {
        set_error("panib_index_build: cap=%lld / entries=%lld / rank %d of %d inconsistent (padded form needs "
                  "entries == n*cap)", (long long)cap, (long long)total, rank, world);
        return PANIB_E_ARG;
    }
    if (!d_work || work_bytes < (int64_t)L.bytes || !d_stats || max_hash >= 0xFFFFFFFFFFFFFFF0ull) {
        set_error("panib_index_build: workspace of %lld bytes given, %zu needed (or max_hash too large)",
                  (long long)work_bytes, L.bytes);
        return PANIB_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char *base = static_cast<char *>(d_work);
    auto *tkeys = reinterpret_cast<unsigned long long *>(base + L.off_keys);
    auto *tcount = reinterpret_cast<uint32_t *>(base + L.off_count);
    auto *taux = reinterpret_cast<uint32_t *>(base + L.off_aux);
    auto *eslot = reinterpret_cast<uint32_t *>(base + L.off_eslot);
    auto *eidx = reinterpret_cast<uint32_t *>(base + L.off_eidx);
    // order-preserving maps of the hash range [0, max_hash]: onto the ranks, and onto world * slots table
    // positions (a rank's slice then covers its own table once)
    const unsigned __int128 range = (unsigned __int128)max_hash + 1;
    auto scale = [&](unsigned __int128 parts) -> uint64_t {
        const unsigned __int128 m = (parts << 64) / range;
        return m > (unsigned __int128)UINT64_MAX ? UINT64_MAX : (uint64_t)m;
    };
    IndexShard sh;
    sh.wmul = scale((unsigned __int128)world);
    sh.smul = scale((unsigned __int128)world * (unsigned __int128)L.slots);
    sh.mask = (uint32_t)(L.slots - 1);
    sh.rank = rank;
    sh.world = world;

    PANIB_CUDA(cudaMemsetAsync(d_stats, 0, 4 * sizeof(uint64_t), st));
    PANIB_CUDA(cudaMemsetAsync(tkeys, 0xFF, (size_t)L.slots * 8, st));
    PANIB_CUDA(cudaMemsetAsync(tcount, 0, (size_t)L.slots * 4, st));
    const dim3 grid = entry_grid(n, cap);
    index_insert_kernel<<<grid, kIdxThreads, 0, st>>>(d_rows, d_counts, stride, (int)n, (int)cap, d_offsets, L.total,
                                                      max_hash, sh, tkeys, tcount, eslot, eidx, d_status);
    rc = check_launch("index_insert_kernel");
    if (rc) return rc;
    // fewer, looping CTAs here: the kernel ends with two statistics atomics per warp
    dim3 cgrid = grid;
    if (cgrid.x > 296) cgrid.x = 296;
    if (cgrid.y > 8) cgrid.y = 8;
    index_classify_kernel<<<cgrid, kIdxThreads, 0, st>>>(d_counts, (int)n, (int)cap, d_offsets, L.total, tau, tcount,
                                                        taux, eslot, eidx,
                                                        reinterpret_cast<unsigned long long *>(d_stats));
    return check_launch("index_classify_kernel");
}

// Phase 2: intersection sizes from the index built by panib_index_build with the SAME arguments.
// d_ov (uint32 [n x ld_ov]) is fully overwritten: counts for i != j (mirrored), sketch sizes on the
// diagonal (rank 0 only, as panib_intersect); the ranks' matrices sum to the full result.
extern "C" int panib_index_count(const int32_t *d_counts, int64_t n, uint64_t max_hash, int64_t cap, int64_t total,
                                 const int64_t *d_offsets, int tau, void *d_work, int64_t work_bytes,
                                 const uint64_t *d_stats, uint32_t *d_ov, int64_t ld_ov, int rank, int world,
                                 int32_t *d_status, void *stream) {
    (void)max_hash;
    IndexLayout L;
    int rc = index_layout(n, total, tau, world, &L);
    if (rc) return rc;
    if (!d_work || work_bytes < (int64_t)L.bytes || ld_ov < n || rank < 0 || rank >= world || cap <= 0) {
        set_error("panib_index_count: bad workspace / ld_ov / rank arguments");
        return PANIB_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char *base = static_cast<char *>(d_work);
    const auto *tcount = reinterpret_cast<const uint32_t *>(base + L.off_count);
    const auto *taux = reinterpret_cast<const uint32_t *>(base + L.off_aux);
    const auto *eslot = reinterpret_cast<const uint32_t *>(base + L.off_eslot);
    const auto *eidx = reinterpret_cast<const uint32_t *>(base + L.off_eidx);
    auto *lists = reinterpret_cast<uint32_t *>(base + L.off_lists);
    uint32_t *bits = reinterpret_cast<uint32_t *>(base + L.off_bits);
    const unsigned long long *stats = reinterpret_cast<const unsigned long long *>(d_stats);

    PANIB_CUDA(cudaMemsetAsync(d_ov, 0, (size_t)n * ld_ov * sizeof(uint32_t), st));
    PANIB_CUDA(cudaMemsetAsync(bits, 0, (size_t)n * L.wcap * sizeof(uint32_t), st));
    const dim3 grid = entry_grid(n, cap);
    index_emit_kernel<<<grid, kIdxThreads, 0, st>>>(d_counts, (int)n, (int)cap, d_offsets, L.total, tau, tcount, taux,
                                                    eslot, eidx, lists, L.total, bits, L.wcap, d_status);
    rc = check_launch("index_emit_kernel");
    if (rc) return rc;
    index_sparse_kernel<<<grid, kIdxThreads, 0, st>>>(d_counts, (int)n, (int)cap, d_offsets, L.total, tau, tcount,
                                                      taux, eslot, eidx, lists, L.total, d_ov, ld_ov);
    rc = check_launch("index_sparse_kernel");
    if (rc) return rc;
    const unsigned nt = (unsigned)((n + kDenseTile - 1) / kDenseTile);
    index_dense_kernel<<<dim3(nt, nt), 256, 0, st>>>(bits, L.wcap, stats, (int)n, d_ov, ld_ov, rank, world);
    rc = check_launch("index_dense_kernel");
    if (rc) return rc;
    const unsigned nm = (unsigned)((n + 31) / 32);
    index_mirror_kernel<<<dim3(nm, nm), 256, 0, st>>>(d_ov, ld_ov, (int)n, d_counts, rank == 0 ? 1 : 0);
    return check_launch("index_mirror_kernel");
}
