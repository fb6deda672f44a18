// index.cu -- kernel K2, inverted-index form ("which genomes hold this hash?").
//
// The probing kernel in pairwise.cu touches every element of both sketches for every pair: its cost is
// N^2 * |sketch| whatever the genomes share.  This form costs what the genomes SHARE.  All sketches are
// flattened to (hash, genome) entries and sorted by hash (CUB radix sort); a run of equal hashes is the
// list of genomes holding that hash, and it contributes +1 to ov[i][j] for every pair of its members:
//
//   * frequent hashes (run length >= tau) become columns of a genome x column BIT MATRIX, and
//     ov[i][j] += popcount(row_i AND row_j) is computed for all pairs by a tiled AND+POPC kernel
//     (64 x 64 genome tiles, 4 x 4 pairs per thread, 32-word chunks staged in shared memory);
//   * rare hashes (2 <= run length < tau) are expanded pair by pair with atomicAdd into ov
//     (at most tau - 1 adds per entry);
//   * hashes held by one genome contribute nothing and are dropped.
//
// The entries are either exact (offsets = prefix sums of the sketch sizes, computed by the caller) or
// padded to `cap` slots per genome with a key above max_hash (no host knowledge of the sizes needed:
// the form a captured CUDA graph uses).
//
// Results are identical to the probing kernel (exact integer counts); which form is cheaper depends
// on the data, so the host chooses (engine.py: Engine.intersect(method="auto")) from the statistics
// panib_index_build returns.  Replaces the same reference step as pairwise.cu: the external
// `sourmash scripts manysearch` call in pyani_plus/methods/sourmash.py:184-200.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace panib {

constexpr int kIdxThreads = 256;
constexpr int kDenseTile = 64;    // genomes per tile side of the AND+POPC kernel
constexpr int kDenseChunk = 32;   // bit-matrix words staged per step

// entry e = g * cap + i: hash i of genome g, or the pad key (> max_hash) beyond the genome's size
__global__ void __launch_bounds__(kIdxThreads)
index_flatten_kernel(const uint64_t *__restrict__ rows, const int32_t *__restrict__ counts, int64_t stride, int n,
                     int cap, uint64_t pad_key, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                     int32_t *status) {
    const int64_t e = (int64_t)blockIdx.x * kIdxThreads + threadIdx.x;
    if (e >= (int64_t)n * cap) return;
    const int g = (int)(e / cap), i = (int)(e - (int64_t)g * cap);
    const int c = counts[g];
    if (i == 0 && c > cap && status) atomicOr(status, PANIB_ST_SEGMENT_OVERFLOW);  // cap too small: re-plan
    keys[e] = i < c ? rows[(size_t)g * stride + i] : pad_key;
    vals[e] = (uint32_t)g;
}

// exact-size form: genome g's hashes go to entries offsets[g] .. offsets[g] + counts[g] (no pad entries);
// grid = (chunks of the largest sketch, genomes)
__global__ void __launch_bounds__(kIdxThreads)
index_flatten_exact_kernel(const uint64_t *__restrict__ rows, const int32_t *__restrict__ counts, int64_t stride,
                           int n, int cap, const int64_t *__restrict__ offsets, int64_t total,
                           uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, int32_t *status) {
    // `cap` only sizes the grid: a genome larger than the hint is still flattened completely (strided loop)
    (void)cap;
    for (int g = blockIdx.y; g < n; g += gridDim.y) {
        const int c = counts[g];
        const int64_t at = offsets[g];
        if (blockIdx.x == 0 && threadIdx.x == 0 && (at < 0 || at + c > total) && status)
            atomicOr(status, PANIB_ST_SEGMENT_OVERFLOW);  // offsets do not match the sizes
        for (int i = blockIdx.x * kIdxThreads + threadIdx.x; i < c; i += gridDim.x * kIdxThreads) {
            if (at + i < total) {
                keys[at + i] = rows[(size_t)g * stride + i];
                vals[at + i] = (uint32_t)g;
            }
        }
    }
}

__global__ void __launch_bounds__(kIdxThreads)
index_heads_kernel(const uint64_t *__restrict__ keys, int64_t total, int32_t *__restrict__ head) {
    const int64_t e = (int64_t)blockIdx.x * kIdxThreads + threadIdx.x;
    if (e >= total) return;
    head[e] = (e == 0 || keys[e] != keys[e - 1]) ? 1 : 0;
}

// gidx = inclusive scan of the head flags: entry e belongs to run gidx[e] - 1; start[r] = first entry of run r
__global__ void __launch_bounds__(kIdxThreads)
index_starts_kernel(const int32_t *__restrict__ gidx, int64_t total, int32_t *__restrict__ start,
                    unsigned long long *__restrict__ stats) {
    const int64_t e = (int64_t)blockIdx.x * kIdxThreads + threadIdx.x;
    if (e >= total) return;
    const int r = gidx[e];
    if (e == 0 || gidx[e - 1] != r) start[r - 1] = (int32_t)e;
    if (e == total - 1) {
        start[r] = (int32_t)total;
        stats[2] = (unsigned long long)r;  // number of runs (the pad run included)
    }
}

// per run: frequent -> a bit-matrix column, rare -> pair expansion, single / pad -> nothing
__global__ void __launch_bounds__(kIdxThreads)
index_classify_kernel(const uint64_t *__restrict__ keys, const int32_t *__restrict__ start,
                      const int32_t *__restrict__ gidx, int64_t total, uint64_t max_hash, int tau,
                      int32_t *__restrict__ densecol, unsigned long long *__restrict__ stats) {
    const int64_t r = (int64_t)blockIdx.x * kIdxThreads + threadIdx.x;
    const bool live = r < gidx[total - 1];
    unsigned long long pairs = 0ull;
    if (live) {
        const int s = start[r], m = start[r + 1] - s;
        int col = -1;
        if (keys[s] <= max_hash) {
            if (m >= tau) col = (int)atomicAdd(stats + 0, 1ull);
            else if (m >= 2) pairs = (unsigned long long)m * (unsigned long long)(m - 1) / 2ull;
        }
        densecol[r] = col;
    }
    // one atomic per warp for the rare-pair total (millions of runs would otherwise queue on one address)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(0xFFFFFFFFu, pairs, o);
    if ((threadIdx.x & 31) == 0 && pairs) atomicAdd(stats + 1, pairs);
}

__global__ void __launch_bounds__(kIdxThreads)
index_bits_kernel(const uint32_t *__restrict__ vals, const int32_t *__restrict__ gidx,
                  const int32_t *__restrict__ densecol, int64_t total, int n, uint32_t *__restrict__ bits,
                  int64_t wcap, int32_t *status) {
    const int64_t e = (int64_t)blockIdx.x * kIdxThreads + threadIdx.x;
    if (e >= total) return;
    const int col = densecol[gidx[e] - 1];
    if (col < 0) return;
    if ((col >> 5) >= wcap || vals[e] >= (uint32_t)n) {  // cannot happen with a consistent index
        if (status) atomicOr(status, PANIB_ST_SEGMENT_OVERFLOW);
        return;
    }
    atomicOr(bits + (size_t)vals[e] * wcap + (col >> 5), 1u << (col & 31));
}

// rare hashes: entry e pairs with the later members of its run (genome ids ascend inside a run because
// the radix sort is stable and the flattening is genome-major), so only i < j is written
__global__ void __launch_bounds__(kIdxThreads)
index_sparse_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                    const int32_t *__restrict__ gidx, const int32_t *__restrict__ start,
                    const int32_t *__restrict__ densecol, int64_t total, int n, uint64_t max_hash,
                    uint32_t *__restrict__ ov, int64_t ld, int rank, int world) {
    const int64_t e = (int64_t)blockIdx.x * kIdxThreads + threadIdx.x;
    if (e >= total) return;
    const int r = gidx[e] - 1;
    if (densecol[r] >= 0 || keys[e] > max_hash) return;
    if (world > 1 && r % world != rank) return;
    const int end = start[r + 1];
    const uint32_t i = vals[e];
    if (i >= (uint32_t)n) return;
    for (int64_t x = e + 1; x < end && x < total; x++) {
        const uint32_t j = vals[x];
        if (j < (uint32_t)n) atomicAdd(ov + (size_t)i * ld + j, 1u);
    }
}

// frequent hashes: ov[i][j] += popcount(bits[i] & bits[j]) over the used words, for i < j.
// One CTA per 64 x 64 tile on or above the diagonal; thread (ty, tx) owns pairs (ty + 16p, tx + 16q).
__global__ void __launch_bounds__(256)
index_dense_kernel(const uint32_t *__restrict__ bits, int64_t wcap, const unsigned long long *__restrict__ stats,
                   int n, uint32_t *__restrict__ ov, int64_t ld, int rank, int world) {
    const int I = blockIdx.y, J = blockIdx.x;
    if (I > J) return;
    if (world > 1 && (int)(((int64_t)I * gridDim.x + J) % world) != rank) return;
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
    int64_t total, wcap;
    size_t off_keys[2], off_vals[2], off_bits, off_temp, temp_bytes, bytes;
};

static int index_layout(int64_t n, int64_t total, int tau, IndexLayout *L) {
    if (n <= 0 || total <= 0 || tau < 2 || total >= 0x7FFFFF00LL) {
        set_error("panib_index: n=%lld entries=%lld tau=%d out of range (entries must be < 2^31)", (long long)n,
                  (long long)total, tau);
        return PANIB_E_ARG;
    }
    L->total = total;
    L->wcap = (L->total / tau + 31) / 32 + 1;
    size_t sort_bytes = 0, scan_bytes = 0;
    if (cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                        (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)L->total, 0,
                                        64) != cudaSuccess ||
        cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (const int32_t *)nullptr, (int32_t *)nullptr,
                                      (int)L->total) != cudaSuccess) {
        set_error("panib_index: CUB temporary-storage query failed");
        return PANIB_E_CUDA;
    }
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t at = 0;
    for (int b = 0; b < 2; b++) { L->off_keys[b] = at; at = align(at + (size_t)(L->total + 2) * 8); }
    for (int b = 0; b < 2; b++) { L->off_vals[b] = at; at = align(at + (size_t)(L->total + 2) * 4); }
    L->off_bits = at; at = align(at + (size_t)n * L->wcap * 4);
    L->temp_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    L->off_temp = at; at = align(at + L->temp_bytes);
    L->bytes = at;
    return PANIB_OK;
}

}  // namespace panib

using namespace panib;

extern "C" int panib_index_workspace_bytes(int64_t n, int64_t total, int tau, int64_t *bytes) {
    IndexLayout L;
    int rc = index_layout(n, total, tau, &L);
    if (rc) return rc;
    *bytes = (int64_t)L.bytes;
    return PANIB_OK;
}

// Phase 1: flatten, sort, find the runs, classify them.  d_stats (uint64[4]) receives
// [0] frequent hashes (bit-matrix columns), [1] pairs the rare hashes expand to, [2] runs, [3] unused.
extern "C" int panib_index_build(const uint64_t *d_rows, const int32_t *d_counts, int64_t stride, int64_t n,
                                 uint64_t max_hash, int64_t cap, int64_t total, const int64_t *d_offsets, int tau,
                                 void *d_work, int64_t work_bytes, uint64_t *d_stats, int32_t *d_status,
                                 void *stream) {
    IndexLayout L;
    int rc = index_layout(n, total, tau, &L);
    if (rc) return rc;
    if (cap <= 0 || cap > 0x7FFFFFFFLL || (!d_offsets && total != n * cap)) {
        set_error("panib_index_build: cap=%lld / entries=%lld inconsistent (padded form needs entries == n*cap)",
                  (long long)cap, (long long)total);
        return PANIB_E_ARG;
    }
    if (!d_work || work_bytes < (int64_t)L.bytes || !d_stats || max_hash >= 0xFFFFFFFFFFFFFFF0ull) {
        set_error("panib_index_build: workspace of %lld bytes given, %zu needed (or max_hash too large)",
                  (long long)work_bytes, L.bytes);
        return PANIB_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char *base = static_cast<char *>(d_work);
    uint64_t *keys[2] = {reinterpret_cast<uint64_t *>(base + L.off_keys[0]),
                         reinterpret_cast<uint64_t *>(base + L.off_keys[1])};
    uint32_t *vals[2] = {reinterpret_cast<uint32_t *>(base + L.off_vals[0]),
                         reinterpret_cast<uint32_t *>(base + L.off_vals[1])};
    const int64_t T = L.total;
    const unsigned blocks = (unsigned)((T + kIdxThreads - 1) / kIdxThreads);
    const uint64_t pad_key = max_hash + 1;
    int end_bit = 1;
    while (end_bit < 64 && (pad_key >> end_bit)) end_bit++;

    PANIB_CUDA(cudaMemsetAsync(d_stats, 0, 4 * sizeof(uint64_t), st));
    if (d_offsets) {
        const dim3 grid((unsigned)((cap + kIdxThreads - 1) / kIdxThreads), (unsigned)(n < 65535 ? n : 65535));
        index_flatten_exact_kernel<<<grid, kIdxThreads, 0, st>>>(d_rows, d_counts, stride, (int)n, (int)cap, d_offsets,
                                                                 T, keys[0], vals[0], d_status);
    } else {
        index_flatten_kernel<<<blocks, kIdxThreads, 0, st>>>(d_rows, d_counts, stride, (int)n, (int)cap, pad_key,
                                                             keys[0], vals[0], d_status);
    }
    rc = check_launch("index_flatten_kernel");
    if (rc) return rc;
    // keys[0]/vals[0] = flattened input, keys[1]/vals[1] = sorted output; after the sort the input
    // buffers are re-used: vals[0] -> head flags, then densecol; keys[0] -> gidx (T+1 ints) + start (T+1 ints)
    size_t temp = L.temp_bytes;
    PANIB_CUDA(cub::DeviceRadixSort::SortPairs(base + L.off_temp, temp, keys[0], keys[1], vals[0], vals[1], (int)T,
                                               0, end_bit, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const uint64_t *skeys = keys[1];
    int32_t *head = reinterpret_cast<int32_t *>(vals[0]);
    int32_t *gidx = reinterpret_cast<int32_t *>(keys[0]);
    int32_t *start = gidx + (T + 1);
    index_heads_kernel<<<blocks, kIdxThreads, 0, st>>>(skeys, T, head);
    rc = check_launch("index_heads_kernel");
    if (rc) return rc;
    temp = L.temp_bytes;
    PANIB_CUDA(cub::DeviceScan::InclusiveSum(base + L.off_temp, temp, head, gidx, (int)T, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    index_starts_kernel<<<blocks, kIdxThreads, 0, st>>>(gidx, T, start,
                                                        reinterpret_cast<unsigned long long *>(d_stats));
    rc = check_launch("index_starts_kernel");
    if (rc) return rc;
    index_classify_kernel<<<blocks, kIdxThreads, 0, st>>>(skeys, start, gidx, T, max_hash, tau, head,
                                                          reinterpret_cast<unsigned long long *>(d_stats));
    return check_launch("index_classify_kernel");
}

// Phase 2: intersection sizes from the index built by panib_index_build with the SAME arguments.
// d_ov (uint32 [n x ld_ov]) is fully overwritten: counts for i != j (mirrored), sketch sizes on the
// diagonal (rank 0 only, as panib_intersect); the ranks' matrices sum to the full result.
extern "C" int panib_index_count(const int32_t *d_counts, int64_t n, uint64_t max_hash, int64_t total, int tau,
                                 void *d_work, int64_t work_bytes, const uint64_t *d_stats, uint32_t *d_ov,
                                 int64_t ld_ov, int rank, int world, int32_t *d_status, void *stream) {
    IndexLayout L;
    int rc = index_layout(n, total, tau, &L);
    if (rc) return rc;
    if (!d_work || work_bytes < (int64_t)L.bytes || ld_ov < n || world < 1 || rank < 0 || rank >= world) {
        set_error("panib_index_count: bad workspace / ld_ov / rank arguments");
        return PANIB_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char *base = static_cast<char *>(d_work);
    const int64_t T = L.total;
    const unsigned blocks = (unsigned)((T + kIdxThreads - 1) / kIdxThreads);
    const uint64_t *skeys = reinterpret_cast<const uint64_t *>(base + L.off_keys[1]);
    const uint32_t *svals = reinterpret_cast<const uint32_t *>(base + L.off_vals[1]);
    const int32_t *densecol = reinterpret_cast<const int32_t *>(base + L.off_vals[0]);
    const int32_t *gidx = reinterpret_cast<const int32_t *>(base + L.off_keys[0]);
    const int32_t *start = gidx + (T + 1);
    uint32_t *bits = reinterpret_cast<uint32_t *>(base + L.off_bits);
    const unsigned long long *stats = reinterpret_cast<const unsigned long long *>(d_stats);

    PANIB_CUDA(cudaMemsetAsync(d_ov, 0, (size_t)n * ld_ov * sizeof(uint32_t), st));
    PANIB_CUDA(cudaMemsetAsync(bits, 0, (size_t)n * L.wcap * sizeof(uint32_t), st));
    index_bits_kernel<<<blocks, kIdxThreads, 0, st>>>(svals, gidx, densecol, T, (int)n, bits, L.wcap, d_status);
    rc = check_launch("index_bits_kernel");
    if (rc) return rc;
    index_sparse_kernel<<<blocks, kIdxThreads, 0, st>>>(skeys, svals, gidx, start, densecol, T, (int)n, max_hash, d_ov,
                                                        ld_ov, rank, world);
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
