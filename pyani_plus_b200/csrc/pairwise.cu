// pairwise.cu -- stage 2 (kernel K2: all-vs-all sorted-sketch intersection) and stage 3 (ANI).
//
// Replaces the arithmetic of `sourmash scripts manysearch -m DNA -t 0` (reference call site:
// pyani_plus/methods/sourmash.py:184-200, compute_sourmash_tile :147-206): for every
// (query, subject) pair the exact size of the intersection of two sorted, duplicate-free u64 lists.
//
// Design (DESIGN.md "K2").  One CTA = one work item = (query row i, value cell c, block of JB subject
// columns).  The CTA stages the query's segment (its hashes inside cell c) in shared memory ONCE,
// builds a monotone bucket index over it (idx[b] = first position whose bucket >= b; sketches are
// uniformly distributed hashes, so a multiplicative map gives ~0.3 elements per bucket), and then
// each warp streams one subject column from global/L2 with coalesced 8-byte loads; every streamed
// hash costs one index load and, on average, ~0.3 compares.  Because the streamed column is sorted,
// neighbouring lanes touch neighbouring index entries and segment slots (few bank conflicts).
// A staged query is re-used by all JB columns of the item, and work items are ordered so that
// concurrently running CTAs share the same super-block of columns, which therefore stays L2
// resident: HBM traffic is far below the algorithmic 8(|A|+|B|) bytes per pair.
#include <math.h>

#include <sched.h>

#include <thread>
#include <vector>

#include "common.cuh"

namespace panib {

struct K2Args {
    const uint64_t *q_rows;
    const int32_t *q_counts;
    int64_t q_stride;
    int nq;
    const uint64_t *s_rows;
    const int32_t *s_counts;
    int64_t s_stride;
    int ns;
    const int32_t *q_fence;  // [nq][G+1] positions of the cell boundaries in each query row (G > 1)
    const int32_t *s_fence;  // [ns][G+1]
    int G;                   // number of value cells
    uint64_t cellw;          // width of a value cell (cell c = [c*cellw, (c+1)*cellw)); unused when G == 1
    int pre;                 // bucket(x) = umulhi((x - base) >> pre, mul)
    uint32_t mul;
    int R;                   // index buckets
    int seg_cap;             // shared-memory capacity (elements) of a staged segment
    int symmetric;
    int JB;                  // subject columns per work item
    int SB;                  // work items (column blocks) per L2 super-block
    int nJB;                 // number of column blocks
    uint32_t *ov;
    int64_t ld;
    int rank, world;
    int32_t *status;
};

// positions of the cell boundaries inside every row: fence[r][c] = lower_bound(row r, c * cellw)
__global__ void __launch_bounds__(256)
fence_kernel(const uint64_t *__restrict__ rows, const int32_t *__restrict__ counts, int64_t stride, int n, int G,
             uint64_t cellw, int32_t *__restrict__ fence) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * (G + 1)) return;
    const int r = (int)(t / (G + 1)), c = (int)(t % (G + 1));
    const int cnt = counts[r];
    int pos;
    if (c == 0) pos = 0;
    else if (c == G) pos = cnt;
    else {
        const uint64_t v = (uint64_t)c * cellw;
        const uint64_t *row = rows + (size_t)r * stride;
        int lo = 0, hi = cnt;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (row[mid] < v) lo = mid + 1; else hi = mid;
        }
        pos = lo;
    }
    fence[t] = pos;
}

__global__ void __launch_bounds__(256)
diag_kernel(const int32_t *__restrict__ counts, int n, uint32_t *__restrict__ ov, int64_t ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ov[(size_t)i * ld + i] = (uint32_t)counts[i];
}

constexpr int kThreadsK2 = 512;
constexpr uint32_t kIdxMany = 0x8000u;  // flag in an index entry: the bucket holds more than 2 elements

// bucket of a hash inside its cell: monotone in x, < R (mul <= 2^32 R / (vmax+1), see make_plan)
template <bool kCells>
__device__ __forceinline__ uint32_t k2_bucket(uint64_t x, uint64_t base, int pre, uint32_t mul) {
    if (kCells) x -= base;
    // low 32 bits of x >> pre for 0 <= pre <= 32 in one funnel shift (the shift amount clamps at 32)
    return __umulhi(__funnelshift_rc((uint32_t)x, (uint32_t)(x >> 32), (uint32_t)pre), mul);
}

// (A variant that kept the staged segment as two 32-bit planes and compared low words first was
// measured 12 % SLOWER: with ~3 % of probes matching, most warps take the divergent "verify" branch.)

template <bool kCells>
__global__ void __launch_bounds__(kThreadsK2) intersect_kernel(const K2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *seg = reinterpret_cast<uint64_t *>(smem_raw);
    uint16_t *idx = reinterpret_cast<uint16_t *>(seg + a.seg_cap);

    // ---- decode the work item: (super-block, query row, cell, column block inside the super-block)
    int64_t id = (int64_t)blockIdx.x + (int64_t)blockIdx.y * gridDim.x;
    const int jl = (int)(id % a.SB);
    int64_t t = id / a.SB;
    const int c = kCells ? (int)(t % a.G) : 0;
    if (kCells) t /= a.G;
    const int i = (int)(t % a.nq);
    const int sb = (int)(t / a.nq);
    const int jb = sb * a.SB + jl;
    if (jb >= a.nJB) return;
    int j0 = jb * a.JB;
    const int j1 = min(a.ns, j0 + a.JB);
    if (a.symmetric) {
        if (j1 - 1 <= i) return;  // only q < s is computed
        j0 = max(j0, i + 1);
    }
    if (a.world > 1 && (int)(id % a.world) != a.rank) return;

    // ---- stage the query segment and build its bucket index
    const int tid = threadIdx.x;
    const uint64_t *qrow = a.q_rows + (size_t)i * a.q_stride;
    int q0 = 0, q1 = a.q_counts[i];
    if (kCells) {
        q0 = a.q_fence[(size_t)i * (a.G + 1) + c];
        q1 = a.q_fence[(size_t)i * (a.G + 1) + c + 1];
    }
    // the all-ones hash (possible only when max_hash == 2^64-1) doubles as the shared-memory
    // sentinel, so it is stripped from both lists and counted on its own
    const bool q_ones = q1 > q0 && qrow[q1 - 1] == kEmpty;
    if (q_ones) --q1;
    const int n = q1 - q0;
    if (n + 2 > a.seg_cap) {
        if (tid == 0) atomicOr(a.status, PANIB_ST_SEGMENT_OVERFLOW);
        return;
    }
    const uint64_t base = kCells ? (uint64_t)c * a.cellw : 0ull;
    const int pre = a.pre;
    const uint32_t mul = a.mul;
    const int R = a.R;
    for (int p = tid; p < n + 2; p += kThreadsK2) seg[p] = p < n ? qrow[q0 + p] : kEmpty;
    __syncthreads();
    // idx[b] = first position whose bucket is >= b (every entry is written exactly once)
    for (int p = tid; p < n; p += kThreadsK2) {
        const int b = (int)k2_bucket<kCells>(seg[p], base, pre, mul);
        const int bprev = p > 0 ? (int)k2_bucket<kCells>(seg[p - 1], base, pre, mul) : -1;
        for (int q = bprev + 1; q <= b; q++) idx[q] = (uint16_t)p;
    }
    {
        const int blast = n > 0 ? (int)k2_bucket<kCells>(seg[n - 1], base, pre, mul) : -1;
        for (int q = blast + 1 + tid; q <= R; q += kThreadsK2) idx[q] = (uint16_t)n;
    }
    __syncthreads();
    // flag the buckets with more than two elements.  A neighbour's entry is read to decide, so
    // deciding and writing are separated by a barrier, 32 buckets per thread (one mask register) at a
    // time: one round for R <= 16384.
    for (int q0 = 0; q0 < R; q0 += 32 * kThreadsK2) {
        uint32_t many = 0;
#pragma unroll 4
        for (int it = 0; it < 32; ++it) {
            const int q = q0 + it * kThreadsK2 + tid;
            if (q < R && (uint32_t)idx[q + 1] - (uint32_t)idx[q] > 2u) many |= 1u << it;
        }
        __syncthreads();
        for (; many; many &= many - 1) {
            const int q = q0 + (__ffs(many) - 1) * kThreadsK2 + tid;
            idx[q] = (uint16_t)(idx[q] | kIdxMany);
        }
        __syncthreads();
    }

    // ---- each warp streams subject columns (sorted, so neighbouring lanes probe neighbouring slots)
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kThreadsK2 / 32;
    for (int col = j0 + warp; col < j1; col += nwarps) {
        const uint64_t *srow = a.s_rows + (size_t)col * a.s_stride;
        int s0 = 0, s1 = a.s_counts[col];
        if (kCells) {
            s0 = a.s_fence[(size_t)col * (a.G + 1) + c];
            s1 = a.s_fence[(size_t)col * (a.G + 1) + c + 1];
        }
        uint32_t cnt = 0;
        if (s1 > s0 && __ldg(srow + s1 - 1) == kEmpty) {
            --s1;
            cnt = (lane == 0 && q_ones) ? 1u : 0u;
        }
        auto probe = [&](uint64_t x) {
            const uint32_t b = k2_bucket<kCells>(x, base, pre, mul);
            const int e = reinterpret_cast<const int16_t *>(idx)[b];  // sign bit = "crowded bucket" flag
            const uint32_t lo = (uint32_t)e & 0x7FFFu;
            // a sketch holds distinct hashes, so at most one slot can match: one predicated add
            const bool found = (seg[lo] == x) | (seg[lo + 1] == x);
            if (found) ++cnt;
            if (e < 0) {  // rare: walk the rest of a crowded bucket
                const uint32_t hi = idx[b + 1] & 0x7FFFu;
                for (uint32_t q = lo + 2; q < hi; q++) cnt += (seg[q] == x) ? 1u : 0u;
            }
        };
        int p = s0 + lane;
        for (; p + 96 < s1; p += 128) {  // four independent loads in flight per lane
            const uint64_t x0 = __ldg(srow + p), x1 = __ldg(srow + p + 32), x2 = __ldg(srow + p + 64),
                           x3 = __ldg(srow + p + 96);
            probe(x0); probe(x1); probe(x2); probe(x3);
        }
        for (; p < s1; p += 32) probe(__ldg(srow + p));
        cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
        if (lane == 0) {
            if (kCells) {
                if (cnt) {
                    atomicAdd(&a.ov[(size_t)i * a.ld + col], cnt);
                    if (a.symmetric) atomicAdd(&a.ov[(size_t)col * a.ld + i], cnt);
                }
            } else {
                a.ov[(size_t)i * a.ld + col] = cnt;
                if (a.symmetric) a.ov[(size_t)col * a.ld + i] = cnt;
            }
        }
    }
}

// ---- containment -> ANI (sourmash ani_from_containment; SURVEY.md 8c convention 11) -------------
__host__ __device__ inline double ani_from_containment(double c, int k) {
    if (c == 0.0) return 0.0;
    if (c == 1.0) return 1.0;
    return 1.0 - (1.0 - pow(c, 1.0 / (double)k));
}

__host__ __device__ inline void pair_ani(uint32_t ov, int32_t nq, int32_t ns, int k, double &identity,
                                         double &cov_query) {
    if (ov == 0 || nq <= 0 || ns <= 0) {  // branchwater prints no row -> NULL in the database
        identity = nan("");
        cov_query = nan("");
        return;
    }
    const double qani = ani_from_containment((double)ov / (double)nq, k);
    const double mani = ani_from_containment((double)ov / (double)ns, k);
    identity = qani > mani ? qani : mani;  // max_containment_ani
    cov_query = qani;                      // query_containment_ani
}

__global__ void __launch_bounds__(256)
ani_kernel(const uint32_t *__restrict__ ov, int64_t ld, const int32_t *__restrict__ qc, int64_t nq,
           const int32_t *__restrict__ sc, int64_t ns, int k, double *__restrict__ identity,
           double *__restrict__ cov_query) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * ns) return;
    const int64_t q = t / ns, s = t % ns;
    double id, cv;
    pair_ani(ov[q * ld + s], qc[q], sc[s], k, id, cv);
    identity[t] = id;
    cov_query[t] = cv;
}

}  // namespace panib

// ================================================================================================
// C ABI
// ================================================================================================
using namespace panib;

namespace {

struct K2Plan {
    int G, seg_cap, R, pre;
    uint32_t mul;
    uint64_t cellw;
};

int bitlen64(uint64_t v) {
    int n = 0;
    while (v) { n++; v >>= 1; }
    return n;
}

// choose segmentation and the bucket map; 0 values mean "choose for me"
//   seg_cap : elements of a query segment staged in shared memory (+2 sentinels), <= 32767
//   R       : index buckets, ~3 per element (0.33 elements per bucket on average)
// Auto policy: one segment when the largest sketch has <= 6144 hashes (<= ~75 KB of shared memory,
// three CTAs per SM); larger sketches are cut into value cells of ~4900 hashes.
K2Plan make_plan(uint64_t max_hash, int64_t max_count, int n_cells, int seg_cap, int idx_buckets) {
    K2Plan p;
    if (max_count < 1) max_count = 1;
    if (n_cells <= 0) n_cells = max_count <= 6144 ? 1 : (int)((max_count * 5 + 4 * 6144 - 1) / (4 * 6144));
    if (seg_cap <= 0) {
        // expected cell population + 6 sigma (cells are equal slices of a uniform hash range)
        const double mean = (double)max_count / n_cells;
        const double want = n_cells == 1 ? (double)max_count : mean + 6.0 * sqrt(mean) + 16.0;
        seg_cap = (int)((int64_t)(want + 2 + 63) / 64 * 64);
    }
    if (seg_cap > 32767) seg_cap = 32767;
    if (idx_buckets <= 0) {
        // ~3 buckets per element, trimmed (not below 2 per element) when that keeps a third CTA
        // resident on the SM: 3 x (74.5 KB + 1 KB reserved) fit the 227 KB of shared memory
        idx_buckets = 3 * seg_cap;
        const int fit = (76288 - seg_cap * 8) / 2 - 2;
        if (fit < idx_buckets && fit >= 2 * seg_cap) idx_buckets = fit;
    }
    p.G = n_cells;
    p.seg_cap = seg_cap;
    p.R = idx_buckets;
    // cell width = ceil((max_hash + 1) / G), computed without overflowing for max_hash = 2^64-1
    const unsigned __int128 universe = (unsigned __int128)max_hash + 1;
    const unsigned __int128 cw = (universe + (unsigned)p.G - 1) / (unsigned)p.G;
    const uint64_t vmax_full = (uint64_t)(cw - 1);  // largest x - base inside a cell
    p.cellw = p.G > 1 ? (uint64_t)cw : 0;
    const int bl = bitlen64(vmax_full);
    p.pre = bl > 32 ? bl - 32 : 0;
    const uint64_t vmax = vmax_full >> p.pre;  // < 2^32
    // bucket(v) = floor(v * mul / 2^32) with mul <= 2^32 * R / (vmax + 1): monotone, < R
    unsigned __int128 m = (((unsigned __int128)p.R) << 32) / ((unsigned __int128)vmax + 1);
    if (m > 0xFFFFFFFFull) m = 0xFFFFFFFFull;
    p.mul = (uint32_t)m;
    return p;
}

}  // namespace

extern "C" int64_t panib_intersect_fence_entries(int64_t nq, int64_t ns, uint64_t max_hash, int64_t max_count,
                                                 int n_cells, int seg_cap) {
    const K2Plan p = make_plan(max_hash, max_count, n_cells, seg_cap, 0);
    return p.G > 1 ? (nq + ns) * (int64_t)(p.G + 1) : 0;
}

extern "C" int panib_intersect(const uint64_t *d_q_rows, const int32_t *d_q_counts, int64_t q_stride, int64_t nq,
                               const uint64_t *d_s_rows, const int32_t *d_s_counts, int64_t s_stride, int64_t ns,
                               int symmetric, uint64_t max_hash, int64_t max_count, int n_cells, int seg_cap,
                               int idx_buckets, int32_t *d_fence, uint32_t *d_ov, int64_t ld_ov, int rank,
                               int world, int32_t *d_status, void *stream) {
    if (nq <= 0 || ns <= 0) return PANIB_OK;
    if (nq > 0x7FFFFFFF || ns > 0x7FFFFFFF || ld_ov < ns || world < 1 || rank < 0 || rank >= world) {
        set_error("panib_intersect: bad sizes nq=%lld ns=%lld ld=%lld rank=%d world=%d", (long long)nq,
                  (long long)ns, (long long)ld_ov, rank, world);
        return PANIB_E_ARG;
    }
    if (symmetric && (d_q_rows != d_s_rows || nq != ns)) {
        set_error("panib_intersect: symmetric mode needs identical query and subject tables");
        return PANIB_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const K2Plan p = make_plan(max_hash, max_count, n_cells, seg_cap, idx_buckets);
    if (p.seg_cap > 32767 || p.seg_cap < 3 || p.R < 1) {
        set_error("panib_intersect: seg_cap=%d / idx_buckets=%d out of range", p.seg_cap, p.R);
        return PANIB_E_ARG;
    }
    const size_t smem = (size_t)p.seg_cap * 8 + ((size_t)p.R + 2) * 2;
    if (smem > 227 * 1024) {
        set_error("panib_intersect: %zu bytes of shared memory needed (max 227 KB)", smem);
        return PANIB_E_ARG;
    }
    static size_t smem_set = 0;
    if (smem > smem_set) {
        PANIB_CUDA(cudaFuncSetAttribute(intersect_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
        PANIB_CUDA(cudaFuncSetAttribute(intersect_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
        smem_set = smem;
    }
    PANIB_CUDA(cudaMemsetAsync(d_ov, 0, (size_t)nq * ld_ov * sizeof(uint32_t), st));

    K2Args a;
    a.q_rows = d_q_rows; a.q_counts = d_q_counts; a.q_stride = q_stride; a.nq = (int)nq;
    a.s_rows = d_s_rows; a.s_counts = d_s_counts; a.s_stride = s_stride; a.ns = (int)ns;
    a.q_fence = nullptr; a.s_fence = nullptr;
    a.G = p.G; a.cellw = p.cellw; a.pre = p.pre; a.mul = p.mul; a.R = p.R; a.seg_cap = p.seg_cap;
    a.symmetric = symmetric ? 1 : 0;
    a.ov = d_ov; a.ld = ld_ov; a.rank = rank; a.world = world; a.status = d_status;

    if (p.G > 1) {
        if (!d_fence) {
            set_error("panib_intersect: d_fence scratch required for %d cells", p.G);
            return PANIB_E_ARG;
        }
        const int64_t nqf = nq * (p.G + 1);
        fence_kernel<<<(unsigned)((nqf + 255) / 256), 256, 0, st>>>(d_q_rows, d_q_counts, q_stride, (int)nq, p.G,
                                                                    p.cellw, d_fence);
        int rc = check_launch("fence_kernel(q)");
        if (rc) return rc;
        a.q_fence = d_fence;
        if (symmetric) {
            a.s_fence = d_fence;
        } else {
            const int64_t nsf = ns * (p.G + 1);
            fence_kernel<<<(unsigned)((nsf + 255) / 256), 256, 0, st>>>(d_s_rows, d_s_counts, s_stride, (int)ns,
                                                                        p.G, p.cellw, d_fence + nqf);
            rc = check_launch("fence_kernel(s)");
            if (rc) return rc;
            a.s_fence = d_fence + nqf;
        }
    }

    // columns per work item (16 warps, one column each at a time): aim for a few thousand items so
    // that 148 SMs x 3 CTAs stay busy, while a staged query is re-used by as many columns as possible
    const double pairs = (double)nq * (double)ns * (symmetric ? 0.5 : 1.0) * p.G / (double)world;
    int JB = (int)(pairs / 3600.0);
    JB = JB < 16 ? 16 : (JB > 256 ? 256 : JB);
    JB = (JB + 15) & ~15;
    a.JB = JB;
    a.nJB = (int)((ns + JB - 1) / JB);
    int SB = 1024 / JB;  // ~1024 columns (a few tens of MB of sketches) per L2 super-block
    if (SB > a.nJB) SB = a.nJB;
    if (SB < 1) SB = 1;
    a.SB = SB;
    const int64_t nSB = (a.nJB + SB - 1) / SB;
    const int64_t items = nSB * nq * p.G * SB;
    const int64_t gx = items < (1 << 30) ? items : (1 << 30);
    const int64_t gy = (items + gx - 1) / gx;
    // (ids >= items decode to a column block >= nJB and exit immediately)
    if (p.G > 1) intersect_kernel<true><<<dim3((unsigned)gx, (unsigned)gy), kThreadsK2, smem, st>>>(a);
    else intersect_kernel<false><<<dim3((unsigned)gx, (unsigned)gy), kThreadsK2, smem, st>>>(a);
    int rc = check_launch("intersect_kernel");
    if (rc) return rc;
    if (symmetric && rank == 0) {
        diag_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(d_q_counts, (int)nq, d_ov, ld_ov);
        rc = check_launch("diag_kernel");
    }
    return rc;
}

extern "C" int panib_ani_device(const uint32_t *d_ov, int64_t ld_ov, const int32_t *d_q_counts, int64_t nq,
                                const int32_t *d_s_counts, int64_t ns, int k, double *d_identity,
                                double *d_cov_query, void *stream) {
    if (nq <= 0 || ns <= 0) return PANIB_OK;
    const int64_t n = nq * ns;
    ani_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_ov, ld_ov, d_q_counts, nq,
                                                                               d_s_counts, ns, k, d_identity,
                                                                               d_cov_query);
    return check_launch("ani_kernel");
}

extern "C" int panib_ani_host(const uint32_t *h_ov, int64_t ld_ov, const int32_t *h_q_counts, int64_t nq,
                              const int32_t *h_s_counts, int64_t ns, int k, double *h_identity,
                              double *h_cov_query) {
    // two libm pow per pair: 10^8 pairs at 10,000 genomes, so the query rows are spread over the host threads
    // the process may use (the values do not depend on the split: every pair is computed independently)
    auto rows = [&](int64_t q0, int64_t q1) {
        for (int64_t q = q0; q < q1; q++)
            for (int64_t s = 0; s < ns; s++)
                pair_ani(h_ov[q * ld_ov + s], h_q_counts[q], h_s_counts[s], k, h_identity[q * ns + s],
                         h_cov_query[q * ns + s]);
    };
    int threads = 1;
    if (nq * ns >= (1 << 16)) {
        cpu_set_t set;
        threads = sched_getaffinity(0, sizeof set, &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
        if (threads > 64) threads = 64;
        if (threads > nq) threads = (int)nq;
        if (threads < 1) threads = 1;
    }
    if (threads == 1) {
        rows(0, nq);
        return PANIB_OK;
    }
    std::vector<std::thread> pool;
    const int64_t per = (nq + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        const int64_t q0 = t * per, q1 = q0 + per < nq ? q0 + per : nq;
        if (q0 < q1) pool.emplace_back(rows, q0, q1);
    }
    for (auto &th : pool) th.join();
    return PANIB_OK;
}
