// hostpack.cpp -- host side of the ingest pipeline of libpanib200.so: ASCII base stream -> 2-bit packed words +
// validity mask on a persistent pool of host threads (AVX-512 / AVX2 / scalar, chosen at run time).
//
// Why on the host: the sketch path is fed over PCIe.  An ASCII base is 1 byte, its packed form 0.375 byte
// (2 bits + 1 validity bit), so packing BEFORE the copy cuts the bytes on the link by 2.7x (SURVEY.md 8f-1:
// "pack+mask emitted by the reader, pinned buffers, overlap with K1").  The reference reads FASTA through
// pyani_plus/utils.py:40-90 (fasta_bytes_iterator) and hands whole files to `sourmash scripts singlesketch`
// (pyani_plus/methods/sourmash.py:67-83); this is the ingest stage of the replacement.
//
// The packed format is the device format of pack.cuh / panib_pack_ascii: 16 bases per uint32, base i at
// bits 2i, A=0 C=1 G=2 T=3 (case-insensitive), invalid bases packed as 0 with their bit set in the mask
// (1 bit per base, 32 bases per uint32).  tests/test_abi.py checks the three code paths against each other;
// the GPU tests check the host form against panib_pack_ascii.
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include <pthread.h>
#include <sched.h>

#include "../../include/panib200.h"

namespace panib {
void set_error(const char *fmt, ...);

// ---- one block of 32*m bases ----------------------------------------------------------------------------
static void pack_block_scalar(const uint8_t *a, int64_t n_groups, uint32_t *packed, uint32_t *mask) {
    static uint8_t lut[256];
    static std::atomic<bool> ready{false};
    if (!ready.load(std::memory_order_acquire)) {
        uint8_t t[256];
        memset(t, 0x80, sizeof t);  // bit 7 = invalid
        t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3;
        memcpy(lut, t, sizeof t);
        ready.store(true, std::memory_order_release);
    }
    for (int64_t g = 0; g < n_groups; g++) {
        uint64_t p = 0;
        uint32_t inv = 0;
        for (int i = 0; i < 32; i++) {
            const uint8_t c = lut[a[32 * g + i]];
            p |= (uint64_t)(c & 3u) << (2 * i);
            inv |= (uint32_t)(c >> 7) << i;
        }
        packed[2 * g] = (uint32_t)p;
        packed[2 * g + 1] = (uint32_t)(p >> 32);
        mask[g] = inv;
    }
}

__attribute__((target("avx2"))) static void pack_block_avx2(const uint8_t *a, int64_t n_groups, uint32_t *packed,
                                                           uint32_t *mask) {
    const __m256i up = _mm256_set1_epi8((char)0xDF);
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'),
                  cT = _mm256_set1_epi8('T');
    const __m256i m3 = _mm256_set1_epi8(3), m1 = _mm256_set1_epi8(1);
    const __m256i w14 = _mm256_set1_epi16(0x0401);      // bytes (1, 4): c0 + 4 c1
    const __m256i w116 = _mm256_set1_epi32(0x00100001);  // words (1, 16): + 16 (c2 + 4 c3)
    const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                          0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    for (int64_t g = 0; g < n_groups; g++) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(a + 32 * g));
        const __m256i u = _mm256_and_si256(v, up);
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, cA), _mm256_cmpeq_epi8(u, cC)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(u, cG), _mm256_cmpeq_epi8(u, cT)));
        // code = ((u >> 1) & 3) ^ ((u >> 2) & 1), zero where invalid
        __m256i code = _mm256_xor_si256(_mm256_and_si256(_mm256_srli_epi16(u, 1), m3),
                                        _mm256_and_si256(_mm256_srli_epi16(u, 2), m1));
        code = _mm256_and_si256(code, ok);
        const __m256i b = _mm256_madd_epi16(_mm256_maddubs_epi16(code, w14), w116);  // one byte per 4 bases
        const __m256i q = _mm256_shuffle_epi8(b, pick);                              // 4 bytes per 128-bit lane
        packed[2 * g] = (uint32_t)_mm256_extract_epi32(q, 0);
        packed[2 * g + 1] = (uint32_t)_mm256_extract_epi32(q, 4);
        mask[g] = ~(uint32_t)_mm256_movemask_epi8(ok);
    }
}

// AVX-512: 256 bases per iteration.  Code and validity both come from 16-entry byte tables indexed by the low
// nibble of the character (VPSHUFB ignores bits 4-6, and a byte with bit 7 set selects 0): 'A' 0x41 -> 1,
// 'C' 0x43 -> 3, 'T' 0x54 -> 4, 'G' 0x47 -> 7, lower case alike.  A base is valid iff its upper-cased byte
// equals the letter its nibble stands for.  The four 16-byte results are written as one 64-byte line with a
// non-temporal store when the destination is aligned: the packed form is only ever read by the DMA engine, so
// it should neither be read for ownership nor stay in cache (the CALLER fences: pack_fence()).  One core's
// demand misses do not saturate DRAM, so the ASCII stream is prefetched 4 KB ahead (measured on a Sapphire
// Rapids host, 8 threads: 32 -> 49 GB/s of ASCII with the stores and the prefetch together).
template <bool kStream>
__attribute__((target("avx512f,avx512bw,avx512vl"))) static void pack_block_avx512_t(const uint8_t *a, int64_t n_groups,
                                                                                     uint32_t *packed, uint32_t *mask) {
    const __m512i up = _mm512_set1_epi8((char)0xDF);
    const __m512i lut_code = _mm512_broadcast_i32x4(_mm_setr_epi8(0, 0, 0, 1, 3, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i lut_chr = _mm512_broadcast_i32x4(
        _mm_setr_epi8(-1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1));
    const __m512i w14 = _mm512_set1_epi16(0x0401), w116 = _mm512_set1_epi32(0x00100001);
    const bool nt = kStream && ((uintptr_t)packed & 63) == 0;
    int64_t g = 0;
    for (; g + 8 <= n_groups; g += 8) {
        __m128i q[4];
        uint64_t inv[4];
        const char *pf = reinterpret_cast<const char *>(a) + 32 * g + 4096;  // prefetches never fault
        _mm_prefetch(pf, _MM_HINT_T0);
        _mm_prefetch(pf + 64, _MM_HINT_T0);
        _mm_prefetch(pf + 128, _MM_HINT_T0);
        _mm_prefetch(pf + 192, _MM_HINT_T0);
#pragma GCC unroll 4
        for (int i = 0; i < 4; i++) {
            const __m512i v = _mm512_loadu_si512(a + 32 * g + 64 * i);
            const __mmask64 ok = _mm512_cmpeq_epi8_mask(_mm512_and_si512(v, up), _mm512_shuffle_epi8(lut_chr, v));
            const __m512i code = _mm512_maskz_shuffle_epi8(ok, lut_code, v);
            q[i] = _mm512_cvtepi32_epi8(_mm512_madd_epi16(_mm512_maddubs_epi16(code, w14), w116));
            inv[i] = ~(uint64_t)ok;
        }
        __m512i line = _mm512_castsi128_si512(q[0]);
        line = _mm512_inserti32x4(line, q[1], 1);
        line = _mm512_inserti32x4(line, q[2], 2);
        line = _mm512_inserti32x4(line, q[3], 3);
        if (nt) _mm512_stream_si512(reinterpret_cast<__m512i *>(packed + 2 * g), line);
        else _mm512_storeu_si512(packed + 2 * g, line);
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(mask + g),
                            _mm256_set_epi64x((long long)inv[3], (long long)inv[2], (long long)inv[1],
                                              (long long)inv[0]));
    }
    if (g < n_groups) pack_block_scalar(a + 32 * g, n_groups - g, packed + 2 * g, mask + g);
}
// streaming form: for a destination that is written once and read by the DMA engine much later (a whole-stream
// buffer); cached form: for a small ring that is re-used while it is still in the last-level cache, so that
// neither the stores nor the DMA reads go to DRAM
static void pack_block_avx512(const uint8_t *a, int64_t n_groups, uint32_t *packed, uint32_t *mask) {
    pack_block_avx512_t<true>(a, n_groups, packed, mask);
}
static void pack_block_avx512_cached(const uint8_t *a, int64_t n_groups, uint32_t *packed, uint32_t *mask) {
    pack_block_avx512_t<false>(a, n_groups, packed, mask);
}
// non-temporal stores are weakly ordered: fence before the packed words are handed to another thread / the DMA engine
static inline void pack_fence() { _mm_sfence(); }


using PackFn = void (*)(const uint8_t *, int64_t, uint32_t *, uint32_t *);
static PackFn choose_pack(int force, bool cached = false) {
    // force: 0 = best available, 1 = scalar, 2 = AVX2, 3 = AVX-512 (tests); unavailable -> nullptr
    __builtin_cpu_init();
    if (force == 0) {  // PANIB_PACK_ISA=scalar|avx2|avx512 pins the pool's code path (diagnostics)
        const char *e = getenv("PANIB_PACK_ISA");
        if (e) force = !strcmp(e, "scalar") ? 1 : !strcmp(e, "avx2") ? 2 : !strcmp(e, "avx512") ? 3 : 0;
    }
    const bool has512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                        __builtin_cpu_supports("avx512vl");
    const bool has2 = __builtin_cpu_supports("avx2");
    switch (force) {
    case 1: return pack_block_scalar;
    case 2: return has2 ? pack_block_avx2 : nullptr;
    case 3: return has512 ? pack_block_avx512 : nullptr;
    default: return has512 ? (cached ? pack_block_avx512_cached : pack_block_avx512) : has2 ? pack_block_avx2 : pack_block_scalar;
    }
}

// ---- persistent worker pool --------------------------------------------------------------------------------
// A job is a range of equal blocks; workers and the submitting thread draw block indices from one atomic
// counter.  done[] lets a consumer wait for a PREFIX of the blocks (the ingest pipeline copies chunk c while
// the pool packs chunk c+1).
class Pool {
public:
    static Pool &get() {
        static Pool p;
        return p;
    }
    int size() const { return (int)workers_.size() + 1; }

    struct Job {
        const uint8_t *ascii;
        uint32_t *packed, *mask;
        int64_t n_groups, groups_per_block, n_blocks;
        PackFn fn;
        // ingest pipeline only (sketch.cu): per-tile dirty flags instead of a dense mask, and chunk ownership
        uint8_t *tile_dirty = nullptr;  // [tiles]: 1 = the tile holds an invalid base and its 128 mask words were written
        int64_t blocks_per_chunk = 0;   // > 0: chunk_state[] decides who handles a chunk
        std::vector<std::atomic<uint8_t>> chunk_state;  // 0 free, 1 being packed by the pool, 2 taken raw by the submitter
        // packed output as a RING of ring_blocks blocks (0 = the whole stream, block b at its own offset): block b may
        // be written once every block < b - ring_blocks + 1 has been copied away, which the submitter reports in
        // free_upto (blocks < free_upto are free again)
        int64_t ring_blocks = 0;
        std::atomic<int64_t> free_upto{0};
        std::atomic<bool> abort{false};
        std::atomic<int64_t> next{0}, finished{0};
        std::vector<std::atomic<uint8_t>> done;
        Job(int64_t blocks, int64_t chunks) : chunk_state((size_t)chunks), done((size_t)blocks) {
            for (auto &d : done) d.store(0, std::memory_order_relaxed);
            for (auto &c : chunk_state) c.store(0, std::memory_order_relaxed);
        }
        bool run_one() {
            const int64_t b = next.fetch_add(1, std::memory_order_relaxed);
            if (b >= n_blocks) return false;
            bool skip = false;
            if (blocks_per_chunk > 0) {  // the first block of a chunk to get here claims the chunk for the pool
                auto &st = chunk_state[(size_t)(b / blocks_per_chunk)];
                uint8_t cur = st.load(std::memory_order_acquire);
                if (cur == 0 && st.compare_exchange_strong(cur, 1, std::memory_order_acq_rel)) cur = 1;
                skip = cur == 2;
            }
            if (!skip && ring_blocks > 0) {  // wait for the ring slot (never on the submitting thread: it frees them)
                while (b >= free_upto.load(std::memory_order_acquire) + ring_blocks) {
                    if (abort.load(std::memory_order_relaxed)) { skip = true; break; }
                    _mm_pause();
                }
            }
            if (!skip) {
                const int64_t g0 = b * groups_per_block;
                const int64_t ng = g0 + groups_per_block <= n_groups ? groups_per_block : n_groups - g0;
                // where the packed words of group g go: their own offset, or the block's slot of the ring
                uint32_t *const pk = ring_blocks > 0 ? packed + 2 * ((b % ring_blocks) * groups_per_block - g0) : packed;
                if (!tile_dirty) {
                    fn(ascii + 32 * g0, ng, pk + 2 * g0, mask + g0);
                } else {  // tile by tile: the mask of a clean tile is never written
                    constexpr int64_t kTileGroups = PANIB_TILE_BASES / 32;
                    alignas(64) uint32_t local[kTileGroups];
                    for (int64_t g = g0; g < g0 + ng; g += kTileGroups) {
                        fn(ascii + 32 * g, kTileGroups, pk + 2 * g, local);
                        uint32_t any = 0;
                        for (int i = 0; i < kTileGroups; i++) any |= local[i];
                        tile_dirty[g / kTileGroups] = any ? 1 : 0;
                        if (any) memcpy(mask + g, local, sizeof local);
                    }
                }
                pack_fence();
            }
            done[(size_t)b].store(1, std::memory_order_release);
            finished.fetch_add(1, std::memory_order_release);
            return true;
        }
    };

    // start `job` on up to `threads` threads (workers only; the caller may help with job->run_one())
    void start(Job *job, int threads) {
        std::lock_guard<std::mutex> lk(mu_);
        job_ = job;
        want_ = threads - 1 < (int)workers_.size() ? (threads > 1 ? threads - 1 : 0) : (int)workers_.size();
        generation_++;
        active_ = want_;
        cv_.notify_all();
    }
    // wait until every worker has left the job (the caller has seen all blocks finished)
    void finish(Job *job) {
        while (job->finished.load(std::memory_order_acquire) < job->n_blocks) {
            if (!job->run_one()) std::this_thread::yield();
        }
        std::unique_lock<std::mutex> lk(mu_);
        idle_cv_.wait(lk, [&] { return active_ == 0; });
        job_ = nullptr;
    }

private:
    // One pool per process.  Under a one-process-per-GPU launcher (LOCAL_RANK / LOCAL_WORLD_SIZE set, as
    // torchrun does) a process takes its own share of the CPUs it may run on, so that the ranks' pools do not
    // sit on each other.  Workers are pinned to one CPU each: freshly woken threads otherwise stay on the
    // waker's CPU until the scheduler's load balancer moves them, which on a virtualised host took about a
    // second (measured: 4 threads packed at the speed of 1 for the first ~15 calls).  PANIB_PIN=0 disables.
    Pool() {
        std::vector<int> cpus;
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof set, &set) == 0)
            for (int c = 0; c < CPU_SETSIZE; c++)
                if (CPU_ISSET(c, &set)) cpus.push_back(c);
        if (cpus.empty()) {
            int n = (int)std::thread::hardware_concurrency();
            for (int c = 0; c < (n > 0 ? n : 1); c++) cpus.push_back(c);
        }
        const char *lr = getenv("LOCAL_RANK"), *lw = getenv("LOCAL_WORLD_SIZE"), *pin = getenv("PANIB_PIN");
        const int world = lw ? atoi(lw) : 1, rank = lr ? atoi(lr) : 0;
        if (world > 1 && rank >= 0 && rank < world && (int)cpus.size() >= world) {
            const size_t lo = cpus.size() * (size_t)rank / (size_t)world, hi = cpus.size() * (size_t)(rank + 1) / (size_t)world;
            cpus = std::vector<int>(cpus.begin() + (long)lo, cpus.begin() + (long)hi);
        }
        if (cpus.size() > 64) cpus.resize(64);
        const bool do_pin = !(pin && pin[0] == '0');
        const int n = (int)cpus.size();
        for (int i = 0; i < n - 1; i++) {
            workers_.emplace_back([this, i] { loop(i); });
            if (do_pin) {  // worker i on the (i+1)-th CPU of the share; the submitting thread is left alone
                cpu_set_t one;
                CPU_ZERO(&one);
                CPU_SET(cpus[(size_t)i + 1], &one);
                pthread_setaffinity_np(workers_.back().native_handle(), sizeof one, &one);
            }
        }
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            cv_.notify_all();
        }
        for (auto &t : workers_) t.join();
    }
    void loop(int idx) {
        uint64_t seen = 0;
        for (;;) {
            Job *job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                if (idx >= want_) continue;
                job = job_;
            }
            while (job->run_one()) {
            }
            std::lock_guard<std::mutex> lk(mu_);
            if (--active_ == 0) idle_cv_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, idle_cv_;
    std::vector<std::thread> workers_;
    Job *job_ = nullptr;
    uint64_t generation_ = 0;
    int want_ = 0, active_ = 0;
    bool stop_ = false;
};

static std::mutex g_pack_mutex;  // one packing job at a time per process

// Asynchronous packing job used by the ingest pipeline (sketch.cu).  Blocks are `bases_per_block` bases (whole
// tiles), chunks are `blocks_per_chunk` blocks.  The pool packs chunks from the front; the submitting thread may
// take chunks that no pool thread has touched yet for itself (host_pack_claim_raw: those go over PCIe as ASCII
// and are packed by the GPU).  With tile_dirty != NULL the mask words of a tile are written to h_mask only when
// the tile holds an invalid base.
struct HostPackJob {
    Pool::Job job;
    std::unique_lock<std::mutex> lock;
    std::vector<int64_t> cursor;  // per chunk: blocks [first, cursor) are known to be done
    HostPackJob(int64_t blocks, int64_t chunks) : job(blocks, chunks), lock(g_pack_mutex), cursor((size_t)chunks, 0) {}
};

HostPackJob *host_pack_start(const uint8_t *h_ascii, int64_t n_bases, uint32_t *h_packed, uint32_t *h_mask,
                             int64_t bases_per_block, int threads, uint8_t *tile_dirty, int64_t blocks_per_chunk,
                             int64_t ring_blocks) {
    const int64_t n_groups = n_bases / 32, gpb = bases_per_block / 32;
    const int64_t blocks = (n_groups + gpb - 1) / gpb;
    const int64_t chunks = blocks_per_chunk > 0 ? (blocks + blocks_per_chunk - 1) / blocks_per_chunk : 0;
    auto *hp = new HostPackJob(blocks, chunks);
    hp->job.ascii = h_ascii;
    hp->job.packed = h_packed;
    hp->job.mask = h_mask;
    hp->job.n_groups = n_groups;
    hp->job.groups_per_block = gpb;
    hp->job.n_blocks = blocks;
    hp->job.fn = choose_pack(0, ring_blocks > 0);
    hp->job.ring_blocks = ring_blocks;
    hp->job.tile_dirty = tile_dirty;
    hp->job.blocks_per_chunk = blocks_per_chunk;
    Pool &pool = Pool::get();
    pool.start(&hp->job, threads > 0 ? threads : pool.size());
    return hp;
}
void host_pack_wait_prefix(HostPackJob *hp, int64_t blocks) {
    if (blocks > hp->job.n_blocks) blocks = hp->job.n_blocks;
    for (int64_t b = 0; b < blocks; b++) {
        while (!hp->job.done[(size_t)b].load(std::memory_order_acquire)) {
            if (!hp->job.run_one()) std::this_thread::yield();  // help, then spin politely
        }
    }
}
// all blocks of chunk c packed?  (never true for a chunk taken raw: its blocks are flagged done without packing,
// the submitter does not ask)
bool host_pack_chunk_ready(HostPackJob *hp, int64_t c) {
    const int64_t bpc = hp->job.blocks_per_chunk;
    const int64_t b1 = (c + 1) * bpc < hp->job.n_blocks ? (c + 1) * bpc : hp->job.n_blocks;
    int64_t &cur = hp->cursor[(size_t)c];
    if (cur < c * bpc) cur = c * bpc;
    while (cur < b1 && hp->job.done[(size_t)cur].load(std::memory_order_acquire)) cur++;
    return cur >= b1;
}
// take chunk c out of the pool's hands (only succeeds while no pool thread has started on it)
bool host_pack_claim_raw(HostPackJob *hp, int64_t c) {
    uint8_t expect = 0;
    return hp->job.chunk_state[(size_t)c].compare_exchange_strong(expect, 2, std::memory_order_acq_rel);
}
// the submitting thread lends a hand -- except in ring mode, where it is the one that frees the slots the
// packers wait for and must not wait itself
bool host_pack_help(HostPackJob *hp) { return hp->job.ring_blocks > 0 ? false : hp->job.run_one(); }
// blocks [b0, b1) all packed?
bool host_pack_blocks_ready(HostPackJob *hp, int64_t b0, int64_t b1) {
    for (int64_t b = b0; b < b1; b++)
        if (!hp->job.done[(size_t)b].load(std::memory_order_acquire)) return false;
    return true;
}
// ring mode: every block below `blocks` has been copied away (or will never be packed)
void host_pack_release(HostPackJob *hp, int64_t blocks) { hp->job.free_upto.store(blocks, std::memory_order_release); }
void host_pack_finish(HostPackJob *hp) {
    // an early exit (error path) must not leave workers waiting for ring slots
    if (hp->job.finished.load(std::memory_order_acquire) < hp->job.n_blocks) hp->job.abort.store(true);
    Pool::get().finish(&hp->job);
    delete hp;
}

}  // namespace panib

using namespace panib;

extern "C" __attribute__((visibility("default"))) int panib_host_threads(void) { return Pool::get().size(); }

extern "C" __attribute__((visibility("default"))) int panib_pack_host(const uint8_t *h_ascii, int64_t n_bases,
                                                                      uint32_t *h_packed, uint32_t *h_mask,
                                                                      int threads) {
    if (n_bases < 0 || (n_bases & 31) || (n_bases && (!h_ascii || !h_packed || !h_mask))) {
        set_error("panib_pack_host: n_bases=%lld must be a non-negative multiple of 32 with non-NULL buffers",
                  (long long)n_bases);
        return PANIB_E_ARG;
    }
    if (n_bases == 0) return PANIB_OK;
    if (threads < 0) {  // tests: -1 scalar, -2 AVX2, -3 AVX-512 on the calling thread
        PackFn fn = choose_pack(-threads);
        if (!fn) {
            set_error("panib_pack_host: this CPU lacks the requested instruction set");
            return PANIB_E_ARG;
        }
        fn(h_ascii, n_bases / 32, h_packed, h_mask);
        pack_fence();
        return PANIB_OK;
    }
    HostPackJob *hp = host_pack_start(h_ascii, n_bases, h_packed, h_mask, 1 << 18, threads, nullptr, 0, 0);
    host_pack_finish(hp);
    return PANIB_OK;
}

// sparse-mask form of panib_pack_host (what the ingest pipeline runs): h_tile_dirty[t] = 1 when tile t holds an
// invalid base; only then are the tile's PANIB_TILE_BASES/32 mask words written to h_mask.
extern "C" __attribute__((visibility("default"))) int panib_pack_host_tiles(const uint8_t *h_ascii, int64_t n_bases,
                                                                            uint32_t *h_packed, uint32_t *h_mask,
                                                                            uint8_t *h_tile_dirty, int threads) {
    if (n_bases < 0 || (n_bases % PANIB_TILE_BASES) || (n_bases && (!h_ascii || !h_packed || !h_mask || !h_tile_dirty))) {
        set_error("panib_pack_host_tiles: n_bases=%lld must be a non-negative multiple of %d with non-NULL buffers",
                  (long long)n_bases, PANIB_TILE_BASES);
        return PANIB_E_ARG;
    }
    if (n_bases == 0) return PANIB_OK;
    HostPackJob *hp = host_pack_start(h_ascii, n_bases, h_packed, h_mask, 64 * PANIB_TILE_BASES, threads, h_tile_dirty, 0, 0);
    host_pack_finish(hp);
    return PANIB_OK;
}
