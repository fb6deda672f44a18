/*
 * panib200.h -- C ABI of libpanib200.so, the B200 (sm_100a) engine behind pyani-plus's
 * `sourmash` method.  Plain pointers and sizes only; no torch / C++ types.
 *
 * What it replaces.  In the reference the arithmetic of this path is done by two external
 * commands spawned from Python:
 *   sourmash scripts singlesketch -I DNA -p k=K,scaled=S    pyani_plus/methods/sourmash.py:67-83
 *   sourmash scripts manysearch   -m DNA -t 0               pyani_plus/methods/sourmash.py:184-200
 * (called from prepare_genomes :34-84 and compute_sourmash_tile :147-206, driven by
 * private_cli.prepare private_cli.py:714-754 and private_cli.compute_sourmash :1803-1902).
 * The entry points below are what a reference-side binding (ctypes) would
 * call instead of those two subprocesses.
 *
 * Conventions
 *  - every function returns 0 on success, a negative PANIB_E_* code on failure; the text of the
 *    last failure on the calling thread is available from panib_last_error();
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all device work is
 *    enqueued on it and the call returns without synchronising unless stated otherwise;
 *  - pointers named d_* are device pointers (the caller owns all memory -- the Python host
 *    allocates them as torch tensors), h_* are host pointers (pinned for async copies);
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Device data layout
 *  base stream   All genomes of a batch are laid out in one stream of bases cut into tiles of
 *                PANIB_TILE_BASES positions.  Genome g owns tiles [tile_off[g], tile_off[g+1]);
 *                its records are written back to back with ONE invalid separator base between
 *                records, and the rest of its last tile is invalid padding (at least one base),
 *                so no k-mer window can span two records or two genomes.  The stream is followed
 *                by one extra all-invalid tile (halo for the last real tile).
 *                ASCII form: 1 byte per base ('N' or any non-ACGT byte = invalid).
 *                Packed form: 2 bits per base, 16 bases per uint32, base i of a word at bits
 *                [2i+1:2i], A=0 C=1 G=2 T=3; plus a validity mask, 1 bit per base, 32 bases per
 *                uint32, bit set = invalid.
 *  sketch table  Row g (row_stride uint64 slots, row_stride = max nb[g] * PANIB_BUCKET_SLOTS)
 *                first serves as nb[g] value-range buckets of open-addressing hash sets while
 *                k-mers are hashed, and is then sorted / de-duplicated / compacted IN PLACE so
 *                that row g holds counts[g] ascending distinct hashes starting at slot 0.
 *  overlap       ov[q * ld + s] = |sketch_q n sketch_s| as uint32.
 */
#ifndef PANIB200_H
#define PANIB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PANIB_API __attribute__((visibility("default")))
#else
#define PANIB_API
#endif

#define PANIB_VERSION_STRING "0.1.0"
#define PANIB_TILE_BASES 4096      /* k-mer start positions per tile / per CTA of the sketch kernel */
#define PANIB_BUCKET_SLOTS 1024    /* uint64 slots per value-range bucket of a sketch-table row    */
#define PANIB_MAX_FAST_K 32        /* k handled by the register-resident kernel (k=21,31 built)     */
#define PANIB_MAX_K 255

#define PANIB_OK 0
#define PANIB_E_CUDA (-1)          /* a CUDA runtime call or kernel launch failed                   */
#define PANIB_E_ARG (-2)           /* invalid argument                                             */
#define PANIB_E_NODEVICE (-3)      /* no CUDA device: this library has no CPU path                  */

/* d_status is int32[4], zero-initialised by the caller: [0] = PANIB_ST_* bits OR-ed in by kernels,
 * [1] = largest sketch size seen by the finalize kernel (atomicMax), [2] = scratch of the sketch kernel
 * (ticket counter of its dynamic tile schedule; zeroed by the library before each launch, so one status
 * block must not be shared by sketch calls running concurrently on different streams), [3] reserved.
 * The finalize kernel also stores each row's size in the row's LAST slot (row[row_stride-1]), so a
 * single all-gather of the rows carries the sizes along. */
#define PANIB_ST_BUCKET_OVERFLOW 1 /* a sketch bucket filled up: re-run with more buckets           */
#define PANIB_ST_SEGMENT_OVERFLOW 2/* a pairwise segment exceeded seg_cap: re-run with more cells    */
#define PANIB_ST_INDEX_OVERFLOW 4  /* a rank's slice of the inverted index outgrew its table: re-run unsharded */

/* ---- library ------------------------------------------------------------------------------ */
PANIB_API const char *panib_version(void);                    /* "0.1.0 sm_100a ..."                        */
PANIB_API int panib_last_error(char *buf, size_t n);           /* copies the last error text; returns length */
PANIB_API int panib_device_count(void);                        /* 0 when no usable CUDA device               */
PANIB_API uint64_t panib_launch_count(void);                   /* kernels launched by this library so far    */

/* max_hash of a FracMinHash `scaled`: 0->0, 1->2^64-1, else (uint64)(2^64 / (double)scaled).
 * Replaces sourmash's max_hash_for_scaled; pinned by the "max_hash" field of the fixture .sig files. */
PANIB_API uint64_t panib_max_hash(uint64_t scaled);

/* Bucket plan for one genome with n_kmers candidate windows: number of buckets nb (>=1) and the
 * multiplier bmul with bucket(h) = mulhi64(h, bmul) (monotone in h, < nb for h <= max_hash). */
PANIB_API int panib_plan_buckets(int64_t n_kmers, uint64_t scaled, double slack, int32_t *nb, uint64_t *bmul);

/* ---- ingest (host): FASTA text -> base-stream form ------------------------------------------ */
/* Parses decompressed FASTA text exactly as pyani_plus/utils.py:40-90 (fasta_bytes_iterator) does and
 * writes the records' sequences back to back with one 'N' between records.  dst may be NULL to only
 * measure; n bytes are always enough for dst.  Returns the stream-form length, or PANIB_E_ARG.  out4 = n_records, total_bases,
 * offset and length (in text) of the first record's title. */
PANIB_API int64_t panib_fasta_to_stream(const uint8_t *text, int64_t n, uint8_t *dst, int64_t dst_cap,
                                        int64_t *out4);

/* ---- results (host): decimal text of a sketch ---------------------------------------------- */
/* Writes n unsigned 64-bit integers in decimal, `sep` between them (sep == 0: back to back): the "mins" array
 * of a sourmash signature file and the input of its "md5sum" (what `sourmash scripts singlesketch` writes for
 * pyani_plus/methods/sourmash.py:57-83).  Returns the bytes written, or PANIB_E_ARG if dst_cap is too small
 * (21 * n always suffices). */
PANIB_API int64_t panib_format_u64(const uint64_t *values, int64_t n, int sep, char *dst, int64_t dst_cap);

/* ---- stage 0: ASCII base stream -> packed 2-bit + validity mask --------------------------- */
/* n_bases must be a multiple of 32.  Upper-cases; any byte other than A,C,G,T is invalid. */
PANIB_API int panib_pack_ascii(const uint8_t *d_ascii, int64_t n_bases, uint32_t *d_packed, uint32_t *d_mask,
                     void *stream);

/* ---- stage 1 (kernel K1): FracMinHash sketching -- replaces `sourmash scripts singlesketch` -- */
/* d_packed/d_mask: stream of (n_tiles+1) tiles.  d_tile_off[n_genomes+1]: int64 tile offsets.
 * d_nb[n_genomes], d_bmul[n_genomes]: bucket plan.  d_table: n_genomes*row_stride uint64, need not be
 * initialised.  d_flags[n_genomes]: int32 scratch.  On completion (stream order) row g holds the
 * sorted distinct hashes h = murmur3_x64_128(canonical k-mer, seed).h1 with 0 < h <= max_hash and
 * d_counts[g] (int32) their number.  d_status: int32, PANIB_ST_* bits are OR-ed in.            */
PANIB_API int panib_sketch_stream(const uint32_t *d_packed, const uint32_t *d_mask, const int64_t *d_tile_off,
                        int64_t n_genomes, int64_t n_tiles, int k, uint32_t seed, uint64_t max_hash,
                        const int32_t *d_nb, const uint64_t *d_bmul, uint64_t *d_table,
                        int64_t row_stride, int32_t *d_counts, int32_t *d_flags, int32_t *d_status,
                        void *stream);

/* Optional scratch for the sketch calls of the CURRENT device (d_ptr stays owned by the caller and must outlive
 * them; NULL / 0 unregisters; sketch calls of one device that use it must not run concurrently).  With it K1
 * parks the surviving hashes per CTA and a second kernel inserts them into the table rows afterwards, instead
 * of inserting one by one from the hashing loop; results are identical.  2 * 12 bytes per expected survivor
 * (bases / scaled) of the largest single call, plus a few KB, is plenty; a smaller one is simply not used. */
PANIB_API int panib_set_workspace(void *d_ptr, int64_t bytes);

/* The two halves of panib_sketch_stream, exposed for profiling / benchmarking. */
PANIB_API int panib_sketch_hash_only(const uint32_t *d_packed, const uint32_t *d_mask, const int64_t *d_tile_off,
                           int64_t n_genomes, int64_t n_tiles, int k, uint32_t seed, uint64_t max_hash,
                           const int32_t *d_nb, const uint64_t *d_bmul, uint64_t *d_table,
                           int64_t row_stride, int32_t *d_flags, int32_t *d_status, void *stream);
PANIB_API int panib_sketch_finalize(uint64_t *d_table, int64_t row_stride, int64_t n_genomes, const int32_t *d_nb,
                          int32_t *d_counts, const int32_t *d_flags, int32_t *d_status, void *stream);

/* Multi-GPU finalize FUSED with the all-gather of sketches: sorts the buckets of this rank's rows, then
 * writes every genome's sorted sketch straight into row (rank*per_rank + g) of the gathered table of
 * every rank through peer-mapped pointers (h_peer_tables[world], host array of device pointers that are
 * valid on this device: CUDA IPC / symmetric memory), size in the row's last slot.  The caller issues a
 * cross-rank barrier afterwards (multi_gpu.SymmetricGather).  Replaces panib_sketch_finalize + an NCCL
 * all-gather of the rows. */
PANIB_API int panib_sketch_finalize_gather(uint64_t *d_table, int64_t row_stride, int64_t n_genomes,
                                 const int32_t *d_nb, int32_t *d_counts, const int32_t *d_flags,
                                 int32_t *d_status, const uint64_t *const *h_peer_tables, int world, int rank,
                                 int64_t per_rank, void *stream);

/* Host-buffer form (what a caller holding FASTA bytes uses): copies the pinned ASCII stream
 * host->device on `stream`, packs it and sketches it.  d_ascii is device scratch of n_bases bytes,
 * n_bases = (n_tiles+1)*PANIB_TILE_BASES. */
PANIB_API int panib_sketch_ascii_host(const uint8_t *h_ascii, uint8_t *d_ascii, int64_t n_bases,
                            uint32_t *d_packed, uint32_t *d_mask, const int64_t *d_tile_off,
                            int64_t n_genomes, int64_t n_tiles, int k, uint32_t seed, uint64_t max_hash,
                            const int32_t *d_nb, const uint64_t *d_bmul, uint64_t *d_table,
                            int64_t row_stride, int32_t *d_counts, int32_t *d_flags, int32_t *d_status,
                            void *stream);

/* As panib_sketch_ascii_host without the finalize step (rows stay bucketed hash sets). */
PANIB_API int panib_sketch_ascii_host_hash_only(const uint8_t *h_ascii, uint8_t *d_ascii, int64_t n_bases,
                                      uint32_t *d_packed, uint32_t *d_mask, const int64_t *d_tile_off,
                                      int64_t n_genomes, int64_t n_tiles, int k, uint32_t seed, uint64_t max_hash,
                                      const int32_t *d_nb, const uint64_t *d_bmul, uint64_t *d_table,
                                      int64_t row_stride, int32_t *d_flags, int32_t *d_status, void *stream);

/* ---- ingest (host): ASCII base stream -> packed 2-bit + validity mask on host threads ---------- */
/* The host form of panib_pack_ascii (same packed words and mask bits), run on the library's pool of host
 * threads (threads = 0: all the process may use; panib_host_threads() tells how many that is).  This is
 * the reader side of the replacement for the reference's per-file hand-over of FASTA bytes to
 * `sourmash scripts singlesketch` (pyani_plus/methods/sourmash.py:67-83; pyani_plus/utils.py:40-90):
 * packing before the copy puts 0.375 instead of 1 byte per base on PCIe.  AVX-512 / AVX2 / scalar code
 * chosen at run time; threads = -1 / -2 / -3 runs the scalar / AVX2 / AVX-512 form on the calling thread
 * (tests).  n_bases must be a multiple of 32. */
PANIB_API int panib_host_threads(void);
PANIB_API int panib_pack_host(const uint8_t *h_ascii, int64_t n_bases, uint32_t *h_packed, uint32_t *h_mask,
                    int threads);
/* Sparse-mask form (what the ingest pipeline runs on its pool): n_bases a multiple of PANIB_TILE_BASES;
 * h_tile_dirty[t] = 1 when tile t holds an invalid base, and only then are the tile's PANIB_TILE_BASES/32 mask
 * words written to h_mask (a clean tile's mask is all zero and is never touched). */
PANIB_API int panib_pack_host_tiles(const uint8_t *h_ascii, int64_t n_bases, uint32_t *h_packed, uint32_t *h_mask,
                          uint8_t *h_tile_dirty, int threads);

/* Packed host-buffer form of panib_sketch_stream: the ingest pipeline.  h_packed / h_mask are (pinned) host
 * buffers of n_bases/16 and n_bases/32 words.  With h_ascii == NULL they already hold the packed stream and are
 * copied as they are.  With h_ascii != NULL the ASCII stream is packed into them by the host threads chunk by
 * chunk, overlapped with the host->device copies of the packed chunks and with K1 on the chunks already on the
 * device; h_packed / h_mask are then scratch (their content after the call is unspecified).
 * d_scratch (device, scratch_bytes; may be NULL / 0) widens the pipeline, panib_ingest_scratch_bytes(n_bases)
 * tells how much serves everything:
 *   - with room for the mask staging ring the validity mask is sent sparsely: only the 512-byte masks of tiles
 *     that hold an invalid base cross PCIe (0.25 instead of 0.375 byte per base), a kernel scatters them;
 *   - with room for at least one more chunk of ASCII, and h_ascii page-locked, chunks are also taken from the
 *     END of the stream as plain ASCII whenever the link would otherwise wait for the host threads, and packed
 *     on the GPU: host cores and link both stay busy whatever their relative speeds.
 * The host threads write the packed words either to their place in h_packed (streaming stores; the DMA engine
 * reads them back from DRAM) or through a 16 MB ring at the start of h_packed that stays in the last-level cache
 * (no DRAM traffic for the packed form at all).  Which is faster depends on the host, so per calling thread and
 * stream size the first call warms up, the second and third time one form each, and later calls use the faster;
 * PANIB_INGEST_RING_MB=0 / N pins the choice.
 * d_counts == NULL skips the finalize step (rows stay bucketed hash sets, as after panib_sketch_hash_only).
 * The call returns when all host work is done and all device work is enqueued on `stream`.  The sketches do
 * not depend on which way a chunk travelled. */
PANIB_API int64_t panib_ingest_scratch_bytes(int64_t n_bases);
/* What the calling thread's last panib_sketch_packed_host moved: out6 = {bytes copied host->device, chunks,
 * chunks that travelled as ASCII, tiles whose mask was sent, bytes of the packed-word ring it used (0 = the
 * whole-stream buffer), 1 if the call was one of the two that time the two forms against each other}. */
PANIB_API int panib_ingest_last(int64_t *out6);
PANIB_API int panib_sketch_packed_host(const uint8_t *h_ascii, uint32_t *h_packed, uint32_t *h_mask, int64_t n_bases,
                             uint8_t *d_scratch, int64_t scratch_bytes,
                             uint32_t *d_packed, uint32_t *d_mask, const int64_t *d_tile_off,
                             int64_t n_genomes, int64_t n_tiles, int k, uint32_t seed, uint64_t max_hash,
                             const int32_t *d_nb, const uint64_t *d_bmul, uint64_t *d_table,
                             int64_t row_stride, int32_t *d_counts, int32_t *d_flags, int32_t *d_status,
                             int host_threads, void *stream);

/* ---- stage 2 (kernel K2): all-vs-all sorted-sketch intersection -- replaces `manysearch` ---- */
/* Queries are rows of (d_q_rows, d_q_counts, q_stride), subjects rows of (d_s_rows, ...); they may be
 * the same table.  symmetric != 0 requires the same table and computes only q < s, mirroring the
 * result and writing the diagonal (= sketch size).  n_cells / seg_cap / idx_buckets choose the
 * shared-memory segmentation; pass 0 for all three to let the library choose from
 * max_count (the largest sketch size).  d_fence: int32 scratch of (nq+ns)*(n_cells+1) entries
 * (unused when n_cells == 1, may be NULL).  d_ov: uint32 [nq x ld_ov], fully overwritten.
 * Multi-GPU: rank r of world w processes the work items it owns (item % w == r) and leaves the
 * rest of d_ov zero, so the ranks' matrices sum to the full result. */
PANIB_API int panib_intersect(const uint64_t *d_q_rows, const int32_t *d_q_counts, int64_t q_stride, int64_t nq,
                    const uint64_t *d_s_rows, const int32_t *d_s_counts, int64_t s_stride, int64_t ns,
                    int symmetric, uint64_t max_hash, int64_t max_count, int n_cells, int seg_cap,
                    int idx_buckets, int32_t *d_fence, uint32_t *d_ov, int64_t ld_ov, int rank, int world,
                    int32_t *d_status, void *stream);
/* Size (in int32 entries) of the d_fence scratch panib_intersect needs for these arguments. */
PANIB_API int64_t panib_intersect_fence_entries(int64_t nq, int64_t ns, uint64_t max_hash, int64_t max_count,
                                      int n_cells, int seg_cap);

/* ---- stage 2, inverted-index form (csrc/index.cu): same counts as panib_intersect(symmetric=1), cost
 * proportional to what the genomes SHARE instead of N^2 x sketch size.  Replaces the same reference step
 * (pyani_plus/methods/sourmash.py:184-200).  A sketch is a duplicate-free set, so the genomes holding a hash
 * are found by a group-by over all (hash, genome) entries: a hand-written open-addressing hash table in
 * global memory (atomicCAS on the key, atomicAdd for the entry's arrival index in its group; no sort, no
 * library).  Hashes held by >= tau genomes become columns of a bit matrix (AND+POPC over all genome pairs),
 * rarer shared hashes are expanded pair by pair from their member lists.
 * Entries: exact (d_offsets = int64[n+1] prefix sums of the sketch sizes, entries = their total) or padded
 * (d_offsets NULL, entries = n * cap: needs no host knowledge of the sizes, so it can be captured in a CUDA
 * graph; cap >= the largest sketch, a larger one raises PANIB_ST_SEGMENT_OVERFLOW).  In the exact form cap
 * only sizes the launch grid.  d_work: scratch of panib_index_workspace_bytes(n, entries, tau, world) bytes,
 * d_stats: uint64[4] written by panib_index_build = {bit-matrix columns, pairs expanded from rare hashes,
 * distinct hashes, member-list words} of THIS rank, so that the caller can compare the cost with the probing
 * kernel before calling panib_index_count with the SAME n / cap / entries / d_offsets / tau / rank / world /
 * d_work.  Requires entries < 2^31 and max_hash < 2^64 - 16 (scaled >= 2).  Offsets that do not match the
 * sizes raise PANIB_ST_SEGMENT_OVERFLOW in d_status.
 * Multi-GPU: the HASH RANGE is sharded -- rank r of world indexes only the hashes in its 1/world slice of
 * [0, max_hash], so every step shrinks with the number of ranks; the ranks' matrices sum to the result (as
 * panib_intersect).  A slice holding more than twice its share of the entries raises
 * PANIB_ST_INDEX_OVERFLOW (the caller then uses another form). */
PANIB_API int panib_index_workspace_bytes(int64_t n, int64_t entries, int tau, int world, int64_t *bytes);
PANIB_API int panib_index_build(const uint64_t *d_rows, const int32_t *d_counts, int64_t stride, int64_t n,
                      uint64_t max_hash, int64_t cap, int64_t entries, const int64_t *d_offsets, int tau,
                      int rank, int world, void *d_work, int64_t work_bytes, uint64_t *d_stats,
                      int32_t *d_status, void *stream);
PANIB_API int panib_index_count(const int32_t *d_counts, int64_t n, uint64_t max_hash, int64_t cap, int64_t entries,
                      const int64_t *d_offsets, int tau, void *d_work, int64_t work_bytes,
                      const uint64_t *d_stats, uint32_t *d_ov, int64_t ld_ov, int rank, int world,
                      int32_t *d_status, void *stream);

/* ---- stage 3: containment -> ANI ------------------------------------------------------------ */
/* identity[q,s] = max(ani(ov/|Q|), ani(ov/|S|)), cov_query[q,s] = ani(ov/|Q|),
 * ani(c) = 0 if c==0, 1 if c==1, else 1-(1-c^(1/k)); NaN where branchwater prints no row (ov==0).
 * pyani-plus mapping: private_cli.py:1875-1887.  Device form uses CUDA's double pow (<= 2 ulp);
 * the host form uses libm pow and is the one the drop-in CLI path uses (repr-exact floats). */
PANIB_API int panib_ani_device(const uint32_t *d_ov, int64_t ld_ov, const int32_t *d_q_counts, int64_t nq,
                     const int32_t *d_s_counts, int64_t ns, int k, double *d_identity,
                     double *d_cov_query, void *stream);
PANIB_API int panib_ani_host(const uint32_t *h_ov, int64_t ld_ov, const int32_t *h_q_counts, int64_t nq,
                   const int32_t *h_s_counts, int64_t ns, int k, double *h_identity, double *h_cov_query);

/* ---- synthetic genomes (BASELINE.json configs 2-5) ---------------------------- */
/* Writes genomes g0..g0+n_genomes-1 (each `length` bases, one record) into an ASCII base stream
 * with the layout above: genome i at tile i*tiles_per_genome, tiles_per_genome = length/TILE + 1. */
PANIB_API int panib_synth_ascii(uint64_t seed, int64_t g0, int64_t n_genomes, int64_t length, uint8_t *d_ascii,
                      void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PANIB200_H */
