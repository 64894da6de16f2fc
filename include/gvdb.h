/* gvdb.h — C ABI of the B200-native quantized-search engine for grape-vector-db.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or
 * torch types.  One gvdb_index is ONE SHARD on ONE GPU: a contiguous range of corpus
 * rows held in HBM as (a) 1-bit codes, (b) the original f32 rows, (c) per-row L2
 * norms and (d) a live bitmap.  The reference interfaces each entry point replaces
 * are cited as /root/reference file:line.  Ids are dense row numbers; the String ids
 * of `trait VectorIndex` (src/index.rs:35-62) stay on the host side of the FFI
 * (id_to_index / index_to_id maps, src/index.rs:333-334) — see INTEGRATION.md.
 *
 * Threading: search entry points are re-entrant (the reference calls
 * VectorIndex::search under a tokio READ guard from many worker threads,
 * src/lib.rs:469-477).  Mutating entry points (add/remove/clear/reserve) require
 * exclusivity, exactly what the reference's write guard gives (src/lib.rs:351-352).
 *
 * Errors: every call returns a gvdb_status; nothing aborts or throws across the
 * boundary.  gvdb_last_error() returns the calling thread's last message.
 * There is NO CPU fallback: without a usable CUDA device gvdb_create fails.
 */
#ifndef GVDB_H
#define GVDB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GVDB_ABI_VERSION 5u

#if defined(__GNUC__)
#define GVDB_API __attribute__((visibility("default")))
#else
#define GVDB_API
#endif

/* Status codes map 1:1 onto the VectorDbError variants this path emits
 * (src/types.rs:859-920). */
typedef enum gvdb_status {
    GVDB_OK = 0,
    GVDB_ERR_INDEX_NOT_BUILT = 1,          /* VectorDbError::IndexNotBuilt  (src/index.rs:213,621-623) */
    GVDB_ERR_DIMENSION_MISMATCH = 2,       /* ::DimensionMismatch{expected,actual} (src/index.rs:590-594) */
    GVDB_ERR_INVALID_VECTOR_DIMENSION = 3, /* ::InvalidVectorDimension (src/quantization.rs:131-133,335) */
    GVDB_ERR_QUANTIZATION = 4,             /* ::QuantizationError(String) (src/quantization.rs:158-162) */
    GVDB_ERR_INDEX = 5,                    /* ::IndexError(String): CUDA / allocation / internal failures */
    GVDB_ERR_INVALID_ARGUMENT = 6,         /* ::ConfigError(String) */
    GVDB_ERR_NOT_IMPLEMENTED = 7           /* ::NotImplemented(String) */
} gvdb_status;

#define GVDB_NO_ID UINT64_MAX /* unfilled id slot; its score is -inf (distance +inf) */

typedef struct gvdb_index gvdb_index; /* opaque: one shard on one GPU */

/* BinaryQuantizationConfig (src/quantization.rs:11-31) + placement.  Zero-initialise,
 * set struct_size = sizeof(gvdb_config), then fill. */
typedef struct gvdb_config {
    uint32_t struct_size;
    uint32_t dim;            /* vector dimension (> 0) */
    float threshold;         /* BinaryQuantizationConfig::threshold  (default 0.0) */
    float rescore_ratio;     /* BinaryQuantizationConfig::rescore_ratio (default 0.1) */
    int32_t device;          /* CUDA device ordinal */
    uint32_t flags;          /* GVDB_FLAG_* */
    uint64_t capacity_rows;  /* rows to reserve in HBM up front (0 = grow on demand) */
    uint64_t row_base;       /* global row number of this shard's local row 0 */
    /* GVDB_FLAG_ROW_WINDOW: keep the f32 originals only for local rows
     * [window_first, window_first + window_count); every row still gets its 1-bit code and norm.
     * This is the "codes replicated, rows sharded" multi-GPU layout (codes are 32x smaller than
     * the rows): each GPU scans the whole corpus for ITS share of the queries and rescoring is
     * done by the GPU that owns the row (gvdb_stage1_device / gvdb_rescore_keys_device /
     * gvdb_finish_owned_device).  Entry points that need every row's f32 data fail with
     * GVDB_ERR_INVALID_ARGUMENT on a windowed index whose window does not cover all rows. */
    uint64_t window_first;
    uint64_t window_count;
} gvdb_config;
#define GVDB_FLAG_ROW_WINDOW 1u

/* IndexStats (src/index.rs:82-88) plus the HBM footprint. */
typedef struct gvdb_stats {
    uint64_t vector_count;   /* live rows == VectorIndex::len() */
    uint64_t rows;           /* rows ever added (live + tombstoned) */
    uint64_t dimension;
    uint64_t memory_usage;   /* rows * dim * 4, the reference's figure (src/index.rs:676-679) */
    uint64_t hbm_bytes;      /* bytes actually reserved on the device */
    uint64_t code_bytes_per_row;
} gvdb_stats;

/* ---- lifecycle ------------------------------------------------------------------ */
GVDB_API uint32_t gvdb_abi_version(void);
GVDB_API const char* gvdb_last_error(void);
/* replaces FaissVectorIndex::new / HnswVectorIndex::new at the wiring point src/lib.rs:256-261 */
GVDB_API gvdb_status gvdb_create(const gvdb_config* cfg, gvdb_index** out);
GVDB_API void gvdb_destroy(gvdb_index* h);

/* ---- ingest (exclusive) ----------------------------------------------------------- */
/* VectorIndex::add_vectors (src/index.rs:612-617; caller src/lib.rs:350-353).  `rows` is
 * n x dim row-major f32 in HOST memory; rows are copied to HBM, quantised
 * (BinaryQuantizer::quantize_batch, src/quantization.rs:125-127) and normed on the GPU.
 * *first_row_out = local row number of rows[0]. */
GVDB_API gvdb_status gvdb_add(gvdb_index* h, const float* rows, uint64_t n, uint64_t* first_row_out);
/* Same with `rows` already in DEVICE memory of h's GPU; work is enqueued on `stream`
 * (a cudaStream_t, NULL = default stream) and completed before return. */
GVDB_API gvdb_status gvdb_add_device(gvdb_index* h, void* stream, const float* rows_dev, uint64_t n,
                            uint64_t* first_row_out);
GVDB_API gvdb_status gvdb_reserve(gvdb_index* h, uint64_t capacity_rows);
/* VectorIndex::remove_vector (src/index.rs:642-650): tombstone; *was_live_out = Ok(bool). */
GVDB_API gvdb_status gvdb_remove(gvdb_index* h, uint64_t local_row, int32_t* was_live_out);
/* VectorIndex::clear (src/index.rs:664-670) */
GVDB_API gvdb_status gvdb_clear(gvdb_index* h);
/* VectorIndex::len / get_stats (src/index.rs:652-658,672-681) */
GVDB_API uint64_t gvdb_len(const gvdb_index* h);
GVDB_API gvdb_status gvdb_get_stats(const gvdb_index* h, gvdb_stats* out);

/* ---- persistence ------------------------------------------------ */
/* One flat little-endian file per shard, readable with mmap:
 *   [  0, 64)  header: "GVDBIDX1", u32 version = 1, u32 dim, f32 threshold, f32 rescore_ratio,
 *              u64 rows, u64 live rows, u64 row_base, u32 code bytes per row = ceil(dim/8),
 *              u32 flags (bit 0: the checksum is present), u64 checksum
 *   codes      rows x ceil(dim/8) bytes, BinaryVector::to_bytes() layout (src/quantization.rs:54-56),
 *              padded to a multiple of 64 bytes
 *   norms      rows x f32 (sequential-fold L2 norms), padded to 64
 *   live       ceil(rows/32) x u32 bitmap (bit r of word t = row 32t + r; 0 = tombstone), padded to 64
 *   rows       rows x dim f32, row-major
 * checksum = 64-bit FNV-1a taken over 8-byte little-endian WORDS instead of bytes: h = 0xcbf29ce484222325;
 * for every u64 word w of the four sections' bytes (file order, without their padding, as one byte
 * stream, the last partial word zero-extended) h = (h ^ w) * 0x100000001b3; finally h = (h ^ stream
 * length in bytes) * 0x100000001b3.  gvdb_load refuses a file whose checksum does not match; a file
 * with flags = 0 (written before the checksum existed) is loaded unchecked.  gvdb_save writes to
 * `path`.tmp, fsyncs and renames, so a crash mid-save leaves the previous file intact.
 * The reference reserves a `quantized` sled tree but never writes it and re-inserts vectors one
 * at a time on restart (src/advanced_storage.rs:52-62,105-112; src/query.rs:282-409); loading
 * this file restores the shard without re-quantising. */
GVDB_API gvdb_status gvdb_save(gvdb_index* h, const char* path);
/* Creates a new index on `device` from a file written by gvdb_save. */
GVDB_API gvdb_status gvdb_load(const char* path, int32_t device, gvdb_index** out);

/* ---- quantizer pieces (parity / BinaryVector interop) ----------------------------- */
/* BinaryQuantizer::quantize_batch (src/quantization.rs:86-127) on the GPU: n x dim f32
 * (host) -> n x ceil(dim/8) bytes (host) in BinaryVector::to_bytes() layout
 * (bit j -> byte j/8, bit 7-(j%8); src/quantization.rs:37,54-56). */
GVDB_API gvdb_status gvdb_quantize(gvdb_index* h, const float* x, uint64_t n, uint8_t* codes_out);
/* Stored codes of local rows [first, first+n) in the same byte layout. */
GVDB_API gvdb_status gvdb_get_codes(gvdb_index* h, uint64_t first, uint64_t n, uint8_t* codes_out);
/* BinaryQuantizer::hamming_distance (src/quantization.rs:130-141) of each query code
 * against every stored row: q_codes nq x ceil(dim/8) bytes (host) -> dist_out nq x rows
 * u32 (host).  Tombstoned rows are included (this is the raw scan). */
GVDB_API gvdb_status gvdb_hamming(gvdb_index* h, const uint8_t* q_codes, uint32_t nq, uint32_t* dist_out);
/* rescore_count of src/quantization.rs:178-179: min(n, (usize)((f32)n * ratio)). */
GVDB_API uint64_t gvdb_rescore_count(uint64_t n, float ratio);

/* ---- two-stage search (re-entrant) -------------------------------------------------- */
/* BinaryQuantizer::multi_stage_search (src/quantization.rs:151-193) for a batch, wired as
 * docs/architecture.md:355-372 describes: quantise the query, 1-bit Hamming scan over all
 * live rows, keep the top `rescore_count` by (hamming asc, row asc) [= the stable sort of
 * :175], exact f32 cosine of those [:181-187, :206-216], order by (cosine desc, hamming
 * asc, row asc) [= the stable sort of :190] and return the first k (k <= rescore_count;
 * the reference returns all rescore_count, i.e. k == rescore_count).
 *   queries        nq x dim f32
 *   ids_out        nq x k   u64 GLOBAL row numbers (row_base + local), GVDB_NO_ID if unfilled
 *   scores_out     nq x k   f32 cosine similarity, -inf if unfilled
 *   cand_ids_out   optional nq x rescore_count u64: the stage-1 candidate list in order
 *   cand_ham_out   optional nq x rescore_count u32: their Hamming distances (UINT32_MAX unfilled)
 * Host-pointer form: copies inside, results ready on return. */
GVDB_API gvdb_status gvdb_search_batch(gvdb_index* h, const float* queries, uint32_t nq, uint32_t k,
                              uint32_t rescore_count, uint64_t* ids_out, float* scores_out,
                              uint64_t* cand_ids_out, uint32_t* cand_ham_out);
/* Device-pointer form: every pointer is DEVICE memory on h's GPU; kernels are enqueued on
 * `stream`; the call returns after the stream has been checked for candidate-buffer
 * overflow (one 4-byte read-back), so results are complete on return.
 * Both forms always answer exactly: a batch of >= 64 queries runs ONE tensor-core pass under per-query
 * thresholds estimated from a strided sample of the rows; the device checks that every query got its
 * rescore_count candidates and the call is repeated with an exact segment schedule otherwise
 * (gvdb_profile.optimistic_reruns), and a corpus on which even that overflows a candidate buffer
 * (thousands of rows tying below a query's threshold, a tombstoned or filtered-out prefix) is answered by
 * the cut by counting, which has no capacity limit (gvdb_profile.overflow_fallbacks). */
GVDB_API gvdb_status gvdb_search_batch_device(gvdb_index* h, void* stream, const float* queries_dev,
                                     uint32_t nq, uint32_t k, uint32_t rescore_count,
                                     uint64_t* ids_out_dev, float* scores_out_dev,
                                     uint64_t* cand_ids_out_dev, uint32_t* cand_ham_out_dev);

/* ---- exact flat search (re-entrant) ------------------------------------------------- */
/* FaissVectorIndex::search (src/index.rs:620-640) + cosine_distance (:686-700) for a batch:
 * distance = 1 - cos (+inf on a zero norm) over all live rows, ascending, ties by row.
 *   ids_out nq x k (GVDB_NO_ID unfilled), dist_out nq x k (+inf unfilled).
 * Always answers, like the reference: rows are scanned in growing segments under the k-th best distance
 * so far; when that schedule overflows a candidate buffer (the rows seen first say nothing about the rest:
 * a tombstoned or filtered-out prefix, a corpus stored in order of relevance) the call is repeated with
 * fixed segments that cannot overflow (slower; gvdb_profile.overflow_fallbacks counts it). */
GVDB_API gvdb_status gvdb_flat_search_batch(gvdb_index* h, const float* queries, uint32_t nq, uint32_t k,
                                   uint64_t* ids_out, float* dist_out);
GVDB_API gvdb_status gvdb_flat_search_batch_device(gvdb_index* h, void* stream, const float* queries_dev,
                                          uint32_t nq, uint32_t k, uint64_t* ids_out_dev,
                                          float* dist_out_dev);

/* BasicVectorStore::vector_search (src/storage.rs:296-339) + its cosine_similarity (:851-865) for a
 * batch: similarity = cos (0.0 on a zero norm) over all live rows, rows with similarity < threshold
 * skipped when use_threshold != 0, descending, ties by row (the reference: sled key order), first k.
 *   ids_out nq x k (GVDB_NO_ID unfilled), sims_out nq x k (-inf unfilled). */
GVDB_API gvdb_status gvdb_similarity_search_batch(gvdb_index* h, const float* queries, uint32_t nq, uint32_t k,
                                                  float threshold, int32_t use_threshold,
                                                  uint64_t* ids_out, float* sims_out);
GVDB_API gvdb_status gvdb_similarity_search_batch_device(gvdb_index* h, void* stream, const float* queries_dev,
                                                         uint32_t nq, uint32_t k, float threshold,
                                                         int32_t use_threshold, uint64_t* ids_out_dev,
                                                         float* sims_out_dev);

/* ---- row-sharded search: the two halves around the exchange step -------------------- */
/* Failure mode of the sharded / staged entry points (gvdb_search_shard[_sliced]_device, gvdb_stage1_device,
 * gvdb_search_exchange_device): rescore_count <= 2048 (larger counts: the ratio-mode calls below), and they have no
 * cut-by-counting fallback — a shard on
 * which the exact segment schedule overflows a candidate buffer (see above) returns GVDB_ERR_INDEX for that
 * call.  In the peer exchange such a rank publishes an empty candidate list for the step, keeps serving
 * its peers and reports the error for its own batch only; the exchange stays in step. */
/* A shard's answer for nq queries is ONE packed record buffer (so the exchange is one
 * all-gather):  [ ids u64 nq x R | ham u32 nq x R | score f32 nq x R ],  16 * nq * R bytes.
 * Per query the R records are in (hamming asc, global row asc) order; unfilled slots are
 * (GVDB_NO_ID, UINT32_MAX, -inf). */
GVDB_API uint64_t gvdb_shard_record_bytes(uint32_t nq, uint32_t rescore_count);
/* Replaces the scatter side of ShardManager::search_vectors (src/distributed/shard.rs:760-775):
 * local stage 1 + stage 2 of this shard -> records_dev (DEVICE, gvdb_shard_record_bytes). */
GVDB_API gvdb_status gvdb_search_shard_device(gvdb_index* h, void* stream, const float* queries_dev,
                                              uint32_t nq, uint32_t rescore_count, void* records_dev);
/* Same, with the answer laid out for a query-sliced exchange: queries are cut into n_slices equal
 * consecutive slices (nq % n_slices == 0) and records_dev receives n_slices packed buffers back
 * to back, buffer s = gvdb_shard_record_bytes(nq / n_slices, R) bytes for slice s.  With one
 * slice per rank an all-to-all hands every rank all shards' records for ITS slice of the
 * queries, so the merge (and the rest of the per-query work) is divided across ranks too. */
GVDB_API gvdb_status gvdb_search_shard_sliced_device(gvdb_index* h, void* stream, const float* queries_dev,
                                                     uint32_t nq, uint32_t rescore_count, uint32_t n_slices,
                                                     void* records_dev);
/* gvdb_search_shard_sliced_device without its host synchronisation (dim % 4 == 0, rescore_count <= 256): the
 * single pass is enqueued on `stream` and the call returns, so the exchange and the merge can be queued behind it
 * without the GPU idling on the host.  gvdb_search_shard_verify, called once the step's other work is enqueued,
 * waits for the stream; *rerun_out = 1 means the device refused the pass's thresholds (rare) and the step must be
 * repeated with gvdb_search_shard_sliced_device.  gvdb_search_shard_verify reads ONE pinned pair of words per index:
 * a caller that goes through it keeps one enqueue in flight per index.  Several enqueues on the same stream may be
 * in flight when each gets its own verdict_out_dev and the caller reads those itself (the multi-tile batches of
 * grape-vector-db_b200/dist.py: every query tile's step is enqueued before one host wait).
 * verdict_out_dev (optional, DEVICE, 2 x u32): the same verdict written on the stream (any word nonzero = repeat),
 * so that the ranks of a sharded deployment can pass their verdicts round with the answers (one collective, one
 * host wait per step) instead of agreeing through the host. */
GVDB_API gvdb_status gvdb_search_shard_sliced_enqueue_device(gvdb_index* h, void* stream, const float* queries_dev,
                                                             uint32_t nq, uint32_t rescore_count, uint32_t n_slices,
                                                             void* records_dev, uint32_t* verdict_out_dev);
GVDB_API gvdb_status gvdb_search_shard_verify(gvdb_index* h, void* stream, int32_t* rerun_out);
/* Replaces the gather side (concat + sort + truncate, src/distributed/shard.rs:776-783) with the
 * rule that reproduces the single-index result: over the n_shards x R gathered records of
 * each query keep the global top R by (hamming, global row), then order by (cosine desc,
 * hamming asc, row asc) and emit k.  `records_dev` is n_shards packed buffers back to back,
 * in rank order — exactly what an all-gather of the gvdb_search_shard_device outputs yields. */
/* gvdb_merge_shards_device, gvdb_rescore_keys_device and gvdb_finish_owned_device are
 * stream-ordered: they enqueue on `stream` and return; outputs are complete once the stream
 * reaches that point (they have nothing to read back on the host). */
GVDB_API gvdb_status gvdb_merge_shards_device(gvdb_index* h, void* stream, uint32_t n_shards,
                                              const void* records_dev, uint32_t nq,
                                              uint32_t rescore_count, uint32_t k,
                                              uint64_t* ids_out_dev, float* scores_out_dev);

/* Ratio mode across row shards (rescore_count > 2048, the reference's default rescore_ratio = 0.1 on a sharded corpus;
 * SURVEY.md §8e).  Three calls around two gathers; the answer equals the single-index search's bit for bit:
 *   1. gvdb_shard_hist_device: per query, the histogram of this shard's live rows over the Hamming distance
 *      (gvdb_shard_hist_bins(h) u32 bins per query: distances 0 .. code bits).
 *   2. gather every shard's histograms in rank order: hists_all = [n_shards][nq][bins].
 *   3. gvdb_search_shard_ratio_device: every shard derives the same global cut from hists_all — the bin holding the
 *      rescore_count-th smallest key (hamming, global row) of the union; ties inside that bin go to the shards in rank
 *      order, i.e. by ascending global row — rescoring its own members of the global top rescore_count exactly, and
 *      writes its best k of them by (cosine desc, hamming asc, row asc) as a packed record buffer
 *      (gvdb_shard_record_bytes(nq, k); per query in (hamming, row) order, unfilled slots as above).  k <= 1024.
 *   4. gather the record buffers in rank order; gvdb_merge_shards_ratio_device orders the n_shards x k records of each
 *      query by (cosine desc, hamming asc, row asc) and emits k (n_shards * k <= 4096).
 * One query at a time inside (like the single-index cut by counting); results complete on return of call 3. */
GVDB_API uint32_t gvdb_shard_hist_bins(const gvdb_index* h);
GVDB_API gvdb_status gvdb_shard_hist_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                                            uint32_t* hist_out_dev);
GVDB_API gvdb_status gvdb_search_shard_ratio_device(gvdb_index* h, void* stream, const float* queries_dev,
                                                    uint32_t nq, uint64_t rescore_count, uint32_t k,
                                                    const uint32_t* hists_all_dev, uint32_t n_shards,
                                                    uint32_t my_shard, void* records_dev);
GVDB_API gvdb_status gvdb_merge_shards_ratio_device(gvdb_index* h, void* stream, uint32_t n_shards,
                                                    const void* records_dev, uint32_t nq, uint32_t k,
                                                    uint64_t* ids_out_dev, float* scores_out_dev);

/* ---- query-parallel search over replicated codes + row-sharded originals ----------------- */
/* Keys are hamming << 40 | global row (rows < 2^40), GVDB_NO_ID when unfilled.
 * Stage 1 only (quantise, Hamming scan, exact top rescore_count by (hamming, row)):
 *   keys_out_dev  nq x rescore_count u64, ascending per query.   rescore_count <= 2048. */
GVDB_API gvdb_status gvdb_stage1_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                                        uint32_t rescore_count, uint64_t* keys_out_dev);
/* Owner-computes rescoring: for every key whose row lies in THIS index's resident window, the
 * exact f32 cosine of (queries_dev[q], row) (src/quantization.rs:206-216); 0.0f elsewhere.
 *   queries_dev nq x dim, keys_dev / scores_out_dev nq x rescore_count. */
GVDB_API gvdb_status gvdb_rescore_keys_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                                              uint32_t rescore_count, const uint64_t* keys_dev,
                                              float* scores_out_dev);
/* Final order for nq queries: scores_by_owner_dev holds n_owners score arrays (nq x rescore_count
 * each, back to back, owner g = rows [g * rows_per_owner, (g+1) * rows_per_owner)); each key takes
 * the score of the owner of its row, then (cosine desc, hamming asc, row asc), first k. */
GVDB_API gvdb_status gvdb_finish_owned_device(gvdb_index* h, void* stream, const uint64_t* keys_dev,
                                              const float* scores_by_owner_dev, uint32_t n_owners,
                                              uint64_t rows_per_owner, uint32_t nq, uint32_t rescore_count,
                                              uint32_t k, uint64_t* ids_out_dev, float* scores_out_dev);

/* ---- peer rows: rescoring straight out of the other GPUs' HBM over NVLink ------------------ */
/* With GVDB_FLAG_ROW_WINDOW every GPU holds all codes and one contiguous share of the f32 rows:
 * owner o keeps rows [o * rows_per_owner, (o+1) * rows_per_owner).  Once the owners' row buffers
 * are attached, gvdb_search_batch(_device) works on the windowed index: the rescoring kernel
 * copies each candidate row from its owner's memory (cp.async on the peer-mapped address, i.e.
 * NVLink/NVSwitch loads inside the kernel) — no collective in the data path.
 * Across processes: exchange gvdb_export_rows_ipc handles (any transport; this repository
 * all-gathers them with torch.distributed) and call gvdb_attach_peer_rows_ipc.  Inside one
 * process (several indexes, tests): gvdb_rows_device_ptr + gvdb_attach_peer_rows_ptr. */
#define GVDB_IPC_HANDLE_BYTES 64
GVDB_API gvdb_status gvdb_export_rows_ipc(gvdb_index* h, uint8_t* handle_out /* 64 bytes */);
GVDB_API gvdb_status gvdb_attach_peer_rows_ipc(gvdb_index* h, uint32_t n_owners, uint64_t rows_per_owner,
                                               uint32_t my_owner, const uint8_t* handles /* n_owners x 64 */);
GVDB_API const void* gvdb_rows_device_ptr(const gvdb_index* h);
GVDB_API gvdb_status gvdb_attach_peer_rows_ptr(gvdb_index* h, uint32_t n_owners, uint64_t rows_per_owner,
                                               uint32_t my_owner, const void* const* row_ptrs /* n_owners */);

/* ---- filtered search --------------------------------------------------
 * The row bitmap the scans consult — today the tombstones of remove_vector (src/index.rs:629,642-650)
 * — ANDed with a per-call allow-list, e.g. the ids FilterEngine::execute_filter returns
 * (src/filtering.rs:374) mapped to row numbers.  allow_bits: ceil(rows / 32) words, bit (r % 32) of
 * word r / 32 set when local row r may be returned.  The answer is the search over the sub-corpus of
 * allowed live rows (same order, same scores); rescore_count counts allowed rows. */
GVDB_API gvdb_status gvdb_search_batch_filtered(gvdb_index* h, const float* queries, const uint32_t* allow_bits,
                                                uint32_t nq, uint32_t k, uint32_t rescore_count,
                                                uint64_t* ids_out, float* scores_out);
GVDB_API gvdb_status gvdb_search_batch_filtered_device(gvdb_index* h, void* stream, const float* queries_dev,
                                                       const uint32_t* allow_bits_dev, uint32_t nq, uint32_t k,
                                                       uint32_t rescore_count, uint64_t* ids_out_dev,
                                                       float* scores_out_dev);
GVDB_API gvdb_status gvdb_flat_search_batch_filtered(gvdb_index* h, const float* queries, const uint32_t* allow_bits,
                                                     uint32_t nq, uint32_t k, uint64_t* ids_out, float* dist_out);

/* ---- peer exchange: the same layout with NO collective library in the data path ----------------
 * Codes replicated, f32 rows sharded, queries partitioned (as gvdb_stage1_device /
 * gvdb_rescore_keys_device / gvdb_finish_owned_device), but the three exchanges of a step (queries
 * and candidate keys to the owners, cosines back) are posted stores into the peers' HBM over
 * NVLink, ordered by release/acquire flags; one call per step, nothing on the host in between.
 * Replaces the scatter/gather the reference sketches in src/distributed/shard.rs:760-786.
 *   create   on every rank, after the rows are loaded: allocates this rank's mailbox for batches of
 *            at most nq_max queries and rescore_max candidates; the index must hold owner `rank`'s
 *            row window (GVDB_FLAG_ROW_WINDOW, window_first = rank * rows_per_owner)
 *   export / attach (IPC)   exchange the 64-byte handles (one all-gather, once), attach, then
 *            BARRIER before the first step
 *   mailbox_ptr / attach_ptr   the same wiring for several indexes inside one process
 *   search   every rank calls it once per step with the same nq and rescore_count, each with ITS
 *            batch; answers land in ids_out_dev / scores_out_dev (nq x k) on `stream`.  A peer
 *            that does not show up within GVDB_XCHG_TIMEOUT_MS (default 20000) makes the step's
 *            answers invalid and gvdb_exchange_status (which waits for the steps in flight) and
 *            every later step return IndexError — never a hang. */
GVDB_API gvdb_status gvdb_exchange_create(gvdb_index* h, uint32_t world, uint32_t rank, uint64_t rows_per_owner,
                                          uint32_t nq_max, uint32_t rescore_max);
GVDB_API gvdb_status gvdb_exchange_export_ipc(gvdb_index* h, uint8_t* handle_out /* GVDB_IPC_HANDLE_BYTES */);
GVDB_API gvdb_status gvdb_exchange_attach_ipc(gvdb_index* h, const uint8_t* handles /* world * GVDB_IPC_HANDLE_BYTES */);
GVDB_API void* gvdb_exchange_mailbox_ptr(const gvdb_index* h);
GVDB_API gvdb_status gvdb_exchange_attach_ptr(gvdb_index* h, void* const* mailboxes /* world */);
GVDB_API gvdb_status gvdb_search_exchange_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                                                 uint32_t k, uint32_t rescore_count, uint64_t* ids_out_dev,
                                                 float* scores_out_dev);
GVDB_API gvdb_status gvdb_exchange_status(gvdb_index* h, uint32_t* timed_out_kinds /* optional */);

/* ---- sparse side of the hybrid search: BM25 over CSR postings ---------- */
/* SparseIndex (src/sparse.rs:31-222) as an immutable snapshot: postings in CSR form by term id,
 * documents numbered 0..n_docs-1, ascending inside each term's list.
 *   post_off n_terms+1, post_doc / post_tf one entry per posting, doc_len n_docs (all HOST).
 * average_document_length follows the reference (:96-104): the sum of document_length over ALL
 * postings entries (a document counts once per distinct term), in CSR order, divided by n_docs. */
typedef struct gvdb_sparse gvdb_sparse;
GVDB_API gvdb_status gvdb_sparse_create(int32_t device, float k1, float b, gvdb_sparse** out);
GVDB_API void gvdb_sparse_destroy(gvdb_sparse* s);
GVDB_API gvdb_status gvdb_sparse_build(gvdb_sparse* s, uint64_t n_docs, uint32_t n_terms, const uint64_t* post_off,
                                       const uint32_t* post_doc, const float* post_tf, const float* doc_len);
GVDB_API float gvdb_sparse_average_document_length(const gvdb_sparse* s);
/* SparseIndex::search_bm25 (:153-199) for a batch of sparse query vectors (CSR: q_off nq+1,
 * q_terms / q_tfs), all HOST.  doc_out nq x limit (GVDB_NO_ID unfilled), score_out nq x limit
 * (-inf unfilled); order: score descending, ties by document number. */
GVDB_API gvdb_status gvdb_sparse_search_bm25_batch(gvdb_sparse* s, uint32_t nq, const uint64_t* q_off,
                                                   const uint32_t* q_terms, const float* q_tfs, uint32_t limit,
                                                   uint64_t* doc_out, float* score_out);
/* The same with the answers left in DEVICE buffers, asynchronous on `stream` (the query CSR is still
 * HOST memory: it is a few bytes per query and the idf is computed on the host). */
GVDB_API gvdb_status gvdb_sparse_search_bm25_batch_device(gvdb_sparse* s, void* stream, uint32_t nq,
                                                          const uint64_t* q_off, const uint32_t* q_terms,
                                                          const float* q_tfs, uint32_t limit,
                                                          uint64_t* doc_out_dev, float* score_out_dev);
GVDB_API uint64_t gvdb_sparse_launches(const gvdb_sparse* s);   /* kernels launched by this handle's searches */

/* rrf_fusion (src/hybrid.rs:422-488) for nq queries at once.  Each list is nq x n_* document numbers,
 * best first; GVDB_NO_ID ends a query's list early (the padding the search calls leave); n_* = 0
 * skips a list.  score(doc) = sum over the lists, in the order dense, sparse, text, of
 * 1.0 / (k + (rank + 1) as f32): the dense loop inserts, the others add (same f32 order as the
 * reference).  Output: nq x limit, score descending; exact score ties — unspecified in the
 * reference, which sorts a HashMap's entries — by first appearance (dense, then sparse, then text).
 * n_dense + n_sparse + n_text <= 4096. */
GVDB_API gvdb_status gvdb_rrf_fusion_batch(int32_t device, const uint64_t* dense, uint32_t n_dense,
                                           const uint64_t* sparse, uint32_t n_sparse, const uint64_t* text,
                                           uint32_t n_text, uint32_t nq, float k, uint32_t limit,
                                           uint64_t* ids_out, float* scores_out);
GVDB_API gvdb_status gvdb_rrf_fusion_batch_device(int32_t device, void* stream, const uint64_t* dense_dev,
                                                  uint32_t n_dense, const uint64_t* sparse_dev, uint32_t n_sparse,
                                                  const uint64_t* text_dev, uint32_t n_text, uint32_t nq, float k,
                                                  uint32_t limit, uint64_t* ids_out_dev, float* scores_out_dev);

/* linear_fusion (src/hybrid.rs:491-566) and normalized_fusion (:568-616; normalize != 0: every list's scores go through
 * normalize_scores first — (score - min) / (max - min) over the list, 1.0 when max - min is not > 0) for nq queries at
 * once.  Lists as for rrf_fusion, each with its scores (nq x n_* f32: the dense similarities, the BM25 scores, the
 * text scores).  A document's fused score is built in the reference's order: the dense loop inserts
 * similarity * dense_weight (a repeated document overwrites), the sparse and the text loop add theirs; output ordered
 * by fused score descending, exact ties by first appearance (the reference's order among ties is its HashMap's). */
GVDB_API gvdb_status gvdb_weighted_fusion_batch(int32_t device, const uint64_t* dense, const float* dense_scores,
                                                uint32_t n_dense, const uint64_t* sparse, const float* sparse_scores,
                                                uint32_t n_sparse, const uint64_t* text, const float* text_scores,
                                                uint32_t n_text, uint32_t nq, float dense_weight, float sparse_weight,
                                                float text_weight, int32_t normalize, uint32_t limit,
                                                uint64_t* ids_out, float* scores_out);
GVDB_API gvdb_status gvdb_weighted_fusion_batch_device(int32_t device, void* stream, const uint64_t* dense_dev,
                                                       const float* dense_scores_dev, uint32_t n_dense,
                                                       const uint64_t* sparse_dev, const float* sparse_scores_dev,
                                                       uint32_t n_sparse, const uint64_t* text_dev,
                                                       const float* text_scores_dev, uint32_t n_text, uint32_t nq,
                                                       float dense_weight, float sparse_weight, float text_weight,
                                                       int32_t normalize, uint32_t limit, uint64_t* ids_out_dev,
                                                       float* scores_out_dev);

/* ---- measurement hooks ---------------------------------------------------------------- */
/* When enabled, every kernel the library launches is bracketed by CUDA events on the stream
 * it is launched on; times are accumulated at the call's final synchronisation. */
typedef struct gvdb_profile {
    uint64_t launches;        /* all kernels launched by this index */
    uint64_t scan_launches;   /* scan_kernel launches */
    double scan_ms;           /* summed scan_kernel device time */
    double scan_bytes;        /* algorithmic code bytes streamed by those launches:
                                 rows_in_segment * code_bytes_per_row * query_groups */
    double scan_pairs;        /* (row, query) pairs evaluated */
    double select_ms, rescore_ms, topk_ms, prep_ms, flat_ms, merge_ms;
    uint64_t tc_launches;     /* tc_scan_kernel (tcgen05) launches */
    double tc_ms;             /* summed tc_scan_kernel device time */
    double tc_macs;           /* algorithmic multiply-accumulates of those launches:
                                 rows * padded queries * code bits (one per code bit and pair) */
    double tc_bytes;          /* code bytes those launches read: rows * code_bytes_per_row * query slices */
    double scatter_ms;        /* tc_scatter_kernel (survivor records -> candidate keys) */
    uint64_t optimistic_reruns; /* calls repeated because the device refuted the single-pass threshold guess */
    uint64_t overflow_fallbacks; /* calls answered by the cut by counting after a candidate buffer overflowed */
    double exchange_ms;       /* peer exchange: push + signal kernels (gvdb_search_exchange_device) */
    double exchange_wait_ms;  /* peer exchange: time spent waiting for the peers' flags (rank skew included) */
    double sample_ms;         /* tc_scan_kernel in sample mode + tc_tau_kernel (single-pass threshold estimate) */
    double dot_ms;            /* tc_dot_kernel (ratio mode: bf16 tcgen05 GEMM of queries x rows) */
    uint64_t dot_launches;
    double dot_macs;          /* rows x padded queries x padded dims of those launches */
    uint64_t ratio_fallback_queries; /* ratio-mode queries answered by the cut by counting (the filter could not vouch for them) */
} gvdb_profile;
/* Parity / diagnostics of the ratio-mode filter (gvdb_ratio.cuh): dot_out[q * rows + r] = the bf16 tensor-core dot
 * product of query q and stored row r (f32 accumulation; operands rounded to bf16, so |error| <= 2^-8 |q||r|).
 * dim % 64 == 0 (or 96), dim <= 768, no row window. */
GVDB_API gvdb_status gvdb_approx_dot(gvdb_index* h, const float* queries, uint32_t nq, float* dot_out);
GVDB_API gvdb_status gvdb_profile_enable(gvdb_index* h, int32_t on);
/* The dense FP4 tensor-core rate of `device`, measured: back-to-back tcgen05.mma kind::mxf4
 * (M128 x N128 x K64, A in TMEM) on every SM for a few milliseconds, nothing else running.
 * tmacs_per_s_out: multiply-accumulates per second / 1e12 (x2 = FLOP/s); clk_per_mma_out: SM clocks
 * per MMA on one SM (64 = the pipe's rate).  The denominator of the tensor-core scan's roofline:
 * MEASURED_PEAKS.json holds no FP4 figure. */
GVDB_API gvdb_status gvdb_measure_fp4_mma_rate(int32_t device, double* tmacs_per_s_out, double* clk_per_mma_out);
GVDB_API gvdb_status gvdb_profile_read(gvdb_index* h, gvdb_profile* out, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* GVDB_H */
