// gvdb_sparse.cuh — BM25 over CSR postings on the GPU: the sparse list that
// HybridSearchEngine fuses with the dense one (/root/reference/src/hybrid.rs:305-308).
// SparseIndex::search_bm25 (/root/reference/src/sparse.rs:153-222), per query:
//   for each query term i, in query order, for each posting (doc, tf, len) of that term:
//       score[doc] (inserted as 0.0 on first touch) += q_tf[i] * tf' * idf[i]
//       tf' = tf * (k1 + 1) / (tf + k1 * (1 - b + b * (len / avg_len)))
//   sort by score descending, truncate(limit)          (ties: the reference's HashMap order is
//                                                       unspecified; here doc id ascending)
// Exactness: idf = ln((N - df + 0.5) / (df + 0.5)) is computed on the HOST (libm logf, as Rust's
// f32::ln) and passed in; the rest is f32 + - * / in the reference's order (__f*_rn, no FMA).  A
// document appears at most once in a term's postings, so one launch per term rank touches every
// accumulator at most once: additions happen in query-term order, like the reference's loop.
//   bm25_accumulate_kernel   one term rank of every query of the chunk
//   bm25_hist_kernel         radix-select levels (11 + 11 + 10 bits) over the descending score image
//   bm25_cut_kernel          the exact 32-bit image of the limit-th best score
//   bm25_tiecount / tiecut   when more documents tie at that score than fit: the last document kept
//   bm25_compact_kernel      keys (descending image << 32 | doc) inside the cut: exactly
//                            min(limit, touched documents), however many documents tie
//   bm25_topk_kernel         block bitonic sort of those keys, first `limit`
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gvdb_kernels.cuh"
#include "gvdb_tc.cuh"

namespace gvdb {

constexpr uint32_t BM25_ABSENT = 0xFFFFFFFFu;     // accumulator bit pattern of "document not touched" (a NaN)

// per query: the cut found level by level (see bm25_hist_kernel)
struct Bm25Cut {
    uint32_t present;      // documents touched by the query
    uint32_t want_left;    // rank of the wanted entry inside the current prefix (1-based)
    uint32_t prefix;       // key bits fixed so far (image levels, then document levels)
    uint32_t thr32, above; // exact image of the want-th best score, documents strictly better than it
    uint32_t m;            // documents at or above thr32 (m - above tie exactly at the cut)
    uint32_t doc_thr;      // ties are kept up to this document number (ascending = the tie order)
    uint32_t appended;
};

// tf' of every posting, once per index (k1, b and the average length are fixed when the postings are built):
//   (tf * (k1 + 1)) / (tf + k1 * (1 - b + b * (len / avg_len)))       src/sparse.rs:180-186, same operation order
__global__ void __launch_bounds__(256)
bm25_weight_kernel(const uint32_t* __restrict__ post_doc, const float* __restrict__ post_tf,
                   const float* __restrict__ doc_len, uint64_t n_post, float k1, float b, float avg_len,
                   float* __restrict__ post_w) {
    const float k1p1 = __fadd_rn(k1, 1.0f), one_minus_b = __fsub_rn(1.0f, b);
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_post; p += (uint64_t)gridDim.x * blockDim.x) {
        const float tf = post_tf[p], len = __ldg(doc_len + post_doc[p]);
        const float denom = __fadd_rn(tf, __fmul_rn(k1, __fadd_rn(one_minus_b, __fmul_rn(b, __fdiv_rn(len, avg_len)))));
        post_w[p] = __fdiv_rn(__fmul_rn(tf, k1p1), denom);
    }
}

// grid.y = query of the chunk; the query's rank-th term (if it has one) against its postings.
// acc_stride = n_docs rounded up to 4 (the later passes read the accumulators as uint4).
// (The dense-accumulator path: queries of more than BMB_MAX_TERMS terms or limits above BMB_MAX_LIMIT; everything
// else runs bm25_block_kernel below.)
__global__ void __launch_bounds__(256)
bm25_accumulate_kernel(const uint64_t* __restrict__ post_off, const uint32_t* __restrict__ post_doc,
                       const float* __restrict__ post_w, uint32_t n_terms,
                       const uint64_t* __restrict__ q_off, const uint32_t* __restrict__ q_terms,
                       const float* __restrict__ q_tfs, const float* __restrict__ q_idf, uint32_t q0, int rank,
                       uint64_t acc_stride, uint32_t* __restrict__ acc) {
    const uint32_t q = q0 + blockIdx.y;
    const uint64_t t0 = q_off[q], t1 = q_off[q + 1];
    if (t0 + rank >= t1) return;
    const uint32_t term = q_terms[t0 + rank];
    if (term >= n_terms) return;
    const float qtf = q_tfs[t0 + rank], idf = q_idf[t0 + rank];
    const uint64_t p0 = post_off[term], p1 = post_off[term + 1];
    uint32_t* mine = acc + (size_t)blockIdx.y * acc_stride;
    constexpr int U = 4;                       // postings in flight per thread (independent loads first)
    for (uint64_t base = p0 + (uint64_t)blockIdx.x * blockDim.x * U; base < p1; base += (uint64_t)gridDim.x * blockDim.x * U) {
        uint32_t doc[U], old[U];
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t p = base + (uint64_t)u * blockDim.x + threadIdx.x;
            doc[u] = p < p1 ? __ldg(post_doc + p) : 0xFFFFFFFFu;
            w[u] = p < p1 ? __ldg(post_w + p) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (doc[u] != 0xFFFFFFFFu) old[u] = mine[doc[u]];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (doc[u] == 0xFFFFFFFFu) continue;
            const float sc = __fmul_rn(__fmul_rn(qtf, w[u]), idf);
            const float base_v = old[u] == BM25_ABSENT ? 0.0f : __uint_as_float(old[u]);
            mine[doc[u]] = __float_as_uint(__fadd_rn(base_v, sc));
        }
    }
}

// descending image of a present score; a NaN score cannot be produced by finite inputs
__device__ __forceinline__ uint32_t bm25_desc_image(uint32_t bits) { return ~f32_asc_key(__uint_as_float(bits)); }

// Exact cut without sorting the accumulator.  The descending score image is fixed 11 + 11 + 10 bits at a
// time (radix select, BM25_LEVEL_BINS = 2048 bins per level, three uint4 passes over the accumulators):
// thr32 = the image of the want-th best score, above = documents strictly better, m = documents at or
// above it.  When more documents tie at thr32 than fit (m > want), the lowest document numbers win:
// bm25_tiecount_kernel counts the ties of each CTA's CONTIGUOUS slice of the documents (one more pass),
// bm25_tiecut_kernel finds the slice holding the last tie kept and walks it for doc_thr.  Exactly
// want = min(limit, present) documents then pass "image < thr32 or (image == thr32 and doc <= doc_thr)",
// however many share the cut score (BM25 over tf = count/len data ties by the million).
// Histograms are per CTA in shared memory, lanes with equal bins are combined first (__match_any_sync):
// tying scores put millions of documents into one bin, which serialises global atomics.
constexpr int BM25_LEVEL_BINS = 2048;
__host__ __device__ __forceinline__ constexpr int bm25_level_shift(int level) { return level == 0 ? 21 : level == 1 ? 10 : 0; }
__host__ __device__ __forceinline__ constexpr int bm25_level_bits(int level) { return level == 2 ? 10 : 11; }

template <int LEVEL>
__global__ void __launch_bounds__(256)
bm25_hist_kernel(const uint32_t* __restrict__ acc, uint64_t acc_stride, uint32_t* __restrict__ hist,
                 Bm25Cut* __restrict__ cut) {
    __shared__ uint32_t sh[BM25_LEVEL_BINS];
    const uint4* mine = reinterpret_cast<const uint4*>(acc + (size_t)blockIdx.y * acc_stride);
    uint32_t* h = hist + (size_t)blockIdx.y * BM25_LEVEL_BINS;
    const uint32_t prefix = cut[blockIdx.y].prefix;
    if (LEVEL >= 1 && cut[blockIdx.y].present == 0) return;
    for (int i = threadIdx.x; i < BM25_LEVEL_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    constexpr int shift = bm25_level_shift(LEVEL), bits = bm25_level_bits(LEVEL);
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n4 = acc_stride / 4;
    uint32_t present = 0;
    // two uint4 (8 accumulators) in flight per thread; the padding past n_docs reads as "absent"
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x * 2; base < n4; base += (uint64_t)gridDim.x * blockDim.x * 2) {
        uint4 v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint64_t i = base + (uint64_t)u * blockDim.x + threadIdx.x;
            v[u] = i < n4 ? __ldg(mine + i) : make_uint4(BM25_ABSENT, BM25_ABSENT, BM25_ABSENT, BM25_ABSENT);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                bool ok = w[e] != BM25_ABSENT;
                const uint32_t img = bm25_desc_image(w[e]);
                if constexpr (LEVEL == 0) present += ok;
                else ok = ok && (img >> (shift + bits)) == prefix;
                const uint32_t bin = ok ? (img >> shift) & ((1u << bits) - 1u) : 0xFFFFFFFFu;
                if (__any_sync(0xffffffffu, ok)) {
                    const uint32_t peers = __match_any_sync(0xffffffffu, bin);
                    if (ok && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&sh[bin], (uint32_t)__popc(peers));
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BM25_LEVEL_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&h[i], sh[i]);
    if (LEVEL == 0) {
        for (int o = 16; o > 0; o >>= 1) present += __shfl_xor_sync(0xffffffffu, present, o);
        if (lane == 0 && present) atomicAdd(&cut[blockIdx.y].present, present);
    }
}

// one CTA per query: the bin holding the want-th entry of this level's histogram (level 0..2)
__global__ void __launch_bounds__(256)
bm25_cut_kernel(const uint32_t* __restrict__ hist, uint32_t limit, Bm25Cut* __restrict__ cut, int level) {
    __shared__ uint32_t part[256];
    const uint32_t* h = hist + (size_t)blockIdx.x * BM25_LEVEL_BINS;
    Bm25Cut* c = cut + blockIdx.x;
    const uint32_t want_all = min(limit, c->present);
    const uint32_t want = level == 0 ? want_all : c->want_left;
    constexpr int PER = BM25_LEVEL_BINS / 256;
    uint32_t sum = 0;
    for (int i = 0; i < PER; ++i) sum += h[threadIdx.x * PER + i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (want_all == 0) { c->want_left = 0; c->prefix = 0; c->thr32 = 0; c->above = 0; c->m = 0; c->doc_thr = 0; return; }
        uint32_t run = 0;
        int t = 0;
        for (; t < 255; ++t) { if (run + part[t] >= want) break; run += part[t]; }
        for (int i = 0; i < PER; ++i) {
            const uint32_t hv = h[t * PER + i];
            if (run + hv >= want) {
                const uint32_t bin = (uint32_t)(t * PER + i);
                const uint32_t prefix = level == 0 ? bin : (c->prefix << bm25_level_bits(level)) | bin;
                c->prefix = prefix;
                c->want_left = want - run;                     // rank inside the chosen bin
                if (level == 2) {                              // the image is complete
                    c->thr32 = prefix;
                    c->above = want_all - (want - run);
                    c->m = want_all - (want - run) + hv;
                    c->doc_thr = 0xFFFFFFFFu;                  // every tie is kept unless bm25_tiecut_kernel says otherwise
                }
                break;
            }
            run += hv;
        }
    }
}

// surplus ties only (m > want): ties inside CTA x's contiguous slice of query y's accumulators
__global__ void __launch_bounds__(256)
bm25_tiecount_kernel(const uint32_t* __restrict__ acc, uint64_t acc_stride, uint32_t limit, const Bm25Cut* __restrict__ cut,
                     uint32_t* __restrict__ counts) {
    const Bm25Cut c = cut[blockIdx.y];
    if (c.m == min(limit, c.present)) return;
    const uint4* mine = reinterpret_cast<const uint4*>(acc + (size_t)blockIdx.y * acc_stride);
    const uint64_t n4 = acc_stride / 4, per = (n4 + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = (uint64_t)blockIdx.x * per, hi = min(n4, lo + per);
    // the stored bit pattern whose image is thr32 is unique up to -0.0 / +0.0: compare images
    uint32_t n = 0;
    for (uint64_t base = lo; base < hi; base += (uint64_t)blockDim.x * 2) {
        uint4 v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint64_t i = base + (uint64_t)u * blockDim.x + threadIdx.x;
            v[u] = i < hi ? __ldg(mine + i) : make_uint4(BM25_ABSENT, BM25_ABSENT, BM25_ABSENT, BM25_ABSENT);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) n += (w[e] != BM25_ABSENT && bm25_desc_image(w[e]) == c.thr32) ? 1u : 0u;
        }
    }
    __shared__ uint32_t warp_n[8];
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) warp_n[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += warp_n[w];
        counts[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

// one CTA per query: the slice holding the want_left-th tie (document order), then that document
__global__ void __launch_bounds__(256)
bm25_tiecut_kernel(const uint32_t* __restrict__ acc, uint64_t acc_stride, uint32_t limit, Bm25Cut* __restrict__ cut,
                   const uint32_t* __restrict__ counts, uint32_t n_slices) {
    Bm25Cut* c = cut + blockIdx.x;
    const uint32_t thr = c->thr32;
    if (c->m == min(limit, c->present)) return;                  // doc_thr stays "all ties"
    __shared__ uint32_t s_slice, s_rank, s_warp[8], s_found;
    if (threadIdx.x == 0) {
        uint32_t run = 0, want = c->want_left, b = 0;
        for (; b + 1 < n_slices; ++b) {
            const uint32_t t = counts[(size_t)blockIdx.x * n_slices + b];
            if (run + t >= want) break;
            run += t;
        }
        s_slice = b; s_rank = want - run; s_found = 0xFFFFFFFFu;
    }
    __syncthreads();
    const uint4* mine = reinterpret_cast<const uint4*>(acc + (size_t)blockIdx.x * acc_stride);
    const uint64_t n4 = acc_stride / 4, per = (n4 + n_slices - 1) / n_slices;
    const uint64_t lo = (uint64_t)s_slice * per, hi = min(n4, lo + per);
    uint32_t need = s_rank;                                       // ties still to pass, block-uniform
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint64_t base = lo; base < hi; base += blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        const uint4 v = i < hi ? __ldg(mine + i) : make_uint4(BM25_ABSENT, BM25_ABSENT, BM25_ABSENT, BM25_ABSENT);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t tie[4], mine_n = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) { tie[e] = (w[e] != BM25_ABSENT && bm25_desc_image(w[e]) == thr) ? 1u : 0u; mine_n += tie[e]; }
        uint32_t incl = mine_n;                                   // inclusive scan over the CTA, thread order = document order
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w8 = 0; w8 < 8; ++w8) { if (w8 < (int)wid) before += s_warp[w8]; total += s_warp[w8]; }
        const uint32_t excl = before + incl - mine_n;
        if (mine_n && excl < need && excl + mine_n >= need) {      // the need-th tie is one of my four
            uint32_t seen = excl;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (tie[e] && ++seen == need) s_found = (uint32_t)(i * 4 + e);
        }
        __syncthreads();
        if (s_found != 0xFFFFFFFFu || total >= need) break;       // block-uniform
        need -= total;
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_found != 0xFFFFFFFFu) c->doc_thr = s_found;
}

// keys (image << 32 | doc) of the documents inside the cut: exactly min(limit, present) of them
__global__ void __launch_bounds__(256)
bm25_compact_kernel(const uint32_t* __restrict__ acc, uint64_t acc_stride, Bm25Cut* __restrict__ cut,
                    uint64_t* __restrict__ keys, uint32_t key_cap) {
    const uint4* mine = reinterpret_cast<const uint4*>(acc + (size_t)blockIdx.y * acc_stride);
    Bm25Cut* c = cut + blockIdx.y;
    if (c->m == 0) return;
    const uint32_t thr = c->thr32, doc_thr = c->doc_thr;
    uint64_t* out = keys + (size_t)blockIdx.y * key_cap;
    const uint64_t n4 = acc_stride / 4;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x * 2; base < n4; base += (uint64_t)gridDim.x * blockDim.x * 2) {
        uint4 v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint64_t i = base + (uint64_t)u * blockDim.x + threadIdx.x;
            v[u] = i < n4 ? __ldg(mine + i) : make_uint4(BM25_ABSENT, BM25_ABSENT, BM25_ABSENT, BM25_ABSENT);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint64_t i = base + (uint64_t)u * blockDim.x + threadIdx.x;
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (w[e] == BM25_ABSENT) continue;
                const uint32_t img = bm25_desc_image(w[e]);
                const uint32_t doc = (uint32_t)(i * 4 + e);
                if (img < thr || (img == thr && doc <= doc_thr)) {
                    const uint32_t pos = atomicAdd(&c->appended, 1u);
                    if (pos < key_cap) out[pos] = ((uint64_t)img << 32) | doc;
                }
            }
        }
    }
}

// one CTA per query: sort the compacted keys (m <= SORT_N), emit (doc, score) for the first `limit`;
// the exact score bits come back from the accumulator (the image folds -0.0 into +0.0)
__global__ void __launch_bounds__(SORT_THREADS)
bm25_topk_kernel(const uint64_t* __restrict__ keys, uint32_t key_cap, const Bm25Cut* __restrict__ cut,
                 const uint32_t* __restrict__ acc, uint64_t acc_stride, uint32_t limit, uint64_t* __restrict__ doc_out,
                 float* __restrict__ score_out) {
    extern __shared__ __align__(16) uint64_t skeys[];
    const uint32_t q = blockIdx.x;
    const uint32_t m = min(min(cut[q].present, limit), key_cap);
    const uint32_t n_eff = max(64u, next_pow2(m));
    for (uint32_t i = threadIdx.x; i < n_eff; i += blockDim.x) skeys[i] = i < m ? keys[(size_t)q * key_cap + i] : UINT64_MAX;
    __syncthreads();
    bitonic_sort_smem(skeys, n_eff);
    for (uint32_t t = threadIdx.x; t < limit; t += blockDim.x) {
        if (t < m) {
            const uint32_t doc = (uint32_t)skeys[t];
            doc_out[(size_t)q * limit + t] = doc;
            score_out[(size_t)q * limit + t] = __uint_as_float(acc[(size_t)q * acc_stride + doc]);
        } else {
            doc_out[(size_t)q * limit + t] = UINT64_MAX;
            score_out[(size_t)q * limit + t] = -INFINITY;
        }
    }
}

// ---- the blocked path (round 2): accumulators in shared memory, no per-query image of the corpus in HBM --------
// The dense path above costs ~450 MB of memory traffic per query on 5M documents (a 20 MB accumulator array that is
// cleared, read-modify-written through 32-byte sectors and read five more times by the cut).  Here the documents are
// cut into blocks of BMB_DOCS; a CTA owns (query, segment of consecutive blocks) and, block after block, keeps the
// block's accumulators in SHARED memory:
//   * at the start it finds, for every query term, where each of its blocks begins inside the term's postings
//     (documents ascend inside a term: one binary search per (term, block boundary), all in parallel);
//   * the segment's postings then form ONE stream of pieces (block, term, <= BMB_CH postings) that thread 0 keeps
//     BMB_NS pieces ahead of the consumers with TMA bulk copies (documents and weights, two copies per piece into a
//     shared-memory ring, completion on an mbarrier): the stream runs on across block boundaries, so the memory
//     latency is paid once per segment, not once per term and block.  The threads add a piece's postings into the
//     accumulators (documents are distinct inside a term: no conflicts); the __syncthreads after every piece frees
//     its stage AND keeps each accumulator's additions in query-term order, the reference's order;
//   * the block's candidates — present scores whose key (descending score image << 32 | document) is not above the
//     running bound — are counted with float compares (an absent accumulator is a NaN and fails them), four
//     documents per lane and step, and appended in document order (warp scans, no atomics) to a candidate buffer of
//     2 x LP keys; when it fills, a bitonic sort keeps the best `limit` and the limit-th key becomes the new bound
//     (shared with the query's other segments through one atomicMin).  A block with more candidates than the buffer
//     holds (the first block of a segment; corpora whose scores tie by the thousand) first finds its own limit-th key
//     exactly with a 4 x 8-bit radix select over the block and a walk of the ties in document order.
// Every segment ends with its best `limit` keys; bm25_merge_kernel orders the segments' keys and rebuilds the score
// bits from the image (the image is one-to-one here: sums that start from +0.0 never produce -0.0).
// Items are ordered segment-major, so the CTAs running together work on the same blocks for different queries and
// share the postings of common terms in L2.
constexpr int BMB_DOCS = 8192;
constexpr int BMB_WARPS = 4;                            // consumer warps
constexpr int BMB_CONSUMERS = 32 * BMB_WARPS;
constexpr int BMB_THREADS = BMB_CONSUMERS + 32;         // + the producer warp (one lane issues the TMA copies)
constexpr int BMB_PER_WARP = BMB_DOCS / BMB_WARPS;      // 2048 consecutive documents per warp
constexpr int BMB_STEPS = BMB_PER_WARP / 32;            // block select: 64 steps of 32 lanes
constexpr int BMB_CSTEPS = BMB_PER_WARP / 128;          // candidate count: 16 steps of 32 lanes x 4 documents
constexpr int BMB_CH = 512;                             // postings per staged piece
constexpr int BMB_NS = 4;                               // pieces in flight (shared-memory ring)
constexpr int BMB_MAX_TERMS = 64;
constexpr int BMB_MAX_LIMIT = 1024;
constexpr int BMB_MAX_BPS = 80;                         // blocks per segment (boundary table in shared memory)
constexpr uint32_t BMB_END = 0xFFFFFFFFu;               // piece descriptor: the stream is over

__host__ __device__ inline size_t bmb_smem_bytes(uint32_t LP, uint32_t t_cap, uint32_t bps) {
    size_t b = (size_t)BMB_DOCS * 4 + (size_t)BMB_NS * BMB_CH * 8 + (size_t)2 * LP * 8;   // acc, ring, cand
    b += (size_t)t_cap * 8 + 2 * BMB_NS * 8 + BMB_NS * 16 + 256 * 4;                     // s_p0, barriers, descriptors, hist
    b += (size_t)t_cap * (bps + 1) * 4;                                                  // tb
    b += (size_t)t_cap * 12;                                                             // s_len, s_qtf, s_idf
    return b + 64;
}

__device__ __forceinline__ bool bmb_pass(uint32_t bits, uint32_t doc, uint64_t bound, uint64_t& key) {
    key = ((uint64_t)bm25_desc_image(bits) << 32) | doc;
    return bits != BM25_ABSENT && key <= bound;
}
// mbarrier wait that lets the hardware park the thread (suspend-time hint in ns) instead of spinning on issue slots
__device__ __forceinline__ void bmb_wait_parked(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(100000u) : "memory");
}
// barrier of the consumer warps only (the producer warp never joins it)
__device__ __forceinline__ void bmb_sync() { asm volatile("bar.sync 1, %0;" :: "n"(BMB_CONSUMERS) : "memory"); }
__device__ __forceinline__ void bmb_sort(uint64_t* s, uint32_t n, uint32_t tid) {   // bitonic_sort_smem for the consumers
    for (uint32_t k = 2; k <= n; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = tid; t < (n >> 1); t += BMB_CONSUMERS) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                const bool up = (i & k) == 0;
                const uint64_t a = s[i], b = s[l];
                if ((a > b) == up) { s[i] = b; s[l] = a; }
            }
            bmb_sync();
        }
}

// the pieces of a segment in (block, term, offset) order
struct BmbIter { uint32_t j, i, rel; bool started; };
__device__ __forceinline__ bool bmb_next(BmbIter& it, uint32_t nb, uint32_t T, const uint32_t* tb, uint32_t tbs,
                                         const uint64_t* s_p0, uint64_t& ga, uint32_t& n_el, uint32_t& pj, uint32_t& pi) {
    while (it.j < nb) {
        const uint32_t lo = tb[it.i * tbs + it.j], hi = tb[it.i * tbs + it.j + 1];
        if (!it.started) { it.rel = lo; it.started = true; }
        if (it.rel < hi) {
            ga = s_p0[it.i] + it.rel;
            const uint64_t ge = min((uint64_t)(s_p0[it.i] + hi), (uint64_t)((ga & ~(uint64_t)3) + BMB_CH));   // copies start 16-byte aligned
            n_el = (uint32_t)(ge - ga);
            pj = it.j; pi = it.i;
            it.rel += n_el;
            return true;
        }
        it.started = false;
        if (++it.i >= T) { it.i = 0; ++it.j; }
    }
    return false;
}

__global__ void __launch_bounds__(BMB_THREADS, 4)
bm25_block_kernel(const uint64_t* __restrict__ post_off, const uint32_t* __restrict__ post_doc,
                  const float* __restrict__ post_w, uint32_t n_terms, const uint64_t* __restrict__ q_off,
                  const uint32_t* __restrict__ q_terms, const float* __restrict__ q_tfs, const float* __restrict__ q_idf,
                  uint32_t nq, uint32_t n_blocks, uint32_t n_seg, uint32_t bps, uint32_t t_cap, uint32_t limit,
                  uint32_t LP, unsigned long long* __restrict__ bound, uint64_t* __restrict__ seg_keys) {
    extern __shared__ __align__(128) uint8_t bmb_smem[];
    uint32_t* acc = reinterpret_cast<uint32_t*>(bmb_smem);
    uint32_t* sdoc = acc + BMB_DOCS;                              // [BMB_NS][BMB_CH]
    float* sw = reinterpret_cast<float*>(sdoc + BMB_NS * BMB_CH); // [BMB_NS][BMB_CH]
    uint64_t* cand = reinterpret_cast<uint64_t*>(sw + BMB_NS * BMB_CH);
    uint4* desc = reinterpret_cast<uint4*>(cand + 2 * LP);        // per stage: {postings, offset of the first, term, block}
    uint64_t* full = reinterpret_cast<uint64_t*>(desc + BMB_NS);  // BMB_NS "piece landed" barriers
    uint64_t* empty = full + BMB_NS;                              // BMB_NS "stage free" barriers (one arrival per consumer warp)
    uint64_t* s_p0 = empty + BMB_NS;
    uint32_t* hist = reinterpret_cast<uint32_t*>(s_p0 + t_cap);
    uint32_t* tb = hist + 256;
    uint32_t* s_len = tb + (size_t)t_cap * (bps + 1);
    float* s_qtf = reinterpret_cast<float*>(s_len + t_cap);
    float* s_idf = s_qtf + t_cap;
    __shared__ uint32_t s_warp[BMB_WARPS];
    __shared__ uint32_t s_bin, s_need, s_thr_idx;
    __shared__ unsigned long long s_gb;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t seg = blockIdx.x / nq, q = blockIdx.x % nq;
    const uint32_t b0 = seg * bps, b1 = min(n_blocks, b0 + bps);
    const uint32_t tbs = bps + 1;
    uint64_t* out = seg_keys + ((size_t)q * n_seg + seg) * limit;
    const uint64_t t0 = q_off[q];
    const uint32_t T = (uint32_t)min((uint64_t)t_cap, q_off[q + 1] - t0);
    const uint32_t nb = T ? b1 - b0 : 0u;

    for (uint32_t i = tid; i < T; i += BMB_THREADS) {
        const uint32_t term = q_terms[t0 + i];
        const uint64_t p0 = term < n_terms ? post_off[term] : 0, p1 = term < n_terms ? post_off[term + 1] : 0;
        s_p0[i] = p0; s_len[i] = (uint32_t)(p1 - p0);
        s_qtf[i] = q_tfs[t0 + i]; s_idf[i] = q_idf[t0 + i];
    }
    if (tid == 0) {
        for (int s_ = 0; s_ < BMB_NS; ++s_) { mbar_init(smem_u32(full + s_), 1); mbar_init(smem_u32(empty + s_), BMB_WARPS); }
        fence_mbar_init();
    }
    __syncthreads();
    // where block b0 + j begins inside term i's postings (j = nb: where the segment ends)
    for (uint32_t x = tid; x < T * (nb + 1); x += BMB_THREADS) {
        const uint32_t i = x / (nb + 1), j = x % (nb + 1);
        const uint64_t target = (uint64_t)(b0 + j) * BMB_DOCS;
        const uint32_t* docs = post_doc + s_p0[i];
        uint32_t lo = 0, hi = s_len[i];
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((uint64_t)__ldg(docs + mid) < target) lo = mid + 1; else hi = mid;
        }
        tb[i * tbs + j] = lo;
    }
    __syncthreads();

    if (warp == BMB_WARPS) {
        // ===================== producer: one lane keeps the ring full =====================
        if (lane == 0) {
            BmbIter it{0, 0, 0, false};
            uint64_t ga; uint32_t n_el, pj, pi, n = 0;
            for (;; ++n) {
                const bool have = bmb_next(it, nb, T, tb, tbs, s_p0, ga, n_el, pj, pi);
                const uint32_t s_ = n % BMB_NS;
                if (n >= BMB_NS) bmb_wait_parked(smem_u32(empty + s_), ((n / BMB_NS) - 1) & 1u);
                const uint32_t bar = smem_u32(full + s_);
                if (!have) {
                    desc[s_] = make_uint4(BMB_END, 0, 0, 0);
                    mbar_arrive(bar);
                    break;
                }
                const uint64_t g0 = ga & ~(uint64_t)3;
                const uint32_t bytes = (uint32_t)(((ga - g0) + n_el + 3) & ~(uint64_t)3) * 4u;
                desc[s_] = make_uint4(n_el, (uint32_t)(ga - g0), pi, pj);
                mbar_expect_tx(bar, 2 * bytes);
                tma_bulk_g2s(smem_u32(sdoc + s_ * BMB_CH), post_doc + g0, bytes, bar);
                tma_bulk_g2s(smem_u32(sw + s_ * BMB_CH), post_w + g0, bytes, bar);
            }
        }
        return;
    }

    // ===================== consumers =====================
    uint32_t n_cand = 0;
    uint64_t bound_key = UINT64_MAX;

    auto sort_truncate = [&]() {
        for (uint32_t i = n_cand + tid; i < 2 * LP; i += BMB_CONSUMERS) cand[i] = UINT64_MAX;
        bmb_sync();
        bmb_sort(cand, 2 * LP, tid);
        n_cand = min(n_cand, limit);
        if (n_cand == limit) {
            const uint64_t nbk = cand[limit - 1];
            if (nbk < bound_key) {
                bound_key = nbk;
                if (tid == 0) atomicMin(bound + q, (unsigned long long)nbk);
            }
        }
    };

    // ---- candidates of the finished block: warp w owns documents [w * 2048, (w + 1) * 2048) ----
    auto candidates = [&](uint32_t block_start) {
        if ((uint64_t)s_gb < bound_key) bound_key = s_gb;
        const uint32_t wbase = warp * BMB_PER_WARP;
        const uint4* a4 = reinterpret_cast<const uint4*>(acc + wbase) + lane;
        for (int round = 0;; ++round) {
            if (round) bmb_sync();                              // s_warp of the previous round was read
            // pass <=> score > thr, or score == thr and document <= bdoc (an absent accumulator is a NaN: fails both)
            float thr_f = -INFINITY;
            uint32_t bdoc = 0xFFFFFFFFu;
            if (bound_key != UINT64_MAX) {
                const uint32_t asc = ~(uint32_t)(bound_key >> 32);
                thr_f = __uint_as_float((asc & 0x80000000u) ? (asc & 0x7fffffffu) : ~asc);
                bdoc = (uint32_t)bound_key;
            }
            // cheap look first: the warp's best score (fmaxf drops NaNs) against the threshold
            float best = -INFINITY;
#pragma unroll
            for (int e = 0; e < BMB_CSTEPS; ++e) {
                const uint4 v = a4[e * 32];
                best = fmaxf(fmaxf(best, fmaxf(__uint_as_float(v.x), __uint_as_float(v.y))),
                             fmaxf(__uint_as_float(v.z), __uint_as_float(v.w)));
            }
            uint32_t wcount = 0;
            if (__any_sync(0xffffffffu, best >= thr_f)) {
                uint32_t cnt = 0;
#pragma unroll
                for (int e = 0; e < BMB_CSTEPS; ++e) {
                    const uint4 v = a4[e * 32];
                    const uint32_t doc0 = block_start + wbase + e * 128 + lane * 4;
                    const float f0 = __uint_as_float(v.x), f1 = __uint_as_float(v.y), f2 = __uint_as_float(v.z), f3 = __uint_as_float(v.w);
                    cnt += (f0 > thr_f || (f0 == thr_f && doc0 <= bdoc)) ? 1u : 0u;
                    cnt += (f1 > thr_f || (f1 == thr_f && doc0 + 1 <= bdoc)) ? 1u : 0u;
                    cnt += (f2 > thr_f || (f2 == thr_f && doc0 + 2 <= bdoc)) ? 1u : 0u;
                    cnt += (f3 > thr_f || (f3 == thr_f && doc0 + 3 <= bdoc)) ? 1u : 0u;
                }
                wcount = __reduce_add_sync(0xffffffffu, cnt);
            }
            if (lane == 0) s_warp[warp] = wcount;
            bmb_sync();
            uint32_t total_c = 0, woff = 0;
#pragma unroll
            for (int w8 = 0; w8 < BMB_WARPS; ++w8) { const uint32_t c = s_warp[w8]; total_c += c; if (w8 < (int)warp) woff += c; }
            if (total_c == 0) break;
            if (total_c <= 2 * LP - n_cand) {                   // append in document order
                if (wcount) {
                    uint32_t off = n_cand + woff;
                    for (int e = 0; e < BMB_CSTEPS; ++e) {
                        const uint4 v = a4[e * 32];
                        const uint32_t doc0 = block_start + wbase + e * 128 + lane * 4;
                        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
                        uint32_t pm = 0;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float f = __uint_as_float(w4[c]);
                            pm |= (f > thr_f || (f == thr_f && doc0 + c <= bdoc)) ? (1u << c) : 0u;
                        }
                        if (!__any_sync(0xffffffffu, pm != 0)) continue;
                        const uint32_t mine_n = __popc(pm);
                        uint32_t incl = mine_n;
                        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
                        uint32_t at = off + incl - mine_n;
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if ((pm >> c) & 1u) cand[at++] = ((uint64_t)bm25_desc_image(w4[c]) << 32) | (doc0 + c);
                        off += __shfl_sync(0xffffffffu, incl, 31);
                    }
                }
                n_cand += total_c;
                break;
            }
            if (n_cand > limit) {                               // make room: keep the best `limit`, tighten the bound
                bmb_sync();
                sort_truncate();
                continue;
            }
            // more candidates in this one block than the buffer holds: the block's own limit-th key, exactly
            uint32_t prefix = 0, need = limit;
            for (int level = 0; level < 4; ++level) {
                const int shift = 24 - 8 * level;
                bmb_sync();
                for (uint32_t i = tid; i < 256; i += BMB_CONSUMERS) hist[i] = 0;
                bmb_sync();
                for (int e = 0; e < BMB_STEPS; ++e) {
                    const uint32_t idx = wbase + e * 32 + lane;
                    uint64_t key;
                    bool ok = bmb_pass(acc[idx], block_start + idx, bound_key, key);
                    const uint32_t img = (uint32_t)(key >> 32);
                    if (level > 0) ok = ok && (img >> (shift + 8)) == prefix;
                    const uint32_t bin = ok ? (img >> shift) & 255u : 0xFFFFFFFFu;
                    if (__any_sync(0xffffffffu, ok)) {
                        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
                        if (ok && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
                    }
                }
                bmb_sync();
                if (warp == 0) {                                // the bin holding the need-th key of this level
                    uint32_t sum = 0;
                    for (int i = 0; i < 8; ++i) sum += hist[lane * 8 + i];
                    uint32_t incl = sum;
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
                    const uint32_t before = incl - sum;
                    if (before < need && need <= incl) {
                        uint32_t run = before;
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t hv = hist[lane * 8 + i];
                            if (run + hv >= need) { s_bin = lane * 8 + i; s_need = need - run; break; }
                            run += hv;
                        }
                    }
                }
                bmb_sync();
                prefix = (prefix << 8) | s_bin;
                need = s_need;
            }
            // ties at the image `prefix`: the first `need` of them in document order stay
            {
                uint32_t tmask[BMB_STEPS / 32] = {}, tcount = 0;
                for (int e = 0; e < BMB_STEPS; ++e) {
                    const uint32_t idx = wbase + e * 32 + lane;
                    uint64_t key;
                    const bool ok = bmb_pass(acc[idx], block_start + idx, bound_key, key) && (uint32_t)(key >> 32) == prefix;
                    const uint32_t m = __ballot_sync(0xffffffffu, ok);
#pragma unroll
                    for (int h2 = 0; h2 < BMB_STEPS / 32; ++h2)
                        if ((int)lane + 32 * h2 == e) tmask[h2] = m;
                    tcount += __popc(m);
                }
                bmb_sync();
                if (lane == 0) s_warp[warp] = tcount;
                bmb_sync();
                uint32_t before = 0;
                for (int w8 = 0; w8 < (int)warp; ++w8) before += s_warp[w8];
                if (before < need && need <= before + tcount) { // this warp holds the last tie kept
                    uint32_t r = need - before;
                    for (int e = 0; e < BMB_STEPS; ++e) {
                        uint32_t m = 0;
#pragma unroll
                        for (int h2 = 0; h2 < BMB_STEPS / 32; ++h2) {
                            const uint32_t mm = __shfl_sync(0xffffffffu, tmask[h2], e & 31);
                            if ((e >> 5) == h2) m = mm;
                        }
                        const uint32_t c = __popc(m);
                        if (r <= c) {
                            if (lane == 0) {
                                uint32_t mm = m;
                                for (uint32_t x = 1; x < r; ++x) mm &= mm - 1;
                                s_thr_idx = wbase + e * 32 + (uint32_t)(__ffs(mm) - 1);
                            }
                            break;
                        }
                        r -= c;
                    }
                }
                bmb_sync();
                const uint64_t nbk = ((uint64_t)prefix << 32) | (block_start + s_thr_idx);
                bound_key = nbk;                                // <= the old bound: it is one of the passing keys
                if (tid == 0) atomicMin(bound + q, (unsigned long long)nbk);
            }
        }
    };

    // ---- the posting stream ----
    uint32_t cur_j = 0xFFFFFFFFu, cur_i = 0;
    for (uint32_t n = 0;; ++n) {
        const uint32_t s_ = n % BMB_NS;
        mbar_wait(smem_u32(full + s_), (n / BMB_NS) & 1u);
        const uint4 d = desc[s_];                                // {postings, offset of the first, term, block}
        if (d.x == BMB_END || d.w != cur_j) {
            bmb_sync();                                          // every warp has added the previous block's last piece
            if (cur_j != 0xFFFFFFFFu) { candidates((b0 + cur_j) * BMB_DOCS); bmb_sync(); }
            if (d.x == BMB_END) break;
            cur_j = d.w; cur_i = d.z;
            uint4* a4 = reinterpret_cast<uint4*>(acc);
            const uint4 absent = make_uint4(BM25_ABSENT, BM25_ABSENT, BM25_ABSENT, BM25_ABSENT);
#pragma unroll
            for (int i = 0; i < BMB_DOCS / 4 / BMB_CONSUMERS; ++i) a4[i * BMB_CONSUMERS + tid] = absent;
            if (tid == 0) s_gb = *reinterpret_cast<volatile unsigned long long*>(bound + q);
            bmb_sync();
        } else if (d.z != cur_i) {
            bmb_sync();                                          // additions stay in query-term order
            cur_i = d.z;
        }
        {
            const uint32_t block_start = (b0 + cur_j) * BMB_DOCS;
            const uint32_t* sd = sdoc + s_ * BMB_CH + d.y;
            const float* swp = sw + s_ * BMB_CH + d.y;
            const float qtf = s_qtf[d.z], idf = s_idf[d.z];
            // documents are distinct inside a term: the four read-modify-writes of a thread are independent, so
            // all loads go first (the compiler cannot know and would chain them)
            constexpr int R = BMB_CH / BMB_CONSUMERS;
            const uint32_t rtid = (tid + 32u * n) % BMB_CONSUMERS;   // partial pieces start at a different warp each time
            uint32_t dd[R], old[R];
            float sc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t x = r * BMB_CONSUMERS + rtid;
                dd[r] = x < d.x ? sd[x] - block_start : 0xFFFFFFFFu;
                sc[r] = x < d.x ? swp[x] : 0.0f;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                old[r] = dd[r] != 0xFFFFFFFFu ? acc[dd[r]] : 0u;
                sc[r] = __fmul_rn(__fmul_rn(qtf, sc[r]), idf);
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (dd[r] != 0xFFFFFFFFu)
                    acc[dd[r]] = __float_as_uint(__fadd_rn(old[r] == BM25_ABSENT ? 0.0f : __uint_as_float(old[r]), sc[r]));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(empty + s_));        // this warp is done with the stage
    }
    sort_truncate();
    for (uint32_t i = tid; i < limit; i += BMB_CONSUMERS) out[i] = i < n_cand ? cand[i] : UINT64_MAX;
}

// one CTA per query: the segments' keys -> the best `limit`, in order; scores rebuilt from the image
__global__ void __launch_bounds__(256)
bm25_merge_kernel(const uint64_t* __restrict__ seg_keys, uint32_t n_seg, uint32_t limit, uint32_t LP,
                  uint64_t* __restrict__ doc_out, float* __restrict__ score_out) {
    extern __shared__ __align__(16) uint64_t mk[];               // 2 x LP
    const uint32_t q = blockIdx.x, total = n_seg * limit;
    const uint64_t* in = seg_keys + (size_t)q * total;
    for (uint32_t i = threadIdx.x; i < LP; i += blockDim.x) mk[i] = UINT64_MAX;
    for (uint32_t c0 = 0; c0 < total; c0 += LP) {
        for (uint32_t i = threadIdx.x; i < LP; i += blockDim.x) mk[LP + i] = c0 + i < total ? in[c0 + i] : UINT64_MAX;
        __syncthreads();
        bitonic_sort_smem(mk, 2 * LP);
    }
    for (uint32_t t = threadIdx.x; t < limit; t += blockDim.x) {
        const uint64_t key = mk[t];
        if (key != UINT64_MAX) {
            const uint32_t asc = ~(uint32_t)(key >> 32);         // f32_asc_key of the score
            const uint32_t bits = (asc & 0x80000000u) ? (asc & 0x7fffffffu) : ~asc;
            doc_out[(size_t)q * limit + t] = (uint32_t)key;
            score_out[(size_t)q * limit + t] = __uint_as_float(bits);
        } else {
            doc_out[(size_t)q * limit + t] = UINT64_MAX;
            score_out[(size_t)q * limit + t] = -INFINITY;
        }
    }
}

// limits above the block bitonic capacity: the query's keys were sorted by a device radix sort; emit (doc, score)
__global__ void __launch_bounds__(256)
bm25_emit_sorted_kernel(const uint64_t* __restrict__ sorted_keys, const Bm25Cut* __restrict__ cut,
                        const uint32_t* __restrict__ acc, uint32_t limit, uint64_t* __restrict__ doc_out,
                        float* __restrict__ score_out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= limit) return;
    const uint32_t m = min(cut->present, limit);
    if (t < m) {
        const uint32_t doc = (uint32_t)sorted_keys[t];
        doc_out[t] = doc;
        score_out[t] = __uint_as_float(acc[doc]);
    } else {
        doc_out[t] = UINT64_MAX;
        score_out[t] = -INFINITY;
    }
}

__global__ void fill_f32_kernel(float* __restrict__ p, size_t n, float v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- rrf_fusion / linear_fusion / normalized_fusion (/root/reference/src/hybrid.rs:422-616) on the GPU ----------
// One CTA per query.  The three lists (dense, sparse, text; document numbers, GVDB_NO_ID ends a list
// early) are laid end to end in shared memory.  Every entry first gets its CONTRIBUTION:
//   MODE 0 (rrf_fusion, :422-488)         1 / (k + rank + 1)
//   MODE 1 (linear_fusion, :491-566)      score * weight of its list
//   MODE 2 (normalized_fusion, :568-616)  normalize_scores first: (score - min) / (max - min) over the list's valid
//                                         entries (1.0 when max - min is not > 0), then * weight
// Entry p is a HEAD when no earlier entry carries its document; a head's score is built in the reference's order:
// the dense loop INSERTS its contribution (a repeated document overwrites), the sparse and text loops add theirs in
// list order.  Heads are ordered by score descending; exact ties (unspecified in the reference: HashMap iteration)
// by first appearance.  n_d + n_s + n_t <= RRF_MAX.
constexpr uint32_t RRF_MAX = 4096;

struct FusionScores {                     // MODE 1, 2: the lists' scores and weights (unused by MODE 0)
    const float* dense; const float* sparse; const float* text;
    float w_dense, w_sparse, w_text;
};

template <int MODE>
__global__ void __launch_bounds__(SORT_THREADS)
list_fusion_kernel(const uint64_t* __restrict__ dense, uint32_t n_d, const uint64_t* __restrict__ sparse, uint32_t n_s,
                   const uint64_t* __restrict__ text, uint32_t n_t, float k, FusionScores fs, uint32_t limit,
                   uint64_t* __restrict__ ids_out, float* __restrict__ scores_out) {
    extern __shared__ __align__(16) uint64_t rrf_smem[];
    const uint32_t n = n_d + n_s + n_t;
    const uint32_t n_eff = max(64u, next_pow2(n));
    uint64_t* ids = rrf_smem;                 // [n]
    uint64_t* keys = rrf_smem + n;            // [n_eff]
    float* sc = reinterpret_cast<float*>(keys + n_eff);   // [n] fused score of a head
    float* con = sc + n;                      // [n] contribution of an entry
    __shared__ uint32_t len[3];
    __shared__ uint32_t s_min[3], s_max[3];   // MODE 2: ascending images of the lists' min / max score
    const uint32_t q = blockIdx.x;
    if (threadIdx.x < 3) {
        len[threadIdx.x] = threadIdx.x == 0 ? n_d : threadIdx.x == 1 ? n_s : n_t;
        s_min[threadIdx.x] = f32_asc_key(INFINITY); s_max[threadIdx.x] = f32_asc_key(-INFINITY);   // the folds' start values
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < n; p += blockDim.x) {
        uint64_t id;
        if (p < n_d) { id = dense[(size_t)q * n_d + p]; if (id == UINT64_MAX) atomicMin(&len[0], p); }
        else if (p < n_d + n_s) { id = sparse[(size_t)q * n_s + (p - n_d)]; if (id == UINT64_MAX) atomicMin(&len[1], p - n_d); }
        else { id = text[(size_t)q * n_t + (p - n_d - n_s)]; if (id == UINT64_MAX) atomicMin(&len[2], p - n_d - n_s); }
        ids[p] = id;
    }
    __syncthreads();
    const uint32_t l0 = len[0], l1 = len[1], l2 = len[2];
    auto list_of = [&](uint32_t p, uint32_t& pos) { if (p < n_d) { pos = p; return 0; } if (p < n_d + n_s) { pos = p - n_d; return 1; } pos = p - n_d - n_s; return 2; };
    auto score_of = [&](int l, uint32_t pos) {
        return l == 0 ? fs.dense[(size_t)q * n_d + pos] : l == 1 ? fs.sparse[(size_t)q * n_s + pos] : fs.text[(size_t)q * n_t + pos];
    };
    if (MODE == 2) {
        // fold(NEG_INFINITY, f32::max) / fold(INFINITY, f32::min) over the list's valid entries (NaNs are ignored by both)
        for (uint32_t p = threadIdx.x; p < n; p += blockDim.x) {
            uint32_t pos; const int l = list_of(p, pos);
            if (pos < len[l]) {
                const float v = score_of(l, pos);
                if (v == v) { atomicMin(&s_min[l], f32_asc_key(v)); atomicMax(&s_max[l], f32_asc_key(v)); }
            }
        }
        __syncthreads();
    }
    for (uint32_t p = threadIdx.x; p < n; p += blockDim.x) {
        uint32_t pos; const int l = list_of(p, pos);
        float c = 0.0f;
        if (pos < len[l]) {
            if (MODE == 0) c = __fdiv_rn(1.0f, __fadd_rn(k, (float)(pos + 1)));
            else {
                float v = score_of(l, pos);
                if (MODE == 2) {
                    const uint32_t a = s_min[l], b = s_max[l];                       // images back to floats (+0.0 for either zero)
                    const float mn = __uint_as_float((a & 0x80000000u) ? (a & 0x7fffffffu) : ~a);
                    const float mx = __uint_as_float((b & 0x80000000u) ? (b & 0x7fffffffu) : ~b);
                    const float range = __fsub_rn(mx, mn);
                    v = range > 0.0f ? __fdiv_rn(__fsub_rn(v, mn), range) : 1.0f;
                }
                c = __fmul_rn(v, l == 0 ? fs.w_dense : l == 1 ? fs.w_sparse : fs.w_text);
            }
        }
        con[p] = c;
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < n_eff; p += blockDim.x) {
        uint64_t key = UINT64_MAX;
        if (p < n) {
            const uint32_t pos = p < n_d ? p : p < n_d + n_s ? p - n_d : p - n_d - n_s;
            const bool valid = p < n_d ? pos < l0 : p < n_d + n_s ? pos < l1 : pos < l2;
            if (valid) {
                const uint64_t id = ids[p];
                bool head = true, have = false;
                float s = 0.0f;
                for (uint32_t j = 0; j < l0; ++j)
                    if (ids[j] == id) {
                        if (j < p) head = false;
                        s = con[j];                                                     // insert: overwrites
                        have = true;
                    }
                for (uint32_t j = 0; j < l1 && head; ++j)
                    if (ids[n_d + j] == id) {
                        if (n_d + j < p) head = false;
                        const float r = con[n_d + j];
                        s = have ? __fadd_rn(s, r) : r;
                        have = true;
                    }
                for (uint32_t j = 0; j < l2 && head; ++j)
                    if (ids[n_d + n_s + j] == id) {
                        if (n_d + n_s + j < p) head = false;
                        const float r = con[n_d + n_s + j];
                        s = have ? __fadd_rn(s, r) : r;
                        have = true;
                    }
                if (head) {
                    sc[p] = s;
                    key = ((uint64_t)(~f32_asc_key(s)) << 32) | p;
                }
            }
        }
        keys[p] = key;
    }
    __syncthreads();
    bitonic_sort_smem(keys, n_eff);
    for (uint32_t t = threadIdx.x; t < limit; t += blockDim.x) {
        const uint64_t key = t < n_eff ? keys[t] : UINT64_MAX;
        if (key != UINT64_MAX) {
            const uint32_t p = (uint32_t)key;
            ids_out[(size_t)q * limit + t] = ids[p];
            scores_out[(size_t)q * limit + t] = sc[p];
        } else {
            ids_out[(size_t)q * limit + t] = UINT64_MAX;
            scores_out[(size_t)q * limit + t] = -INFINITY;
        }
    }
}

}  // namespace gvdb
