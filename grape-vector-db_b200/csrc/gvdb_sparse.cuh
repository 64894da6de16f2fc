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
//   bm25_hist_kernel         radix-select levels (11 + 11 + 10 bits): the descending score image, then
//                            the document number among the documents tying at the cut
//   bm25_cut_kernel          the exact 32-bit image of the limit-th best score and the last tying
//                            document kept
//   bm25_compact_kernel      keys (descending image << 32 | doc) inside the cut: exactly
//                            min(limit, touched documents), however many documents tie
//   bm25_topk_kernel         block bitonic sort of those keys, first `limit`
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gvdb_kernels.cuh"

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

// grid.y = query of the chunk; the query's rank-th term (if it has one) against its postings.
// acc_stride = n_docs rounded up to 4 (the later passes read the accumulators as uint4).
__global__ void __launch_bounds__(256)
bm25_accumulate_kernel(const uint64_t* __restrict__ post_off, const uint32_t* __restrict__ post_doc,
                       const float* __restrict__ post_tf, const float* __restrict__ doc_len, uint32_t n_terms,
                       const uint64_t* __restrict__ q_off, const uint32_t* __restrict__ q_terms,
                       const float* __restrict__ q_tfs, const float* __restrict__ q_idf, uint32_t q0, int rank,
                       float k1, float b, float avg_len, uint64_t acc_stride, uint32_t* __restrict__ acc) {
    const uint32_t q = q0 + blockIdx.y;
    const uint64_t t0 = q_off[q], t1 = q_off[q + 1];
    if (t0 + rank >= t1) return;
    const uint32_t term = q_terms[t0 + rank];
    if (term >= n_terms) return;
    const float qtf = q_tfs[t0 + rank], idf = q_idf[t0 + rank];
    const uint64_t p0 = post_off[term], p1 = post_off[term + 1];
    uint32_t* mine = acc + (size_t)blockIdx.y * acc_stride;
    const float k1p1 = __fadd_rn(k1, 1.0f), one_minus_b = __fsub_rn(1.0f, b);
    constexpr int U = 4;                       // postings in flight per thread (independent loads first)
    for (uint64_t base = p0 + (uint64_t)blockIdx.x * blockDim.x * U; base < p1; base += (uint64_t)gridDim.x * blockDim.x * U) {
        uint32_t doc[U], old[U];
        float tf[U], len[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t p = base + (uint64_t)u * blockDim.x + threadIdx.x;
            doc[u] = p < p1 ? __ldg(post_doc + p) : 0xFFFFFFFFu;
            tf[u] = p < p1 ? __ldg(post_tf + p) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (doc[u] != 0xFFFFFFFFu) { len[u] = __ldg(doc_len + doc[u]); old[u] = mine[doc[u]]; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (doc[u] == 0xFFFFFFFFu) continue;
            // (tf * (k1 + 1)) / (tf + k1 * (1 - b + b * (len / avg_len)))
            const float denom = __fadd_rn(tf[u], __fmul_rn(k1, __fadd_rn(one_minus_b, __fmul_rn(b, __fdiv_rn(len[u], avg_len)))));
            const float tfc = __fdiv_rn(__fmul_rn(tf[u], k1p1), denom);
            const float sc = __fmul_rn(__fmul_rn(qtf, tfc), idf);
            const float base_v = old[u] == BM25_ABSENT ? 0.0f : __uint_as_float(old[u]);
            mine[doc[u]] = __float_as_uint(__fadd_rn(base_v, sc));
        }
    }
}

// descending image of a present score; a NaN score cannot be produced by finite inputs
__device__ __forceinline__ uint32_t bm25_desc_image(uint32_t bits) { return ~f32_asc_key(__uint_as_float(bits)); }

// Radix select of the exact cut without sorting the accumulator.  A 32-bit key is fixed 11 + 11 + 10
// bits at a time (BM25_LEVEL_BINS = 2048 bins per level): levels 0-2 over the descending score image
// give thr32 (the want-th best score), above and m; when more documents tie at thr32 than fit
// (m > want), levels 3-5 do the same over the DOCUMENT NUMBER of the tying documents and give doc_thr.
// Exactly want = min(limit, present) documents then pass "image < thr32 or (image == thr32 and
// doc <= doc_thr)", however many share the cut score (BM25 over tf = count/len data ties by the million).
// Histograms are per CTA in shared memory, lanes with equal bins are combined first (__match_any_sync):
// tying scores put millions of documents into one bin, which serialises global atomics.
constexpr int BM25_LEVEL_BINS = 2048;
__host__ __device__ __forceinline__ int bm25_level_shift(int level) { return (level % 3) == 0 ? 21 : (level % 3) == 1 ? 10 : 0; }
__host__ __device__ __forceinline__ int bm25_level_bits(int level) { return (level % 3) == 2 ? 10 : 11; }

__global__ void __launch_bounds__(256)
bm25_hist_kernel(const uint32_t* __restrict__ acc, uint64_t acc_stride, uint32_t limit, uint32_t* __restrict__ hist,
                 Bm25Cut* __restrict__ cut, int level) {
    __shared__ uint32_t sh[BM25_LEVEL_BINS];
    const uint4* mine = reinterpret_cast<const uint4*>(acc + (size_t)blockIdx.y * acc_stride);
    uint32_t* h = hist + (size_t)blockIdx.y * BM25_LEVEL_BINS;
    const Bm25Cut c = cut[blockIdx.y];
    if (level >= 1 && c.present == 0) return;
    if (level >= 3 && c.m == min(limit, c.present)) return;          // no surplus ties: every tie is kept
    for (int i = threadIdx.x; i < BM25_LEVEL_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const int shift = bm25_level_shift(level), bits = bm25_level_bits(level);
    const bool first = (level % 3) == 0;
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
            const uint64_t i = base + (uint64_t)u * blockDim.x + threadIdx.x;
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                bool ok = w[e] != BM25_ABSENT;
                uint32_t key = 0;
                if (ok) {
                    const uint32_t img = bm25_desc_image(w[e]);
                    if (level < 3) key = img;
                    else { ok = img == c.thr32; key = (uint32_t)(i * 4 + e); }
                    if (level == 0) ++present;
                }
                if (ok && !first) ok = (key >> (shift + bits)) == c.prefix;
                const uint32_t bin = ok ? (key >> shift) & ((1u << bits) - 1u) : 0xFFFFFFFFu;
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
    if (level == 0) {
        for (int o = 16; o > 0; o >>= 1) present += __shfl_xor_sync(0xffffffffu, present, o);
        if (lane == 0 && present) atomicAdd(&cut[blockIdx.y].present, present);
    }
}

// one CTA per query: the bin holding the want-th entry of this level's histogram
__global__ void __launch_bounds__(256)
bm25_cut_kernel(const uint32_t* __restrict__ hist, uint32_t limit, Bm25Cut* __restrict__ cut, int level) {
    __shared__ uint32_t part[256];
    const uint32_t* h = hist + (size_t)blockIdx.x * BM25_LEVEL_BINS;
    Bm25Cut* c = cut + blockIdx.x;
    const uint32_t want_all = min(limit, c->present);
    if (level >= 3 && c->m == want_all) { if (threadIdx.x == 0) c->doc_thr = 0xFFFFFFFFu; return; }
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
                const uint32_t prefix = (level % 3) == 0 ? bin : (c->prefix << bm25_level_bits(level)) | bin;
                c->prefix = prefix;
                c->want_left = want - run;                     // rank inside the chosen bin
                if (level == 2) {                              // the image is complete
                    c->thr32 = prefix;
                    c->above = want_all - (want - run);
                    c->m = want_all - (want - run) + hv;
                    c->want_left = want - run;                 // ties to keep, if they do not all fit
                }
                if (level == 5) c->doc_thr = prefix;
                break;
            }
            run += hv;
        }
    }
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

__global__ void fill_f32_kernel(float* __restrict__ p, size_t n, float v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- rrf_fusion (/root/reference/src/hybrid.rs:422-488) on the GPU ------------------------------
// One CTA per query.  The three lists (dense, sparse, text; document numbers, GVDB_NO_ID ends a list
// early) are laid end to end in shared memory.  Entry p is a HEAD when no earlier entry carries its
// document; a head's score is built in the reference's order: the dense loop INSERTS 1/(k + rank+1)
// (a repeated document overwrites), the sparse and text loops add theirs in list order.  Heads are
// ordered by score descending; exact ties (unspecified in the reference: HashMap iteration) by first
// appearance.  n_d + n_s + n_t <= RRF_MAX.
constexpr uint32_t RRF_MAX = 4096;

__global__ void __launch_bounds__(SORT_THREADS)
rrf_fusion_kernel(const uint64_t* __restrict__ dense, uint32_t n_d, const uint64_t* __restrict__ sparse, uint32_t n_s,
                  const uint64_t* __restrict__ text, uint32_t n_t, float k, uint32_t limit,
                  uint64_t* __restrict__ ids_out, float* __restrict__ scores_out) {
    extern __shared__ __align__(16) uint64_t rrf_smem[];
    const uint32_t n = n_d + n_s + n_t;
    const uint32_t n_eff = max(64u, next_pow2(n));
    uint64_t* ids = rrf_smem;                 // [n]
    uint64_t* keys = rrf_smem + n;            // [n_eff]
    float* sc = reinterpret_cast<float*>(keys + n_eff);   // [n]
    __shared__ uint32_t len[3];
    const uint32_t q = blockIdx.x;
    if (threadIdx.x < 3) len[threadIdx.x] = threadIdx.x == 0 ? n_d : threadIdx.x == 1 ? n_s : n_t;
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
                        s = __fdiv_rn(1.0f, __fadd_rn(k, (float)(j + 1)));          // insert: overwrites
                        have = true;
                    }
                for (uint32_t j = 0; j < l1 && head; ++j)
                    if (ids[n_d + j] == id) {
                        if (n_d + j < p) head = false;
                        const float r = __fdiv_rn(1.0f, __fadd_rn(k, (float)(j + 1)));
                        s = have ? __fadd_rn(s, r) : r;
                        have = true;
                    }
                for (uint32_t j = 0; j < l2 && head; ++j)
                    if (ids[n_d + n_s + j] == id) {
                        if (n_d + n_s + j < p) head = false;
                        const float r = __fdiv_rn(1.0f, __fadd_rn(k, (float)(j + 1)));
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
