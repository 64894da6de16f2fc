// gvdb_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A plain, single-file restatement of the arithmetic of the reference's
// quantized-search path, written from the reference's *behaviour* (citations
// below are into /root/reference, file:line).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library; the
// product (grape-vector-db_b200/) never links, imports or calls it.
//
// PARITY STATUS: "parity unpinned" by the reference's own suite.  The reference
// is Rust and neither cargo nor rustc exists in this image, so it cannot be
// executed here, and its tests hold no golden vector for this path (they assert
// shapes only: src/quantization.rs:361-400, src/hybrid.rs:991-1025,
// src/sparse.rs:383-390).  The oracle is pinned instead to known-answer values
// derived by hand from the inputs of those same tests (tests/test_oracle_kat.py,
// SURVEY.md §8c).  Third-party arithmetic restated from its published semantics:
//   hamming 0.1.3  (Cargo.lock:1204-1206)  distance(a,b) = number of differing bits
//   bitvec  1.0.1  (Cargo.lock:326-328)    BitVec<u8,Msb0>: bit j -> byte j/8, bit 7-(j%8)
//
// Floating-point rules: IEEE binary32 everywhere, every sum is a left-to-right
// fold, multiply and add are separate roundings.  Build with
//   g++ -O2 -ffp-contract=off  (never -ffast-math)   — see oracle/Makefile.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>
#include <unordered_map>
#include <vector>

#define GVO_API extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------
// quantize — src/quantization.rs:97-101 (cached) and :114-118 (uncached): one
// `push(value > threshold)` per element into BitVec<u8, Msb0>.
// Strict '>' : threshold itself, -0.0 (for threshold 0) and NaN give bit 0.
// Pad bits of the last byte are 0.
GVO_API void gvo_quantize(const float* x, size_t dim, float threshold, uint8_t* code) {
    size_t nbytes = (dim + 7) / 8;                       // byte_size(), :49-51
    std::memset(code, 0, nbytes);
    for (size_t j = 0; j < dim; ++j)
        if (x[j] > threshold) code[j >> 3] |= (uint8_t)(0x80u >> (j & 7));
}

GVO_API void gvo_quantize_batch(const float* x, size_t n, size_t dim, float threshold,
                                uint8_t* codes) {       // :125-127
    size_t nbytes = (dim + 7) / 8;
    for (size_t i = 0; i < n; ++i) gvo_quantize(x + i * dim, dim, threshold, codes + i * nbytes);
}

// hamming_distance — src/quantization.rs:130-141; hamming::distance counts the
// differing bits of two equal-length byte slices.
GVO_API uint64_t gvo_hamming(const uint8_t* a, const uint8_t* b, size_t nbytes) {
    uint64_t d = 0;
    size_t i = 0;
    for (; i + 8 <= nbytes; i += 8) {
        uint64_t x, y;
        std::memcpy(&x, a + i, 8);
        std::memcpy(&y, b + i, 8);
        d += (uint64_t)__builtin_popcountll(x ^ y);
    }
    for (; i < nbytes; ++i) d += (uint64_t)__builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

// similarity — src/quantization.rs:144-148: 1.0 - (distance as f32 / dim as f32)
GVO_API float gvo_similarity(const uint8_t* a, const uint8_t* b, size_t dim) {
    float distance = (float)gvo_hamming(a, b, (dim + 7) / 8);
    float max_distance = (float)dim;
    return 1.0f - (distance / max_distance);
}

// cosine_similarity_manual — src/quantization.rs:206-216 (identical arithmetic in
// src/storage.rs:851-865): three separate sequential folds, sqrt, 0.0 on a zero norm.
GVO_API float gvo_cosine_similarity(const float* a, const float* b, size_t dim) {
    float dot = 0.0f, sa = 0.0f, sb = 0.0f;
    for (size_t j = 0; j < dim; ++j) dot = dot + a[j] * b[j];
    for (size_t j = 0; j < dim; ++j) sa = sa + a[j] * a[j];
    for (size_t j = 0; j < dim; ++j) sb = sb + b[j] * b[j];
    float na = std::sqrt(sa), nb = std::sqrt(sb);
    if (na == 0.0f || nb == 0.0f) return 0.0f;
    return dot / (na * nb);
}

// cosine_distance — src/index.rs:686-700: INFINITY on a zero norm, else 1 - cos.
// (The length-mismatch INFINITY case cannot occur with one fixed dim.)
GVO_API float gvo_cosine_distance(const float* a, const float* b, size_t dim) {
    float dot = 0.0f, sa = 0.0f, sb = 0.0f;
    for (size_t j = 0; j < dim; ++j) dot = dot + a[j] * b[j];
    for (size_t j = 0; j < dim; ++j) sa = sa + a[j] * a[j];
    for (size_t j = 0; j < dim; ++j) sb = sb + b[j] * b[j];
    float na = std::sqrt(sa), nb = std::sqrt(sb);
    if (na == 0.0f || nb == 0.0f) return std::numeric_limits<float>::infinity();
    return 1.0f - (dot / (na * nb));
}

// rescore_count — src/quantization.rs:178-179: (N as f32 * ratio) as usize, min N.
// Rust's float->usize cast saturates (NaN -> 0, negative -> 0).
GVO_API size_t gvo_rescore_count(size_t n, float ratio) {
    float p = (float)n * ratio;
    size_t r;
    if (!(p > 0.0f)) r = 0;
    else if (p >= 18446744073709551616.0f) r = std::numeric_limits<size_t>::max();
    else r = (size_t)p;
    return r < n ? r : n;
}

// multi_stage_search — src/quantization.rs:151-193.
//   stage 1: similarity for ALL n candidates, stable sort by similarity desc (:165-175)
//   slice [..rescore_count]                                              (:178-179)
//   stage 2: cosine for those, stable sort by cosine desc                (:181-190)
// returns all rescore_count pairs (no truncation to k, :192).
// `rescore_count` is passed as an integer so callers can use either the ratio form
// (gvo_rescore_count) or the documented k*oversample form (docs/architecture.md:364-365).
// Optional outputs expose the stage-1 candidate list (index + Hamming distance).
// Return: number of results, or -1 if a NaN reaches partial_cmp().unwrap() (the
// reference panics there, :175/:190).
GVO_API int64_t gvo_multi_stage_search(const uint8_t* qcode, const uint8_t* codes, const float* q,
                                       const float* rows, size_t n, size_t dim,
                                       size_t rescore_count, uint64_t* out_idx, float* out_score,
                                       uint64_t* cand_idx_out, uint32_t* cand_ham_out) {
    size_t nbytes = (dim + 7) / 8;
    std::vector<std::pair<size_t, float>> s1(n);
    std::vector<uint32_t> ham(n);
    for (size_t i = 0; i < n; ++i) {
        ham[i] = (uint32_t)gvo_hamming(qcode, codes + i * nbytes, nbytes);
        float distance = (float)ham[i];
        s1[i] = {i, 1.0f - (distance / (float)dim)};
        if (std::isnan(s1[i].second)) return -1;
    }
    std::stable_sort(s1.begin(), s1.end(),
                     [](const std::pair<size_t, float>& a, const std::pair<size_t, float>& b) {
                         return a.second > b.second;   // b.1.partial_cmp(&a.1) => descending
                     });
    size_t r = rescore_count < n ? rescore_count : n;
    std::vector<std::pair<size_t, float>> s2(r);
    for (size_t t = 0; t < r; ++t) {
        size_t idx = s1[t].first;
        if (cand_idx_out) cand_idx_out[t] = idx;
        if (cand_ham_out) cand_ham_out[t] = ham[idx];
        s2[t] = {idx, gvo_cosine_similarity(q, rows + idx * dim, dim)};
        if (std::isnan(s2[t].second)) return -1;
    }
    std::stable_sort(s2.begin(), s2.end(),
                     [](const std::pair<size_t, float>& a, const std::pair<size_t, float>& b) {
                         return a.second > b.second;
                     });
    for (size_t t = 0; t < r; ++t) {
        out_idx[t] = s2[t].first;
        out_score[t] = s2[t].second;
    }
    return (int64_t)r;
}

// Same results as gvo_multi_stage_search, but stage 1 selects the top
// rescore_count by the unique key (ham asc, idx asc) instead of sorting all n.
// Equality of the two is asserted in tests/test_oracle_kat.py; this is the
// "mechanically kinder" CPU baseline of BASELINE.md §2.
GVO_API int64_t gvo_multi_stage_search_select(const uint8_t* qcode, const uint8_t* codes,
                                              const float* q, const float* rows, size_t n,
                                              size_t dim, size_t rescore_count, uint64_t* out_idx,
                                              float* out_score) {
    size_t nbytes = (dim + 7) / 8;
    std::vector<uint64_t> key(n);
    for (size_t i = 0; i < n; ++i)
        key[i] = (gvo_hamming(qcode, codes + i * nbytes, nbytes) << 32) | (uint64_t)i;
    size_t r = rescore_count < n ? rescore_count : n;
    if (r < n) std::nth_element(key.begin(), key.begin() + r, key.end());
    std::sort(key.begin(), key.begin() + r);
    std::vector<std::pair<size_t, float>> s2(r);
    for (size_t t = 0; t < r; ++t) {
        size_t idx = (size_t)(key[t] & 0xffffffffu);
        s2[t] = {idx, gvo_cosine_similarity(q, rows + idx * dim, dim)};
        if (std::isnan(s2[t].second)) return -1;
    }
    std::stable_sort(s2.begin(), s2.end(),
                     [](const std::pair<size_t, float>& a, const std::pair<size_t, float>& b) {
                         return a.second > b.second;
                     });
    for (size_t t = 0; t < r; ++t) {
        out_idx[t] = s2[t].first;
        out_score[t] = s2[t].second;
    }
    return (int64_t)r;
}

// FaissVectorIndex::search — src/index.rs:620-640: cosine_distance for every live
// row (:628-633, rows whose id was removed are skipped), stable sort ascending with
// partial_cmp().unwrap_or(Equal) (:636), truncate(k) (:637).
// live == nullptr means every row is live.  Returns the number of results.
// NaN distances (only possible with non-finite inputs) make unwrap_or(Equal) a
// non-strict-weak order in the reference; the oracle rejects them with -1.
GVO_API int64_t gvo_flat_search(const float* q, const float* rows, const uint8_t* live, size_t n,
                                size_t dim, size_t k, uint64_t* out_idx, float* out_dist) {
    std::vector<std::pair<size_t, float>> res;
    res.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        if (live && !live[i]) continue;
        float d = gvo_cosine_distance(q, rows + i * dim, dim);
        if (std::isnan(d)) return -1;
        res.push_back({i, d});
    }
    std::stable_sort(res.begin(), res.end(),
                     [](const std::pair<size_t, float>& a, const std::pair<size_t, float>& b) {
                         return a.second < b.second;
                     });
    size_t r = k < res.size() ? k : res.size();
    for (size_t t = 0; t < r; ++t) {
        out_idx[t] = res[t].first;
        out_dist[t] = res[t].second;
    }
    return (int64_t)r;
}

// BasicVectorStore::vector_search — src/storage.rs:296-339 with its own cosine_similarity
// (:851-865, identical arithmetic to the quantizer's: 0.0 on a zero norm): every stored vector in
// iteration order (here: row order; the reference iterates sled keys), `continue` when a
// threshold is given and similarity < threshold (:312-316), stable sort DESCENDING with
// partial_cmp().unwrap_or(Equal) (:328-332), truncate(limit) (:333).
GVO_API int64_t gvo_similarity_search(const float* q, const float* rows, const uint8_t* live, size_t n,
                                      size_t dim, size_t limit, float threshold, int use_threshold,
                                      uint64_t* out_idx, float* out_sim) {
    std::vector<std::pair<size_t, float>> res;
    res.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        if (live && !live[i]) continue;
        float s = gvo_cosine_similarity(q, rows + i * dim, dim);
        if (std::isnan(s)) return -1;
        if (use_threshold && s < threshold) continue;
        res.push_back({i, s});
    }
    std::stable_sort(res.begin(), res.end(),
                     [](const std::pair<size_t, float>& a, const std::pair<size_t, float>& b) {
                         return a.second > b.second;
                     });
    size_t r = limit < res.size() ? limit : res.size();
    for (size_t t = 0; t < r; ++t) {
        out_idx[t] = res[t].first;
        out_sim[t] = res[t].second;
    }
    return (int64_t)r;
}

// ---------------------------------------------------------------------------
// Batched, multi-threaded drivers for the CPU baseline.  The reference's only
// parallelism on this path is "one query per rayon task"
// (src/performance/parallel_search.rs:129-137); same here.  mode 0 = faithful
// full stable sort, mode 1 = select variant.
GVO_API int64_t gvo_multi_stage_search_batch(const float* queries, size_t nq, const uint8_t* codes,
                                             const float* rows, size_t n, size_t dim,
                                             float threshold, size_t rescore_count, size_t k,
                                             uint64_t* out_idx /*nq*k*/, float* out_score,
                                             int nthreads, int mode) {
    size_t nbytes = (dim + 7) / 8;
    size_t r = rescore_count < n ? rescore_count : n;
    size_t kk = k < r ? k : r;
    std::atomic<size_t> next(0);
    std::atomic<int> bad(0);
    auto work = [&]() {
        std::vector<uint8_t> qcode(nbytes);
        std::vector<uint64_t> idx(r ? r : 1);
        std::vector<float> sc(r ? r : 1);
        for (;;) {
            size_t qi = next.fetch_add(1);
            if (qi >= nq) break;
            const float* q = queries + qi * dim;
            gvo_quantize(q, dim, threshold, qcode.data());
            int64_t got = mode == 0
                ? gvo_multi_stage_search(qcode.data(), codes, q, rows, n, dim, r, idx.data(),
                                         sc.data(), nullptr, nullptr)
                : gvo_multi_stage_search_select(qcode.data(), codes, q, rows, n, dim, r,
                                                idx.data(), sc.data());
            if (got < 0) { bad.store(1); continue; }
            for (size_t t = 0; t < k; ++t) {
                out_idx[qi * k + t] = t < kk ? idx[t] : UINT64_MAX;
                out_score[qi * k + t] = t < kk ? sc[t] : -std::numeric_limits<float>::infinity();
            }
        }
    };
    if (nthreads <= 1) work();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    return bad.load() ? -1 : (int64_t)kk;
}

GVO_API int64_t gvo_flat_search_batch(const float* queries, size_t nq, const float* rows,
                                      const uint8_t* live, size_t n, size_t dim, size_t k,
                                      uint64_t* out_idx, float* out_dist, int nthreads) {
    std::atomic<size_t> next(0);
    std::atomic<int> bad(0);
    auto work = [&]() {
        std::vector<uint64_t> idx(k ? k : 1);
        std::vector<float> ds(k ? k : 1);
        for (;;) {
            size_t qi = next.fetch_add(1);
            if (qi >= nq) break;
            int64_t got = gvo_flat_search(queries + qi * dim, rows, live, n, dim, k, idx.data(),
                                          ds.data());
            if (got < 0) { bad.store(1); continue; }
            for (size_t t = 0; t < k; ++t) {
                out_idx[qi * k + t] = t < (size_t)got ? idx[t] : UINT64_MAX;
                out_dist[qi * k + t] =
                    t < (size_t)got ? ds[t] : std::numeric_limits<float>::infinity();
            }
        }
    };
    if (nthreads <= 1) work();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    return bad.load() ? -1 : 0;
}

// ---------------------------------------------------------------------------
// Cross-shard merge of the two-stage path.  The reference's
// scatter/gather rule is concat + sort by score desc + truncate
// (src/distributed/shard.rs:776-783); for the two-stage path to equal the
// single-index result the merge must first re-apply the stage-1 cut: keep the global
// top-R by (ham asc, global idx asc) among the shards' local top-R lists, and only
// then order by (cos desc, ham asc, idx asc).  Records are (ham, idx, score).
GVO_API int64_t gvo_shard_merge(const uint32_t* ham, const uint64_t* idx, const float* score,
                                size_t nrec, size_t rescore_count, size_t k, uint64_t* out_idx,
                                float* out_score) {
    std::vector<size_t> ord;
    ord.reserve(nrec);
    for (size_t i = 0; i < nrec; ++i)
        if (idx[i] != UINT64_MAX) ord.push_back(i);
    std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
        if (ham[a] != ham[b]) return ham[a] < ham[b];
        return idx[a] < idx[b];
    });
    if (ord.size() > rescore_count) ord.resize(rescore_count);
    std::stable_sort(ord.begin(), ord.end(),
                     [&](size_t a, size_t b) { return score[a] > score[b]; });
    size_t r = k < ord.size() ? k : ord.size();
    for (size_t t = 0; t < r; ++t) {
        out_idx[t] = idx[ord[t]];
        out_score[t] = score[ord[t]];
    }
    return (int64_t)r;
}

// Plain scatter/gather rule of src/distributed/shard.rs:776-783: concat, stable sort
// by score descending, truncate(limit).  Input order = concat order.
GVO_API int64_t gvo_concat_sort_truncate(const uint64_t* idx, const float* score, size_t nrec,
                                         size_t limit, uint64_t* out_idx, float* out_score) {
    std::vector<size_t> ord(nrec);
    std::iota(ord.begin(), ord.end(), (size_t)0);
    std::stable_sort(ord.begin(), ord.end(),
                     [&](size_t a, size_t b) { return score[a] > score[b]; });
    size_t r = limit < nrec ? limit : nrec;
    for (size_t t = 0; t < r; ++t) {
        out_idx[t] = idx[ord[t]];
        out_score[t] = score[ord[t]];
    }
    return (int64_t)r;
}

// ---------------------------------------------------------------------------
// rrf_fusion — src/hybrid.rs:422-488.  score(id) = sum over the lists (dense, sparse,
// text, in that order) of 1.0 / (k + (rank+1) as f32); the first list INSERTS
// (a duplicate id inside the dense list overwrites, :439), later lists add.
// The reference collects a HashMap into a Vec and stable-sorts by score desc, so the
// order among exact ties is unspecified there; the oracle breaks ties by first
// appearance (dense, then sparse, then text) and tests compare per tie group.
// Ids are u64 document numbers (the FFI crate maps them to the String ids).
GVO_API int64_t gvo_rrf_fusion(const uint64_t* dense, size_t nd, const uint64_t* sparse, size_t ns,
                               const uint64_t* text, size_t nt, float k, uint64_t* out_idx,
                               float* out_score, size_t out_cap) {
    std::unordered_map<uint64_t, size_t> pos;
    std::vector<uint64_t> ids;
    std::vector<float> sc;
    for (size_t rank = 0; rank < nd; ++rank) {
        float rrf = 1.0f / (k + (float)(rank + 1));
        auto it = pos.find(dense[rank]);
        if (it == pos.end()) { pos[dense[rank]] = ids.size(); ids.push_back(dense[rank]); sc.push_back(rrf); }
        else sc[it->second] = rrf;                       // HashMap::insert overwrites
    }
    const uint64_t* lists[2] = {sparse, text};
    size_t lens[2] = {ns, nt};
    for (int l = 0; l < 2; ++l)
        for (size_t rank = 0; rank < lens[l]; ++rank) {
            float rrf = 1.0f / (k + (float)(rank + 1));
            uint64_t id = lists[l][rank];
            auto it = pos.find(id);
            if (it == pos.end()) { pos[id] = ids.size(); ids.push_back(id); sc.push_back(rrf); }
            else sc[it->second] = sc[it->second] + rrf;
        }
    std::vector<size_t> ord(ids.size());
    std::iota(ord.begin(), ord.end(), (size_t)0);
    std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return sc[a] > sc[b]; });
    size_t r = out_cap < ord.size() ? out_cap : ord.size();
    for (size_t t = 0; t < r; ++t) {
        out_idx[t] = ids[ord[t]];
        out_score[t] = sc[ord[t]];
    }
    return (int64_t)ids.size();
}

// ---------------------------------------------------------------------------
// linear_fusion — src/hybrid.rs:491-566 — and normalized_fusion — :568-587 with normalize_scores
// :589-616.  weighted = score * weight; the dense loop INSERTS (a duplicate overwrites, :511), the
// sparse and the text loop add (`*current_score += weighted`, :518,:535) or insert; the result is
// sorted by score descending (ties: HashMap order in the reference, first appearance here).
// normalize != 0: each list goes through normalize_scores first: max = fold(NEG_INFINITY, f32::max),
// min = fold(INFINITY, f32::min), range = max - min, score' = range > 0 ? (score - min) / range : 1.0.
static std::vector<float> gvo_normalize_scores(const float* sc, size_t n) {
    std::vector<float> out(n);
    if (n == 0) return out;
    float mx = -INFINITY, mn = INFINITY;
    for (size_t i = 0; i < n; ++i) { if (!(sc[i] != sc[i])) { mx = sc[i] > mx ? sc[i] : mx; mn = sc[i] < mn ? sc[i] : mn; } }   // f32::max / f32::min ignore NaN
    const float range = mx - mn;
    for (size_t i = 0; i < n; ++i) out[i] = range > 0.0f ? (sc[i] - mn) / range : 1.0f;
    return out;
}

GVO_API int64_t gvo_weighted_fusion(const uint64_t* dense, const float* dense_sc, size_t nd, const uint64_t* sparse,
                                    const float* sparse_sc, size_t ns, const uint64_t* text, const float* text_sc,
                                    size_t nt, float w_dense, float w_sparse, float w_text, int normalize,
                                    uint64_t* out_idx, float* out_score, size_t out_cap) {
    std::vector<float> nd_sc, ns_sc, nt_sc;
    if (normalize) {
        nd_sc = gvo_normalize_scores(dense_sc, nd); ns_sc = gvo_normalize_scores(sparse_sc, ns); nt_sc = gvo_normalize_scores(text_sc, nt);
        dense_sc = nd_sc.data(); sparse_sc = ns_sc.data(); text_sc = nt_sc.data();
    }
    std::unordered_map<uint64_t, size_t> pos;
    std::vector<uint64_t> ids;
    std::vector<float> sc;
    for (size_t i = 0; i < nd; ++i) {
        const float w = dense_sc[i] * w_dense;
        auto it = pos.find(dense[i]);
        if (it == pos.end()) { pos[dense[i]] = ids.size(); ids.push_back(dense[i]); sc.push_back(w); }
        else sc[it->second] = w;                          // HashMap::insert overwrites
    }
    const uint64_t* lists[2] = {sparse, text};
    const float* lsc[2] = {sparse_sc, text_sc};
    const float lw[2] = {w_sparse, w_text};
    size_t lens[2] = {ns, nt};
    for (int l = 0; l < 2; ++l)
        for (size_t i = 0; i < lens[l]; ++i) {
            const float w = lsc[l][i] * lw[l];
            auto it = pos.find(lists[l][i]);
            if (it == pos.end()) { pos[lists[l][i]] = ids.size(); ids.push_back(lists[l][i]); sc.push_back(w); }
            else sc[it->second] = sc[it->second] + w;
        }
    std::vector<size_t> ord(ids.size());
    std::iota(ord.begin(), ord.end(), (size_t)0);
    std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return sc[a] > sc[b]; });
    size_t r = out_cap < ord.size() ? out_cap : ord.size();
    for (size_t t = 0; t < r; ++t) {
        out_idx[t] = ids[ord[t]];
        out_score[t] = sc[ord[t]];
    }
    return (int64_t)ids.size();
}

// ---------------------------------------------------------------------------
// BM25 — src/sparse.rs:153-222 over a CSR restatement of the inverted index
// (HashMap<u32, Vec<InvertedIndexEntry>>, :31-38).  For term t the postings are
// post_doc/post_tf[post_off[t] .. post_off[t+1]) in insertion (document) order;
// doc_len[d] is DocumentSparseRepresentation::document_length.
//   idf   = ln((N - df + 0.5) / (df + 0.5))                         (:202-204)
//   score = q_tf * (tf*(k1+1)) / (tf + k1*(1 - b + b*(len/avg))) * idf   (:207-222)
// accumulated per document in query-term order (:166-192); df = postings length
// (document_frequencies is incremented once per (doc, term), :87-89).
// average_document_length is what add_document recomputes (:96-104): the sum of
// document_length over ALL POSTINGS ENTRIES (a document is counted once per distinct
// term) divided by total_documents.  The reference sums in HashMap iteration order
// (unspecified); the oracle sums in term order, in f32.
// Ties in the final sort are unspecified in the reference (HashMap -> Vec); the
// oracle orders ties by document number.
GVO_API float gvo_bm25_avg_len(const uint64_t* post_off, const uint32_t* post_doc,
                               const float* doc_len, size_t n_terms, size_t n_docs) {
    float total = 0.0f;
    for (size_t t = 0; t < n_terms; ++t)
        for (uint64_t p = post_off[t]; p < post_off[t + 1]; ++p) total = total + doc_len[post_doc[p]];
    return n_docs ? total / (float)n_docs : 0.0f;
}

GVO_API int64_t gvo_bm25_search(const uint32_t* q_terms, const float* q_tf, size_t nqt,
                                const uint64_t* post_off, const uint32_t* post_doc,
                                const float* post_tf, const float* doc_len, size_t n_terms,
                                size_t n_docs, float avg_len, float k1, float b, size_t limit,
                                uint64_t* out_doc, float* out_score) {
    if (n_docs == 0) return 0;                                      // :161-163
    std::unordered_map<uint32_t, float> acc;
    std::vector<uint32_t> order;
    for (size_t i = 0; i < nqt; ++i) {
        uint32_t t = q_terms[i];
        if (t >= n_terms || post_off[t] == post_off[t + 1]) continue;   // :169 (term absent)
        size_t df = (size_t)(post_off[t + 1] - post_off[t]);
        float idf = std::log(((float)n_docs - (float)df + 0.5f) / ((float)df + 0.5f));
        for (uint64_t p = post_off[t]; p < post_off[t + 1]; ++p) {
            float tf = post_tf[p], len = doc_len[post_doc[p]];
            float tf_component = (tf * (k1 + 1.0f)) / (tf + k1 * (1.0f - b + b * (len / avg_len)));
            float s = q_tf[i] * tf_component * idf;
            auto it = acc.find(post_doc[p]);
            if (it == acc.end()) { acc[post_doc[p]] = 0.0f + s; order.push_back(post_doc[p]); }
            else it->second = it->second + s;
        }
    }
    std::sort(order.begin(), order.end());
    std::stable_sort(order.begin(), order.end(),
                     [&](uint32_t a, uint32_t c) { return acc[a] > acc[c]; });
    size_t r = limit < order.size() ? limit : order.size();
    for (size_t t = 0; t < r; ++t) {
        out_doc[t] = order[t];
        out_score[t] = acc[order[t]];
    }
    return (int64_t)r;
}

GVO_API int gvo_hardware_threads() { return (int)std::thread::hardware_concurrency(); }
