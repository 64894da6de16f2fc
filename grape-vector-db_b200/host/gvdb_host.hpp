// gvdb_host.hpp — host side above the C ABI, mirroring the reference's Rust interfaces for
// this path with the same names, argument meaning and error behaviour, so that code written
// against the reference reads the same here.  (The reference is Rust; this image has no
// cargo/rustc, so the host mirror is C++17.  The Rust FFI crate a maintainer would add is in
// ../rust/ and INTEGRATION.md.)
//
//   reference (file:line under /root/reference)                 here
//   BinaryQuantizationConfig   src/quantization.rs:11-31        gvdb::BinaryQuantizationConfig
//   BinaryVector               src/quantization.rs:35-63        gvdb::BinaryVector
//   BinaryQuantizer            src/quantization.rs:67-241       gvdb::BinaryQuantizer
//   trait VectorIndex          src/index.rs:35-62               gvdb::VectorIndex
//   IndexStats                 src/index.rs:82-88               gvdb::IndexStats
//   FaissVectorIndex (Flat)    src/index.rs:330-683             gvdb::GpuVectorIndex (exact or two-stage)
//   VectorDbError              src/types.rs:859-920             gvdb::VectorDbError
//   rrf_fusion                 src/hybrid.rs:422-488            gvdb::rrf_fusion
//   linear_ / normalized_fusion src/hybrid.rs:491-616           gvdb::linear_fusion, gvdb::normalized_fusion (GPU: gvdb_weighted_fusion_batch)
//   SparseIndex / BM25         src/sparse.rs:31-222             gvdb::SparseIndex (host, as the reference),
//                                                               gvdb::GpuSparseIndex (postings scored on the GPU)
//   shard merge                src/distributed/shard.rs:776-783 gvdb::concat_sort_truncate
//
// All arithmetic on the dense path runs on the GPU through include/gvdb.h.  RRF is host code in
// the reference and stays host code here ("consumes the GPU top-k unchanged"); BM25 is available
// both as the reference's host loop (SparseIndex) and on the GPU (GpuSparseIndex, gvdb_sparse_*).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "gvdb.h"

namespace gvdb {

// ---- VectorDbError (src/types.rs:859-920): the variants this path emits -------------------
struct VectorDbError : std::runtime_error {
    enum Kind { IndexNotBuilt, DimensionMismatch, InvalidVectorDimension, QuantizationError,
                IndexError, ConfigError, NotImplemented };
    Kind kind;
    size_t expected = 0, actual = 0;   // DimensionMismatch { expected, actual }
    VectorDbError(Kind k, const std::string& m, size_t e = 0, size_t a = 0)
        : std::runtime_error(m), kind(k), expected(e), actual(a) {}
};

inline void check(gvdb_status st) {
    if (st == GVDB_OK) return;
    const std::string msg = gvdb_last_error();
    switch (st) {
        case GVDB_ERR_INDEX_NOT_BUILT: throw VectorDbError(VectorDbError::IndexNotBuilt, msg);
        case GVDB_ERR_DIMENSION_MISMATCH: throw VectorDbError(VectorDbError::DimensionMismatch, msg);
        case GVDB_ERR_INVALID_VECTOR_DIMENSION: throw VectorDbError(VectorDbError::InvalidVectorDimension, msg);
        case GVDB_ERR_QUANTIZATION: throw VectorDbError(VectorDbError::QuantizationError, msg);
        case GVDB_ERR_INVALID_ARGUMENT: throw VectorDbError(VectorDbError::ConfigError, msg);
        case GVDB_ERR_NOT_IMPLEMENTED: throw VectorDbError(VectorDbError::NotImplemented, msg);
        default: throw VectorDbError(VectorDbError::IndexError, msg);
    }
}

// ---- BinaryQuantizationConfig (src/quantization.rs:11-31), same defaults -------------------
struct BinaryQuantizationConfig {
    float threshold = 0.0f;
    bool enable_simd = true;      // dead flag in the reference (:196-203); kept for signature parity
    float rescore_ratio = 0.1f;
    bool enable_cache = true;     // the reference's Debug-string cache (:89) is not reproduced
    int device = 0;               // placement (not in the reference)
};

// ---- BinaryVector (src/quantization.rs:35-63) ---------------------------------------------------
struct BinaryVector {
    std::vector<uint8_t> data;    // BitVec<u8, Msb0> storage: bit j -> byte j/8, bit 7-(j%8)
    size_t dimension = 0;
    size_t byte_size() const { return (dimension + 7) / 8; }                 // :49-51
    std::vector<uint8_t> to_bytes() const { return data; }                   // :54-56
    static BinaryVector from_bytes(std::vector<uint8_t> bytes, size_t dimension) {   // :59-62
        return BinaryVector{std::move(bytes), dimension};
    }
};

struct CacheStats { bool enabled; size_t size; size_t capacity; };            // :244-249

class Handle {   // RAII over gvdb_index*
public:
    Handle(uint32_t dim, float threshold, float ratio, int device, uint64_t capacity = 0, uint64_t row_base = 0) {
        gvdb_config cfg{};
        cfg.struct_size = sizeof(cfg);
        cfg.dim = dim; cfg.threshold = threshold; cfg.rescore_ratio = ratio; cfg.device = device;
        cfg.capacity_rows = capacity; cfg.row_base = row_base;
        check(gvdb_create(&cfg, &h_));
    }
    ~Handle() { gvdb_destroy(h_); }
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;
    gvdb_index* get() const { return h_; }
private:
    gvdb_index* h_ = nullptr;
};

// ---- BinaryQuantizer (src/quantization.rs:67-241) --------------------------------------------------
class BinaryQuantizer {
public:
    explicit BinaryQuantizer(BinaryQuantizationConfig config) : config_(config) {}

    // quantize(&mut self, &[f32]) -> Result<BinaryVector>   (:86-122)
    BinaryVector quantize(const std::vector<float>& vector) {
        BinaryVector out;
        out.dimension = vector.size();
        out.data.assign(out.byte_size(), 0);
        if (vector.empty()) return out;
        Handle& h = handle(vector.size());
        check(gvdb_quantize(h.get(), vector.data(), 1, out.data.data()));
        return out;
    }
    // quantize_batch(&mut self, &[Vec<f32>]) -> Result<Vec<BinaryVector>>   (:125-127)
    std::vector<BinaryVector> quantize_batch(const std::vector<std::vector<float>>& vectors) {
        std::vector<BinaryVector> out(vectors.size());
        size_t i = 0;
        while (i < vectors.size()) {            // one GPU call per run of equal dimensions
            size_t j = i, dim = vectors[i].size();
            while (j < vectors.size() && vectors[j].size() == dim) ++j;
            if (dim == 0) { for (; i < j; ++i) out[i] = BinaryVector{{}, 0}; continue; }
            std::vector<float> flat((j - i) * dim);
            for (size_t r = i; r < j; ++r) std::copy(vectors[r].begin(), vectors[r].end(), flat.begin() + (r - i) * dim);
            const size_t nb = (dim + 7) / 8;
            std::vector<uint8_t> codes((j - i) * nb);
            check(gvdb_quantize(handle(dim).get(), flat.data(), j - i, codes.data()));
            for (size_t r = i; r < j; ++r)
                out[r] = BinaryVector{std::vector<uint8_t>(codes.begin() + (r - i) * nb, codes.begin() + (r - i + 1) * nb), dim};
            i = j;
        }
        return out;
    }
    // hamming_distance(&self, a, b) -> Result<f32>   (:130-141): InvalidVectorDimension on mismatch
    float hamming_distance(const BinaryVector& a, const BinaryVector& b) {
        if (a.dimension != b.dimension)
            throw VectorDbError(VectorDbError::InvalidVectorDimension, "invalid vector dimension");
        if (a.dimension == 0) return 0.0f;
        // one stored row (b) scanned by one query code (a) on the GPU
        Handle h((uint32_t)a.dimension, config_.threshold, config_.rescore_ratio, config_.device);
        add_codes_as_rows(h, {b});
        uint32_t d = 0;
        check(gvdb_hamming(h.get(), a.data.data(), 1, &d));
        return (float)d;
    }
    // similarity(&self, a, b) -> Result<f32>   (:144-148)
    float similarity(const BinaryVector& a, const BinaryVector& b) {
        float distance = hamming_distance(a, b);
        float max_distance = (float)a.dimension;
        return 1.0f - (distance / max_distance);
    }
    // multi_stage_search(&self, query_binary, candidates_binary, original_query,
    //                    original_candidates) -> Result<Vec<(usize, f32)>>   (:151-193)
    // Drop-in form: the caller owns the candidates, so they are uploaded per call; a resident
    // corpus should use GpuVectorIndex instead.  query_binary / candidates_binary must be the
    // quantizations of the originals under this config (they are what quantize() returns);
    // the GPU re-derives them from the originals, bit-identically.
    std::vector<std::pair<size_t, float>> multi_stage_search(
        const BinaryVector& query_binary, const std::vector<BinaryVector>& candidates_binary,
        const std::vector<float>& original_query, const std::vector<std::vector<float>>& original_candidates) {
        if (candidates_binary.size() != original_candidates.size())
            throw VectorDbError(VectorDbError::QuantizationError,
                                "Mismatch between binary and original candidate counts");   // :158-162
        const size_t n = original_candidates.size(), dim = original_query.size();
        (void)query_binary;
        const uint64_t r = gvdb_rescore_count(n, config_.rescore_ratio);                      // :178-179
        if (r == 0 || n == 0) return {};
        Handle h((uint32_t)dim, config_.threshold, config_.rescore_ratio, config_.device, n);
        std::vector<float> flat(n * dim);
        for (size_t i = 0; i < n; ++i) {
            if (original_candidates[i].size() != dim)
                throw VectorDbError(VectorDbError::DimensionMismatch, "dimension mismatch", dim, original_candidates[i].size());
            std::copy(original_candidates[i].begin(), original_candidates[i].end(), flat.begin() + i * dim);
        }
        check(gvdb_add(h.get(), flat.data(), n, nullptr));
        std::vector<uint64_t> ids(r);
        std::vector<float> sc(r);
        check(gvdb_search_batch(h.get(), original_query.data(), 1, (uint32_t)r, (uint32_t)r, ids.data(), sc.data(), nullptr, nullptr));
        std::vector<std::pair<size_t, float>> out;
        for (size_t t = 0; t < r && ids[t] != GVDB_NO_ID; ++t) out.emplace_back((size_t)ids[t], sc[t]);
        return out;
    }
    void clear_cache() {}                                                                      // :219-223
    CacheStats get_cache_stats() const { return CacheStats{false, 0, 0}; }                     // :226-240

private:
    Handle& handle(size_t dim) {
        if (!h_ || dim_ != dim) {
            h_.reset(new Handle((uint32_t)dim, config_.threshold, config_.rescore_ratio, config_.device));
            dim_ = dim;
        }
        return *h_;
    }
    // Store codes as rows whose signs reproduce the bits (value = bit ? thr+1 : thr-1).
    void add_codes_as_rows(Handle& h, const std::vector<BinaryVector>& v) {
        const size_t dim = v[0].dimension;
        std::vector<float> flat(v.size() * dim);
        for (size_t i = 0; i < v.size(); ++i)
            for (size_t j = 0; j < dim; ++j)
                flat[i * dim + j] = ((v[i].data[j >> 3] >> (7 - (j & 7))) & 1) ? config_.threshold + 1.0f : config_.threshold - 1.0f;
        check(gvdb_add(h.get(), flat.data(), v.size(), nullptr));
    }
    BinaryQuantizationConfig config_;
    std::unique_ptr<Handle> h_;
    size_t dim_ = 0;
};

// ---- IndexStats (src/index.rs:82-88) and trait VectorIndex (src/index.rs:35-62) ----------------------
struct IndexStats {
    size_t vector_count;
    size_t dimension;
    std::string index_type;
    size_t memory_usage;
};

class VectorIndex {
public:
    virtual ~VectorIndex() = default;
    virtual void add_vector(const std::string& id, const std::vector<float>& vector) = 0;
    virtual void add_vectors(const std::vector<std::pair<std::string, std::vector<float>>>& vectors) = 0;
    virtual std::vector<std::pair<std::string, float>> search(const std::vector<float>& query, size_t k) = 0;
    virtual bool remove_vector(const std::string& id) = 0;
    virtual size_t len() const = 0;
    virtual bool is_empty() const = 0;
    virtual void optimize() = 0;
    virtual void clear() = 0;
    virtual IndexStats get_stats() const = 0;
};

// The GPU index behind `dyn VectorIndex` (wiring point src/lib.rs:256-261).  Score convention is
// the exact reference path's (FaissVectorIndex: distance = 1 - cos ascending, src/index.rs:630-637).
//   Mode::Exact     flat scan                       == FaissVectorIndex::search
//   Mode::TwoStage  1-bit scan + rescoring, R = k*oversample (docs/architecture.md:364-365),
//                   returned as distance = 1 - cos of the rescored candidates
class GpuVectorIndex : public VectorIndex {
public:
    enum class Mode { Exact, TwoStage };
    explicit GpuVectorIndex(Mode mode = Mode::TwoStage, uint32_t oversample = 4, BinaryQuantizationConfig cfg = {})
        : mode_(mode), oversample_(oversample), cfg_(cfg) {}

    void add_vector(const std::string& id, const std::vector<float>& vector) override {
        add_vectors({{id, vector}});
    }
    void add_vectors(const std::vector<std::pair<std::string, std::vector<float>>>& vectors) override {
        if (vectors.empty()) return;
        if (!h_) {
            if (vectors[0].second.empty())
                throw VectorDbError(VectorDbError::InvalidVectorDimension, "invalid vector dimension");
            dim_ = vectors[0].second.size();
            h_.reset(new Handle((uint32_t)dim_, cfg_.threshold, cfg_.rescore_ratio, cfg_.device));
        }
        std::vector<float> flat(vectors.size() * dim_);
        for (size_t i = 0; i < vectors.size(); ++i) {
            if (vectors[i].second.size() != dim_)      // src/index.rs:590-594
                throw VectorDbError(VectorDbError::DimensionMismatch, "dimension mismatch", dim_, vectors[i].second.size());
            std::copy(vectors[i].second.begin(), vectors[i].second.end(), flat.begin() + i * dim_);
        }
        uint64_t first = 0;
        check(gvdb_add(h_->get(), flat.data(), vectors.size(), &first));
        for (size_t i = 0; i < vectors.size(); ++i) {
            const std::string& id = vectors[i].first;
            auto it = id_to_index_.find(id);
            if (it != id_to_index_.end()) {           // re-insert: the old row is orphaned, like
                int32_t was = 0;                      // the overwritten map entries of :607-608
                check(gvdb_remove(h_->get(), it->second, &was));
                index_to_id_.erase(it->second);
            }
            id_to_index_[id] = first + i;
            index_to_id_[first + i] = id;
        }
    }
    std::vector<std::pair<std::string, float>> search(const std::vector<float>& query, size_t k) override {
        if (!h_) throw VectorDbError(VectorDbError::IndexNotBuilt, "index not built");        // :621-623
        if (query.size() != dim_)
            throw VectorDbError(VectorDbError::DimensionMismatch, "dimension mismatch", dim_, query.size());
        std::vector<std::pair<std::string, float>> out;
        if (k == 0) return out;
        std::vector<uint64_t> ids(k);
        std::vector<float> val(k);
        if (mode_ == Mode::Exact) {
            check(gvdb_flat_search_batch(h_->get(), query.data(), 1, (uint32_t)k, ids.data(), val.data()));
            for (size_t t = 0; t < k && ids[t] != GVDB_NO_ID; ++t) out.emplace_back(index_to_id_.at(ids[t]), val[t]);
        } else {
            const uint32_t r = (uint32_t)(k * oversample_);
            check(gvdb_search_batch(h_->get(), query.data(), 1, (uint32_t)k, r, ids.data(), val.data(), nullptr, nullptr));
            for (size_t t = 0; t < k && ids[t] != GVDB_NO_ID; ++t) out.emplace_back(index_to_id_.at(ids[t]), 1.0f - val[t]);
        }
        return out;
    }
    // search restricted to the documents a filter allows — the id list FilterEngine::execute_filter returns
    // (src/filtering.rs:374).  Unknown ids are ignored; the answer is the search over the allowed live rows.
    std::vector<std::pair<std::string, float>> search_filtered(const std::vector<float>& query, size_t k,
                                                               const std::vector<std::string>& allowed_ids) {
        if (!h_) throw VectorDbError(VectorDbError::IndexNotBuilt, "index not built");
        if (query.size() != dim_)
            throw VectorDbError(VectorDbError::DimensionMismatch, "dimension mismatch", dim_, query.size());
        std::vector<std::pair<std::string, float>> out;
        if (k == 0) return out;
        gvdb_stats st{};
        check(gvdb_get_stats(h_->get(), &st));
        std::vector<uint32_t> allow((st.rows + 31) / 32, 0u);
        for (const std::string& id : allowed_ids) {
            auto it = id_to_index_.find(id);
            if (it != id_to_index_.end()) allow[it->second >> 5] |= 1u << (it->second & 31);
        }
        std::vector<uint64_t> ids(k);
        std::vector<float> val(k);
        if (mode_ == Mode::Exact) {
            check(gvdb_flat_search_batch_filtered(h_->get(), query.data(), allow.data(), 1, (uint32_t)k, ids.data(), val.data()));
            for (size_t t = 0; t < k && ids[t] != GVDB_NO_ID; ++t) out.emplace_back(index_to_id_.at(ids[t]), val[t]);
        } else {
            const uint32_t r = (uint32_t)(k * oversample_);
            check(gvdb_search_batch_filtered(h_->get(), query.data(), allow.data(), 1, (uint32_t)k, r, ids.data(), val.data()));
            for (size_t t = 0; t < k && ids[t] != GVDB_NO_ID; ++t) out.emplace_back(index_to_id_.at(ids[t]), 1.0f - val[t]);
        }
        return out;
    }
    bool remove_vector(const std::string& id) override {                                       // :642-650
        auto it = id_to_index_.find(id);
        if (it == id_to_index_.end()) return false;
        int32_t was = 0;
        check(gvdb_remove(h_->get(), it->second, &was));
        index_to_id_.erase(it->second);
        id_to_index_.erase(it);
        return true;
    }
    size_t len() const override { return id_to_index_.size(); }                               // :652-654
    bool is_empty() const override { return id_to_index_.empty(); }
    void optimize() override {}
    void clear() override {                                                                    // :664-670
        if (h_) check(gvdb_clear(h_->get()));
        id_to_index_.clear(); index_to_id_.clear();
        h_.reset(); dim_ = 0;
    }
    IndexStats get_stats() const override {                                                    // :672-681
        gvdb_stats s{};
        if (h_) check(gvdb_get_stats(h_->get(), &s));
        return IndexStats{len(), dim_, mode_ == Mode::Exact ? "GpuFlat" : "GpuBinaryTwoStage", (size_t)s.memory_usage};
    }

private:
    Mode mode_;
    uint32_t oversample_;
    BinaryQuantizationConfig cfg_;
    std::unique_ptr<Handle> h_;
    size_t dim_ = 0;
    std::unordered_map<std::string, uint64_t> id_to_index_;
    std::unordered_map<uint64_t, std::string> index_to_id_;
};

// ---- rrf_fusion (src/hybrid.rs:422-488) ------------------------------------------------------------------
struct ScoreBreakdown {                       // src/types.rs:436-446
    bool has_dense = false, has_sparse = false, has_text = false;
    float dense_score = 0, sparse_score = 0, text_score = 0, final_score = 0;
};
struct Fused { std::string id; float score; ScoreBreakdown breakdown; };

// score(id) = sum over (dense, sparse, text) of 1.0 / (k + (rank+1) as f32); the dense list
// inserts (a repeated id overwrites, :439), later lists add.  Sorted by score descending; the
// reference's order among exact ties is HashMap iteration order (unspecified) — here ties keep
// first-appearance order, which is one of the orders the reference can produce.
inline std::vector<Fused> rrf_fusion(const std::vector<std::pair<std::string, float>>& dense_results,
                                     const std::vector<std::pair<std::string, float>>& sparse_results,
                                     const std::vector<std::pair<std::string, float>>& text_results, float k) {
    std::unordered_map<std::string, size_t> pos;
    std::vector<Fused> docs;
    for (size_t rank = 0; rank < dense_results.size(); ++rank) {
        const float rrf = 1.0f / (k + (float)(rank + 1));
        ScoreBreakdown b; b.has_dense = true; b.dense_score = dense_results[rank].second; b.final_score = rrf;
        auto it = pos.find(dense_results[rank].first);
        if (it == pos.end()) { pos[dense_results[rank].first] = docs.size(); docs.push_back({dense_results[rank].first, rrf, b}); }
        else { docs[it->second].score = rrf; docs[it->second].breakdown = b; }
    }
    auto add = [&](const std::vector<std::pair<std::string, float>>& list, int which) {
        for (size_t rank = 0; rank < list.size(); ++rank) {
            const float rrf = 1.0f / (k + (float)(rank + 1));
            auto it = pos.find(list[rank].first);
            if (it == pos.end()) {
                ScoreBreakdown b; b.final_score = rrf;
                if (which == 1) { b.has_sparse = true; b.sparse_score = list[rank].second; }
                else { b.has_text = true; b.text_score = list[rank].second; }
                pos[list[rank].first] = docs.size();
                docs.push_back({list[rank].first, rrf, b});
            } else {
                Fused& d = docs[it->second];
                d.score = d.score + rrf;
                if (which == 1) { d.breakdown.has_sparse = true; d.breakdown.sparse_score = list[rank].second; }
                else { d.breakdown.has_text = true; d.breakdown.text_score = list[rank].second; }
                d.breakdown.final_score = d.score;
            }
        }
    };
    add(sparse_results, 1);
    add(text_results, 2);
    std::stable_sort(docs.begin(), docs.end(), [](const Fused& a, const Fused& b) { return a.score > b.score; });
    return docs;
}

// ---- linear_fusion / normalized_fusion (src/hybrid.rs:491-616) on the GPU -----------------------------------------
// The String ids are numbered by first appearance, the three lists go to gvdb_weighted_fusion_batch (one request),
// and the fused order comes back; the ScoreBreakdown is filled from the input lists as the reference does
// (dense_score / sparse_score / text_score = the list's score — the normalised one for normalized_fusion —,
// final_score = the fused score).
namespace detail {
inline std::vector<float> normalize_scores(const std::vector<std::pair<std::string, float>>& r) {   // :589-616, for the breakdown
    std::vector<float> out(r.size());
    if (r.empty()) return out;
    float mx = -INFINITY, mn = INFINITY;
    for (auto& e : r) { if (e.second == e.second) { mx = e.second > mx ? e.second : mx; mn = e.second < mn ? e.second : mn; } }
    const float range = mx - mn;
    for (size_t i = 0; i < r.size(); ++i) out[i] = range > 0.0f ? (r[i].second - mn) / range : 1.0f;
    return out;
}
inline std::vector<Fused> weighted_fusion(const std::vector<std::pair<std::string, float>>& dense_results,
                                          const std::vector<std::pair<std::string, float>>& sparse_results,
                                          const std::vector<std::pair<std::string, float>>& text_results,
                                          float dense_weight, float sparse_weight, float text_weight, bool normalize, int device) {
    const std::vector<std::pair<std::string, float>>* lists[3] = {&dense_results, &sparse_results, &text_results};
    std::unordered_map<std::string, uint64_t> number;
    std::vector<std::string> name;
    std::vector<uint64_t> ids[3];
    std::vector<float> sc[3], shown[3];
    for (int l = 0; l < 3; ++l) {
        shown[l] = normalize ? normalize_scores(*lists[l]) : std::vector<float>();
        for (size_t i = 0; i < lists[l]->size(); ++i) {
            auto it = number.find((*lists[l])[i].first);
            if (it == number.end()) { it = number.emplace((*lists[l])[i].first, (uint64_t)name.size()).first; name.push_back((*lists[l])[i].first); }
            ids[l].push_back(it->second);
            sc[l].push_back((*lists[l])[i].second);
            if (!normalize) shown[l].push_back((*lists[l])[i].second);
        }
    }
    std::vector<Fused> out;
    if (name.empty()) return out;
    std::vector<uint64_t> oi(name.size());
    std::vector<float> os(name.size());
    check(gvdb_weighted_fusion_batch(device, ids[0].data(), sc[0].data(), (uint32_t)ids[0].size(), ids[1].data(), sc[1].data(),
                                     (uint32_t)ids[1].size(), ids[2].data(), sc[2].data(), (uint32_t)ids[2].size(), 1,
                                     dense_weight, sparse_weight, text_weight, normalize ? 1 : 0, (uint32_t)name.size(),
                                     oi.data(), os.data()));
    std::unordered_map<uint64_t, size_t> at;
    for (size_t t = 0; t < oi.size() && oi[t] != GVDB_NO_ID; ++t) {
        ScoreBreakdown b; b.final_score = os[t];
        at[oi[t]] = out.size();
        out.push_back({name[oi[t]], os[t], b});
    }
    for (int l = 0; l < 3; ++l)
        for (size_t i = 0; i < ids[l].size(); ++i) {
            ScoreBreakdown& b = out[at[ids[l][i]]].breakdown;
            if (l == 0) { b.has_dense = true; b.dense_score = shown[l][i]; }
            else if (l == 1) { b.has_sparse = true; b.sparse_score = shown[l][i]; }
            else { b.has_text = true; b.text_score = shown[l][i]; }
        }
    return out;
}
}  // namespace detail

inline std::vector<Fused> linear_fusion(const std::vector<std::pair<std::string, float>>& dense_results,
                                        const std::vector<std::pair<std::string, float>>& sparse_results,
                                        const std::vector<std::pair<std::string, float>>& text_results,
                                        float dense_weight, float sparse_weight, float text_weight, int device = 0) {
    return detail::weighted_fusion(dense_results, sparse_results, text_results, dense_weight, sparse_weight, text_weight, false, device);
}
inline std::vector<Fused> normalized_fusion(const std::vector<std::pair<std::string, float>>& dense_results,
                                            const std::vector<std::pair<std::string, float>>& sparse_results,
                                            const std::vector<std::pair<std::string, float>>& text_results,
                                            float dense_weight, float sparse_weight, float text_weight, int device = 0) {
    return detail::weighted_fusion(dense_results, sparse_results, text_results, dense_weight, sparse_weight, text_weight, true, device);
}

// ---- SparseIndex / BM25 (src/sparse.rs:31-222) --------------------------------------------------------------
struct BM25Parameters { float k1 = 1.2f; float b = 0.75f; };                  // :42-53
struct SparseVector { std::vector<uint32_t> indices; std::vector<float> values; size_t dimension = 0; };
struct DocumentSparseRepresentation {
    std::string document_id;
    std::vector<std::pair<uint32_t, float>> term_frequencies;   // HashMap<u32,f32> in the reference
    float document_length = 0;
};

class SparseIndex {
public:
    explicit SparseIndex(BM25Parameters p = {}) : params_(p) {}
    virtual ~SparseIndex() = default;
    // add_document (:71-107).  average_document_length is recomputed as the reference does: the
    // sum of document_length over ALL postings entries (a document counts once per distinct
    // term) divided by total_documents.  The reference sums in HashMap order (unspecified); a
    // running f32 total in insertion order is used here.
    virtual void add_document(const DocumentSparseRepresentation& doc) {
        for (auto& tf : doc.term_frequencies) {
            postings_[tf.first].push_back(Entry{doc.document_id, tf.second, doc.document_length});
            total_length_ = total_length_ + doc.document_length;
        }
        total_documents_ += 1;
        if (total_documents_ > 0) average_document_length_ = total_length_ / (float)total_documents_;
    }
    virtual size_t total_documents() const { return total_documents_; }
    virtual float average_document_length() const { return average_document_length_; }
    // search_bm25 (:153-199)
    virtual std::vector<std::pair<std::string, float>> search_bm25(const SparseVector& query_vector, size_t limit) const {
        std::vector<std::pair<std::string, float>> out;
        if (total_documents_ == 0) return out;                                   // :161-163
        std::unordered_map<std::string, size_t> pos;
        for (size_t i = 0; i < query_vector.indices.size() && i < query_vector.values.size(); ++i) {
            auto it = postings_.find(query_vector.indices[i]);
            if (it == postings_.end()) continue;
            const size_t df = it->second.size();                                 // :170-174
            const float idf = std::log(((float)total_documents_ - (float)df + 0.5f) / ((float)df + 0.5f));   // :202-204
            for (const Entry& e : it->second) {
                const float k1 = params_.k1, b = params_.b;
                const float tf_component = (e.term_frequency * (k1 + 1.0f)) /
                    (e.term_frequency + k1 * (1.0f - b + b * (e.document_length / average_document_length_)));
                const float s = query_vector.values[i] * tf_component * idf;      // :207-222
                auto p = pos.find(e.document_id);
                if (p == pos.end()) { pos[e.document_id] = out.size(); out.emplace_back(e.document_id, 0.0f + s); }
                else out[p->second].second = out[p->second].second + s;
            }
        }
        std::stable_sort(out.begin(), out.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
        if (out.size() > limit) out.resize(limit);                               // :196
        return out;
    }
protected:
    BM25Parameters params_;
private:
    struct Entry { std::string document_id; float term_frequency; float document_length; };
    std::unordered_map<uint32_t, std::vector<Entry>> postings_;
    size_t total_documents_ = 0;
    float total_length_ = 0.0f, average_document_length_ = 0.0f;
};

// SparseIndex with the postings scored on the GPU (gvdb_sparse_*).  Documents
// are numbered in insertion order; the postings are frozen into CSR and uploaded on the first
// search after a change.  Score ties come out by document number (the reference's order among
// ties is its HashMap's, i.e. unspecified).  Batched searches share one call.
class GpuSparseIndex : public SparseIndex {
public:
    explicit GpuSparseIndex(BM25Parameters p = {}, int device = 0) : SparseIndex(p) {
        check(gvdb_sparse_create(device, p.k1, p.b, &h_));
    }
    ~GpuSparseIndex() override { gvdb_sparse_destroy(h_); }
    GpuSparseIndex(const GpuSparseIndex&) = delete;
    GpuSparseIndex& operator=(const GpuSparseIndex&) = delete;

    void add_document(const DocumentSparseRepresentation& doc) override {
        const uint32_t d = (uint32_t)ids_.size();
        ids_.push_back(doc.document_id);
        doc_len_.push_back(doc.document_length);
        for (auto& tf : doc.term_frequencies) {
            if (tf.first >= by_term_.size()) by_term_.resize((size_t)tf.first + 1);
            auto& list = by_term_[tf.first];
            if (!list.empty() && list.back().first == d) continue;    // one entry per (term, document)
            list.emplace_back(d, tf.second);
        }
        dirty_ = true;
    }
    size_t total_documents() const override { return ids_.size(); }
    float average_document_length() const override { freeze(); return gvdb_sparse_average_document_length(h_); }

    std::vector<std::pair<std::string, float>> search_bm25(const SparseVector& query_vector, size_t limit) const override {
        return search_bm25_batch({query_vector}, limit)[0];
    }
    std::vector<std::vector<std::pair<std::string, float>>> search_bm25_batch(const std::vector<SparseVector>& queries,
                                                                              size_t limit) const {
        std::vector<std::vector<std::pair<std::string, float>>> out(queries.size());
        if (queries.empty() || limit == 0 || ids_.empty()) return out;
        freeze();
        std::vector<uint64_t> q_off(queries.size() + 1, 0);
        std::vector<uint32_t> q_terms;
        std::vector<float> q_tfs;
        for (size_t q = 0; q < queries.size(); ++q) {
            const size_t n = std::min(queries[q].indices.size(), queries[q].values.size());     // zip (:166)
            q_terms.insert(q_terms.end(), queries[q].indices.begin(), queries[q].indices.begin() + n);
            q_tfs.insert(q_tfs.end(), queries[q].values.begin(), queries[q].values.begin() + n);
            q_off[q + 1] = q_terms.size();
        }
        std::vector<uint64_t> docs(queries.size() * limit);
        std::vector<float> scores(queries.size() * limit);
        check(gvdb_sparse_search_bm25_batch(h_, (uint32_t)queries.size(), q_off.data(), q_terms.data(), q_tfs.data(),
                                            (uint32_t)limit, docs.data(), scores.data()));
        for (size_t q = 0; q < queries.size(); ++q)
            for (size_t t = 0; t < limit && docs[q * limit + t] != GVDB_NO_ID; ++t)
                out[q].emplace_back(ids_[docs[q * limit + t]], scores[q * limit + t]);
        return out;
    }
private:
    void freeze() const {
        if (!dirty_) return;
        std::vector<uint64_t> post_off(by_term_.size() + 1, 0);
        std::vector<uint32_t> post_doc;
        std::vector<float> post_tf;
        for (size_t t = 0; t < by_term_.size(); ++t) {
            for (auto& e : by_term_[t]) { post_doc.push_back(e.first); post_tf.push_back(e.second); }
            post_off[t + 1] = post_doc.size();
        }
        check(gvdb_sparse_build(h_, ids_.size(), (uint32_t)by_term_.size(), post_off.data(), post_doc.data(),
                                post_tf.data(), doc_len_.data()));
        dirty_ = false;
    }
    gvdb_sparse* h_ = nullptr;
    std::vector<std::string> ids_;
    std::vector<float> doc_len_;
    std::vector<std::vector<std::pair<uint32_t, float>>> by_term_;
    mutable bool dirty_ = true;
};

// ---- HybridSearchEngine (src/hybrid.rs:166-356), dense side behind `VectorIndex` ------------------------------
// The reference holds the concrete HnswVectorIndex (:191-206); with the field typed as the trait
//  the GPU index slots in and rrf_fusion consumes its list unchanged.
struct HybridSearchRequest {                    // src/types.rs HybridSearchRequest, the fields this path reads
    std::vector<float> dense_vector;            // empty = None
    SparseVector sparse_vector;                 // empty indices = None
    size_t limit = 10;
};

class HybridSearchEngine {
public:
    HybridSearchEngine(std::shared_ptr<VectorIndex> dense_engine, std::shared_ptr<SparseIndex> sparse_engine,
                       float rrf_k = 60.0f)     // FusionStrategy::RRF { k } (:61-66), default 60.0
        : dense_(std::move(dense_engine)), sparse_(std::move(sparse_engine)), k_(rrf_k) {}
    // add_document (:225-247)
    void add_document(const std::string& id, const std::vector<float>& embedding,
                      const DocumentSparseRepresentation* sparse_repr) {
        dense_->add_vector(id, embedding);
        if (sparse_repr) sparse_->add_document(*sparse_repr);
    }
    // search (:286-356): dense list of 2*limit (:295-298), BM25 list of 2*limit (:305-308), fusion
    // (:331-333), first `limit` (:336)
    std::vector<Fused> search(const HybridSearchRequest& request) const {
        std::vector<std::pair<std::string, float>> dense_results, sparse_results;
        if (!request.dense_vector.empty()) dense_results = dense_->search(request.dense_vector, request.limit * 2);
        if (!request.sparse_vector.indices.empty())
            sparse_results = sparse_->search_bm25(request.sparse_vector, request.limit * 2);
        std::vector<Fused> fused = rrf_fusion(dense_results, sparse_results, {}, k_);
        if (fused.size() > request.limit) fused.resize(request.limit);
        return fused;
    }
private:
    std::shared_ptr<VectorIndex> dense_;
    std::shared_ptr<SparseIndex> sparse_;
    float k_;
};

// ---- scatter/gather merge rule (src/distributed/shard.rs:776-783) -----------------------------------------------
inline std::vector<std::pair<std::string, float>> concat_sort_truncate(
    const std::vector<std::vector<std::pair<std::string, float>>>& shard_results, size_t limit) {
    std::vector<std::pair<std::string, float>> all;
    for (auto& s : shard_results) all.insert(all.end(), s.begin(), s.end());
    std::stable_sort(all.begin(), all.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
    if (all.size() > limit) all.resize(limit);
    return all;
}

}  // namespace gvdb
