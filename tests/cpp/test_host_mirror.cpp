// Reads like the reference's own unit tests (src/quantization.rs:356-401, src/hybrid.rs:982-1026,
// src/sparse.rs:378-421) but runs against the C++ host mirror over the C ABI, with the VALUES
// pinned (the reference asserts only shapes).  Needs a B200: `./test_host_mirror` exits 0 on success.
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "gvdb_host.hpp"

using namespace gvdb;

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

static void test_binary_quantization() {            // src/quantization.rs:361-371
    BinaryQuantizationConfig config;
    BinaryQuantizer quantizer(config);
    std::vector<float> vector = {0.5f, -0.3f, 0.8f, -0.1f, 0.2f};
    BinaryVector binary_vec = quantizer.quantize(vector);
    REQUIRE(binary_vec.dimension == 5);
    REQUIRE(binary_vec.data.size() == 1 && binary_vec.byte_size() == 1);
    REQUIRE(binary_vec.data[0] == 0xA8);             // [1,0,1,0,1] Msb0 (the comment at :369)
}

static void test_hamming_distance() {               // src/quantization.rs:374-386
    BinaryQuantizer quantizer(BinaryQuantizationConfig{});
    BinaryVector bin1 = quantizer.quantize({1.0f, -1.0f, 1.0f, -1.0f});
    BinaryVector bin2 = quantizer.quantize({1.0f, 1.0f, -1.0f, -1.0f});
    float distance = quantizer.hamming_distance(bin1, bin2);
    REQUIRE(distance > 0.0f);                        // the reference's assertion
    REQUIRE(distance == 2.0f);                       // pinned
    REQUIRE(quantizer.similarity(bin1, bin2) == 0.5f);
    BinaryVector other = quantizer.quantize({1.0f, 2.0f, 3.0f});
    bool threw = false;
    try { quantizer.hamming_distance(bin1, other); }
    catch (const VectorDbError& e) { threw = e.kind == VectorDbError::InvalidVectorDimension; }
    REQUIRE(threw);                                  // src/quantization.rs:131-133
}

static void test_multi_stage_search() {             // src/quantization.rs:151-193
    BinaryQuantizationConfig config;
    config.rescore_ratio = 0.5f;
    BinaryQuantizer quantizer(config);
    std::vector<std::vector<float>> cands = {{1, 1, 1, 1}, {1, 1, 1, 1}, {1, -1, 1, 1}, {2, 2, 2, 2}, {-1, -1, -1, -1},
                                             {1, 1, -1, -1}, {-1, 1, 1, 1}, {3, 3, 3, 3}};
    std::vector<float> q = {1, 1, 1, 1};
    auto cb = quantizer.quantize_batch(cands);
    auto qb = quantizer.quantize(q);
    auto res = quantizer.multi_stage_search(qb, cb, q, cands);
    REQUIRE(res.size() == 4);                        // (8 as f32 * 0.5) as usize
    // stage 1 keeps the four Hamming-0 rows 0,1,3,7; all have cosine 1.0 -> stage-1 order kept
    REQUIRE(res[0].first == 0 && res[1].first == 1 && res[2].first == 3 && res[3].first == 7);
    REQUIRE(res[0].second == 1.0f);
    cb.pop_back();
    bool threw = false;
    try { quantizer.multi_stage_search(qb, cb, q, cands); }
    catch (const VectorDbError& e) { threw = e.kind == VectorDbError::QuantizationError; }
    REQUIRE(threw);                                  // :158-162
}

static void test_vector_index_trait() {             // trait VectorIndex, src/index.rs:35-62
    for (auto mode : {GpuVectorIndex::Mode::Exact, GpuVectorIndex::Mode::TwoStage}) {
        GpuVectorIndex index(mode, 4);
        bool threw = false;
        try { index.search({1, 0, 0}, 1); } catch (const VectorDbError& e) { threw = e.kind == VectorDbError::IndexNotBuilt; }
        REQUIRE(threw);                              // src/index.rs:621-623
        index.add_vector("doc1", {1.0f, 0.0f, 0.0f});
        index.add_vectors({{"doc2", {0.9f, 0.1f, 0.0f}}, {"doc3", {-1.0f, 0.0f, 0.0f}}, {"doc4", {0.0f, 1.0f, 0.0f}}});
        REQUIRE(index.len() == 4 && !index.is_empty());
        threw = false;
        try { index.add_vector("bad", {1.0f}); } catch (const VectorDbError& e) { threw = e.kind == VectorDbError::DimensionMismatch && e.expected == 3 && e.actual == 1; }
        REQUIRE(threw);                              // src/index.rs:590-594
        auto res = index.search({1.0f, 0.0f, 0.0f}, 2);
        REQUIRE(res.size() == 2 && res[0].first == "doc1" && res[1].first == "doc2");
        REQUIRE(res[0].second == 0.0f);              // distance = 1 - cos (src/index.rs:699)
        REQUIRE(index.remove_vector("doc1") && !index.remove_vector("doc1"));
        res = index.search({1.0f, 0.0f, 0.0f}, 1);
        REQUIRE(res.size() == 1 && res[0].first == "doc2");
        IndexStats st = index.get_stats();
        REQUIRE(st.vector_count == 3 && st.dimension == 3);
        index.clear();
        REQUIRE(index.is_empty());
    }
}

static void test_filtered_search() {                // FilterEngine::execute_filter ids as a pre-filter
    GpuVectorIndex exact(GpuVectorIndex::Mode::Exact, 4), two_stage(GpuVectorIndex::Mode::TwoStage, 4);
    for (GpuVectorIndex* idx : {&exact, &two_stage}) {
        idx->add_vector("a", {1.0f, 0.0f, 0.0f, 0.0f});
        idx->add_vector("b", {0.9f, 0.1f, 0.0f, 0.0f});
        idx->add_vector("c", {0.0f, 1.0f, 0.0f, 0.0f});
        idx->add_vector("d", {0.7f, 0.7f, 0.0f, 0.0f});
        auto all = idx->search({1.0f, 0.0f, 0.0f, 0.0f}, 2);
        REQUIRE(all.size() == 2 && all[0].first == "a" && all[1].first == "b");
        auto some = idx->search_filtered({1.0f, 0.0f, 0.0f, 0.0f}, 2, {"c", "d", "nobody"});
        REQUIRE(some.size() == 2 && some[0].first == "d" && some[1].first == "c");
        REQUIRE(idx->search_filtered({1.0f, 0.0f, 0.0f, 0.0f}, 2, {}).empty());
        REQUIRE(idx->remove_vector("d"));
        auto left = idx->search_filtered({1.0f, 0.0f, 0.0f, 0.0f}, 2, {"c", "d"});
        REQUIRE(left.size() == 1 && left[0].first == "c");
    }
}

static void test_rrf_fusion() {                     // src/hybrid.rs:991-1025
    std::vector<std::pair<std::string, float>> dense = {{"doc1", 0.9f}, {"doc2", 0.8f}, {"doc3", 0.7f}};
    std::vector<std::pair<std::string, float>> sparse = {{"doc2", 0.95f}, {"doc1", 0.85f}, {"doc4", 0.75f}};
    auto result = rrf_fusion(dense, sparse, {}, 60.0f);
    REQUIRE(!result.empty());
    float doc1 = 0, doc3 = 0;
    for (auto& r : result) { if (r.id == "doc1") doc1 = r.score; if (r.id == "doc3") doc3 = r.score; }
    REQUIRE(doc1 > doc3);                            // the reference's assertion (:1022-1024)
    REQUIRE(doc1 == 1.0f / 61.0f + 1.0f / 62.0f);    // pinned
    REQUIRE(doc3 == 1.0f / 63.0f);
    REQUIRE(result.size() == 4 && result[0].id == "doc1" && result[1].id == "doc2");
}

static void test_weighted_fusions() {               // src/hybrid.rs:491-616, through the GPU library
    std::vector<std::pair<std::string, float>> dense = {{"doc1", 0.9f}, {"doc2", 0.5f}};
    std::vector<std::pair<std::string, float>> sparse = {{"doc2", 4.0f}, {"doc3", 1.0f}};
    auto lin = linear_fusion(dense, sparse, {}, 0.7f, 0.2f, 0.1f);
    REQUIRE(lin.size() == 3 && lin[0].id == "doc2" && lin[1].id == "doc1" && lin[2].id == "doc3");
    REQUIRE(lin[0].score == 0.5f * 0.7f + 4.0f * 0.2f && lin[1].score == 0.9f * 0.7f && lin[2].score == 1.0f * 0.2f);
    REQUIRE(lin[0].breakdown.has_dense && lin[0].breakdown.has_sparse && lin[0].breakdown.sparse_score == 4.0f);
    auto nrm = normalized_fusion(dense, sparse, {{"doc7", 123.0f}}, 0.7f, 0.2f, 0.1f);
    REQUIRE(nrm.size() == 4 && nrm[0].id == "doc1" && nrm[1].id == "doc2" && nrm[2].id == "doc7" && nrm[3].id == "doc3");
    REQUIRE(nrm[0].score == 0.7f && nrm[1].score == 0.0f + 0.2f && nrm[2].score == 0.1f && nrm[3].score == 0.0f);
    REQUIRE(nrm[2].breakdown.has_text && nrm[2].breakdown.text_score == 1.0f);
}

static void test_sparse_bm25() {                    // src/sparse.rs:153-222
    SparseIndex idx;
    idx.add_document({"a", {{0, 1.0f}}, 1.0f});
    idx.add_document({"b", {{0, 2.0f}}, 2.0f});
    idx.add_document({"c", {{1, 1.0f}}, 1.0f});
    REQUIRE(idx.total_documents() == 3);
    REQUIRE(idx.average_document_length() == 4.0f / 3.0f);
    SparseVector q{{0}, {1.0f}, 10};
    auto r = idx.search_bm25(q, 10);
    REQUIRE(r.size() == 2);
    const float idf = std::log((3.0f - 2.0f + 0.5f) / (2.0f + 0.5f));   // negative: df > N/2
    REQUIRE(idf < 0.0f && r[0].second >= r[1].second);
}

static void test_gpu_sparse_bm25() {                // same documents on the GPU postings index
    SparseIndex host;
    GpuSparseIndex gpu;
    const DocumentSparseRepresentation docs[] = {
        {"a", {{0, 1.0f}}, 1.0f}, {"b", {{0, 2.0f}, {3, 0.5f}}, 2.5f}, {"c", {{1, 1.0f}}, 1.0f},
        {"d", {{3, 0.25f}, {1, 0.75f}}, 1.0f}, {"e", {{2, 1.0f}}, 1.0f}};
    for (auto& d : docs) { host.add_document(d); gpu.add_document(d); }
    REQUIRE(gpu.total_documents() == 5);
    // postings are summed in term order on the GPU side, in insertion order on the host side
    REQUIRE(std::fabs(gpu.average_document_length() - host.average_document_length()) < 1e-6f);
    for (const SparseVector& q : {SparseVector{{0}, {1.0f}, 10}, SparseVector{{3, 1}, {0.5f, 0.5f}, 10},
                                  SparseVector{{9}, {1.0f}, 10}, SparseVector{{1, 1, 2}, {0.25f, 0.5f, 0.25f}, 10}}) {
        auto h = host.search_bm25(q, 3), g = gpu.search_bm25(q, 3);
        REQUIRE(h.size() == g.size());
        for (size_t i = 0; i < h.size(); ++i) REQUIRE(std::fabs(h[i].second - g[i].second) <= 1e-6f * std::fabs(h[i].second));
    }
    auto batch = gpu.search_bm25_batch({SparseVector{{0}, {1.0f}, 10}, SparseVector{{2}, {1.0f}, 10}}, 4);
    REQUIRE(batch.size() == 2 && batch[0].size() == 2 && batch[1].size() == 1 && batch[1][0].first == "e");
}

static void test_hybrid_search() {                 // src/hybrid.rs:286-356 on the GPU dense index
    auto dense = std::make_shared<GpuVectorIndex>(GpuVectorIndex::Mode::Exact, 4);
    std::shared_ptr<SparseIndex> sparse = std::make_shared<GpuSparseIndex>();     // BM25 list from the GPU too
    HybridSearchEngine engine(dense, sparse, 60.0f);
    // dense order for the query (1,0,0): doc1 (cos 1), doc2, doc3, doc4
    DocumentSparseRepresentation s1{"doc1", {{7, 1.0f}}, 1.0f}, s3{"doc3", {{7, 3.0f}}, 3.0f};
    engine.add_document("doc1", {1.0f, 0.0f, 0.0f}, &s1);
    engine.add_document("doc2", {0.9f, 0.2f, 0.0f}, nullptr);
    engine.add_document("doc3", {0.5f, 0.5f, 0.0f}, &s3);
    engine.add_document("doc4", {0.0f, 1.0f, 0.0f}, nullptr);
    HybridSearchRequest req;
    req.dense_vector = {1.0f, 0.0f, 0.0f};
    req.sparse_vector = SparseVector{{7}, {1.0f}, 10};
    req.limit = 3;
    auto dense_only = dense->search(req.dense_vector, 6);
    REQUIRE(dense_only.size() == 4 && dense_only[0].first == "doc1" && dense_only[1].first == "doc2" &&
            dense_only[2].first == "doc3" && dense_only[3].first == "doc4");
    auto bm25 = sparse->search_bm25(req.sparse_vector, 6);
    REQUIRE(bm25.size() == 2);
    // idf = ln((2 - 2 + 0.5) / (2 + 0.5)) < 0: the higher tf scores LOWER, so doc1 leads the sparse list
    REQUIRE(bm25[0].first == "doc1" && bm25[1].first == "doc3");
    auto res = engine.search(req);
    REQUIRE(res.size() == 3);
    REQUIRE(res[0].id == "doc1" && res[0].score == 1.0f / 61.0f + 1.0f / 61.0f);
    REQUIRE(res[1].id == "doc3" && res[1].score == 1.0f / 63.0f + 1.0f / 62.0f);
    REQUIRE(res[2].id == "doc2" && res[2].score == 1.0f / 62.0f);
}

int main() {
    test_binary_quantization();
    test_hamming_distance();
    test_multi_stage_search();
    test_vector_index_trait();
    test_filtered_search();
    test_rrf_fusion();
    test_weighted_fusions();
    test_sparse_bm25();
    test_gpu_sparse_bm25();
    test_hybrid_search();
    std::printf("host mirror tests: all passed\n");
    return 0;
}
