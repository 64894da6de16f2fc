// tc_probe.cu — standalone check + timing of the tcgen05 Hamming scan (gvdb_tc.cuh) against a
// plain popc kernel.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo
//   -I grape-vector-db_b200/csrc -o gpurun_out/tc_probe tools/tc_probe.cu     Run on a B200.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "gvdb_tc.cuh"
using namespace gvdb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

template <int NCHUNK>
__global__ void ref_kernel(const uint4* codes, uint32_t ntiles, const uint32_t* qpack, int qs, uint32_t nq, uint32_t* dist, uint64_t stride, uint64_t n_rows) {
    uint32_t tile = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (tile >= ntiles) return;
    uint32_t row = tile * 32 + lane;
    uint4 r[NCHUNK];
    for (int c = 0; c < NCHUNK; ++c) r[c] = codes[((size_t)tile * NCHUNK + c) * 32 + lane];
    for (uint32_t q = 0; q < nq; ++q) {
        uint32_t d = 0;
        for (int c = 0; c < NCHUNK; ++c) {
            const uint32_t* w = qpack + (size_t)q * qs + c * 4;
            d += __popc(r[c].x ^ w[0]) + __popc(r[c].y ^ w[1]) + __popc(r[c].z ^ w[2]) + __popc(r[c].w ^ w[3]);
        }
        if (row < n_rows) dist[(size_t)q * stride + row] = d;
    }
}

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 16); }

int main(int argc, char** argv) {
#ifndef PROBE_NCHUNK
#define PROBE_NCHUNK 6
#endif
    constexpr int NCHUNK = PROBE_NCHUNK;
    constexpr int TC_QBLOCKS = tc_qblocks(NCHUNK);
    const int qs = NCHUNK * 4 + 4;
    uint64_t n_rows = argc > 1 ? atoll(argv[1]) : 148 * 128 * 2 + 77;
    uint32_t nq = argc > 2 ? atoi(argv[2]) : 300;
    int timing = argc > 3 ? atoi(argv[3]) : 0;
    uint32_t ntiles = (uint32_t)((n_rows + 31) / 32);
    uint32_t nq_pad = (nq + TC_NQ - 1) / TC_NQ * TC_NQ;
    size_t code_words = (size_t)ntiles * NCHUNK * 32 * 4;
    std::vector<uint32_t> h_codes(code_words), h_qpack((size_t)nq * qs), h_live(ntiles, 0xffffffffu);
    for (auto& w : h_codes) w = rnd();
    if (n_rows % 32) h_live[ntiles - 1] = (1u << (n_rows % 32)) - 1u;   // rows past the end are not live
    for (uint32_t q = 0; q < nq; ++q) {
        for (int w = 0; w < NCHUNK * 4; ++w) h_qpack[(size_t)q * qs + w] = rnd();
        h_qpack[(size_t)q * qs + NCHUNK * 4] = TAU_ALL;
        for (int w = 1; w < 4; ++w) h_qpack[(size_t)q * qs + NCHUNK * 4 + w] = 0;
    }
    uint4* d_codes; uint32_t *d_qpack, *d_live, *d_ref, *d_tc, *d_qpop, *d_cnt, *d_flag; int8_t* d_qexp; int32_t* d_qbase; uint64_t* d_buf; int32_t* d_tilemin; uint2* d_recs; uint32_t* d_lc;
    CK(cudaMalloc(&d_codes, code_words * 4)); CK(cudaMalloc(&d_qpack, h_qpack.size() * 4)); CK(cudaMalloc(&d_live, ntiles * 4));
    const bool check = !timing;
    size_t dist_bytes = check ? (size_t)nq * n_rows * 4 : 4;
    CK(cudaMalloc(&d_ref, dist_bytes)); CK(cudaMalloc(&d_tc, dist_bytes));
    CK(cudaMalloc(&d_qexp, (size_t)(nq_pad / TC_NQ) * tc_qblock_bytes(NCHUNK))); CK(cudaMalloc(&d_qbase, nq_pad * 4)); CK(cudaMalloc(&d_qpop, nq_pad * 4));
    const uint32_t cap = 8192;
    const uint32_t rec_cap = 1u << 17;
    CK(cudaMalloc(&d_recs, (size_t)148 * TC_EPI_WARPS * rec_cap * 8)); CK(cudaMalloc(&d_lc, 148 * TC_EPI_WARPS * 4)); CK(cudaMemset(d_lc, 0, 148 * TC_EPI_WARPS * 4));
    CK(cudaMalloc(&d_cnt, nq_pad * 4 * CNT_STRIDE)); CK(cudaMalloc(&d_flag, 256)); CK(cudaMalloc(&d_buf, (size_t)nq_pad * cap * 8));
    CK(cudaMemcpy(d_codes, h_codes.data(), code_words * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_qpack, h_qpack.data(), h_qpack.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_live, h_live.data(), ntiles * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_tc, 0xff, dist_bytes)); CK(cudaMemset(d_cnt, 0, nq_pad * 4 * CNT_STRIDE)); CK(cudaMemset(d_flag, 0, 256));

    tc_expand_queries_kernel<<<nq_pad, 64>>>(d_qpack, qs, NCHUNK, nq, nq_pad, d_qexp, d_qpop);
    CK(cudaGetLastError());
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint32_t ngroups = (ntiles + 3) / 4;
    CK(cudaMalloc(&d_tilemin, (size_t)32 * 128 * nq_pad * 4));
    const uint32_t qb_item = argc > 6 ? (uint32_t)atoi(argv[6]) : (uint32_t)TC_QBLOCKS;
    const uint32_t n_qsl = (nq_pad / TC_NQ + qb_item - 1) / qb_item;
    uint32_t n_rsl = argc > 5 ? atoi(argv[5]) : 0;
    if (!n_rsl) { n_rsl = sms / n_qsl; if (!n_rsl) n_rsl = 1; }
    if (n_rsl > ngroups) n_rsl = ngroups;
    uint32_t grid = n_qsl * n_rsl < (uint32_t)sms ? n_qsl * n_rsl : sms;
    size_t smem = TC_QBLOCKS * tc_qblock_bytes(NCHUNK);
    printf("split: %u query slices x %u row slices, grid %u\n", n_qsl, n_rsl, grid);
    if (check) {
        CK(cudaFuncSetAttribute(tc_scan_kernel<NCHUNK, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_bias_kernel<<<(nq_pad + 7) / 8, 256>>>(d_qpack, qs, NCHUNK, d_qpop, nq, nq_pad, d_qexp, d_qbase, 1);
        ref_kernel<NCHUNK><<<(ntiles + 7) / 8, 256>>>(d_codes, ntiles, d_qpack, qs, nq, d_ref, n_rows, n_rows);
        CK(cudaGetLastError());
        tc_scan_kernel<NCHUNK, 1><<<grid, TC_THREADS, smem>>>(d_codes, d_live, 0, ntiles, ngroups, 1, d_qexp, d_qbase, nq, nq_pad, n_qsl, n_rsl, qb_item,
                                                             nullptr, 0, nullptr, d_flag, d_tc, n_rows, n_rows, nullptr);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> a((size_t)nq * n_rows), b((size_t)nq * n_rows);
        CK(cudaMemcpy(a.data(), d_ref, dist_bytes, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), d_tc, dist_bytes, cudaMemcpyDeviceToHost));
        size_t bad = 0, first = (size_t)-1;
        for (size_t i = 0; i < a.size(); ++i) if (a[i] != b[i]) { if (!bad) first = i; ++bad; }
        printf("check: rows=%llu nq=%u mismatches=%zu of %zu\n", (unsigned long long)n_rows, nq, bad, a.size());
        if (bad) {
            printf("first mismatch at q=%zu row=%zu ref=%u tc=%u\n", first / n_rows, first % n_rows, a[first], b[first]);
            for (int i = 0; i < 8; ++i) printf("  [%d] ref=%u tc=%u\n", i, a[i], b[i]);
            return 1;
        }
        // ---- sample mode (MODE 2): per-class minima of D = hamming - popc(q); class = (row slice, row % 128) ----
        std::vector<uint32_t> h_pop(nq_pad);
        CK(cudaMemcpy(h_pop.data(), d_qpop, nq_pad * 4, cudaMemcpyDeviceToHost));
        CK(cudaFuncSetAttribute(tc_scan_kernel<NCHUNK, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            const uint32_t nqb = nq_pad / TC_NQ;
            uint32_t rs = (uint32_t)sms / nqb; if (rs < 1) rs = 1; if (rs > 128) rs = 128; if (rs > ngroups) rs = ngroups;
            const uint32_t g2 = nqb * rs < (uint32_t)sms ? nqb * rs : (uint32_t)sms;
            tc_scan_kernel<NCHUNK, 2><<<g2, TC_THREADS, smem>>>(d_codes, d_live, 0, ntiles, ngroups, 1, d_qexp, d_qbase, nq, nq_pad, nqb, rs, 1,
                                                               nullptr, 0, nullptr, d_flag, nullptr, 0, n_rows, d_tilemin);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            std::vector<int32_t> tm((size_t)rs * 32 * nq_pad);
            CK(cudaMemcpy(tm.data(), d_tilemin, tm.size() * 4, cudaMemcpyDeviceToHost));
            size_t badt = 0;
            for (uint32_t sl = 0; sl < rs; ++sl) {
                const uint64_t g_lo = (uint64_t)ngroups * sl / rs, g_hi = (uint64_t)ngroups * (sl + 1) / rs;
                for (uint32_t c = 0; c < 32; ++c)             // class c: rows 4c .. 4c+3 of every group
                    for (uint32_t q = 0; q < nq; ++q) {
                        int32_t want = TC_TILEMIN_NONE;
                        for (uint64_t g = g_lo; g < g_hi; ++g)
                            for (uint32_t k = 0; k < 4; ++k) {
                                const uint64_t r = g * 128 + c * 4 + k;
                                if (r < n_rows) want = std::min<int32_t>(want, (int32_t)a[(size_t)q * n_rows + r] - (int32_t)h_pop[q]);
                            }
                        const int32_t got = tm[((size_t)sl * 32 + c) * nq_pad + q];
                        if (got != want) { if (!badt) printf("class-min mismatch slice=%u class=%u q=%u got=%d want=%d\n", sl, c, q, got, want); ++badt; }
                    }
            }
            printf("sample-mode check: %u classes x %u queries, mismatches=%zu\n", rs * 32, nq, badt);
            if (badt) return 1;
        }
        // ---- search mode (MODE 0): survivors of a finite threshold must be exactly {ham < tau} ----
        const uint32_t tau = NCHUNK * 64 - 32;
        for (uint32_t q = 0; q < nq; ++q) h_qpack[(size_t)q * qs + NCHUNK * 4] = (q % 7 == 3) ? TAU_ALL : tau + (q % 5);
        CK(cudaMemcpy(d_qpack, h_qpack.data(), h_qpack.size() * 4, cudaMemcpyHostToDevice));
        tc_bias_kernel<<<(nq_pad + 7) / 8, 256>>>(d_qpack, qs, NCHUNK, d_qpop, nq, nq_pad, d_qexp, d_qbase, 0);
        CK(cudaFuncSetAttribute(tc_scan_kernel<NCHUNK, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const uint32_t bigcap = 65536;
        uint64_t* d_buf2; CK(cudaMalloc(&d_buf2, (size_t)nq_pad * bigcap * 8));
        CK(cudaMemset(d_cnt, 0, nq_pad * 4 * CNT_STRIDE));
        tc_scan_kernel<NCHUNK, 0><<<grid, TC_THREADS, smem>>>(d_codes, d_live, 0, ntiles, ngroups, 1, d_qexp, d_qbase, nq, nq_pad, n_qsl, n_rsl, qb_item,
                                                             d_recs, rec_cap, d_lc, d_flag, nullptr, 0, n_rows, nullptr);
        tc_scatter_kernel<<<dim3(TC_SCATTER_X, grid * TC_EPI_WARPS), 256>>>(d_recs, rec_cap, d_lc, d_codes, NCHUNK, d_qpack, qs, d_cnt, d_buf2, bigcap, d_flag);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> h_cnt_s((size_t)nq_pad * CNT_STRIDE), h_cnt(nq_pad);
        CK(cudaMemcpy(h_cnt_s.data(), d_cnt, (size_t)nq_pad * 4 * CNT_STRIDE, cudaMemcpyDeviceToHost));
        for (uint32_t q = 0; q < nq_pad; ++q) h_cnt[q] = h_cnt_s[(size_t)q * CNT_STRIDE];
        std::vector<uint64_t> h_buf((size_t)nq_pad * bigcap);
        CK(cudaMemcpy(h_buf.data(), d_buf2, h_buf.size() * 8, cudaMemcpyDeviceToHost));
        size_t badq = 0, total = 0;
        for (uint32_t q = 0; q < nq_pad; ++q) {
            std::vector<uint64_t> want;
            if (q < nq) {
                uint32_t t = h_qpack[(size_t)q * qs + NCHUNK * 4];
                for (uint64_t r = 0; r < n_rows; ++r) { uint32_t d = a[(size_t)q * n_rows + r]; if (t == TAU_ALL || d < t) want.push_back(((uint64_t)d << 32) | r); }
            }
            std::vector<uint64_t> got(h_buf.begin() + (size_t)q * bigcap, h_buf.begin() + (size_t)q * bigcap + std::min<uint32_t>(h_cnt[q], bigcap));
            std::sort(got.begin(), got.end()); std::sort(want.begin(), want.end());
            total += want.size();
            if (q < nq && h_qpack[(size_t)q * qs + NCHUNK * 4] == TAU_ALL && n_rows > bigcap) continue;   // overflow by construction
            if (got != want) { if (!badq) printf("search-mode mismatch q=%u got=%zu want=%zu\n", q, got.size(), want.size()); ++badq; }
        }
        printf("search-mode check: %zu survivors, bad queries=%zu\n", total, badq);
        if (badq) return 1;
    } else {
        // timing in search mode; tau (argv[4]) relative to NCHUNK*64: 768 bits: 200 no survivors, 352 ~1%, 335 ~2e-4
        const uint32_t tau_t = argc > 4 ? atoi(argv[4]) : 200;
        for (uint32_t q = 0; q < nq; ++q) h_qpack[(size_t)q * qs + NCHUNK * 4] = tau_t;
        CK(cudaMemcpy(d_qpack, h_qpack.data(), h_qpack.size() * 4, cudaMemcpyHostToDevice));
        tc_bias_kernel<<<(nq_pad + 7) / 8, 256>>>(d_qpack, qs, NCHUNK, d_qpop, nq, nq_pad, d_qexp, d_qbase, 0);
        CK(cudaFuncSetAttribute(tc_scan_kernel<NCHUNK, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; ++rep)
            tc_scan_kernel<NCHUNK, 0><<<grid, TC_THREADS, smem>>>(d_codes, d_live, 0, ntiles, ngroups, 1, d_qexp, d_qbase, nq, nq_pad, n_qsl, n_rsl, qb_item,
                                                                 d_recs, rec_cap, d_lc, d_flag, nullptr, 0, n_rows, nullptr);
        CK(cudaDeviceSynchronize());
        const int reps = 10;
        unsigned long long* d_prof; CK(cudaMalloc(&d_prof, 16 * 8)); CK(cudaMemset(d_prof, 0, 16 * 8));
        for (int dbg : {0, 4, 1}) {
        // the cycle accounting costs time (clock reads in every warp): one launch with it, the timed ones without
        cudaMemsetAsync(d_cnt, 0, nq_pad * 4 * CNT_STRIDE);
        tc_scan_kernel<NCHUNK, 0><<<grid, TC_THREADS, smem>>>(d_codes, d_live, 0, ntiles, ngroups, 1, d_qexp, d_qbase, nq, nq_pad, n_qsl, n_rsl, qb_item,
                                                             d_recs, rec_cap, d_lc, d_flag, nullptr, 0, n_rows, nullptr, dbg, d_prof);
        cudaEventRecord(e0);
        for (int rep = 0; rep < reps; ++rep) {
            cudaMemsetAsync(d_cnt, 0, nq_pad * 4 * CNT_STRIDE);
            tc_scan_kernel<NCHUNK, 0><<<grid, TC_THREADS, smem>>>(d_codes, d_live, 0, ntiles, ngroups, 1, d_qexp, d_qbase, nq, nq_pad, n_qsl, n_rsl, qb_item,
                                                                 d_recs, rec_cap, d_lc, d_flag, nullptr, 0, n_rows, nullptr, dbg, nullptr);
        }
        cudaEventRecord(e1);
        cudaEvent_t e2; cudaEventCreate(&e2);
        tc_scatter_kernel<<<dim3(TC_SCATTER_X, grid * TC_EPI_WARPS), 256>>>(d_recs, rec_cap, d_lc, d_codes, NCHUNK, d_qpack, qs, d_cnt, d_buf, cap, d_flag);
        cudaEventRecord(e2);
        CK(cudaDeviceSynchronize());
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        float ms_sc = 0; cudaEventElapsedTime(&ms_sc, e1, e2);
        { std::vector<uint32_t> lc(grid * TC_EPI_WARPS); CK(cudaMemcpy(lc.data(), d_lc, lc.size() * 4, cudaMemcpyDeviceToHost)); unsigned long long tot = 0; for (auto c : lc) tot += c;
          printf("  survivors %llu (%.1f per query), scatter kernel %.3f ms\n", tot, (double)tot / nq, ms_sc); }
        double macs = (double)n_rows * nq_pad * NCHUNK * 128;
        unsigned long long hp[16]; CK(cudaMemcpy(hp, d_prof, 16 * 8, cudaMemcpyDeviceToHost));
        const double grp = (double)((ntiles + 3) / 4) / n_rsl * ((n_qsl * n_rsl + grid - 1) / grid);
        printf("  per group (clk): expander: wait a_free %.0f  expand %.0f | epilogue(one warp): wait acc_full %.0f  work %.0f (loads + hand-back %.0f) | mma: wait acc_empty %.0f  a_ready %.0f  b_full(total) %llu | total %.0f clk/group, SM clock %.0f MHz\n",
               hp[0] / grp, hp[2] / grp, hp[4] / grp, hp[5] / grp, hp[6] / grp, hp[8] / grp, hp[9] / grp, hp[11],
               hp[14] / grp, hp[15] ? 1e3 * (double)hp[14] / (double)hp[15] : 0.0);
        printf("timing dbg=%d (1=no-accumulator-reads 2=no-A-stores): rows=%llu nq=%u  %.3f ms/launch  %.1f TMAC/s = %.1f%% of 148 SMs x 16384 MAC/clk @1.965GHz\n",
               dbg, (unsigned long long)n_rows, nq, ms, macs / ms / 1e9, 100.0 * macs / (ms * 1e-3) / (148.0 * 16384 * 1.965e9));
        CK(cudaMemset(d_prof, 0, 16 * 8));
        }
        uint32_t flag = 0, c0 = 0; CK(cudaMemcpy(&flag, d_flag, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&c0, d_cnt, 4, cudaMemcpyDeviceToHost));
        printf("overflow=%u cnt[0]=%u\n", flag, c0);
    }
    printf("OK\n");
    return 0;
}
