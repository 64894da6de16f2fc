// gvdb.cu — C ABI (include/gvdb.h) over the sm_100a kernels in gvdb_kernels.cuh.
//
// One gvdb_index = one shard in one GPU's HBM:
//   rows   f32  [cap][dim]                       the originals (stage-2 operand)
//   codes  uint4[(cap/32)][nchunk][32]           1-bit codes, blocked (see gvdb_kernels.cuh)
//   norms  f32  [cap]                            sequential-fold L2 norms
//   live   u32  [cap/32]                         tombstone bitmap
// Search is a fixed kernel sequence per query tile (no host round trips in between):
//   ingest_kernel<QUERY> -> { scan_kernel -> select_kernel } per row segment
//   -> rescore_kernel -> topk_kernel
// Row segments grow geometrically: the first segment emits every row, select_kernel turns
// what was emitted into the exact R-th smallest distance so far, and later segments only
// emit rows that beat it — the expected emission per segment stays <= cap/4.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <unistd.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost a pointer test unless a tool (nsys, ncu --nvtx) is attached

#include "gvdb.h"
#include "gvdb_flat.cuh"
#include "gvdb_kernels.cuh"
#include "gvdb_tc.cuh"
#include "gvdb_bigr.cuh"
#include "gvdb_sparse.cuh"
#include "gvdb_xchg.cuh"
#include "gvdb_ratio.cuh"

namespace {

using namespace gvdb;

thread_local std::string g_err;

struct Err {
    gvdb_status st;
    std::string msg;
};
[[noreturn]] void fail(gvdb_status st, const std::string& msg) { throw Err{st, msg}; }

#define CU(x)                                                                              \
    do {                                                                                   \
        cudaError_t e__ = (x);                                                             \
        if (e__ != cudaSuccess)                                                            \
            fail(GVDB_ERR_INDEX, std::string(#x) + ": " + cudaGetErrorString(e__));        \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    void ensure(size_t need) {
        if (need <= bytes) return;
        if (p) CU(cudaFree(p));
        p = nullptr; bytes = 0;
        CU(cudaMalloc(&p, need));
        bytes = need;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

enum Kind { K_SCAN = 0, K_SELECT, K_RESCORE, K_TOPK, K_PREP, K_FLAT, K_MERGE, K_TCSCAN, K_SAMPLE, K_SCATTER, K_XCHG, K_XCHG_WAIT, K_TCDOT, K_COUNT };
struct ProfRec {
    int kind;
    cudaEvent_t e0, e1;
    double bytes, pairs;
};

struct Workspace {
    std::vector<cudaEvent_t> ev_pool;   // timing events, reused across calls
    size_t ev_used = 0;
    std::vector<ProfRec> recs;
    cudaEvent_t next_event() {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev_pool.push_back(e);
        }
        return ev_pool[ev_used++];
    }
    cudaStream_t stream = nullptr;   // used by the host-pointer entry points
    cudaEvent_t idle = nullptr;      // recorded after the last kernel that touched the buffers
    bool used = false;
    DevBuf qpack, qnorm, cnt, flag, buf, rec_ham, rec_ids, rec_score;
    DevBuf q_in, ids_out, sc_out, codes_tmp, misc;
    DevBuf qexp, qpop, qbase;        // tcgen05 path: pre-expanded queries (+ bias digits), popc(q), popc(q) + bias
    DevBuf tilemin;                  // tcgen05 path: per-(sample tile, query) minima of the single-pass search
    DevBuf tc_recs, list_counts;     // tcgen05 path: warp-private survivor records
    DevBuf r_q16, r_thr, r_state, r_counts, r_ecnt, r_ekeys, r_esc, r_fb, r_topk_i, r_topk_s, r_list, r_misc;   // ratio mode (gvdb_ratio.cuh)
    uint32_t* h_fb = nullptr;        // pinned: per-query fallback flags of a ratio-mode tile
    size_t h_fb_bytes = 0;
    DevBuf big_keys, big_keys2, big_aux, big_k32, big_v32, big_tmp, big_cut;   // large-R path (gvdb_bigr.cuh)
    uint32_t* h_flag = nullptr;      // pinned
    const uint32_t* live_eff = nullptr;   // per-call row filter ANDed with the tombstone bitmap (filtered searches)
    DevBuf filt, allow_in;
    void* h_res = nullptr;           // pinned staging for the host-pointer entry point's results
    size_t h_res_bytes = 0;
    ~Workspace() {
        for (DevBuf* b : {&qpack, &qnorm, &cnt, &flag, &buf, &rec_ham, &rec_ids, &rec_score, &q_in,
                          &ids_out, &sc_out, &codes_tmp, &misc, &qexp, &qpop, &qbase, &tilemin, &tc_recs, &list_counts, &r_q16, &r_thr, &r_state, &r_counts, &r_ecnt, &r_ekeys,
                          &r_esc, &r_fb, &r_topk_i, &r_topk_s, &r_list, &r_misc,
                          &big_keys, &big_keys2, &big_aux, &big_k32, &big_v32, &big_tmp, &big_cut, &filt, &allow_in}) b->release();
        if (h_flag) cudaFreeHost(h_flag);
        if (h_res) cudaFreeHost(h_res);
        if (h_fb) cudaFreeHost(h_fb);
        for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
        if (idle) cudaEventDestroy(idle);
        if (stream) cudaStreamDestroy(stream);
    }
};

const int kSupportedChunks[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32};

}  // namespace

// Peer exchange state of one rank (gvdb_xchg.cuh has the protocol).
struct Exchange {
    uint32_t world = 0, rank = 0, nq_max = 0, r_max = 0;
    uint64_t rows_per_owner = 0;
    uint8_t* mailbox = nullptr;            // my mailbox (cudaMalloc), mapped by every peer
    size_t mailbox_bytes = 0, set_bytes = 0, q_off = 0, keys_off = 0, sc_off = 0, flags_off = 0;
    std::vector<uint8_t*> peers;           // every rank's mailbox as mapped here (own included)
    uint8_t** peers_dev = nullptr;
    std::vector<void*> opened;             // IPC mappings to close
    bool attached = false;
    uint32_t epoch = 0;
    uint64_t limit_ns = 20ull * 1000 * 1000 * 1000;
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    DevBuf my_keys, err, done;             // done: one block counter per signal kind (128 bytes apart)
    uint32_t* h_err = nullptr;             // pinned copy of err, refreshed at the end of every step
    bool broken = false;                   // a step failed on the host after its epoch began: peers are out of step
    bool dma_queries = true;               // GVDB_XCHG_DMA=0: queries pushed by a kernel instead of the copy engines
    std::mutex mu;                         // steps of one rank are issued one at a time
    ~Exchange() {                          // the device of the owning index is current (gvdb_destroy / create)
        for (void* p : opened) cudaIpcCloseMemHandle(p);
        if (mailbox) cudaFree(mailbox);
        if (peers_dev) cudaFree(peers_dev);
        my_keys.release(); err.release(); done.release();
        if (h_err) cudaFreeHost(h_err);
        if (side) cudaStreamDestroy(side);
        if (fork) cudaEventDestroy(fork);
        if (join) cudaEventDestroy(join);
    }
};

struct gvdb_index {
    gvdb_config cfg{};
    int dim = 0, nchunk = 0, qs = 0, nbytes = 0;
    uint64_t n_rows = 0, n_live = 0, cap_rows = 0;
    float* rows = nullptr;
    uint4* codes = nullptr;
    float* norms = nullptr;
    uint32_t* live = nullptr;
    // ratio mode (gvdb_ratio.cuh): bf16 copy of the rows blocked for the tensor-core pass + 1 / |row|, built on
    // first use and rebuilt after the row count changed
    uint4* rows16 = nullptr;
    float* rinv = nullptr;
    uint64_t rows16_rows = 0, rows16_cap = 0;
    std::mutex rows16_mu;
    bool windowed = false;           // GVDB_FLAG_ROW_WINDOW: f32 rows kept for [win_first, win_first + win_count) only
    uint64_t win_first = 0, win_count = 0;
    // rows[(r - win_first) * dim] is local row r's data; kernels index with the local row number
    const float* rows_base() const { return windowed ? rows - win_first * (size_t)dim : rows; }
    bool rows_cover_all() const { return !windowed || (win_first == 0 && win_count >= n_rows); }
    // peer rows (gvdb_attach_peer_rows_*): owner o's buffer holds rows [o * peer_per, (o+1) * peer_per)
    const float** peer_rows_dev = nullptr;   // device array of n_peers base pointers (own buffer included)
    uint32_t n_peers = 0;
    uint64_t peer_per = 0;
    std::vector<void*> ipc_opened;
    struct Exchange* xchg = nullptr;     // peer exchange (gvdb_exchange_*), see gvdb_xchg.cuh
    bool rows_reachable() const { return rows_cover_all() || (peer_rows_dev && (uint64_t)n_peers * peer_per >= n_rows); }
    int sm_count = 148;
    std::mutex pool_mu;
    std::vector<std::unique_ptr<Workspace>> pool;
    uint32_t query_tile = 1024;
    int scan_ctas_per_sm = 16;
    int scan_variant = -1;     // GVDB_SCAN_NCSA: adder-count override for tuning (768-d only)
    uint32_t tc_min_q = 64;    // GVDB_TC_MIN_Q: query-tile size from which the tcgen05 scan is used
    uint32_t seg0_rows = 4096; // GVDB_SEG0_ROWS: rows of the first ("emit everything") segment
    uint32_t tc_qb_force = 0;  // GVDB_TC_QB: force the query blocks per tensor-core work item (0 = model)
    uint32_t opt_m = 5;        // GVDB_OPT_M: smallest order statistic of the single-pass threshold (0 = single pass off)
    uint32_t* h_async_flag = nullptr;   // pinned: flags of the last gvdb_search_shard_sliced_enqueue_device
    bool async_pending = false;
    bool ratio_tc = true;      // GVDB_RATIO_TC=0: ratio mode always by the cut by counting
    uint32_t sample_div = 16;  // GVDB_SAMPLE_DIV: the single-pass sample is 1/sample_div of the row groups
    uint32_t seg_growth = 16;  // GVDB_SEG_GROWTH: cap on the geometric segment growth (0 = cap/(4R) only)
    std::atomic<int> profile_on{0};
    std::atomic<uint64_t> launches{0};
    std::atomic<uint64_t> optimistic_reruns{0};
    std::atomic<uint64_t> overflow_fallbacks{0};
    std::mutex prof_mu;
    gvdb_profile prof{};
};

namespace {

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        CU(cudaSetDevice(dev));
    }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

// cudaFuncSetAttribute applies to the CURRENT device: one flag bit per device ordinal, so a process that
// drives several GPUs (one thread each) raises the limit on every one of them.
template <class F>
void ensure_dyn_smem(std::atomic<uint64_t>& done, F kernel, int bytes) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return;
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.fetch_or(bit, std::memory_order_release);
}

// Launch with programmatic stream serialization (GVDB_PDL=1; off by default): the kernel may be scheduled while its
// predecessor in the stream drains (it calls pdl_wait() before touching global memory).  Measured on the 1M x 768 step,
// same box, alternating runs: 0.4625 ms with, 0.456-0.460 ms without, and the two-caller end-to-end rate drops from
// 2.37M to 2.13M QPS (CTAs parked at griddepcontrol.wait take SM resources from the other caller's kernels); letting
// the scatter kernel in beside the tensor-core scan costs 0.05 ms.  The launches stay plain.
inline bool pdl_enabled() {
    static const bool on = [] { const char* e = std::getenv("GVDB_PDL"); return e && e[0] == '1'; }();
    return on;
}
template <typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    CU(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}

// Brackets one kernel launch with events (only when profiling is on) and counts it.
// NVTX range names of the stages (tracing), one per Kind
const char* const kKindName[K_COUNT] = {"gvdb:scan(popc)", "gvdb:select", "gvdb:rescore", "gvdb:topk", "gvdb:query_prep", "gvdb:flat_scan",
                                        "gvdb:merge_shards", "gvdb:scan(tcgen05)", "gvdb:sample+thresholds", "gvdb:scatter",
                                        "gvdb:exchange", "gvdb:exchange_wait", "gvdb:ratio_filter(tcgen05 bf16)"};
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct Timed {
    Workspace* ws; cudaStream_t st; bool on; ProfRec rec; NvtxRange nv;
    Timed(gvdb_index* h, Workspace* ws_, cudaStream_t st_, int kind, double bytes = 0, double pairs = 0)
        : ws(ws_), st(st_), on(h->profile_on.load(std::memory_order_relaxed) != 0), nv(kKindName[kind]) {
        h->launches.fetch_add(1, std::memory_order_relaxed);
        if (on) {
            rec = ProfRec{kind, ws->next_event(), ws->next_event(), bytes, pairs};
            cudaEventRecord(rec.e0, st);
        }
    }
    ~Timed() {
        if (on) {
            cudaEventRecord(rec.e1, st);
            ws->recs.push_back(rec);
        }
    }
};

// Call after the stream has been synchronised.
void flush_profile(gvdb_index* h, Workspace* ws) {
    if (ws->recs.empty()) { ws->ev_used = 0; return; }
    std::lock_guard<std::mutex> lk(h->prof_mu);
    for (const ProfRec& r : ws->recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
        switch (r.kind) {
            case K_SCAN: h->prof.scan_ms += ms; h->prof.scan_launches += 1;
                         h->prof.scan_bytes += r.bytes; h->prof.scan_pairs += r.pairs; break;
            case K_SELECT: h->prof.select_ms += ms; break;
            case K_RESCORE: h->prof.rescore_ms += ms; break;
            case K_TOPK: h->prof.topk_ms += ms; break;
            case K_PREP: h->prof.prep_ms += ms; break;
            case K_FLAT: h->prof.flat_ms += ms; break;
            case K_MERGE: h->prof.merge_ms += ms; break;
            case K_TCSCAN: h->prof.tc_ms += ms; h->prof.tc_launches += 1;
                           h->prof.tc_bytes += r.bytes; h->prof.tc_macs += r.pairs; break;
            case K_SAMPLE: h->prof.sample_ms += ms; break;
            case K_SCATTER: h->prof.scatter_ms += ms; break;
            case K_TCDOT: h->prof.dot_ms += ms; h->prof.dot_launches += 1; h->prof.dot_macs += r.pairs; break;
            case K_XCHG: h->prof.exchange_ms += ms; break;
            case K_XCHG_WAIT: h->prof.exchange_wait_ms += ms; break;
        }
    }
    ws->recs.clear();
    ws->ev_used = 0;
}

struct WsLease {
    gvdb_index* h;
    Workspace* ws;
    cudaStream_t stream;   // the stream this lease runs on
    WsLease(gvdb_index* h_, cudaStream_t user_stream, bool use_user) : h(h_), ws(nullptr) {
        {
            std::lock_guard<std::mutex> lk(h->pool_mu);
            if (!h->pool.empty()) {
                ws = h->pool.back().release();
                h->pool.pop_back();
            }
        }
        if (!ws) {
            ws = new Workspace();
            CU(cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&ws->idle, cudaEventDisableTiming));
            CU(cudaMallocHost(&ws->h_flag, 64));
        }
        stream = use_user ? user_stream : ws->stream;
        if (ws->used) CU(cudaStreamWaitEvent(stream, ws->idle, 0));
    }
    ~WsLease() {
        cudaEventRecord(ws->idle, stream);
        ws->used = true;
        std::lock_guard<std::mutex> lk(h->pool_mu);
        h->pool.emplace_back(ws);
    }
};

size_t tiles_for(uint64_t rows) { return (size_t)((rows + 31) / 32); }

// the row bitmap the scans consult: tombstones, ANDed with the call's filter when there is one
inline const uint32_t* live_of(const gvdb_index* h, const Workspace* ws) { return ws->live_eff ? ws->live_eff : h->live; }

void grow(gvdb_index* h, uint64_t need_rows) {
    if (need_rows <= h->cap_rows) return;
    if (need_rows > 0xfffffff0ull) fail(GVDB_ERR_INVALID_ARGUMENT, "at most 2^32-16 rows per shard");   // before anything is allocated
    uint64_t new_cap = std::max<uint64_t>(need_rows, h->cap_rows + h->cap_rows / 2);
    new_cap = std::max<uint64_t>(new_cap, 1024);
    new_cap = (new_cap + 31) / 32 * 32;
    float* rows = nullptr; uint4* codes = nullptr; float* norms = nullptr; uint32_t* live = nullptr;
    // the new buffers are freed again if a later allocation or copy of this call fails
    struct Undo { float*& r; uint4*& c; float*& n; uint32_t*& l; bool armed = true;
                  ~Undo() { if (armed) { if (r) cudaFree(r); if (c) cudaFree(c); if (n) cudaFree(n); if (l) cudaFree(l); } } } undo{rows, codes, norms, live};
    size_t code_bytes = tiles_for(new_cap) * h->nchunk * 32 * sizeof(uint4);
    if (!h->windowed) CU(cudaMalloc(&rows, new_cap * (size_t)h->dim * sizeof(float)));
    CU(cudaMalloc(&codes, code_bytes));
    CU(cudaMalloc(&norms, new_cap * sizeof(float)));
    CU(cudaMalloc(&live, tiles_for(new_cap) * sizeof(uint32_t)));
    CU(cudaMemset(codes, 0, code_bytes));
    CU(cudaMemset(norms, 0, new_cap * sizeof(float)));
    CU(cudaMemset(live, 0, tiles_for(new_cap) * sizeof(uint32_t)));
    if (h->n_rows) {
        if (!h->windowed) CU(cudaMemcpy(rows, h->rows, h->n_rows * (size_t)h->dim * sizeof(float), cudaMemcpyDeviceToDevice));
        CU(cudaMemcpy(codes, h->codes, tiles_for(h->n_rows) * h->nchunk * 32 * sizeof(uint4), cudaMemcpyDeviceToDevice));
        CU(cudaMemcpy(norms, h->norms, h->n_rows * sizeof(float), cudaMemcpyDeviceToDevice));
        CU(cudaMemcpy(live, h->live, tiles_for(h->n_rows) * sizeof(uint32_t), cudaMemcpyDeviceToDevice));
    }
    CU(cudaDeviceSynchronize());
    if (h->rows && !h->windowed) cudaFree(h->rows);
    if (h->codes) cudaFree(h->codes);
    if (h->norms) cudaFree(h->norms);
    if (h->live) cudaFree(h->live);
    if (!h->windowed) h->rows = rows;
    h->codes = codes; h->norms = norms; h->live = live; h->cap_rows = new_cap;
    undo.armed = false;
}

void need_all_rows(const gvdb_index* h) {
    if (!h->rows_cover_all())
        fail(GVDB_ERR_INVALID_ARGUMENT,
             "this entry point needs the f32 data of every row, but the index keeps a row window "
             "(GVDB_FLAG_ROW_WINDOW): use gvdb_stage1_device / gvdb_rescore_keys_device / gvdb_finish_owned_device");
}

// ---- kernel dispatch on NCHUNK (and, for tuning, on the carry-save adder count) ------------
template <int NCHUNK, int MODE, int NCSA, bool AGG>
void launch_scan_t(cudaStream_t st, dim3 grid, size_t smem, const uint4* codes, const uint32_t* live,
                   uint32_t tile_lo, uint32_t tile_hi, const uint32_t* qpack, int nq, int qgroup,
                   uint32_t* cnt, uint64_t* buf, uint32_t cap, uint32_t* overflow, uint32_t* dist_out,
                   uint64_t dist_stride, uint64_t n_rows) {
    static std::atomic<uint64_t> attr_done{0};
    ensure_dyn_smem(attr_done, scan_kernel<NCHUNK, MODE, NCSA, AGG>, 100 * 1024);
    scan_kernel<NCHUNK, MODE, NCSA, AGG><<<grid, SCAN_THREADS, smem, st>>>(
        codes, live, tile_lo, tile_hi, qpack, nq, qgroup, cnt, buf, cap, overflow, dist_out,
        dist_stride, n_rows);
    CU(cudaGetLastError());
}

// default adder count: ~0.44 adders per word balances the XU (popc) and ALU (LOP3) pipes
constexpr int default_ncsa(int nchunk) { return nchunk * 2; }   // W/2, tuned on B200 (profiles/r01_ncsa_tuning.txt)

template <int MODE>
void launch_scan(int nchunk, int variant, cudaStream_t st, dim3 grid, size_t smem, const uint4* codes,
                 const uint32_t* live, uint32_t tile_lo, uint32_t tile_hi, const uint32_t* qpack,
                 int nq, int qgroup, uint32_t* cnt, uint64_t* buf, uint32_t cap, uint32_t* overflow,
                 uint32_t* dist_out, uint64_t dist_stride, uint64_t n_rows) {
#define GVDB_ARGS st, grid, smem, codes, live, tile_lo, tile_hi, qpack, nq, qgroup, cnt, buf, cap, overflow, dist_out, dist_stride, n_rows
#define GVDB_CASE(N) case N: if (agg) launch_scan_t<N, MODE, default_ncsa(N), MODE == 0>(GVDB_ARGS); \
                             else launch_scan_t<N, MODE, default_ncsa(N), false>(GVDB_ARGS); break;
    // many queries per pass: survivors of one tile are appended with one atomic per (tile, query)
    const bool agg = MODE == 0 && nq >= 8;
    if (nchunk == 6 && MODE == 0 && variant >= 0) {   // tuning variants for the 768-d kernel
        switch (variant) {
            case 0: launch_scan_t<6, MODE, 0, false>(GVDB_ARGS); return;
            case 8: launch_scan_t<6, MODE, 8, false>(GVDB_ARGS); return;
            case 9: launch_scan_t<6, MODE, 9, false>(GVDB_ARGS); return;
            case 10: launch_scan_t<6, MODE, 10, false>(GVDB_ARGS); return;
            case 12: launch_scan_t<6, MODE, 12, false>(GVDB_ARGS); return;
            case 13: launch_scan_t<6, MODE, 13, false>(GVDB_ARGS); return;
            case 14: launch_scan_t<6, MODE, 14, false>(GVDB_ARGS); return;
            default: break;
        }
    }
    switch (nchunk) {
        GVDB_CASE(1) GVDB_CASE(2) GVDB_CASE(3) GVDB_CASE(4) GVDB_CASE(6) GVDB_CASE(8)
        GVDB_CASE(12) GVDB_CASE(16) GVDB_CASE(24) GVDB_CASE(32)
        default: fail(GVDB_ERR_INDEX, "unsupported code width");
    }
#undef GVDB_CASE
#undef GVDB_ARGS
}

// ---- tcgen05 scan dispatch (codes up to 1536 bits: the resident query block must fit shared memory) ---
bool tc_supported(int nchunk) { return tc_supported_chunks(nchunk); }

// Work split of one tcgen05 scan launch: items = query slices x row slices, one CTA per SM looping
// over items.  Pick the row-slice count that fills whole waves of SMs.
struct TcSplit { uint32_t qslices, rslices, grid, qb_item; };
constexpr uint32_t kSampleMaxSlices = TC_TAU_MAX_TILES / 32;    // sample mode: 32 row classes per row slice
TcSplit tc_split(const gvdb_index* h, uint32_t ngroups, uint32_t nq_pad, bool sample = false) {
    const uint32_t sms = (uint32_t)h->sm_count;
    const uint32_t nqb = nq_pad / TC_NQ;
    if (sample) {      // one query block per item (the running minima live in registers), <= 128 row slices
        const uint32_t r = std::max<uint32_t>(1, std::min<uint32_t>(std::min<uint32_t>(ngroups, kSampleMaxSlices), std::max<uint32_t>(1, sms / nqb)));
        return TcSplit{nqb, r, (uint32_t)std::min<uint64_t>((uint64_t)nqb * r, sms), 1};
    }
    uint32_t qb_max = (uint32_t)tc_qblocks(h->nchunk);
    uint32_t qb_min = 1;
    if (h->tc_qb_force) qb_min = qb_max = std::min<uint32_t>(qb_max, h->tc_qb_force);   // GVDB_TC_QB (tuning)
    TcSplit best{1, 1, 1, 1};
    double best_cost = 1e300;
    // makespan model: waves x (row groups per item) x (blocks per item + the per-group A expansion share).
    // The share (0.45 of a block's MMAs) is fitted to configs[1] on a B200 (tools/knob_sweep.sh, 1M x 768 x 1024
    // queries): items of 4 / 3 / 2 query blocks take 0.258 / 0.296 / 0.277 ms — with the earlier 0.1 the model
    // preferred 2 blocks x 37 row slices over 4 x 74 by 0.3 % where 4 x 74 is 7 % faster.
    for (uint32_t qb = qb_min; qb <= qb_max; ++qb) {
        const uint32_t qsl = (nqb + qb - 1) / qb;
        const uint32_t rmax = std::max<uint32_t>(1, std::min<uint32_t>(ngroups, std::max<uint32_t>(1, 8 * sms / qsl)));
        for (uint32_t r = 1; r <= rmax; ++r) {
            const uint64_t items = (uint64_t)qsl * r;
            const uint64_t waves = (items + sms - 1) / sms;
            const double cost = (double)waves * (std::ceil((double)ngroups / r) * (qb + 0.45) + 6.0 * qb);
            if (cost < best_cost - 1e-9) {
                best_cost = cost;
                best = TcSplit{qsl, r, (uint32_t)std::min<uint64_t>(items, sms), qb};
            }
        }
    }
    return best;
}

// MODE 0: survivors (hamming < tau, qthr holds the per-query constants) -> per-query candidate buffers.
// MODE 1: every distance to dist_out (parity).
// MODE 2: per-(tile, query) minima of the strided sample -> ws->tilemin.
// Row groups: `ngroups` groups of 128 rows starting at tile_lo, `group_stride` groups apart.
// expect_per_query (MODE 0): survivors the caller expects per query, sizes the record lists.
template <int MODE>
void launch_tc_scan(gvdb_index* h, Workspace* ws, cudaStream_t st, uint32_t tile_lo, uint32_t tile_hi,
                    uint32_t ngroups, uint32_t group_stride, uint32_t nq, uint32_t nq_pad, uint32_t* cnt,
                    uint64_t* buf, uint32_t cap, uint32_t* overflow, uint32_t* dist_out, uint64_t dist_stride,
                    uint32_t expect_per_query = 0) {
    const TcSplit sp = tc_split(h, ngroups, nq_pad, MODE == 2 || MODE == 3);   // one query block per item: per-lane state
    const uint32_t grid = sp.grid;
    const uint32_t nlists = grid * TC_EPI_WARPS;
    const size_t smem = (size_t)tc_qblocks(h->nchunk) * tc_qblock_bytes(h->nchunk);
    uint32_t rec_cap = 0;
    if (MODE == 0) {
        // 4x head-room over the expected survivors per list (a list that fills up raises the overflow flag)
        const uint64_t want = 4ull * ((uint64_t)nq * std::max<uint32_t>(expect_per_query, 64)) / nlists;
        rec_cap = 2048;
        while (rec_cap < want && rec_cap < (1u << 20)) rec_cap <<= 1;
        ws->tc_recs.ensure((size_t)nlists * rec_cap * sizeof(uint2));
        ws->list_counts.ensure((size_t)h->sm_count * TC_EPI_WARPS * 4);
    }
    const int8_t* qexp = ws->qexp.as<int8_t>();
    const int32_t* qbase = ws->qbase.as<int32_t>();
    int32_t* tilemin = ws->tilemin.as<int32_t>();
    uint2* recs = ws->tc_recs.as<uint2>();
    uint32_t* lc = ws->list_counts.as<uint32_t>();
    const double rows = (double)ngroups * TC_ROWS;
    {
    // algorithmic work of the launch: rows x padded queries x code bits (one MAC per code bit and pair)
    Timed t(h, ws, st, MODE == 2 ? K_SAMPLE : K_TCSCAN, rows * h->nchunk * 16.0 * sp.qslices,
            rows * (double)nq_pad * (h->nchunk * 128.0));
#define GVDB_TC_CASE(N)                                                                              \
    case N: {                                                                                        \
        static std::atomic<uint64_t> attr_done{0};                                                   \
        ensure_dyn_smem(attr_done, tc_scan_kernel<N, MODE>, (int)(tc_qblocks(N) * tc_qblock_bytes(N))); \
        launch_pdl(tc_scan_kernel<N, MODE>, dim3(grid), dim3(TC_THREADS), smem, st, h->codes, live_of(h, ws), tile_lo, tile_hi, ngroups, group_stride, \
                   qexp, qbase, nq, nq_pad, sp.qslices, sp.rslices, sp.qb_item,        \
                   recs, rec_cap, lc, overflow, dist_out, dist_stride, h->n_rows, tilemin, 0, (unsigned long long*)nullptr); \
        break;                                                                                       \
    }
    switch (h->nchunk) {
        GVDB_TC_CASE(1) GVDB_TC_CASE(2) GVDB_TC_CASE(3) GVDB_TC_CASE(4) GVDB_TC_CASE(6) GVDB_TC_CASE(8) GVDB_TC_CASE(12)
        GVDB_TC_CASE(16) GVDB_TC_CASE(24)
        default: fail(GVDB_ERR_INDEX, "tcgen05 scan: unsupported code width");
    }
#undef GVDB_TC_CASE
    }
    CU(cudaGetLastError());
    if (MODE == 0) {
        Timed t(h, ws, st, K_SCATTER);
        // CTAs per list: enough threads for the survivors a list is expected to hold (a second CTA on a list of
        // ~170 records only adds a wave of blocks that find nothing to do)
        const uint64_t per_list = ((uint64_t)nq * std::max<uint32_t>(expect_per_query, 1) + nlists - 1) / nlists;
        const unsigned sx = (unsigned)std::min<uint64_t>(8, std::max<uint64_t>(1, (per_list + 255) / 256));
        launch_pdl(tc_scatter_kernel, dim3(sx, nlists), dim3(256), 0, st,
                   recs, rec_cap, lc, h->codes, h->nchunk, ws->qpack.as<uint32_t>(), h->qs, cnt, buf, cap, overflow);
        CU(cudaGetLastError());
    }
}

void tc_ensure_query_buffers(gvdb_index* h, Workspace* ws, uint32_t nq_pad) {
    ws->qexp.ensure((size_t)(nq_pad / TC_NQ) * tc_qblock_bytes(h->nchunk));
    ws->qpop.ensure((size_t)nq_pad * 4);
    ws->qbase.ensure((size_t)nq_pad * 4);
}

// qpack -> expanded query blocks + popcounts (the multi-segment schedule and gvdb_hamming; the single-pass
// search does this inside query_prep_tc_kernel)
void tc_prepare_queries(gvdb_index* h, Workspace* ws, cudaStream_t st, uint32_t nq, uint32_t nq_pad) {
    tc_ensure_query_buffers(h, ws, nq_pad);
    h->launches.fetch_add(1, std::memory_order_relaxed);
    tc_expand_queries_kernel<<<nq_pad, 64, 0, st>>>(ws->qpack.as<uint32_t>(), h->qs, h->nchunk, nq, nq_pad,
                                                    ws->qexp.as<int8_t>(), ws->qpop.as<uint32_t>());
    CU(cudaGetLastError());
}

// the current tau words of qpack -> the bias digits of the resident blocks + popc(q) + bias (zero_bias: MODE 1)
void tc_update_bias(gvdb_index* h, Workspace* ws, cudaStream_t st, uint32_t nq, uint32_t nq_pad, int zero_bias) {
    h->launches.fetch_add(1, std::memory_order_relaxed);
    tc_bias_kernel<<<(nq_pad + 7) / 8, 256, 0, st>>>(ws->qpack.as<uint32_t>(), h->qs, h->nchunk, ws->qpop.as<uint32_t>(),
                                                     nq, nq_pad, ws->qexp.as<int8_t>(), ws->qbase.as<int32_t>(), zero_bias);
    CU(cudaGetLastError());
}

constexpr int kQGroup = 128;   // queries staged per CTA (at most)

// Queries staged per CTA for a scan over `ntiles` tiles: the full group when the row range alone
// fills the GPU; for short ranges (the first segment) smaller groups, so that the launch still
// spreads over every SM (the per-CTA work is a serial walk over its query group).
int pick_qgroup(const gvdb_index* h, uint32_t ntiles, uint32_t nq) {
    const uint32_t warps = SCAN_THREADS / 32;
    const uint32_t x_full = (ntiles + warps - 1) / warps;
    int g = kQGroup;
    while (g > 8 && (uint64_t)x_full * ((nq + g - 1) / g) < 4ull * h->sm_count) g >>= 1;
    return g;
}

dim3 scan_grid(const gvdb_index* h, uint32_t ntiles, uint32_t nq, int qgroup = kQGroup) {
    const uint32_t warps = SCAN_THREADS / 32;
    uint32_t y = (nq + qgroup - 1) / qgroup;
    uint32_t x_full = (ntiles + warps - 1) / warps;
    uint32_t x = x_full;
    if (y > 2) x = std::min<uint32_t>(x_full, std::max<uint32_t>(1, (uint32_t)(h->sm_count * h->scan_ctas_per_sm) / y));
    x = std::max<uint32_t>(x, 1);
    return dim3(x, y, 1);
}

uint32_t next_pow2_host(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
uint32_t pick_cap(uint32_t R) { return R <= 1024 ? 8192u : 16384u; }
constexpr uint32_t kMaxR = SORT_N / 2;

// Plan of the single-pass search (tensor-core scan): a strided SAMPLE of the row groups gives, per query,
// the minima of its 32-row tiles; the m-th smallest of them + 1 is the threshold of ONE pass over all rows
// (tc_tau_kernel).  Expected survivors per query: m * rows / sampled rows; the pass found the true top R iff
// at least R candidates came out (checked on the device, the call is rerun with the segment schedule otherwise).
//   sample = every row (small corpora): m = R and the guarantee is deterministic — the m-th smallest tile
//            minimum is at least the m-th smallest distance, so at least R rows lie at or below it.
//   else:    the pass misses iff m of the true top R rows fall into the sample, probability
//            P(Poisson(R * sampled fraction) >= m); the smallest m >= GVDB_OPT_M keeping that under 1e-6.
struct SinglePass { bool ok; uint32_t n_sgroups, gstride, m; };
SinglePass plan_single_pass(const gvdb_index* h, uint32_t ntiles, uint32_t R, uint32_t cap) {
    SinglePass sp{false, 0, 1, 0};
    if (h->opt_m == 0) return sp;
    const uint32_t ngroups = (ntiles + 3) / 4;
    uint32_t n_s = std::min<uint32_t>(1024, std::max<uint32_t>(64, ngroups / h->sample_div));
    if (n_s >= ngroups) {                                     // the sample is the corpus
        // the rows fall into at most min(ngroups, 128) x 32 classes: want R of them to be plenty
        if (4ull * R > 32ull * std::min<uint32_t>(ngroups, kSampleMaxSlices)) return sp;
        sp = SinglePass{true, ngroups, 1, R};
        return sp;
    }
    const uint32_t gstride = ngroups / n_s;
    const double x = (double)R * n_s / (double)ngroups;
    double term = std::exp(-x), below = 0.0;                  // below = P(Poisson(x) < m)
    uint32_t m = 0;
    for (; m < 1024; ++m) {
        if (m >= h->opt_m && 1.0 - below <= 1e-6) break;
        below += term;
        term *= x / (double)(m + 1);
    }
    if (m >= 1024 || m > 128) return sp;                      // m classes out of several hundred: collisions stay rare
    if ((double)m * ngroups / n_s > cap / 3.0) return sp;     // expected survivors per query
    sp = SinglePass{true, n_s, gstride, m};
    return sp;
}

// Final ordering fused into the rescoring kernel (rescore_topk_kernel): the caller's k-lists are written
// by search_core itself and the record arrays only when asked for.  `done` reports whether the fused
// kernel ran (it needs dim % 4 == 0 and a rescore count whose staging fits shared memory).
struct FusedTopk {
    uint32_t k = 0;
    uint64_t* ids_out = nullptr;
    float* scores_out = nullptr;
    bool want_records = true;
    uint32_t slice_q = 0;        // > 0: records packed per slice of slice_q queries at rec_ids (gvdb_search_shard_sliced_device)
    bool done = false;
};
bool rescore_topk_fits(const gvdb_index* h, uint32_t R);
size_t rescore_topk_smem(int dim, uint32_t R) {
    uint32_t n_eff = 64;
    while (n_eff < R) n_eff <<= 1;
    return ((size_t)dim + (size_t)R * 2 * RT_STRIDE) * sizeof(float) + (size_t)n_eff * 8;   // two slabs of R staged rows
}

bool rescore_topk_fits(const gvdb_index* h, uint32_t R) {
    return (h->dim & 3) == 0 && R <= (uint32_t)RT_MAX_R && rescore_topk_smem(h->dim, R) <= 200 * 1024;
}

// Stage 1 + stage 2 for queries [0,nq) (device pointers): fills rec_* [nq][R].
void search_core(gvdb_index* h, Workspace* ws, cudaStream_t st, const float* queries_dev,
                 uint32_t nq, uint32_t R, uint32_t* rec_ham, uint64_t* rec_ids, float* rec_score,
                 bool reset_overflow_flag = true, bool allow_optimistic = true, bool* used_optimistic = nullptr,
                 uint64_t* keys_out = nullptr /* stage 1 only: hamming << 40 | global row, nq x R */,
                 FusedTopk* fused = nullptr) {
    if (!keys_out && !h->rows_reachable()) need_all_rows(h);
    if (h->n_rows == 0) fail(GVDB_ERR_INDEX_NOT_BUILT, "index not built: search before any add");
    if (R == 0) fail(GVDB_ERR_INVALID_ARGUMENT, "rescore_count must be >= 1");
    if (R > kMaxR)
        fail(GVDB_ERR_NOT_IMPLEMENTED, "rescore_count > 2048 is not implemented for the sharded entry points");
    const uint32_t cap = pick_cap(R);
    const uint32_t ntiles = (uint32_t)tiles_for(h->n_rows);
    const uint32_t QT = h->query_tile;
    const uint32_t qt_max = std::min(QT, nq);
    ws->qpack.ensure((size_t)qt_max * h->qs * 4);
    ws->qnorm.ensure((size_t)qt_max * 4);
    ws->cnt.ensure((size_t)qt_max * 4 * CNT_STRIDE);
    ws->flag.ensure(256);
    ws->buf.ensure((size_t)qt_max * cap * 8);
    if (reset_overflow_flag) CU(cudaMemsetAsync(ws->flag.p, 0, 256, st));
    // geometric segments; growth keeps expected emission R*(g-1) <= cap/4
    const uint32_t seg0_tiles = h->seg0_rows / 32;
    uint32_t g = std::max<uint32_t>(2, cap / (4 * R));
    // the cap trades survivors per segment (epilogue + select work) for launches: only worth it
    // where the tensor-core scan runs; few-query passes are HBM-bound and want few launches
    if (h->seg_growth >= 2 && std::min(QT, nq) >= h->tc_min_q && tc_supported(h->nchunk)) g = std::min(g, h->seg_growth);
    const uint32_t r_pow2 = std::max<uint32_t>(32, next_pow2_host(R));
    const uint32_t nbins = (uint32_t)h->nchunk * 128 + 1;
    const size_t selh_smem = (size_t)r_pow2 * 8 + (size_t)SORT_N * 8 + (size_t)nbins * 4;
    {
        static std::atomic<uint64_t> attr_done{0};
        ensure_dyn_smem(attr_done, select_hist_kernel, 72 * 1024);
    }
    for (uint32_t qt0 = 0; qt0 < nq; qt0 += QT) {
        const uint32_t nqt = std::min(QT, nq - qt0);
        // Large query tiles: the scan is a dense contraction -> tcgen05 (gvdb_tc.cuh).
        const bool tc_ok = nqt >= h->tc_min_q && tc_supported(h->nchunk);
        const uint32_t nq_pad = (nqt + TC_NQ - 1) / TC_NQ * TC_NQ;
        SinglePass sp{false, 0, 1, 0};
        if (allow_optimistic && tc_ok && (h->dim & 3) == 0) sp = plan_single_pass(h, ntiles, R, cap);
        if (sp.ok) {
            // ---- single pass: prep -> sample (tile minima) -> thresholds -> one scan over all rows -> cut ----
            if (used_optimistic) *used_optimistic = true;
            tc_ensure_query_buffers(h, ws, nq_pad);
            const uint32_t n_stiles = tc_split(h, sp.n_sgroups, nq_pad, true).rslices * 32;    // row classes of the sample
            ws->tilemin.ensure((size_t)n_stiles * nq_pad * 4);
            {
                Timed t(h, ws, st, K_PREP);
                const size_t smem = (size_t)QPT_WARPS * ((size_t)h->dim * 4 + (size_t)h->nchunk * 16);
                static std::atomic<uint64_t> attr_done{0};
                ensure_dyn_smem(attr_done, query_prep_tc_kernel, QPT_WARPS * (4096 * 4 + 32 * 16));
                query_prep_tc_kernel<<<(nq_pad + QPT_WARPS - 1) / QPT_WARPS, 32 * QPT_WARPS, smem, st>>>(
                    queries_dev + (size_t)qt0 * h->dim, nqt, nq_pad, h->dim, h->cfg.threshold, h->nchunk,
                    ws->qnorm.as<float>(), ws->qpack.as<uint32_t>(), h->qs, ws->qexp.as<int8_t>(), ws->qpop.as<uint32_t>(),
                    ws->cnt.as<uint32_t>(), nullptr);
            }
            CU(cudaGetLastError());
            launch_tc_scan<2>(h, ws, st, 0, ntiles, sp.n_sgroups, sp.gstride, nqt, nq_pad, nullptr, nullptr, 0,
                              ws->flag.as<uint32_t>(), nullptr, 0);
            {
                Timed t(h, ws, st, K_SAMPLE);
                static std::atomic<uint64_t> attr_done{0};
                ensure_dyn_smem(attr_done, tc_tau_kernel, TC_TAU_WARPS * (TC_TAU_MAX_TILES + 2) * 2);
                launch_pdl(tc_tau_kernel, dim3((nq_pad + TC_TAU_WARPS - 1) / TC_TAU_WARPS), dim3(32 * TC_TAU_WARPS),
                           (size_t)TC_TAU_WARPS * (n_stiles + 2) * 2, st,
                           ws->tilemin.as<int32_t>(), n_stiles, nqt, nq_pad, sp.m, ws->qpop.as<uint32_t>(),
                           ws->qpack.as<uint32_t>(), h->qs, h->nchunk, ws->qexp.as<int8_t>(), ws->qbase.as<int32_t>());
            }
            CU(cudaGetLastError());
            launch_tc_scan<0>(h, ws, st, 0, ntiles, (ntiles + 3) / 4, 1, nqt, nq_pad, ws->cnt.as<uint32_t>(),
                              ws->buf.as<uint64_t>(), cap, ws->flag.as<uint32_t>(), nullptr, 0,
                              (uint32_t)std::min<uint64_t>(cap, (uint64_t)sp.m * ((ntiles + 3) / 4) / sp.n_sgroups + 1));
            {
                Timed t(h, ws, st, K_SELECT);
                // the tie area follows the survivors expected per query (a few hundred at R = 40: small CTAs, one wave)
                uint32_t tie_cap = 1024;
                while (tie_cap < (uint32_t)SORT_N && tie_cap < (uint64_t)sp.m * ((ntiles + 3) / 4) / sp.n_sgroups + 1) tie_cap <<= 1;
                launch_pdl(select_hist_kernel, dim3(nqt), dim3(SELH_THREADS),
                           (size_t)r_pow2 * 8 + (size_t)tie_cap * 8 + (size_t)nbins * 4, st,
                           ws->buf.as<uint64_t>(), cap, ws->cnt.as<uint32_t>(), R, r_pow2, nbins,
                           ws->qpack.as<uint32_t>(), h->qs, h->nchunk * 4, 0u, 2, ws->flag.as<uint32_t>(), tie_cap);
            }
            CU(cudaGetLastError());
        } else {
            // ---- segment schedule: the first segment emits every row, later ones only rows below the exact
            //      R-th smallest distance so far (tcgen05 scan for large tiles, xor + popc otherwise) ----
            {
                Timed t(h, ws, st, K_PREP);
                if ((h->dim & 3) == 0)
                    query_prep_direct_kernel<<<(nqt + 31) / 32, 32, 0, st>>>(
                        queries_dev + (size_t)qt0 * h->dim, nqt, h->dim, h->cfg.threshold, h->nchunk,
                        ws->qnorm.as<float>(), ws->qpack.as<uint32_t>(), h->qs);
                else
                    ingest_kernel<true><<<(nqt + STAGE_ROWS - 1) / STAGE_ROWS, STAGE_ROWS, 0, st>>>(
                        queries_dev + (size_t)qt0 * h->dim, nqt, h->dim, h->cfg.threshold, h->nchunk, 0, nullptr,
                        ws->qnorm.as<float>(), nullptr, ws->qpack.as<uint32_t>(), h->qs);
            }
            CU(cudaGetLastError());
            CU(cudaMemsetAsync(ws->cnt.p, 0, (size_t)nqt * 4 * CNT_STRIDE, st));
            const bool use_tc = tc_ok && ntiles > seg0_tiles;
            if (use_tc) tc_prepare_queries(h, ws, st, nqt, nq_pad);
            uint32_t lo = 0;
            while (lo < ntiles) {
                uint64_t hi64 = lo == 0 ? seg0_tiles : (uint64_t)lo * g;
                uint32_t hi = (uint32_t)std::min<uint64_t>(hi64, ntiles);
                const double seg_rows = (double)(hi - lo) * 32.0;
                if (use_tc && lo > 0) {
                    tc_update_bias(h, ws, st, nqt, nq_pad, 0);
                    launch_tc_scan<0>(h, ws, st, lo, hi, (hi - lo + 3) / 4, 1, nqt, nq_pad, ws->cnt.as<uint32_t>(),
                                      ws->buf.as<uint64_t>(), cap, ws->flag.as<uint32_t>(), nullptr, 0, cap / 4);
                } else {
                    const int qg = pick_qgroup(h, hi - lo, nqt);
                    dim3 grid = scan_grid(h, hi - lo, nqt, qg);
                    Timed t(h, ws, st, K_SCAN, seg_rows * h->nchunk * 16.0 * grid.y, seg_rows * nqt);
                    launch_scan<0>(h->nchunk, h->scan_variant, st, grid, (size_t)qg * h->qs * 4, h->codes, live_of(h, ws), lo, hi,
                                   ws->qpack.as<uint32_t>(), (int)nqt, qg, ws->cnt.as<uint32_t>(),
                                   ws->buf.as<uint64_t>(), cap, ws->flag.as<uint32_t>(), nullptr, 0, h->n_rows);
                }
                {
                    Timed t(h, ws, st, K_SELECT);
                    select_hist_kernel<<<nqt, SELH_THREADS, selh_smem, st>>>(
                        ws->buf.as<uint64_t>(), cap, ws->cnt.as<uint32_t>(), R, r_pow2, nbins,
                        ws->qpack.as<uint32_t>(), h->qs, h->nchunk * 4, 0u, 0, ws->flag.as<uint32_t>(), (uint32_t)SORT_N);
                }
                CU(cudaGetLastError());
                lo = hi;
            }
        }
        const uint64_t pairs = (uint64_t)nqt * R;
        if (keys_out) {
            Timed t(h, ws, st, K_RESCORE);
            emit_keys_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(
                ws->buf.as<uint64_t>(), cap, ws->cnt.as<uint32_t>(), R, nqt, h->cfg.row_base, keys_out + (size_t)qt0 * R);
        } else if (rescore_topk_fits(h, R)) {
            // rescoring (+ the final ordering when the caller handed its k-lists down): one CTA per query
            Timed t(h, ws, st, K_RESCORE);
            static std::atomic<uint64_t> attr_done{0};
            ensure_dyn_smem(attr_done, rescore_topk_kernel, 200 * 1024);
            uint32_t n_eff = 64;
            while (n_eff < R) n_eff <<= 1;
            const bool topk = fused && fused->k > 0;
            const bool recs = !fused || fused->want_records;
            const uint32_t k = topk ? fused->k : 0u;
            const uint32_t slice_q = fused ? fused->slice_q : 0u;
            if (slice_q && (qt0 % slice_q != 0 || QT % slice_q != 0))
                fail(GVDB_ERR_INVALID_ARGUMENT, "queries per slice must divide the query tile");
            // sliced layout: the tile starts at slice qt0 / slice_q of the packed buffer
            uint64_t* ids_base = !recs ? nullptr
                               : slice_q ? reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(rec_ids) + (size_t)(qt0 / slice_q) * slice_q * R * 16)
                                         : rec_ids + (size_t)qt0 * R;
            launch_pdl(rescore_topk_kernel, dim3(nqt), dim3(32 * ((R + 31) / 32)), rescore_topk_smem(h->dim, R), st,
                h->rows, h->norms, h->cfg.row_base, h->dim, queries_dev + (size_t)qt0 * h->dim, ws->qnorm.as<float>(),
                ws->buf.as<uint64_t>(), cap, ws->cnt.as<uint32_t>(), R, n_eff, k,
                topk ? fused->ids_out + (size_t)qt0 * k : nullptr, topk ? fused->scores_out + (size_t)qt0 * k : nullptr,
                recs && !slice_q ? rec_ham + (size_t)qt0 * R : nullptr, ids_base,
                recs && !slice_q ? rec_score + (size_t)qt0 * R : nullptr,
                h->rows_cover_all() ? nullptr : h->peer_rows_dev, h->peer_per, slice_q);
            if (fused) fused->done = true;
        } else {
            Timed t(h, ws, st, K_RESCORE);
            if ((h->dim & 3) == 0) {
                const int cols = std::min(h->dim, RS_SLAB);
                const int stride = ((cols >> 2) & 1) ? cols : cols + 4;      // stride/4 odd: conflict-free LDS.128
                // 32 consecutive pairs touch at most 31/R + 2 consecutive queries
                const int q_slots = (int)std::min<uint32_t>(32, 31 / R + 2);
                const size_t smem = (size_t)(32 + q_slots) * stride * sizeof(float);
                static std::atomic<uint64_t> attr_done{0};
                ensure_dyn_smem(attr_done, rescore_slab_kernel<false>, 64 * (RS_SLAB + 4) * (int)sizeof(float));
                rescore_slab_kernel<false><<<(unsigned)((pairs + 31) / 32), 32, smem, st>>>(
                    h->rows, h->norms, h->cfg.row_base, h->dim, stride, q_slots, queries_dev + (size_t)qt0 * h->dim,
                    ws->qnorm.as<float>(), ws->buf.as<uint64_t>(), cap, ws->cnt.as<uint32_t>(), R, nqt,
                    rec_ham + (size_t)qt0 * R, rec_ids + (size_t)qt0 * R, rec_score + (size_t)qt0 * R, 0, 0,
                    h->rows_cover_all() ? nullptr : h->peer_rows_dev, h->peer_per);
            }
            else
                rescore_kernel<<<(unsigned)((pairs + STAGE_ROWS - 1) / STAGE_ROWS), STAGE_ROWS, 0, st>>>(
                    h->rows, h->norms, h->cfg.row_base, h->dim, queries_dev + (size_t)qt0 * h->dim,
                    ws->qnorm.as<float>(), ws->buf.as<uint64_t>(), cap, ws->cnt.as<uint32_t>(), R, nqt,
                    rec_ham + (size_t)qt0 * R, rec_ids + (size_t)qt0 * R, rec_score + (size_t)qt0 * R);
        }
        CU(cudaGetLastError());
    }
}

// End of a call that has nothing to read back: synchronise only when per-launch timing is on
// (the work is ordered on the caller's stream; the stream-ordered contract of the *_device calls).
void finish_async(gvdb_index* h, Workspace* ws, cudaStream_t st) {
    if (h->profile_on.load(std::memory_order_relaxed) != 0 || !ws->recs.empty()) {
        CU(cudaStreamSynchronize(st));
        flush_profile(h, ws);
    }
}

// Final synchronisation of a search call.  Returns true when the call must be run again without
// the optimistic threshold (the device refuted the guess, or a candidate buffer overflowed under it).
// overflowed (optional out): a candidate buffer overflowed under the plain (geometric) schedule —
// thousands of rows tie below a query's threshold (a heavily duplicated corpus).  Callers that can
// fall back to the cut by counting (search_big_r, no capacity limit) ask for the flag; the others
// report IndexError.
bool check_overflow(gvdb_index* h, Workspace* ws, cudaStream_t st, bool was_optimistic = false,
                    bool* overflowed = nullptr) {
    CU(cudaMemcpyAsync(ws->h_flag, ws->flag.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    flush_profile(h, ws);
    if (overflowed) *overflowed = false;
    if (was_optimistic && (ws->h_flag[0] || ws->h_flag[1])) {
        h->optimistic_reruns.fetch_add(1, std::memory_order_relaxed);
        return true;
    }
    if (ws->h_flag[0]) {
        if (overflowed) { *overflowed = true; return false; }
        fail(GVDB_ERR_INDEX,
             "candidate buffer overflow in the Hamming scan (thousands of rows tie below a query's threshold); "
             "use gvdb_search_batch(_device), which falls back to the cut by counting");
    }
    return false;
}

void launch_topk(gvdb_index* h, Workspace* ws, cudaStream_t st, const uint64_t* rec_ids, const float* rec_score, uint32_t nq,
                 uint32_t R, uint32_t k, uint64_t* ids_out, float* scores_out) {
    uint32_t n_eff = 64;
    while (n_eff < R) n_eff <<= 1;
    uint32_t threads = std::min<uint32_t>(1024, std::max<uint32_t>(32, n_eff / 2));
    {
        Timed t(h, ws, st, K_TOPK);
        topk_kernel<<<nq, threads, n_eff * 8, st>>>(rec_ids, rec_score, R, n_eff, k, ids_out, scores_out);
    }
    CU(cudaGetLastError());
}

// Large rescore counts (R > 2048): the cut by counting of gvdb_bigr.cuh, one query at a time
// (distances are produced for chunks of queries).  Same outputs and ordering as search_device.
void search_big_r(gvdb_index* h, Workspace* ws, cudaStream_t st, const float* q_dev, uint32_t nq, uint32_t k,
                  uint32_t R, uint64_t* ids_out, float* scores_out, uint64_t* cand_ids, uint32_t* cand_ham) {
    need_all_rows(h);
    if (h->n_rows == 0) fail(GVDB_ERR_INDEX_NOT_BUILT, "index not built: search before any add");
    if ((h->dim & 3) != 0) fail(GVDB_ERR_NOT_IMPLEMENTED, "rescore_count > 2048 needs dim % 4 == 0");
    const uint64_t N = h->n_rows;
    const uint32_t ntiles = (uint32_t)tiles_for(N);
    const uint32_t nbins = (uint32_t)h->nchunk * 128 + 1;
    int key_bits = 32;
    while ((1u << (key_bits - 32)) < nbins) ++key_bits;
    const uint32_t QC = std::min<uint32_t>(nq, 16);
    ws->qpack.ensure((size_t)QC * h->qs * 4);
    ws->qnorm.ensure((size_t)QC * 4);
    ws->misc.ensure((size_t)QC * N * 4);
    ws->big_keys.ensure(N * 8);
    ws->big_keys2.ensure(N * 8);
    ws->big_aux.ensure((size_t)nbins * 4 + 64);
    ws->big_k32.ensure((size_t)R * 8);
    ws->big_v32.ensure((size_t)R * 8);
    ws->rec_ham.ensure((size_t)R * 4);
    ws->rec_ids.ensure((size_t)R * 8);
    ws->rec_score.ensure((size_t)R * 4);
    uint32_t* hist = ws->big_aux.as<uint32_t>();
    BigRCut* cut = reinterpret_cast<BigRCut*>(ws->big_aux.as<uint8_t>() + (size_t)nbins * 4 + (16 - (nbins * 4) % 16) % 16);
    uint32_t* k32 = ws->big_k32.as<uint32_t>();
    uint32_t* k32o = k32 + R;
    uint32_t* v32 = ws->big_v32.as<uint32_t>();
    uint32_t* v32o = v32 + R;
    size_t tmp_a = 0, tmp_b = 0;
    CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp_a, ws->big_keys.as<uint64_t>(), ws->big_keys2.as<uint64_t>(),
                                      (int64_t)N, 0, key_bits, st));
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_b, k32, k32o, v32, v32o, (int64_t)R, 0, 32, st));
    ws->big_tmp.ensure(std::max(tmp_a, tmp_b) + 256);
    static std::atomic<uint64_t> attr_done{0};
    ensure_dyn_smem(attr_done, rescore_slab_kernel<false>, 64 * (RS_SLAB + 4) * (int)sizeof(float));
    const int cols = std::min(h->dim, RS_SLAB);
    const int stride = ((cols >> 2) & 1) ? cols : cols + 4;
    const int hist_grid = (int)std::min<uint64_t>((N + 255) / 256, (uint64_t)h->sm_count * 8);
    for (uint32_t q0 = 0; q0 < nq; q0 += QC) {
        const uint32_t m = std::min(QC, nq - q0);
        query_prep_direct_kernel<<<(m + 31) / 32, 32, 0, st>>>(q_dev + (size_t)q0 * h->dim, m, h->dim, h->cfg.threshold,
                                                              h->nchunk, ws->qnorm.as<float>(), ws->qpack.as<uint32_t>(), h->qs);
        CU(cudaGetLastError());
        {
            const int qg = pick_qgroup(h, ntiles, m);
            dim3 grid = scan_grid(h, ntiles, m, qg);
            launch_scan<1>(h->nchunk, -1, st, grid, (size_t)qg * h->qs * 4, h->codes, live_of(h, ws), 0, ntiles,
                           ws->qpack.as<uint32_t>(), (int)m, qg, nullptr, nullptr, 0, nullptr, ws->misc.as<uint32_t>(), N, N);
        }
        h->launches.fetch_add(2, std::memory_order_relaxed);
        for (uint32_t qi = 0; qi < m; ++qi) {
            const uint32_t* dist = ws->misc.as<uint32_t>() + (size_t)qi * N;
            const uint32_t gq = q0 + qi;
            CU(cudaMemsetAsync(ws->big_aux.p, 0, ws->big_aux.bytes, st));
            dist_hist_kernel<<<hist_grid, 256, nbins * 4, st>>>(dist, live_of(h, ws), N, nbins, hist);
            cut_kernel<<<1, 32, 0, st>>>(hist, nbins, R, cut);
            CU(cudaGetLastError());
            BigRCut hc;
            CU(cudaMemcpyAsync(&hc, cut, sizeof(hc), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            if (hc.r_eff > 0) {
                cut_compact_kernel<<<hist_grid, 256, 0, st>>>(dist, live_of(h, ws), N, cut, ws->big_keys.as<uint64_t>());
                size_t tb = ws->big_tmp.bytes;
                CU(cub::DeviceRadixSort::SortKeys(ws->big_tmp.p, tb, ws->big_keys.as<uint64_t>(), ws->big_keys2.as<uint64_t>(),
                                                  (int64_t)hc.m, 0, key_bits, st));
            }
            // candidates = the first r_eff sorted keys; rescoring writes the query's R records
            const int q_slots = 2;
            rescore_slab_kernel<false><<<(R + 31) / 32, 32, (size_t)(32 + q_slots) * stride * sizeof(float), st>>>(
                h->rows, h->norms, h->cfg.row_base, h->dim, stride, q_slots, q_dev + (size_t)gq * h->dim,
                ws->qnorm.as<float>() + qi, ws->big_keys2.as<uint64_t>(), 0, &cut->r_eff, R, 1,
                ws->rec_ham.as<uint32_t>(), ws->rec_ids.as<uint64_t>(), ws->rec_score.as<float>(), 0, 0);
            cos_key_kernel<<<(R + 255) / 256, 256, 0, st>>>(ws->rec_score.as<float>(), R, k32, v32);
            size_t tb = ws->big_tmp.bytes;
            CU(cub::DeviceRadixSort::SortPairs(ws->big_tmp.p, tb, k32, k32o, v32, v32o, (int64_t)R, 0, 32, st));
            bigr_emit_kernel<<<(k + 255) / 256, 256, 0, st>>>(v32o, cut, ws->rec_ids.as<uint64_t>(), ws->rec_score.as<float>(), k,
                                                            ids_out + (size_t)gq * k, scores_out + (size_t)gq * k);
            CU(cudaGetLastError());
            if (cand_ids) CU(cudaMemcpyAsync(cand_ids + (size_t)gq * R, ws->rec_ids.p, (size_t)R * 8, cudaMemcpyDeviceToDevice, st));
            if (cand_ham) CU(cudaMemcpyAsync(cand_ham + (size_t)gq * R, ws->rec_ham.p, (size_t)R * 4, cudaMemcpyDeviceToDevice, st));
            h->launches.fetch_add(7, std::memory_order_relaxed);
        }
    }
    CU(cudaStreamSynchronize(st));
    flush_profile(h, ws);
}

// ---- ratio mode across row shards: per-shard histograms, then the shard's members of the global top R ----
// (the protocol of gvdb_bigr.cuh: gvdb_shard_hist_device -> gather -> gvdb_search_shard_ratio_device -> gather ->
// gvdb_merge_shards_ratio_device).  One query at a time like search_big_r; distances in chunks of 16 queries.
void shard_hist(gvdb_index* h, Workspace* ws, cudaStream_t st, const float* q_dev, uint32_t nq, uint32_t* hist_out) {
    const uint32_t nbins = (uint32_t)h->nchunk * 128 + 1;
    CU(cudaMemsetAsync(hist_out, 0, (size_t)nq * nbins * 4, st));
    const uint64_t N = h->n_rows;
    if (N == 0) return;
    if ((h->dim & 3) != 0) fail(GVDB_ERR_NOT_IMPLEMENTED, "rescore_count > 2048 needs dim % 4 == 0");
    const uint32_t ntiles = (uint32_t)tiles_for(N);
    const uint32_t QC = std::min<uint32_t>(nq, 16);
    ws->qpack.ensure((size_t)QC * h->qs * 4);
    ws->qnorm.ensure((size_t)QC * 4);
    ws->misc.ensure((size_t)QC * N * 4);
    const int hist_grid = (int)std::min<uint64_t>((N + 255) / 256, (uint64_t)h->sm_count * 8);
    for (uint32_t q0 = 0; q0 < nq; q0 += QC) {
        const uint32_t m = std::min(QC, nq - q0);
        query_prep_direct_kernel<<<(m + 31) / 32, 32, 0, st>>>(q_dev + (size_t)q0 * h->dim, m, h->dim, h->cfg.threshold,
                                                              h->nchunk, ws->qnorm.as<float>(), ws->qpack.as<uint32_t>(), h->qs);
        const int qg = pick_qgroup(h, ntiles, m);
        dim3 grid = scan_grid(h, ntiles, m, qg);
        launch_scan<1>(h->nchunk, -1, st, grid, (size_t)qg * h->qs * 4, h->codes, live_of(h, ws), 0, ntiles,
                       ws->qpack.as<uint32_t>(), (int)m, qg, nullptr, nullptr, 0, nullptr, ws->misc.as<uint32_t>(), N, N);
        for (uint32_t qi = 0; qi < m; ++qi)
            dist_hist_kernel<<<hist_grid, 256, nbins * 4, st>>>(ws->misc.as<uint32_t>() + (size_t)qi * N, live_of(h, ws), N, nbins,
                                                                hist_out + (size_t)(q0 + qi) * nbins);
        CU(cudaGetLastError());
        h->launches.fetch_add(2 + m, std::memory_order_relaxed);
    }
}

void search_shard_ratio(gvdb_index* h, Workspace* ws, cudaStream_t st, const float* q_dev, uint32_t nq, uint64_t R64,
                        uint32_t kr, const uint32_t* hists_all, uint32_t n_shards, uint32_t my, uint8_t* records) {
    need_all_rows(h);
    if ((h->dim & 3) != 0) fail(GVDB_ERR_NOT_IMPLEMENTED, "rescore_count > 2048 needs dim % 4 == 0");
    const uint64_t N = h->n_rows;
    const uint64_t nrec = (uint64_t)nq * kr;
    uint64_t* out_ids = reinterpret_cast<uint64_t*>(records);
    uint32_t* out_ham = reinterpret_cast<uint32_t*>(records + nrec * 8);
    float* out_score = reinterpret_cast<float*>(records + nrec * 12);
    const uint32_t nbins = (uint32_t)h->nchunk * 128 + 1;
    ws->big_cut.ensure((size_t)nq * sizeof(BigRCut));
    BigRCut* cuts = ws->big_cut.as<BigRCut>();
    shard_cut_kernel<<<nq, 32, 0, st>>>(hists_all, n_shards, my, nq, nbins, (unsigned long long)R64, cuts);
    CU(cudaGetLastError());
    std::vector<BigRCut> hc(nq);
    CU(cudaMemcpyAsync(hc.data(), cuts, (size_t)nq * sizeof(BigRCut), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // this shard's largest share over the batch sizes the per-query arrays (>= 1 so that the launches are well formed)
    uint32_t Rl = 1, Ml = 1;
    for (const BigRCut& c : hc) { Rl = std::max(Rl, c.r_eff); Ml = std::max(Ml, c.m); }
    const uint32_t ntiles = (uint32_t)tiles_for(std::max<uint64_t>(N, 1));
    int key_bits = 32;
    while ((1u << (key_bits - 32)) < nbins) ++key_bits;
    const uint32_t QC = std::min<uint32_t>(nq, 16);
    ws->qpack.ensure((size_t)QC * h->qs * 4);
    ws->qnorm.ensure((size_t)QC * 4);
    ws->misc.ensure((size_t)QC * std::max<uint64_t>(N, 1) * 4);
    ws->big_keys.ensure((size_t)Ml * 8);
    ws->big_keys2.ensure((size_t)Ml * 8);
    ws->big_k32.ensure((size_t)Rl * 8);
    ws->big_v32.ensure((size_t)Rl * 8);
    ws->rec_ham.ensure((size_t)Rl * 4);
    ws->rec_ids.ensure((size_t)Rl * 8);
    ws->rec_score.ensure((size_t)Rl * 4);
    uint32_t* k32 = ws->big_k32.as<uint32_t>();
    uint32_t* k32o = k32 + Rl;
    uint32_t* v32 = ws->big_v32.as<uint32_t>();
    uint32_t* v32o = v32 + Rl;
    size_t tmp_a = 0, tmp_b = 0;
    CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp_a, ws->big_keys.as<uint64_t>(), ws->big_keys2.as<uint64_t>(), (int64_t)Ml, 0, key_bits, st));
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_b, k32, k32o, v32, v32o, (int64_t)Rl, 0, 32, st));
    ws->big_tmp.ensure(std::max(tmp_a, tmp_b) + 256);
    static std::atomic<uint64_t> attr_done{0};
    ensure_dyn_smem(attr_done, rescore_slab_kernel<false>, 64 * (RS_SLAB + 4) * (int)sizeof(float));
    const int cols = std::min(h->dim, RS_SLAB);
    const int stride = ((cols >> 2) & 1) ? cols : cols + 4;
    const int hist_grid = (int)std::min<uint64_t>((N + 255) / 256, (uint64_t)h->sm_count * 8);
    uint32_t kr_eff = 64;
    while (kr_eff < kr) kr_eff <<= 1;
    for (uint32_t q0 = 0; q0 < nq; q0 += QC) {
        const uint32_t m = std::min(QC, nq - q0);
        if (N) {
            query_prep_direct_kernel<<<(m + 31) / 32, 32, 0, st>>>(q_dev + (size_t)q0 * h->dim, m, h->dim, h->cfg.threshold,
                                                                  h->nchunk, ws->qnorm.as<float>(), ws->qpack.as<uint32_t>(), h->qs);
            const int qg = pick_qgroup(h, ntiles, m);
            dim3 grid = scan_grid(h, ntiles, m, qg);
            launch_scan<1>(h->nchunk, -1, st, grid, (size_t)qg * h->qs * 4, h->codes, live_of(h, ws), 0, ntiles,
                           ws->qpack.as<uint32_t>(), (int)m, qg, nullptr, nullptr, 0, nullptr, ws->misc.as<uint32_t>(), N, N);
        }
        for (uint32_t qi = 0; qi < m; ++qi) {
            const uint32_t gq = q0 + qi;
            BigRCut* cut = cuts + gq;
            if (hc[gq].r_eff > 0) {
                const uint32_t* dist = ws->misc.as<uint32_t>() + (size_t)qi * N;
                cut_compact_kernel<<<hist_grid, 256, 0, st>>>(dist, live_of(h, ws), N, cut, ws->big_keys.as<uint64_t>());
                size_t tb = ws->big_tmp.bytes;
                CU(cub::DeviceRadixSort::SortKeys(ws->big_tmp.p, tb, ws->big_keys.as<uint64_t>(), ws->big_keys2.as<uint64_t>(),
                                                  (int64_t)hc[gq].m, 0, key_bits, st));
                // candidates = the first r_eff sorted keys: this shard's members of the global top R
                const int q_slots = 2;
                rescore_slab_kernel<false><<<(Rl + 31) / 32, 32, (size_t)(32 + q_slots) * stride * sizeof(float), st>>>(
                    h->rows, h->norms, h->cfg.row_base, h->dim, stride, q_slots, q_dev + (size_t)gq * h->dim,
                    ws->qnorm.as<float>() + qi, ws->big_keys2.as<uint64_t>(), 0, &cut->r_eff, Rl, 1,
                    ws->rec_ham.as<uint32_t>(), ws->rec_ids.as<uint64_t>(), ws->rec_score.as<float>(), 0, 0);
                cos_key_kernel<<<(Rl + 255) / 256, 256, 0, st>>>(ws->rec_score.as<float>(), Rl, k32, v32);
                tb = ws->big_tmp.bytes;
                CU(cub::DeviceRadixSort::SortPairs(ws->big_tmp.p, tb, k32, k32o, v32, v32o, (int64_t)Rl, 0, 32, st));
            }
            bigr_emit_records_kernel<<<1, 256, (size_t)kr_eff * 8, st>>>(
                v32o, cut, ws->rec_ham.as<uint32_t>(), ws->rec_ids.as<uint64_t>(), ws->rec_score.as<float>(), kr,
                out_ids + (size_t)gq * kr, out_ham + (size_t)gq * kr, out_score + (size_t)gq * kr);
            CU(cudaGetLastError());
            h->launches.fetch_add(6, std::memory_order_relaxed);
        }
    }
    CU(cudaStreamSynchronize(st));
    flush_profile(h, ws);
}

__global__ void ratio_rinv_kernel(const float* __restrict__ norms, uint64_t n, float* __restrict__ rinv) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rinv[i] = norms[i] > 0.0f ? 1.0f / norms[i] : 0.0f;
}
int ratio_units(const gvdb_index* h) { return (h->dim + 31) / 32; }
bool ratio_supported(const gvdb_index* h) {
    const int nu = ratio_units(h);
    return !h->windowed && (h->dim & 3) == 0 && (nu == 2 || nu == 4 || nu == 8 || nu == 12 || nu == 16 || nu == 24) &&
           tc_supported(h->nchunk);
}
// bf16 rows + reciprocal norms for every stored row (rebuilt when rows were added; searches hold the caller's read guard)
void ensure_rows16(gvdb_index* h, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(h->rows16_mu);
    if (h->rows16_rows == h->n_rows) return;
    const int nu = ratio_units(h);
    if (h->rows16_cap < h->n_rows) {
        if (h->rows16) CU(cudaFree(h->rows16));
        if (h->rinv) CU(cudaFree(h->rinv));
        h->rows16 = nullptr; h->rinv = nullptr; h->rows16_cap = 0;
        const uint64_t cap = (std::max<uint64_t>(h->cap_rows, h->n_rows) + 31) / 32 * 32;
        CU(cudaMalloc((void**)&h->rows16, cap * (size_t)nu * 64));
        CU(cudaMalloc((void**)&h->rinv, cap * sizeof(float)));
        h->rows16_cap = cap;
        h->rows16_rows = 0;
    }
    const uint64_t first = h->rows16_rows, n = h->n_rows - first;
    // rows of a partially filled last tile are rewritten together with the new ones (their slots hold stale data otherwise)
    if (first == 0) CU(cudaMemsetAsync(h->rows16, 0, h->rows16_cap * (size_t)nu * 64, st));
    const uint64_t work = n * (uint64_t)nu;
    ratio_rows16_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(h->rows, first, n, h->dim, nu, h->rows16);
    ratio_rinv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->norms + first, n, h->rinv + first);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    h->rows16_rows = h->n_rows;
}

template <int MODE>
void launch_tc_dot(gvdb_index* h, Workspace* ws, cudaStream_t st, const int8_t* q16, const float* thr, uint32_t nq,
                   uint32_t nq_pad, uint2* recs, uint32_t rec_cap, uint32_t* list_counts, uint32_t* overflow,
                   float* dot_out, uint64_t dot_stride, uint32_t* grid_out) {
    const int nu = ratio_units(h);
    const uint32_t n_tiles = (uint32_t)tiles_for(h->n_rows), ngroups = (n_tiles + 3) / 4, nqb = nq_pad / TC_NQ;
    const uint32_t sms = (uint32_t)h->sm_count;
    // row slices: a few items per CTA, never more than one slice per row group
    const uint32_t rsl = std::max<uint32_t>(1, std::min<uint32_t>(ngroups, (4 * sms + nqb - 1) / nqb));
    const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)rsl * nqb, sms);
    if (grid_out) *grid_out = grid;
    const size_t smem = (size_t)((nu + 1) / 2) * TC_STAGE_BYTES;
    Timed t(h, ws, st, K_TCDOT, (double)h->n_rows * nu * 64.0 * nqb, (double)h->n_rows * nq_pad * (nu * 32.0));
#define GVDB_DOT_CASE(N)                                                                             \
    case N: {                                                                                        \
        static std::atomic<uint64_t> attr_done{0};                                                   \
        ensure_dyn_smem(attr_done, tc_dot_kernel<N, MODE>, (int)(((N + 1) / 2) * TC_STAGE_BYTES));   \
        tc_dot_kernel<N, MODE><<<grid, TC_THREADS, smem, st>>>(h->rows16, live_of(h, ws), h->rinv, n_tiles, ngroups, q16, thr, nq, \
                                                               nq_pad, rsl, recs, rec_cap, list_counts, overflow, dot_out,        \
                                                               dot_stride, h->n_rows);                                            \
        break;                                                                                       \
    }
    switch (nu) {
        GVDB_DOT_CASE(2) GVDB_DOT_CASE(4) GVDB_DOT_CASE(8) GVDB_DOT_CASE(12) GVDB_DOT_CASE(16) GVDB_DOT_CASE(24)
        default: fail(GVDB_ERR_NOT_IMPLEMENTED, "tensor-core dot products: unsupported dimension");
    }
#undef GVDB_DOT_CASE
    CU(cudaGetLastError());
}

// ---- ratio mode on the tensor cores (gvdb_ratio.cuh): one tile of <= query_tile queries ------------------------------
// Returns false when the tile was not handled (the caller runs the cut by counting for it).
bool search_ratio_tile(gvdb_index* h, Workspace* ws, cudaStream_t st, const float* q_dev, uint32_t nq, uint32_t k, uint32_t R,
                       uint64_t* ids_out, float* scores_out) {
    constexpr uint32_t Rf = 256;                              // rescore count of the fast path that seeds the bound
    const uint32_t K = (uint32_t)h->nchunk * 128;
    const uint32_t nq_pad = (nq + TC_NQ - 1) / TC_NQ * TC_NQ;
    const uint32_t ntiles = (uint32_t)tiles_for(h->n_rows), ngroups = (ntiles + 3) / 4;
    const int nu = ratio_units(h);
    // 1. fast path: records (stage-1 order) + top k of the Rf closest rows in Hamming distance
    ws->rec_ham.ensure((size_t)nq * Rf * 4);
    ws->rec_ids.ensure((size_t)nq * Rf * 8);
    ws->rec_score.ensure((size_t)nq * Rf * 4);
    ws->r_topk_i.ensure((size_t)nq * k * 8);
    ws->r_topk_s.ensure((size_t)nq * k * 4);
    for (int attempt = 0; attempt < 2; ++attempt) {
        bool optimistic = false;
        FusedTopk fused;
        fused.k = k; fused.ids_out = ws->r_topk_i.as<uint64_t>(); fused.scores_out = ws->r_topk_s.as<float>();
        fused.want_records = true;
        search_core(h, ws, st, q_dev, nq, Rf, ws->rec_ham.as<uint32_t>(), ws->rec_ids.as<uint64_t>(), ws->rec_score.as<float>(),
                    true, attempt == 0, &optimistic, nullptr, &fused);
        if (!fused.done)
            launch_topk(h, ws, st, ws->rec_ids.as<uint64_t>(), ws->rec_score.as<float>(), nq, Rf, k, ws->r_topk_i.as<uint64_t>(),
                        ws->r_topk_s.as<float>());
        bool overflowed = false;
        if (check_overflow(h, ws, st, optimistic, &overflowed)) continue;
        if (overflowed) return false;
        break;
    }
    ensure_rows16(h, st);
    // 2. b*: counting passes of the FP4 scan.  The query blocks (ws->qexp, qpop) are the fast path's; the first guess
    //    comes from the distances of a strided sample of row groups.
    ws->r_state.ensure((size_t)nq * sizeof(RatioState));
    ws->r_counts.ensure((size_t)nq_pad * 4 + 256);
    ws->r_fb.ensure((size_t)nq_pad * 4);
    CU(cudaMemsetAsync(ws->r_fb.p, 0, (size_t)nq_pad * 4, st));
    uint32_t* n_active = ws->r_counts.as<uint32_t>() + nq_pad;      // one word behind the counts
    {
        const uint32_t n_s = std::min<uint32_t>(ngroups, 64), stride_g = std::max<uint32_t>(1, ngroups / n_s);
        const uint32_t n_sample = n_s * TC_ROWS;
        ws->r_misc.ensure((size_t)nq * n_sample * 4);
        tc_update_bias(h, ws, st, nq, nq_pad, 1);
        // group_stride 1 writes by row number: the compact form needs a stride > 1; a corpus of <= 64 groups is its own sample
        if (stride_g > 1)
            launch_tc_scan<1>(h, ws, st, 0, ntiles, n_s, stride_g, nq, nq_pad, nullptr, nullptr, 0, nullptr, ws->r_misc.as<uint32_t>(), n_sample);
        else {
            CU(cudaMemsetAsync(ws->r_misc.p, 0xFF, (size_t)nq * n_sample * 4, st));
            launch_tc_scan<1>(h, ws, st, 0, ntiles, ngroups, 1, nq, nq_pad, nullptr, nullptr, 0, nullptr, ws->r_misc.as<uint32_t>(), n_sample);
        }
        ratio_bstar_init_kernel<<<nq, 256, (K + 1) * 4, st>>>(ws->r_misc.as<uint32_t>(), n_sample, stride_g > 1 ? n_sample : (uint32_t)std::min<uint64_t>(n_sample, h->n_rows),
                                                              K + 1, R, stride_g > 1 ? h->n_rows : h->n_rows, ws->r_state.as<RatioState>(),
                                                              ws->qpack.as<uint32_t>(), h->qs, h->nchunk * 4);
        CU(cudaGetLastError());
    }
    bool converged = false;
    for (int pass = 0; pass < 24 && !converged; ++pass) {
        tc_update_bias(h, ws, st, nq, nq_pad, 0);
        CU(cudaMemsetAsync(ws->r_counts.p, 0, (size_t)nq_pad * 4 + 256, st));
        launch_tc_scan<3>(h, ws, st, 0, ntiles, ngroups, 1, nq, nq_pad, nullptr, nullptr, 0, nullptr, ws->r_counts.as<uint32_t>(), 0);
        ratio_bstar_update_kernel<<<(nq + 127) / 128, 128, 0, st>>>(ws->r_state.as<RatioState>(), ws->r_counts.as<uint32_t>(), nq, R, K,
                                                                   ws->qpack.as<uint32_t>(), h->qs, h->nchunk * 4, n_active);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(ws->h_flag + 4, n_active, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        converged = ws->h_flag[4] == 0;
    }
    if (!converged) return false;
    // 3. the dense filter: rows whose approximate cosine beats (the fast path's k-th best) - eps
    ws->r_thr.ensure((size_t)nq_pad * 4);
    ratio_threshold_kernel<<<(nq_pad + 127) / 128, 128, 0, st>>>(ws->r_topk_s.as<float>(), k, ws->qnorm.as<float>(), nq, nq_pad, RATIO_EPS,
                                                                ws->r_thr.as<float>(), ws->r_fb.as<uint32_t>());
    ws->r_q16.ensure((size_t)(nq_pad / TC_NQ) * tc_qblock_bytes(nu));
    {
        const uint64_t words = (uint64_t)nq_pad * nu * 16;
        ratio_q16_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(q_dev, nq, nq_pad, h->dim, nu, ws->r_q16.as<int8_t>());
    }
    const uint32_t rec_cap = 16384;
    const uint32_t nlists_max = (uint32_t)h->sm_count * TC_EPI_WARPS;
    ws->tc_recs.ensure((size_t)nlists_max * rec_cap * sizeof(uint2));
    ws->list_counts.ensure((size_t)nlists_max * 4);
    CU(cudaMemsetAsync(ws->list_counts.p, 0, (size_t)nlists_max * 4, st));
    uint32_t grid = 0;
    launch_tc_dot<0>(h, ws, st, ws->r_q16.as<int8_t>(), ws->r_thr.as<float>(), nq, nq_pad, ws->tc_recs.as<uint2>(), rec_cap,
                     ws->list_counts.as<uint32_t>(), ws->flag.as<uint32_t>(), nullptr, 0, &grid);
    // 4. survivors -> E lists -> exact rescoring -> merge with the fast path's records
    ws->r_ecnt.ensure((size_t)nq * 4);
    ws->r_ekeys.ensure((size_t)nq * RATIO_E_CAP * 8);
    ws->r_esc.ensure((size_t)nq * RATIO_E_CAP * 4);
    CU(cudaMemsetAsync(ws->r_ecnt.p, 0, (size_t)nq * 4, st));
    CU(cudaMemsetAsync(ws->r_ekeys.p, 0xFF, (size_t)nq * RATIO_E_CAP * 8, st));
    h->launches.fetch_add(6, std::memory_order_relaxed);
    ratio_scatter_kernel<<<dim3(4, grid * TC_EPI_WARPS), 256, 0, st>>>(
        ws->tc_recs.as<uint2>(), rec_cap, ws->list_counts.as<uint32_t>(), h->codes, h->nchunk, ws->qpack.as<uint32_t>(), h->qs,
        ws->r_state.as<RatioState>(), ws->rec_ham.as<uint32_t>(), ws->rec_ids.as<uint64_t>(), Rf, h->cfg.row_base, K,
        ws->r_ecnt.as<uint32_t>(), ws->r_ekeys.as<uint64_t>(), RATIO_E_CAP, ws->r_fb.as<uint32_t>());
    CU(cudaGetLastError());
    {
        const uint64_t pairs = (uint64_t)nq * RATIO_E_CAP;
        ws->r_list.ensure(pairs * 4 + 256);
        uint32_t* list = ws->r_list.as<uint32_t>() + 64;
        uint32_t* count = ws->r_list.as<uint32_t>();
        CU(cudaMemsetAsync(count, 0, 4, st));
        OwnedPair pred{ws->r_ekeys.as<uint64_t>(), h->cfg.row_base, 0, h->n_rows};
        owned_compact_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(pred, (uint32_t)pairs, list, count);
        static std::atomic<uint64_t> attr_done{0};
        ensure_dyn_smem(attr_done, rescore_owned_ring_kernel, (int)RO_SMEM);
        Timed t(h, ws, st, K_RESCORE);
        // the list is short (a few survivors per query): 4096 warps cover 131072 entries, the kernel's warps beyond the list exit
        const unsigned blocks = (unsigned)std::min<uint64_t>((pairs + 32 * RO_WARPS - 1) / (32 * RO_WARPS), 65536);
        rescore_owned_ring_kernel<<<blocks, 32 * RO_WARPS, RO_SMEM, st>>>(
            h->rows_base(), h->norms, h->cfg.row_base, h->dim, q_dev, ws->r_ekeys.as<uint64_t>(), list, count, RATIO_E_CAP,
            ws->r_esc.as<float>(), nullptr, 0, 1);
    }
    CU(cudaGetLastError());
    {
        uint32_t n_eff = 64;
        while (n_eff < Rf + RATIO_E_CAP) n_eff <<= 1;
        static std::atomic<uint64_t> attr_done{0};
        ensure_dyn_smem(attr_done, ratio_finish_kernel, (int)((RATIO_E_CAP + 4096) * 8));
        Timed t(h, ws, st, K_TOPK);
        ratio_finish_kernel<<<nq, 1024, (size_t)(RATIO_E_CAP + n_eff) * 8, st>>>(
            ws->rec_ids.as<uint64_t>(), ws->rec_score.as<float>(), Rf, ws->r_ekeys.as<uint64_t>(), ws->r_esc.as<float>(),
            ws->r_ecnt.as<uint32_t>(), RATIO_E_CAP, n_eff, ws->r_fb.as<uint32_t>(), k, ids_out, scores_out);
    }
    CU(cudaGetLastError());
    // 5. queries the filter could not vouch for (a tie at b*, a full list, a non-positive bound): the cut by counting
    if (ws->h_fb_bytes < (size_t)nq * 4) {
        if (ws->h_fb) cudaFreeHost(ws->h_fb);
        ws->h_fb = nullptr; ws->h_fb_bytes = 0;
        CU(cudaMallocHost((void**)&ws->h_fb, (size_t)nq * 4));
        ws->h_fb_bytes = (size_t)nq * 4;
    }
    CU(cudaMemcpyAsync(ws->h_fb, ws->r_fb.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ws->h_flag, ws->flag.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    flush_profile(h, ws);
    if (ws->h_flag[0]) return false;                          // a record list of the dense pass overflowed
    uint64_t n_fb = 0;
    for (uint32_t q = 0; q < nq; ++q)
        if (ws->h_fb[q]) {
            ++n_fb;
            search_big_r(h, ws, st, q_dev + (size_t)q * h->dim, 1, k, R, ids_out + (size_t)q * k, scores_out + (size_t)q * k, nullptr, nullptr);
        }
    if (n_fb) {
        std::lock_guard<std::mutex> lk(h->prof_mu);
        h->prof.ratio_fallback_queries += n_fb;
    }
    return true;
}

// h_ids / h_scores (optional, pinned): the k-lists are also copied there BEFORE the call's one
// synchronisation, so the host-pointer entry point pays a single stream sync.
void search_device(gvdb_index* h, Workspace* ws, cudaStream_t st, const float* q_dev, uint32_t nq,
                   uint32_t k, uint32_t R, uint64_t* ids_out, float* scores_out, uint64_t* cand_ids,
                   uint32_t* cand_ham, uint64_t* h_ids = nullptr, float* h_scores = nullptr) {
    if (k > R) fail(GVDB_ERR_INVALID_ARGUMENT, "k must be <= rescore_count");
    if (R > kMaxR) {
        // ratio mode: batches the tensor cores are worth it for go through the dense filter (gvdb_ratio.cuh), tile by tile;
        // everything else (few queries, candidate lists asked for, k > 256, dims it does not support) is the cut by counting
        const bool dense = h->ratio_tc && ratio_supported(h) && nq >= h->tc_min_q && !cand_ids && !cand_ham && k >= 1 && k <= 256 &&
                           h->n_rows >= 65536 && (uint64_t)R <= h->n_rows;
        if (dense) {
            const uint32_t QT = h->query_tile;
            for (uint32_t q0 = 0; q0 < nq; q0 += QT) {
                const uint32_t m = std::min(QT, nq - q0);
                if (m < h->tc_min_q || !search_ratio_tile(h, ws, st, q_dev + (size_t)q0 * h->dim, m, k, R, ids_out + (size_t)q0 * k,
                                                          scores_out + (size_t)q0 * k))
                    search_big_r(h, ws, st, q_dev + (size_t)q0 * h->dim, m, k, R, ids_out + (size_t)q0 * k, scores_out + (size_t)q0 * k,
                                 nullptr, nullptr);
            }
        } else
        search_big_r(h, ws, st, q_dev, nq, k, R, ids_out, scores_out, cand_ids, cand_ham);
        if (h_ids) {
            CU(cudaMemcpyAsync(h_ids, ids_out, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(h_scores, scores_out, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        return;
    }
    ws->rec_ham.ensure((size_t)nq * R * 4);
    ws->rec_ids.ensure((size_t)nq * R * 8);
    ws->rec_score.ensure((size_t)nq * R * 4);
    for (int attempt = 0; attempt < 2; ++attempt) {
        bool optimistic = false;
        FusedTopk fused;
        fused.k = k; fused.ids_out = ids_out; fused.scores_out = scores_out;
        fused.want_records = cand_ids != nullptr || cand_ham != nullptr;
        search_core(h, ws, st, q_dev, nq, R, ws->rec_ham.as<uint32_t>(), ws->rec_ids.as<uint64_t>(),
                    ws->rec_score.as<float>(), true, attempt == 0, &optimistic, nullptr, k > 0 ? &fused : nullptr);
        if (!fused.done)
            launch_topk(h, ws, st, ws->rec_ids.as<uint64_t>(), ws->rec_score.as<float>(), nq, R, k, ids_out, scores_out);
        if (cand_ids) CU(cudaMemcpyAsync(cand_ids, ws->rec_ids.p, (size_t)nq * R * 8, cudaMemcpyDeviceToDevice, st));
        if (cand_ham) CU(cudaMemcpyAsync(cand_ham, ws->rec_ham.p, (size_t)nq * R * 4, cudaMemcpyDeviceToDevice, st));
        if (h_ids) {
            CU(cudaMemcpyAsync(h_ids, ids_out, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(h_scores, scores_out, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
        }
        bool overflowed = false;
        if (check_overflow(h, ws, st, optimistic, &overflowed)) continue;
        if (overflowed) {
            // exact fallback without capacity limits: the cut by counting (slow path)
            if ((h->dim & 3) != 0)
                fail(GVDB_ERR_INDEX, "candidate buffer overflow in the Hamming scan and dim % 4 != 0: no fallback");
            h->overflow_fallbacks.fetch_add(1, std::memory_order_relaxed);
            search_big_r(h, ws, st, q_dev, nq, k, R, ids_out, scores_out, cand_ids, cand_ham);
            if (h_ids) {
                CU(cudaMemcpyAsync(h_ids, ids_out, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
                CU(cudaMemcpyAsync(h_scores, scores_out, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
            }
        }
        break;
    }
}

// Exact flat search: same segment/select machinery with key = image(1 - cos) << 32 | row.
void flat_device(gvdb_index* h, Workspace* ws, cudaStream_t st, const float* q_dev, uint32_t nq,
                 uint32_t k, uint64_t* ids_out, float* dist_out, bool sim = false, float sim_threshold = 0.f,
                 bool use_threshold = false) {
    need_all_rows(h);
    if (h->n_rows == 0) fail(GVDB_ERR_INDEX_NOT_BUILT, "index not built: search before any add");
    if (k == 0) fail(GVDB_ERR_INVALID_ARGUMENT, "k must be >= 1");
    if (k > kMaxR) fail(GVDB_ERR_NOT_IMPLEMENTED, "flat search with k > 2048 is not implemented yet");
    const uint32_t cap = pick_cap(k);
    const uint32_t QT = h->query_tile;
    const uint32_t qt_max = std::min(QT, nq);
    ws->qpack.ensure((size_t)qt_max * h->qs * 4);
    ws->qnorm.ensure((size_t)qt_max * 4);
    ws->cnt.ensure((size_t)qt_max * 4 * CNT_STRIDE);
    ws->flag.ensure(256);
    ws->buf.ensure((size_t)qt_max * cap * 8);
    ws->misc.ensure((size_t)qt_max * 4);   // per-query key thresholds
    // Two schedules.  Fast: the first 4096 rows emit everything, later segments grow geometrically and emit only
    // rows that beat the k-th best so far — few launches, but a segment can overflow a query's candidate buffer
    // when the rows seen so far say nothing about the rest (a tombstoned or filtered-out prefix, a corpus stored in
    // order of relevance).  Safe: segments of cap - k rows; a buffer holds the k kept keys plus at most one segment,
    // so it cannot overflow whatever the data — the call is repeated with it when the fast one overflowed.
    for (int attempt = 0; attempt < 2; ++attempt) {
    const bool safe = attempt == 1;
    CU(cudaMemsetAsync(ws->flag.p, 0, 256, st));
    const uint64_t seg0 = safe ? cap - k : 4096;
    const uint64_t g = std::max<uint32_t>(2, cap / (4 * k));
    for (uint32_t qt0 = 0; qt0 < nq; qt0 += QT) {
        const uint32_t nqt = std::min(QT, nq - qt0);
        const float* qd = q_dev + (size_t)qt0 * h->dim;
        {
            Timed t(h, ws, st, K_PREP);
            ingest_kernel<true><<<(nqt + STAGE_ROWS - 1) / STAGE_ROWS, STAGE_ROWS, 0, st>>>(
                qd, nqt, h->dim, h->cfg.threshold, h->nchunk, 0, nullptr, ws->qnorm.as<float>(), nullptr,
                ws->qpack.as<uint32_t>(), h->qs);
        }
        CU(cudaGetLastError());
        CU(cudaMemsetAsync(ws->cnt.p, 0, (size_t)nqt * 4 * CNT_STRIDE, st));
        CU(cudaMemsetAsync(ws->misc.p, 0xff, (size_t)nqt * 4, st));   // TAU_ALL
        uint64_t lo = 0;
        while (lo < h->n_rows) {
            uint64_t hi = std::min<uint64_t>(safe ? lo + seg0 : (lo == 0 ? seg0 : lo * g), h->n_rows);
            dim3 grid((unsigned)((hi - lo + FLAT_TM - 1) / FLAT_TM), (nqt + FLAT_TN - 1) / FLAT_TN, 1);
            {
                Timed t(h, ws, st, K_FLAT);
                if (sim)
                    flat_scan_kernel<true><<<grid, FLAT_THREADS, 0, st>>>(
                        h->rows_base(), h->norms, live_of(h, ws), lo, hi, h->dim, qd, ws->qnorm.as<float>(), nqt,
                        ws->misc.as<uint32_t>(), ws->cnt.as<uint32_t>(), ws->buf.as<uint64_t>(), cap,
                        ws->flag.as<uint32_t>(), sim_threshold, use_threshold ? 1 : 0);
                else
                    flat_scan_kernel<false><<<grid, FLAT_THREADS, 0, st>>>(
                        h->rows_base(), h->norms, live_of(h, ws), lo, hi, h->dim, qd, ws->qnorm.as<float>(), nqt,
                        ws->misc.as<uint32_t>(), ws->cnt.as<uint32_t>(), ws->buf.as<uint64_t>(), cap,
                        ws->flag.as<uint32_t>(), 0.f, 0);
            }
            CU(cudaGetLastError());
            {
                Timed t(h, ws, st, K_SELECT);
                select_kernel<<<nqt, SORT_THREADS, SORT_N * 8, st>>>(
                    ws->buf.as<uint64_t>(), cap, ws->cnt.as<uint32_t>(), k, ws->misc.as<uint32_t>(), 1, 0);
            }
            CU(cudaGetLastError());
            lo = hi;
        }
        h->launches.fetch_add(1, std::memory_order_relaxed);
        flat_emit_kernel<<<(nqt * k + 255) / 256, 256, 0, st>>>(
            ws->buf.as<uint64_t>(), cap, ws->cnt.as<uint32_t>(), nqt, k, h->cfg.row_base,
            ids_out + (size_t)qt0 * k, dist_out + (size_t)qt0 * k, sim ? 1 : 0);
        CU(cudaGetLastError());
    }
    bool overflowed = false;
    check_overflow(h, ws, st, false, &overflowed);
    if (!overflowed) break;
    if (safe) fail(GVDB_ERR_INDEX, "flat search: candidate buffer overflow under the safe schedule (internal error)");
    h->overflow_fallbacks.fetch_add(1, std::memory_order_relaxed);
    }
}

// Rows [n_rows, n_rows + n) from DEVICE memory `src`: codes + norms + live bits for all of them,
// f32 copies for the rows the index keeps (all of them, or the window's share).  `src` may be the
// rows' final place (src_is_final: the host path copies straight into it).
void add_device_impl(gvdb_index* h, cudaStream_t st, const float* src, uint64_t n, bool src_is_final,
                     uint64_t* first_out) {
    if (first_out) *first_out = h->n_rows;
    if (n == 0) return;
    if (h->n_rows + n > 0xfffffff0ull) fail(GVDB_ERR_INDEX, "a shard holds at most 2^32-16 rows");
    grow(h, h->n_rows + n);
    const uint64_t first = h->n_rows;
    if (!src_is_final) {
        uint64_t lo = first, hi = first + n;                 // rows to keep
        if (h->windowed) { lo = std::max(lo, h->win_first); hi = std::min(hi, h->win_first + h->win_count); }
        if (lo < hi)
            CU(cudaMemcpyAsync(h->rows + (lo - (h->windowed ? h->win_first : 0)) * (size_t)h->dim,
                               src + (lo - first) * (size_t)h->dim, (hi - lo) * (size_t)h->dim * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
    }
    ingest_kernel<false><<<(unsigned)((n + STAGE_ROWS - 1) / STAGE_ROWS), STAGE_ROWS, 0, st>>>(
        src, n, h->dim, h->cfg.threshold, h->nchunk, first, h->codes, h->norms, h->live, nullptr, 0);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    h->n_rows += n;
    h->n_live += n;
}

template <class F>
gvdb_status guarded(F&& f) {
    try {
        f();
        return GVDB_OK;
    } catch (const Err& e) {
        g_err = e.msg;
        return e.st;
    } catch (const std::exception& e) {
        g_err = e.what();
        return GVDB_ERR_INDEX;
    } catch (...) {
        g_err = "unknown error";
        return GVDB_ERR_INDEX;
    }
}

void need(const void* p, const char* what) {
    if (!p) fail(GVDB_ERR_INVALID_ARGUMENT, std::string(what) + " is NULL");
}

}  // namespace

// =========================================================================================
extern "C" {

uint32_t gvdb_abi_version(void) { return GVDB_ABI_VERSION; }
const char* gvdb_last_error(void) { return g_err.c_str(); }

gvdb_status gvdb_create(const gvdb_config* cfg, gvdb_index** out) {
    return guarded([&] {
        need(cfg, "cfg"); need(out, "out");
        *out = nullptr;
        if (cfg->struct_size != sizeof(gvdb_config))
            fail(GVDB_ERR_INVALID_ARGUMENT, "gvdb_config.struct_size mismatch");
        if (cfg->dim == 0) fail(GVDB_ERR_INVALID_VECTOR_DIMENSION, "dimension must be > 0");
        if (cfg->dim > 4096) fail(GVDB_ERR_INVALID_ARGUMENT, "dimension > 4096 is not supported");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            fail(GVDB_ERR_INDEX, std::string("no usable CUDA device (there is no CPU fallback): ") +
                                     cudaGetErrorString(e));
        if (cfg->device < 0 || cfg->device >= ndev) fail(GVDB_ERR_INVALID_ARGUMENT, "bad device ordinal");
        DeviceGuard dg(cfg->device);
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, cfg->device));
        if (prop.major < 10)
            fail(GVDB_ERR_INDEX, "this library is built for sm_100a (B200) only");
        std::unique_ptr<gvdb_index> h(new gvdb_index());
        h->cfg = *cfg;
        h->dim = (int)cfg->dim;
        h->nbytes = (h->dim + 7) / 8;
        int want = (h->dim + 127) / 128;
        for (int c : kSupportedChunks) if (c >= want) { h->nchunk = c; break; }
        h->qs = h->nchunk * 4 + 4;
        h->sm_count = prop.multiProcessorCount;
        if (const char* s = getenv("GVDB_QUERY_TILE")) h->query_tile = std::max(1, atoi(s));
        if (const char* s = getenv("GVDB_SCAN_NCSA")) h->scan_variant = atoi(s);
        if (const char* s = getenv("GVDB_TC_MIN_Q")) h->tc_min_q = (uint32_t)std::max(1, atoi(s));
        if (const char* s = getenv("GVDB_SEG0_ROWS")) h->seg0_rows = (uint32_t)std::max(32, atoi(s)) / 32 * 32;
        if (const char* s = getenv("GVDB_TC_QB")) h->tc_qb_force = (uint32_t)std::max(0, atoi(s));
        if (const char* s = getenv("GVDB_OPT_M")) h->opt_m = (uint32_t)std::max(0, atoi(s));
        if (const char* s = getenv("GVDB_SAMPLE_DIV")) h->sample_div = (uint32_t)std::max(1, atoi(s));
        if (const char* s = getenv("GVDB_RATIO_TC")) h->ratio_tc = atoi(s) != 0;
        h->query_tile = std::min<uint32_t>(h->query_tile, 32768);   // survivor records carry the query in 16 bits
        if (const char* s = getenv("GVDB_SEG_GROWTH")) h->seg_growth = (uint32_t)std::max(0, atoi(s));
        if (const char* s = getenv("GVDB_SCAN_CTAS_PER_SM")) h->scan_ctas_per_sm = std::max(1, atoi(s));
        if (cfg->flags & GVDB_FLAG_ROW_WINDOW) {
            h->windowed = true;
            h->win_first = cfg->window_first;
            h->win_count = cfg->window_count;
            if (h->win_count) CU(cudaMalloc(&h->rows, h->win_count * (size_t)h->dim * sizeof(float)));
        }
        if (cfg->capacity_rows) grow(h.get(), cfg->capacity_rows);
        *out = h.release();
    });
}

void gvdb_destroy(gvdb_index* h) {
    if (!h) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    if (h->peer_rows_dev) cudaFree((void*)h->peer_rows_dev);
    delete h->xchg;
    h->pool.clear();
    if (h->rows) cudaFree(h->rows);
    if (h->codes) cudaFree(h->codes);
    if (h->norms) cudaFree(h->norms);
    if (h->live) cudaFree(h->live);
    if (h->rows16) cudaFree(h->rows16);
    if (h->rinv) cudaFree(h->rinv);
    if (h->h_async_flag) cudaFreeHost(h->h_async_flag);
    delete h;
    cudaSetDevice(prev);
}

gvdb_status gvdb_reserve(gvdb_index* h, uint64_t capacity_rows) {
    return guarded([&] {
        need(h, "index");
        DeviceGuard dg(h->cfg.device);
        grow(h, capacity_rows);
    });
}

gvdb_status gvdb_add(gvdb_index* h, const float* rows, uint64_t n, uint64_t* first_row_out) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_add");
        need(h, "index");
        if (n) need(rows, "rows");
        DeviceGuard dg(h->cfg.device);
        if (first_row_out) *first_row_out = h->n_rows;
        if (n == 0) return;
        grow(h, h->n_rows + n);
        WsLease lease(h, nullptr, false);
        if (!h->windowed) {
            // rows go straight into their final place in HBM; ingest reads them from there
            float* dst = h->rows + h->n_rows * (size_t)h->dim;
            CU(cudaMemcpyAsync(dst, rows, n * (size_t)h->dim * sizeof(float), cudaMemcpyHostToDevice, lease.stream));
            add_device_impl(h, lease.stream, dst, n, true, nullptr);
        } else {
            // windowed: stage chunks, keep only the window's share
            const uint64_t chunk = 65536;
            lease.ws->q_in.ensure(std::min(chunk, n) * (size_t)h->dim * 4);
            for (uint64_t i0 = 0; i0 < n; i0 += chunk) {
                const uint64_t m = std::min(chunk, n - i0);
                CU(cudaMemcpyAsync(lease.ws->q_in.p, rows + i0 * h->dim, m * (size_t)h->dim * 4, cudaMemcpyHostToDevice, lease.stream));
                add_device_impl(h, lease.stream, lease.ws->q_in.as<float>(), m, false, nullptr);
            }
        }
    });
}

gvdb_status gvdb_add_device(gvdb_index* h, void* stream, const float* rows_dev, uint64_t n,
                            uint64_t* first_row_out) {
    return guarded([&] {
        need(h, "index");
        if (n) need(rows_dev, "rows_dev");
        DeviceGuard dg(h->cfg.device);
        add_device_impl(h, (cudaStream_t)stream, rows_dev, n, false, first_row_out);
    });
}

gvdb_status gvdb_remove(gvdb_index* h, uint64_t local_row, int32_t* was_live_out) {
    return guarded([&] {
        need(h, "index");
        DeviceGuard dg(h->cfg.device);
        int32_t was = 0;
        if (local_row < h->n_rows) {
            uint32_t word = 0;
            CU(cudaMemcpy(&word, h->live + (local_row >> 5), 4, cudaMemcpyDeviceToHost));
            if (word & (1u << (local_row & 31))) {
                was = 1;
                word &= ~(1u << (local_row & 31));
                CU(cudaMemcpy(h->live + (local_row >> 5), &word, 4, cudaMemcpyHostToDevice));
                h->n_live -= 1;
            }
        }
        if (was_live_out) *was_live_out = was;
    });
}

gvdb_status gvdb_clear(gvdb_index* h) {
    return guarded([&] {
        need(h, "index");
        DeviceGuard dg(h->cfg.device);
        CU(cudaDeviceSynchronize());
        if (h->cap_rows) {
            CU(cudaMemset(h->codes, 0, tiles_for(h->cap_rows) * h->nchunk * 32 * sizeof(uint4)));
            CU(cudaMemset(h->live, 0, tiles_for(h->cap_rows) * sizeof(uint32_t)));
        }
        h->n_rows = 0;
        h->n_live = 0;
    });
}

uint64_t gvdb_len(const gvdb_index* h) { return h ? h->n_live : 0; }

gvdb_status gvdb_get_stats(const gvdb_index* h, gvdb_stats* out) {
    return guarded([&] {
        need(h, "index"); need(out, "out");
        out->vector_count = h->n_live;
        out->rows = h->n_rows;
        out->dimension = (uint64_t)h->dim;
        out->memory_usage = h->n_rows * (uint64_t)h->dim * 4;
        out->code_bytes_per_row = (uint64_t)h->nchunk * 16;
        out->hbm_bytes = h->cap_rows * ((uint64_t)h->dim * 4 + 4) +
                         tiles_for(h->cap_rows) * ((uint64_t)h->nchunk * 512 + 4);
    });
}

// ---- persistence ------------------------------------------------------------------------------------
namespace {
struct FileHeader {
    char magic[8];
    uint32_t version, dim;
    float threshold, rescore_ratio;
    uint64_t rows, live_rows, row_base;
    uint32_t code_bytes, flags;      // flags bit 0: `checksum` covers the four sections
    uint64_t checksum;
};
static_assert(sizeof(FileHeader) == 64, "header is 64 bytes");
constexpr uint32_t FILE_FLAG_CHECKSUM = 1u;

// FNV-1a over the little-endian u64 words of the four sections' bytes (codes, norms, live, rows, in file order,
// WITHOUT their zero padding, taken as one byte stream; the last partial word zero-extended), then over the
// stream's length in bytes.  Documented in include/gvdb.h so that any reader of the file can check it.
struct SectionHash {
    uint64_t h = 0xcbf29ce484222325ull, total = 0;
    uint8_t carry[8];
    size_t nc = 0;
    static constexpr uint64_t P = 0x100000001b3ull;
    void update(const void* data, size_t n) {
        const uint8_t* p = static_cast<const uint8_t*>(data);
        total += n;
        if (nc) {
            while (n && nc < 8) { carry[nc++] = *p++; --n; }
            if (nc < 8) return;
            uint64_t w; memcpy(&w, carry, 8);
            h = (h ^ w) * P; nc = 0;
        }
        for (; n >= 8; n -= 8, p += 8) { uint64_t w; memcpy(&w, p, 8); h = (h ^ w) * P; }
        while (n) { carry[nc++] = *p++; --n; }
    }
    uint64_t digest() const {
        uint64_t r = h;
        if (nc) { uint64_t w = 0; memcpy(&w, carry, nc); r = (r ^ w) * P; }
        return (r ^ total) * P;
    }
};
uint64_t pad64(uint64_t n) { return (n + 63) / 64 * 64; }

struct File {
    FILE* f = nullptr;
    File(const char* path, const char* mode) : f(fopen(path, mode)) {}
    ~File() { if (f) fclose(f); }
};

// device <-> file through a pinned staging buffer
void dev_to_file(FILE* f, const void* dev, uint64_t bytes, uint64_t padded, void* stage, size_t stage_bytes,
                 SectionHash& hash) {
    for (uint64_t off = 0; off < bytes; off += stage_bytes) {
        const size_t m = (size_t)std::min<uint64_t>(stage_bytes, bytes - off);
        CU(cudaMemcpy(stage, static_cast<const uint8_t*>(dev) + off, m, cudaMemcpyDeviceToHost));
        hash.update(stage, m);
        if (fwrite(stage, 1, m, f) != m) fail(GVDB_ERR_INDEX, "gvdb_save: short write");
    }
    static const uint8_t zeros[64] = {0};
    if (padded > bytes && fwrite(zeros, 1, (size_t)(padded - bytes), f) != padded - bytes)
        fail(GVDB_ERR_INDEX, "gvdb_save: short write");
}
void file_to_dev(FILE* f, void* dev, uint64_t bytes, uint64_t padded, void* stage, size_t stage_bytes,
                 SectionHash& hash) {
    for (uint64_t off = 0; off < bytes; off += stage_bytes) {
        const size_t m = (size_t)std::min<uint64_t>(stage_bytes, bytes - off);
        if (fread(stage, 1, m, f) != m) fail(GVDB_ERR_INDEX, "gvdb_load: truncated file");
        hash.update(stage, m);
        CU(cudaMemcpy(static_cast<uint8_t*>(dev) + off, stage, m, cudaMemcpyHostToDevice));
    }
    if (padded > bytes && fseek(f, (long)(padded - bytes), SEEK_CUR) != 0) fail(GVDB_ERR_INDEX, "gvdb_load: truncated file");
}
}  // namespace

gvdb_status gvdb_save(gvdb_index* h, const char* path) {
    return guarded([&] {
        need(h, "index"); need(path, "path");
        need_all_rows(h);
        DeviceGuard dg(h->cfg.device);
        CU(cudaDeviceSynchronize());
        // written beside the target and renamed over it once complete and on disk: a crash mid-save leaves the
        // previous shard file intact
        const std::string tmp_path = std::string(path) + ".tmp";
        struct Unlink { const std::string& p; bool armed = true; ~Unlink() { if (armed) ::remove(p.c_str()); } } unlink_tmp{tmp_path};
        File file(tmp_path.c_str(), "wb");
        if (!file.f) fail(GVDB_ERR_INDEX, std::string("gvdb_save: cannot open ") + tmp_path);
        FileHeader hd{};
        memcpy(hd.magic, "GVDBIDX1", 8);
        hd.version = 1; hd.dim = (uint32_t)h->dim;
        hd.threshold = h->cfg.threshold; hd.rescore_ratio = h->cfg.rescore_ratio;
        hd.rows = h->n_rows; hd.live_rows = h->n_live; hd.row_base = h->cfg.row_base;
        hd.code_bytes = (uint32_t)h->nbytes;
        hd.flags = FILE_FLAG_CHECKSUM;          // the checksum itself is written once the sections are out
        if (fwrite(&hd, sizeof(hd), 1, file.f) != 1) fail(GVDB_ERR_INDEX, "gvdb_save: short write");
        SectionHash hash;
        const size_t stage_bytes = 32u << 20;
        void* stage = nullptr;
        CU(cudaMallocHost(&stage, stage_bytes));
        struct Free { void* p; ~Free() { cudaFreeHost(p); } } guard{stage};
        const uint64_t N = h->n_rows;
        {   // codes: blocked -> reference byte layout, in chunks
            const uint64_t chunk = stage_bytes / (uint64_t)h->nbytes;
            DevBuf tmp; tmp.ensure(std::max<uint64_t>(1, std::min(chunk, N)) * h->nbytes);
            uint64_t written = 0;
            for (uint64_t i0 = 0; i0 < N; i0 += chunk) {
                const uint64_t m = std::min(chunk, N - i0);
                unblock_codes_kernel<<<(unsigned)((m + 255) / 256), 256>>>(h->codes, h->nchunk, i0, m, h->nbytes, tmp.as<uint8_t>());
                CU(cudaGetLastError());
                dev_to_file(file.f, tmp.p, m * h->nbytes, m * h->nbytes, stage, stage_bytes, hash);
                written += m * h->nbytes;
            }
            tmp.release();
            static const uint8_t zeros[64] = {0};
            const uint64_t padn = pad64(written) - written;
            if (padn && fwrite(zeros, 1, (size_t)padn, file.f) != padn) fail(GVDB_ERR_INDEX, "gvdb_save: short write");
        }
        dev_to_file(file.f, h->norms, N * 4, pad64(N * 4), stage, stage_bytes, hash);
        dev_to_file(file.f, h->live, tiles_for(N) * 4, pad64(tiles_for(N) * 4), stage, stage_bytes, hash);
        dev_to_file(file.f, h->rows, N * (uint64_t)h->dim * 4, N * (uint64_t)h->dim * 4, stage, stage_bytes, hash);
        hd.checksum = hash.digest();
        if (fseek(file.f, 0, SEEK_SET) != 0 || fwrite(&hd, sizeof(hd), 1, file.f) != 1)
            fail(GVDB_ERR_INDEX, "gvdb_save: cannot write the checksum");
        if (fflush(file.f) != 0 || fsync(fileno(file.f)) != 0) fail(GVDB_ERR_INDEX, "gvdb_save: flush failed");
        fclose(file.f); file.f = nullptr;
        if (::rename(tmp_path.c_str(), path) != 0) fail(GVDB_ERR_INDEX, std::string("gvdb_save: cannot rename onto ") + path);
        unlink_tmp.armed = false;
    });
}

gvdb_status gvdb_load(const char* path, int32_t device, gvdb_index** out) {
    return guarded([&] {
        need(path, "path"); need(out, "out");
        *out = nullptr;
        File file(path, "rb");
        if (!file.f) fail(GVDB_ERR_INDEX, std::string("gvdb_load: cannot open ") + path);
        FileHeader hd{};
        if (fread(&hd, sizeof(hd), 1, file.f) != 1 || memcmp(hd.magic, "GVDBIDX1", 8) != 0 || hd.version != 1)
            fail(GVDB_ERR_INDEX, "gvdb_load: not a gvdb index file");
        if (hd.code_bytes != (hd.dim + 7) / 8) fail(GVDB_ERR_INDEX, "gvdb_load: inconsistent header");
        if (hd.rows > 0xfffffff0ull || hd.live_rows > hd.rows) fail(GVDB_ERR_INDEX, "gvdb_load: inconsistent header (row counts)");
        gvdb_config cfg{};
        cfg.struct_size = sizeof(cfg); cfg.dim = hd.dim; cfg.threshold = hd.threshold;
        cfg.rescore_ratio = hd.rescore_ratio; cfg.device = device; cfg.capacity_rows = hd.rows; cfg.row_base = hd.row_base;
        gvdb_index* h = nullptr;
        if (gvdb_create(&cfg, &h) != GVDB_OK) fail(GVDB_ERR_INDEX, std::string("gvdb_load: ") + g_err);
        std::unique_ptr<gvdb_index, void (*)(gvdb_index*)> owner(h, gvdb_destroy);
        DeviceGuard dg(device);
        const uint64_t N = hd.rows;
        const size_t stage_bytes = 32u << 20;
        void* stage = nullptr;
        CU(cudaMallocHost(&stage, stage_bytes));
        struct Free { void* p; ~Free() { cudaFreeHost(p); } } guard{stage};
        SectionHash hash;
        if (N) {
            const uint64_t chunk = stage_bytes / (uint64_t)h->nbytes;
            DevBuf tmp; tmp.ensure(std::min(chunk, N) * h->nbytes);
            uint64_t readn = 0;
            for (uint64_t i0 = 0; i0 < N; i0 += chunk) {
                const uint64_t m = std::min(chunk, N - i0);
                file_to_dev(file.f, tmp.p, m * h->nbytes, m * h->nbytes, stage, stage_bytes, hash);
                block_codes_kernel<<<(unsigned)((m + 255) / 256), 256>>>(tmp.as<uint8_t>(), h->nchunk, i0, m, h->nbytes, h->codes);
                CU(cudaGetLastError());
                CU(cudaDeviceSynchronize());
                readn += m * h->nbytes;
            }
            tmp.release();
            if (pad64(readn) > readn && fseek(file.f, (long)(pad64(readn) - readn), SEEK_CUR) != 0)
                fail(GVDB_ERR_INDEX, "gvdb_load: truncated file");
            file_to_dev(file.f, h->norms, N * 4, pad64(N * 4), stage, stage_bytes, hash);
            file_to_dev(file.f, h->live, tiles_for(N) * 4, pad64(tiles_for(N) * 4), stage, stage_bytes, hash);
            file_to_dev(file.f, h->rows, N * (uint64_t)h->dim * 4, N * (uint64_t)h->dim * 4, stage, stage_bytes, hash);
        }
        // files written before the checksum existed carry flags = 0 and are taken as they are
        if ((hd.flags & FILE_FLAG_CHECKSUM) && hash.digest() != hd.checksum)
            fail(GVDB_ERR_INDEX, "gvdb_load: checksum mismatch (the file is corrupt or was cut short)");
        h->n_rows = N;
        // the bitmap is the truth: bits past the last row are cleared and the live count is taken from it
        uint64_t live_rows = 0;
        if (N) {
            const uint64_t nt = tiles_for(N);
            std::vector<uint32_t> bits(nt);
            CU(cudaMemcpy(bits.data(), h->live, nt * 4, cudaMemcpyDeviceToHost));
            if (N % 32) {
                const uint32_t keep = (1u << (N % 32)) - 1u;
                if (bits[nt - 1] & ~keep) {
                    bits[nt - 1] &= keep;
                    CU(cudaMemcpy(h->live + (nt - 1), &bits[nt - 1], 4, cudaMemcpyHostToDevice));
                }
            }
            for (uint32_t w : bits) live_rows += (uint64_t)__builtin_popcount(w);
        }
        if (live_rows != hd.live_rows) fail(GVDB_ERR_INDEX, "gvdb_load: the header's live row count disagrees with the tombstone bitmap");
        h->n_live = live_rows;
        *out = owner.release();
    });
}

uint64_t gvdb_rescore_count(uint64_t n, float ratio) {
    // (candidates.len() as f32 * rescore_ratio) as usize, then .min(len)  — quantization.rs:178-179
    float p = (float)n * ratio;
    uint64_t r;
    if (!(p > 0.0f)) r = 0;
    else if (p >= 18446744073709551616.0f) r = UINT64_MAX;
    else r = (uint64_t)p;
    return r < n ? r : n;
}

gvdb_status gvdb_quantize(gvdb_index* h, const float* x, uint64_t n, uint8_t* codes_out) {
    return guarded([&] {
        need(h, "index");
        if (n == 0) return;
        need(x, "x"); need(codes_out, "codes_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, nullptr, false);
        Workspace* ws = lease.ws;
        cudaStream_t st = lease.stream;
        const uint64_t chunk = 65536;
        ws->q_in.ensure(chunk * (size_t)h->dim * 4);
        ws->codes_tmp.ensure(tiles_for(chunk) * h->nchunk * 512 + chunk * 4 + tiles_for(chunk) * 4 +
                             chunk * (size_t)h->nbytes);
        uint8_t* base = ws->codes_tmp.as<uint8_t>();
        uint4* codes = reinterpret_cast<uint4*>(base);
        float* norms = reinterpret_cast<float*>(base + tiles_for(chunk) * h->nchunk * 512);
        uint32_t* live = reinterpret_cast<uint32_t*>(norms + chunk);
        uint8_t* flat = reinterpret_cast<uint8_t*>(live + tiles_for(chunk));
        for (uint64_t i0 = 0; i0 < n; i0 += chunk) {
            uint64_t m = std::min(chunk, n - i0);
            CU(cudaMemcpyAsync(ws->q_in.p, x + i0 * h->dim, m * (size_t)h->dim * 4, cudaMemcpyHostToDevice, st));
            CU(cudaMemsetAsync(live, 0, tiles_for(chunk) * 4, st));
            ingest_kernel<false><<<(unsigned)((m + STAGE_ROWS - 1) / STAGE_ROWS), STAGE_ROWS, 0, st>>>(
                ws->q_in.as<float>(), m, h->dim, h->cfg.threshold, h->nchunk, 0, codes, norms, live, nullptr, 0);
            unblock_codes_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(codes, h->nchunk, 0, m, h->nbytes, flat);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(codes_out + i0 * h->nbytes, flat, m * (size_t)h->nbytes, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
    });
}

gvdb_status gvdb_get_codes(gvdb_index* h, uint64_t first, uint64_t n, uint8_t* codes_out) {
    return guarded([&] {
        need(h, "index");
        if (n == 0) return;
        need(codes_out, "codes_out");
        if (first + n > h->n_rows) fail(GVDB_ERR_INVALID_ARGUMENT, "row range out of bounds");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, nullptr, false);
        Workspace* ws = lease.ws;
        cudaStream_t st = lease.stream;
        const uint64_t chunk = 1 << 20;
        ws->codes_tmp.ensure(std::min(chunk, n) * (size_t)h->nbytes);
        for (uint64_t i0 = 0; i0 < n; i0 += chunk) {
            uint64_t m = std::min(chunk, n - i0);
            unblock_codes_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(
                h->codes, h->nchunk, first + i0, m, h->nbytes, ws->codes_tmp.as<uint8_t>());
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(codes_out + i0 * h->nbytes, ws->codes_tmp.p, m * (size_t)h->nbytes, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
    });
}

gvdb_status gvdb_hamming(gvdb_index* h, const uint8_t* q_codes, uint32_t nq, uint32_t* dist_out) {
    return guarded([&] {
        need(h, "index");
        if (nq == 0 || h->n_rows == 0) return;
        need(q_codes, "q_codes"); need(dist_out, "dist_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, nullptr, false);
        Workspace* ws = lease.ws;
        cudaStream_t st = lease.stream;
        const uint64_t N = h->n_rows;
        uint32_t qchunk = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(nq, (256ull << 20) / (N * 4)));
        qchunk = std::min<uint32_t>(qchunk, 1024);
        ws->q_in.ensure((size_t)qchunk * h->nbytes);
        ws->qpack.ensure((size_t)qchunk * h->qs * 4);
        ws->misc.ensure((size_t)qchunk * N * 4);
        const uint32_t ntiles = (uint32_t)tiles_for(N);
        for (uint32_t q0 = 0; q0 < nq; q0 += qchunk) {
            uint32_t m = std::min(qchunk, nq - q0);
            CU(cudaMemcpyAsync(ws->q_in.p, q_codes + (size_t)q0 * h->nbytes, (size_t)m * h->nbytes, cudaMemcpyHostToDevice, st));
            pack_query_codes_kernel<<<m, 64, 0, st>>>(ws->q_in.as<uint8_t>(), m, h->nbytes, h->nchunk, ws->qpack.as<uint32_t>(), h->qs);
            CU(cudaGetLastError());
            if (m >= h->tc_min_q && tc_supported(h->nchunk)) {     // same kernel the batched search uses
                const uint32_t m_pad = (m + TC_NQ - 1) / TC_NQ * TC_NQ;
                tc_prepare_queries(h, ws, st, m, m_pad);
                tc_update_bias(h, ws, st, m, m_pad, 1);
                launch_tc_scan<1>(h, ws, st, 0, ntiles, (ntiles + 3) / 4, 1, m, m_pad, nullptr, nullptr, 0, nullptr,
                                  ws->misc.as<uint32_t>(), N);
            } else {
                dim3 grid = scan_grid(h, ntiles, m);
                launch_scan<1>(h->nchunk, -1, st, grid, (size_t)kQGroup * h->qs * 4, h->codes, h->live, 0, ntiles,
                               ws->qpack.as<uint32_t>(), (int)m, kQGroup, nullptr, nullptr, 0, nullptr,
                               ws->misc.as<uint32_t>(), N, N);
            }
            CU(cudaMemcpyAsync(dist_out + (size_t)q0 * N, ws->misc.p, (size_t)m * N * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
    });
}

gvdb_status gvdb_search_batch_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                                     uint32_t k, uint32_t rescore_count, uint64_t* ids_out_dev,
                                     float* scores_out_dev, uint64_t* cand_ids_out_dev,
                                     uint32_t* cand_ham_out_dev) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_search_batch_device");
        need(h, "index");
        if (nq == 0) return;
        need(queries_dev, "queries"); need(ids_out_dev, "ids_out"); need(scores_out_dev, "scores_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        search_device(h, lease.ws, lease.stream, queries_dev, nq, k, rescore_count, ids_out_dev,
                      scores_out_dev, cand_ids_out_dev, cand_ham_out_dev);
    });
}

gvdb_status gvdb_search_batch(gvdb_index* h, const float* queries, uint32_t nq, uint32_t k,
                              uint32_t rescore_count, uint64_t* ids_out, float* scores_out,
                              uint64_t* cand_ids_out, uint32_t* cand_ham_out) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_search_batch");
        need(h, "index");
        if (nq == 0) return;
        need(queries, "queries"); need(ids_out, "ids_out"); need(scores_out, "scores_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, nullptr, false);
        Workspace* ws = lease.ws;
        cudaStream_t st = lease.stream;
        const uint32_t R = rescore_count;
        ws->q_in.ensure((size_t)nq * h->dim * 4);
        ws->ids_out.ensure((size_t)nq * std::max(k, 1u) * 8);
        ws->sc_out.ensure((size_t)nq * std::max(k, 1u) * 4);
        CU(cudaMemcpyAsync(ws->q_in.p, queries, (size_t)nq * h->dim * 4, cudaMemcpyHostToDevice, st));
        // the cut-by-counting path (large R, or the overflow fallback) keeps one query's records at a
        // time: give the call full candidate buffers whenever candidates are asked for
        uint64_t* cand_ids_dev = nullptr;
        uint32_t* cand_ham_dev = nullptr;
        if (cand_ids_out || cand_ham_out) {
            ws->codes_tmp.ensure((size_t)nq * R * 12);
            cand_ids_dev = ws->codes_tmp.as<uint64_t>();
            cand_ham_dev = reinterpret_cast<uint32_t*>(ws->codes_tmp.as<uint8_t>() + (size_t)nq * R * 8);
        }
        // results land in a pinned staging buffer before the call's single synchronisation
        const size_t res_bytes = (size_t)nq * k * 12;
        if (ws->h_res_bytes < res_bytes) {
            if (ws->h_res) cudaFreeHost(ws->h_res);
            ws->h_res = nullptr; ws->h_res_bytes = 0;
            CU(cudaMallocHost(&ws->h_res, res_bytes));
            ws->h_res_bytes = res_bytes;
        }
        uint64_t* h_ids = static_cast<uint64_t*>(ws->h_res);
        float* h_sc = reinterpret_cast<float*>(static_cast<uint8_t*>(ws->h_res) + (size_t)nq * k * 8);
        search_device(h, ws, st, ws->q_in.as<float>(), nq, k, R, ws->ids_out.as<uint64_t>(),
                      ws->sc_out.as<float>(), cand_ids_dev, cand_ham_dev, h_ids, h_sc);
        memcpy(ids_out, h_ids, (size_t)nq * k * 8);
        memcpy(scores_out, h_sc, (size_t)nq * k * 4);
        if (!cand_ids_out && !cand_ham_out) return;
        if (cand_ids_out) CU(cudaMemcpyAsync(cand_ids_out, cand_ids_dev, (size_t)nq * R * 8, cudaMemcpyDeviceToHost, st));
        if (cand_ham_out) CU(cudaMemcpyAsync(cand_ham_out, cand_ham_dev, (size_t)nq * R * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    });
}

gvdb_status gvdb_flat_search_batch_device(gvdb_index* h, void* stream, const float* queries_dev,
                                          uint32_t nq, uint32_t k, uint64_t* ids_out_dev,
                                          float* dist_out_dev) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_flat_search_batch_device");
        need(h, "index");
        if (nq == 0) return;
        need(queries_dev, "queries"); need(ids_out_dev, "ids_out"); need(dist_out_dev, "dist_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        flat_device(h, lease.ws, lease.stream, queries_dev, nq, k, ids_out_dev, dist_out_dev);
    });
}

gvdb_status gvdb_flat_search_batch(gvdb_index* h, const float* queries, uint32_t nq, uint32_t k,
                                   uint64_t* ids_out, float* dist_out) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_flat_search_batch");
        need(h, "index");
        if (nq == 0) return;
        need(queries, "queries"); need(ids_out, "ids_out"); need(dist_out, "dist_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, nullptr, false);
        Workspace* ws = lease.ws;
        cudaStream_t st = lease.stream;
        ws->q_in.ensure((size_t)nq * h->dim * 4);
        ws->ids_out.ensure((size_t)nq * std::max(k, 1u) * 8);
        ws->sc_out.ensure((size_t)nq * std::max(k, 1u) * 4);
        CU(cudaMemcpyAsync(ws->q_in.p, queries, (size_t)nq * h->dim * 4, cudaMemcpyHostToDevice, st));
        flat_device(h, ws, st, ws->q_in.as<float>(), nq, k, ws->ids_out.as<uint64_t>(), ws->sc_out.as<float>());
        CU(cudaMemcpyAsync(ids_out, ws->ids_out.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(dist_out, ws->sc_out.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    });
}

// ---- filtered search: a per-call allow-list ANDed with the tombstones ------------------------------------
namespace {
// Sets the workspace's effective row bitmap for the lifetime of the scope.  The allow-list covers the
// stored rows: bit (r % 32) of word r / 32 = 1 when local row r may be returned (the layout of the
// tombstone bitmap); ceil(rows / 32) words.
struct FilterScope {
    Workspace* ws;
    FilterScope(gvdb_index* h, Workspace* ws_, cudaStream_t st, const uint32_t* allow_dev) : ws(ws_) {
        const size_t words = tiles_for(h->n_rows);
        ws->filt.ensure(std::max<size_t>(4, tiles_for(h->cap_rows) * 4));
        if (words) {
            and_bitmap_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(h->live, allow_dev, words, ws->filt.as<uint32_t>());
            CU(cudaGetLastError());
        }
        ws->live_eff = ws->filt.as<uint32_t>();
    }
    ~FilterScope() { ws->live_eff = nullptr; }
};
}  // namespace

gvdb_status gvdb_search_batch_filtered_device(gvdb_index* h, void* stream, const float* queries_dev,
                                              const uint32_t* allow_bits_dev, uint32_t nq, uint32_t k,
                                              uint32_t rescore_count, uint64_t* ids_out_dev, float* scores_out_dev) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_search_batch_filtered_device");
        need(h, "index");
        if (nq == 0) return;
        need(queries_dev, "queries"); need(allow_bits_dev, "allow_bits"); need(ids_out_dev, "ids_out"); need(scores_out_dev, "scores_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        FilterScope fs(h, lease.ws, lease.stream, allow_bits_dev);
        search_device(h, lease.ws, lease.stream, queries_dev, nq, k, rescore_count, ids_out_dev, scores_out_dev, nullptr, nullptr);
    });
}

gvdb_status gvdb_search_batch_filtered(gvdb_index* h, const float* queries, const uint32_t* allow_bits, uint32_t nq,
                                       uint32_t k, uint32_t rescore_count, uint64_t* ids_out, float* scores_out) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_search_batch_filtered");
        need(h, "index");
        if (nq == 0) return;
        need(queries, "queries"); need(allow_bits, "allow_bits"); need(ids_out, "ids_out"); need(scores_out, "scores_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, nullptr, false);
        Workspace* ws = lease.ws;
        cudaStream_t st = lease.stream;
        const size_t words = tiles_for(h->n_rows);
        ws->q_in.ensure((size_t)nq * h->dim * 4);
        ws->ids_out.ensure((size_t)nq * std::max(k, 1u) * 8);
        ws->sc_out.ensure((size_t)nq * std::max(k, 1u) * 4);
        ws->allow_in.ensure(std::max<size_t>(4, words * 4));
        CU(cudaMemcpyAsync(ws->q_in.p, queries, (size_t)nq * h->dim * 4, cudaMemcpyHostToDevice, st));
        if (words) CU(cudaMemcpyAsync(ws->allow_in.p, allow_bits, words * 4, cudaMemcpyHostToDevice, st));
        FilterScope fs(h, ws, st, ws->allow_in.as<uint32_t>());
        search_device(h, ws, st, ws->q_in.as<float>(), nq, k, rescore_count, ws->ids_out.as<uint64_t>(), ws->sc_out.as<float>(),
                      nullptr, nullptr);
        CU(cudaMemcpyAsync(ids_out, ws->ids_out.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(scores_out, ws->sc_out.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    });
}

gvdb_status gvdb_flat_search_batch_filtered(gvdb_index* h, const float* queries, const uint32_t* allow_bits, uint32_t nq,
                                            uint32_t k, uint64_t* ids_out, float* dist_out) {
    return guarded([&] {
        need(h, "index");
        if (nq == 0) return;
        need(queries, "queries"); need(allow_bits, "allow_bits"); need(ids_out, "ids_out"); need(dist_out, "dist_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, nullptr, false);
        Workspace* ws = lease.ws;
        cudaStream_t st = lease.stream;
        const size_t words = tiles_for(h->n_rows);
        ws->q_in.ensure((size_t)nq * h->dim * 4);
        ws->ids_out.ensure((size_t)nq * std::max(k, 1u) * 8);
        ws->sc_out.ensure((size_t)nq * std::max(k, 1u) * 4);
        ws->allow_in.ensure(std::max<size_t>(4, words * 4));
        CU(cudaMemcpyAsync(ws->q_in.p, queries, (size_t)nq * h->dim * 4, cudaMemcpyHostToDevice, st));
        if (words) CU(cudaMemcpyAsync(ws->allow_in.p, allow_bits, words * 4, cudaMemcpyHostToDevice, st));
        FilterScope fs(h, ws, st, ws->allow_in.as<uint32_t>());
        flat_device(h, ws, st, ws->q_in.as<float>(), nq, k, ws->ids_out.as<uint64_t>(), ws->sc_out.as<float>());
        CU(cudaMemcpyAsync(ids_out, ws->ids_out.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(dist_out, ws->sc_out.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    });
}

gvdb_status gvdb_similarity_search_batch_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                                                uint32_t k, float threshold, int32_t use_threshold,
                                                uint64_t* ids_out_dev, float* sims_out_dev) {
    return guarded([&] {
        need(h, "index");
        if (nq == 0) return;
        need(queries_dev, "queries"); need(ids_out_dev, "ids_out"); need(sims_out_dev, "sims_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        flat_device(h, lease.ws, lease.stream, queries_dev, nq, k, ids_out_dev, sims_out_dev, true, threshold,
                    use_threshold != 0);
    });
}

gvdb_status gvdb_similarity_search_batch(gvdb_index* h, const float* queries, uint32_t nq, uint32_t k,
                                         float threshold, int32_t use_threshold, uint64_t* ids_out, float* sims_out) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_similarity_search_batch");
        need(h, "index");
        if (nq == 0) return;
        need(queries, "queries"); need(ids_out, "ids_out"); need(sims_out, "sims_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, nullptr, false);
        Workspace* ws = lease.ws;
        cudaStream_t st = lease.stream;
        ws->q_in.ensure((size_t)nq * h->dim * 4);
        ws->ids_out.ensure((size_t)nq * std::max(k, 1u) * 8);
        ws->sc_out.ensure((size_t)nq * std::max(k, 1u) * 4);
        CU(cudaMemcpyAsync(ws->q_in.p, queries, (size_t)nq * h->dim * 4, cudaMemcpyHostToDevice, st));
        flat_device(h, ws, st, ws->q_in.as<float>(), nq, k, ws->ids_out.as<uint64_t>(), ws->sc_out.as<float>(), true,
                    threshold, use_threshold != 0);
        CU(cudaMemcpyAsync(ids_out, ws->ids_out.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(sims_out, ws->sc_out.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    });
}

uint64_t gvdb_shard_record_bytes(uint32_t nq, uint32_t rescore_count) {
    return (uint64_t)nq * rescore_count * 16ull;
}

gvdb_status gvdb_search_shard_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                                     uint32_t rescore_count, void* records_dev) {
    return guarded([&] {
        need(h, "index");
        if (nq == 0) return;
        need(queries_dev, "queries"); need(records_dev, "records");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        const uint64_t nr = (uint64_t)nq * rescore_count;
        uint8_t* base = static_cast<uint8_t*>(records_dev);
        for (int attempt = 0; attempt < 2; ++attempt) {
            bool optimistic = false;
            search_core(h, lease.ws, lease.stream, queries_dev, nq, rescore_count,
                        reinterpret_cast<uint32_t*>(base + nr * 8), reinterpret_cast<uint64_t*>(base),
                        reinterpret_cast<float*>(base + nr * 12), true, attempt == 0, &optimistic);
            if (!check_overflow(h, lease.ws, lease.stream, optimistic)) break;
        }
    });
}

gvdb_status gvdb_search_shard_sliced_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                                            uint32_t rescore_count, uint32_t n_slices, void* records_dev) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_search_shard_sliced_device");
        need(h, "index");
        if (nq == 0) return;
        need(queries_dev, "queries"); need(records_dev, "records");
        if (n_slices == 0 || nq % n_slices != 0)
            fail(GVDB_ERR_INVALID_ARGUMENT, "nq must be a positive multiple of n_slices");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        const uint32_t per = nq / n_slices;
        const uint64_t nr = (uint64_t)per * rescore_count;
        if (rescore_count >= 1 && rescore_count <= kMaxR && rescore_topk_fits(h, rescore_count) && h->query_tile % per == 0) {
            // one pass over the shard for ALL queries; the rescoring kernel writes the slice-packed layout
            for (int attempt = 0; attempt < 2; ++attempt) {
                bool optimistic = false;
                FusedTopk fused;
                fused.slice_q = per;
                search_core(h, lease.ws, lease.stream, queries_dev, nq, rescore_count, nullptr,
                            static_cast<uint64_t*>(records_dev), nullptr, true, attempt == 0, &optimistic, nullptr, &fused);
                if (!check_overflow(h, lease.ws, lease.stream, optimistic)) break;
            }
            return;
        }
        for (int attempt = 0; attempt < 2; ++attempt) {
            bool optimistic = false;
            for (uint32_t s = 0; s < n_slices; ++s) {
                uint8_t* base = static_cast<uint8_t*>(records_dev) + (uint64_t)s * nr * 16;
                search_core(h, lease.ws, lease.stream, queries_dev + (size_t)s * per * h->dim, per, rescore_count,
                            reinterpret_cast<uint32_t*>(base + nr * 8), reinterpret_cast<uint64_t*>(base),
                            reinterpret_cast<float*>(base + nr * 12), s == 0, attempt == 0, &optimistic);
            }
            if (!check_overflow(h, lease.ws, lease.stream, optimistic)) break;
        }
    });
}

// The sliced shard search WITHOUT its host synchronisation: the single pass is enqueued on `stream` together with
// a copy of its verdict flags into pinned memory, and the call returns — the caller can queue the exchange and the
// merge behind it at once.  gvdb_search_shard_verify (after the step's other work was enqueued) waits for the
// stream and says whether the device accepted the pass; if not, the step must be repeated with the synchronous
// entry point.  One enqueue in flight per index.
gvdb_status gvdb_search_shard_sliced_enqueue_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                                                    uint32_t rescore_count, uint32_t n_slices, void* records_dev,
                                                    uint32_t* verdict_out_dev) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_search_shard_sliced_enqueue_device");
        need(h, "index");
        if (nq == 0) return;
        need(queries_dev, "queries"); need(records_dev, "records");
        if (n_slices == 0 || nq % n_slices != 0)
            fail(GVDB_ERR_INVALID_ARGUMENT, "nq must be a positive multiple of n_slices");
        const uint32_t per = nq / n_slices;
        if (!(rescore_count >= 1 && rescore_count <= kMaxR && rescore_topk_fits(h, rescore_count) && h->query_tile % per == 0))
            fail(GVDB_ERR_NOT_IMPLEMENTED, "the enqueue form needs dim % 4 == 0, rescore_count <= 256 and slices that divide the query tile");
        DeviceGuard dg(h->cfg.device);
        if (!h->h_async_flag) CU(cudaMallocHost((void**)&h->h_async_flag, 64));
        WsLease lease(h, (cudaStream_t)stream, true);
        bool optimistic = false;
        FusedTopk fused;
        fused.slice_q = per;
        search_core(h, lease.ws, lease.stream, queries_dev, nq, rescore_count, nullptr, static_cast<uint64_t*>(records_dev), nullptr,
                    true, true, &optimistic, nullptr, &fused);
        CU(cudaMemcpyAsync(h->h_async_flag, lease.ws->flag.p, 8, cudaMemcpyDeviceToHost, lease.stream));
        if (verdict_out_dev)
            CU(cudaMemcpyAsync(verdict_out_dev, lease.ws->flag.p, 8, cudaMemcpyDeviceToDevice, lease.stream));
        h->async_pending = true;
        if (h->profile_on.load(std::memory_order_relaxed) != 0) {   // per-launch timing wants the events resolved
            CU(cudaStreamSynchronize(lease.stream));
            flush_profile(h, lease.ws);
        }
    });
}

gvdb_status gvdb_search_shard_verify(gvdb_index* h, void* stream, int32_t* rerun_out) {
    return guarded([&] {
        need(h, "index"); need(rerun_out, "rerun_out");
        *rerun_out = 0;
        if (!h->async_pending) return;
        DeviceGuard dg(h->cfg.device);
        CU(cudaStreamSynchronize((cudaStream_t)stream));
        h->async_pending = false;
        if (h->h_async_flag[0] || h->h_async_flag[1]) {
            *rerun_out = 1;
            h->optimistic_reruns.fetch_add(1, std::memory_order_relaxed);
        }
    });
}

gvdb_status gvdb_merge_shards_device(gvdb_index* h, void* stream, uint32_t n_shards,
                                     const void* records_dev, uint32_t nq, uint32_t rescore_count,
                                     uint32_t k, uint64_t* ids_out_dev, float* scores_out_dev) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_merge_shards_device");
        need(h, "index");
        if (nq == 0) return;
        need(records_dev, "records"); need(ids_out_dev, "ids_out"); need(scores_out_dev, "scores_out");
        const uint32_t R = rescore_count;
        if (R == 0 || R > kMaxR) fail(GVDB_ERR_INVALID_ARGUMENT, "rescore_count must be in [1, 2048]");
        if (k > R) fail(GVDB_ERR_INVALID_ARGUMENT, "k must be <= rescore_count");
        if (n_shards == 0) fail(GVDB_ERR_INVALID_ARGUMENT, "n_shards must be >= 1");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        Workspace* ws = lease.ws;
        cudaStream_t st = lease.stream;
        ws->rec_ham.ensure((size_t)nq * R * 4);
        ws->rec_ids.ensure((size_t)nq * R * 8);
        ws->rec_score.ensure((size_t)nq * R * 4);
        ShardRecords rec{static_cast<const uint8_t*>(records_dev), gvdb_shard_record_bytes(nq, R), nq, R, R};
        {
            Timed t(h, ws, st, K_MERGE);
            merge_select_kernel<<<nq, SORT_THREADS, SORT_N * 8, st>>>(
                n_shards, rec, ws->rec_ids.as<uint64_t>(), ws->rec_score.as<float>(),
                ws->rec_ham.as<uint32_t>());
        }
        CU(cudaGetLastError());
        launch_topk(h, ws, st, ws->rec_ids.as<uint64_t>(), ws->rec_score.as<float>(), nq, R, k, ids_out_dev, scores_out_dev);
        finish_async(h, ws, st);
    });
}

uint32_t gvdb_shard_hist_bins(const gvdb_index* h) { return h ? (uint32_t)h->nchunk * 128 + 1 : 0; }

gvdb_status gvdb_shard_hist_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                                   uint32_t* hist_out_dev) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_shard_hist_device");
        need(h, "index");
        if (nq == 0) return;
        need(queries_dev, "queries"); need(hist_out_dev, "hist_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        shard_hist(h, lease.ws, lease.stream, queries_dev, nq, hist_out_dev);
        finish_async(h, lease.ws, lease.stream);
    });
}

gvdb_status gvdb_search_shard_ratio_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                                           uint64_t rescore_count, uint32_t k, const uint32_t* hists_all_dev,
                                           uint32_t n_shards, uint32_t my_shard, void* records_dev) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_search_shard_ratio_device");
        need(h, "index");
        if (nq == 0) return;
        need(queries_dev, "queries"); need(hists_all_dev, "hists_all"); need(records_dev, "records");
        if (rescore_count == 0) fail(GVDB_ERR_INVALID_ARGUMENT, "rescore_count must be >= 1");
        if (k == 0 || k > 1024) fail(GVDB_ERR_INVALID_ARGUMENT, "k must be in [1, 1024]");
        if (n_shards == 0 || my_shard >= n_shards) fail(GVDB_ERR_INVALID_ARGUMENT, "my_shard must be < n_shards");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        search_shard_ratio(h, lease.ws, lease.stream, queries_dev, nq, rescore_count, k, hists_all_dev, n_shards, my_shard,
                           static_cast<uint8_t*>(records_dev));
    });
}

gvdb_status gvdb_merge_shards_ratio_device(gvdb_index* h, void* stream, uint32_t n_shards, const void* records_dev,
                                           uint32_t nq, uint32_t k, uint64_t* ids_out_dev, float* scores_out_dev) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_merge_shards_ratio_device");
        need(h, "index");
        if (nq == 0) return;
        need(records_dev, "records"); need(ids_out_dev, "ids_out"); need(scores_out_dev, "scores_out");
        if (k == 0 || n_shards == 0 || (uint64_t)n_shards * k > (uint64_t)SORT_N)
            fail(GVDB_ERR_INVALID_ARGUMENT, "n_shards * k must be in [1, 4096]");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        Workspace* ws = lease.ws;
        cudaStream_t st = lease.stream;
        const uint32_t keep = n_shards * k;
        ws->rec_ham.ensure((size_t)nq * keep * 4);
        ws->rec_ids.ensure((size_t)nq * keep * 8);
        ws->rec_score.ensure((size_t)nq * keep * 4);
        // every record stays (the shards applied the global cut): order by (hamming, row), then by cosine
        ShardRecords rec{static_cast<const uint8_t*>(records_dev), gvdb_shard_record_bytes(nq, k), nq, k, keep};
        {
            Timed t(h, ws, st, K_MERGE);
            merge_select_kernel<<<nq, SORT_THREADS, SORT_N * 8, st>>>(
                n_shards, rec, ws->rec_ids.as<uint64_t>(), ws->rec_score.as<float>(), ws->rec_ham.as<uint32_t>());
        }
        CU(cudaGetLastError());
        launch_topk(h, ws, st, ws->rec_ids.as<uint64_t>(), ws->rec_score.as<float>(), nq, keep, k, ids_out_dev, scores_out_dev);
        finish_async(h, ws, st);
    });
}

gvdb_status gvdb_stage1_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                               uint32_t rescore_count, uint64_t* keys_out_dev) {
    return guarded([&] {
        need(h, "index");
        if (nq == 0) return;
        need(queries_dev, "queries"); need(keys_out_dev, "keys_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        for (int attempt = 0; attempt < 2; ++attempt) {
            bool optimistic = false;
            search_core(h, lease.ws, lease.stream, queries_dev, nq, rescore_count, nullptr, nullptr, nullptr, true,
                        attempt == 0, &optimistic, keys_out_dev);
            if (!check_overflow(h, lease.ws, lease.stream, optimistic)) break;
        }
    });
}

namespace {
// Owner-side rescoring of `nq` queries' candidate keys (all asynchronous on st).
void rescore_keys_core(gvdb_index* h, Workspace* ws, cudaStream_t st, const float* queries_dev, uint32_t nq,
                       uint32_t R, const uint64_t* keys_dev, float* scores_out_dev, uint8_t* const* peers = nullptr,
                       uint64_t peer_off = 0, uint32_t pairs_per_peer = 1, OwnedSignal sg = OwnedSignal{0, 0, 0, 0, 0, 0, nullptr}) {
    if ((h->dim & 3) != 0) fail(GVDB_ERR_NOT_IMPLEMENTED, "owner-side rescoring needs dim % 4 == 0");
    const int cols = std::min(h->dim, RS_SLAB);
    const int stride = ((cols >> 2) & 1) ? cols : cols + 4;
    const int q_slots = (int)std::min<uint32_t>(32, 31 / R + 2);
    static std::atomic<uint64_t> attr_done_a{0}, attr_done_b{0};
    ensure_dyn_smem(attr_done_a, rescore_slab_kernel<true>, 64 * (RS_SLAB + 4) * (int)sizeof(float));
    ensure_dyn_smem(attr_done_b, rescore_owned_list_kernel, 64 * (RS_SLAB + 4) * (int)sizeof(float));
    const uint64_t lo = h->windowed ? h->win_first : 0;
    const uint64_t hi = h->windowed ? std::min(h->n_rows, h->win_first + h->win_count) : h->n_rows;
    const uint64_t pairs = (uint64_t)nq * R;
    if (pairs > 0x7fffffffull) fail(GVDB_ERR_INVALID_ARGUMENT, "nq * rescore_count too large");
    if (peers || (h->windowed && !h->rows_cover_all())) {
        // this GPU owns a fraction of the rows: list the owned pairs, then score 32 of them per warp —
        // the work is 1/G of the pairs, not 1/G of every warp
        ws->big_v32.ensure(pairs * 4);
        ws->big_aux.ensure(256);
        uint32_t* list = ws->big_v32.as<uint32_t>();
        uint32_t* count = ws->big_aux.as<uint32_t>();
        OwnedPair pred{keys_dev, h->cfg.row_base, lo, hi};
        Timed t(h, ws, st, K_RESCORE);
        CU(cudaMemsetAsync(count, 0, 4, st));
        if (!peers) CU(cudaMemsetAsync(scores_out_dev, 0, pairs * 4, st));
        owned_compact_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(pred, (uint32_t)pairs, list, count);
        (void)sg;
        static std::atomic<uint64_t> attr_done_c{0};
        ensure_dyn_smem(attr_done_c, rescore_owned_ring_kernel, (int)RO_SMEM);
        rescore_owned_ring_kernel<<<(unsigned)((pairs + 32 * RO_WARPS - 1) / (32 * RO_WARPS)), 32 * RO_WARPS, RO_SMEM, st>>>(
            h->rows_base(), h->norms, h->cfg.row_base, h->dim, queries_dev, keys_dev, list, count, R,
            scores_out_dev, peers, peer_off, pairs_per_peer);
    } else {
        Timed t(h, ws, st, K_RESCORE);
        rescore_slab_kernel<true><<<(unsigned)((pairs + 31) / 32), 32, (size_t)(32 + q_slots) * stride * sizeof(float), st>>>(
            h->rows_base(), h->norms, h->cfg.row_base, h->dim, stride, q_slots, queries_dev, nullptr, keys_dev, 0,
            nullptr, R, nq, nullptr, nullptr, scores_out_dev, lo, hi);
    }
    CU(cudaGetLastError());
}

// Each key takes its owner's score; order by (cosine desc, hamming asc, row asc); first k.
void finish_owned_core(gvdb_index* h, Workspace* ws, cudaStream_t st, const uint64_t* keys_dev,
                       const float* scores_by_owner_dev, uint32_t n_owners, uint64_t rows_per_owner, uint32_t nq,
                       uint32_t R, uint32_t k, uint64_t* ids_out_dev, float* scores_out_dev) {
    if (R == 0 || R > kMaxR) fail(GVDB_ERR_INVALID_ARGUMENT, "rescore_count must be in [1, 2048]");
    if (k > R) fail(GVDB_ERR_INVALID_ARGUMENT, "k must be <= rescore_count");
    if (n_owners == 0 || rows_per_owner == 0) fail(GVDB_ERR_INVALID_ARGUMENT, "n_owners and rows_per_owner must be >= 1");
    uint32_t n_eff = 64;
    while (n_eff < R) n_eff <<= 1;
    const uint32_t threads = std::min<uint32_t>(1024, std::max<uint32_t>(32, n_eff / 2));
    {
        Timed t(h, ws, st, K_TOPK);
        topk_owned_kernel<<<nq, threads, n_eff * 8, st>>>(keys_dev, scores_by_owner_dev, n_owners, rows_per_owner, nq, R,
                                                          n_eff, k, ids_out_dev, scores_out_dev);
    }
    CU(cudaGetLastError());
}
}  // namespace

gvdb_status gvdb_rescore_keys_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                                     uint32_t rescore_count, const uint64_t* keys_dev, float* scores_out_dev) {
    return guarded([&] {
        need(h, "index");
        if (nq == 0 || rescore_count == 0) return;
        need(queries_dev, "queries"); need(keys_dev, "keys"); need(scores_out_dev, "scores_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        rescore_keys_core(h, lease.ws, lease.stream, queries_dev, nq, rescore_count, keys_dev, scores_out_dev);
        finish_async(h, lease.ws, lease.stream);
    });
}

gvdb_status gvdb_finish_owned_device(gvdb_index* h, void* stream, const uint64_t* keys_dev,
                                     const float* scores_by_owner_dev, uint32_t n_owners, uint64_t rows_per_owner,
                                     uint32_t nq, uint32_t rescore_count, uint32_t k, uint64_t* ids_out_dev,
                                     float* scores_out_dev) {
    return guarded([&] {
        need(h, "index");
        if (nq == 0) return;
        need(keys_dev, "keys"); need(scores_by_owner_dev, "scores_by_owner");
        need(ids_out_dev, "ids_out"); need(scores_out_dev, "scores_out");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        finish_owned_core(h, lease.ws, lease.stream, keys_dev, scores_by_owner_dev, n_owners, rows_per_owner, nq,
                          rescore_count, k, ids_out_dev, scores_out_dev);
        finish_async(h, lease.ws, lease.stream);
    });
}

namespace {
void attach_peers(gvdb_index* h, uint32_t n_owners, uint64_t rows_per_owner, uint32_t my_owner,
                  const std::vector<const float*>& ptrs) {
    if (!h->windowed) fail(GVDB_ERR_INVALID_ARGUMENT, "peer rows need an index created with GVDB_FLAG_ROW_WINDOW");
    if (n_owners == 0 || rows_per_owner == 0 || my_owner >= n_owners)
        fail(GVDB_ERR_INVALID_ARGUMENT, "bad owner partition");
    if (h->win_first != (uint64_t)my_owner * rows_per_owner || h->win_count > rows_per_owner)
        fail(GVDB_ERR_INVALID_ARGUMENT, "this index's row window is not owner my_owner's share");
    if (rows_per_owner > 0xffffffffull) fail(GVDB_ERR_INVALID_ARGUMENT, "rows_per_owner too large");
    if (h->peer_rows_dev) { cudaFree((void*)h->peer_rows_dev); h->peer_rows_dev = nullptr; }
    CU(cudaMalloc((void**)&h->peer_rows_dev, n_owners * sizeof(float*)));
    CU(cudaMemcpy((void*)h->peer_rows_dev, ptrs.data(), n_owners * sizeof(float*), cudaMemcpyHostToDevice));
    h->n_peers = n_owners;
    h->peer_per = rows_per_owner;
}
}  // namespace

gvdb_status gvdb_export_rows_ipc(gvdb_index* h, uint8_t* handle_out) {
    return guarded([&] {
        need(h, "index"); need(handle_out, "handle_out");
        static_assert(sizeof(cudaIpcMemHandle_t) == GVDB_IPC_HANDLE_BYTES, "IPC handle size");
        if (!h->rows) fail(GVDB_ERR_INDEX_NOT_BUILT, "no f32 rows on this index");
        DeviceGuard dg(h->cfg.device);
        cudaIpcMemHandle_t hd;
        CU(cudaIpcGetMemHandle(&hd, h->rows));
        memcpy(handle_out, &hd, sizeof(hd));
    });
}

gvdb_status gvdb_attach_peer_rows_ipc(gvdb_index* h, uint32_t n_owners, uint64_t rows_per_owner, uint32_t my_owner,
                                      const uint8_t* handles) {
    return guarded([&] {
        need(h, "index"); need(handles, "handles");
        DeviceGuard dg(h->cfg.device);
        std::vector<const float*> ptrs(n_owners, nullptr);
        for (uint32_t o = 0; o < n_owners; ++o) {
            if (o == my_owner) { ptrs[o] = h->rows; continue; }
            cudaIpcMemHandle_t hd;
            memcpy(&hd, handles + (size_t)o * GVDB_IPC_HANDLE_BYTES, sizeof(hd));
            void* p = nullptr;
            CU(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
            h->ipc_opened.push_back(p);
            ptrs[o] = static_cast<const float*>(p);
        }
        attach_peers(h, n_owners, rows_per_owner, my_owner, ptrs);
    });
}

const void* gvdb_rows_device_ptr(const gvdb_index* h) { return h ? h->rows : nullptr; }

gvdb_status gvdb_attach_peer_rows_ptr(gvdb_index* h, uint32_t n_owners, uint64_t rows_per_owner, uint32_t my_owner,
                                      const void* const* row_ptrs) {
    return guarded([&] {
        need(h, "index"); need(row_ptrs, "row_ptrs");
        DeviceGuard dg(h->cfg.device);
        std::vector<const float*> ptrs(n_owners, nullptr);
        for (uint32_t o = 0; o < n_owners; ++o) ptrs[o] = o == my_owner ? h->rows : static_cast<const float*>(row_ptrs[o]);
        attach_peers(h, n_owners, rows_per_owner, my_owner, ptrs);
    });
}

// ---- peer exchange (gvdb_xchg.cuh) -------------------------------------------------------------------
namespace {
size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// push + signal(kind) in one launch: the last block to finish publishes the flag (gvdb_xchg.cuh)
void xchg_push(gvdb_index* h, Workspace* ws, cudaStream_t st, Exchange* x, uint64_t dst_off, const void* src,
               uint64_t src_stride, uint64_t bytes, uint32_t kind) {
    if (bytes == 0) return;
    const XchgSignal sg{x->peers_dev, x->flags_off, kind, x->world, x->rank, x->epoch, x->done.as<uint32_t>() + kind * 32};
    const bool wide = ((dst_off | (uint64_t)(uintptr_t)src | src_stride | bytes) & 15) == 0;
    const uint64_t units = bytes / (wide ? 16 : 4);
    const unsigned gx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((units + 255) / 256, 4 * (uint64_t)h->sm_count / x->world + 1));
    Timed t(h, ws, st, K_XCHG);
    if (wide)
        xchg_push_kernel<uint4><<<dim3(gx, x->world), 256, 0, st>>>(x->peers_dev, dst_off, static_cast<const uint8_t*>(src), src_stride, bytes, sg);
    else
        xchg_push_kernel<uint32_t><<<dim3(gx, x->world), 256, 0, st>>>(x->peers_dev, dst_off, static_cast<const uint8_t*>(src), src_stride, bytes, sg);
}
void xchg_signal(gvdb_index* h, Workspace* ws, cudaStream_t st, Exchange* x, uint32_t kind) {
    Timed t(h, ws, st, K_XCHG);
    xchg_signal_kernel<<<1, 32, 0, st>>>(x->peers_dev, x->flags_off, kind, x->world, x->rank, x->epoch);
}
void xchg_wait(gvdb_index* h, Workspace* ws, cudaStream_t st, Exchange* x, uint32_t kind_mask) {
    Timed t(h, ws, st, K_XCHG_WAIT);
    xchg_wait_kernel<<<1, XCHG_KINDS * XCHG_MAX_WORLD, 0, st>>>(
        reinterpret_cast<const uint32_t*>(x->mailbox + x->flags_off), kind_mask, x->world, x->epoch, x->limit_ns,
        x->err.as<uint32_t>());
}
}  // namespace

gvdb_status gvdb_exchange_create(gvdb_index* h, uint32_t world, uint32_t rank, uint64_t rows_per_owner,
                                 uint32_t nq_max, uint32_t rescore_max) {
    return guarded([&] {
        need(h, "index");
        if (h->xchg) fail(GVDB_ERR_INVALID_ARGUMENT, "this index already has a peer exchange");
        if (world == 0 || world > XCHG_MAX_WORLD || rank >= world) fail(GVDB_ERR_INVALID_ARGUMENT, "bad world / rank");
        if (nq_max == 0 || rescore_max == 0 || rescore_max > kMaxR)
            fail(GVDB_ERR_INVALID_ARGUMENT, "nq_max >= 1 and rescore_max in [1, 2048]");
        if (rows_per_owner == 0) fail(GVDB_ERR_INVALID_ARGUMENT, "rows_per_owner must be >= 1");
        if ((h->dim & 3) != 0) fail(GVDB_ERR_NOT_IMPLEMENTED, "the peer exchange needs dim % 4 == 0");
        if (world > 1) {
            if (!h->windowed) fail(GVDB_ERR_INVALID_ARGUMENT, "the peer exchange needs an index created with GVDB_FLAG_ROW_WINDOW");
            if (h->win_first != (uint64_t)rank * rows_per_owner || h->win_count > rows_per_owner)
                fail(GVDB_ERR_INVALID_ARGUMENT, "this index's row window is not owner `rank`'s share");
        }
        DeviceGuard dg(h->cfg.device);
        std::unique_ptr<Exchange> x(new Exchange());
        x->world = world; x->rank = rank; x->nq_max = nq_max; x->r_max = rescore_max; x->rows_per_owner = rows_per_owner;
        const size_t slots = (size_t)world * nq_max;
        x->q_off = 0;
        x->keys_off = align256(slots * h->dim * 4);
        x->sc_off = x->keys_off + align256(slots * rescore_max * 8);
        x->set_bytes = x->sc_off + align256(slots * rescore_max * 4);
        x->flags_off = 2 * x->set_bytes;
        x->mailbox_bytes = x->flags_off + (size_t)XCHG_KINDS * world * XCHG_FLAG_STRIDE * 4;
        if (const char* e = getenv("GVDB_XCHG_TIMEOUT_MS")) x->limit_ns = (uint64_t)std::max(1, atoi(e)) * 1000000ull;
        if (const char* e = getenv("GVDB_XCHG_DMA")) x->dma_queries = e[0] == '1';
        // plain cudaMalloc (not a pool allocation): the block is exported with cudaIpcGetMemHandle
        CU(cudaMalloc((void**)&x->mailbox, x->mailbox_bytes));
        CU(cudaMemset(x->mailbox, 0, x->mailbox_bytes));
        CU(cudaMalloc((void**)&x->peers_dev, world * sizeof(uint8_t*)));
        x->my_keys.ensure((size_t)nq_max * rescore_max * 8);
        x->err.ensure(256);
        CU(cudaMemset(x->err.p, 0, 256));
        x->done.ensure(XCHG_KINDS * 128);
        CU(cudaMemset(x->done.p, 0, XCHG_KINDS * 128));
        CU(cudaMallocHost((void**)&x->h_err, 64));
        x->h_err[0] = 0;
        CU(cudaStreamCreateWithFlags(&x->side, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&x->fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&x->join, cudaEventDisableTiming));
        CU(cudaDeviceSynchronize());
        x->peers.assign(world, nullptr);
        x->peers[rank] = x->mailbox;
        if (world == 1) {
            CU(cudaMemcpy(x->peers_dev, x->peers.data(), sizeof(uint8_t*), cudaMemcpyHostToDevice));
            x->attached = true;
        }
        h->xchg = x.release();
    });
}

gvdb_status gvdb_exchange_export_ipc(gvdb_index* h, uint8_t* handle_out) {
    return guarded([&] {
        need(h, "index"); need(handle_out, "handle_out");
        if (!h->xchg) fail(GVDB_ERR_INVALID_ARGUMENT, "call gvdb_exchange_create first");
        DeviceGuard dg(h->cfg.device);
        cudaIpcMemHandle_t hd;
        CU(cudaIpcGetMemHandle(&hd, h->xchg->mailbox));
        memcpy(handle_out, &hd, sizeof(hd));
    });
}

namespace {
void xchg_attach(gvdb_index* h, const std::vector<uint8_t*>& ptrs) {
    Exchange* x = h->xchg;
    for (uint32_t w = 0; w < x->world; ++w) x->peers[w] = w == x->rank ? x->mailbox : ptrs[w];
    CU(cudaMemcpy(x->peers_dev, x->peers.data(), x->world * sizeof(uint8_t*), cudaMemcpyHostToDevice));
    x->attached = true;
}
}  // namespace

gvdb_status gvdb_exchange_attach_ipc(gvdb_index* h, const uint8_t* handles) {
    return guarded([&] {
        need(h, "index"); need(handles, "handles");
        Exchange* x = h->xchg;
        if (!x) fail(GVDB_ERR_INVALID_ARGUMENT, "call gvdb_exchange_create first");
        DeviceGuard dg(h->cfg.device);
        std::vector<uint8_t*> ptrs(x->world, nullptr);
        for (uint32_t w = 0; w < x->world; ++w) {
            if (w == x->rank) continue;
            cudaIpcMemHandle_t hd;
            memcpy(&hd, handles + (size_t)w * GVDB_IPC_HANDLE_BYTES, sizeof(hd));
            void* p = nullptr;
            CU(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
            x->opened.push_back(p);
            ptrs[w] = static_cast<uint8_t*>(p);
        }
        xchg_attach(h, ptrs);
    });
}

void* gvdb_exchange_mailbox_ptr(const gvdb_index* h) { return h && h->xchg ? h->xchg->mailbox : nullptr; }

gvdb_status gvdb_exchange_attach_ptr(gvdb_index* h, void* const* mailboxes) {
    return guarded([&] {
        need(h, "index"); need(mailboxes, "mailboxes");
        Exchange* x = h->xchg;
        if (!x) fail(GVDB_ERR_INVALID_ARGUMENT, "call gvdb_exchange_create first");
        DeviceGuard dg(h->cfg.device);
        std::vector<uint8_t*> ptrs(x->world, nullptr);
        for (uint32_t w = 0; w < x->world; ++w) {
            ptrs[w] = static_cast<uint8_t*>(mailboxes[w]);
            if (w != x->rank && !ptrs[w]) fail(GVDB_ERR_INVALID_ARGUMENT, "null peer mailbox");
        }
        xchg_attach(h, ptrs);
    });
}

gvdb_status gvdb_exchange_status(gvdb_index* h, uint32_t* timed_out_kinds) {
    return guarded([&] {
        need(h, "index");
        Exchange* x = h->xchg;
        if (!x) fail(GVDB_ERR_INVALID_ARGUMENT, "call gvdb_exchange_create first");
        DeviceGuard dg(h->cfg.device);
        uint32_t e = 0;
        CU(cudaMemcpy(&e, x->err.p, 4, cudaMemcpyDeviceToHost));   // waits for the steps in flight
        if (timed_out_kinds) *timed_out_kinds = e;
        if (e) fail(GVDB_ERR_INDEX, "peer exchange timed out waiting for a peer (kinds bit mask " + std::to_string(e) +
                                    ": 1 queries, 2 keys, 4 scores); the answers of that step are invalid");
    });
}

gvdb_status gvdb_search_exchange_device(gvdb_index* h, void* stream, const float* queries_dev, uint32_t nq,
                                        uint32_t k, uint32_t rescore_count, uint64_t* ids_out_dev,
                                        float* scores_out_dev) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_search_exchange_device");
        need(h, "index");
        Exchange* x = h->xchg;
        if (!x || !x->attached) fail(GVDB_ERR_INVALID_ARGUMENT, "peer exchange not created / peers not attached");
        need(queries_dev, "queries"); need(ids_out_dev, "ids_out"); need(scores_out_dev, "scores_out");
        const uint32_t R = rescore_count, W = x->world;
        if (nq == 0 || nq > x->nq_max) fail(GVDB_ERR_INVALID_ARGUMENT, "nq must be in [1, nq_max] (every rank calls with the same nq)");
        if (R == 0 || R > x->r_max) fail(GVDB_ERR_INVALID_ARGUMENT, "rescore_count must be in [1, rescore_max]");
        if (k > R) fail(GVDB_ERR_INVALID_ARGUMENT, "k must be <= rescore_count");
        if (h->n_live == 0) fail(GVDB_ERR_INDEX_NOT_BUILT, "index is empty");
        std::lock_guard<std::mutex> lk(x->mu);
        if (x->h_err[0] || x->broken)
            fail(GVDB_ERR_INDEX, "an earlier peer-exchange step timed out or failed; the exchange is out of step");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, (cudaStream_t)stream, true);
        Workspace* ws = lease.ws;
        cudaStream_t st = lease.stream;
        x->epoch += 1;
        // from here on the peers expect this rank's pushes for the epoch: a host-side failure leaves them
        // waiting (they time out) and this rank out of step, which later calls report at once
        struct Guard { Exchange* x; bool ok = false; ~Guard() { if (!ok) x->broken = true; } } guard{x};
        uint8_t* set = x->mailbox + (size_t)(x->epoch & 1) * x->set_bytes;
        const uint64_t set_off = (uint64_t)(x->epoch & 1) * x->set_bytes;
        const uint64_t q_bytes = (uint64_t)nq * h->dim * 4, key_bytes = (uint64_t)nq * R * 8, sc_bytes = (uint64_t)nq * R * 4;
        // my queries to every owner, on the side stream, under stage 1
        CU(cudaEventRecord(x->fork, st));
        CU(cudaStreamWaitEvent(x->side, x->fork, 0));
        if (x->dma_queries) {
            // the queries (the big message: nq x dim x 4 bytes to every peer) go by the copy engines, so that the push
            // takes no SM from query prep and the sample pass running beside it; the flag follows in stream order
            Timed t(h, ws, x->side, K_XCHG);
            for (uint32_t w = 0; w < x->world; ++w)
                CU(cudaMemcpyAsync(x->peers[w] + set_off + x->q_off + x->rank * q_bytes, queries_dev, q_bytes,
                                   cudaMemcpyDeviceToDevice, x->side));
            xchg_signal_kernel<<<1, 32, 0, x->side>>>(x->peers_dev, x->flags_off, XCHG_Q, x->world, x->rank, x->epoch);
            CU(cudaGetLastError());
        } else {
            xchg_push(h, ws, x->side, x, set_off + x->q_off + x->rank * q_bytes, queries_dev, 0, q_bytes, XCHG_Q);
        }
        CU(cudaEventRecord(x->join, x->side));
        // stage 1 on my batch
        uint64_t* my_keys = x->my_keys.as<uint64_t>();
        // A stage 1 that fails on this rank (candidate buffer overflow on an adversarial corpus) must not take
        // the exchange down: the peers already expect this epoch's pushes.  The rank then publishes an EMPTY key
        // list (nothing to score for its queries), serves its peers as usual, and reports the error for its own
        // batch at the end of the step.
        std::string stage1_error;
        try {
            for (int attempt = 0; attempt < 2; ++attempt) {
                bool optimistic = false;
                search_core(h, ws, st, queries_dev, nq, R, nullptr, nullptr, nullptr, true, attempt == 0, &optimistic, my_keys);
                if (!check_overflow(h, ws, st, optimistic)) break;
            }
        } catch (const Err& e) {
            stage1_error = e.msg;
            cudaGetLastError();
            CU(cudaMemsetAsync(my_keys, 0xFF, key_bytes, st));
        }
        xchg_push(h, ws, st, x, set_off + x->keys_off + x->rank * key_bytes, my_keys, 0, key_bytes, XCHG_K);
        CU(cudaStreamWaitEvent(st, x->join, 0));
        // every rank's queries and keys are here: score the candidates whose rows I own
        xchg_wait(h, ws, st, x, (1u << XCHG_Q) | (1u << XCHG_K));
        // ... and store each cosine straight into its requester's mailbox (slot `rank` of its sc_in)
        rescore_keys_core(h, ws, st, reinterpret_cast<const float*>(set + x->q_off), W * nq, R,
                          reinterpret_cast<const uint64_t*>(set + x->keys_off), nullptr, x->peers_dev,
                          set_off + x->sc_off + x->rank * sc_bytes, nq * R);
        // (a signal folded into the rescoring kernel costs a system-scope fence per block — each block owns an SM —
        //  and measured +0.07 ms per step at N = 2: the cosines are signalled by a kernel of their own)
        xchg_signal(h, ws, st, x, XCHG_S);
        xchg_wait(h, ws, st, x, 1u << XCHG_S);
        finish_owned_core(h, ws, st, my_keys, reinterpret_cast<const float*>(set + x->sc_off), W, x->rows_per_owner, nq, R,
                          k, ids_out_dev, scores_out_dev);
        CU(cudaMemcpyAsync(x->h_err, x->err.p, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaGetLastError());
        guard.ok = true;
        finish_async(h, ws, st);
        if (!stage1_error.empty())
            fail(GVDB_ERR_INDEX, "peer exchange: stage 1 failed on this rank (its answers of this step are empty; the "
                                 "exchange stays in step): " + stage1_error);
    });
}

// ---- BM25 ---------------------------------------------------------------------------------------------
struct gvdb_sparse {
    int device = 0;
    float k1 = 1.2f, b = 0.75f, avg_len = 0.0f;
    uint64_t n_docs = 0, n_post = 0;
    uint32_t n_terms = 0;
    std::vector<uint64_t> h_post_off;          // host copy: document frequencies for the idf
    DevBuf post_off, post_doc, post_w;         // CSR postings: documents and tf' (bm25_weight_kernel), by term
    DevBuf acc, hist, cut, keys, tie_counts, q_off, q_terms, q_tfs, q_idf, doc_out, score_out;
    DevBuf bound, seg_keys;                    // blocked path: per-query bound, per-(query, segment) best keys
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    cudaEvent_t idle = nullptr;                // recorded after the last kernel that touched the scratch buffers
    bool used = false;
    uint64_t launches = 0;                     // kernels launched by the searches of this handle
    bool force_dense = false;                  // GVDB_BM25_DENSE=1: the dense-accumulator path for every query (tests)
    std::mutex mu;                             // one search at a time per handle
};

gvdb_status gvdb_sparse_create(int32_t device, float k1, float b, gvdb_sparse** out) {
    return guarded([&] {
        need(out, "out");
        *out = nullptr;
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            fail(GVDB_ERR_INDEX, "no usable CUDA device (there is no CPU fallback)");
        if (device < 0 || device >= ndev) fail(GVDB_ERR_INVALID_ARGUMENT, "bad device ordinal");
        DeviceGuard dg(device);
        std::unique_ptr<gvdb_sparse> s(new gvdb_sparse());
        s->device = device; s->k1 = k1; s->b = b;
        if (const char* e = std::getenv("GVDB_BM25_DENSE")) s->force_dense = e[0] == '1';
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        s->sm_count = prop.multiProcessorCount;
        CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s->idle, cudaEventDisableTiming));
        *out = s.release();
    });
}

void gvdb_sparse_destroy(gvdb_sparse* s) {
    if (!s) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    for (DevBuf* b : {&s->post_off, &s->post_doc, &s->post_w, &s->acc, &s->hist, &s->cut, &s->keys, &s->tie_counts,
                      &s->q_off, &s->q_terms, &s->q_tfs, &s->q_idf, &s->doc_out, &s->score_out, &s->bound, &s->seg_keys}) b->release();
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->idle) cudaEventDestroy(s->idle);
    delete s;
    cudaSetDevice(prev);
}

gvdb_status gvdb_sparse_build(gvdb_sparse* s, uint64_t n_docs, uint32_t n_terms, const uint64_t* post_off,
                              const uint32_t* post_doc, const float* post_tf, const float* doc_len) {
    return guarded([&] {
        need(s, "sparse index"); need(post_off, "post_off");
        if (n_docs > 0xfffffff0ull) fail(GVDB_ERR_INVALID_ARGUMENT, "at most 2^32-16 documents");
        const uint64_t n_post = post_off[n_terms];
        if (n_post) { need(post_doc, "post_doc"); need(post_tf, "post_tf"); need(doc_len, "doc_len"); }
        std::lock_guard<std::mutex> lk(s->mu);
        DeviceGuard dg(s->device);
        s->h_post_off.assign(post_off, post_off + n_terms + 1);
        // the reference's average: every postings entry adds its document's length (src/sparse.rs:96-104)
        float total = 0.0f;
        for (uint32_t t = 0; t < n_terms; ++t) {
            if (post_off[t + 1] < post_off[t]) fail(GVDB_ERR_INVALID_ARGUMENT, "post_off must be non-decreasing");
            for (uint64_t p = post_off[t]; p < post_off[t + 1]; ++p) {
                if (post_doc[p] >= n_docs) fail(GVDB_ERR_INVALID_ARGUMENT, "posting refers to a document >= n_docs");
                // one entry per (term, document): the accumulate kernel touches each accumulator once per launch
                if (p > post_off[t] && post_doc[p] <= post_doc[p - 1])
                    fail(GVDB_ERR_INVALID_ARGUMENT, "documents must be strictly ascending inside a term's postings");
                total = total + doc_len[post_doc[p]];
            }
        }
        s->avg_len = n_docs ? total / (float)n_docs : 0.0f;
        s->n_docs = n_docs; s->n_terms = n_terms; s->n_post = n_post;
        s->post_off.ensure((size_t)(n_terms + 1) * 8);
        s->post_doc.ensure(n_post * 4 + 16);                          // + 16: the staged copies round their end up to 16 bytes
        s->post_w.ensure(n_post * 4 + 16);
        CU(cudaMemcpy(s->post_off.p, post_off, (size_t)(n_terms + 1) * 8, cudaMemcpyHostToDevice));
        if (n_post) {
            // k1, b and the average length are fixed from here on: the tf' factor of every posting is computed once
            // (the operations of src/sparse.rs:180-186 in their order); tfs and lengths need not stay in HBM
            DevBuf tf_tmp, len_tmp;
            struct Rel { DevBuf& a; DevBuf& b; ~Rel() { a.release(); b.release(); } } rel{tf_tmp, len_tmp};
            tf_tmp.ensure(n_post * 4);
            len_tmp.ensure(std::max<size_t>(4, n_docs * 4));
            CU(cudaMemcpy(s->post_doc.p, post_doc, n_post * 4, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(tf_tmp.p, post_tf, n_post * 4, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(len_tmp.p, doc_len, n_docs * 4, cudaMemcpyHostToDevice));
            bm25_weight_kernel<<<(unsigned)s->sm_count * 8, 256, 0, s->stream>>>(
                s->post_doc.as<uint32_t>(), tf_tmp.as<float>(), len_tmp.as<float>(), n_post, s->k1, s->b, s->avg_len,
                s->post_w.as<float>());
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(s->stream));
        }
    });
}

float gvdb_sparse_average_document_length(const gvdb_sparse* s) { return s ? s->avg_len : 0.0f; }

namespace {
// SparseIndex::search_bm25 for a batch; query CSR on the host, answers written to DEVICE buffers on `st`
// (asynchronous; the caller synchronises).  Holds s->mu.
void bm25_core(gvdb_sparse* s, cudaStream_t st, uint32_t nq, const uint64_t* q_off, const uint32_t* q_terms,
               const float* q_tfs, uint32_t limit, uint64_t* doc_out_dev, float* score_out_dev) {
    // the scratch buffers are shared by every call on this handle: a call on another stream waits for the last one
    if (s->used) CU(cudaStreamWaitEvent(st, s->idle, 0));
    struct Done { gvdb_sparse* s; cudaStream_t st; ~Done() { cudaEventRecord(s->idle, st); s->used = true; } } done{s, st};
    const uint64_t nt = q_off[nq];
    if (nt) { need(q_terms, "q_terms"); need(q_tfs, "q_tfs"); }
    // idf on the host (libm logf == Rust's f32::ln); absent terms are dropped (:169)
    std::vector<uint32_t> terms(nt);
    std::vector<float> idf(nt);
    uint32_t max_terms = 0;
    for (uint32_t q = 0; q < nq; ++q) max_terms = std::max<uint32_t>(max_terms, (uint32_t)(q_off[q + 1] - q_off[q]));
    for (uint64_t i = 0; i < nt; ++i) {
        const uint32_t t = q_terms[i];
        const uint64_t df = t < s->n_terms ? s->h_post_off[t + 1] - s->h_post_off[t] : 0;
        terms[i] = df ? t : s->n_terms;                            // n_terms = "absent"
        idf[i] = df ? std::log(((float)s->n_docs - (float)df + 0.5f) / ((float)df + 0.5f)) : 0.0f;
    }
    s->q_off.ensure((size_t)(nq + 1) * 8);
    s->q_terms.ensure(std::max<size_t>(4, nt * 4));
    s->q_tfs.ensure(std::max<size_t>(4, nt * 4));
    s->q_idf.ensure(std::max<size_t>(4, nt * 4));
    // pageable host sources: cudaMemcpyAsync stages them before it returns
    CU(cudaMemcpyAsync(s->q_off.p, q_off, (size_t)(nq + 1) * 8, cudaMemcpyHostToDevice, st));
    if (nt) {
        CU(cudaMemcpyAsync(s->q_terms.p, terms.data(), nt * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(s->q_tfs.p, q_tfs, nt * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(s->q_idf.p, idf.data(), nt * 4, cudaMemcpyHostToDevice, st));
    }
    if (limit <= (uint32_t)BMB_MAX_LIMIT && max_terms <= (uint32_t)BMB_MAX_TERMS && !s->force_dense) {
        // ---- the blocked path: accumulators in shared memory, two launches per batch (gvdb_sparse.cuh) ----
        const uint32_t n_blocks = (uint32_t)((s->n_docs + BMB_DOCS - 1) / BMB_DOCS);
        const uint32_t want_seg = (uint32_t)std::min<uint64_t>(n_blocks, std::max<uint64_t>(1, ((uint64_t)s->sm_count * 8 + nq - 1) / nq));
        uint32_t max_bps = BMB_MAX_BPS;
        if (const char* e = std::getenv("GVDB_BM25_BPS")) max_bps = std::max(1, std::min(BMB_MAX_BPS, atoi(e)));
        max_bps = std::max<uint32_t>(1, std::min<uint32_t>(max_bps, 4096 / std::max<uint32_t>(1, max_terms) - 1));   // boundary table <= 16 KB
        const uint32_t bps = std::min<uint32_t>(max_bps, (n_blocks + want_seg - 1) / want_seg);
        const uint32_t n_seg = (n_blocks + bps - 1) / bps;
        const uint32_t t_cap = std::max<uint32_t>(1, max_terms);
        uint32_t LP = 64;
        while (LP < limit) LP <<= 1;
        const size_t smem = bmb_smem_bytes(LP, t_cap, bps);
        static std::atomic<uint64_t> attr_a{0}, attr_b{0};
        ensure_dyn_smem(attr_a, bm25_block_kernel, 200 * 1024);
        ensure_dyn_smem(attr_b, bm25_merge_kernel, 2 * BMB_MAX_LIMIT * 8);
        s->bound.ensure((size_t)nq * 8);
        s->seg_keys.ensure((size_t)nq * n_seg * limit * 8);
        CU(cudaMemsetAsync(s->bound.p, 0xFF, (size_t)nq * 8, st));
        bm25_block_kernel<<<n_seg * nq, BMB_THREADS, smem, st>>>(
            s->post_off.as<uint64_t>(), s->post_doc.as<uint32_t>(), s->post_w.as<float>(), s->n_terms,
            s->q_off.as<uint64_t>(), s->q_terms.as<uint32_t>(), s->q_tfs.as<float>(), s->q_idf.as<float>(), nq, n_blocks,
            n_seg, bps, t_cap, limit, LP, s->bound.as<unsigned long long>(), s->seg_keys.as<uint64_t>());
        bm25_merge_kernel<<<nq, 256, (size_t)2 * LP * 8, st>>>(s->seg_keys.as<uint64_t>(), n_seg, limit, LP, doc_out_dev,
                                                             score_out_dev);
        CU(cudaGetLastError());
        s->launches += 2;
        return;
    }
    const uint64_t stride = (s->n_docs + 3) / 4 * 4;               // accumulators per query: uint4-readable, padding stays "absent"
    const uint32_t QC = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(nq, 64), (512ull << 20) / (stride * 4)));
    // keys inside the cut: exactly min(limit, touched) per query; up to SORT_N they are ordered by a block bitonic sort,
    // above that by a device radix sort, one query at a time
    const uint32_t key_cap = std::max<uint32_t>(SORT_N, limit);
    const bool big_limit = limit > (uint32_t)SORT_N;
    size_t sort_tmp = 0;
    if (big_limit) {
        CU(cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int64_t)key_cap, 0, 64, st));
        s->seg_keys.ensure((size_t)key_cap * 8 + sort_tmp + 256);      // sorted keys of one query + the sort's scratch
    }
    s->acc.ensure((size_t)QC * stride * 4);
    s->hist.ensure((size_t)QC * BM25_LEVEL_BINS * 4);
    s->cut.ensure((size_t)QC * sizeof(Bm25Cut));
    s->keys.ensure((size_t)QC * key_cap * 8);
    s->tie_counts.ensure((size_t)QC * s->sm_count * 2 * 4);
    const unsigned gx = (unsigned)s->sm_count * 2;
    for (uint32_t q0 = 0; q0 < nq; q0 += QC) {
        const uint32_t m = std::min(QC, nq - q0);
        CU(cudaMemsetAsync(s->acc.p, 0xFF, (size_t)m * stride * 4, st));
        CU(cudaMemsetAsync(s->cut.p, 0, (size_t)m * sizeof(Bm25Cut), st));
        if (big_limit) CU(cudaMemsetAsync(s->keys.p, 0xFF, (size_t)m * key_cap * 8, st));
        for (uint32_t rank = 0; rank < max_terms; ++rank)
            bm25_accumulate_kernel<<<dim3(gx, m), 256, 0, st>>>(
                s->post_off.as<uint64_t>(), s->post_doc.as<uint32_t>(), s->post_w.as<float>(),
                s->n_terms, s->q_off.as<uint64_t>(), s->q_terms.as<uint32_t>(), s->q_tfs.as<float>(),
                s->q_idf.as<float>(), q0, (int)rank, stride, s->acc.as<uint32_t>());
        for (int level = 0; level < 3; ++level) {
            CU(cudaMemsetAsync(s->hist.p, 0, (size_t)m * BM25_LEVEL_BINS * 4, st));
            uint32_t* hist = s->hist.as<uint32_t>();
            Bm25Cut* cut = s->cut.as<Bm25Cut>();
            if (level == 0) bm25_hist_kernel<0><<<dim3(gx, m), 256, 0, st>>>(s->acc.as<uint32_t>(), stride, hist, cut);
            else if (level == 1) bm25_hist_kernel<1><<<dim3(gx, m), 256, 0, st>>>(s->acc.as<uint32_t>(), stride, hist, cut);
            else bm25_hist_kernel<2><<<dim3(gx, m), 256, 0, st>>>(s->acc.as<uint32_t>(), stride, hist, cut);
            bm25_cut_kernel<<<m, 256, 0, st>>>(hist, limit, cut, level);
        }
        // surplus ties at the cut score: the lowest document numbers are kept (both kernels return at once otherwise)
        bm25_tiecount_kernel<<<dim3(gx, m), 256, 0, st>>>(s->acc.as<uint32_t>(), stride, limit, s->cut.as<Bm25Cut>(),
                                                         s->tie_counts.as<uint32_t>());
        bm25_tiecut_kernel<<<m, 256, 0, st>>>(s->acc.as<uint32_t>(), stride, limit, s->cut.as<Bm25Cut>(),
                                             s->tie_counts.as<uint32_t>(), gx);
        bm25_compact_kernel<<<dim3(gx, m), 256, 0, st>>>(s->acc.as<uint32_t>(), stride, s->cut.as<Bm25Cut>(),
                                                        s->keys.as<uint64_t>(), key_cap);
        if (!big_limit) {
            bm25_topk_kernel<<<m, SORT_THREADS, SORT_N * 8, st>>>(s->keys.as<uint64_t>(), key_cap, s->cut.as<Bm25Cut>(),
                                                                 s->acc.as<uint32_t>(), stride, limit,
                                                                 doc_out_dev + (size_t)q0 * limit, score_out_dev + (size_t)q0 * limit);
        } else {
            // unwritten key slots must sort last: the compaction wrote min(limit, touched) keys, the rest is set here
            uint64_t* sorted = s->seg_keys.as<uint64_t>();
            void* tmp = s->seg_keys.as<uint8_t>() + (size_t)key_cap * 8;
            for (uint32_t qi = 0; qi < m; ++qi) {
                size_t tb = sort_tmp;
                CU(cub::DeviceRadixSort::SortKeys(tmp, tb, s->keys.as<uint64_t>() + (size_t)qi * key_cap, sorted, (int64_t)key_cap, 0, 64, st));
                bm25_emit_sorted_kernel<<<(limit + 255) / 256, 256, 0, st>>>(
                    sorted, s->cut.as<Bm25Cut>() + qi, s->acc.as<uint32_t>() + (size_t)qi * stride, limit,
                    doc_out_dev + (size_t)(q0 + qi) * limit, score_out_dev + (size_t)(q0 + qi) * limit);
            }
        }
        CU(cudaGetLastError());
    }
    s->launches += (uint64_t)((nq + QC - 1) / QC) * (max_terms + 10);
}
}  // namespace

gvdb_status gvdb_sparse_search_bm25_batch(gvdb_sparse* s, uint32_t nq, const uint64_t* q_off, const uint32_t* q_terms,
                                          const float* q_tfs, uint32_t limit, uint64_t* doc_out, float* score_out) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_sparse_search_bm25_batch");
        need(s, "sparse index");
        if (nq == 0 || limit == 0) return;
        need(q_off, "q_off"); need(doc_out, "doc_out"); need(score_out, "score_out");
        std::lock_guard<std::mutex> lk(s->mu);
        for (uint64_t i = 0; i < (uint64_t)nq * limit; ++i) { doc_out[i] = UINT64_MAX; score_out[i] = -INFINITY; }
        if (s->n_docs == 0) return;                                   // src/sparse.rs:161-163
        DeviceGuard dg(s->device);
        s->doc_out.ensure((size_t)nq * limit * 8);
        s->score_out.ensure((size_t)nq * limit * 4);
        bm25_core(s, s->stream, nq, q_off, q_terms, q_tfs, limit, s->doc_out.as<uint64_t>(), s->score_out.as<float>());
        CU(cudaMemcpyAsync(doc_out, s->doc_out.p, (size_t)nq * limit * 8, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaMemcpyAsync(score_out, s->score_out.p, (size_t)nq * limit * 4, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    });
}

gvdb_status gvdb_sparse_search_bm25_batch_device(gvdb_sparse* s, void* stream, uint32_t nq, const uint64_t* q_off,
                                                 const uint32_t* q_terms, const float* q_tfs, uint32_t limit,
                                                 uint64_t* doc_out_dev, float* score_out_dev) {
    return guarded([&] {
        NvtxRange nvtx_call("gvdb_sparse_search_bm25_batch_device");
        need(s, "sparse index");
        if (nq == 0 || limit == 0) return;
        need(q_off, "q_off"); need(doc_out_dev, "doc_out"); need(score_out_dev, "score_out");
        std::lock_guard<std::mutex> lk(s->mu);
        DeviceGuard dg(s->device);
        cudaStream_t st = (cudaStream_t)stream;
        if (s->n_docs == 0) {                                         // src/sparse.rs:161-163
            CU(cudaMemsetAsync(doc_out_dev, 0xFF, (size_t)nq * limit * 8, st));
            fill_f32_kernel<<<(unsigned)(((size_t)nq * limit + 255) / 256), 256, 0, st>>>(score_out_dev, (size_t)nq * limit, -INFINITY);
            return;
        }
        bm25_core(s, st, nq, q_off, q_terms, q_tfs, limit, doc_out_dev, score_out_dev);
    });
}

uint64_t gvdb_sparse_launches(const gvdb_sparse* s) { return s ? s->launches : 0; }

// ---- rrf_fusion ---------------------------------------------------------------------------------------
namespace {
// mode 0: rrf_fusion; 1: linear_fusion; 2: normalized_fusion (gvdb_sparse.cuh)
void fusion_launch(int mode, cudaStream_t st, const uint64_t* dense, uint32_t n_d, const uint64_t* sparse, uint32_t n_s,
                   const uint64_t* text, uint32_t n_t, uint32_t nq, float k, FusionScores fs, uint32_t limit,
                   uint64_t* ids_out, float* scores_out) {
    const uint32_t n = n_d + n_s + n_t;
    if (n == 0 || n > RRF_MAX) fail(GVDB_ERR_INVALID_ARGUMENT, "fusion: the three list lengths must sum to 1..4096");
    if ((n_d && !dense) || (n_s && !sparse) || (n_t && !text)) fail(GVDB_ERR_INVALID_ARGUMENT, "fusion: null list");
    if (mode != 0 && ((n_d && !fs.dense) || (n_s && !fs.sparse) || (n_t && !fs.text)))
        fail(GVDB_ERR_INVALID_ARGUMENT, "fusion: null score list");
    uint32_t n_eff = 64;
    while (n_eff < n) n_eff <<= 1;
    const size_t smem = (size_t)n * 8 + (size_t)n_eff * 8 + (size_t)n * 8;
    static std::atomic<uint64_t> attr0{0}, attr1{0}, attr2{0};
    const unsigned threads = n_eff <= 512 ? 256 : SORT_THREADS;
    if (mode == 0) {
        ensure_dyn_smem(attr0, list_fusion_kernel<0>, (int)(RRF_MAX * 24));
        list_fusion_kernel<0><<<nq, threads, smem, st>>>(dense, n_d, sparse, n_s, text, n_t, k, fs, limit, ids_out, scores_out);
    } else if (mode == 1) {
        ensure_dyn_smem(attr1, list_fusion_kernel<1>, (int)(RRF_MAX * 24));
        list_fusion_kernel<1><<<nq, threads, smem, st>>>(dense, n_d, sparse, n_s, text, n_t, k, fs, limit, ids_out, scores_out);
    } else {
        ensure_dyn_smem(attr2, list_fusion_kernel<2>, (int)(RRF_MAX * 24));
        list_fusion_kernel<2><<<nq, threads, smem, st>>>(dense, n_d, sparse, n_s, text, n_t, k, fs, limit, ids_out, scores_out);
    }
    CU(cudaGetLastError());
}

// host lists in, host lists out (shared by the rrf and the weighted entry points)
void fusion_host(int mode, int32_t device, const uint64_t* dense, const float* dense_sc, uint32_t n_dense,
                 const uint64_t* sparse, const float* sparse_sc, uint32_t n_sparse, const uint64_t* text,
                 const float* text_sc, uint32_t n_text, uint32_t nq, float k, float w_d, float w_s, float w_t,
                 uint32_t limit, uint64_t* ids_out, float* scores_out) {
    if (nq == 0 || limit == 0) return;
    need(ids_out, "ids_out"); need(scores_out, "scores_out");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        fail(GVDB_ERR_INDEX, "no usable CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) fail(GVDB_ERR_INVALID_ARGUMENT, "bad device ordinal");
    if ((n_dense && !dense) || (n_sparse && !sparse) || (n_text && !text)) fail(GVDB_ERR_INVALID_ARGUMENT, "fusion: null list");
    if (mode != 0 && ((n_dense && !dense_sc) || (n_sparse && !sparse_sc) || (n_text && !text_sc)))
        fail(GVDB_ERR_INVALID_ARGUMENT, "fusion: null score list");
    DeviceGuard dg(device);
    const size_t nl = (size_t)nq * (n_dense + n_sparse + n_text);
    DevBuf in, insc, out;
    struct Rel { DevBuf &a, &b, &c; ~Rel() { a.release(); b.release(); c.release(); } } rel{in, insc, out};
    in.ensure(std::max<size_t>(8, nl * 8));
    insc.ensure(std::max<size_t>(8, nl * 4));
    out.ensure((size_t)nq * limit * 12);
    uint64_t* d = in.as<uint64_t>();
    uint64_t* sp = d + (size_t)nq * n_dense;
    uint64_t* tx = sp + (size_t)nq * n_sparse;
    float* ds = insc.as<float>();
    float* ss = ds + (size_t)nq * n_dense;
    float* ts = ss + (size_t)nq * n_sparse;
    if (n_dense) CU(cudaMemcpy(d, dense, (size_t)nq * n_dense * 8, cudaMemcpyHostToDevice));
    if (n_sparse) CU(cudaMemcpy(sp, sparse, (size_t)nq * n_sparse * 8, cudaMemcpyHostToDevice));
    if (n_text) CU(cudaMemcpy(tx, text, (size_t)nq * n_text * 8, cudaMemcpyHostToDevice));
    if (mode != 0) {
        if (n_dense) CU(cudaMemcpy(ds, dense_sc, (size_t)nq * n_dense * 4, cudaMemcpyHostToDevice));
        if (n_sparse) CU(cudaMemcpy(ss, sparse_sc, (size_t)nq * n_sparse * 4, cudaMemcpyHostToDevice));
        if (n_text) CU(cudaMemcpy(ts, text_sc, (size_t)nq * n_text * 4, cudaMemcpyHostToDevice));
    }
    uint64_t* oi = out.as<uint64_t>();
    float* os = reinterpret_cast<float*>(oi + (size_t)nq * limit);
    fusion_launch(mode, nullptr, n_dense ? d : nullptr, n_dense, n_sparse ? sp : nullptr, n_sparse, n_text ? tx : nullptr, n_text,
                  nq, k, FusionScores{ds, ss, ts, w_d, w_s, w_t}, limit, oi, os);
    CU(cudaMemcpy(ids_out, oi, (size_t)nq * limit * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(scores_out, os, (size_t)nq * limit * 4, cudaMemcpyDeviceToHost));
}
}  // namespace

gvdb_status gvdb_rrf_fusion_batch_device(int32_t device, void* stream, const uint64_t* dense_dev, uint32_t n_dense,
                                         const uint64_t* sparse_dev, uint32_t n_sparse, const uint64_t* text_dev,
                                         uint32_t n_text, uint32_t nq, float k, uint32_t limit, uint64_t* ids_out_dev,
                                         float* scores_out_dev) {
    return guarded([&] {
        if (nq == 0 || limit == 0) return;
        need(ids_out_dev, "ids_out"); need(scores_out_dev, "scores_out");
        DeviceGuard dg(device);
        fusion_launch(0, (cudaStream_t)stream, dense_dev, n_dense, sparse_dev, n_sparse, text_dev, n_text, nq, k,
                      FusionScores{nullptr, nullptr, nullptr, 0.f, 0.f, 0.f}, limit, ids_out_dev, scores_out_dev);
    });
}

gvdb_status gvdb_rrf_fusion_batch(int32_t device, const uint64_t* dense, uint32_t n_dense, const uint64_t* sparse,
                                  uint32_t n_sparse, const uint64_t* text, uint32_t n_text, uint32_t nq, float k,
                                  uint32_t limit, uint64_t* ids_out, float* scores_out) {
    return guarded([&] {
        fusion_host(0, device, dense, nullptr, n_dense, sparse, nullptr, n_sparse, text, nullptr, n_text, nq, k, 0.f, 0.f, 0.f,
                    limit, ids_out, scores_out);
    });
}

gvdb_status gvdb_weighted_fusion_batch_device(int32_t device, void* stream, const uint64_t* dense_dev,
                                              const float* dense_scores_dev, uint32_t n_dense, const uint64_t* sparse_dev,
                                              const float* sparse_scores_dev, uint32_t n_sparse, const uint64_t* text_dev,
                                              const float* text_scores_dev, uint32_t n_text, uint32_t nq, float dense_weight,
                                              float sparse_weight, float text_weight, int32_t normalize, uint32_t limit,
                                              uint64_t* ids_out_dev, float* scores_out_dev) {
    return guarded([&] {
        if (nq == 0 || limit == 0) return;
        need(ids_out_dev, "ids_out"); need(scores_out_dev, "scores_out");
        DeviceGuard dg(device);
        fusion_launch(normalize ? 2 : 1, (cudaStream_t)stream, dense_dev, n_dense, sparse_dev, n_sparse, text_dev, n_text, nq, 0.f,
                      FusionScores{dense_scores_dev, sparse_scores_dev, text_scores_dev, dense_weight, sparse_weight, text_weight},
                      limit, ids_out_dev, scores_out_dev);
    });
}

gvdb_status gvdb_weighted_fusion_batch(int32_t device, const uint64_t* dense, const float* dense_scores, uint32_t n_dense,
                                       const uint64_t* sparse, const float* sparse_scores, uint32_t n_sparse,
                                       const uint64_t* text, const float* text_scores, uint32_t n_text, uint32_t nq,
                                       float dense_weight, float sparse_weight, float text_weight, int32_t normalize,
                                       uint32_t limit, uint64_t* ids_out, float* scores_out) {
    return guarded([&] {
        fusion_host(normalize ? 2 : 1, device, dense, dense_scores, n_dense, sparse, sparse_scores, n_sparse, text, text_scores,
                    n_text, nq, 0.f, dense_weight, sparse_weight, text_weight, limit, ids_out, scores_out);
    });
}


gvdb_status gvdb_approx_dot(gvdb_index* h, const float* queries, uint32_t nq, float* dot_out) {
    return guarded([&] {
        need(h, "index");
        if (nq == 0 || h->n_rows == 0) return;
        need(queries, "queries"); need(dot_out, "dot_out");
        if (!ratio_supported(h)) fail(GVDB_ERR_NOT_IMPLEMENTED, "tensor-core dot products need dim % 64 == 0, dim <= 768 and an index without a row window");
        DeviceGuard dg(h->cfg.device);
        WsLease lease(h, nullptr, false);
        Workspace* ws = lease.ws;
        cudaStream_t st = lease.stream;
        ensure_rows16(h, st);
        const int nu = ratio_units(h);
        const uint64_t N = h->n_rows;
        const uint32_t qchunk = (uint32_t)std::max<uint64_t>(TC_NQ, std::min<uint64_t>(1024, ((256ull << 20) / (N * 4)) / TC_NQ * TC_NQ));
        ws->q_in.ensure((size_t)qchunk * h->dim * 4);
        ws->qexp.ensure((size_t)(qchunk / TC_NQ) * tc_qblock_bytes(nu));
        ws->misc.ensure((size_t)qchunk * N * 4);
        for (uint32_t q0 = 0; q0 < nq; q0 += qchunk) {
            const uint32_t m = std::min(qchunk, nq - q0), m_pad = (m + TC_NQ - 1) / TC_NQ * TC_NQ;
            CU(cudaMemcpyAsync(ws->q_in.p, queries + (size_t)q0 * h->dim, (size_t)m * h->dim * 4, cudaMemcpyHostToDevice, st));
            const uint64_t words = (uint64_t)m_pad * nu * 16;
            ratio_q16_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(ws->q_in.as<float>(), m, m_pad, h->dim, nu, ws->qexp.as<int8_t>());
            launch_tc_dot<1>(h, ws, st, ws->qexp.as<int8_t>(), nullptr, m, m_pad, nullptr, 0, nullptr, nullptr, ws->misc.as<float>(), N, nullptr);
            CU(cudaMemcpyAsync(dot_out + (size_t)q0 * N, ws->misc.p, (size_t)m * N * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        flush_profile(h, ws);
    });
}

gvdb_status gvdb_measure_fp4_mma_rate(int32_t device, double* tmacs_per_s_out, double* clk_per_mma_out) {
    return guarded([&] {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            fail(GVDB_ERR_INDEX, "no usable CUDA device (there is no CPU fallback)");
        if (device < 0 || device >= ndev) fail(GVDB_ERR_INVALID_ARGUMENT, "bad device ordinal");
        DeviceGuard dg(device);
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        const int sms = prop.multiProcessorCount;
        DevBuf clk;
        struct Rel { DevBuf& b; ~Rel() { b.release(); } } rel{clk};
        clk.ensure(64);
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        struct Ev { cudaEvent_t a, b; ~Ev() { cudaEventDestroy(a); cudaEventDestroy(b); } } ev{e0, e1};
        const int niter = 40000, reps = 4;                       // ~1.4 ms per launch at 64 clk per MMA
        tc_mma_rate_kernel<<<sms, 128, 4096>>>(niter, clk.as<long long>());      // warm-up
        CU(cudaGetLastError());
        CU(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) tc_mma_rate_kernel<<<sms, 128, 4096>>>(niter, clk.as<long long>());
        CU(cudaEventRecord(e1));
        CU(cudaDeviceSynchronize());
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        long long c = 0;
        CU(cudaMemcpy(&c, clk.p, 8, cudaMemcpyDeviceToHost));
        const double macs = (double)reps * sms * niter * 128.0 * 128.0 * 64.0;
        if (tmacs_per_s_out) *tmacs_per_s_out = macs / (ms * 1e-3) / 1e12;
        if (clk_per_mma_out) *clk_per_mma_out = (double)c / niter;
    });
}

gvdb_status gvdb_profile_enable(gvdb_index* h, int32_t on) {
    return guarded([&] {
        need(h, "index");
        h->profile_on.store(on ? 1 : 0);
    });
}

gvdb_status gvdb_profile_read(gvdb_index* h, gvdb_profile* out, int32_t reset) {
    return guarded([&] {
        need(h, "index"); need(out, "out");
        std::lock_guard<std::mutex> lk(h->prof_mu);
        h->prof.launches = h->launches.load();
        h->prof.optimistic_reruns = h->optimistic_reruns.load();
        h->prof.overflow_fallbacks = h->overflow_fallbacks.load();
        *out = h->prof;
        if (reset) {
            h->prof = gvdb_profile{};
            h->launches.store(0);
            h->optimistic_reruns.store(0);
            h->overflow_fallbacks.store(0);
        }
    });
}

}  // extern "C"
