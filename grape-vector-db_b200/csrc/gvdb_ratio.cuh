// gvdb_ratio.cuh — ratio mode on the tensor cores (K3b): rescore_count = (N as f32 * rescore_ratio) as usize,
// the reference's DEFAULT (rescore_ratio = 0.1: 100 000 candidates per query on 1M rows,
// /root/reference/src/quantization.rs:22-31,178-216).  Rescoring that many candidates per query by gathering
// their f32 rows costs 307 MB of HBM traffic per query; for a batch of queries it really is a dense GEMM
// (queries x rows), so it runs on the tensor cores — as a FILTER whose survivors are then rescored exactly:
//
//   1. the fast path with R' = 256 (gvdb_tc.cuh): the exact top k by cosine among the 256 rows closest in
//      Hamming distance.  They all belong to the R candidates (256 < R), so their k-th best cosine c_k is a
//      LOWER bound of the answer's k-th best cosine.
//   2. b*[q], the Hamming distance of the R-th candidate: a few passes of the FP4 scan that only COUNT the rows
//      below a per-query threshold (MODE 3), steered by ratio_bstar_kernel (bracket + bisection from a sampled
//      quantile).  Membership of a row in the candidate set: hamming < b* yes, > b* no, == b* a tie whose
//      membership depends on its row number (the stable sort) — left to the exact path (next point).
//   3. tc_dot_kernel: every (row, query) dot product on tcgen05.mma kind::f16 (bf16 operands, f32 accumulate, A in
//      TMEM, the query block resident in shared memory).  bf16 rounding moves a cosine by at most 2^-8
//      (|sum q_j r_j| <= |q||r|, relative error 2^-9 per operand), so every row whose approximate cosine
//      is below c_k - eps cannot be in the answer; the few rows above it leave as (row, query) records.
//   4. ratio_scatter_kernel recomputes the records' Hamming distances from the codes, drops non-members and
//      rows the fast path already holds, and marks the query for the exact fallback if a tie at b* shows up
//      or a list overflows; the survivors are rescored EXACTLY (sequential-fold f32 cosine,
//      rescore_owned_ring_kernel) and merged with the fast path's records (ratio_finish_kernel).
// The answer is the reference's, bit for bit (ids and cosines); queries the filter cannot vouch
// for fall back to the cut by counting (gvdb_bigr.cuh).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "gvdb_tc.cuh"

namespace gvdb {

constexpr float RATIO_EPS = 0.0045f;          // 2^-8 (bf16 operands) + slack for the norms and the f32 accumulation
constexpr int RATIO_MAX_UNITS = 24;           // 64-byte A units per row: 32 dims each -> dim <= 768 (the query block must fit shared memory)
constexpr uint32_t RATIO_E_CAP = 2048;        // filter survivors kept per query
constexpr uint64_t RATIO_TIE_BIT = 1ull << 63; // E key flag: hamming == b*, membership undecided

// f32 rows -> bf16, blocked for the expanders: a UNIT is 32 consecutive dims of one row (64 bytes = the A operand
// of two K=16 MMAs); piece j (16 bytes) of unit u of row r of tile t at uint4 index ((t*NU + u)*4 + j)*32 + r.
// Dims beyond `dim` (up to NU*32) are zero.  One thread per (row, unit).
__global__ void __launch_bounds__(256)
ratio_rows16_kernel(const float* __restrict__ rows, uint64_t first, uint64_t n, int dim, int nu, uint4* __restrict__ rows16) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * (uint64_t)nu) return;
    const uint64_t r = first + i / nu;
    const int u = (int)(i % nu);
    const float* src = rows + r * (uint64_t)dim + u * 32;
    uint32_t w[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int d0 = u * 32 + 2 * e;
        const float a = d0 < dim ? src[2 * e] : 0.0f, b = d0 + 1 < dim ? src[2 * e + 1] : 0.0f;
        const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
        w[e] = *reinterpret_cast<const uint32_t*>(&v);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        rows16[(((r >> 5) * (uint64_t)nu + u) * 4 + j) * 32 + (r & 31)] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}

// queries -> bf16 in the byte order the resident block needs (the layout of tc_qexp_offset with K bytes = 2 per
// dim): K byte kb of query q.  One thread per (query, 4 K bytes = 2 dims).  Padding queries / dims are zero.
__global__ void __launch_bounds__(256)
ratio_q16_kernel(const float* __restrict__ queries, uint32_t nq, uint32_t nq_pad, int dim, int nu, int8_t* __restrict__ q16) {
    const uint32_t words = (uint32_t)nu * 16;                       // 32-bit words per query
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)nq_pad * words) return;
    const uint32_t q = (uint32_t)(i / words), o = (uint32_t)(i % words);
    const int d0 = (int)o * 2;
    float a = 0.0f, b = 0.0f;
    if (q < nq) {
        if (d0 < dim) a = queries[(size_t)q * dim + d0];
        if (d0 + 1 < dim) b = queries[(size_t)q * dim + d0 + 1];
    }
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    // nchunk of the FP4 layout == nu here: both count 64-byte K slices of the operand (2 per 16 KB sub-block... see tc_qexp_offset)
    *reinterpret_cast<uint32_t*>(q16 + tc_qexp_offset(q, (int)o, nu)) = *reinterpret_cast<const uint32_t*>(&v);
}

__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// UMMA instruction descriptor, kind::f16: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9 / 10-12 = 1), both K-major
__host__ __device__ constexpr uint32_t tc_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- the dense pass ----------------------------------------------------------------------------------
// Same roles as tc_scan_kernel (4 expander warps, 8 epilogue warps, 1 loader/MMA warp; A ring of four TMEM
// slots, two accumulator buffers) with one resident query block per item: item = (row slice, query block),
// query blocks of one row slice on neighbouring CTAs so that they read the slice's rows out of L2 together.
//   MODE 0  filter: row r passes for query q iff dot * rinv[r] > thr[q]  (thr = (c_k - eps) * |q|; rinv = 1/|r|,
//           0 for a zero row) -> (row, query) records in warp-private lists (as tc_scan_kernel MODE 0).
//   MODE 1  every approximate dot product to dot_out[q * stride + row] (tests).
template <int NU, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_dot_kernel(const uint4* __restrict__ rows16, const uint32_t* __restrict__ live, const float* __restrict__ rinv,
              uint32_t n_tiles, uint32_t ngroups, const int8_t* __restrict__ q16, const float* __restrict__ thr,
              uint32_t nq, uint32_t nq_pad, uint32_t n_rslices, uint2* __restrict__ recs, uint32_t rec_cap,
              uint32_t* __restrict__ list_counts, uint32_t* __restrict__ overflow, float* __restrict__ dot_out,
              uint64_t dot_stride, uint64_t n_rows) {
    constexpr int NBUF = 2, NSLOT = 4;
    constexpr int SC = NU % 3 == 0 ? 3 : (NU % 2 == 0 ? 2 : 1);   // units per A slot
    constexpr int PH = NU / SC;
    constexpr int UNIT_COLS = 16;                                   // 64 bytes per row
    constexpr int SLOT_COLS = SC * UNIT_COLS;
    constexpr int A_COLS = NSLOT * SLOT_COLS;
    constexpr uint32_t IDESC = tc_idesc_bf16(TC_ROWS, TC_NQ);
    constexpr uint32_t QBLOCK_BYTES = (uint32_t)((NU + 1) / 2) * TC_STAGE_BYTES + TC_BIAS_BYTES;   // tc_qblock_bytes(NU)
    static_assert(A_COLS + NBUF * TC_NQ <= 512, "TMEM budget");
    static_assert(QBLOCK_BYTES <= 216 * 1024, "the resident query block must fit shared memory");

    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ float s_thr[TC_NQ];
    __shared__ uint32_t s_qrow[MODE == 0 ? TC_EPI_WARPS * TC_QUEUE : 1];
    __shared__ uint32_t s_qq[MODE == 0 ? TC_EPI_WARPS * TC_QUEUE : 1];
    __shared__ __align__(8) uint64_t bars[2 + 2 * NSLOT + 2 * NBUF];
    __shared__ uint32_t s_tmem_base;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t b_full = bar0, b_free = bar0 + 8;
    auto a_ready = [&](uint32_t s) { return bar0 + 8u * (2 + s); };
    auto a_free = [&](uint32_t s) { return bar0 + 8u * (2 + NSLOT + s); };
    auto acc_full = [&](uint32_t b) { return bar0 + 8u * (2 + 2 * NSLOT + b); };
    auto acc_empty = [&](uint32_t b) { return bar0 + 8u * (2 + 2 * NSLOT + NBUF + b); };

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t nqb = nq_pad / TC_NQ;
    const uint32_t n_items = nqb * n_rslices;
    if (threadIdx.x == 0) {
        mbar_init(b_full, 1); mbar_init(b_free, 1);
        for (int s = 0; s < NSLOT; ++s) { mbar_init(a_ready(s), 4); mbar_init(a_free(s), 1); }
        for (int b = 0; b < NBUF; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), TC_EPI_WARPS); }
        fence_mbar_init();
    }
    constexpr int MMA_WARP = 4 + TC_EPI_WARPS;
    if (warp == MMA_WARP) tc_alloc(smem_u32(&s_tmem_base), TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (s_tmem_base != 0u) { if (threadIdx.x == 0 && overflow) overflow[2] = 1u; __trap(); }
    constexpr uint32_t tmem_a = 0, tmem_d = A_COLS;
    const uint32_t lane_taddr = (uint32_t)((warp & 3) * 32) << 16;

    // item -> (query block, row groups): item = row slice * nqb + query block
    auto item_range = [&](uint32_t item, uint32_t& qb, uint32_t& g_lo, uint32_t& g_hi) {
        const uint32_t rsl = item / nqb;
        qb = item % nqb;
        g_lo = (uint32_t)((uint64_t)ngroups * rsl / n_rslices);
        g_hi = (uint32_t)((uint64_t)ngroups * (rsl + 1) / n_rslices);
    };

    if (warp < 4) {
        // ===================== expanders: bf16 rows -> A operand in TMEM (no expansion, 64 B per unit) =====================
        uint32_t p = 0;
        auto load_phase = [&](uint32_t g, int ph, uint4 (&r)[SC * 4]) {
            const uint32_t tile = g * 4u + warp;
            const bool in_range = tile < n_tiles;
#pragma unroll
            for (int i = 0; i < SC * 4; ++i)
                r[i] = in_range ? ldg_stream(rows16 + (((size_t)tile * NU + ph * SC) * 4 + i) * 32 + lane) : make_uint4(0, 0, 0, 0);
        };
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            uint32_t qb, g_lo, g_hi;
            item_range(item, qb, g_lo, g_hi);
            uint4 r[SC * 4];
            if (g_lo < g_hi) load_phase(g_lo, 0, r);
            for (uint32_t g = g_lo; g < g_hi; ++g) {
#pragma unroll 1
                for (int ph = 0; ph < PH; ++ph, ++p) {
                    uint4 rn[SC * 4];
                    const bool more = ph + 1 < PH || g + 1 < g_hi;
                    if (more) load_phase(ph + 1 < PH ? g : g + 1, ph + 1 < PH ? ph + 1 : 0, rn);   // in flight while this phase is stored
                    const uint32_t s = p % NSLOT;
                    mbar_wait(a_free(s), ((p / NSLOT) & 1u) ^ 1u);
                    tc_fence_after();
#pragma unroll
                    for (int i = 0; i < SC * 2; ++i) {                  // 8 columns (32 bytes) per store
                        const uint32_t v[8] = {r[2 * i].x, r[2 * i].y, r[2 * i].z, r[2 * i].w,
                                               r[2 * i + 1].x, r[2 * i + 1].y, r[2 * i + 1].z, r[2 * i + 1].w};
                        tc_st8(tmem_a + lane_taddr + s * SLOT_COLS + (uint32_t)(i * 8), v);
                    }
                    tc_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_ready(s));
#pragma unroll
                    for (int i = 0; i < SC * 4; ++i) r[i] = rn[i];
                }
            }
        }
    } else if (warp < MMA_WARP) {
        // ===================== epilogue =====================
        const int ew = warp - 4;
        const uint32_t half = (uint32_t)(ew >> 2);
        uint32_t it = 0;
        uint32_t* q_row = s_qrow + (MODE == 0 ? ew * TC_QUEUE : 0);
        uint32_t* q_q = s_qq + (MODE == 0 ? ew * TC_QUEUE : 0);
        uint32_t q_head = 0, q_count = 0, n_written = 0;
        uint2* my_list = recs + (size_t)(blockIdx.x * TC_EPI_WARPS + ew) * rec_cap;
        const uint32_t lane_lt = (1u << lane) - 1u;
        auto flush = [&](uint32_t n_take) {
            __syncwarp();
            if ((uint32_t)lane < n_take && n_written + lane < rec_cap) {
                const uint32_t at = (q_head + lane) % TC_QUEUE;
                my_list[n_written + lane] = make_uint2(q_row[at], q_q[at]);
            }
            n_written += n_take;
            q_head = (q_head + n_take) % TC_QUEUE;
            q_count -= n_take;
            __syncwarp();
        };
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            uint32_t qb, g_lo, g_hi;
            item_range(item, qb, g_lo, g_hi);
            if (MODE == 0) {
                asm volatile("bar.sync 1, %0;" :: "n"(32 * TC_EPI_WARPS) : "memory");
                for (uint32_t i = threadIdx.x - 128; i < TC_NQ; i += 32 * TC_EPI_WARPS) s_thr[i] = thr[qb * TC_NQ + i];
                asm volatile("bar.sync 1, %0;" :: "n"(32 * TC_EPI_WARPS) : "memory");
            }
            for (uint32_t g = g_lo; g < g_hi; ++g, ++it) {
                const uint32_t tile = g * 4u + (ew & 3);
                const bool in_range = tile < n_tiles;
                const uint32_t row = tile * 32u + lane;
                const bool alive = in_range && ((__ldg(live + (in_range ? tile : 0)) >> lane) & 1u);
                const float ri = (alive && row < n_rows) ? __ldg(rinv + row) : 0.0f;
                const uint32_t b = it % NBUF;
                const uint32_t qbase_q = qb * TC_NQ + half * 64;
                mbar_wait(acc_full(b), (it / NBUF) & 1u);
                tc_fence_after();
                uint32_t v[64];
                {
                    uint32_t v0[32], v1[32];
                    const uint32_t col0 = tmem_d + lane_taddr + b * TC_NQ + half * 64;
                    tc_ld32(col0, v0);
                    tc_ld32(col0 + 32, v1);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) { v[j] = v0[j]; v[32 + j] = v1[j]; }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(b));
                if (MODE == 0) {
                    // y = thr - dot * rinv: a sign bit means the row's approximate cosine beats the bound
                    const float4* c4 = reinterpret_cast<const float4*>(s_thr + half * 64);
                    uint64_t mask = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t y[16];
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 c = c4[4 * k + j4];
                            y[4 * j4 + 0] = __float_as_uint(fmaf(-__uint_as_float(v[16 * k + 4 * j4 + 0]), ri, c.x));
                            y[4 * j4 + 1] = __float_as_uint(fmaf(-__uint_as_float(v[16 * k + 4 * j4 + 1]), ri, c.y));
                            y[4 * j4 + 2] = __float_as_uint(fmaf(-__uint_as_float(v[16 * k + 4 * j4 + 2]), ri, c.z));
                            y[4 * j4 + 3] = __float_as_uint(fmaf(-__uint_as_float(v[16 * k + 4 * j4 + 3]), ri, c.w));
                        }
                        uint32_t o = y[0];
#pragma unroll
                        for (int j = 1; j < 15; j += 2) o |= y[j] | y[j + 1];
                        o |= y[15];
                        if (__any_sync(0xffffffffu, (int32_t)o < 0 && alive)) {
                            uint32_t mk = 0;
#pragma unroll
                            for (int j = 0; j < 16; ++j) mk = __funnelshift_l(y[j], mk, 1);
                            if (alive) mask |= (uint64_t)mk << (48 - 16 * k);
                        }
                    }
                    while (__any_sync(0xffffffffu, mask != 0)) {
                        const bool has = mask != 0;
                        const uint32_t m = __ballot_sync(0xffffffffu, has);
                        if (has) {
                            const int e = __clzll((long long)mask);
                            mask &= ~(0x8000000000000000ull >> e);
                            const uint32_t at = (q_head + q_count + __popc(m & lane_lt)) % TC_QUEUE;
                            q_row[at] = row;
                            q_q[at] = qbase_q + (uint32_t)e;
                        }
                        q_count += __popc(m);
                        if (q_count >= 32) flush(32);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 64; ++j) {
                        const uint32_t q = qbase_q + j;
                        if (q < nq && in_range && row < n_rows) dot_out[(size_t)q * dot_stride + row] = __uint_as_float(v[j]);
                    }
                }
            }
        }
        if (MODE == 0) {
            if (q_count) flush(q_count);
            if (lane == 0) {
                list_counts[blockIdx.x * TC_EPI_WARPS + ew] = min(n_written, rec_cap);
                if (n_written > rec_cap) *overflow = 1u;
            }
        }
    } else {
        // ===================== query loader (TMA) + MMA issuer =====================
        uint32_t bfull_phase = 0, bfree_phase = 0;
        uint32_t it = 0, p = 0;
        bool first_item = true;
        const uint32_t smem_base = smem_u32(smem);
        const uint64_t bdesc0 = tc_smem_desc(smem_base, 128, 1024);
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            uint32_t qb, g_lo, g_hi;
            item_range(item, qb, g_lo, g_hi);
            if (g_lo >= g_hi) continue;
            if (!first_item) { mbar_wait(b_free, bfree_phase); bfree_phase ^= 1u; }
            first_item = false;
            if (elect_one()) {
                constexpr uint32_t OP_BYTES = (uint32_t)((NU + 1) / 2) * TC_STAGE_BYTES;
                mbar_expect_tx(b_full, OP_BYTES);
                const int8_t* src = q16 + (size_t)qb * QBLOCK_BYTES;
                for (uint32_t off = 0; off < OP_BYTES; off += TC_STAGE_BYTES)
                    tma_bulk_g2s(smem_base + off, src + off, TC_STAGE_BYTES, b_full);
            }
            __syncwarp();
            mbar_wait(b_full, bfull_phase);
            bfull_phase ^= 1u;
            for (uint32_t g = g_lo; g < g_hi; ++g, ++it, p += PH) {
                const uint32_t b = it % NBUF;
                mbar_wait(acc_empty(b), ((it / NBUF) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_addr = tmem_d + b * TC_NQ;
#pragma unroll
                for (int ph = 0; ph < PH; ++ph) {
                    const uint32_t s = (p + ph) % NSLOT;
                    mbar_wait(a_ready(s), ((p + ph) / NSLOT) & 1u);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_addr = tmem_a + s * SLOT_COLS;
#pragma unroll
                        for (int kc = 0; kc < SC; ++kc) {
                            const int ks = ph * SC + kc;
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                tc_mma_bf16_ts(d_addr, a_addr + (uint32_t)((kc * 2 + j) * 8),
                                               bdesc0 + (uint64_t)(((ks / 2) * TC_STAGE_BYTES + ((ks % 2) * 4 + j * 2) * 128) >> 4),
                                               IDESC, (ks | j) != 0 ? 1u : 0u);
                        }
                        tc_commit(a_free(s));
                        if (ph == PH - 1) tc_commit(acc_full(b));
                    }
                    __syncwarp();
                }
            }
            if (elect_one()) tc_commit(b_free);
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        __syncwarp();
        tc_dealloc(0u, TC_TMEM_COLS);
    }
}

// ---- b*: the Hamming distance of the R-th candidate, per query ------------------------------------------------
// State of the search: c(T) = number of live rows with hamming < T is non-decreasing in T; b* is the largest T with
// c(T) < R.  lo / hi bracket it (c(lo) < R <= c(hi)); `T` is the threshold the next counting pass evaluates.
struct RatioState {
    uint32_t lo, c_lo;       // c(lo) < R                      (starts at 0, 0)
    uint32_t hi, c_hi;       // c(hi) >= R, or hi = K + 1 with c_hi = live rows (everything)
    uint32_t T, step;        // threshold under evaluation; outward step while one side is still its initial bound
    uint32_t done;           // 1: bstar / need are final
    uint32_t bstar, below;   // b* (K + 1: fewer than R live rows, every live row is a candidate); below = c(b*)
};

// First guess from a sample: dist[q * stride + i], i < n_sample (0xffffffff: no row), the compact output of the FP4
// scan's MODE 1 over strided row groups.  T = the smallest threshold whose sampled count reaches R * sample / rows.
// One CTA per query, histogram in shared memory.
__global__ void __launch_bounds__(256)
ratio_bstar_init_kernel(const uint32_t* __restrict__ dist, uint64_t stride, uint32_t n_sample, uint32_t nbins, uint64_t R,
                        uint64_t n_rows, RatioState* __restrict__ state, uint32_t* __restrict__ qpack, int qs, int tau_word) {
    extern __shared__ uint32_t sh[];
    const uint32_t q = blockIdx.x;
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) sh[i] = 0;
    __shared__ uint32_t s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_sample; i += blockDim.x) {
        const uint32_t d = dist[(size_t)q * stride + i];
        if (d != 0xffffffffu) { atomicAdd(&sh[min(d, nbins - 1)], 1u); atomicAdd(&s_n, 1u); }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double want = (double)R * (double)s_n / (double)max((uint64_t)1, n_rows);
        uint32_t run = 0, T = nbins;                       // nbins = K + 1
        for (uint32_t b = 0; b < nbins; ++b) {
            run += sh[b];
            if ((double)run >= want) { T = b + 1; break; }  // c_sample(b + 1) = rows with hamming <= b
        }
        RatioState st{0u, 0u, nbins, 0xffffffffu, min(max(T, 1u), nbins - 1), 1u, 0u, 0u, 0u};
        state[q] = st;
        qpack[(size_t)q * qs + tau_word] = st.T;
    }
}

// After a counting pass at state.T: narrow the bracket, pick the next threshold (step outwards while a side is
// unknown, bisect once both are known), finish when hi == lo + 1.  *n_active = queries still searching.
__global__ void ratio_bstar_update_kernel(RatioState* __restrict__ state, const uint32_t* __restrict__ counts, uint32_t nq,
                                          uint32_t R, uint32_t K, uint32_t* __restrict__ qpack, int qs, int tau_word,
                                          uint32_t* __restrict__ n_active) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    RatioState st = state[q];
    if (st.done) return;
    const uint32_t c = counts[q];
    if (c < R) { st.lo = st.T; st.c_lo = c; } else { st.hi = st.T; st.c_hi = c; }
    const bool hi_known = st.c_hi != 0xffffffffu;
    if (!hi_known && st.lo >= K + 1) {                      // even "everything" holds fewer than R live rows
        st.done = 1; st.bstar = K + 1; st.below = st.c_lo;
    } else if (hi_known && st.hi == st.lo + 1) {
        st.done = 1; st.bstar = st.lo; st.below = st.c_lo;
    } else {
        if (!hi_known) { st.T = min(st.lo + st.step, K + 1); st.step *= 2; }
        else if (st.lo == 0 && st.c_lo == 0 && st.step < 0x40000000u && st.hi > st.step) {   // lower side still the initial bound
            st.T = st.hi - min(st.step, st.hi - 1); st.step *= 2;
            if (st.T <= st.lo) st.T = st.lo + (st.hi - st.lo) / 2;
        } else st.T = st.lo + (st.hi - st.lo) / 2;
        if (st.T <= st.lo) st.T = st.lo + 1;
        if (hi_known && st.T >= st.hi) st.T = st.hi - 1;
        atomicAdd(n_active, 1u);
    }
    state[q] = st;
    qpack[(size_t)q * qs + tau_word] = st.T;
}

// thr[q] = (c_k - eps) * |q| for the dense filter, c_k = the k-th best cosine of the fast path (its lists are
// sorted); a query whose list is short (fewer than k rows) or whose bound is not positive goes to the fallback.
__global__ void ratio_threshold_kernel(const float* __restrict__ topk_scores, uint32_t k, const float* __restrict__ qnorm,
                                       uint32_t nq, uint32_t nq_pad, float eps, float* __restrict__ thr,
                                       uint32_t* __restrict__ fallback) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq_pad) return;
    float t = 3.0e38f;                                      // padding queries: nothing passes
    if (q < nq) {
        const float ck = topk_scores[(size_t)q * k + (k - 1)];
        if (!(ck > eps) || !(qnorm[q] > 0.0f)) { fallback[q] = 1u; }
        else t = (ck - eps) * qnorm[q];
    }
    thr[q] = t;
}

// Records (row, query) of the dense filter -> E lists.  The Hamming distance is recomputed from the codes; a record
// is dropped if the fast path already holds it (key <= the last key of its record list) or it is not a candidate
// (hamming > b*); a tie at b* or a full list sends the query to the exact fallback.  E keys: hamming << 40 | global row.
__global__ void __launch_bounds__(256)
ratio_scatter_kernel(const uint2* __restrict__ recs, uint32_t rec_cap, const uint32_t* __restrict__ list_counts,
                     const uint4* __restrict__ codes, int nchunk, const uint32_t* __restrict__ qpack, int qs,
                     const RatioState* __restrict__ state, const uint32_t* __restrict__ f_ham, const uint64_t* __restrict__ f_ids,
                     uint32_t Rf, uint64_t row_base, uint32_t K, uint32_t* __restrict__ e_cnt, uint64_t* __restrict__ e_keys,
                     uint32_t e_cap, uint32_t* __restrict__ fallback) {
    const uint32_t n_list = list_counts[blockIdx.y];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_list; i += gridDim.x * blockDim.x) {
        const uint2 r = recs[(size_t)blockIdx.y * rec_cap + i];
        const uint32_t row = r.x, q = r.y;
        const uint4* rc = codes + ((size_t)(row >> 5) * nchunk) * 32 + (row & 31);
        const uint4* qc = reinterpret_cast<const uint4*>(qpack + (size_t)q * qs);
        uint32_t d = 0;
        for (int c = 0; c < nchunk; ++c) {
            const uint4 a = rc[c * 32], b = qc[c];
            d += __popc(a.x ^ b.x) + __popc(a.y ^ b.y) + __popc(a.z ^ b.z) + __popc(a.w ^ b.w);
        }
        const uint64_t grow = row_base + row;
        const uint32_t fh = f_ham[(size_t)q * Rf + (Rf - 1)];
        const uint64_t fi = f_ids[(size_t)q * Rf + (Rf - 1)];
        if (fi == UINT64_MAX || d < fh || (d == fh && grow <= fi)) continue;         // one of the fast path's records
        const uint32_t bstar = state[q].bstar;
        uint64_t tie = 0;
        if (bstar <= K) {
            if (d > bstar) continue;                                                    // not a candidate
            // hamming == b*: a candidate or not depending on its place among the ties (the stable sort).  It is kept
            // PROVISIONALLY (bit 63); only if it makes the final top k does the query go to the exact fallback.
            if (d == bstar) tie = RATIO_TIE_BIT;
        }
        const uint32_t pos = atomicAdd(&e_cnt[q], 1u);
        if (pos < e_cap) e_keys[(size_t)q * e_cap + pos] = tie | ((uint64_t)d << 40) | grow;
        else fallback[q] = 1u;
    }
}

// Final order of one query: the fast path's records (stage-1 order) followed by the filter's survivors sorted by
// (hamming, row) are the candidates in stage-1 order; key = descending cosine image << 32 | that position, first k.
// One CTA per query; dynamic shared memory: e_cap u64 (E sort) + n_eff u64 (final sort).
__global__ void __launch_bounds__(1024)
ratio_finish_kernel(const uint64_t* __restrict__ f_ids, const float* __restrict__ f_score, uint32_t Rf,
                    const uint64_t* __restrict__ e_keys, const float* __restrict__ e_score, const uint32_t* __restrict__ e_cnt,
                    uint32_t e_cap, uint32_t n_eff, uint32_t* __restrict__ fallback, uint32_t k,
                    uint64_t* __restrict__ ids_out, float* __restrict__ scores_out) {
    extern __shared__ __align__(16) uint64_t rf_smem[];
    uint64_t* ekeys = rf_smem;                              // (key << 12 | slot) sorted: slot = position in e_keys
    uint64_t* skeys = rf_smem + e_cap;
    const uint32_t q = blockIdx.x;
    if (fallback[q]) return;
    const uint32_t ne = min(e_cnt[q], e_cap);
    const uint32_t ne_pow2 = max(32u, next_pow2(ne));
    for (uint32_t i = threadIdx.x; i < ne_pow2; i += blockDim.x)
        ekeys[i] = i < ne ? (((e_keys[(size_t)q * e_cap + i] & ~RATIO_TIE_BIT) << 12) | i) : UINT64_MAX;   // keys fit 52 bits (12 + 40)
    __syncthreads();
    if (ne > 1) bitonic_sort_smem(ekeys, ne_pow2);
    for (uint32_t i = threadIdx.x; i < n_eff; i += blockDim.x) {
        uint64_t key = UINT64_MAX;
        if (i < Rf) {
            if (f_ids[(size_t)q * Rf + i] != UINT64_MAX) key = ((uint64_t)(~f32_asc_key(f_score[(size_t)q * Rf + i])) << 32) | i;
        } else if (i - Rf < ne) {
            const uint32_t slot = (uint32_t)(ekeys[i - Rf] & 0xfffu);
            key = ((uint64_t)(~f32_asc_key(e_score[(size_t)q * e_cap + slot])) << 32) | i;
        }
        skeys[i] = key;
    }
    __syncthreads();
    bitonic_sort_smem(skeys, n_eff);
    for (uint32_t t = threadIdx.x; t < k; t += blockDim.x) {
        const uint64_t key = t < n_eff ? skeys[t] : UINT64_MAX;
        uint64_t id = UINT64_MAX;
        float sc = -INFINITY;
        if (key != UINT64_MAX) {
            const uint32_t pos = (uint32_t)key;
            if (pos < Rf) { id = f_ids[(size_t)q * Rf + pos]; sc = f_score[(size_t)q * Rf + pos]; }
            else {
                const uint32_t slot = (uint32_t)(ekeys[pos - Rf] & 0xfffu);
                const uint64_t ek = e_keys[(size_t)q * e_cap + slot];
                if (ek & RATIO_TIE_BIT) fallback[q] = 1u;              // an undecided tie made the top k: the exact path answers
                id = ek & ((1ull << 40) - 1);
                sc = e_score[(size_t)q * e_cap + slot];
            }
        }
        ids_out[(size_t)q * k + t] = id;
        scores_out[(size_t)q * k + t] = sc;
    }
}

}  // namespace gvdb
