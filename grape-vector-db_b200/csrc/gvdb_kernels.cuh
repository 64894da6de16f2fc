// gvdb_kernels.cuh — sm_100a kernels of the quantized-search path.
//
//   ingest_kernel     f32 rows -> 1-bit codes (blocked layout) + sequential L2 norms
//                     [BinaryQuantizer::quantize, /root/reference/src/quantization.rs:97-101]
//   scan_kernel       1-bit Hamming scan, XOR + popc, 128-bit coalesced loads, query codes
//                     staged in shared memory by TMA bulk copy, fused threshold filter
//                     [hamming_distance/similarity + the stage-1 loop, quantization.rs:130-148,165-172]
//   select_kernel     exact top-R by the unique key (hamming << 32 | row): block bitonic sort
//                     [the stable sort + slice of quantization.rs:175-179]
//   rescore_kernel    exact-order f32 cosine of the kept candidates (bit-identical folds)
//                     [cosine_similarity_manual, quantization.rs:206-216]
//   topk_kernel       order by (cosine desc, hamming asc, row asc), emit k
//                     [the stable sort of quantization.rs:190]
//   merge_select_kernel  cross-shard stage-1 cut over gathered records
//                     [replaces concat+sort+truncate, src/distributed/shard.rs:776-783]
//
// HBM layout of the codes ("blocked"): rows are grouped in tiles of 32; a row's code is
// NCHUNK 16-byte chunks; chunk c of row r of tile t lives at uint4 index
// (t*NCHUNK + c)*32 + r.  A warp that owns a tile therefore reads every chunk as one fully
// coalesced 512-byte request (LDG.128 per lane), and a tile is NCHUNK*512 contiguous bytes.
// Inside a chunk the bytes are the reference's BinaryVector::to_bytes() bytes
// (bit j -> byte j/8, bit 7-(j%8)); Hamming distance is invariant to that choice.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gvdb {

constexpr uint32_t TAU_ALL = 0xffffffffu;   // "emit every live row"
// Per-query candidate counters live one per 128-byte line (cnt[q * CNT_STRIDE]): appends are L2
// atomics, which serialise per line; packed counters put 1024 hot addresses on 32 lines.
constexpr int CNT_STRIDE = 32;
constexpr int SCAN_THREADS = 256;           // 8 warps = 8 tiles of 32 rows in flight per CTA
constexpr int SORT_N = 4096;                // block bitonic capacity (u64 keys, 32 KB smem)
constexpr int SORT_THREADS = 1024;
constexpr int STAGE_ROWS = 128;             // rows per CTA in the transposing kernels

// ---------------------------------------------------------------------------------------
// small PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {   // read-once data: skip L1
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Programmatic dependent launch (the hot single-pass chain is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization): a kernel lets its successor's CTAs be scheduled as soon as every
// CTA of its own has started, and waits for its predecessor to have completed (and flushed) before it touches global
// memory.  Both are no-ops in a normally launched kernel.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Orderable image of an f32: ascending u32 order == ascending float order; -0.0 and +0.0
// share one image (Rust's partial_cmp calls them Equal, so they must tie).
__device__ __forceinline__ uint32_t f32_asc_key(float f) {
    if (f == 0.0f) f = 0.0f;   // canonicalise -0.0
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---------------------------------------------------------------------------------------
// Block-wide bitonic sort of n (power of two, <= SORT_N) u64 keys in shared memory,
// ascending.  Keys are unique by construction, so the result is deterministic.
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* s, uint32_t n) {
    for (uint32_t k = 2; k <= n; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                uint32_t l = i | j;
                bool up = (i & k) == 0;
                uint64_t a = s[i], b = s[l];
                if ((a > b) == up) { s[i] = b; s[l] = a; }
            }
            __syncthreads();
        }
    }
}
__device__ __forceinline__ uint32_t next_pow2(uint32_t v) {
    return v <= 1 ? 1u : 1u << (32 - __clz(v - 1));
}

// ---------------------------------------------------------------------------------------
// ingest_kernel: one thread per row, 32-column slabs transposed through shared memory so
// global reads are full 128-byte lines and each thread still walks ITS row left to right
// (the norm is a sequential f32 fold: sum = sum + x*x, mul and add rounded separately).
// QUERY=false: writes blocked codes + norms + live bits for corpus rows.
// QUERY=true : writes qpack[i*qs .. ] = code words, then [tau=TAU_ALL, 0, 0, 0], and qnorm.
template <bool QUERY>
__global__ void __launch_bounds__(STAGE_ROWS)
ingest_kernel(const float* __restrict__ x, uint64_t n, int dim, float thr, int nchunk,
              uint64_t first_row, uint4* __restrict__ codes, float* __restrict__ norms,
              uint32_t* __restrict__ live, uint32_t* __restrict__ qpack, int qs) {
    __shared__ float tile[STAGE_ROWS][33];
    const int tid = threadIdx.x;
    const uint64_t i0 = (uint64_t)blockIdx.x * STAGE_ROWS;
    const uint64_t i = i0 + tid;
    const bool valid = i < n;
    const bool vec4 = (dim & 3) == 0;
    float ss = 0.0f;
    const uint64_t g = first_row + i;
    for (int c = 0; c < nchunk; ++c) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        for (int sub = 0; sub < 4; ++sub) {
            const int j0 = (c * 4 + sub) * 32;
            if (j0 < dim) {   // block-uniform
                __syncthreads();
                if (vec4) {
                    for (int idx = tid; idx < STAGE_ROWS * 8; idx += STAGE_ROWS) {
                        int r = idx >> 3, seg = idx & 7;
                        int j = j0 + seg * 4;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (i0 + r < n && j < dim)
                            v = *reinterpret_cast<const float4*>(x + (i0 + r) * (uint64_t)dim + j);
                        tile[r][seg * 4 + 0] = v.x; tile[r][seg * 4 + 1] = v.y;
                        tile[r][seg * 4 + 2] = v.z; tile[r][seg * 4 + 3] = v.w;
                    }
                } else {
                    for (int idx = tid; idx < STAGE_ROWS * 32; idx += STAGE_ROWS) {
                        int r = idx >> 5, e = idx & 31;
                        int j = j0 + e;
                        tile[r][e] = (i0 + r < n && j < dim) ? x[(i0 + r) * (uint64_t)dim + j] : 0.f;
                    }
                }
                __syncthreads();
                const int lim = min(32, dim - j0);
                uint32_t word = 0;
                for (int e = 0; e < lim; ++e) {
                    float v = tile[tid][e];
                    ss = __fadd_rn(ss, __fmul_rn(v, v));
                    // Msb0 inside each byte, bytes in memory order (little-endian word)
                    word |= (uint32_t)(v > thr) << ((e & ~7) | (7 - (e & 7)));
                }
                w[sub] = word;
            }
        }
        if (valid) {
            uint4 pk = make_uint4(w[0], w[1], w[2], w[3]);
            if (QUERY) *reinterpret_cast<uint4*>(qpack + i * (uint64_t)qs + c * 4) = pk;
            else codes[((g >> 5) * (uint64_t)nchunk + c) * 32 + (g & 31)] = pk;
        }
    }
    if (valid) {
        norms[QUERY ? i : g] = __fsqrt_rn(ss);
        if (QUERY) *reinterpret_cast<uint4*>(qpack + i * (uint64_t)qs + nchunk * 4) =
                       make_uint4(TAU_ALL, 0u, 0u, 0u);
        else atomicOr(&live[g >> 5], 1u << (g & 31));
    }
}

// Query codes supplied by the caller (reference byte layout, nbytes per query) -> qpack.
__global__ void pack_query_codes_kernel(const uint8_t* __restrict__ qcodes, uint32_t nq,
                                        int nbytes, int nchunk, uint32_t* __restrict__ qpack,
                                        int qs) {
    uint32_t q = blockIdx.x;
    if (q >= nq) return;
    for (int w = threadIdx.x; w < qs; w += blockDim.x) {
        uint32_t v = 0;
        if (w < nchunk * 4) {
            for (int b = 0; b < 4; ++b) {
                int byte = w * 4 + b;
                if (byte < nbytes) v |= (uint32_t)qcodes[(uint64_t)q * nbytes + byte] << (8 * b);
            }
        } else if (w == nchunk * 4) v = TAU_ALL;
        qpack[(uint64_t)q * qs + w] = v;
    }
}

// blocked codes -> reference byte layout (for gvdb_get_codes / parity)
__global__ void unblock_codes_kernel(const uint4* __restrict__ codes, int nchunk, uint64_t first,
                                     uint64_t n, int nbytes, uint8_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t g = first + i;
    for (int c = 0; c < nchunk; ++c) {
        uint4 v = codes[((g >> 5) * (uint64_t)nchunk + c) * 32 + (g & 31)];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        for (int b = 0; b < 16; ++b) {
            int byte = c * 16 + b;
            if (byte < nbytes) out[i * (uint64_t)nbytes + byte] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
        }
    }
}

// reference byte layout (nbytes per row, rows [first, first+n)) -> blocked codes (gvdb_load)
__global__ void block_codes_kernel(const uint8_t* __restrict__ in, int nchunk, uint64_t first, uint64_t n,
                                   int nbytes, uint4* __restrict__ codes) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t g = first + i;
    for (int c = 0; c < nchunk; ++c) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        for (int b = 0; b < 16; ++b) {
            int byte = c * 16 + b;
            if (byte < nbytes) w[b >> 2] |= (uint32_t)in[i * (uint64_t)nbytes + byte] << (8 * (b & 3));
        }
        codes[((g >> 5) * (uint64_t)nchunk + c) * 32 + (g & 31)] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---------------------------------------------------------------------------------------
// Hamming distance of W xor-ed words with NCSA carry-save adders in front of the popcounts.
// B200 issues 16 popc/clk/SM (XU pipe) but 64 LOP3/clk/SM (ALU pipe); the plain loop is
// XU-bound.  A full adder (sum = a^b^c, carry = maj(a,b,c): two LOP3) turns three words of
// weight w into one of weight w and one of weight 2w, i.e. removes one popc for two LOP3.
// NCSA adders are spent greedily on the lowest weight class that still has three words;
// the result is exact: d = P1 + 2*P2 + 4*P4 with Pw the popcount sum of weight class w.
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

template <int W, int NCSA>
__device__ __forceinline__ uint32_t hamming_csa(const uint32_t (&x)[W]) {
    // stage A: adders over consecutive triples of the inputs (weight 1 -> 1 + 2)
    constexpr int nA = (NCSA < W / 3) ? NCSA : W / 3;
    constexpr int n1 = nA + (W - 3 * nA);           // weight-1 words after stage A
    uint32_t w1[n1 > 0 ? n1 : 1];
    uint32_t w2a[nA > 0 ? nA : 1];
#pragma unroll
    for (int i = 0; i < nA; ++i) {
        w1[i] = lop3_xor3(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
        w2a[i] = lop3_maj(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
    }
#pragma unroll
    for (int i = 3 * nA; i < W; ++i) w1[nA + i - 3 * nA] = x[i];
    // stage B: adders over triples of the weight-1 words
    constexpr int remB = NCSA - nA;
    constexpr int nB = (remB < n1 / 3) ? remB : n1 / 3;
    constexpr int n1b = nB + (n1 - 3 * nB);
    uint32_t v1[n1b > 0 ? n1b : 1];
    uint32_t w2b[nB > 0 ? nB : 1];
#pragma unroll
    for (int i = 0; i < nB; ++i) {
        v1[i] = lop3_xor3(w1[3 * i], w1[3 * i + 1], w1[3 * i + 2]);
        w2b[i] = lop3_maj(w1[3 * i], w1[3 * i + 1], w1[3 * i + 2]);
    }
#pragma unroll
    for (int i = 3 * nB; i < n1; ++i) v1[nB + i - 3 * nB] = w1[i];
    // stage C: adders over triples of the weight-2 words (2 -> 2 + 4)
    constexpr int n2 = nA + nB;
    uint32_t w2[n2 > 0 ? n2 : 1];
#pragma unroll
    for (int i = 0; i < nA; ++i) w2[i] = w2a[i];
#pragma unroll
    for (int i = 0; i < nB; ++i) w2[nA + i] = w2b[i];
    constexpr int remC = NCSA - nA - nB;
    constexpr int nC = (remC < n2 / 3) ? remC : n2 / 3;
    uint32_t p1 = 0, p2 = 0, p4 = 0;
#pragma unroll
    for (int i = 0; i < n1b; ++i) p1 += __popc(v1[i]);
#pragma unroll
    for (int i = 0; i < nC; ++i) {
        p2 += __popc(lop3_xor3(w2[3 * i], w2[3 * i + 1], w2[3 * i + 2]));
        p4 += __popc(lop3_maj(w2[3 * i], w2[3 * i + 1], w2[3 * i + 2]));
    }
#pragma unroll
    for (int i = 3 * nC; i < n2; ++i) p2 += __popc(w2[i]);
    return p1 + 2u * p2 + 4u * p4;
}

// ---------------------------------------------------------------------------------------
// scan_kernel — THE hot loop.  One warp owns a tile of 32 rows: each lane keeps its row's
// code in NCHUNK uint4 registers (one coalesced 512 B request per chunk) and walks the query
// group staged in shared memory (one TMA bulk copy per CTA; every lane reads the same
// address, i.e. a broadcast LDS.128).  Per (row, query): NCHUNK*4 XOR + popc.
// MODE 0 (search): rows that beat the query's current threshold are appended to the query's
//                  candidate buffer as key = hamming << 32 | row; distances never touch HBM.
// MODE 1 (parity): every distance is written out (gvdb_hamming).
// Algorithmic bytes per launch: (tile_hi - tile_lo) * 32 * NCHUNK * 16 (codes streamed once
// per query group) — see DESIGN.md §5.
template <int NCHUNK, int MODE, int NCSA, bool AGG>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const uint4* __restrict__ codes, const uint32_t* __restrict__ live, uint32_t tile_lo,
            uint32_t tile_hi, const uint32_t* __restrict__ qpack, int nq, int qgroup,
            uint32_t* __restrict__ cnt, uint64_t* __restrict__ buf, uint32_t cap,
            uint32_t* __restrict__ overflow, uint32_t* __restrict__ dist_out,
            uint64_t dist_stride, uint64_t n_rows) {
    constexpr int QS = NCHUNK * 4 + 4;   // words per staged query: code, then [tau,0,0,0]
    extern __shared__ __align__(128) uint32_t sq[];
    __shared__ __align__(8) uint64_t mbar;

    const int q0 = blockIdx.y * qgroup;
    const int nql = min(qgroup, nq - q0);
    const uint32_t bar = smem_u32(&mbar);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)nql * QS * 4u;
        mbar_expect_tx(bar, bytes);
        tma_bulk_g2s(smem_u32(sq), qpack + (size_t)q0 * QS, bytes, bar);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WARPS = SCAN_THREADS / 32;
    bool staged = false;

    for (uint32_t tile = tile_lo + blockIdx.x * WARPS + warp; tile < tile_hi;
         tile += gridDim.x * WARPS) {
        uint4 r[NCHUNK];
        const uint4* base = codes + ((size_t)tile * NCHUNK) * 32 + lane;
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) r[c] = ldg_stream(base + c * 32);
        const bool alive = (live[tile] >> lane) & 1u;
        const uint32_t row = tile * 32u + lane;
        if (!staged) {                       // row loads are in flight while the TMA lands
            while (!mbar_try_wait(bar, 0)) {}
            staged = true;
        }
#pragma unroll 2
        for (int q = 0; q < nql; ++q) {
            const uint4* qp = reinterpret_cast<const uint4*>(sq + q * QS);
            uint32_t x[NCHUNK * 4];
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                const uint4 v = qp[c];
                x[4 * c + 0] = r[c].x ^ v.x; x[4 * c + 1] = r[c].y ^ v.y;
                x[4 * c + 2] = r[c].z ^ v.z; x[4 * c + 3] = r[c].w ^ v.w;
            }
            const uint32_t d = hamming_csa<NCHUNK * 4, NCSA>(x);
            if (MODE == 0) {
                const uint32_t tau = sq[q * QS + NCHUNK * 4];
                // strict '<': tau is the R-th smallest distance over EARLIER rows, so a later
                // row that ties it has a larger row number and cannot enter the top R.
                const bool hit = alive && d < tau;
                if (AGG) {
                    // warp-aggregated append: one atomic per (tile, query) instead of one per row
                    const uint32_t m = __ballot_sync(0xffffffffu, hit);
                    if (m) {
                        uint32_t base = 0;
                        const int leader = __ffs(m) - 1;
                        if (lane == leader) base = atomicAdd(&cnt[(size_t)(q0 + q) * CNT_STRIDE], (uint32_t)__popc(m));
                        base = __shfl_sync(0xffffffffu, base, leader);
                        if (hit) {
                            const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
                            if (pos < cap) buf[(size_t)(q0 + q) * cap + pos] = ((uint64_t)d << 32) | row;
                            else *overflow = 1u;
                        }
                    }
                } else if (hit) {
                    const uint32_t pos = atomicAdd(&cnt[(size_t)(q0 + q) * CNT_STRIDE], 1u);
                    if (pos < cap) buf[(size_t)(q0 + q) * cap + pos] = ((uint64_t)d << 32) | row;
                    else *overflow = 1u;
                }
            } else {
                if (row < n_rows) dist_out[(size_t)(q0 + q) * dist_stride + row] = d;
            }
        }
    }
    if (!staged) { while (!mbar_try_wait(bar, 0)) {} }   // never exit with a copy in flight
}

// ---------------------------------------------------------------------------------------
// select_kernel: one CTA per query.  Streams the query's candidate keys through a
// SORT_N-wide block bitonic sort, keeping the R smallest; writes them back sorted, sets
// cnt = kept and the query's next threshold (R-th smallest distance, or TAU_ALL while
// fewer than R live rows have been seen).  R <= SORT_N/2.
__global__ void __launch_bounds__(SORT_THREADS)
select_kernel(uint64_t* __restrict__ buf, uint32_t cap, uint32_t* __restrict__ cnt, uint32_t R,
              uint32_t* __restrict__ qpack, int qs, int tau_word) {
    extern __shared__ __align__(16) uint64_t skeys[];
    const uint32_t q = blockIdx.x;
    uint64_t* mine = buf + (size_t)q * cap;
    const uint32_t n_in = min(cnt[(size_t)q * CNT_STRIDE], cap);
    uint32_t have = 0, consumed = 0;
    while (consumed < n_in) {
        const uint32_t take = min((uint32_t)SORT_N - have, n_in - consumed);
        const uint32_t total = have + take;
        const uint32_t n_eff = max(64u, next_pow2(total));
        for (uint32_t i = threadIdx.x; i < n_eff - have; i += blockDim.x)
            skeys[have + i] = i < take ? mine[consumed + i] : UINT64_MAX;
        __syncthreads();
        bitonic_sort_smem(skeys, n_eff);
        have = min(R, total);
        consumed += take;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < have; i += blockDim.x) mine[i] = skeys[i];
    if (threadIdx.x == 0) {
        cnt[(size_t)q * CNT_STRIDE] = have;
        qpack[(size_t)q * qs + tau_word] = (have >= R && R > 0) ? (uint32_t)(skeys[R - 1] >> 32) : TAU_ALL;
    }
}

// ---------------------------------------------------------------------------------------
// select_hist_kernel: the same contract as select_kernel for keys whose high word is a Hamming
// distance (a small integer < nbins): exact top-R by (hamming, row) without sorting the input.
//   1. histogram of the distances in shared memory, prefix scan -> threshold bin b* (the bin that
//      holds the R-th smallest key) and how many keys of that bin are still needed;
//   2. keys below b* are kept outright; the keys IN b* compete on their row number: a small tie
//      set is sorted in shared memory, a large one goes through a 4 x 8-bit radix select on the row;
//   3. the R kept keys are sorted (bitonic over next_pow2(R)) and written back.
// Dynamic shared memory: r_pow2 u64 (kept) + SORT_N u64 (ties) + nbins u32 (histogram).
constexpr int SELH_THREADS = 256;

// Threshold for the next scan segment after a select that kept `have` sorted keys.
//   tau[0]  = the R-th smallest distance so far (TAU_ALL while fewer than R rows were seen): only
//             rows strictly below it can still enter the top R.
//   opt_m   > 0 (optimistic single pass): tau[0] = min(tau[0], d_(m) + 1) with d_(m) the m-th
//             smallest distance of the rows seen so far, remembered in tau[1].  If the rows seen so
//             far are a fair sample, about m * (rows left / rows seen) rows lie below it.  The
//             guess is checked, not trusted:
//   verify  1: after the pass: the top R is exact iff at least R candidates lie strictly below tau[1]
//             (then every row of the true top R was emitted); otherwise flags[1] asks for a rerun.
//           2: single pass over ALL rows under tau[1] (tc_tau_kernel): every candidate lies below
//             tau[1], so the top R is exact iff R of them were emitted, or tau[1] let everything pass.
__device__ __forceinline__ void select_set_tau(const uint64_t* sel, uint32_t have, uint32_t R, uint32_t* tau,
                                               uint32_t opt_m, int verify, uint32_t* flags) {
    const uint32_t tau_r = (have >= R && R > 0) ? (uint32_t)(sel[R - 1] >> 32) : TAU_ALL;
    if (verify == 2) {
        if (!(have >= R || tau[1] == TAU_ALL)) flags[1] = 1u;
        tau[0] = tau_r;
        return;
    }
    if (verify) {
        if (!(have >= R && tau_r < tau[1])) flags[1] = 1u;
        tau[0] = tau_r;
        return;
    }
    uint32_t t = tau_r;
    if (opt_m > 0 && have >= opt_m) {
        const uint32_t tau_opt = (uint32_t)(sel[opt_m - 1] >> 32) + 1u;
        tau[1] = tau_opt;
        t = min(t, tau_opt);
    }
    tau[0] = t;
}

__global__ void __launch_bounds__(SELH_THREADS)
select_hist_kernel(uint64_t* __restrict__ buf, uint32_t cap, uint32_t* __restrict__ cnt, uint32_t R,
                   uint32_t r_pow2, uint32_t nbins, uint32_t* __restrict__ qpack, int qs, int tau_word,
                   uint32_t opt_m, int verify, uint32_t* __restrict__ flags, uint32_t tie_cap) {
    // tie_cap (a power of two in [256, SORT_N]): keys of the threshold bin that are ordered in shared memory; a larger
    // tie set goes through the radix select.  The single-pass search expects a few hundred keys per query and asks for
    // 1024: the CTA then needs 12 KB instead of 36 KB and all 1024 queries of a tile are resident at once.
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) uint64_t sel[];
    uint64_t* tie = sel + r_pow2;
    uint32_t* hist = reinterpret_cast<uint32_t*>(tie + tie_cap);
    __shared__ uint32_t s_bstar, s_below, s_nsel, s_ntie, s_prefix, s_need;
    const uint32_t q = blockIdx.x;
    uint64_t* mine = buf + (size_t)q * cap;
    const uint32_t n = min(cnt[(size_t)q * CNT_STRIDE], cap);
    const uint32_t tid = threadIdx.x;

    if (n <= R) {                                   // everything is kept: just order it
        for (uint32_t i = tid; i < r_pow2; i += SELH_THREADS) sel[i] = i < n ? mine[i] : UINT64_MAX;
        __syncthreads();
        bitonic_sort_smem(sel, r_pow2);
        for (uint32_t i = tid; i < n; i += SELH_THREADS) mine[i] = sel[i];
        if (tid == 0) {
            cnt[(size_t)q * CNT_STRIDE] = n;
            select_set_tau(sel, n, R, qpack + (size_t)q * qs + tau_word, opt_m, verify, flags);
        }
        return;
    }
    for (uint32_t i = tid; i < nbins; i += SELH_THREADS) hist[i] = 0;
    if (tid == 0) { s_nsel = 0; s_ntie = 0; }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += SELH_THREADS)
        atomicAdd(&hist[min((uint32_t)(mine[i] >> 32), nbins - 1)], 1u);
    __syncthreads();
    if (tid < 32) {                                 // warp 0: which bin holds the R-th smallest key?
        const uint32_t per = (nbins + 31) / 32;
        const uint32_t b0 = tid * per, b1 = min(nbins, b0 + per);
        uint32_t sum = 0;
        for (uint32_t b = b0; b < b1; ++b) sum += hist[b];
        uint32_t incl = sum;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)tid >= o) incl += t;
        }
        uint32_t run = incl - sum;                  // keys in bins before b0
        if (run < R && incl >= R) {                 // exactly one lane
            for (uint32_t b = b0; b < b1; ++b) {
                const uint32_t h = hist[b];
                if (run + h >= R) { s_bstar = b; s_below = run; s_need = R - run; break; }
                run += h;
            }
        }
    }
    __syncthreads();
    const uint32_t bstar = s_bstar, below = s_below, need = s_need;
    const uint32_t ntie_all = hist[bstar];
    const bool small_ties = ntie_all <= tie_cap;
    for (uint32_t i = tid; i < n; i += SELH_THREADS) {
        const uint64_t key = mine[i];
        const uint32_t h = min((uint32_t)(key >> 32), nbins - 1);
        if (h < bstar) sel[atomicAdd(&s_nsel, 1u)] = key;
        else if (h == bstar && small_ties) tie[atomicAdd(&s_ntie, 1u)] = key;
    }
    __syncthreads();
    if (small_ties) {
        if (ntie_all > need) {                      // order the ties by row, keep the first `need`
            const uint32_t n_eff = max(32u, next_pow2(ntie_all));
            for (uint32_t i = ntie_all + tid; i < n_eff; i += SELH_THREADS) tie[i] = UINT64_MAX;
            __syncthreads();
            bitonic_sort_smem(tie, n_eff);
        }
    } else {
        // radix select on the row number among the keys of bin b*: find the need-th smallest row
        uint32_t prefix = 0, want = need;           // rows matching `prefix` in the bits decided so far
        uint32_t* h256 = reinterpret_cast<uint32_t*>(tie);
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (uint32_t i = tid; i < 256; i += SELH_THREADS) h256[i] = 0;
            __syncthreads();
            const uint32_t hi_mask = shift == 24 ? 0u : ~((1u << (shift + 8)) - 1u);
            for (uint32_t i = tid; i < n; i += SELH_THREADS) {
                const uint64_t key = mine[i];
                const uint32_t row = (uint32_t)key;
                if (min((uint32_t)(key >> 32), nbins - 1) == bstar && (row & hi_mask) == prefix)
                    atomicAdd(&h256[(row >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t run = 0, d = 0;
                for (; d < 256; ++d) { if (run + h256[d] >= want) break; run += h256[d]; }
                s_prefix = prefix | (d << shift);
                s_need = want - run;
            }
            __syncthreads();
            prefix = s_prefix; want = s_need;
            __syncthreads();
        }
        // rows are unique, so exactly `need` keys of the bin have row <= prefix
        for (uint32_t i = tid; i < n; i += SELH_THREADS) {
            const uint64_t key = mine[i];
            if (min((uint32_t)(key >> 32), nbins - 1) == bstar && (uint32_t)key <= prefix)
                sel[below + atomicAdd(&s_ntie, 1u)] = key;
        }
        __syncthreads();
    }
    if (small_ties)
        for (uint32_t i = tid; i < need; i += SELH_THREADS) sel[below + i] = tie[i];
    for (uint32_t i = R + tid; i < r_pow2; i += SELH_THREADS) sel[i] = UINT64_MAX;
    __syncthreads();
    bitonic_sort_smem(sel, r_pow2);
    for (uint32_t i = tid; i < R; i += SELH_THREADS) mine[i] = sel[i];
    if (tid == 0) {
        cnt[(size_t)q * CNT_STRIDE] = R;
        select_set_tau(sel, R, R, qpack + (size_t)q * qs + tau_word, opt_m, verify, flags);
    }
}

// ---------------------------------------------------------------------------------------
// rescore_kernel: one thread per (query, candidate).  32-column slabs of the candidate's row
// and of its query are transposed through shared memory; each thread then folds its own
// pair strictly left to right:  dot = dot + q[j]*c[j]  (separate roundings, never FMA),
// which is bit-identical to the reference's iterator sum.  ||q|| and ||c|| are the same
// sequential folds, computed once (prep / ingest) instead of once per pair.
// Outputs the shard record (hamming, global row, cosine) in stage-1 order.
__global__ void __launch_bounds__(STAGE_ROWS)
rescore_kernel(const float* __restrict__ rows, const float* __restrict__ norms, uint64_t row_base,
               int dim, const float* __restrict__ queries, const float* __restrict__ qnorm,
               const uint64_t* __restrict__ buf, uint32_t cap, const uint32_t* __restrict__ cnt,
               uint32_t R, uint32_t nq, uint32_t* __restrict__ out_ham,
               uint64_t* __restrict__ out_ids, float* __restrict__ out_score) {
    __shared__ float tc[STAGE_ROWS][33];
    __shared__ float tq[STAGE_ROWS][33];
    __shared__ uint32_t s_row[STAGE_ROWS];
    __shared__ uint32_t s_q[STAGE_ROWS];
    const int tid = threadIdx.x;
    const uint64_t p = (uint64_t)blockIdx.x * STAGE_ROWS + tid;
    const uint32_t q = (uint32_t)(p / R), r = (uint32_t)(p % R);
    const bool slot = q < nq;
    const bool valid = slot && r < cnt[(size_t)q * CNT_STRIDE];
    uint64_t key = valid ? buf[(size_t)q * cap + r] : UINT64_MAX;
    s_row[tid] = valid ? (uint32_t)key : 0xffffffffu;
    s_q[tid] = valid ? q : 0xffffffffu;
    const bool vec4 = (dim & 3) == 0;
    float dot = 0.0f;
    for (int j0 = 0; j0 < dim; j0 += 32) {
        __syncthreads();
        if (vec4) {
            for (int idx = tid; idx < STAGE_ROWS * 8; idx += STAGE_ROWS) {
                int rr = idx >> 3, seg = idx & 7;
                int j = j0 + seg * 4;
                uint32_t crow = s_row[rr], cq = s_q[rr];
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                if (crow != 0xffffffffu && j < dim) {
                    a = *reinterpret_cast<const float4*>(rows + (uint64_t)crow * dim + j);
                    b = *reinterpret_cast<const float4*>(queries + (uint64_t)cq * dim + j);
                }
                tc[rr][seg * 4 + 0] = a.x; tc[rr][seg * 4 + 1] = a.y;
                tc[rr][seg * 4 + 2] = a.z; tc[rr][seg * 4 + 3] = a.w;
                tq[rr][seg * 4 + 0] = b.x; tq[rr][seg * 4 + 1] = b.y;
                tq[rr][seg * 4 + 2] = b.z; tq[rr][seg * 4 + 3] = b.w;
            }
        } else {
            for (int idx = tid; idx < STAGE_ROWS * 32; idx += STAGE_ROWS) {
                int rr = idx >> 5, e = idx & 31;
                int j = j0 + e;
                uint32_t crow = s_row[rr], cq = s_q[rr];
                bool ok = crow != 0xffffffffu && j < dim;
                tc[rr][e] = ok ? rows[(uint64_t)crow * dim + j] : 0.f;
                tq[rr][e] = ok ? queries[(uint64_t)cq * dim + j] : 0.f;
            }
        }
        __syncthreads();
        const int lim = min(32, dim - j0);
        for (int e = 0; e < lim; ++e) dot = __fadd_rn(dot, __fmul_rn(tq[tid][e], tc[tid][e]));
    }
    if (slot) {
        float cosv = -INFINITY;
        if (valid) {
            const float na = qnorm[q], nb = norms[(uint32_t)key];
            cosv = (na == 0.0f || nb == 0.0f) ? 0.0f : __fdiv_rn(dot, __fmul_rn(na, nb));
        }
        out_ham[p] = valid ? (uint32_t)(key >> 32) : 0xffffffffu;
        out_ids[p] = valid ? row_base + (uint32_t)key : UINT64_MAX;
        out_score[p] = cosv;
    }
}

// ---------------------------------------------------------------------------------------
// Direct variants for dim % 4 == 0 (every BASELINE config): no shared-memory transposition.
// Each thread streams ITS row with 128-bit loads, DEPTH loads in flight (the loads are
// independent; only the f32 adds form the sequential chain the reference's fold dictates),
// so the kernels run at memory-system speed instead of one global-latency per 32-column slab.
constexpr int DIRECT_DEPTH = 8;      // float4 loads in flight per operand per thread

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// queries -> qpack (code words, then [tau = TAU_ALL, 0, 0, 0]) + sequential-fold norms.
// Same outputs as ingest_kernel<true>; one thread per query, 32 queries per CTA.
// 32 float4 in flight per thread: a 768-d query is six global-memory round trips (was 24 with
// DIRECT_DEPTH = 8; the kernel is latency-bound, 1024 threads on the whole GPU)
constexpr int QPREP_DEPTH = 32;

__global__ void __launch_bounds__(32)
query_prep_direct_kernel(const float* __restrict__ x, uint32_t n, int dim, float thr, int nchunk,
                         float* __restrict__ norms, uint32_t* __restrict__ qpack, int qs) {
    const uint32_t i = blockIdx.x * 32u + threadIdx.x;
    if (i >= n) return;
    const float* row = x + (size_t)i * dim;
    const int nv = dim >> 2;                       // float4 per row
    float ss = 0.0f;
    uint32_t word = 0;
    uint32_t* out = qpack + (size_t)i * qs;
    for (int v0 = 0; v0 < nv; v0 += QPREP_DEPTH) {
        float4 buf[QPREP_DEPTH];
#pragma unroll
        for (int u = 0; u < QPREP_DEPTH; ++u)
            buf[u] = v0 + u < nv ? ldg_f4(row + 4 * (v0 + u)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < QPREP_DEPTH; ++u) {
            if (v0 + u < nv) {
                const float e4[4] = {buf[u].x, buf[u].y, buf[u].z, buf[u].w};
                const int j0 = 4 * (v0 + u);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float v = e4[t];
                    ss = __fadd_rn(ss, __fmul_rn(v, v));
                    const int e = (j0 + t) & 31;
                    // Msb0 inside each byte, bytes in memory order (little-endian word)
                    word |= (uint32_t)(v > thr) << ((e & ~7) | (7 - (e & 7)));
                }
                if (((j0 + 4) & 31) == 0) { out[j0 >> 5] = word; word = 0; }
            }
        }
    }
    const int full_words = dim >> 5;
    if (dim & 31) out[full_words] = word;          // trailing partial word (pad bits 0)
    for (int w = (dim + 31) >> 5; w < nchunk * 4; ++w) out[w] = 0u;
    out[nchunk * 4] = TAU_ALL; out[nchunk * 4 + 1] = 0u; out[nchunk * 4 + 2] = 0u; out[nchunk * 4 + 3] = 0u;
    norms[i] = __fsqrt_rn(ss);
}

// ---------------------------------------------------------------------------------------
// rescore_slab_kernel (dim % 4 == 0): one WARP per 32 (query, candidate) pairs.  The warp copies
// its 32 candidate rows into shared memory with cp.async, every row as contiguous 512-byte
// requests (a gather of whole 3 KB rows runs at HBM speed; per-lane 16-byte reads of 32
// different rows do not), then lane l folds pair l strictly left to right from shared memory.
// Rows sit `stride` floats apart with stride/4 odd, so the per-lane LDS.128 are conflict-free.
// Slabs of up to RS_SLAB columns bound the shared memory for large dims.
constexpr int RS_SLAB = 768;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// OWNED = true (owner-computes rescoring, gvdb_rescore_keys_device): buf holds nq x R keys
// hamming << 40 | GLOBAL row; a pair is scored iff its row lies in [win_lo, win_hi) (local row
// numbers, local = global - row_base), everything else gets 0.0f; only out_score is written and
// the query norms are folded here (qnorm == nullptr).
template <bool OWNED>
__global__ void __launch_bounds__(32)
rescore_slab_kernel(const float* __restrict__ rows, const float* __restrict__ norms, uint64_t row_base,
                    int dim, int stride, int q_slots, const float* __restrict__ queries,
                    const float* __restrict__ qnorm, const uint64_t* __restrict__ buf, uint32_t cap,
                    const uint32_t* __restrict__ cnt, uint32_t R, uint32_t nq, uint32_t* __restrict__ out_ham,
                    uint64_t* __restrict__ out_ids, float* __restrict__ out_score,
                    uint64_t win_lo, uint64_t win_hi,
                    const float* const* __restrict__ peer_rows = nullptr, uint64_t rows_per_owner = 0) {
    // peer_rows != nullptr: row r lives in owner r / rows_per_owner's buffer (this GPU's or a
    // peer's, mapped over NVLink) at offset (r % rows_per_owner) * dim — gvdb_attach_peer_rows_*.
    extern __shared__ __align__(16) float srow[];            // 32 candidate rows, then q_slots query rows
    float* sq = srow + (size_t)32 * stride;
    const int lane = threadIdx.x;
    const uint64_t p0 = (uint64_t)blockIdx.x * 32u;
    const uint64_t p = p0 + lane;
    const uint32_t q = (uint32_t)(p / R), r = (uint32_t)(p % R);
    const uint32_t q_first = (uint32_t)(p0 / R);             // the warp's pairs cover queries q_first ...
    const bool slot = q < nq;
    bool valid;
    uint64_t key;
    uint32_t my_row;
    if (OWNED) {
        key = slot ? buf[(size_t)q * R + r] : UINT64_MAX;
        const uint64_t local = (key & ((1ull << 40) - 1)) - row_base;
        valid = key != UINT64_MAX && local >= win_lo && local < win_hi;
        my_row = valid ? (uint32_t)local : 0xffffffffu;
    } else {
        valid = slot && r < cnt[(size_t)q * CNT_STRIDE];
        key = valid ? buf[(size_t)q * cap + r] : UINT64_MAX;
        my_row = (uint32_t)key;
    }
    float dot = 0.0f, qq2 = 0.0f;
    if (OWNED && !__any_sync(0xffffffffu, valid)) {          // none of these rows lives here
        if (slot) out_score[p] = 0.0f;
        return;
    }
    // Staging: ONE TMA bulk copy per row (lane l fetches its own candidate row, the first q_slots
    // lanes also fetch a query row), completion on an mbarrier.  Whole-row requests matter when the
    // row lives in a peer GPU's HBM: NVLink carries them as large packets.
    __shared__ __align__(8) uint64_t s_bar;
    const uint32_t bar = smem_u32(&s_bar);
    if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
    uint32_t phase = 0;
    const bool have_row = my_row != 0xffffffffu;
    const bool have_q = lane < q_slots && q_first + lane < nq;
    const uint32_t n_copies = __popc(__ballot_sync(0xffffffffu, have_row)) + __popc(__ballot_sync(0xffffffffu, have_q));
    for (int c0 = 0; c0 < dim; c0 += RS_SLAB) {
        const int cols = min(RS_SLAB, dim - c0);
        const int nv = cols >> 2;                            // float4 per row in this slab
        const uint32_t row_bytes = (uint32_t)cols * 4u;
        if (c0) {                                            // everyone is done reading the previous slab,
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // order those reads before the TMA writes
            __syncwarp();
        }
        if (lane == 0) mbar_expect_tx(bar, n_copies * row_bytes);
        __syncwarp();
        if (have_row) {
            const float* src = peer_rows ? peer_rows[my_row / rows_per_owner] + (size_t)(my_row % rows_per_owner) * dim + c0
                                         : rows + (size_t)my_row * dim + c0;
            tma_bulk_g2s(smem_u32(srow + (size_t)lane * stride), src, row_bytes, bar);
        }
        if (have_q)
            tma_bulk_g2s(smem_u32(sq + (size_t)lane * stride), queries + (size_t)(q_first + lane) * dim + c0, row_bytes, bar);
        while (!mbar_try_wait(bar, phase)) {}
        phase ^= 1u;
        if (valid) {
            const float4* mine = reinterpret_cast<const float4*>(srow + (size_t)lane * stride);
            const float4* myq = reinterpret_cast<const float4*>(sq + (size_t)(q - q_first) * stride);
#pragma unroll 4
            for (int v = 0; v < nv; ++v) {
                const float4 a = myq[v], b = mine[v];
                dot = __fadd_rn(dot, __fmul_rn(a.x, b.x));
                dot = __fadd_rn(dot, __fmul_rn(a.y, b.y));
                dot = __fadd_rn(dot, __fmul_rn(a.z, b.z));
                dot = __fadd_rn(dot, __fmul_rn(a.w, b.w));
                if (OWNED) {                                 // ||q||^2, the same sequential fold
                    qq2 = __fadd_rn(qq2, __fmul_rn(a.x, a.x));
                    qq2 = __fadd_rn(qq2, __fmul_rn(a.y, a.y));
                    qq2 = __fadd_rn(qq2, __fmul_rn(a.z, a.z));
                    qq2 = __fadd_rn(qq2, __fmul_rn(a.w, a.w));
                }
            }
        }
    }
    if (slot) {
        if (OWNED) {
            float cosv = 0.0f;
            if (valid) {
                const float na = __fsqrt_rn(qq2), nb = norms[my_row];
                cosv = (na == 0.0f || nb == 0.0f) ? 0.0f : __fdiv_rn(dot, __fmul_rn(na, nb));
            }
            out_score[p] = cosv;
        } else {
            float cosv = -INFINITY;
            if (valid) {
                const float na = qnorm[q], nb = norms[my_row];
                cosv = (na == 0.0f || nb == 0.0f) ? 0.0f : __fdiv_rn(dot, __fmul_rn(na, nb));
            }
            out_ham[p] = valid ? (uint32_t)(key >> 32) : 0xffffffffu;
            out_ids[p] = valid ? row_base + my_row : UINT64_MAX;
            out_score[p] = cosv;
        }
    }
}

// ---------------------------------------------------------------------------------------
// rescore_topk_kernel (dim % 4 == 0, R <= RT_MAX_R): ONE CTA per query, one warp per 32 of its R
// candidates.  The query row is staged once per CTA; every warp streams its 32 candidate rows through
// a two-deep ring of RT_SLAB-column slabs (one TMA bulk copy per row and slab, the next slab in flight
// while lane l folds pair l strictly left to right: separately rounded multiply and add, the
// reference's iterator sum, /root/reference/src/quantization.rs:206-216).  The cosines then go through
// the block-wide ordering of topk_kernel in the same CTA (key = descending cosine image << 32 |
// stage-1 position: equal cosines keep stage-1 order, the reference's second stable sort, :190) and
// the first k leave as the answer.  rec_* (optional): the shard records in stage-1 order.
// Replaces rescore_slab_kernel + topk_kernel on the single-index path: twice the rows in flight per SM
// and no record round trip through HBM.
constexpr int RT_SLAB = 64;            // columns per slab: 32 rows x 256 B per warp and slab (ten warps per SM)
constexpr int RT_MAX_R = 256;          // up to 8 warps per CTA
constexpr int RT_STRIDE = RT_SLAB + 4; // floats between staged rows: stride / 4 odd -> conflict-free LDS.128

__global__ void __launch_bounds__(RT_MAX_R)
rescore_topk_kernel(const float* __restrict__ rows, const float* __restrict__ norms, uint64_t row_base, int dim,
                    const float* __restrict__ queries, const float* __restrict__ qnorm,
                    const uint64_t* __restrict__ buf, uint32_t cap, const uint32_t* __restrict__ cnt, uint32_t R,
                    uint32_t n_eff, uint32_t k, uint64_t* __restrict__ ids_out, float* __restrict__ scores_out,
                    uint32_t* __restrict__ rec_ham, uint64_t* __restrict__ rec_ids, float* __restrict__ rec_score,
                    const float* const* __restrict__ peer_rows, uint64_t rows_per_owner,
                    uint32_t slice_q /* > 0: packed shard records grouped by slices of slice_q queries, rec_ids = base */) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float rt_smem[];         // query row | per warp: 2 slabs of 32 x RT_STRIDE | sort keys
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x;
    float* s_q = rt_smem;
    // every warp stages two slabs of ITS rows: 32, except the last warp of the CTA, which holds R - 32 (nwarps - 1)
    // (R = 40: 8 rows — the CTA then needs 25 KB instead of 38 KB and all 1024 queries of a tile are resident at once)
    const int rw = min(32, (int)R - warp * 32);
    float* s_rows = rt_smem + dim + (size_t)warp * 2 * 32 * RT_STRIDE;
    uint64_t* skeys = reinterpret_cast<uint64_t*>(rt_smem + dim + (size_t)R * 2 * RT_STRIDE);
    __shared__ __align__(8) uint64_t s_bar[1 + 2 * (RT_MAX_R / 32)];
    const uint32_t bar_q = smem_u32(&s_bar[0]);
    const uint32_t bar0 = smem_u32(&s_bar[1 + 2 * warp]);
    if (threadIdx.x == 0) { mbar_init(bar_q, 1); }
    if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
    if (threadIdx.x == 0 || lane == 0) fence_mbar_init();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar_q, (uint32_t)dim * 4u);
        tma_bulk_g2s(smem_u32(s_q), queries + (size_t)q * dim, (uint32_t)dim * 4u, bar_q);
    }
    const uint32_t r = threadIdx.x;                          // stage-1 position of this thread's candidate
    const bool valid = r < R && r < cnt[(size_t)q * CNT_STRIDE];
    const uint64_t key = valid ? buf[(size_t)q * cap + r] : UINT64_MAX;
    const uint32_t my_row = (uint32_t)key;
    const float* src = nullptr;
    if (valid) src = peer_rows ? peer_rows[my_row / rows_per_owner] + (size_t)(my_row % rows_per_owner) * dim
                               : rows + (size_t)my_row * dim;
    const uint32_t n_valid = __popc(__ballot_sync(0xffffffffu, valid));
    const int n_slabs = (dim + RT_SLAB - 1) / RT_SLAB;
    // One TMA bulk copy per row and slab (lane l fetches its own candidate's piece).  Measured on the
    // 1M x 768, R = 40 batch: 52 us with these copies, 62 us with the same pieces staged by cp.async
    // (one 512-byte LDGSTS per row and slab): the gather is bound by the latency of ~6 warps per SM,
    // not by the copy engine.
    auto issue = [&](int sl) {                               // slab sl -> ring slot sl & 1
        const int c0 = sl * RT_SLAB;
        const uint32_t bytes = (uint32_t)min(RT_SLAB, dim - c0) * 4u;
        const uint32_t bar = bar0 + 8u * (sl & 1);
        if (lane == 0) mbar_expect_tx(bar, n_valid * bytes);
        __syncwarp();
        if (valid) tma_bulk_g2s(smem_u32(s_rows + ((size_t)(sl & 1) * rw + lane) * RT_STRIDE), src + c0, bytes, bar);
    };
    float dot = 0.0f;
    if (n_valid) {
        issue(0);
        while (!mbar_try_wait(bar_q, 0)) {}
        for (int sl = 0; sl < n_slabs; ++sl) {
            if (sl + 1 < n_slabs) {
                // the slot of slab sl + 1 was read by slab sl - 1: order those reads before the TMA writes
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                issue(sl + 1);
            }
            while (!mbar_try_wait(bar0 + 8u * (sl & 1), (uint32_t)(sl >> 1) & 1u)) {}
            if (valid) {
                const int nv = min(RT_SLAB, dim - sl * RT_SLAB) >> 2;
                const float4* mine = reinterpret_cast<const float4*>(s_rows + ((size_t)(sl & 1) * rw + lane) * RT_STRIDE);
                const float4* myq = reinterpret_cast<const float4*>(s_q + sl * RT_SLAB);
#pragma unroll 8
                for (int v = 0; v < nv; ++v) {
                    const float4 a = myq[v], b = mine[v];
                    dot = __fadd_rn(dot, __fmul_rn(a.x, b.x));
                    dot = __fadd_rn(dot, __fmul_rn(a.y, b.y));
                    dot = __fadd_rn(dot, __fmul_rn(a.z, b.z));
                    dot = __fadd_rn(dot, __fmul_rn(a.w, b.w));
                }
            }
        }
    }
    float cosv = -INFINITY;
    if (valid) {
        const float na = qnorm[q], nb = norms[my_row];
        cosv = (na == 0.0f || nb == 0.0f) ? 0.0f : __fdiv_rn(dot, __fmul_rn(na, nb));
    }
    if (rec_ids && r < R) {
        uint64_t* o_ids = rec_ids; uint32_t* o_ham = rec_ham; float* o_sc = rec_score;
        size_t p = (size_t)q * R + r;
        if (slice_q) {
            // slice s = q / slice_q is one packed buffer [ids u64 | ham u32 | score f32] x slice_q x R
            // (the layout gvdb_merge_shards_device consumes after the all-to-all)
            const size_t nr = (size_t)slice_q * R;
            uint8_t* base = reinterpret_cast<uint8_t*>(rec_ids) + (size_t)(q / slice_q) * nr * 16;
            o_ids = reinterpret_cast<uint64_t*>(base);
            o_ham = reinterpret_cast<uint32_t*>(base + nr * 8);
            o_sc = reinterpret_cast<float*>(base + nr * 12);
            p = (size_t)(q % slice_q) * R + r;
        }
        o_ham[p] = valid ? (uint32_t)(key >> 32) : 0xffffffffu;
        o_ids[p] = valid ? row_base + my_row : UINT64_MAX;
        o_sc[p] = cosv;
    }
    if (k == 0) return;
    // order: descending cosine image, then stage-1 position
    for (uint32_t i = threadIdx.x; i < n_eff; i += blockDim.x) skeys[i] = UINT64_MAX;
    __syncthreads();
    if (valid) skeys[r] = ((uint64_t)(~f32_asc_key(cosv)) << 32) | r;
    // the cosine and row of position r stay in this thread's registers: the winners are fetched by position
    __shared__ float s_cos[RT_MAX_R];
    __shared__ uint32_t s_rowid[RT_MAX_R];
    s_cos[threadIdx.x] = cosv;
    s_rowid[threadIdx.x] = my_row;
    __syncthreads();
    bitonic_sort_smem(skeys, n_eff);
    for (uint32_t t = threadIdx.x; t < k; t += blockDim.x) {
        const uint64_t kk = t < n_eff ? skeys[t] : UINT64_MAX;
        if (kk == UINT64_MAX) {
            ids_out[(size_t)q * k + t] = UINT64_MAX;
            scores_out[(size_t)q * k + t] = -INFINITY;
        } else {
            const uint32_t pos = (uint32_t)kk;
            ids_out[(size_t)q * k + t] = row_base + s_rowid[pos];
            scores_out[(size_t)q * k + t] = s_cos[pos];
        }
    }
}

// rescore_owned_list_kernel: owner-computes rescoring over a COMPACTED list of pair indices
// (the pairs whose row this GPU holds: 1/G of them, in ascending order).  One warp per 32 list
// entries; lane l TMA-copies its candidate row and its query row (entries of one warp may belong
// to many queries), folds ||q||^2 and the dot product left to right, writes out_score[pair] (or, for
// the peer exchange, the pair's slot in its requester's mailbox).
// done_counter / flag fields (peer exchange): the last block to finish publishes flag[kind][rank] := epoch in
// every peer's mailbox — the cosines this kernel stored into the peers' mailboxes are complete.
struct OwnedSignal {
    uint64_t flags_off;
    uint32_t kind, world, rank, epoch, flag_stride;
    uint32_t* done;          // nullptr: no signal
};
__device__ __forceinline__ void owned_block_done(const OwnedSignal& sg, uint8_t* const* peers, uint32_t total_blocks) {
    if (!sg.done) return;
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(sg.done, 1u);
        if (prev + 1 == total_blocks) {
            __threadfence_system();
            for (uint32_t w = 0; w < sg.world; ++w) {
                uint32_t* f = reinterpret_cast<uint32_t*>(peers[w] + sg.flags_off) + (size_t)(sg.kind * sg.world + sg.rank) * sg.flag_stride;
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(sg.epoch) : "memory");
            }
            *sg.done = 0u;
        }
    }
}

__global__ void __launch_bounds__(32)
rescore_owned_list_kernel(const float* __restrict__ rows, const float* __restrict__ norms, uint64_t row_base,
                          int dim, int stride, const float* __restrict__ queries, const uint64_t* __restrict__ keys,
                          const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_count, uint32_t R,
                          float* __restrict__ out_score, uint8_t* const* __restrict__ peers, uint64_t peer_off,
                          uint32_t pairs_per_peer, OwnedSignal sg) {
    extern __shared__ __align__(16) float srow[];            // 32 candidate rows, then 32 query rows
    float* sq = srow + (size_t)32 * stride;
    const int lane = threadIdx.x;
    const uint32_t n_list = *list_count;
    const uint32_t i = blockIdx.x * 32u + lane;
    if (blockIdx.x * 32u >= n_list) { owned_block_done(sg, peers, gridDim.x); return; }   // warp-uniform
    const bool valid = i < n_list;
    const uint32_t p = valid ? list[i] : 0u;
    const uint32_t q = p / R;
    const uint64_t key = valid ? keys[p] : 0ull;
    const uint32_t my_row = (uint32_t)((key & ((1ull << 40) - 1)) - row_base);
    __shared__ __align__(8) uint64_t s_bar;
    const uint32_t bar = smem_u32(&s_bar);
    if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
    uint32_t phase = 0;
    // lanes scoring pairs of the same query share one staged copy of it (the list is ascending, so a
    // query's pairs sit in neighbouring lanes): the first such lane copies, the others read its slot
    const uint32_t same_q = __match_any_sync(0xffffffffu, valid ? q : 0xFFFFFFFFu - (uint32_t)lane);
    const int q_lane = __ffs(same_q) - 1;
    const bool q_leader = valid && lane == q_lane;
    const uint32_t n_copies = __popc(__ballot_sync(0xffffffffu, valid)) + __popc(__ballot_sync(0xffffffffu, q_leader));
    float dot = 0.0f, qq2 = 0.0f;
    for (int c0 = 0; c0 < dim; c0 += RS_SLAB) {
        const int cols = min(RS_SLAB, dim - c0);
        const int nv = cols >> 2;
        const uint32_t row_bytes = (uint32_t)cols * 4u;
        if (c0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
        }
        if (lane == 0) mbar_expect_tx(bar, n_copies * row_bytes);
        __syncwarp();
        if (valid) {
            tma_bulk_g2s(smem_u32(srow + (size_t)lane * stride), rows + (size_t)my_row * dim + c0, row_bytes, bar);
            if (q_leader) tma_bulk_g2s(smem_u32(sq + (size_t)lane * stride), queries + (size_t)q * dim + c0, row_bytes, bar);
        }
        while (!mbar_try_wait(bar, phase)) {}
        phase ^= 1u;
        if (valid) {
            const float4* mine = reinterpret_cast<const float4*>(srow + (size_t)lane * stride);
            const float4* myq = reinterpret_cast<const float4*>(sq + (size_t)q_lane * stride);
#pragma unroll 4
            for (int v = 0; v < nv; ++v) {
                const float4 a = myq[v], b = mine[v];
                dot = __fadd_rn(dot, __fmul_rn(a.x, b.x));
                dot = __fadd_rn(dot, __fmul_rn(a.y, b.y));
                dot = __fadd_rn(dot, __fmul_rn(a.z, b.z));
                dot = __fadd_rn(dot, __fmul_rn(a.w, b.w));
                qq2 = __fadd_rn(qq2, __fmul_rn(a.x, a.x));
                qq2 = __fadd_rn(qq2, __fmul_rn(a.y, a.y));
                qq2 = __fadd_rn(qq2, __fmul_rn(a.z, a.z));
                qq2 = __fadd_rn(qq2, __fmul_rn(a.w, a.w));
            }
        }
    }
    if (valid) {
        const float na = __fsqrt_rn(qq2), nb = norms[my_row];
        const float cosv = (na == 0.0f || nb == 0.0f) ? 0.0f : __fdiv_rn(dot, __fmul_rn(na, nb));
        // peer exchange: pair p belongs to rank p / pairs_per_peer; its cosine is stored straight into that
        // rank's mailbox (a posted store over NVLink), at peer_off + (p % pairs_per_peer) floats
        if (peers) reinterpret_cast<float*>(peers[p / pairs_per_peer] + peer_off)[p % pairs_per_peer] = cosv;
        else out_score[p] = cosv;
    }
    owned_block_done(sg, peers, gridDim.x);
}

// rescore_owned_ring_kernel: rescore_owned_list_kernel with the staging of rescore_topk_kernel — one warp per
// 32 list entries, rows AND query rows streamed through a two-deep ring of RO_SLAB-column slabs (one TMA
// bulk copy per row and slab), so that ten warps fit an SM instead of one.  Entries of a warp that share a
// query share one staged copy of it (the list is nearly sorted by query); a warp whose entries span more
// than RO_QSLOTS queries goes through its entries in rounds of RO_QSLOTS queries.
constexpr int RO_SLAB = 64;
constexpr int RO_STRIDE = RO_SLAB + 4;
constexpr int RO_QSLOTS = 8;
constexpr int RO_WARPS = 2;
constexpr size_t RO_SMEM = (size_t)RO_WARPS * 2 * (32 + RO_QSLOTS) * RO_STRIDE * sizeof(float);

__global__ void __launch_bounds__(32 * RO_WARPS)
rescore_owned_ring_kernel(const float* __restrict__ rows, const float* __restrict__ norms, uint64_t row_base, int dim,
                          const float* __restrict__ queries, const uint64_t* __restrict__ keys,
                          const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_count, uint32_t R,
                          float* __restrict__ out_score, uint8_t* const* __restrict__ peers, uint64_t peer_off,
                          uint32_t pairs_per_peer) {
    extern __shared__ __align__(16) float ro_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* s_ring = ro_smem + (size_t)warp * 2 * (32 + RO_QSLOTS) * RO_STRIDE;     // slot s: 32 rows then RO_QSLOTS queries
    __shared__ __align__(8) uint64_t s_bar[2 * RO_WARPS];
    const uint32_t bar0 = smem_u32(&s_bar[2 * warp]);
    if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); fence_mbar_init(); }
    __syncwarp();
    const uint32_t n_list = *list_count;
    const uint32_t i = (blockIdx.x * RO_WARPS + warp) * 32u + lane;
    if ((blockIdx.x * RO_WARPS + warp) * 32u >= n_list) return;          // warp-uniform
    const bool valid = i < n_list;
    const uint32_t p = valid ? list[i] : 0u;
    const uint32_t q = p / R;
    const uint64_t key = valid ? keys[p] : 0ull;
    const uint32_t my_row = (uint32_t)((key & ((1ull << 40) - 1)) - row_base);
    // slot of my query among the warp's distinct queries (leaders in lane order)
    const uint32_t same_q = __match_any_sync(0xffffffffu, valid ? q : 0xFFFFFFFFu - (uint32_t)lane);
    const int q_lane = __ffs(same_q) - 1;
    const bool q_leader = valid && lane == q_lane;
    const uint32_t leaders = __ballot_sync(0xffffffffu, q_leader);
    const int my_slot = __popc(leaders & ((1u << q_lane) - 1u));
    const int n_rounds = (__popc(leaders) + RO_QSLOTS - 1) / RO_QSLOTS;
    const int n_slabs = (dim + RO_SLAB - 1) / RO_SLAB;
    uint32_t uses = 0;                                       // ring slot uses so far (barrier phases)
    float cosv = 0.0f;
    for (int round = 0; round < n_rounds; ++round) {
        const bool mine = valid && my_slot / RO_QSLOTS == round;
        const bool lead = q_leader && my_slot / RO_QSLOTS == round;
        const int qs = my_slot % RO_QSLOTS;
        const uint32_t n_copies = __popc(__ballot_sync(0xffffffffu, mine)) + __popc(__ballot_sync(0xffffffffu, lead));
        auto issue = [&](int sl, uint32_t u) {
            const int c0 = sl * RO_SLAB;
            const uint32_t bytes = (uint32_t)min(RO_SLAB, dim - c0) * 4u;
            const uint32_t bar = bar0 + 8u * (u & 1u);
            float* base = s_ring + (size_t)(u & 1u) * (32 + RO_QSLOTS) * RO_STRIDE;
            if (lane == 0) mbar_expect_tx(bar, n_copies * bytes);
            __syncwarp();
            if (mine) tma_bulk_g2s(smem_u32(base + (size_t)lane * RO_STRIDE), rows + (size_t)my_row * dim + c0, bytes, bar);
            if (lead) tma_bulk_g2s(smem_u32(base + (size_t)(32 + qs) * RO_STRIDE), queries + (size_t)q * dim + c0, bytes, bar);
        };
        float dot = 0.0f, qq2 = 0.0f;
        if (n_copies) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            issue(0, uses);
            for (int sl = 0; sl < n_slabs; ++sl) {
                const uint32_t u = uses + (uint32_t)sl;
                if (sl + 1 < n_slabs) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    issue(sl + 1, u + 1);
                }
                while (!mbar_try_wait(bar0 + 8u * (u & 1u), (u >> 1) & 1u)) {}
                if (mine) {
                    const float* base = s_ring + (size_t)(u & 1u) * (32 + RO_QSLOTS) * RO_STRIDE;
                    const int nv = min(RO_SLAB, dim - sl * RO_SLAB) >> 2;
                    const float4* r4 = reinterpret_cast<const float4*>(base + (size_t)lane * RO_STRIDE);
                    const float4* q4 = reinterpret_cast<const float4*>(base + (size_t)(32 + qs) * RO_STRIDE);
#pragma unroll 8
                    for (int v = 0; v < nv; ++v) {
                        const float4 a = q4[v], b = r4[v];
                        dot = __fadd_rn(dot, __fmul_rn(a.x, b.x));
                        dot = __fadd_rn(dot, __fmul_rn(a.y, b.y));
                        dot = __fadd_rn(dot, __fmul_rn(a.z, b.z));
                        dot = __fadd_rn(dot, __fmul_rn(a.w, b.w));
                        qq2 = __fadd_rn(qq2, __fmul_rn(a.x, a.x));
                        qq2 = __fadd_rn(qq2, __fmul_rn(a.y, a.y));
                        qq2 = __fadd_rn(qq2, __fmul_rn(a.z, a.z));
                        qq2 = __fadd_rn(qq2, __fmul_rn(a.w, a.w));
                    }
                }
            }
            uses += (uint32_t)n_slabs;
        }
        if (mine) {
            const float na = __fsqrt_rn(qq2), nb = norms[my_row];
            cosv = (na == 0.0f || nb == 0.0f) ? 0.0f : __fdiv_rn(dot, __fmul_rn(na, nb));
        }
    }
    if (valid) {
        if (peers) reinterpret_cast<float*>(peers[p / pairs_per_peer] + peer_off)[p % pairs_per_peer] = cosv;
        else out_score[p] = cosv;
    }
}

// filtered searches: the bitmap the scans consult = tombstones AND the call's allow-list
__global__ void and_bitmap_kernel(const uint32_t* __restrict__ live, const uint32_t* __restrict__ allow, size_t n_words,
                                  uint32_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_words) out[i] = live[i] & allow[i];
}

// stage-1 result as keys hamming << 40 | global row (gvdb_stage1_device)
__global__ void emit_keys_kernel(const uint64_t* __restrict__ buf, uint32_t cap, const uint32_t* __restrict__ cnt,
                                 uint32_t R, uint32_t nq, uint64_t row_base, uint64_t* __restrict__ keys_out) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t q = (uint32_t)(p / R), r = (uint32_t)(p % R);
    if (q >= nq) return;
    uint64_t out = UINT64_MAX;
    if (r < cnt[(size_t)q * CNT_STRIDE]) {
        const uint64_t key = buf[(size_t)q * cap + r];
        out = ((key >> 32) << 40) | (row_base + (uint32_t)key);
    }
    keys_out[p] = out;
}

// each key takes the score its row's owner computed (gvdb_finish_owned_device)
__global__ void gather_owner_scores_kernel(const uint64_t* __restrict__ keys, const float* __restrict__ by_owner,
                                           uint32_t n_owners, uint64_t rows_per_owner, uint64_t n_pairs,
                                           uint64_t* __restrict__ rec_ids, float* __restrict__ rec_score) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const uint64_t key = keys[p];
    if (key == UINT64_MAX) { rec_ids[p] = UINT64_MAX; rec_score[p] = -INFINITY; return; }
    const uint64_t row = key & ((1ull << 40) - 1);
    const uint64_t owner = min((uint64_t)n_owners - 1, row / rows_per_owner);
    rec_ids[p] = row;
    rec_score[p] = by_owner[owner * n_pairs + p];
}

// ---------------------------------------------------------------------------------------
// topk_kernel: one CTA per query over its R records (already in (hamming, row) order).
// Sort key = (descending image of the cosine) << 32 | position, so equal cosines keep
// stage-1 order — exactly what the reference's second stable sort does.
__global__ void topk_kernel(const uint64_t* __restrict__ rec_ids, const float* __restrict__ rec_score,
                            uint32_t R, uint32_t n_eff, uint32_t k, uint64_t* __restrict__ ids_out,
                            float* __restrict__ scores_out) {
    extern __shared__ __align__(16) uint64_t skeys[];
    const uint32_t q = blockIdx.x;
    const uint64_t* ids = rec_ids + (size_t)q * R;
    const float* sc = rec_score + (size_t)q * R;
    for (uint32_t i = threadIdx.x; i < n_eff; i += blockDim.x) {
        uint64_t key = UINT64_MAX;
        if (i < R && ids[i] != UINT64_MAX) key = ((uint64_t)(~f32_asc_key(sc[i])) << 32) | i;
        skeys[i] = key;
    }
    __syncthreads();
    bitonic_sort_smem(skeys, n_eff);
    for (uint32_t t = threadIdx.x; t < k; t += blockDim.x) {
        uint64_t key = t < n_eff ? skeys[t] : UINT64_MAX;
        if (key == UINT64_MAX) {
            ids_out[(size_t)q * k + t] = UINT64_MAX;
            scores_out[(size_t)q * k + t] = -INFINITY;
        } else {
            uint32_t pos = (uint32_t)key;
            ids_out[(size_t)q * k + t] = ids[pos];
            scores_out[(size_t)q * k + t] = sc[pos];
        }
    }
}

// topk_owned_kernel: topk_kernel for the owner-computes layouts with gather_owner_scores_kernel folded in: the
// score of candidate i of query q is the one its row's owner computed, by_owner[owner][q * R + i].
__global__ void topk_owned_kernel(const uint64_t* __restrict__ keys, const float* __restrict__ by_owner,
                                  uint32_t n_owners, uint64_t rows_per_owner, uint32_t nq, uint32_t R, uint32_t n_eff,
                                  uint32_t k, uint64_t* __restrict__ ids_out, float* __restrict__ scores_out) {
    extern __shared__ __align__(16) uint64_t skeys[];
    const uint32_t q = blockIdx.x;
    const size_t n_pairs = (size_t)nq * R;
    auto score_of = [&](uint32_t i, uint64_t& row) {
        const uint64_t key = keys[(size_t)q * R + i];
        row = key & ((1ull << 40) - 1);
        const uint64_t owner = min((uint64_t)n_owners - 1, row / rows_per_owner);
        return by_owner[owner * n_pairs + (size_t)q * R + i];
    };
    for (uint32_t i = threadIdx.x; i < n_eff; i += blockDim.x) {
        uint64_t key = UINT64_MAX;
        if (i < R && keys[(size_t)q * R + i] != UINT64_MAX) {
            uint64_t row;
            key = ((uint64_t)(~f32_asc_key(score_of(i, row))) << 32) | i;
        }
        skeys[i] = key;
    }
    __syncthreads();
    bitonic_sort_smem(skeys, n_eff);
    for (uint32_t t = threadIdx.x; t < k; t += blockDim.x) {
        const uint64_t key = t < n_eff ? skeys[t] : UINT64_MAX;
        if (key == UINT64_MAX) {
            ids_out[(size_t)q * k + t] = UINT64_MAX;
            scores_out[(size_t)q * k + t] = -INFINITY;
        } else {
            uint64_t row;
            const float sc = score_of((uint32_t)key, row);
            ids_out[(size_t)q * k + t] = row;
            scores_out[(size_t)q * k + t] = sc;
        }
    }
}

// ---------------------------------------------------------------------------------------
// merge_select_kernel: one CTA per query over the gathered shard records
// (n_shards packed buffers back to back, each [ids u64 | ham u32 | score f32] x nq x R).
// Keeps the global top R by (hamming, global row) — the stage-1 cut the single-index search
// would have made — and writes them, in that order, as one merged record list per query.
// Sort key = hamming << 40 | global row (rows < 2^40).  (hamming, row) is unique across
// shards and every shard's list is sorted by it, so the score of a kept key is recovered by
// a binary search in the shard lists instead of carrying a payload through the sort.
struct ShardRecords {
    const uint8_t* base;
    uint64_t shard_bytes;
    uint32_t nq, R;
    uint32_t keep;      // records kept per query: R (the stage-1 cut) or n_shards * R (ratio mode: the shards already
                        // applied the global cut, every record stays)
    __device__ __forceinline__ const uint64_t* ids(uint32_t s) const {
        return reinterpret_cast<const uint64_t*>(base + s * shard_bytes);
    }
    __device__ __forceinline__ const uint32_t* ham(uint32_t s) const {
        return reinterpret_cast<const uint32_t*>(base + s * shard_bytes + (uint64_t)nq * R * 8);
    }
    __device__ __forceinline__ const float* score(uint32_t s) const {
        return reinterpret_cast<const float*>(base + s * shard_bytes + (uint64_t)nq * R * 12);
    }
};

__global__ void __launch_bounds__(SORT_THREADS)
merge_select_kernel(uint32_t n_shards, ShardRecords rec, uint64_t* __restrict__ out_ids,
                    float* __restrict__ out_score, uint32_t* __restrict__ out_ham) {
    extern __shared__ __align__(16) uint64_t skeys[];
    const uint32_t q = blockIdx.x;
    const uint32_t R = rec.R;
    const uint32_t total_in = n_shards * R;
    uint32_t have = 0, consumed = 0;
    while (consumed < total_in) {
        const uint32_t take = min((uint32_t)SORT_N - have, total_in - consumed);
        const uint32_t total = have + take;
        const uint32_t n_eff = max(64u, next_pow2(total));
        for (uint32_t i = threadIdx.x; i < n_eff - have; i += blockDim.x) {
            uint64_t key = UINT64_MAX;
            if (i < take) {
                const uint32_t src = consumed + i;
                const uint32_t s = src / R, r = src % R;
                const size_t at = (size_t)q * R + r;
                const uint64_t id = rec.ids(s)[at];
                if (id != UINT64_MAX) key = ((uint64_t)rec.ham(s)[at] << 40) | id;
            }
            skeys[have + i] = key;
        }
        __syncthreads();
        bitonic_sort_smem(skeys, n_eff);
        have = min(rec.keep, total);
        consumed += take;
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < rec.keep; t += blockDim.x) {
        const uint64_t key = t < have ? skeys[t] : UINT64_MAX;
        uint64_t id = UINT64_MAX;
        uint32_t hm = 0xffffffffu;
        float sc = -INFINITY;
        if (key != UINT64_MAX) {
            id = key & ((1ull << 40) - 1);
            hm = (uint32_t)(key >> 40);
            for (uint32_t s = 0; s < n_shards; ++s) {
                const uint64_t* ids = rec.ids(s) + (size_t)q * R;
                const uint32_t* ham = rec.ham(s) + (size_t)q * R;
                uint32_t lo = 0, hi = R;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    const uint64_t mid_id = ids[mid];
                    const uint64_t mk = mid_id == UINT64_MAX ? UINT64_MAX
                                                             : (((uint64_t)ham[mid] << 40) | mid_id);
                    if (mk < key) lo = mid + 1; else hi = mid;
                }
                if (lo < R && ids[lo] == id && ham[lo] == hm) {
                    sc = rec.score(s)[(size_t)q * R + lo];
                    break;
                }
            }
        }
        out_ids[(size_t)q * rec.keep + t] = id;
        out_ham[(size_t)q * rec.keep + t] = hm;
        out_score[(size_t)q * rec.keep + t] = sc;
    }
}

}  // namespace gvdb
