// gvdb_tc.cuh — the batched Hamming scan on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Used when a corpus pass serves a large query tile (>= 64 queries), where the scan really is a
// dense contraction: rows x queries x dim.  The CUDA-core kernel (scan_kernel) is then bound by
// the integer pipes (16 popc + 64 LOP3 per clk per SM).  A 1-bit code needs no more than 4 bits
// per element on the tensor core, so the contraction runs on the FP4 path:
// tcgen05.mma.kind::mxf4 (e2m1 operands, one UE8M0 scale per 32 elements, all scales 1.0, f32
// accumulation).  Measured with tools/mxf4_probe.cu: 76 clk per M128 x N128 x K64 MMA with A in
// TMEM = 13.8k MAC/clk/SM; kind::i8 and kind::f8f6f4 do 8.2k (tools/mma_floor.cu: 64 clk per K32).
//
// Exactness.  Row codes are expanded on chip to A in {0, 1.0} (nibbles 0x0 / 0x2), query codes
// once per batch to B in {-1.0, +1.0} (0xA where the query bit is 1, 0x2 where it is 0).  With
// S = n11 - n10 over the code bits,  sum_k A_k * B_k = -S  and  hamming = popc(q) - S, so
// hamming < tau  <=>  S + (tau - popc(q)) > 0.  All products and partial sums are integers far
// below 2^24: the f32 accumulation is exact.  Pad bits are 0 in A and contribute nothing.
// The per-query bias v = tau - popc(q) is added ON the tensor core by one extra K=64 MMA per
// accumulator block: A = 64 x 1.0 per row with scale factors (16, 1) for its two 32-element
// halves, B = 32 coarse + 31 fine e2m1 digits with 16 * coarse + fine = -v, plus one digit 0.5.
// The accumulator is then D = -(S + v) + 0.5: never zero, and a survivor is simply D < 0, so the
// epilogue is a pure sign test on registers (a per-element threshold load from shared memory
// measured 3x slower).
// One row word expands with a shift + a LOP3 per output word,
//     out[4*w + i] = ((code_word[w] >> i) << 1) & 0x22222222       (nibble j <- code bit 4j + i),
// a fixed permutation of the code bits applied to rows and queries alike (Hamming distance is
// invariant to it).
//
// Work decomposition.  item = (query slice, row slice).  A query slice is up to tc_qblocks() blocks
// of 128 queries whose expanded codes (+ bias digits) stay RESIDENT in shared memory (3 blocks x
// 52 KB at 768 bits) for the whole item; the item's rows stream through TMEM as the A operand, 128
// at a time.  (Streaming the queries through a shared-memory ring instead needs more L2->SM
// bandwidth than the L2 delivers to 148 SMs at once: that version measured 47 % of the MMA floor.)
//
// Warp roles per CTA (one CTA per SM, persistent over items):
//   warps 0-3  expanders: lane = row.  Load the row's code (coalesced, blocked layout, next group
//              prefetched), expand it and write it into TMEM as the A operand (tcgen05.st).  A is a
//              ring of two slots with their own ready/free barriers; a group goes through it in
//              phases, so rewriting one slot overlaps the MMAs that still read the other.
//   warps 4-7  epilogue: read the f32 accumulators back (tcgen05.ld), sign-test them and append
//              survivors to warp-private record lists (MODE 0), or write every distance (MODE 1).
//   warp 8     warp-uniform control flow, one elected lane: TMA bulk loads of the query slice, then
//              tcgen05.mma issue (M=128, N=128, K=64; A from TMEM, B from shared memory, D in TMEM,
//              one accumulator buffer per resident block).
//   mbarriers  b_full/b_free, a_ready/a_free per A slot, acc_full/acc_empty per accumulator
//              buffer; tcgen05.commit signals MMA completion.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gvdb_kernels.cuh"

namespace gvdb {

constexpr int TC_ROWS = 128;        // rows per group (UMMA M)
constexpr int TC_NQ = 128;          // queries per accumulator block (UMMA N)
constexpr int TC_KSTAGE = 128;      // K bytes per 16 KB query sub-block = 256 e2m1 elements = two 16-byte code chunks
constexpr int TC_STAGE_BYTES = TC_NQ * TC_KSTAGE;   // 16 KB
constexpr int TC_BIAS_BYTES = TC_NQ * 32;           // 4 KB: one K=64 slice of per-query bias digits (e2m1)
constexpr int TC_MAX_QBLOCKS = 3;                   // accumulator buffers that fit TMEM next to A
__host__ __device__ constexpr int tc_subblocks(int nchunk) { return (nchunk + 1) / 2; }
__host__ __device__ constexpr size_t tc_qblock_bytes(int nchunk) {
    return (size_t)tc_subblocks(nchunk) * TC_STAGE_BYTES + TC_BIAS_BYTES;
}
// The A operand lives in a ring of two TMEM slots of tc_slot_chunks() code chunks (16 columns
// each); a row group is expanded and consumed in phases, phase ph using slot ph & 1.  Up to 1536
// bits the ring holds the whole group (two phases); longer codes go through it in 8 phases.
__host__ __device__ constexpr int tc_slot_chunks(int nchunk) {
    return nchunk <= 12 ? (nchunk + 1) / 2 : (nchunk % 3 == 0 ? 3 : 2);
}
__host__ __device__ constexpr int tc_phases(int nchunk) { return (nchunk + tc_slot_chunks(nchunk) - 1) / tc_slot_chunks(nchunk); }
// Resident query blocks per item (each has its own TMEM accumulator buffer): what fits 216 KB of
// shared memory and the TMEM columns left next to the A ring, at most three; a single block when
// the ring cannot hold the whole row group (every block would need all the phases again).
__host__ __device__ constexpr int tc_qblocks(int nchunk) {
    int q = (int)((216 * 1024) / tc_qblock_bytes(nchunk));
    const int by_tmem = (512 - 16 - 2 * tc_slot_chunks(nchunk) * 16) / TC_NQ;
    if (by_tmem < q) q = by_tmem;
    if (TC_MAX_QBLOCKS < q) q = TC_MAX_QBLOCKS;
    if (tc_phases(nchunk) > 2) q = q < 1 ? q : 1;
    return q;
}
__host__ __device__ constexpr bool tc_supported_chunks(int nchunk) {
    return nchunk == 1 || nchunk == 2 || nchunk == 3 || nchunk == 4 || nchunk == 6 || nchunk == 8 || nchunk == 12 ||
           nchunk == 16 || nchunk == 24;
}
constexpr int TC_EPI_WARPS = 8;      // two sets of four epilogue warps, alternating accumulator blocks
constexpr int TC_THREADS = 32 * (4 + TC_EPI_WARPS + 1);   // 4 expander warps + epilogue warps + loader/MMA-issuer warp
constexpr uint32_t TC_TMEM_COLS = 512;

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// One lane of a converged warp; the compiler keeps the surrounding control flow warp-uniform, so
// descriptors and TMEM addresses stay in uniform registers (an `if (lane == 0)` around the MMA
// loop costs a register-to-uniform waterfall per MMA: ~72 clk per issue, above the 64 clk floor).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc], e2m1 x e2m1 -> f32, one UE8M0 scale per 32 elements of K
// (sfa / sfb: TMEM addresses of the scale words; byte 0 scales K[0,32), byte 1 scales K[32,64)).
__device__ __forceinline__ void tc_mma_mxf4_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t sfa_tmem, uint32_t sfb_tmem, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%4], [%5], p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(sfa_tmem), "r"(sfb_tmem), "r"(accumulate) : "memory");
}
// 32 lanes x 8 columns per call (each thread: its own lane, 8 consecutive 32-bit columns)
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}

// UMMA shared-memory descriptor, K-major, no swizzle: 8x16 B core matrices (128 B contiguous);
// LBO = byte stride between core matrices along K, SBO = along N.  (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// UMMA block-scaled instruction descriptor (cute::UMMA::InstrDescriptorBlockScaled): A=B=E2M1,
// K-major both, UE8M0 scales, scale-factor ids 0, dense K=64.
__host__ __device__ constexpr uint32_t tc_idesc_mxf4(int M, int N) {
    return (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(M >> 4) << 24);
}
constexpr uint32_t TC_A_ONE8 = 0x22222222u;   // eight e2m1 1.0
constexpr uint32_t TC_SF_ONE = 0x7F7F7F7Fu;   // UE8M0 1.0 in every byte
constexpr uint32_t TC_SF_BIAS_A = 0x7F7F7F83u;  // byte 0 = 16.0 (K[0,32) of the bias MMA), byte 1 = 1.0
constexpr int TC_BIAS_COARSE = 16;

// e2m1 nibble of a magnitude in {0, 0.5, 1, 1.5, 2, 3, 4, 6} given as 2 * value
__host__ __device__ inline uint32_t tc_e2m1_nibble(int twice, bool negative) {
    const uint32_t mag = twice == 0 ? 0u : twice == 1 ? 1u : twice == 2 ? 2u : twice == 3 ? 3u : twice == 4 ? 4u
                       : twice == 6 ? 5u : twice == 8 ? 6u : 7u /* twice == 12 */;
    return mag | ((negative && twice) ? 8u : 0u);
}
// largest integer in {6, 4, 3, 2, 1} not above mag (mag >= 1)
__host__ __device__ inline int tc_e2m1_int_floor(int mag) { return mag >= 6 ? 6 : mag == 5 ? 4 : mag; }

// ---- query pre-expansion ---------------------------------------------------------------------------
// qpack (code words + tau, as produced by the query prep kernel) -> qexp, e2m1 +-1.0 nibbles in the
// exact byte order the resident blocks need:  sub-block (qb, sb) of 16 KB at
// qb*tc_qblock_bytes + sb*16384 (the last 4 KB of a query block hold the bias digits, written by
// tc_bias_kernel), inside a sub-block
//   offset(n, kk) = (n/8)*1024 + (kk/16)*128 + (n%8)*16 + (kk%16),  n = query in block, kk = K byte
// K element e of a query  <->  code bit 32*w + 4*j + i  with  e = (4*w + i)*8 + j  (see header);
// element e sits in K byte e/2, low nibble for even e.  Query bit 1 -> -1.0 (0xA), 0 -> +1.0 (0x2).
// Queries beyond nq (padding up to a multiple of 128) are all zero.  Also writes popc(q).
__global__ void tc_expand_queries_kernel(const uint32_t* __restrict__ qpack, int qs, int nchunk, uint32_t nq,
                                         uint32_t nq_pad, int8_t* __restrict__ qexp,
                                         uint32_t* __restrict__ qpop) {
    const uint32_t q = blockIdx.x;                 // one CTA per (padded) query
    if (q >= nq_pad) return;
    const uint32_t qb = q / TC_NQ, n = q % TC_NQ;
    const int nwords_out = tc_subblocks(nchunk) * (TC_KSTAGE / 4);   // 32-bit output words per query (incl. pad)
    uint32_t pop = 0;
    for (int o = threadIdx.x; o < nwords_out; o += blockDim.x) {   // output word o = 4*w + i: 8 elements, 4 K bytes
        const int w = o >> 2, i = o & 3;
        uint32_t out = 0;
        if (q < nq && w < nchunk * 4) {
            const uint32_t word = qpack[(size_t)q * qs + w];
            out = TC_A_ONE8 | (((word >> i) & 0x11111111u) << 3);       // 0x2 -> 0xA where the bit is set
            if (i == 0) pop += __popc(word);
        }
        const int kb = o * 4;
        const int sb = kb / TC_KSTAGE, kk = kb % TC_KSTAGE;
        const size_t off = (size_t)qb * tc_qblock_bytes(nchunk) + (size_t)sb * TC_STAGE_BYTES + (n / 8) * 1024 + (kk / 16) * 128 + (n % 8) * 16 + (kk % 16);
        *reinterpret_cast<uint32_t*>(qexp + off) = out;
    }
    // block reduce popcount (blockDim <= 256)
    __shared__ uint32_t red[8];
    for (int o = 16; o > 0; o >>= 1) pop += __shfl_xor_sync(0xffffffffu, pop, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = pop;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int i = 0; i < (int)(blockDim.x + 31) / 32; ++i) s += red[i];
        qpop[q] = q < nq ? s : 0u;
    }
}

// Per-query bias digits for the current thresholds: v = tau - popc(q) (v = K+1 when tau is
// TAU_ALL: everything passes; v = -(K+1) for padding queries: nothing passes; zero_bias: v = 0,
// MODE 1).  The bias MMA must add -v + 0.5: 64 e2m1 digits per query, elements [0,32) scaled by
// 16 (coarse), [32,64) by 1 (fine): 16 * sum(coarse) + sum(fine) = -v, and fine digit 0 is the
// 0.5.  K-major core-matrix order of a 128 x 32 B block:
//   offset(n, kb) = (n/8)*256 + (kb/16)*128 + (n%8)*16 + (kb%16),  element e in byte e/2.
__global__ void tc_bias_kernel(const uint32_t* __restrict__ qpack, int qs, int nchunk,
                               const uint32_t* __restrict__ qpop, uint32_t nq, uint32_t nq_pad,
                               int8_t* __restrict__ qexp, int32_t* __restrict__ qbias, int zero_bias) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq_pad) return;
    const int K = nchunk * 128;
    int v = -(K + 1);
    if (q < nq) {
        const uint32_t tau = qpack[(size_t)q * qs + nchunk * 4];
        v = tau == TAU_ALL ? K + 1 : (int)tau - (int)qpop[q];
        if (v > K + 1) v = K + 1;
        if (v < -(K + 1)) v = -(K + 1);
    }
    if (zero_bias) v = 0;
    qbias[q] = v;
    const uint32_t qb = q / TC_NQ, n = q % TC_NQ;
    uint8_t* blk = reinterpret_cast<uint8_t*>(qexp) + (size_t)qb * tc_qblock_bytes(nchunk) + (size_t)tc_subblocks(nchunk) * TC_STAGE_BYTES;
    const int t = -v;
    const bool neg = t < 0;
    int coarse = (neg ? -t : t) / TC_BIAS_COARSE;          // magnitudes; both parts carry t's sign
    int fine = (neg ? -t : t) % TC_BIAS_COARSE;
    uint32_t nib[64];
    for (int e = 0; e < 32; ++e) {                          // coarse digits, integers from {6,4,3,2,1}
        const int d = coarse > 0 ? tc_e2m1_int_floor(coarse) : 0;
        coarse -= d;
        nib[e] = tc_e2m1_nibble(2 * d, neg);
    }
    nib[32] = tc_e2m1_nibble(1, false);                     // +0.5
    for (int e = 33; e < 64; ++e) {
        const int d = fine > 0 ? tc_e2m1_int_floor(fine) : 0;
        fine -= d;
        nib[e] = tc_e2m1_nibble(2 * d, neg);
    }
    for (int kb = 0; kb < 32; ++kb)
        blk[(n / 8) * 256 + (kb / 16) * 128 + (n % 8) * 16 + (kb % 16)] = (uint8_t)(nib[2 * kb] | (nib[2 * kb + 1] << 4));
}

// ---- the scan ------------------------------------------------------------------------------------------
// MODE 0: append survivors (search).  MODE 1: write all distances (dist_out[q*stride + row]).
// prof (optional, tools/tc_probe): cycle accounting of CTA 0.
template <int NCHUNK, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_scan_kernel(const uint4* __restrict__ codes, const uint32_t* __restrict__ live, uint32_t tile_lo,
               uint32_t tile_hi, const int8_t* __restrict__ qexp, const uint32_t* __restrict__ qpop,
               const int32_t* __restrict__ qbias, uint32_t nq, uint32_t nq_pad, uint32_t n_qslices,
               uint32_t n_rslices, uint32_t qb_item /* query blocks per item, <= tc_qblocks(NCHUNK) */,
               uint2* __restrict__ recs, uint32_t rec_cap,
               uint32_t* __restrict__ list_counts, uint32_t* __restrict__ overflow,
               uint32_t* __restrict__ dist_out, uint64_t dist_stride, uint64_t n_rows, int dbg = 0,
               unsigned long long* __restrict__ prof = nullptr) {
    constexpr int QB = tc_qblocks(NCHUNK);     // resident query blocks
    constexpr int NBUF = QB < 2 ? 2 : QB;      // accumulator buffers (block it uses buffer it % NBUF)
    constexpr int SC = tc_slot_chunks(NCHUNK); // chunks per A slot
    constexpr int PH = (NCHUNK + SC - 1) / SC; // phases per row group
    constexpr int CHUNK_COLS = 16;             // TMEM columns per expanded code chunk (128 e2m1 = 64 B per row)
    constexpr int A_COLS = 2 * SC * CHUNK_COLS;   // TMEM columns of the A ring
    // then: 8 columns bias slice (64 x 1.0), 4 columns scale words 1.0, 4 columns the bias MMA's A scales
    constexpr uint32_t IDESC = tc_idesc_mxf4(TC_ROWS, TC_NQ);
    constexpr uint32_t QBLOCK_BYTES = (uint32_t)tc_qblock_bytes(NCHUNK);
    static_assert(QB >= 1, "the resident query block must fit shared memory");
    static_assert(A_COLS + 16 + NBUF * TC_NQ <= 512, "TMEM budget");
    static_assert(QB * tc_qblock_bytes(NCHUNK) <= 216 * 1024, "shared memory budget");

    extern __shared__ __align__(1024) uint8_t smem[];      // QB resident query blocks
    __shared__ int32_t s_bias[QB * TC_NQ];                   // MODE 1 only
    __shared__ uint32_t s_pop[QB * TC_NQ];
    __shared__ __align__(8) uint64_t bars[6 + 2 * TC_MAX_QBLOCKS];
    __shared__ uint32_t s_tmem_base;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t b_full = bar0, b_free = bar0 + 8;
    auto a_ready = [&](int h) { return bar0 + 8u * (2 + h); };
    auto a_free = [&](int h) { return bar0 + 8u * (4 + h); };
    auto acc_full = [&](int b) { return bar0 + 8u * (6 + b); };
    auto acc_empty = [&](int b) { return bar0 + 8u * (6 + TC_MAX_QBLOCKS + b); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nqb = nq_pad / TC_NQ;
    const uint32_t ngroups = (tile_hi - tile_lo + 3) / 4;
    const uint32_t n_items = n_qslices * n_rslices;

    if (threadIdx.x == 0) {
        mbar_init(b_full, 1); mbar_init(b_free, 1);
        for (int h = 0; h < 2; ++h) { mbar_init(a_ready(h), 4); mbar_init(a_free(h), 1); }
        for (int b = 0; b < NBUF; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 4); }
        fence_mbar_init();
    }
    constexpr int MMA_WARP = 4 + TC_EPI_WARPS;
    if (warp == MMA_WARP) tc_alloc(smem_u32(&s_tmem_base), TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;
    const uint32_t tmem_a = tmem;                 // columns [0, A_COLS): the A ring; [A_COLS, A_COLS+8): bias slice
    const uint32_t tmem_sf_one = tmem + A_COLS + 8;      // 4 columns of scale words 1.0
    const uint32_t tmem_sf_bias = tmem + A_COLS + 12;    // 4 columns: the bias MMA's A scales (16, 1)
    const uint32_t tmem_d = tmem + A_COLS + 16;          // NBUF accumulator buffers of TC_NQ columns
    const uint32_t lane_taddr = (uint32_t)((warp & 3) * 32) << 16;

    unsigned long long pw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define TC_PROF_T0() const long long t0__ = prof ? clock64() : 0
#define TC_PROF_ADD(i) do { if (prof) pw[i] += (unsigned long long)(clock64() - t0__); } while (0)

    // item -> (query blocks, row groups)
    auto item_range = [&](uint32_t item, uint32_t& qb0, uint32_t& nb, uint32_t& g_lo, uint32_t& g_hi) {
        const uint32_t qsl = item / n_rslices, rsl = item % n_rslices;
        qb0 = qsl * qb_item;
        nb = min(qb_item, nqb - qb0);
        g_lo = (uint32_t)((uint64_t)ngroups * rsl / n_rslices);
        g_hi = (uint32_t)((uint64_t)ngroups * (rsl + 1) / n_rslices);
    };

    if (warp < 4) {
        // ===================== expanders: codes -> A operand in TMEM =====================
        uint32_t free_phase[2] = {1, 1};          // first wait on a fresh barrier passes
        {
            uint32_t ones[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) ones[i] = TC_A_ONE8;
            tc_st8(tmem_a + lane_taddr + A_COLS, ones);     // A slice of the bias MMA: 64 x 1.0
#pragma unroll
            for (int i = 0; i < 8; ++i) ones[i] = i < 4 ? TC_SF_ONE : TC_SF_BIAS_A;
            tc_st8(tmem_sf_one + lane_taddr, ones);         // scale words (read by every MMA)
            tc_wait_st();
        }
        auto load_codes = [&](uint32_t g, uint4 (&r)[NCHUNK]) {
            const uint32_t tile = tile_lo + g * 4 + warp;
            const bool in_range = tile < tile_hi;
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c)
                r[c] = in_range ? ldg_stream(codes + ((size_t)tile * NCHUNK + c) * 32 + lane) : make_uint4(0, 0, 0, 0);
        };
        auto expand = [&](const uint4 (&r)[NCHUNK], int c_lo, int c_hi, int h) {   // chunks [c_lo, c_hi) -> slot h
            { TC_PROF_T0(); mbar_wait(a_free(h), free_phase[h]); TC_PROF_ADD(h); }   // MMAs that read this half have retired
            free_phase[h] ^= 1u;
            tc_fence_after();
            TC_PROF_T0();
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                if (c < c_lo || c >= c_hi) continue;
                const uint32_t w4[4] = {r[c].x, r[c].y, r[c].z, r[c].w};
#pragma unroll
                for (int wp = 0; wp < 2; ++wp) {                  // two code words -> 8 columns (one K=64 MMA)
                    uint32_t v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t w = w4[wp * 2 + (i >> 2)];
                        const int sh = i & 3;                        // nibble j <- code bit 4j + sh
                        v[i] = (sh == 0 ? (w << 1) : (w >> (sh - 1))) & TC_A_ONE8;
                    }
                    tc_st8(tmem_a + lane_taddr + (uint32_t)(h * SC * CHUNK_COLS + ((c - c_lo) * 2 + wp) * 8), v);
                }
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready(h));
            TC_PROF_ADD(2 + h);
        };
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            uint32_t qb0, nb, g_lo, g_hi;
            item_range(item, qb0, nb, g_lo, g_hi);
            uint4 r[NCHUNK];
            if (g_lo < g_hi) load_codes(g_lo, r);
            for (uint32_t g = g_lo; g < g_hi; ++g) {
                uint4 rn[NCHUNK];
                if (g + 1 < g_hi) load_codes(g + 1, rn);       // in flight while this group expands
#pragma unroll
                for (int ph = 0; ph < PH; ++ph)
                    expand(r, ph * SC, (ph + 1) * SC < NCHUNK ? (ph + 1) * SC : NCHUNK, ph & 1);
#pragma unroll
                for (int c = 0; c < NCHUNK; ++c) r[c] = rn[c];
            }
        }
        if (prof && blockIdx.x == 0 && threadIdx.x == 0)
            for (int i = 0; i < 4; ++i) prof[i] = pw[i];
    } else if (warp < MMA_WARP) {
        // ===================== epilogue: accumulators -> survivors / distances =====================
        // two sets of four warps (set s takes the accumulator blocks with it % 2 == s): reading and
        // sign-testing a block takes about as long as the FP4 MMAs that produce it
        const int ew = warp - 4;                  // 0 .. TC_EPI_WARPS-1; TMEM lane quarter = ew & 3
        const uint32_t my_set = (uint32_t)(ew >> 2);
        uint32_t it = 0;                          // running accumulator-block counter (buffer = it % NBUF)
        // Survivor records of this warp: a private list, slots handed out with ballot + popc from a
        // register counter.  No shared memory and no atomics in the epilogue.
        uint2* my_list = recs + (size_t)(blockIdx.x * TC_EPI_WARPS + ew) * rec_cap;
        uint32_t my_count = 0;
        const uint32_t lane_lt = (1u << lane) - 1u;
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            uint32_t qb0, nb, g_lo, g_hi;
            item_range(item, qb0, nb, g_lo, g_hi);
            if (MODE == 1) {
                asm volatile("bar.sync 1, %0;" :: "n"(32 * TC_EPI_WARPS) : "memory");      // previous item's readers are done
                for (uint32_t i = threadIdx.x - 128; i < nb * TC_NQ; i += 32 * TC_EPI_WARPS) {
                    const uint32_t q = qb0 * TC_NQ + i;
                    s_bias[i] = qbias[q];
                    s_pop[i] = q < nq ? qpop[q] : 0u;
                }
                asm volatile("bar.sync 1, %0;" :: "n"(32 * TC_EPI_WARPS) : "memory");
            }
            for (uint32_t g = g_lo; g < g_hi; ++g) {
                const uint32_t tile = tile_lo + g * 4 + (ew & 3);
                const bool in_range = tile < tile_hi;
                const uint32_t row = tile * 32u + lane;
                const bool alive = in_range && ((live[in_range ? tile : tile_lo] >> lane) & 1u);
                for (uint32_t blk = 0; blk < nb; ++blk, ++it) {
                    if (TC_EPI_WARPS == 8 && (it & 1u) != my_set) continue;
                    const uint32_t b = it % NBUF;
                    const uint32_t qbase = (qb0 + blk) * TC_NQ, qloc = blk * TC_NQ;
                    { TC_PROF_T0(); mbar_wait(acc_full(b), (it / NBUF) & 1u); TC_PROF_ADD(0); }   // buffer b's (it / NBUF)-th use
                    tc_fence_after();
                    TC_PROF_T0();
#pragma unroll 1
                    for (int half = 0; half < ((dbg & 1) ? 0 : 2); ++half) {
                        uint32_t v0[32], v1[32];
                        const uint32_t col0 = tmem_d + lane_taddr + b * TC_NQ + half * 64;
                        tc_ld32(col0, v0);
                        tc_ld32(col0 + 32, v1);
                        tc_wait_ld();
                        if (MODE == 0) {
                            // D = -(S + bias) + 0.5 (f32): a survivor has D < 0.  Collect the 64 sign bits
                            // with one funnel shift per element (four independent chains); survivors are rare.
                            uint32_t ma = 0, mb = 0, mc = 0, md = 0;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                ma = __funnelshift_l(v0[j], ma, 1);
                                mb = __funnelshift_l(v0[16 + j], mb, 1);
                                mc = __funnelshift_l(v1[j], mc, 1);
                                md = __funnelshift_l(v1[16 + j], md, 1);
                            }
                            // element e of this half-block sits at bit 63 - e
                            uint64_t mask = ((uint64_t)((ma << 16) | mb) << 32) | (uint64_t)((mc << 16) | md);
                            if (!alive) mask = 0;
                            const uint32_t q0 = qbase + half * 64;
                            while (__any_sync(0xffffffffu, mask != 0)) {
                                const bool has = mask != 0;
                                const uint32_t m = __ballot_sync(0xffffffffu, has);
                                if (has) {
                                    const int e = __clzll((long long)mask);
                                    mask &= ~(0x8000000000000000ull >> e);
                                    const uint32_t slot = my_count + __popc(m & lane_lt);
                                    if (slot < rec_cap) my_list[slot] = make_uint2(row, q0 + e);
                                }
                                my_count += __popc(m);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 64; ++j) {
                                const uint32_t q = qbase + half * 64 + j;
                                const uint32_t ql = qloc + half * 64 + j;
                                const int32_t dv = (int32_t)(__uint_as_float(j < 32 ? v0[j & 31] : v1[j & 31]) - 0.5f);   // -(S + bias)
                                // hamming = popc(q) - S = popc(q) + bias + (D - 0.5)
                                if (q < nq && in_range && row < n_rows)
                                    dist_out[(size_t)q * dist_stride + row] = (uint32_t)((int32_t)s_pop[ql] + s_bias[ql] + dv);
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(b));
                    TC_PROF_ADD(1);
                }
            }
        }
        if (MODE == 0 && lane == 0) {
            list_counts[blockIdx.x * TC_EPI_WARPS + ew] = min(my_count, rec_cap);
            if (my_count > rec_cap) *overflow = 1u;
        }
        if (prof && blockIdx.x == 0 && threadIdx.x == 128)
            for (int i = 0; i < 2; ++i) prof[4 + i] = pw[i];
    } else {
        // ===================== query loader (TMA) + MMA issuer: warp 8, one elected lane issues =====================
        uint32_t bfull_phase = 0, bfree_phase = 0;
        uint32_t ready_phase[2] = {0, 0};
        uint32_t empty_phase[NBUF];               // first use of each accumulator buffer passes
#pragma unroll
        for (int b = 0; b < NBUF; ++b) empty_phase[b] = 1u;
        uint32_t it = 0;
        bool first_item = true;
        const long long t_begin = prof ? clock64() : 0;
        unsigned long long ns_begin = 0;
        if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_begin));
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            uint32_t qb0, nb, g_lo, g_hi;
            item_range(item, qb0, nb, g_lo, g_hi);
            if (g_lo >= g_hi) continue;
            if (!first_item) {                    // the previous item's MMAs still read the resident blocks
                mbar_wait(b_free, bfree_phase);
                bfree_phase ^= 1u;
            }
            first_item = false;
            if (elect_one()) {
                mbar_expect_tx(b_full, nb * QBLOCK_BYTES);
                for (uint32_t blk = 0; blk < nb; ++blk) {
                    const int8_t* src = qexp + (size_t)(qb0 + blk) * QBLOCK_BYTES;
                    const uint32_t dst = smem_u32(smem + blk * QBLOCK_BYTES);
                    for (uint32_t off = 0; off < QBLOCK_BYTES; off += TC_STAGE_BYTES)
                        tma_bulk_g2s(dst + off, src + off, min((uint32_t)TC_STAGE_BYTES, QBLOCK_BYTES - off), b_full);
                }
            }
            __syncwarp();
            { TC_PROF_T0(); mbar_wait(b_full, bfull_phase); TC_PROF_ADD(3); }
            bfull_phase ^= 1u;
            for (uint32_t g = g_lo; g < g_hi; ++g) {
                for (uint32_t blk = 0; blk < nb; ++blk, ++it) {
                    const uint32_t b = it % NBUF;
                    const bool last_blk = blk + 1 == nb;
                    { TC_PROF_T0(); mbar_wait(acc_empty(b), empty_phase[b]); TC_PROF_ADD(0); }
                    empty_phase[b] ^= 1u;
                    tc_fence_after();
                    const uint64_t bdesc0 = tc_smem_desc(smem_u32(smem + blk * QBLOCK_BYTES), 128, 1024);
                    const uint32_t d_addr = tmem_d + b * TC_NQ;
#pragma unroll
                    for (int ks = 0; ks < NCHUNK; ++ks) {              // one code chunk = 64 K bytes = two K=64 MMAs
                        const int ph = ks / SC, h = ph & 1, kc = ks % SC;   // phase, slot, chunk in slot
                        if (blk == 0 && kc == 0) {
                            { TC_PROF_T0(); mbar_wait(a_ready(h), ready_phase[h]); TC_PROF_ADD(1 + h); }
                            ready_phase[h] ^= 1u;
                            tc_fence_after();
                        }
                        if (elect_one()) {
#pragma unroll
                            for (int j = 0; j < 2; ++j)                 // the address field counts 16-byte units
                                tc_mma_mxf4_ts(d_addr, tmem_a + (uint32_t)(h * SC * CHUNK_COLS + (kc * 2 + j) * 8),
                                               bdesc0 + (uint64_t)(((ks / 2) * TC_STAGE_BYTES + ((ks % 2) * 4 + j * 2) * 128) >> 4),
                                               IDESC, tmem_sf_one, tmem_sf_one, (ks | j) != 0 ? 1u : 0u);
                            // this slot of A may be rewritten once the MMAs issued so far retire
                            if (last_blk && kc == SC - 1 && ks != NCHUNK - 1) tc_commit(a_free(h));
                        }
                        __syncwarp();
                    }
                    if (elect_one()) {
                        // bias: D += 1.0(128 x 64, scales 16 | 1) * digits(128 queries x 64) = -v + 0.5
                        tc_mma_mxf4_ts(d_addr, tmem_a + A_COLS,
                                       tc_smem_desc(smem_u32(smem + blk * QBLOCK_BYTES + tc_subblocks(NCHUNK) * TC_STAGE_BYTES), 128, 256),
                                       IDESC, tmem_sf_bias, tmem_sf_one, 1u);
                        if (last_blk) tc_commit(a_free(((NCHUNK - 1) / SC) & 1));   // the last phase's slot
                        tc_commit(acc_full(b));
                    }
                    __syncwarp();
                }
            }
            if (elect_one()) tc_commit(b_free);
            __syncwarp();
        }
        if (prof && blockIdx.x == 0 && lane == 0) {
            unsigned long long ns;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
            pw[6] = (unsigned long long)(clock64() - t_begin);
            pw[7] = ns - ns_begin;
            for (int i = 0; i < 8; ++i) prof[8 + i] = pw[i];
        }
    }
#undef TC_PROF_T0
#undef TC_PROF_ADD
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        __syncwarp();
        tc_dealloc(tmem, TC_TMEM_COLS);
    }
}

// Warp-private survivor records (row, query) -> per-query candidate buffers, same contract as
// scan_kernel: key = hamming << 32 | row.  The distance is recomputed from the codes (xor + popc
// over nchunk*4 words, served from L2) — cheaper than carrying it through the epilogue.
// grid = (TC_SCATTER_X, number of lists = 4 x scan CTAs), each CTA strides over its list.
constexpr int TC_SCATTER_X = 4;
__global__ void tc_scatter_kernel(const uint2* __restrict__ recs, uint32_t rec_cap,
                                  const uint32_t* __restrict__ list_counts, const uint4* __restrict__ codes,
                                  int nchunk, const uint32_t* __restrict__ qpack, int qs,
                                  uint32_t* __restrict__ cnt, uint64_t* __restrict__ buf, uint32_t cap,
                                  uint32_t* __restrict__ overflow) {
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
        const uint32_t pos = atomicAdd(&cnt[(size_t)q * CNT_STRIDE], 1u);
        if (pos < cap) buf[(size_t)q * cap + pos] = ((uint64_t)d << 32) | row;
        else *overflow = 1u;
    }
}

}  // namespace gvdb
