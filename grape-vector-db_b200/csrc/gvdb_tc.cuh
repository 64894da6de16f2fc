// gvdb_tc.cuh — the batched Hamming scan on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Used when a corpus pass serves a large query tile (>= 64 queries), where the scan really is a
// dense contraction: rows x queries x dim.  The CUDA-core kernel (scan_kernel) is then bound by
// the integer pipes (16 popc + 64 LOP3 per clk per SM).  A 1-bit code needs no more than 4 bits
// per element on the tensor core, so the contraction runs on the FP4 path:
// tcgen05.mma.kind::mxf4 (e2m1 operands, one UE8M0 scale per 32 elements, all scales 1.0, f32
// accumulation).  Measured with tools/mxf4_probe.cu: one M128 x N128 x K64 MMA with A in TMEM
// retires every 64.1 clk = 16.35k MAC/clk/SM (107 clk with A in shared memory; kind::i8 and
// kind::f8f6f4 do 8.2k, tools/mma_floor.cu).
//
// Exactness.  Row codes are expanded on chip to A in {0, 1.0} (nibbles 0x0 / 0x2), query codes
// once per batch to B in {-1.0, +1.0} (0xA where the query bit is 1, 0x2 where it is 0).  With
// S = n11 - n10 over the code bits,  D = sum_k A_k * B_k = -S  and  hamming = popc(q) - S
// = popc(q) + D.  All products and partial sums are integers far below 2^24: the f32
// accumulation is exact.  Pad bits are 0 in A and contribute nothing.
// One row word expands with a shift + a LOP3 per output word,
//     out[4*w + i] = ((code_word[w] >> i) << 1) & 0x22222222       (nibble j <- code bit 4j + i),
// a fixed permutation of the code bits applied to rows and queries alike (Hamming distance is
// invariant to it).
//
// Three epilogues on the same mainloop:
//   MODE 0  search: one extra K=64 MMA per block adds the per-query bias -(tau - popc(q)) + 0.5, so
//           hamming < tau  <=>  D < 0: the epilogue ORs raw register bits and looks closer only where
//           a sign bit shows up.  Survivors (row, query) leave through per-warp record lists
//           (coalesced stores, no atomics); tc_scatter_kernel turns them into candidate keys.
//   MODE 1  every distance to dist_out (gvdb_hamming, parity tests).
//   MODE 3  count: per query, the number of live rows below its threshold -> dist_out[q] (atomic adds of
//           bit-sliced per-lane counters): the b* search of ratio mode (gvdb_ratio.cuh).
//   MODE 2  sample: per query, the minimum of D over each ROW CLASS (the rows four neighbouring lanes of
//           one CTA see: four rows of every group of its row slice) — tilemin[class][q], one FMNMX per
//           element.  The threshold estimate of the single-pass search comes from these (tc_tau_kernel).
//
// Work decomposition.  item = (query slice, row slice).  A query slice is up to tc_qblocks() blocks
// of 128 queries whose expanded codes stay RESIDENT in shared memory (4 blocks x 48 KB at 768 bits)
// for the whole item; the item's rows stream through TMEM as the A operand, 128 at a time.
// (Streaming the queries through a shared-memory ring instead needs more L2->SM bandwidth than the
// L2 delivers to 148 SMs at once: that version measured 47 % of the MMA floor.)
//
// Warp roles per CTA (one CTA per SM, persistent over items):
//   warps 0-3   expanders: lane = row.  Load the row's code (coalesced, blocked layout, next group
//               prefetched), expand it and write it into TMEM as the A operand (tcgen05.st).  A is a
//               ring of FOUR slots of tc_slot_chunks() code chunks with their own ready/free
//               barriers: up to 768 bits the ring holds two whole row groups, so expanding group
//               g + 1 overlaps every MMA of group g.
//   warps 4-11  epilogue, two sets of four (set s owns accumulator buffer s): tcgen05.ld, test,
//               append / write / reduce.
//   warp 12     query loader (TMA bulk copies) + MMA issuer: M=128, N=128, K=64; A from TMEM, B from
//               shared memory, D in TMEM (two accumulator buffers).  The TMEM base is the constant 0
//               (the CTA allocates all 512 columns), so every operand of the unrolled MMA sequence is
//               a uniform-register value and the issue loop stays ahead of the tensor pipe.
//   mbarriers   b_full/b_free, a_ready/a_free per A slot, acc_full/acc_empty per accumulator
//               buffer; tcgen05.commit signals MMA completion.
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
constexpr int TC_MAX_QBLOCKS = 4;                   // resident query blocks per item (shared memory)
__host__ __device__ constexpr int tc_subblocks(int nchunk) { return (nchunk + 1) / 2; }
__host__ __device__ constexpr size_t tc_qblock_bytes(int nchunk) {
    return (size_t)tc_subblocks(nchunk) * TC_STAGE_BYTES + TC_BIAS_BYTES;
}
// The A operand lives in a ring of FOUR TMEM slots of tc_slot_chunks() code chunks (16 columns each, at
// most 3 chunks: 4 x 48 = 192 columns); a row group is expanded and consumed in tc_phases() phases,
// phase p (a running counter over groups and items) using slot p % 4.  Up to 768 bits the ring holds
// two whole row groups: expanding group g + 1 overlaps every MMA of group g.
// Next to it: 16 scale columns and TWO accumulator buffers of 128 columns.  What was measured on the way
// (1M x 1024 x 768, tools/tc_probe): three accumulators + a one-group ring 0.31 ms; one-chunk slots (12 of
// them) + three accumulators 0.36 ms (the refill chain of a slot is longer than five chunks of MMAs);
// accumulators of 64 columns 0.44 ms (an N = 64 MMA takes as long as an N = 128 one when A changes).
// With two accumulators the epilogue must hand a buffer back fast: every epilogue warp copies its
// 64 columns to registers and releases the buffer BEFORE testing them.
constexpr int TC_NBUF = 2;
constexpr int TC_NSLOT = 4;
__host__ __device__ constexpr int tc_slot_chunks(int nchunk) {
    return nchunk <= 6 ? (nchunk + 1) / 2 : (nchunk % 3 == 0 ? 3 : 2);
}
__host__ __device__ constexpr int tc_phases(int nchunk) { return (nchunk + tc_slot_chunks(nchunk) - 1) / tc_slot_chunks(nchunk); }
// Resident query blocks per item: what fits 216 KB of shared memory, at most four; a single block
// when the ring cannot hold a whole row group (every further block would need all the phases again).
__host__ __device__ constexpr int tc_qblocks(int nchunk) {
    int q = (int)((216 * 1024) / tc_qblock_bytes(nchunk));
    if (TC_MAX_QBLOCKS < q) q = TC_MAX_QBLOCKS;
    if (tc_phases(nchunk) > TC_NSLOT) q = q < 1 ? q : 1;
    return q;
}
__host__ __device__ constexpr bool tc_supported_chunks(int nchunk) {
    return nchunk == 1 || nchunk == 2 || nchunk == 3 || nchunk == 4 || nchunk == 6 || nchunk == 8 || nchunk == 12 ||
           nchunk == 16 || nchunk == 24;
}
constexpr int TC_EPI_WARPS = 8;      // two sets of four epilogue warps, one per accumulator buffer
constexpr int TC_THREADS = 32 * (4 + TC_EPI_WARPS + 1);   // 4 expander warps + epilogue warps + loader/MMA-issuer warp
constexpr uint32_t TC_TMEM_COLS = 512;
constexpr int TC_QUEUE = 64;         // survivor queue entries per epilogue warp (MODE 0)
constexpr int32_t TC_TILEMIN_NONE = 0x7fffffff;     // MODE 2: no live row in the tile

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// One lane of a converged warp; the surrounding control flow stays warp-uniform.
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
// one column: lane l gets TMEM[lane quarter base + l][col]
__device__ __forceinline__ uint32_t tc_ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return v;
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
// digit e (0, 1, 2, ...) of the greedy decomposition of a magnitude into e2m1 integers {6, 4, 3, 2, 1}:
// mag / 6 sixes, then the remainder r < 6 as one digit (r = 5: 4 then 1)
__host__ __device__ inline int tc_bias_digit(int mag, int e) {
    const int n6 = mag / 6, r = mag % 6;
    if (e < n6) return 6;
    if (e == n6) return r == 5 ? 4 : r;
    if (e == n6 + 1) return r == 5 ? 1 : 0;
    return 0;
}
// Per-query bias of the search epilogue, added ON the tensor core by one extra K=64 MMA per accumulator
// block: A = 64 x 1.0 per row with scale factors (16, 1) for its two 32-element halves, B = 32 coarse +
// 31 fine e2m1 digits with 16 * coarse + fine = -v, plus one digit 0.5.  With v = tau - popc(q) the
// accumulator becomes D = -(S + v) + 0.5: never zero, and hamming < tau <=> D < 0 — the epilogue is a
// sign test on raw register bits (no per-element threshold arithmetic, which costs the issue slots the
// MMA warp needs: tools/mma_contention_probe.cu).  v = K+1: everything passes (tau = TAU_ALL);
// v = -(K+1): nothing does (padding queries); v = 0: MODE 1, hamming = popc(q) + D - 0.5.
// One WARP writes the 32 digit bytes of query q (lane = K byte) in the K-major core-matrix order of a
// 128 x 32 B block:  offset(n, kb) = (n/8)*256 + (kb/16)*128 + (n%8)*16 + (kb%16),  element e in byte e/2.
__device__ __forceinline__ void tc_write_bias_digits(int8_t* __restrict__ qexp, int nchunk, uint32_t q, int v, int lane) {
    const uint32_t qb = q / TC_NQ, n = q % TC_NQ;
    uint8_t* blk = reinterpret_cast<uint8_t*>(qexp) + (size_t)qb * tc_qblock_bytes(nchunk) + (size_t)tc_subblocks(nchunk) * TC_STAGE_BYTES;
    const int t = -v;
    const bool neg = t < 0;
    const int mag = neg ? -t : t;
    const int coarse = mag / TC_BIAS_COARSE, fine = mag % TC_BIAS_COARSE;
    uint32_t nib[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int e = 2 * lane + h;                         // element 0..63
        if (e < 32) nib[h] = tc_e2m1_nibble(2 * tc_bias_digit(coarse, e), neg);
        else if (e == 32) nib[h] = tc_e2m1_nibble(1, false);            // +0.5
        else nib[h] = tc_e2m1_nibble(2 * tc_bias_digit(fine, e - 33), neg);
    }
    blk[(n / 8) * 256 + (lane / 16) * 128 + (n % 8) * 16 + (lane % 16)] = (uint8_t)(nib[0] | (nib[1] << 4));
}
__host__ __device__ inline int tc_bias_v(uint32_t tau, uint32_t pop, int K) {
    if (tau == TAU_ALL) return K + 1;
    int v = (int)min(tau, 1u << 20) - (int)pop;
    return v > K + 1 ? K + 1 : (v < -(K + 1) ? -(K + 1) : v);
}

// Byte offset, inside qexp, of output word o (= 4*w + i: code word w, shift i; 8 K elements = 4 K bytes)
// of query q.  Sub-block (qb, sb) of 16 KB at qb*tc_qblock_bytes + sb*16384; inside a sub-block
//   offset(n, kk) = (n/8)*1024 + (kk/16)*128 + (n%8)*16 + (kk%16),  n = query in block, kk = K byte
// K element e of a query  <->  code bit 32*w + 4*j + i  with  e = (4*w + i)*8 + j  (see header);
// element e sits in K byte e/2, low nibble for even e.
__host__ __device__ inline size_t tc_qexp_offset(uint32_t q, int o, int nchunk) {
    const uint32_t qb = q / TC_NQ, n = q % TC_NQ;
    const int kb = o * 4;
    const int sb = kb / TC_KSTAGE, kk = kb % TC_KSTAGE;
    return (size_t)qb * tc_qblock_bytes(nchunk) + (size_t)sb * TC_STAGE_BYTES + (n / 8) * 1024 + (kk / 16) * 128 +
           (n % 8) * 16 + (kk % 16);
}
// Query bit 1 -> -1.0 (0xA), 0 -> +1.0 (0x2).
__device__ __forceinline__ uint32_t tc_qexp_word(uint32_t code_word, int i) {
    return TC_A_ONE8 | (((code_word >> i) & 0x11111111u) << 3);
}

// ---- query pre-expansion (standalone: gvdb_hamming and the multi-segment schedule) ----------------------
// qpack (code words + tau, as produced by the query prep kernels) -> qexp (e2m1 +-1.0 nibbles in the byte
// order the resident blocks need) + popc(q).  Queries beyond nq (padding up to a multiple of 128) are all zero.
__global__ void tc_expand_queries_kernel(const uint32_t* __restrict__ qpack, int qs, int nchunk, uint32_t nq,
                                         uint32_t nq_pad, int8_t* __restrict__ qexp,
                                         uint32_t* __restrict__ qpop) {
    const uint32_t q = blockIdx.x;                 // one CTA per (padded) query
    if (q >= nq_pad) return;
    const int nwords_out = tc_subblocks(nchunk) * (TC_KSTAGE / 4);   // 32-bit output words per query (incl. pad)
    uint32_t pop = 0;
    for (int o = threadIdx.x; o < nwords_out; o += blockDim.x) {   // output word o = 4*w + i: 8 elements, 4 K bytes
        const int w = o >> 2, i = o & 3;
        uint32_t out = 0;
        if (q < nq && w < nchunk * 4) {
            const uint32_t word = qpack[(size_t)q * qs + w];
            out = tc_qexp_word(word, i);
            if (i == 0) pop += __popc(word);
        }
        *reinterpret_cast<uint32_t*>(qexp + tc_qexp_offset(q, o, nchunk)) = out;
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

// Bias digits + qbase[q] = popc(q) + v for the current tau words of qpack (one warp per padded query).
// zero_bias: v = 0 (MODE 1).
__global__ void __launch_bounds__(256)
tc_bias_kernel(const uint32_t* __restrict__ qpack, int qs, int nchunk, const uint32_t* __restrict__ qpop,
               uint32_t nq, uint32_t nq_pad, int8_t* __restrict__ qexp, int32_t* __restrict__ qbase, int zero_bias) {
    const uint32_t q = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq_pad) return;
    const int K = nchunk * 128;
    int v = -(K + 1);
    if (q < nq) v = tc_bias_v(qpack[(size_t)q * qs + nchunk * 4], qpop[q], K);
    if (zero_bias) v = 0;
    if (lane == 0) qbase[q] = (q < nq ? (int32_t)qpop[q] : 0) + v;
    tc_write_bias_digits(qexp, nchunk, q, v, lane);
}

// ---- fused query preparation of the single-pass search (dim % 4 == 0) --------------------------------------
// One WARP per (padded) query: the row is read once with coalesced 128-bit loads and staged in shared
// memory; the lanes build the code words (strict '>' threshold, Msb0 bytes as BinaryVector::to_bytes(),
// /root/reference/src/quantization.rs:97-101), lane 0 folds ||q||^2 strictly left to right (separately
// rounded multiply and add: the reference's iterator sum, :208), then all lanes write qpack (code words
// + [tau = TAU_ALL, 0, 0, 0]), popc(q), the expanded e2m1 block entries and the query's zeroed candidate
// counter.  Replaces query_prep_direct_kernel + tc_expand_queries_kernel + a memset.
constexpr int QPT_WARPS = 8;
__global__ void __launch_bounds__(32 * QPT_WARPS)
query_prep_tc_kernel(const float* __restrict__ x, uint32_t nq, uint32_t nq_pad, int dim, float thr, int nchunk,
                     float* __restrict__ norms, uint32_t* __restrict__ qpack, int qs, int8_t* __restrict__ qexp,
                     uint32_t* __restrict__ qpop, uint32_t* __restrict__ cnt, uint32_t* __restrict__ flags) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float s_rows[];           // QPT_WARPS x dim floats, then QPT_WARPS x nchunk*4 words
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x < 8 && flags) flags[threadIdx.x] = 0u;
    const uint32_t q = blockIdx.x * QPT_WARPS + warp;
    if (q >= nq_pad) return;
    const int nw = nchunk * 4;
    float* mine = s_rows + (size_t)warp * dim;
    uint32_t* words = reinterpret_cast<uint32_t*>(s_rows + (size_t)QPT_WARPS * dim) + warp * nw;
    const bool real = q < nq;
    for (int w = lane; w < nw; w += 32) words[w] = 0u;
    __syncwarp();
    if (real) {
        const float4* src = reinterpret_cast<const float4*>(x + (size_t)q * dim);
        const int nv = dim >> 2;
        for (int v0 = 0; v0 < nv; v0 += 32) {
            const int v = v0 + lane;
            uint32_t bits = 0;
            if (v < nv) {
                const float4 f = __ldg(src + v);
                reinterpret_cast<float4*>(mine)[v] = f;
                const int e0 = (v & 7) * 4;               // element index inside the 32-bit word
                const float e4[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int e = e0 + t;
                    bits |= (uint32_t)(e4[t] > thr) << ((e & ~7) | (7 - (e & 7)));
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {                   // lanes 8k .. 8k+7 make word v0/8 + k
                const uint32_t wk = __reduce_or_sync(0xffffffffu, (lane >> 3) == k ? bits : 0u);
                if (lane == k && (v0 >> 3) + k < nw) words[(v0 >> 3) + k] = wk;
            }
        }
    }
    __syncwarp();
    if (real && lane == 0) {
        float ss = 0.0f;
        const float4* m4 = reinterpret_cast<const float4*>(mine);
        for (int v = 0; v < (dim >> 2); ++v) {
            const float4 f = m4[v];
            ss = __fadd_rn(ss, __fmul_rn(f.x, f.x));
            ss = __fadd_rn(ss, __fmul_rn(f.y, f.y));
            ss = __fadd_rn(ss, __fmul_rn(f.z, f.z));
            ss = __fadd_rn(ss, __fmul_rn(f.w, f.w));
        }
        norms[q] = __fsqrt_rn(ss);
        uint32_t* tau = qpack + (size_t)q * qs + nw;
        tau[0] = TAU_ALL; tau[1] = 0u; tau[2] = 0u; tau[3] = 0u;
        cnt[(size_t)q * CNT_STRIDE] = 0u;
    }
    uint32_t pop = 0;
    for (int w = lane; w < nw; w += 32) {
        const uint32_t word = words[w];
        if (real) qpack[(size_t)q * qs + w] = word;
        pop += __popc(word);
    }
    pop = __reduce_add_sync(0xffffffffu, pop);
    if (lane == 0) qpop[q] = pop;
    const int nwords_out = tc_subblocks(nchunk) * (TC_KSTAGE / 4);
    for (int o = lane; o < nwords_out; o += 32) {
        const int w = o >> 2, i = o & 3;
        const uint32_t out = (real && w < nw) ? tc_qexp_word(words[w], i) : 0u;
        *reinterpret_cast<uint32_t*>(qexp + tc_qexp_offset(q, o, nchunk)) = out;
    }
}

// ---- single-pass threshold from the sample's tile minima ------------------------------------------------
// One warp per query.  tilemin[t * nq_pad + q] (MODE 2) = min over the live rows of sample class t of
// D = hamming - popc(q) (TC_TILEMIN_NONE: no live row).  The minima of different classes belong to
// different rows, so the m-th smallest of them, t_m, is an upper bound of the m-th smallest distance
// among the sampled rows (and equals it unless two of the m closest sampled rows share a class):
//     tau_opt = popc(q) + t_m + 1      (TAU_ALL while fewer than m tiles hold a live row)
// Written to qpack's tau words [tau, tau_opt], as the bias digits of v = t_m + 1 and as qbase = popc(q) + v.
// Whether the pass under tau_opt found the true top R is checked afterwards (select_hist_kernel, verify 2).
constexpr int TC_TAU_WARPS = 8;
constexpr int TC_TAU_MAX_TILES = 4096;
__global__ void __launch_bounds__(32 * TC_TAU_WARPS)
tc_tau_kernel(const int32_t* __restrict__ tilemin, uint32_t n_tiles, uint32_t nq, uint32_t nq_pad, uint32_t m,
              const uint32_t* __restrict__ qpop, uint32_t* __restrict__ qpack, int qs, int nchunk,
              int8_t* __restrict__ qexp, int32_t* __restrict__ qbase) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ int16_t s_min[];                         // TC_TAU_WARPS x (n_tiles + 2)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q0 = blockIdx.x * TC_TAU_WARPS;
    const uint32_t q = q0 + warp;
    const int K = nchunk * 128;
    // the CTA's eight queries are neighbours in tilemin's rows: eight threads read one 32-byte sector
    // six loads in flight per thread (the loop would otherwise wait for every load before issuing the next:
    // 18 L2 round trips one after another were half of this kernel's 20 us)
    constexpr uint32_t TAU_U = 6;
    const uint32_t n_vals = n_tiles * TC_TAU_WARPS;
    for (uint32_t i0 = threadIdx.x; i0 < n_vals; i0 += blockDim.x * TAU_U) {
        int32_t v[TAU_U];
#pragma unroll
        for (uint32_t u = 0; u < TAU_U; ++u) {
            const uint32_t i = i0 + u * blockDim.x;
            const uint32_t t = i / TC_TAU_WARPS, w = i % TC_TAU_WARPS;
            v[u] = (i < n_vals && q0 + w < nq_pad) ? __ldg(tilemin + (size_t)t * nq_pad + q0 + w) : TC_TILEMIN_NONE;
        }
#pragma unroll
        for (uint32_t u = 0; u < TAU_U; ++u) {
            const uint32_t i = i0 + u * blockDim.x;
            if (i < n_vals) {
                const uint32_t t = i / TC_TAU_WARPS, w = i % TC_TAU_WARPS;
                s_min[(size_t)w * (n_tiles + 2) + t] = v[u] == TC_TILEMIN_NONE ? (int16_t)0x7fff : (int16_t)v[u];   // odd word stride
            }
        }
    }
    __syncthreads();
    if (q >= nq_pad) return;
    if (q >= nq) {                                             // padding query: nothing passes
        if (lane == 0) qbase[q] = -(K + 1);
        tc_write_bias_digits(qexp, nchunk, q, -(K + 1), lane);
        return;
    }
    int16_t* mine = s_min + (size_t)warp * (n_tiles + 2);
    int32_t tm = 0x7fff;
    for (uint32_t it = 0; it < m; ++it) {                      // extract the smallest, m times
        int32_t best = 0x7fff;
        uint32_t at = 0xffffffffu;
        for (uint32_t t = lane; t < n_tiles; t += 32) {
            const int32_t v = mine[t];
            if (v < best) { best = v; at = t; }
        }
        const int32_t wmin = __reduce_min_sync(0xffffffffu, best);
        tm = wmin;
        if (wmin == 0x7fff) break;                             // fewer than m live tiles
        const uint32_t owner = __ffs(__ballot_sync(0xffffffffu, best == wmin && at != 0xffffffffu)) - 1;
        if ((uint32_t)lane == owner) mine[at] = (int16_t)0x7fff;
        __syncwarp();
    }
    tm = __shfl_sync(0xffffffffu, tm, 0);
    const int32_t pop = (int32_t)qpop[q];
    int v = K + 1;                                             // fewer than m live tiles: everything passes
    uint32_t t = TAU_ALL;
    if (tm != 0x7fff) { t = (uint32_t)max(pop + tm + 1, 0); v = tc_bias_v(t, (uint32_t)pop, K); }
    if (lane == 0) {
        uint32_t* tau = qpack + (size_t)q * qs + nchunk * 4;
        tau[0] = t; tau[1] = t;
        qbase[q] = pop + v;
    }
    tc_write_bias_digits(qexp, nchunk, q, v, lane);
}

// ---- the scan ------------------------------------------------------------------------------------------
// Row groups: logical group g in [0, ngroups) covers tiles tile_lo + g * group_stride * 4 + {0,1,2,3}
// (group_stride = 1: a contiguous range; > 1: the strided sample of MODE 2).
// prof (optional, tools/tc_probe): cycle accounting of CTA 0.  dbg: timing experiments of tools/tc_probe
// (1 no accumulator reads, 2 no A stores) — 0 in the library.
template <int NCHUNK, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_scan_kernel(const uint4* __restrict__ codes, const uint32_t* __restrict__ live, uint32_t tile_lo,
               uint32_t tile_hi, uint32_t ngroups, uint32_t group_stride, const int8_t* __restrict__ qexp,
               const int32_t* __restrict__ qbase /* popc(q) + v, v = the bias of the query (MODE 0, 1) */, uint32_t nq, uint32_t nq_pad,
               uint32_t n_qslices, uint32_t n_rslices, uint32_t qb_item /* query blocks per item, <= tc_qblocks(NCHUNK) */,
               uint2* __restrict__ recs, uint32_t rec_cap, uint32_t* __restrict__ list_counts, uint32_t* __restrict__ overflow,
               uint32_t* __restrict__ dist_out, uint64_t dist_stride, uint64_t n_rows,
               int32_t* __restrict__ tilemin, int dbg = 0, unsigned long long* __restrict__ prof = nullptr) {
    // the sample pass lets the (small) threshold kernel in early; the main pass does not: scatter CTAs resident beside
    // it cost the MMA warp issue slots (measured: 0.507 ms per step with, 0.441 without)
    if (MODE == 2) pdl_launch_dependents();
    constexpr int QB = tc_qblocks(NCHUNK);     // resident query blocks
    constexpr int NBUF = TC_NBUF;              // accumulator buffers
    constexpr int NSLOT = TC_NSLOT;            // A ring slots
    constexpr int SC = tc_slot_chunks(NCHUNK); // chunks per A slot
    constexpr int PH = (NCHUNK + SC - 1) / SC; // phases per row group
    constexpr int CHUNK_COLS = 16;             // TMEM columns per expanded code chunk (128 e2m1 = 64 B per row)
    constexpr int SLOT_COLS = SC * CHUNK_COLS;
    constexpr int A_COLS = NSLOT * SLOT_COLS;  // TMEM columns of the A ring
    constexpr uint32_t IDESC = tc_idesc_mxf4(TC_ROWS, TC_NQ);
    constexpr uint32_t QBLOCK_BYTES = (uint32_t)tc_qblock_bytes(NCHUNK);
    static_assert(QB >= 1, "the resident query block must fit shared memory");
    static_assert(NBUF >= 2 && A_COLS + 16 + NBUF * TC_NQ <= 512, "TMEM budget");
    static_assert(QB * tc_qblock_bytes(NCHUNK) <= 216 * 1024, "shared memory budget");
    static_assert(QB == 1 || PH <= NSLOT, "several resident blocks need the whole row group in the A ring");

    extern __shared__ __align__(1024) uint8_t smem[];      // QB resident query blocks
    __shared__ int32_t s_base[QB * TC_NQ];                   // MODE 0, 1: popc(q) + v of the item's queries
    __shared__ uint32_t s_qrow[MODE == 0 ? TC_EPI_WARPS * TC_QUEUE : 1];   // MODE 0: survivor queues
    __shared__ uint32_t s_qq[MODE == 0 ? TC_EPI_WARPS * TC_QUEUE : 1];
    __shared__ __align__(8) uint64_t bars[2 + 2 * NSLOT + 2 * NBUF];
    __shared__ uint32_t s_tmem_base;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t b_full = bar0, b_free = bar0 + 8;
    auto a_ready = [&](uint32_t s) { return bar0 + 8u * (2 + s); };
    auto a_free = [&](uint32_t s) { return bar0 + 8u * (2 + NSLOT + s); };
    auto acc_full = [&](uint32_t b) { return bar0 + 8u * (2 + 2 * NSLOT + b); };
    auto acc_empty = [&](uint32_t b) { return bar0 + 8u * (2 + 2 * NSLOT + NBUF + b); };

    // the shuffle tells the compiler the warp index is warp-uniform: the role branches below and everything
    // derived from loop counters inside them can then live in uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t nqb = nq_pad / TC_NQ;
    const uint32_t n_items = n_qslices * n_rslices;

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
    pdl_wait();                                          // barriers and TMEM are set up under the predecessor's tail
    // The CTA owns all 512 columns (one CTA per SM: the shared memory does not admit a second one),
    // so the allocation starts at lane 0, column 0.  Using the constant keeps every TMEM address of
    // the MMA sequence in uniform registers.
    if (s_tmem_base != 0u) { if (threadIdx.x == 0 && overflow) overflow[2] = 1u; __trap(); }
    constexpr uint32_t tmem_a = 0;                       // columns [0, A_COLS): the A ring; [A_COLS, A_COLS+8): bias slice
    constexpr uint32_t tmem_sf_one = A_COLS + 8;         // 4 columns of scale words 1.0 (read by every MMA)
    constexpr uint32_t tmem_sf_bias = A_COLS + 12;       // 4 columns: the bias MMA's A scales (16, 1)
    constexpr uint32_t tmem_d = A_COLS + 16;             // NBUF accumulator buffers of TC_NQ columns
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
    auto group_tile = [&](uint32_t g) { return tile_lo + g * group_stride * 4u; };

    if (warp < 4) {
        // ===================== expanders: codes -> A operand in TMEM =====================
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
        uint32_t p = 0;                                     // running phase counter: slot p % NSLOT
        auto load_codes = [&](uint32_t g, uint4 (&r)[NCHUNK]) {
            const uint32_t tile = group_tile(g) + warp;
            const bool in_range = tile < tile_hi;
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c)
                r[c] = in_range ? ldg_stream(codes + ((size_t)tile * NCHUNK + c) * 32 + lane) : make_uint4(0, 0, 0, 0);
        };
        auto expand = [&](const uint4 (&r)[NCHUNK], int c_lo, int c_hi) {   // chunks [c_lo, c_hi) -> slot p % NSLOT
            const uint32_t s = p % NSLOT;
            { TC_PROF_T0(); mbar_wait(a_free(s), ((p / NSLOT) & 1u) ^ 1u); TC_PROF_ADD(0); }   // the MMAs that read this slot have retired
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
                    if (!(dbg & 2)) tc_st8(tmem_a + lane_taddr + s * SLOT_COLS + (uint32_t)(((c - c_lo) * 2 + wp) * 8), v);
                }
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready(s));
            ++p;
            TC_PROF_ADD(2);
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
                    expand(r, ph * SC, (ph + 1) * SC < NCHUNK ? (ph + 1) * SC : NCHUNK);
#pragma unroll
                for (int c = 0; c < NCHUNK; ++c) r[c] = rn[c];
            }
        }
        if (prof && blockIdx.x == 0 && threadIdx.x == 0)
            for (int i = 0; i < 4; ++i) prof[i] = pw[i];
    } else if (warp < MMA_WARP) {
        // ===================== epilogue: accumulators -> survivors / distances / tile minima =====================
        // Eight warps: warp ew owns TMEM lane quarter ew & 3 (32 rows) and column half ew >> 2 (64 queries)
        // of EVERY accumulator block.  It copies its 64 values to registers, hands the buffer back
        // (acc_empty counts all eight warps) and only then tests them: with two accumulator buffers
        // the MMAs of block i + 2 wait for this hand-over, not for the tests.
        const int ew = warp - 4;                  // 0 .. TC_EPI_WARPS-1
        const uint32_t half = (uint32_t)(ew >> 2);
        uint32_t it = 0;                          // running accumulator-block counter (buffer = it % NBUF)
        // MODE 0: survivors (row, query) wait in a per-warp shared-memory queue and leave 32 at a time as
        // one coalesced 256-byte store to the warp's private record list in global memory: no atomics,
        // no loads, nothing the epilogue has to wait for.  tc_scatter_kernel turns the records into keys.
        uint32_t* q_row = s_qrow + (MODE == 0 ? ew * TC_QUEUE : 0);
        uint32_t* q_q = s_qq + (MODE == 0 ? ew * TC_QUEUE : 0);
        uint32_t q_head = 0, q_count = 0, n_written = 0;
        uint2* my_list = recs + (size_t)(blockIdx.x * TC_EPI_WARPS + ew) * rec_cap;
        const uint32_t lane_lt = (1u << lane) - 1u;
        // MODE 3: bit-sliced per-column counters of this lane; they belong to ONE (query block, column half) at a time
        uint64_t planes[MODE == 3 ? 7 : 1];
#pragma unroll
        for (int i = 0; i < (MODE == 3 ? 7 : 1); ++i) planes[i] = 0ull;
        uint32_t plane_rows = 0, plane_q = 0;
        auto flush_counts = [&](uint32_t qbase_cols) {      // warp-collective: counts of columns qbase_cols .. +63 -> dist_out[q]
#pragma unroll 4
            for (int j = 0; j < 64; ++j) {
                uint32_t c = 0;
#pragma unroll
                for (int i = 0; i < (MODE == 3 ? 7 : 1); ++i) c |= (uint32_t)((planes[i] >> (63 - j)) & 1ull) << i;
                const uint32_t tot = __reduce_add_sync(0xffffffffu, c);
                if (lane == (j & 31) && tot) atomicAdd(&dist_out[qbase_cols + j], tot);
            }
#pragma unroll
            for (int i = 0; i < (MODE == 3 ? 7 : 1); ++i) planes[i] = 0ull;
        };
        auto flush = [&](uint32_t n_take) {                 // warp-collective: write the first n_take queued survivors
            __syncwarp();
            if ((uint32_t)lane < n_take && n_written + lane < rec_cap && !(dbg & 4)) {
                const uint32_t at = (q_head + lane) % TC_QUEUE;
                my_list[n_written + lane] = make_uint2(q_row[at], q_q[at]);
            }
            n_written += n_take;
            q_head = (q_head + n_take) % TC_QUEUE;
            q_count -= n_take;
            __syncwarp();
        };
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            uint32_t qb0, nb, g_lo, g_hi;
            item_range(item, qb0, nb, g_lo, g_hi);
            if (MODE != 2) {
                asm volatile("bar.sync 1, %0;" :: "n"(32 * TC_EPI_WARPS) : "memory");      // previous item's readers are done
                for (uint32_t i = threadIdx.x - 128; i < nb * TC_NQ; i += 32 * TC_EPI_WARPS)
                    s_base[i] = qbase[qb0 * TC_NQ + i];
                asm volatile("bar.sync 1, %0;" :: "n"(32 * TC_EPI_WARPS) : "memory");
            }
            // the tombstone word of the NEXT group is fetched while this one is processed (an L2 round trip)
            auto live_word = [&](uint32_t g) {
                const uint32_t tile = group_tile(g) + (ew & 3);
                return (g < g_hi && tile < tile_hi) ? __ldg(live + tile) : 0u;
            };
            uint32_t live_next = live_word(g_lo);
            float run_min[MODE == 2 ? 64 : 1];                // MODE 2 (one block per item): minima of this lane's rows
            if (MODE == 2) {
#pragma unroll
                for (int j = 0; j < 64; ++j) run_min[j < (MODE == 2 ? 64 : 1) ? j : 0] = __int_as_float(0x7f800000);
            }
            for (uint32_t g = g_lo; g < g_hi; ++g) {
                const uint32_t tile = group_tile(g) + (ew & 3);
                const bool in_range = tile < tile_hi;
                const uint32_t row = tile * 32u + lane;
                const bool alive = (live_next >> lane) & 1u;
                live_next = live_word(g + 1);
                for (uint32_t blk = 0; blk < nb; ++blk, ++it) {
                    const uint32_t b = it % NBUF;
                    const uint32_t qbase_q = (qb0 + blk) * TC_NQ + half * 64, qloc = blk * TC_NQ + half * 64;
                    if (MODE == 3 && plane_rows && plane_q != qbase_q) { flush_counts(plane_q); plane_rows = 0; }
                    { TC_PROF_T0(); mbar_wait(acc_full(b), (it / NBUF) & 1u); TC_PROF_ADD(0); }   // buffer b's (it / NBUF)-th use
                    tc_fence_after();
                    TC_PROF_T0();
                    if (MODE == 2) {
                        // sample mode keeps 64 running minima in registers: take the accumulator in two
                        // halves of 32 columns, then hand the buffer back
                        const uint32_t col0 = tmem_d + lane_taddr + b * TC_NQ + half * 64;
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            uint32_t w[32];
                            tc_ld32(col0 + h2 * 32, w);
                            tc_wait_ld();
                            if (alive) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const int jj = MODE == 2 ? h2 * 32 + j : 0;
                                    run_min[jj] = fminf(run_min[jj], __uint_as_float(w[j]));
                                }
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty(b));
                        TC_PROF_ADD(1);
                        continue;
                    }
                    uint32_t v[64];
                    if (!(dbg & 1)) {
                        uint32_t v0[32], v1[32];
                        const uint32_t col0 = tmem_d + lane_taddr + b * TC_NQ + half * 64;
                        tc_ld32(col0, v0);
                        tc_ld32(col0 + 32, v1);
                        tc_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) { v[j] = v0[j]; v[32 + j] = v1[j]; }
                    } else {
#pragma unroll
                        for (int j = 0; j < 64; ++j) v[j] = 0x3f800000u;
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(b));      // the values are in registers: the buffer is free
                    TC_PROF_ADD(2);
                    if (MODE == 0) {
                        // D = -(S + v) + 0.5 (f32): a survivor has D < 0, i.e. its sign bit set.  Survivors are
                        // rare: OR the raw bits of each quarter (16 columns) with three-input LOP3s and look
                        // closer only where some lane of the warp saw a sign bit.
                        uint64_t mask = 0;                         // element e of this half-block sits at bit 63 - e
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            uint32_t o = v[16 * k];
#pragma unroll
                            for (int j = 1; j < 15; j += 2) o |= v[16 * k + j] | v[16 * k + j + 1];
                            o |= v[16 * k + 15];
                            if (__any_sync(0xffffffffu, (int32_t)o < 0 && alive)) {
                                uint32_t mk = 0;
#pragma unroll
                                for (int j = 0; j < 16; ++j) mk = __funnelshift_l(v[16 * k + j], mk, 1);
                                if (alive) mask |= (uint64_t)mk << (48 - 16 * k);
                            }
                        }
                        // every lane files its own survivors (usually none, rarely more than one)
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
                    } else if (MODE == 3) {
                        // count, per query, the live rows below its threshold: the 64 sign bits of this lane's row
                        // are added into bit-sliced counters (7 planes: up to 127 rows per lane between flushes)
                        uint32_t ma = 0, mb = 0, mc = 0, md = 0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            ma = __funnelshift_l(v[j], ma, 1);
                            mb = __funnelshift_l(v[16 + j], mb, 1);
                            mc = __funnelshift_l(v[32 + j], mc, 1);
                            md = __funnelshift_l(v[48 + j], md, 1);
                        }
                        uint64_t carry = alive ? (((uint64_t)((ma << 16) | mb) << 32) | (uint64_t)((mc << 16) | md)) : 0ull;
#pragma unroll
                        for (int i = 0; i < 7; ++i) {
                            const uint64_t t = planes[MODE == 3 ? i : 0] & carry;
                            planes[MODE == 3 ? i : 0] ^= carry;
                            carry = t;
                        }
                        if (++plane_rows == 127u) { flush_counts(qbase_q); plane_rows = 0; }
                        plane_q = qbase_q;
                    } else if (MODE == 1) {
                        // a strided sample (group_stride > 1) is written compactly: column g * 128 + row-in-group
                        const size_t col = group_stride > 1 ? (size_t)g * TC_ROWS + (uint32_t)(ew & 3) * 32u + lane : (size_t)row;
#pragma unroll
                        for (int j = 0; j < 64; ++j) {
                            const uint32_t q = qbase_q + j;
                            const int32_t dv = (int32_t)(__uint_as_float(v[j]) - 0.5f);   // -(S + v), v = 0
                            if (q < nq && in_range && row < n_rows)
                                dist_out[(size_t)q * dist_stride + col] = (uint32_t)(s_base[qloc + j] + dv);
                            else if (q < nq && group_stride > 1)
                                dist_out[(size_t)q * dist_stride + col] = 0xffffffffu;       // no such row in the sample
                        }
                    }
                    TC_PROF_ADD(1);
                }
            }
            if (MODE == 2) {
                // row class = (row slice, lane quarter, lane / 4): the minima of four neighbouring lanes are
                // folded with two shuffle steps (once per item), leaving 32 classes per row slice; a class
                // that saw no live row reports TC_TILEMIN_NONE.  Lanes 0, 4, 8, ... write 64 consecutive
                // queries (256 B) each.
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const int jj = MODE == 2 ? j : 0;
                    float x = run_min[jj];
                    x = fminf(x, __shfl_xor_sync(0xffffffffu, x, 1));
                    x = fminf(x, __shfl_xor_sync(0xffffffffu, x, 2));
                    run_min[jj] = x;
                }
                const uint32_t rsl = item % n_rslices;
                const size_t cls = ((size_t)rsl * 4u + (uint32_t)(ew & 3)) * 8u + (uint32_t)(lane >> 2);
                int32_t* out = tilemin + cls * nq_pad + (size_t)qb0 * TC_NQ + half * 64;
                if ((lane & 3) == 0)
#pragma unroll
                for (int j4 = 0; j4 < 16; ++j4) {
                    int4 o;
                    const float a = run_min[(MODE == 2) ? 4 * j4 : 0], b = run_min[(MODE == 2) ? 4 * j4 + 1 : 0];
                    const float c = run_min[(MODE == 2) ? 4 * j4 + 2 : 0], d = run_min[(MODE == 2) ? 4 * j4 + 3 : 0];
                    o.x = a < 1e30f ? (int32_t)a : TC_TILEMIN_NONE; o.y = b < 1e30f ? (int32_t)b : TC_TILEMIN_NONE;
                    o.z = c < 1e30f ? (int32_t)c : TC_TILEMIN_NONE; o.w = d < 1e30f ? (int32_t)d : TC_TILEMIN_NONE;
                    reinterpret_cast<int4*>(out)[j4] = o;
                }
            }
        }
        if (MODE == 3 && plane_rows) flush_counts(plane_q);
        if (MODE == 0) {
            if (q_count) flush(q_count);
            if (lane == 0) {
                list_counts[blockIdx.x * TC_EPI_WARPS + ew] = min(n_written, rec_cap);
                if (n_written > rec_cap) *overflow = 1u;
            }
        }
        if (prof && blockIdx.x == 0 && threadIdx.x == 128)
            for (int i = 0; i < 4; ++i) prof[4 + i] = pw[i];
    } else {
        // ===================== query loader (TMA) + MMA issuer: one elected lane issues =====================
        uint32_t bfull_phase = 0, bfree_phase = 0;
        uint32_t it = 0;                          // running accumulator-block counter
        uint32_t p = 0;                           // running phase counter (as in the expanders)
        bool first_item = true;
        const long long t_begin = prof ? clock64() : 0;
        unsigned long long ns_begin = 0;
        if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_begin));
        const uint32_t smem_base = smem_u32(smem);
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
                    const uint32_t dst = smem_base + blk * QBLOCK_BYTES;
                    for (uint32_t off = 0; off < QBLOCK_BYTES; off += TC_STAGE_BYTES)
                        tma_bulk_g2s(dst + off, src + off, min((uint32_t)TC_STAGE_BYTES, QBLOCK_BYTES - off), b_full);
                }
            }
            __syncwarp();
            { TC_PROF_T0(); mbar_wait(b_full, bfull_phase); TC_PROF_ADD(3); }
            bfull_phase ^= 1u;
            for (uint32_t g = g_lo; g < g_hi; ++g, p += PH) {
                for (uint32_t blk = 0; blk < nb; ++blk, ++it) {
                    const uint32_t b = it % NBUF;
                    const bool last_blk = blk + 1 == nb;
                    { TC_PROF_T0(); mbar_wait(acc_empty(b), ((it / NBUF) & 1u) ^ 1u); TC_PROF_ADD(0); }
                    tc_fence_after();
                    const uint64_t bdesc0 = tc_smem_desc(smem_base + blk * QBLOCK_BYTES, 128, 1024);
                    const uint32_t d_addr = tmem_d + b * TC_NQ;
#pragma unroll
                    for (int ph = 0; ph < PH; ++ph) {
                        const uint32_t s = (p + ph) % NSLOT;
                        if (blk == 0) {
                            { TC_PROF_T0(); mbar_wait(a_ready(s), ((p + ph) / NSLOT) & 1u); TC_PROF_ADD(1); }
                            tc_fence_after();
                        }
                        if (elect_one()) {
                            const uint32_t a_addr = tmem_a + s * SLOT_COLS;
#pragma unroll
                            for (int kc = 0; kc < SC; ++kc) {           // one code chunk = 64 K bytes = two K=64 MMAs
                                const int ks = ph * SC + kc;
                                if (ks < NCHUNK) {
#pragma unroll
                                    for (int j = 0; j < 2; ++j)         // the address field counts 16-byte units
                                        tc_mma_mxf4_ts(d_addr, a_addr + (uint32_t)((kc * 2 + j) * 8),
                                                       bdesc0 + (uint64_t)(((ks / 2) * TC_STAGE_BYTES + ((ks % 2) * 4 + j * 2) * 128) >> 4),
                                                       IDESC, tmem_sf_one, tmem_sf_one, (ks | j) != 0 ? 1u : 0u);
                                }
                            }
                            // this slot of A may be rewritten once the MMAs issued so far retire
                            if (last_blk) tc_commit(a_free(s));
                            if (ph == PH - 1) {
                                // bias: D += 1.0(128 x 64, scales 16 | 1) * digits(128 queries x 64) = -v + 0.5
                                if (MODE != 2)
                                    tc_mma_mxf4_ts(d_addr, tmem_a + A_COLS,
                                                   tc_smem_desc(smem_base + blk * QBLOCK_BYTES + tc_subblocks(NCHUNK) * TC_STAGE_BYTES, 128, 256),
                                                   IDESC, tmem_sf_bias, tmem_sf_one, 1u);
                                tc_commit(acc_full(b));
                            }
                        }
                        __syncwarp();
                    }
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
        tc_dealloc(0u, TC_TMEM_COLS);
    }
}

// Warp-private survivor records (row, query) -> per-query candidate buffers, the contract of
// scan_kernel: key = hamming << 32 | row.  The distance is recomputed from the codes (xor + popc
// over nchunk*4 words, served from L2) — cheaper than carrying it through the epilogue.
// grid = (TC_SCATTER_X, number of lists = 8 x scan CTAs), each CTA strides over its list.
constexpr int TC_SCATTER_X = 2;
__global__ void __launch_bounds__(256)
tc_scatter_kernel(const uint2* __restrict__ recs, uint32_t rec_cap, const uint32_t* __restrict__ list_counts,
                  const uint4* __restrict__ codes, int nchunk, const uint32_t* __restrict__ qpack, int qs,
                  uint32_t* __restrict__ cnt, uint64_t* __restrict__ buf, uint32_t cap,
                  uint32_t* __restrict__ overflow) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t n_list = list_counts[blockIdx.y];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_list; i += gridDim.x * blockDim.x) {
        const uint2 r = recs[(size_t)blockIdx.y * rec_cap + i];
        const uint32_t row = r.x, q = r.y;
        const uint4* rc = codes + ((size_t)(row >> 5) * nchunk) * 32 + (row & 31);
        const uint4* qc = reinterpret_cast<const uint4*>(qpack + (size_t)q * qs);
        uint32_t d = 0;
        // six chunk pairs in flight (a loop over the runtime chunk count issued one L2 round trip after the other:
        // 88 % of this kernel's samples waited on them)
        for (int c0 = 0; c0 < nchunk; c0 += 6) {
            uint4 a[6], b[6];
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const bool in = c0 + u < nchunk;
                a[u] = in ? __ldg(rc + (c0 + u) * 32) : make_uint4(0, 0, 0, 0);
                b[u] = in ? __ldg(qc + c0 + u) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 6; ++u)
                d += __popc(a[u].x ^ b[u].x) + __popc(a[u].y ^ b[u].y) + __popc(a[u].z ^ b[u].z) + __popc(a[u].w ^ b[u].w);
        }
        const uint32_t pos = atomicAdd(&cnt[(size_t)q * CNT_STRIDE], 1u);
        if (pos < cap) buf[(size_t)q * cap + pos] = ((uint64_t)d << 32) | row;
        else *overflow = 1u;
    }
}

// ---- the FP4 MMA rate of this part, measured (the roofline denominator of the tensor-core scan) ------------------
// Back-to-back tcgen05.mma kind::mxf4 (M128 x N128 x K64, A in TMEM, B from shared memory, all scales 1.0)
// issued by one thread per SM, alternating between two accumulators; nothing else runs.  clk_out[0] = SM
// clocks CTA 0 needed for `niter` MMAs.  Operand contents do not matter.
__global__ void __launch_bounds__(128, 1)
tc_mma_rate_kernel(int niter, long long* __restrict__ clk_out) {
    extern __shared__ __align__(1024) uint8_t smem[];      // 4 KB B tile (whatever it holds)
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (warp == 0) tc_alloc(smem_u32(&s_tmem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    {
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = TC_SF_ONE;
        tc_st8(tmem + ((uint32_t)(warp * 32) << 16) + 384, v);        // scale words
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = TC_A_ONE8;
        tc_st8(tmem + ((uint32_t)(warp * 32) << 16) + 400, v);        // A: 64 x 1.0 per row
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        const uint64_t bdesc = tc_smem_desc(smem_u32(smem), 128, 256);
        const long long t0 = clock64();
        if (elect_one()) {
            for (int it = 0; it < niter; it += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    tc_mma_mxf4_ts(tmem + (u & 1) * 128, tmem + 400, bdesc, tc_idesc_mxf4(128, 128), tmem + 384, tmem + 388, 1u);
            }
            tc_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        if (lane == 0 && blockIdx.x == 0) clk_out[0] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem, 512);
}

}  // namespace gvdb
