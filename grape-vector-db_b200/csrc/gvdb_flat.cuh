// gvdb_flat.cuh — exact flat search, the FaissVectorIndex::search semantics
// (/root/reference/src/index.rs:620-640 with cosine_distance, :686-700).
//
// flat_scan_kernel is a register-tiled f32 "GEMM" whose inner product is NOT an FMA: every
// output element accumulates dot = dot + q[j]*c[j] for j ascending with separately rounded
// multiply and add, so it is bit-identical to the reference's sequential iterator sum.
// The epilogue forms distance = 1 - dot/(||q||*||c||) (+inf on a zero norm), maps it to an
// order-preserving u32 image and feeds the same threshold/append/select machinery as the
// Hamming scan (key = image << 32 | row), which yields the reference's stable ascending sort
// with ties broken by row number.
#pragma once
#include "gvdb_kernels.cuh"

namespace gvdb {

constexpr int FLAT_TM = 128;       // rows per CTA
constexpr int FLAT_TN = 64;        // queries per CTA
constexpr int FLAT_THREADS = 256;  // 16 x 16 threads, 8 rows x 4 queries each

// SIM = false: FaissVectorIndex::search — key image of distance = 1 - cos (+inf on a zero norm), ascending.
// SIM = true : BasicVectorStore::vector_search (/root/reference/src/storage.rs:296-339 with
//              cosine_similarity, :851-865) — similarity = cos (0.0 on a zero norm), rows with
//              similarity < sim_threshold skipped (when use_threshold), descending: key image = ~image(cos).
template <bool SIM>
__global__ void __launch_bounds__(FLAT_THREADS)
flat_scan_kernel(const float* __restrict__ rows, const float* __restrict__ norms,
                 const uint32_t* __restrict__ live, uint64_t row_lo, uint64_t row_hi, int dim,
                 const float* __restrict__ queries, const float* __restrict__ qnorm, uint32_t nq,
                 const uint32_t* __restrict__ tau, uint32_t* __restrict__ cnt,
                 uint64_t* __restrict__ buf, uint32_t cap, uint32_t* __restrict__ overflow,
                 float sim_threshold, int use_threshold) {
    __shared__ float As[FLAT_TM][33];
    __shared__ float Bs[FLAT_TN][33];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const uint64_t r0 = row_lo + (uint64_t)blockIdx.x * FLAT_TM;
    const uint32_t q0 = blockIdx.y * FLAT_TN;
    const bool vec4 = (dim & 3) == 0;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < dim; k0 += 32) {
        __syncthreads();
        if (vec4) {
            for (int idx = tid; idx < FLAT_TM * 8; idx += FLAT_THREADS) {
                int r = idx >> 3, seg = idx & 7, j = k0 + seg * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r0 + r < row_hi && j < dim)
                    v = *reinterpret_cast<const float4*>(rows + (r0 + r) * (uint64_t)dim + j);
                As[r][seg * 4 + 0] = v.x; As[r][seg * 4 + 1] = v.y;
                As[r][seg * 4 + 2] = v.z; As[r][seg * 4 + 3] = v.w;
            }
            for (int idx = tid; idx < FLAT_TN * 8; idx += FLAT_THREADS) {
                int r = idx >> 3, seg = idx & 7, j = k0 + seg * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (q0 + r < nq && j < dim)
                    v = *reinterpret_cast<const float4*>(queries + (uint64_t)(q0 + r) * dim + j);
                Bs[r][seg * 4 + 0] = v.x; Bs[r][seg * 4 + 1] = v.y;
                Bs[r][seg * 4 + 2] = v.z; Bs[r][seg * 4 + 3] = v.w;
            }
        } else {
            for (int idx = tid; idx < FLAT_TM * 32; idx += FLAT_THREADS) {
                int r = idx >> 5, e = idx & 31, j = k0 + e;
                As[r][e] = (r0 + r < row_hi && j < dim) ? rows[(r0 + r) * (uint64_t)dim + j] : 0.f;
            }
            for (int idx = tid; idx < FLAT_TN * 32; idx += FLAT_THREADS) {
                int r = idx >> 5, e = idx & 31, j = k0 + e;
                Bs[r][e] = (q0 + r < nq && j < dim) ? queries[(uint64_t)(q0 + r) * dim + j] : 0.f;
            }
        }
        __syncthreads();
        const int lim = min(32, dim - k0);
        for (int e = 0; e < lim; ++e) {
            float a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = As[i * 16 + ty][e];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[j * 16 + tx][e];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(b[j], a[i]));
        }
    }
    // epilogue: distance, image, threshold, append
    float qn[4]; uint32_t tq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t q = q0 + j * 16 + tx;
        qn[j] = q < nq ? qnorm[q] : 0.f;
        tq[j] = q < nq ? tau[q] : 0u;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint64_t row = r0 + i * 16 + ty;
        if (row >= row_hi) continue;
        if (!((live[row >> 5] >> (row & 31)) & 1u)) continue;
        const float rn = norms[row];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t q = q0 + j * 16 + tx;
            if (q >= nq) continue;
            uint32_t img;
            if (SIM) {
                const float c = (qn[j] == 0.0f || rn == 0.0f) ? 0.0f : __fdiv_rn(acc[i][j], __fmul_rn(qn[j], rn));
                if (use_threshold && c < sim_threshold) continue;
                img = ~f32_asc_key(c);
            } else {
                const float d = (qn[j] == 0.0f || rn == 0.0f)
                                    ? INFINITY
                                    : __fsub_rn(1.0f, __fdiv_rn(acc[i][j], __fmul_rn(qn[j], rn)));
                img = f32_asc_key(d);
            }
            if (img < tq[j]) {
                const uint32_t pos = atomicAdd(&cnt[(size_t)q * CNT_STRIDE], 1u);
                if (pos < cap) buf[(size_t)q * cap + pos] = ((uint64_t)img << 32) | (uint32_t)row;
                else *overflow = 1u;
            }
        }
    }
}

// sorted keys -> (global row, distance | similarity).  -0.0 comes back as +0.0 (the two tie, as in
// Rust's partial_cmp, so they share a key image).
__global__ void flat_emit_kernel(const uint64_t* __restrict__ buf, uint32_t cap,
                                 const uint32_t* __restrict__ cnt, uint32_t nq, uint32_t k,
                                 uint64_t row_base, uint64_t* __restrict__ ids_out,
                                 float* __restrict__ dist_out, int sim) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nq * k) return;
    uint32_t q = idx / k, t = idx % k;
    if (t < cnt[(size_t)q * CNT_STRIDE]) {
        uint64_t key = buf[(size_t)q * cap + t];
        uint32_t u = (uint32_t)(key >> 32);
        if (sim) u = ~u;
        uint32_t bits = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
        ids_out[idx] = row_base + (uint32_t)key;
        dist_out[idx] = __uint_as_float(bits);
    } else {
        ids_out[idx] = UINT64_MAX;
        dist_out[idx] = sim ? -INFINITY : INFINITY;
    }
}

}  // namespace gvdb
