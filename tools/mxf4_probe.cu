// mxf4_probe.cu — can the Hamming contraction run on the FP4 path (tcgen05.mma kind::mxf4,
// block-scaled, e2m1 operands, all scale factors 1.0)?  One M128 x N128 x K64 MMA with A in {0,1},
// B in {+1,-1}, checked against the CPU; then the issue rate.  A from shared memory (SS) and from
// TMEM (TS).  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
//   -I grape-vector-db_b200/csrc -o tools/bin/mxf4_probe tools/mxf4_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gvdb_tc.cuh"
using namespace gvdb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr uint32_t IDESC_MXF4 = (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_mxf4_ss(uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t sfa, uint32_t sfb, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %6, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%4], [%5], p;\n\t}"
                 :: "r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(sfa), "r"(sfb), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_mxf4_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t sfa, uint32_t sfb, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %6, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%4], [%5], p;\n\t}"
                 :: "r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(sfa), "r"(sfb), "r"(acc) : "memory");
}

// smem: A at 0 (128 rows x 32 B, core-matrix order: (row/8)*256 + (kb/16)*128 + (row%8)*16 + kb%16), B at 4096.
// mode 0: SS check, 1: TS check, 2: SS timing, 3: TS timing
__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* __restrict__ a_bytes, const uint8_t* __restrict__ b_bytes,
                                                        float* __restrict__ d_out, int mode, int niter, long long* clk_out,
                                                        uint32_t sfword_a = 0x7F7F7F7Fu, uint32_t sfword_b = 0x7F7F7F7Fu, uint32_t sf_id_bits = 0,
                                                        uint32_t nprobe = 128, uint32_t nacc = 2) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += 128) { smem[i] = a_bytes[i]; smem[4096 + i] = b_bytes[i]; }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (warp == 0) tc_alloc(smem_u32(&s_tmem), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t lane_taddr = (uint32_t)(warp * 32) << 16;
    {   // scale factors: every byte 0x7F (UE8M0 1.0) in columns [384, 392); A operand (TS) in columns [400, 408)
        uint32_t sf[8];
        for (int i = 0; i < 8; ++i) sf[i] = i < 4 ? sfword_a : sfword_b;
        tc_st8(tmem + lane_taddr + 384, sf);
        uint32_t av[8];
        const int row = threadIdx.x;
        for (int w = 0; w < 8; ++w) {          // 32 bytes of row `row`, K order
            uint32_t x = 0;
            for (int b = 0; b < 4; ++b) {
                const int kb = w * 4 + b;
                x |= (uint32_t)a_bytes[(row / 8) * 256 + (kb / 16) * 128 + (row % 8) * 16 + kb % 16] << (8 * b);
            }
            av[w] = x;
        }
        tc_st8(tmem + lane_taddr + 400, av);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        const uint64_t adesc = tc_smem_desc(smem_u32(smem), 128, 256);
        const uint64_t bdesc = tc_smem_desc(smem_u32(smem + 4096), 128, 256);
        const uint32_t sfa = tmem + 384, sfb = tmem + 388;
        long long t0 = clock64();
        if (mode < 2) {
            if (elect_one()) {
                if (mode & 1) mma_mxf4_ts(tmem, tmem + 400, bdesc, IDESC_MXF4 | sf_id_bits, sfa, sfb, 0u);
                else mma_mxf4_ss(tmem, adesc, bdesc, IDESC_MXF4 | sf_id_bits, sfa, sfb, 0u);
            }
            __syncwarp();
        } else if (mode == 4) {
            // issue rate for other N (timing only: B rows beyond the 4 KB that were filled read whatever follows in smem)
            const uint32_t idesc_n = (1u << 7) | (1u << 10) | ((nprobe >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
            if (elect_one()) {
                for (int it = 0; it < niter; it += 8) {
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        mma_mxf4_ts(tmem + (u % nacc) * nprobe, tmem + 400, bdesc, idesc_n, sfa, sfb, 1u);
                }
            }
        } else if (elect_one()) {
            for (int it = 0; it < niter; it += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (mode & 1) mma_mxf4_ts(tmem + (u & 1) * 128, tmem + 400, bdesc, IDESC_MXF4, sfa, sfb, 1u);
                    else mma_mxf4_ss(tmem + (u & 1) * 128, adesc, bdesc, IDESC_MXF4, sfa, sfb, 1u);
                }
            }
        }
        __syncwarp();
        if (elect_one()) tc_commit(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) clk_out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (mode < 2) {
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tc_ld32(tmem + lane_taddr + c0, v);
            tc_wait_ld();
            for (int j = 0; j < 32; ++j) d_out[(size_t)threadIdx.x * 128 + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem, 512);
}

int main() {
    // A[row][k] in {0,1}, B[n][k] in {+1,-1}; e2m1 nibbles: 0 -> 0x0, +1 -> 0x2, -1 -> 0xA; element 2i in the low nibble
    std::vector<int> A(128 * 64), B(128 * 64);
    srand(7);
    for (auto& x : A) x = rand() & 1;
    for (auto& x : B) x = (rand() & 1) ? 1 : -1;
    std::vector<uint8_t> ab(4096), bb(4096);
    auto off = [](int row, int kb) { return (row / 8) * 256 + (kb / 16) * 128 + (row % 8) * 16 + kb % 16; };
    for (int r = 0; r < 128; ++r)
        for (int kb = 0; kb < 32; ++kb) {
            auto nibA = [&](int k) { return A[r * 64 + k] ? 0x2 : 0x0; };
            auto nibB = [&](int k) { return B[r * 64 + k] > 0 ? 0x2 : 0xA; };
            ab[off(r, kb)] = (uint8_t)(nibA(2 * kb) | (nibA(2 * kb + 1) << 4));
            bb[off(r, kb)] = (uint8_t)(nibB(2 * kb) | (nibB(2 * kb + 1) << 4));
        }
    uint8_t *da, *db; float* dd; long long* dclk;
    CK(cudaMalloc(&da, 4096)); CK(cudaMalloc(&db, 4096)); CK(cudaMalloc(&dd, 128 * 128 * 4)); CK(cudaMalloc(&dclk, 8));
    CK(cudaMemcpy(da, ab.data(), 4096, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, bb.data(), 4096, cudaMemcpyHostToDevice));
    for (int mode = 0; mode < 2; ++mode) {
        CK(cudaMemset(dd, 0xff, 128 * 128 * 4));
        probe_kernel<<<1, 128, 8192>>>(da, db, dd, mode, 1, dclk);
        CK(cudaDeviceSynchronize());
        std::vector<float> D(128 * 128);
        CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 128; ++n) {
                int s = 0;
                for (int k = 0; k < 64; ++k) s += A[m * 64 + k] * B[n * 64 + k];
                if (D[m * 128 + n] != (float)s) { if (bad < 5) printf("  mode %d mismatch D[%d][%d] = %g want %d\n", mode, m, n, D[m * 128 + n], s); ++bad; }
            }
        printf("%s check: %d mismatches of 16384\n", mode ? "TS (A in TMEM)" : "SS (A in smem)", bad);
    }
    // which scale byte scales which half of K?  SFA/SFB words with ONE byte = 0x80 (2.0), TS mode
    for (int side = 0; side < 2; ++side)
        for (int byte = 0; byte < 4; ++byte) {
            const uint32_t w = 0x7F7F7F7Fu + (1u << (8 * byte));
            probe_kernel<<<1, 128, 8192>>>(da, db, dd, 1, 1, dclk, side == 0 ? w : 0x7F7F7F7Fu, side == 1 ? w : 0x7F7F7F7Fu, 0);
            CK(cudaDeviceSynchronize());
            std::vector<float> D(128 * 128);
            CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
            // classify: D == s0 + s1 (no effect), 2*s0 + s1 (first 32 K doubled), s0 + 2*s1, 2*(s0+s1)
            int cnt[5] = {0, 0, 0, 0, 0};
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 128; ++n) {
                    int s0 = 0, s1 = 0;
                    for (int k = 0; k < 32; ++k) s0 += A[m * 64 + k] * B[n * 64 + k];
                    for (int k = 32; k < 64; ++k) s1 += A[m * 64 + k] * B[n * 64 + k];
                    const float d = D[m * 128 + n];
                    if (d == (float)(2 * s0 + s1) && s0 != 0) cnt[1]++;
                    else if (d == (float)(s0 + 2 * s1) && s1 != 0) cnt[2]++;
                    else if (d == (float)(2 * s0 + 2 * s1) && (s0 + s1) != 0) cnt[3]++;
                    else if (d == (float)(s0 + s1)) cnt[0]++;
                    else cnt[4]++;
                }
            printf("SF%c byte %d = 2.0: none %d, K[0,32) x2 %d, K[32,64) x2 %d, both %d, other %d\n", side ? 'B' : 'A', byte, cnt[0], cnt[1], cnt[2], cnt[3], cnt[4]);
        }
    for (int mode = 2; mode < 4; ++mode) {
        const int niter = 4000;
        for (int rep = 0; rep < 2; ++rep) probe_kernel<<<148, 128, 8192>>>(da, db, dd, mode, niter, dclk);
        CK(cudaDeviceSynchronize());
        long long clk; CK(cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost));
        printf("%s rate: %.1f clk per M128 x N128 x K64 MMA = %.0f MAC/clk/SM\n", (mode & 1) ? "TS" : "SS", (double)clk / niter,
               128.0 * 128 * 64 * niter / clk);
    }
    // issue rate against N (accumulators of N columns; A, scales at columns >= 384, so N * nacc <= 384)
    for (uint32_t np : {64u, 128u, 192u, 256u}) {
        for (uint32_t nacc : {1u, 2u}) {
            if (np * nacc > 384) continue;
            const int niter = 4000;
            for (int rep = 0; rep < 2; ++rep)
                probe_kernel<<<148, 128, 16384>>>(da, db, dd, 4, niter, dclk, 0x7F7F7F7Fu, 0x7F7F7F7Fu, 0, np, nacc);
            CK(cudaDeviceSynchronize());
            long long clk; CK(cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost));
            printf("TS rate N=%u, %u accumulator(s): %.1f clk per M128 x N%u x K64 MMA = %.0f MAC/clk/SM\n", np, nacc, (double)clk / niter, np,
                   128.0 * np * 64 * niter / clk);
        }
    }
    // the same loop timed with CUDA events over ~10 ms on every SM: the FP4 MMA rate in TIME units, i.e. at the
    // SM clock the part actually sustains under this load (the roofline denominator bench.py quotes)
    {
        const int niter = 40000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        probe_kernel<<<148, 128, 16384>>>(da, db, dd, 4, niter, dclk, 0x7F7F7F7Fu, 0x7F7F7F7Fu, 0, 128, 2);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int rep = 0; rep < 4; ++rep) probe_kernel<<<148, 128, 16384>>>(da, db, dd, 4, niter, dclk, 0x7F7F7F7Fu, 0x7F7F7F7Fu, 0, 128, 2);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        long long clk; CK(cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost));
        const double macs = 4.0 * 148 * (double)niter * 128 * 128 * 64;
        printf("sustained: %.1f clk per MMA, %.2f ms for 4 launches -> %.1f TMAC/s = %.2f PFLOP/s dense FP4 on 148 SMs (implied SM clock %.0f MHz)\n",
               (double)clk / niter, ms, macs / (ms * 1e-3) / 1e12, 2 * macs / (ms * 1e-3) / 1e15,
               4.0 * clk / (ms * 1e-3) / 1e6);
    }
    printf("OK\n");
    return 0;
}
