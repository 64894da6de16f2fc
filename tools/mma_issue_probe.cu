// mma_issue_probe.cu — what costs issue slots around back-to-back tcgen05.mma kind::mxf4 (M128 x N128 x K64,
// A in TMEM)?  Variants of the MMA warp's loop of gvdb_tc.cuh, timed in isolation on every SM:
//   v0  12 MMAs per "block" with the real B descriptors (48 KB resident block, SBO 1024) and A columns, nothing else
//   v1  v0 + a tcgen05.commit after MMA 6 and MMA 12 (to barriers nobody waits on)
//   v2  v1 + an mbarrier try_wait on an already-completed barrier before MMA 1 and MMA 7
//   v3  v2 with the warp-collective structure of the kernel (elect_one per burst, __syncwarp)
//   v4  v0 but the whole loop issued by one thread with a wait every 12 MMAs
//   v5  v2 + tcgen05.fence::after_thread_sync after each wait (no __syncwarp);  v6  v2 + __syncwarp after each burst (no fence)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I grape-vector-db_b200/csrc
//        -o tools/bin/mma_issue_probe tools/mma_issue_probe.cu
#include <cstdio>
#include <cstdlib>
#include "gvdb_tc.cuh"
using namespace gvdb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr uint32_t IDESC = tc_idesc_mxf4(128, 128);

__global__ void __launch_bounds__(128, 1) probe(int variant, int nblocks, long long* clk_out) {
    extern __shared__ __align__(1024) uint8_t smem[];      // 48 KB "query block" (content irrelevant)
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t done = smem_u32(&bars[0]), sink = smem_u32(&bars[1]), ready = smem_u32(&bars[2]);
    if (threadIdx.x == 0) { mbar_init(done, 1); mbar_init(sink, 1); mbar_init(ready, 1); fence_mbar_init(); }
    if (warp == 0) tc_alloc(smem_u32(&s_tmem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) mbar_arrive(ready);               // phase 0 of `ready` is complete: waits on parity 0 pass
    {
        uint32_t v[8];
        for (int i = 0; i < 8; ++i) v[i] = TC_SF_ONE;
        tc_st8(((uint32_t)(warp * 32) << 16) + 192, v);
        for (int c = 0; c < 192; c += 8) { for (int i = 0; i < 8; ++i) v[i] = TC_A_ONE8; tc_st8(((uint32_t)(warp * 32) << 16) + c, v); }
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        const uint32_t smem_base = smem_u32(smem);
        const uint64_t bdesc0 = tc_smem_desc(smem_base, 128, 1024);
        const long long t0 = clock64();
        if (variant == 4) {
            if (elect_one()) {
                for (int blk = 0; blk < nblocks; ++blk) {
                    const uint32_t d = 208 + (blk & 1) * 128;
#pragma unroll
                    for (int ks = 0; ks < 6; ++ks)
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            tc_mma_mxf4_ts(d, (uint32_t)(((blk & 1) * 96) + (ks * 2 + j) * 8),
                                           bdesc0 + (uint64_t)(((ks / 2) * TC_STAGE_BYTES + ((ks % 2) * 4 + j * 2) * 128) >> 4),
                                           IDESC, 192, 192, (ks | j) != 0 ? 1u : 0u);
                    mbar_wait(ready, 0);
                }
            }
        } else {
            for (int blk = 0; blk < nblocks; ++blk) {
                const uint32_t d = 208 + (blk & 1) * 128;
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                    if (variant >= 2) { mbar_wait(ready, 0); if (variant == 3 || variant == 5) tc_fence_after(); }
                    if (elect_one()) {
#pragma unroll
                        for (int kc = 0; kc < 3; ++kc) {
                            const int ks = ph * 3 + kc;
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                tc_mma_mxf4_ts(d, (uint32_t)(((blk & 1) * 96) + ph * 48 + (kc * 2 + j) * 8),
                                               bdesc0 + (uint64_t)(((ks / 2) * TC_STAGE_BYTES + ((ks % 2) * 4 + j * 2) * 128) >> 4),
                                               IDESC, 192, 192, (ks | j) != 0 ? 1u : 0u);
                        }
                        if (variant >= 1) tc_commit(sink);
                    }
                    if (variant == 3 || variant == 6) __syncwarp();
                }
            }
        }
        __syncwarp();
        if (elect_one()) tc_commit(done);
        __syncwarp();
        mbar_wait(done, 0);
        const long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) clk_out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(s_tmem, 512);
}

int main() {
    long long* dclk; CK(cudaMalloc(&dclk, 8));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
    const int nblocks = 400;
    for (int v = 0; v < 7; ++v) {
        for (int rep = 0; rep < 2; ++rep) probe<<<148, 128, 49152>>>(v, nblocks, dclk);
        CK(cudaDeviceSynchronize());
        long long clk; CK(cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost));
        printf("variant %d: %.1f clk per MMA (%.0f per block of 12)\n", v, (double)clk / (nblocks * 12), (double)clk / nblocks);
    }
    printf("OK\n");
    return 0;
}
