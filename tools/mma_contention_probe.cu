// mma_contention_probe.cu — which concurrent activity of the OTHER warps of the CTA slows back-to-back
// tcgen05.mma kind::mxf4 (M128 x N128 x K64, A in TMEM, B from a 48 KB resident block)?  Warp 12 issues
// the MMAs exactly as gvdb_tc.cuh's MMA warp does; warps 0-11 run a background loop until it is done:
//   bg 0 idle (parked on a barrier)      bg 1 mbarrier try_wait spin        bg 2 FADD + funnel-shift chains
//   bg 3 tcgen05.ld x32 loops (accumulator columns)   bg 4 tcgen05.st x8 loops (A columns)
//   bg 5 LDS.128 broadcast loops          bg 6 = 2 + 3 + 5 (the scan's epilogue mix)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I grape-vector-db_b200/csrc
//        -o tools/bin/mma_contention_probe tools/mma_contention_probe.cu
#include <cstdio>
#include <cstdlib>
#include "gvdb_tc.cuh"
using namespace gvdb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr uint32_t IDESC = tc_idesc_mxf4(128, 128);

__global__ void __launch_bounds__(416, 1) probe(int bg, int nbg_warps, int nblocks, long long* clk_out, float* sink_out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ uint32_t s_tmem;
    __shared__ volatile uint32_t s_done;
    __shared__ float s_c[128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t done = smem_u32(&bars[0]), sink = smem_u32(&bars[1]), never = smem_u32(&bars[2]);
    if (threadIdx.x == 0) { mbar_init(done, 1); mbar_init(sink, 1); mbar_init(never, 1); fence_mbar_init(); s_done = 0; }
    if (threadIdx.x < 128) s_c[threadIdx.x] = 0.5f;
    if (warp == 12) tc_alloc(smem_u32(&s_tmem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t lane_taddr = (uint32_t)((warp & 3) * 32) << 16;
    if (warp < 4) {
        uint32_t v[8];
        for (int i = 0; i < 8; ++i) v[i] = TC_SF_ONE;
        tc_st8(lane_taddr + 192, v);
        for (int c = 0; c < 192; c += 8) { for (int i = 0; i < 8; ++i) v[i] = TC_A_ONE8; tc_st8(lane_taddr + c, v); }
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 12) {
        const uint32_t smem_base = smem_u32(smem);
        const uint64_t bdesc0 = tc_smem_desc(smem_base, 128, 1024);
        const long long t0 = clock64();
        for (int blk = 0; blk < nblocks; ++blk) {
            const uint32_t d = 208 + (blk & 1) * 128;
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {
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
                    tc_commit(sink);
                }
                __syncwarp();
            }
        }
        if (elect_one()) tc_commit(done);
        __syncwarp();
        mbar_wait(done, 0);
        const long long t1 = clock64();
        s_done = 1;
        if (lane == 0 && blockIdx.x == 0) clk_out[0] = t1 - t0;
    } else if (warp < nbg_warps) {
        float acc = 0.f;
        uint32_t m = 0;
        while (!s_done) {
            if (bg == 1) {
                for (int i = 0; i < 8; ++i) if (mbar_try_wait(never, 0)) acc += 1.f;
            }
            if (bg == 2 || bg == 6) {
                float x = acc + (float)lane;
#pragma unroll
                for (int i = 0; i < 64; ++i) { x = x + 1.25f; m = __funnelshift_l(__float_as_uint(x), m, 1); }
                acc = x;
            }
            if (bg == 3 || bg == 6) {
                uint32_t v0[32], v1[32];
                tc_ld32(lane_taddr + 208, v0);
                tc_ld32(lane_taddr + 240, v1);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) m ^= v0[i] ^ v1[i];
            }
            if (bg == 4) {
                uint32_t v[8];
                for (int i = 0; i < 8; ++i) v[i] = TC_A_ONE8;
                for (int c = 0; c < 48; c += 8) tc_st8(lane_taddr + 400 + c, v);
                tc_wait_st();
            }
            if (bg == 5 || bg == 6) {
                const float4* c4 = reinterpret_cast<const float4*>(s_c);
#pragma unroll
                for (int i = 0; i < 16; ++i) { const float4 c = c4[i]; acc += c.x + c.y + c.z + c.w; }
            }
        }
        if (acc == 12345.f || m == 0x12345u) sink_out[threadIdx.x] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tc_dealloc(s_tmem, 512);
}

int main() {
    long long* dclk; float* dsink; CK(cudaMalloc(&dclk, 8)); CK(cudaMalloc(&dsink, 4096));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
    const int nblocks = 400;
    for (int bg = 0; bg < 7; ++bg)
        for (int nw : {4, 12}) {
            for (int rep = 0; rep < 2; ++rep) probe<<<148, 416, 49152>>>(bg, nw, nblocks, dclk, dsink);
            CK(cudaDeviceSynchronize());
            long long clk; CK(cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost));
            printf("background %d on %2d warps: %.1f clk per MMA\n", bg, nw, (double)clk / (nblocks * 12));
        }
    printf("OK\n");
    return 0;
}
