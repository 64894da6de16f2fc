// mma_floor.cu — issue-rate floor of tcgen05.mma for the shapes/layouts the Hamming scan can use.
// One CTA per SM, one thread issues NITER MMAs back to back (operands are whatever is in
// shared memory / TMEM: only the timing matters), commit, wait.  Prints clk per MMA.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I grape-vector-db_b200/csrc
//        -o tools/bin/mma_floor tools/mma_floor.cu
#include <cstdio>
#include <cstdlib>
#include "gvdb_tc.cuh"
using namespace gvdb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

enum { KIND_I8 = 0, KIND_F8 = 1, KIND_F16 = 2 };

__device__ __forceinline__ void tc_mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int KIND, bool A_TMEM>
__device__ __forceinline__ void mma_any(uint32_t d, uint32_t a_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if (A_TMEM) {
        if (KIND == KIND_I8)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                         :: "r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
        else if (KIND == KIND_F8)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
                         :: "r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         :: "r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    } else {
        if (KIND == KIND_I8)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                         :: "r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
        else if (KIND == KIND_F8)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                         :: "r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         :: "r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    }
}

__host__ __device__ constexpr uint32_t idesc_for(int kind, int M, int N) {
    uint32_t base = ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (kind == KIND_I8) return base | (2u << 4) | (1u << 7) | (1u << 10);
    if (kind == KIND_F8) return base | (1u << 4);
    return base | (1u << 4) | (1u << 7) | (1u << 10);   // f16 kind: A=B=BF16, D=F32
}

// swz: 0 = no swizzle (LBO 128, SBO 1024: 128 K-bytes per row block), 1 = SWIZZLE_128B (SBO 1024)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int swz) {
    if (swz == 0) return tc_smem_desc(addr, 128, 1024);
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int KIND, bool A_TMEM, int N, int SWZ>
__global__ void __launch_bounds__(128, 1) floor_kernel(int niter, long long* out_clk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 1);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (warp == 0) tc_alloc(smem_u32(&s_tmem), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32 * 1024);
        constexpr uint32_t IDESC = idesc_for(KIND, 128, N);
        long long t0 = clock64();
        for (int it = 0; it < niter; ++it) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {     // 4 K-steps of 32 bytes inside a 128-byte K block
                const uint32_t koff = SWZ ? j * 32 : j * 256;
                mma_any<KIND, A_TMEM>(tmem + 256, tmem + j * 8, make_desc(sa + koff, SWZ), make_desc(sb + koff, SWZ), IDESC, (it | j) ? 1u : 0u);
            }
        }
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out_clk[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem, 512);
}


// ---- interference test: the scan's MMA stream (i8, A in TMEM, N=128, no swizzle) with other
// activity in the CTA.  flags: 1 = commit every 12 MMAs, 2 = four warps spin on an mbarrier,
// 4 = four warps stream tcgen05.st into A columns, 8 = four warps stream tcgen05.ld from D columns,
// 16 = B descriptors walk over 192 KB instead of 16 KB.
__global__ void __launch_bounds__(288, 1) interfere_kernel(int niter, int flags, long long* out_clk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, bar_dummy, bar_done;
    __shared__ uint32_t s_tmem;
    __shared__ volatile int s_stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 1);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar_dummy), 1); mbar_init(smem_u32(&bar_done), 1); s_stop = 0; fence_mbar_init(); }
    if (warp == 8) tc_alloc(smem_u32(&s_tmem), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t lane_taddr = (uint32_t)((warp & 3) * 32) << 16;
    if (warp == 8 && lane == 0) {
        const uint32_t sb = smem_u32(smem);
        constexpr uint32_t IDESC = idesc_for(KIND_I8, 128, 128);
        long long t0 = clock64();
        int n = 0;
        const int period = (flags >> 8) & 0xff;
        for (int it = 0; it < niter; ++it) {
            const uint32_t base = (flags & 16) ? sb + (uint32_t)(it % 12) * 16384u : sb;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                tc_mma_i8_ts(tmem + 256 + ((flags & 32) ? 0 : (it & 1) * 128), tmem + ((flags & 64) ? j * 8 : ((it % 6) * 4 + j) * 8), tc_smem_desc(base + j * 256, 128, 1024), IDESC, (it | j) ? 1u : 0u);
                if (period && (++n % period) == 0) tc_commit(smem_u32(&bar_dummy));
            }
        }
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out_clk[0] = t1 - t0;
        s_stop = 1;
        mbar_arrive(smem_u32(&bar_done));
    } else if (warp < 4) {
        if (flags & 2) { mbar_wait(smem_u32(&bar_done), 0); }
        else if (flags & 4) {
            uint32_t v[8];
            for (int i = 0; i < 8; ++i) v[i] = 0x80808080u & (lane * 0x01010101u + i);
            int c = 0;
            while (!s_stop) {
                tc_st8(tmem + lane_taddr + 200 + (c & 3) * 8, v);   // columns 200..231: unused by the MMAs
                if ((++c & 7) == 0) tc_wait_st();
            }
            tc_wait_st();
        }
    } else if (warp < 8) {
        if (flags & 2) { mbar_wait(smem_u32(&bar_done), 0); }
        else if (flags & 8) {
            uint32_t acc = 0;
            while (!s_stop) {
                uint32_t v[32];
                tc_ld32(tmem + lane_taddr + 384, v);               // columns 384..415: unused by the MMAs
                tc_wait_ld();
                acc += v[0] ^ v[31];
            }
            if (acc == 0x12345678u) out_clk[1] = acc;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tc_dealloc(tmem, 512);
}

void run_interfere(const char* name, int flags) {
    long long* d; CK(cudaMalloc(&d, 16));
    const int niter = 3000;
    CK(cudaFuncSetAttribute(interfere_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int rep = 0; rep < 2; ++rep) interfere_kernel<<<148, 288, 200 * 1024>>>(niter, flags, d);
    CK(cudaDeviceSynchronize());
    long long clk; CK(cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost));
    printf("%-60s %7.1f clk/MMA\n", name, (double)clk / (niter * 4.0));
    cudaFree(d);
}

template <int KIND, bool A_TMEM, int N, int SWZ>
void run(const char* name, int grid) {
    long long* d; CK(cudaMalloc(&d, 8));
    const int niter = 2000;
    CK(cudaFuncSetAttribute(floor_kernel<KIND, A_TMEM, N, SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int rep = 0; rep < 2; ++rep) floor_kernel<KIND, A_TMEM, N, SWZ><<<grid, 128, 64 * 1024>>>(niter, d);
    CK(cudaDeviceSynchronize());
    long long clk; CK(cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost));
    const double per = (double)clk / (niter * 4.0);
    const int kbytes = 32;
    const double mac = 128.0 * N * (KIND == KIND_F16 ? 16 : kbytes) / per;
    printf("%-44s grid=%3d  %7.1f clk/MMA   %7.0f MAC/clk/SM\n", name, grid, per, mac);
    cudaFree(d);
}

int main() {
    for (int grid : {148}) {
        run<KIND_I8, true, 128, 0>("i8  A=tmem N=128 B no-swizzle", grid);
        run<KIND_I8, true, 128, 1>("i8  A=tmem N=128 B swizzle128", grid);
        run<KIND_I8, true, 256, 1>("i8  A=tmem N=256 B swizzle128", grid);
        run<KIND_I8, false, 128, 1>("i8  A=smem N=128 swizzle128", grid);
        run<KIND_I8, false, 256, 1>("i8  A=smem N=256 swizzle128", grid);
        run<KIND_F8, true, 128, 0>("f8  A=tmem N=128 B no-swizzle", grid);
        run<KIND_F8, true, 128, 1>("f8  A=tmem N=128 B swizzle128", grid);
        run<KIND_F8, true, 256, 1>("f8  A=tmem N=256 B swizzle128", grid);
        run<KIND_F8, false, 128, 1>("f8  A=smem N=128 swizzle128", grid);
        run<KIND_F8, false, 256, 1>("f8  A=smem N=256 swizzle128", grid);
        run<KIND_F16, true, 128, 1>("bf16 A=tmem N=128 B swizzle128 (K=16)", grid);
        run<KIND_F16, false, 256, 1>("bf16 A=smem N=256 swizzle128 (K=16)", grid);
    }
    run_interfere("i8 TS N=128: alone (D alternates per 4 MMAs, A cycles 24 slices)", 0);
    run_interfere("  fixed D buffer", 32);
    run_interfere("  fixed A slices", 64);
    run_interfere("  fixed D + fixed A", 96);
    run_interfere("  commit every 4 MMAs", 4 << 8);
    run_interfere("  commit every 8 MMAs", 8 << 8);
    run_interfere("  commit every 12 MMAs", 12 << 8);
    run_interfere("  commit every 25 MMAs", 25 << 8);
    run_interfere("  commit every 50 MMAs", 50 << 8);
    run_interfere("  commit every 100 MMAs", 100 << 8);
    run_interfere("  commit every 12 MMAs, fixed D + fixed A", (12 << 8) | 96);
    run_interfere("  4 warps streaming tcgen05.st", 4);
    run_interfere("  4 warps streaming tcgen05.ld", 8);
    printf("OK\n");
    return 0;
}
