// gvdb_xchg.cuh — peer exchange over NVLink/NVSwitch peer memory: the cross-GPU step of the
// "codes replicated, rows sharded, queries partitioned" layout without a collective library in the
// data path.  Every GPU owns a MAILBOX in its HBM that all its peers map (CUDA IPC):
//
//   set s in {0,1} (steps alternate, so a fast peer can write step e+1 while I still read step e):
//     q_in   [W][nq][dim] f32   slot w = rank w's query batch
//     keys_in[W][nq][R]   u64   slot w = rank w's stage-1 candidate keys (hamming << 32 | global row)
//     sc_in  [W][nq][R]   f32   slot o = owner o's cosines for MY queries' candidates it owns
//   flags  [3][W] u32, one per 128-byte line: epoch counters, flag[kind][w] = last step whose
//          slot-w data of that kind is complete
//
// A step (gvdb_search_exchange_device) on rank r:
//   push my queries into every peer's q_in[r]  + signal(Q)        side stream, under stage 1
//   stage 1 on my queries (scan, cut)                             → my keys
//   push my keys into every peer's keys_in[r]   + signal(K)
//   wait(Q, K from every peer) → score the candidates whose rows I own (all W batches)
//   scatter chunk w of the scores into peer w's sc_in[r] + signal(S)
//   wait(S from every peer) → each key takes its owner's score → order → top k
// Data moves as posted peer STORES (the cheap direction over NVLink); a signal is one
// st.release.sys per peer issued by the LAST block of the kernel that produced the data (every block
// fences at system scope and counts itself done), a wait is an ld.acquire.sys spin with a time limit
// in a one-block kernel of its own (a missing peer becomes an error, not a hang; and a spinning
// kernel of one block never keeps another rank's kernels off the GPU when ranks share one).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gvdb {

enum XchgKind : uint32_t { XCHG_Q = 0, XCHG_K = 1, XCHG_S = 2, XCHG_KINDS = 3 };
constexpr uint32_t XCHG_FLAG_STRIDE = 32;        // u32 per flag: one 128-byte line each
constexpr uint32_t XCHG_MAX_WORLD = 32;

// Completion signal folded into the producing kernel: every block fences its stores at system scope and
// counts itself done; the block that counts last publishes flag[kind][rank] := epoch in every peer's
// mailbox (st.release.sys).  One launch instead of a push and a signal kernel.
struct XchgSignal {
    uint8_t* const* peers;
    uint64_t flags_off;
    uint32_t kind, world, rank, epoch;
    uint32_t* done;          // device counter of finished blocks (zero before the launch, reset by the last block)
};
__device__ __forceinline__ void xchg_block_done(const XchgSignal& sg, uint32_t total_blocks) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(sg.done, 1u);
        if (prev + 1 == total_blocks) {
            __threadfence_system();
            for (uint32_t w = 0; w < sg.world; ++w) {
                uint32_t* f = reinterpret_cast<uint32_t*>(sg.peers[w] + sg.flags_off) + (size_t)(sg.kind * sg.world + sg.rank) * XCHG_FLAG_STRIDE;
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(sg.epoch) : "memory");
            }
            *sg.done = 0u;
        }
    }
}

// grid (x, W): block column w copies bytes from src + w * src_stride to peers[w] + dst_off.
// src_stride = 0: the same buffer to every peer (push); = bytes: chunk w to peer w (scatter).
// The last block to finish signals `sg.kind` to every peer.
template <typename T>
__global__ void __launch_bounds__(256)
xchg_push_kernel(uint8_t* const* __restrict__ peers, uint64_t dst_off, const uint8_t* __restrict__ src,
                 uint64_t src_stride, uint64_t bytes, XchgSignal sg) {
    const uint32_t w = blockIdx.y;
    const T* s = reinterpret_cast<const T*>(src + (size_t)w * src_stride);
    T* d = reinterpret_cast<T*>(peers[w] + dst_off);
    const uint64_t n = bytes / sizeof(T);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        d[i] = s[i];
    xchg_block_done(sg, gridDim.x * gridDim.y);
}

// one thread per peer: flag[kind][rank] of peer w := epoch (release at system scope; the pushes
// were issued by an earlier kernel of the same stream, so they are performed before this store)
__global__ void xchg_signal_kernel(uint8_t* const* __restrict__ peers, uint64_t flags_off, uint32_t kind,
                                   uint32_t world, uint32_t rank, uint32_t epoch) {
    const uint32_t w = threadIdx.x;
    if (w >= world) return;
    uint32_t* f = reinterpret_cast<uint32_t*>(peers[w] + flags_off) + (size_t)(kind * world + rank) * XCHG_FLAG_STRIDE;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
}

// thread (kind, w): spin until my flag[kind][w] has reached `epoch`; after limit_ns the thread
// gives up and records which kind it was waiting for in *err (device memory)
__global__ void xchg_wait_kernel(const uint32_t* __restrict__ flags, uint32_t kind_mask, uint32_t world, uint32_t epoch,
                                 uint64_t limit_ns, uint32_t* __restrict__ err) {
    const uint32_t kind = threadIdx.x / XCHG_MAX_WORLD, w = threadIdx.x % XCHG_MAX_WORLD;
    if (w >= world || !((kind_mask >> kind) & 1u)) return;
    const uint32_t* f = flags + (size_t)(kind * world + w) * XCHG_FLAG_STRIDE;
    uint64_t t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        uint64_t t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > limit_ns) {
            atomicOr(err, 1u << kind);
            break;
        }
        __nanosleep(64);
    }
}

}  // namespace gvdb
