// gvdb_bigr.cuh — stage 1 + stage 2 for LARGE rescore counts (R > 2048), the reference's default
// "ratio mode": rescore_count = (N as f32 * rescore_ratio) as usize, e.g. 100 000 of 1 M rows
// (/root/reference/src/quantization.rs:178-179).  The block-bitonic path keeps at most 2048
// candidates per query on chip; here the cut is made by counting instead, one query at a time:
//   dist_hist_kernel     histogram of the query's Hamming distances over the live rows
//   cut_kernel           threshold bin b* holding the R-th smallest key; M = #rows with ham <= b*
//   cut_compact_kernel   keys (ham << 32 | row) of those M rows, unordered append
//   radix sort (CUB)     M keys ascending == the reference's stable sort by similarity desc
//                        (/root/reference/src/quantization.rs:175); the first R are the candidates
//   rescore_slab_kernel  exact sequential-fold cosine of the R candidates (gvdb_kernels.cuh)
//   cos_key_kernel + stable radix sort (CUB) by descending cosine: equal cosines keep stage-1
//                        order, i.e. the reference's second stable sort (:190)
//   bigr_emit_kernel     first k of that order
// The sorts are cub::DeviceRadixSort (CUDA toolkit); everything else is hand-written.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "gvdb_kernels.cuh"

namespace gvdb {

struct BigRCut {
    uint32_t bstar;     // threshold bin
    uint32_t m;         // rows with ham <= bstar (>= r_eff)
    uint32_t r_eff;     // min(R, live rows)
    uint32_t appended;  // compaction cursor
};

// hist[b] = number of LIVE rows at distance b.  dist: N u32 (gvdb_hamming layout), one query.
__global__ void __launch_bounds__(256)
dist_hist_kernel(const uint32_t* __restrict__ dist, const uint32_t* __restrict__ live, uint64_t n,
                 uint32_t nbins, uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t sh[];
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if ((live[i >> 5] >> (i & 31)) & 1u) atomicAdd(&sh[min(dist[i], nbins - 1)], 1u);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// One warp: r_eff = min(R, live), b* = first bin whose cumulative count reaches r_eff, m = that count.
__global__ void cut_kernel(const uint32_t* __restrict__ hist, uint32_t nbins, uint32_t R, BigRCut* __restrict__ out) {
    const uint32_t lane = threadIdx.x;
    const uint32_t per = (nbins + 31) / 32;
    const uint32_t b0 = lane * per, b1 = min(nbins, b0 + per);
    uint32_t sum = 0;
    for (uint32_t b = b0; b < b1; ++b) sum += hist[b];
    uint32_t incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t r_eff = min(R, total);
    uint32_t run = incl - sum;
    if (r_eff == 0) {
        if (lane == 0) *out = BigRCut{0u, 0u, 0u, 0u};
        return;
    }
    if (run < r_eff && incl >= r_eff) {
        for (uint32_t b = b0; b < b1; ++b) {
            run += hist[b];
            if (run >= r_eff) { *out = BigRCut{b, run, r_eff, 0u}; break; }
        }
    }
}

// keys_out[cursor++] = ham << 32 | row for every live row with ham <= b* (warp-aggregated append).
__global__ void __launch_bounds__(256)
cut_compact_kernel(const uint32_t* __restrict__ dist, const uint32_t* __restrict__ live, uint64_t n,
                   BigRCut* __restrict__ cut, uint64_t* __restrict__ keys_out) {
    const uint32_t bstar = cut->bstar;
    const bool any = cut->r_eff != 0;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_round = (n + 31) / 32 * 32;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (uint64_t)gridDim.x * blockDim.x) {
        bool hit = false;
        uint32_t d = 0;
        if (any && i < n && ((live[i >> 5] >> (i & 31)) & 1u)) { d = dist[i]; hit = d <= bstar; }
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            uint32_t base = 0;
            const int leader = __ffs(m) - 1;
            if ((int)lane == leader) base = atomicAdd(&cut->appended, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (hit) keys_out[base + __popc(m & ((1u << lane) - 1u))] = ((uint64_t)d << 32) | (uint32_t)i;
        }
    }
}

// sort keys of the final order: descending cosine image; value = position in stage-1 order
__global__ void cos_key_kernel(const float* __restrict__ score, uint32_t r, uint32_t* __restrict__ key,
                               uint32_t* __restrict__ val) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r) return;
    key[i] = ~f32_asc_key(score[i]);
    val[i] = i;
}

// first k of the final order; slots beyond r_eff are unfilled (GVDB_NO_ID, -inf)
__global__ void bigr_emit_kernel(const uint32_t* __restrict__ perm, const BigRCut* __restrict__ cut,
                                 const uint64_t* __restrict__ rec_ids, const float* __restrict__ rec_score,
                                 uint32_t k, uint64_t* __restrict__ ids_out, float* __restrict__ scores_out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    if (t < cut->r_eff) {
        const uint32_t p = perm[t];
        ids_out[t] = rec_ids[p];
        scores_out[t] = rec_score[p];
    } else {
        ids_out[t] = UINT64_MAX;
        scores_out[t] = -INFINITY;
    }
}

// records beyond r_eff in a query's [R] record arrays
__global__ void bigr_fill_tail_kernel(const BigRCut* __restrict__ cut, uint32_t R, uint32_t* __restrict__ rec_ham,
                                      uint64_t* __restrict__ rec_ids, float* __restrict__ rec_score) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R || i < cut->r_eff) return;
    rec_ham[i] = 0xffffffffu;
    rec_ids[i] = UINT64_MAX;
    rec_score[i] = -INFINITY;
}

// ---- ratio mode across row shards (SURVEY §8e "Collective (ratio mode)") --------------------------------------------
// Every shard histograms its live rows' distances per query; the shards' histograms are gathered (one collective);
// from them every shard derives the SAME global cut — the bin b* holding the R-th smallest key (hamming, global row)
// of the union — and its own share of it: all its rows below b*, and of the ties AT b* as many as are left after the
// lower-ranked shards took theirs (ties are ordered by global row = by shard, then by local row).
// hists: [n_shards][nq][nbins].  One warp per query; out[q] is the cut of shard `my`: bstar, m = its rows with
// ham <= b*, r_eff = its members of the global top R (the first r_eff of those m in (hamming, row) order).
__global__ void __launch_bounds__(32)
shard_cut_kernel(const uint32_t* __restrict__ hists, uint32_t n_shards, uint32_t my, uint32_t nq, uint32_t nbins,
                 unsigned long long R, BigRCut* __restrict__ out) {
    const uint32_t q = blockIdx.x, lane = threadIdx.x;
    auto g = [&](uint32_t b) {
        unsigned long long t = 0;
        for (uint32_t s = 0; s < n_shards; ++s) t += hists[((size_t)s * nq + q) * nbins + b];
        return t;
    };
    const uint32_t per = (nbins + 31) / 32;
    const uint32_t b0 = min(nbins, lane * per), b1 = min(nbins, b0 + per);
    unsigned long long sum = 0;
    for (uint32_t b = b0; b < b1; ++b) sum += g(b);
    unsigned long long incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t;
    }
    const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned long long r_eff = min(R, total);
    if (r_eff == 0) {
        if (lane == 0) out[q] = BigRCut{0u, 0u, 0u, 0u};
        return;
    }
    unsigned long long run = incl - sum;
    if (run < r_eff && incl >= r_eff) {
        for (uint32_t b = b0; b < b1; ++b) {
            const unsigned long long gb = g(b);
            if (run + gb >= r_eff) {
                const unsigned long long need = r_eff - run;                    // ties wanted at b* (>= 1)
                unsigned long long before = 0;
                for (uint32_t s = 0; s < my; ++s) before += hists[((size_t)s * nq + q) * nbins + b];
                const uint32_t* mine = hists + ((size_t)my * nq + q) * nbins;
                const unsigned long long ties = mine[b];
                const unsigned long long keep = need > before ? min(ties, need - before) : 0ull;
                unsigned long long below = 0;
                for (uint32_t bb = 0; bb < b; ++bb) below += mine[bb];
                out[q] = BigRCut{b, (uint32_t)(below + ties), (uint32_t)(below + keep), 0u};
                break;
            }
            run += gb;
        }
    }
}

// The first kr entries of the cosine order (perm: stage-1 positions) of a query, listed again in stage-1 order
// ((hamming, global row) ascending): the shard's record list of a ratio-mode search.  One CTA; kr <= 1024.
__global__ void __launch_bounds__(256)
bigr_emit_records_kernel(const uint32_t* __restrict__ perm, const BigRCut* __restrict__ cut,
                         const uint32_t* __restrict__ rec_ham, const uint64_t* __restrict__ rec_ids,
                         const float* __restrict__ rec_score, uint32_t kr, uint64_t* __restrict__ out_ids,
                         uint32_t* __restrict__ out_ham, float* __restrict__ out_score) {
    extern __shared__ __align__(16) uint64_t sk[];
    const uint32_t take = min(kr, cut->r_eff);
    const uint32_t n_eff = max(64u, next_pow2(kr));
    for (uint32_t i = threadIdx.x; i < n_eff; i += blockDim.x) sk[i] = i < take ? (uint64_t)perm[i] : UINT64_MAX;
    __syncthreads();
    bitonic_sort_smem(sk, n_eff);
    for (uint32_t t = threadIdx.x; t < kr; t += blockDim.x) {
        if (t < take) {
            const uint32_t p = (uint32_t)sk[t];
            out_ids[t] = rec_ids[p]; out_ham[t] = rec_ham[p]; out_score[t] = rec_score[p];
        } else {
            out_ids[t] = UINT64_MAX; out_ham[t] = 0xffffffffu; out_score[t] = -INFINITY;
        }
    }
}

// predicate of the owner-side compaction (gvdb_rescore_keys_device): pair p is scored here iff
// its key is filled and its row lies in this index's resident window
struct OwnedPair {
    const uint64_t* keys;
    uint64_t row_base, lo, hi;
    __host__ __device__ bool operator()(uint32_t p) const {
        const uint64_t key = keys[p];
        if (key == UINT64_MAX) return false;
        const uint64_t local = (key & ((1ull << 40) - 1)) - row_base;
        return local >= lo && local < hi;
    }
};

// list of the pair indices this GPU scores (one launch; *count zeroed by the caller).  A warp appends
// its owned pairs as one ascending run, the runs land in arrival order: every pair is scored on its
// own, so the order of the list does not change any result.
__global__ void __launch_bounds__(256)
owned_compact_kernel(OwnedPair pred, uint32_t n_pairs, uint32_t* __restrict__ list, uint32_t* __restrict__ count) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool own = p < n_pairs && pred(p);
    const uint32_t mask = __ballot_sync(0xffffffffu, own);
    if (mask == 0) return;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(count, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (own) list[base + __popc(mask & ((1u << lane) - 1u))] = p;
}

}  // namespace gvdb
