// shard_kernels.cuh -- sort-only on N GPUs (SURVEY.md 8(e), row "depth + sort"): one sortIndexes call split by input position.
//
//   rank g owns positions [lo_g, hi_g) of the sort window indexes[sortStart..renderCount) (sorter.cpp:28); every rank holds all centres.
//   C1  min/max (sorter.cpp:24-25, 70-73) must be GLOBAL before the range map (sorter.cpp:142-146): each rank stores its pair into
//       every peer's block over NVLink and waits for all pairs -- 8 B per rank pair, no NCCL call, no host round trip.
//   local sort of the slice with the global range map; the final radix pass leaves, per key, the run [first, last) of the local order.
//   C2  each rank reads the per-key run lengths of all ranks straight from peer memory and derives where its runs start in the
//       global order: buckets descending; inside a bucket higher input positions first (sorter.cpp:162-167), i.e. rank G-1 ... 0.
//   place: local element j with key k goes to out[sortStart + j + delta_g[k]] in RANK 0's output buffer (P2P stores, 4 B/splat).
//
// The only data that crosses NVLink: 8 B min/max per rank pair, 8 B per key per rank pair of run bounds (reads), 4 B per splat of
// sorted indexes into rank 0.  All waits are bounded (kPeerTimeoutCycles) so a missing peer cannot hang a GPU.
#pragma once
#include "common.cuh"
#include "sort_kernels.cuh"
#include "raster_kernels.cuh"   // ld_acquire_sys_u32, kPeerTimeoutCycles

namespace gs {

constexpr int kMaxShardRanks = 8;

// Header of the block every rank exports (CUDA IPC); `runs` (R x uint2) follows it in the same allocation.
struct ShardHeader {
    int32_t mm[kMaxShardRanks][2];        // [g] = rank g's local (min, max) of the current sort, written BY rank g
    uint32_t mm_seq[kMaxShardRanks];      // [g] = sort number for which mm[g] is valid
    uint32_t runs_seq[kMaxShardRanks];    // [g] = sort number for which rank g's `runs` are complete (rank g writes it into every block)
    uint32_t done_seq[kMaxShardRanks];    // rank 0's block only: rank g's elements of that sort are in rank 0's output
    uint32_t timeout;                     // a bounded wait expired
    uint32_t pad[7];
};
static_assert(sizeof(ShardHeader) % 16 == 0, "runs must stay 8-byte aligned behind the header");

struct ShardPeers {
    ShardHeader *hdr[kMaxShardRanks];     // every rank's block as mapped into THIS process (own block included)
    const uint2 *runs[kMaxShardRanks];
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// lane g of one warp waits until flags[g] has reached `seq`
__device__ __forceinline__ bool wait_all_ranks(const uint32_t *flags, uint32_t world, uint32_t seq) {
    bool ok = true;
    if (threadIdx.x < world) {
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys_u32(flags + threadIdx.x) - seq) < 0) {
            if (clock64() - t0 > kPeerTimeoutCycles) { ok = false; break; }
            __nanosleep(100);
        }
    }
    return __all_sync(0xffffffffu, ok);
}

constexpr int kShardSyncThreads = 256;

// C1: my (min, max) into slot `me` of every rank's block; wait for all N pairs in mine; the global min/max replaces the local one in
// the control block (k_bucket reads it there).  Every peer having announced sort `seq` also means it has finished reading my run table
// of sort seq-1 (stream order on its side), so the table is reset here, by the whole CTA, for this sort.
__global__ void __launch_bounds__(kShardSyncThreads)
k_shard_exchange_minmax(ShardPeers peers, SortControl *ctl, uint32_t me, uint32_t world, uint32_t seq, int reset_error, uint2 *runs, uint32_t R) {
    ShardHeader *own = peers.hdr[me];
    if (threadIdx.x < 32) {
        if (threadIdx.x < world) {
            ShardHeader *h = peers.hdr[threadIdx.x];
            h->mm[me][0] = ctl->dmin;
            h->mm[me][1] = ctl->dmax;
            __threadfence_system();
            st_release_sys_u32(&h->mm_seq[me], seq);
        }
        const bool ok = wait_all_ranks(own->mm_seq, world, seq);
        int32_t lo = 2147483640, hi = -2147483640;   // sorter.cpp:24-25 seeds: an empty slice contributes nothing
        if (threadIdx.x < world) { lo = own->mm[threadIdx.x][0]; hi = own->mm[threadIdx.x][1]; }
        lo = warp_min(lo);
        hi = warp_max(hi);
        if (threadIdx.x == 0) {
            ctl->dmin = lo;
            ctl->dmax = hi;
            if (reset_error) ctl->error = 0;   // empty slice: k_depth, which clears it otherwise, did not run
            if (!ok) own->timeout = 1;
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < R; k += kShardSyncThreads) runs[k] = make_uint2(0xffffffffu, 0u);
}
// C2: my runs are complete (previous kernels of this stream) -> tell everybody, then wait for everybody's
__global__ void k_shard_exchange_runs(ShardPeers peers, uint32_t me, uint32_t world, uint32_t seq) {
    __threadfence_system();
    if (threadIdx.x < world) st_release_sys_u32(&peers.hdr[threadIdx.x]->runs_seq[me], seq);
    ShardHeader *own = peers.hdr[me];
    const bool ok = wait_all_ranks(own->runs_seq, world, seq);
    if (threadIdx.x == 0 && !ok) own->timeout = 1;
}

constexpr int kShardScanThreads = 1024;
__device__ __forceinline__ uint32_t run_length(uint2 r) { return r.y > r.x ? r.y - r.x : 0u; }

// C2, step 1: per key, the number of elements all ranks hold (`total`) and the number held by ranks that come BEFORE me inside the
// bucket (ranks > me: their input positions are higher).  Remote reads of 8 B per key per rank, coalesced.
__global__ void __launch_bounds__(kShardScanThreads)
k_shard_totals(ShardPeers peers, uint32_t me, uint32_t world, uint32_t R, uint32_t *__restrict__ total, uint32_t *__restrict__ ahead,
               uint32_t *__restrict__ block_total) {
    __shared__ uint32_t s_scan[40];
    const uint32_t k = blockIdx.x * kShardScanThreads + threadIdx.x;
    uint32_t t = 0, a = 0;
    if (k < R) {
        for (uint32_t g = 0; g < world; ++g) {
            const uint32_t c = run_length(peers.runs[g][k]);
            t += c;
            if (g > me) a += c;
        }
        total[k] = t;
        ahead[k] = a;
    }
    uint32_t sum;
    block_exclusive_scan<kShardScanThreads>(t, s_scan, sum);
    if (threadIdx.x == 0) block_total[blockIdx.x] = sum;
}
// C2, step 2: delta[k] = (elements of smaller keys, all ranks) + (same key, ranks ahead of me) - (start of my run in my local order)
__global__ void __launch_bounds__(kShardScanThreads)
k_shard_delta(const uint2 *__restrict__ my_runs, uint32_t R, const uint32_t *__restrict__ total, const uint32_t *__restrict__ ahead,
              const uint32_t *__restrict__ block_total, uint32_t *__restrict__ delta) {
    __shared__ uint32_t s_scan[40];
    __shared__ uint32_t s_prefix;
    uint32_t part = 0;
    for (uint32_t b = threadIdx.x; b < blockIdx.x; b += kShardScanThreads) part += block_total[b];
    uint32_t dummy;
    const uint32_t ex_part = block_exclusive_scan<kShardScanThreads>(part, s_scan, dummy);
    (void)ex_part;
    if (threadIdx.x == 0) s_prefix = dummy;
    __syncthreads();
    const uint32_t k = blockIdx.x * kShardScanThreads + threadIdx.x;
    const uint32_t t = k < R ? total[k] : 0u;
    uint32_t sum;
    const uint32_t ex = block_exclusive_scan<kShardScanThreads>(t, s_scan, sum);
    if (k < R) delta[k] = s_prefix + ex + ahead[k] - my_runs[k].x;   // wrapping: only meaningful where my run is not empty
}
// place: element j of my sorted slice -> its slot of the global order, in rank 0's buffer (peer stores; runs of equal keys are contiguous)
template <typename KeyT>
__global__ void __launch_bounds__(256)
k_shard_place(const KeyT *__restrict__ keys_sorted, const uint32_t *__restrict__ vals_sorted, uint32_t n, const uint32_t *__restrict__ delta,
              uint32_t *__restrict__ out /* rank 0's sorted + sortStart */) {
    for (uint32_t j = blockIdx.x * 256u + threadIdx.x; j < n; j += gridDim.x * 256u) out[j + delta[(uint32_t)keys_sorted[j]]] = vals_sorted[j];
}
// my elements are in rank 0's buffer; rank 0 then waits for everybody's
__global__ void k_shard_done(ShardPeers peers, uint32_t me, uint32_t world, uint32_t seq) {
    __threadfence_system();
    if (threadIdx.x == 0) st_release_sys_u32(&peers.hdr[0]->done_seq[me], seq);
    if (me == 0) {
        ShardHeader *own = peers.hdr[0];
        const bool ok = wait_all_ranks(own->done_seq, world, seq);
        if (threadIdx.x == 0 && !ok) own->timeout = 1;
        __threadfence_system();
    }
}

} // namespace gs
