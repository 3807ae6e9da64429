// sort_kernels.cuh -- depth + bucket + one-sweep radix kernels (sm_100a).
//
// Replaces the reference's sortIndexes() (src/worker/sorter.cpp:17-168) stage by stage:
//   k_depth      : distance pass, all six branches            sorter.cpp:29-140   (+ running min/max :24-25)
//   k_bucket     : f32 range map + bucket                      sorter.cpp:142-149  (+ per-digit histograms)
//   k_radix_pass : stable LSD one-sweep scatter, 8-bit digits  sorter.cpp:151-167  (prefix sum + backward scatter)
// The reference's output is reverse(stable ascending by bucket); we sort key = (R-1-bucket) ascending, stably,
// over the REVERSED input sequence, which is the same permutation.
//
// Bit-exactness: all float steps use explicit round-to-nearest intrinsics (no FMA contraction), the *1000.0 and
// *4096.0 products are f64 and truncate toward zero, int32 products wrap (done in uint32).
#pragma once
#include "common.cuh"

namespace gs {

enum DepthMode : int {
    kIntStatic = 0, kIntDynamic = 1, kIntPrecomputed = 2, kFloatStatic = 3, kFloatDynamic = 4, kFloatPrecomputed = 5
};

struct DepthParams {
    int32_t irow[4];   // static integer mode: trunc(mvp[2,6,10] * 1000.0), 1       sorter.cpp:64
    float frow[4];     // static float mode: mvp[2], mvp[6], mvp[10]                 sorter.cpp:130-133
    float mvp[16];     // dynamic modes: full matrix                                  sorter.cpp:48, :114
};

// ---------------------------------------------------------------------------------------------------------------
__global__ void k_sort_init(SortControl *ctl, uint32_t *lookback, size_t lookback_words) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (tid == 0) {
        ctl->dmin = 2147483640;
        ctl->dmax = -2147483640;
        ctl->error = 0;
        for (int i = 0; i < 4; ++i) ctl->ticket[i] = 0;
    }
    uint32_t *h = &ctl->hist[0][0];
    for (size_t i = tid; i < 4 * kRadix; i += stride) h[i] = 0;
    for (size_t i = tid; i < lookback_words; i += stride) lookback[i] = 0;
}

// third row of (mvp * T_scene), f32, left-to-right, unfused (sorter.cpp:11-15)
__device__ __forceinline__ void scene_row(const float *mvp, const float *t, float out[4]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float acc = __fmul_rn(mvp[2], t[4 * c + 0]);
        acc = __fadd_rn(acc, __fmul_rn(mvp[6], t[4 * c + 1]));
        acc = __fadd_rn(acc, __fmul_rn(mvp[10], t[4 * c + 2]));
        acc = __fadd_rn(acc, __fmul_rn(mvp[14], t[4 * c + 3]));
        out[c] = acc;
    }
}
__device__ __forceinline__ int32_t trunc_scaled(float f, double scale) { return __double2int_rz((double)f * scale); }

template <int MODE>
__device__ __forceinline__ int32_t splat_distance(uint32_t g, const void *__restrict__ centers, const void *__restrict__ pre,
                                                  const uint32_t *__restrict__ scene_idx, const DepthParams &P,
                                                  const int32_t (*s_irow)[4], const float (*s_frow)[4]) {
    if (MODE == kIntPrecomputed) return (int32_t)ld_nc_u32((const uint32_t *)pre + g);
    if (MODE == kFloatPrecomputed) return trunc_scaled(__uint_as_float(ld_nc_u32((const uint32_t *)pre + g)), 4096.0);
    const int4 c = ld_nc_v4((const int4 *)centers + g);
    if (MODE == kIntStatic) {
        return (int32_t)((uint32_t)P.irow[0] * (uint32_t)c.x + (uint32_t)P.irow[1] * (uint32_t)c.y + (uint32_t)P.irow[2] * (uint32_t)c.z);
    } else if (MODE == kIntDynamic) {
        const int32_t *r = s_irow[ld_nc_u32(scene_idx + g) & (GS_MAX_SCENES_DEV - 1)];
        return (int32_t)((uint32_t)r[0] * (uint32_t)c.x + (uint32_t)r[1] * (uint32_t)c.y + (uint32_t)r[2] * (uint32_t)c.z +
                         (uint32_t)r[3] * (uint32_t)c.w);
    } else if (MODE == kFloatStatic) {
        float acc = __fmul_rn(P.frow[0], __int_as_float(c.x));
        acc = __fadd_rn(acc, __fmul_rn(P.frow[1], __int_as_float(c.y)));
        acc = __fadd_rn(acc, __fmul_rn(P.frow[2], __int_as_float(c.z)));
        return trunc_scaled(acc, 4096.0);
    } else { // kFloatDynamic
        const float *r = s_frow[ld_nc_u32(scene_idx + g) & (GS_MAX_SCENES_DEV - 1)];
        float acc = __fmul_rn(r[0], __int_as_float(c.x));
        acc = __fadd_rn(acc, __fmul_rn(r[1], __int_as_float(c.y)));
        acc = __fadd_rn(acc, __fmul_rn(r[2], __int_as_float(c.z)));
        acc = __fadd_rn(acc, __fmul_rn(r[3], __int_as_float(c.w)));
        return trunc_scaled(acc, 4096.0);
    }
}

constexpr int kDepthThreads = 256;
constexpr int kDepthItems = 4;

// dist[i] for i in [s0, rc); indexes == nullptr means identity.  One min/max atomic pair per block.
template <int MODE, bool IDENTITY>
__global__ void __launch_bounds__(kDepthThreads)
k_depth(const uint32_t *__restrict__ indexes, const void *__restrict__ centers, const void *__restrict__ pre,
        const uint32_t *__restrict__ scene_idx, const float *__restrict__ transforms, DepthParams P, uint32_t s0,
        uint32_t rc, int32_t *__restrict__ dist, SortControl *ctl) {
    __shared__ int32_t s_irow[GS_MAX_SCENES_DEV][4];
    __shared__ float s_frow[GS_MAX_SCENES_DEV][4];
    __shared__ int s_min[kDepthThreads / 32], s_max[kDepthThreads / 32];
    if (MODE == kIntDynamic || MODE == kFloatDynamic) {
        if (threadIdx.x < GS_MAX_SCENES_DEV) {
            float row[4];
            scene_row(P.mvp, transforms + 16 * threadIdx.x, row);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                s_frow[threadIdx.x][k] = row[k];
                s_irow[threadIdx.x][k] = trunc_scaled(row[k], 1000.0);
            }
        }
        __syncthreads();
    }
    int32_t lmin = 2147483640, lmax = -2147483640;
    const uint32_t tile = kDepthThreads * kDepthItems;
    for (uint64_t base = (uint64_t)s0 + (uint64_t)blockIdx.x * tile; base < rc; base += (uint64_t)gridDim.x * tile) {
        uint32_t g[kDepthItems];
#pragma unroll
        for (int k = 0; k < kDepthItems; ++k) {
            const uint64_t i = base + (uint64_t)k * kDepthThreads + threadIdx.x;
            g[k] = (i < rc) ? (IDENTITY ? (uint32_t)i : ld_nc_u32(indexes + i)) : 0xffffffffu;
        }
        int32_t d[kDepthItems];
#pragma unroll
        for (int k = 0; k < kDepthItems; ++k)
            if (g[k] != 0xffffffffu) d[k] = splat_distance<MODE>(g[k], centers, pre, scene_idx, P, s_irow, s_frow);
#pragma unroll
        for (int k = 0; k < kDepthItems; ++k) {
            const uint64_t i = base + (uint64_t)k * kDepthThreads + threadIdx.x;
            if (i < rc) {
                dist[i] = d[k];
                lmin = min(lmin, d[k]);
                lmax = max(lmax, d[k]);
            }
        }
    }
    lmin = warp_min(lmin);
    lmax = warp_max(lmax);
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = lmin; s_max[threadIdx.x >> 5] = lmax; }
    __syncthreads();
    if (threadIdx.x < 32) {
        lmin = threadIdx.x < kDepthThreads / 32 ? s_min[threadIdx.x] : 2147483640;
        lmax = threadIdx.x < kDepthThreads / 32 ? s_max[threadIdx.x] : -2147483640;
        lmin = warp_min(lmin);
        lmax = warp_max(lmax);
        if (threadIdx.x == 0) {
            atomicMin(&ctl->dmin, lmin);
            atomicMax(&ctl->dmax, lmax);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Bucket + key.  key[j] = (R-1) - bucket(dist[i]) written at the REVERSED position j = rc-1-i, plus the digit
// histograms of every radix pass.  Optionally leaves the bucket in dist[i] (the reference's mappedDistances).
struct PassPlan {
    int npasses;
    int shift[4];
    int bits[4];
};

constexpr int kBucketThreads = 256;
constexpr int kBucketItems = 4;

template <typename KeyT>
__global__ void __launch_bounds__(kBucketThreads)
k_bucket(int32_t *__restrict__ dist, KeyT *__restrict__ keys, uint32_t s0, uint32_t rc, uint32_t R, PassPlan plan,
         int write_buckets, SortControl *ctl) {
    __shared__ uint32_t s_hist[4][kRadix];
    for (int i = threadIdx.x; i < 4 * kRadix; i += kBucketThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const int32_t dmin = ctl->dmin, dmax = ctl->dmax;
    const float span = __fsub_rn(__int2float_rn(dmax), __int2float_rn(dmin)); // sorter.cpp:142
    const float range_map = __fdiv_rn(__uint2float_rn(R - 1u), span);          // sorter.cpp:143
    const bool degenerate = (dmax == dmin);
    uint32_t err = 0;
    const uint32_t tile = kBucketThreads * kBucketItems;
    for (uint64_t base = (uint64_t)s0 + (uint64_t)blockIdx.x * tile; base < rc; base += (uint64_t)gridDim.x * tile) {
#pragma unroll
        for (int k = 0; k < kBucketItems; ++k) {
            const uint64_t i = base + (uint64_t)k * kBucketThreads + threadIdx.x;
            if (i < rc) {
                const int32_t d = dist[i];
                const int32_t rel = (int32_t)((uint32_t)d - (uint32_t)dmin);
                int32_t b = __float2int_rz(__fmul_rn(__int2float_rn(rel), range_map)); // sorter.cpp:146
                if (degenerate) b = 0;               // defined deviation: the reference traps / writes out of bounds
                if (b < 0 || (uint32_t)b >= R) { err |= kErrBucketRange; b = b < 0 ? 0 : (int32_t)(R - 1u); }
                if (write_buckets) dist[i] = b;
                const uint32_t key = (R - 1u) - (uint32_t)b;
                keys[(uint64_t)rc - 1u - i] = (KeyT)key;
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    if (p < plan.npasses) atomicAdd(&s_hist[p][(key >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.npasses * kRadix; i += kBucketThreads) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(&ctl->hist[0][0] + i, v);
    }
    if (err) atomicOr(&ctl->error, err);
    if (degenerate && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&ctl->error, kErrDegenerate);
}

// ---------------------------------------------------------------------------------------------------------------
// One-sweep radix pass (stable, 8-bit digit, decoupled look-back across tiles).
//   tile = kRadixThreads * kRadixItems consecutive elements; warp w owns a contiguous 32*ITEMS run, item k of lane l
//   is element run_base + 32k + l, so (warp, k, lane) order == sequence order and ranks are stable.
constexpr int kRadixThreads = 512;
constexpr int kRadixItems = 8;
constexpr int kRadixTile = kRadixThreads * kRadixItems;
constexpr int kRadixWarps = kRadixThreads / 32;

enum ValMode : int { kValArray = 0, kValArrayReversed = 1, kValIotaReversed = 2 };

constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagPrefix = 2u << 30;
constexpr uint32_t kFlagMask = 3u << 30;

template <typename KeyT, int VALMODE, bool WRITE_KEYS>
__global__ void __launch_bounds__(kRadixThreads)
k_radix_pass(const KeyT *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint32_t iota_top,
             KeyT *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t n_host,
             const unsigned long long *__restrict__ n_dev, unsigned long long n_cap, int shift, int bits,
             const uint32_t *__restrict__ pass_hist, uint32_t *lookback, uint32_t *ticket) {
    // element count: known on the host (depth sort) or only on the device (tile instances; grid sized for capacity)
    const uint32_t n = n_dev ? (uint32_t)min(*n_dev, n_cap) : n_host;
    // phase A: per-warp digit counters; phase B: reorder buffers (aliased)
    __shared__ __align__(16) unsigned char s_raw[kRadixTile * (sizeof(KeyT) + 4) > kRadixWarps * (kRadix + 1) * 4
                                                     ? kRadixTile * (sizeof(KeyT) + 4)
                                                     : kRadixWarps * (kRadix + 1) * 4];
    __shared__ uint32_t s_tile_count[kRadix];  // digit totals of this tile
    __shared__ uint32_t s_tile_start[kRadix];  // exclusive scan of the above (slot in the reorder buffer)
    __shared__ uint32_t s_gbase[kRadix];       // global slot of reorder-slot 0 for each digit (wrapping arithmetic)
    __shared__ uint32_t s_scan[40];
    __shared__ uint32_t s_tile_id;

    uint32_t(*s_whist)[kRadix + 1] = reinterpret_cast<uint32_t(*)[kRadix + 1]>(s_raw);
    uint32_t *s_vals = reinterpret_cast<uint32_t *>(s_raw);
    KeyT *s_keys = reinterpret_cast<KeyT *>(s_raw + (size_t)kRadixTile * 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile_id = atomicAdd(ticket, 1u);
    for (int i = tid; i < kRadixWarps * (kRadix + 1); i += kRadixThreads) (&s_whist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile_id;
    const uint64_t tile_base = (uint64_t)tile * kRadixTile;
    if (tile_base >= n) return; // surplus CTA of a capacity-sized grid (uniform exit)
    const uint32_t dmask = (1u << bits) - 1u;

    // ---- load keys (+ values), rank within the warp ----------------------------------------------------------
    uint32_t key[kRadixItems], val[kRadixItems], rank[kRadixItems];
    const uint64_t run_base = tile_base + (uint64_t)warp * (32 * kRadixItems);
#pragma unroll
    for (int k = 0; k < kRadixItems; ++k) {
        const uint64_t e = run_base + (uint64_t)k * 32 + lane;
        key[k] = (e < n) ? (uint32_t)keys_in[e] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kRadixItems; ++k) {
        const uint64_t e = run_base + (uint64_t)k * 32 + lane;
        if (e < n) {
            if (VALMODE == kValArray) val[k] = ld_nc_u32(vals_in + e);
            else if (VALMODE == kValArrayReversed) val[k] = ld_nc_u32(vals_in + ((uint64_t)n - 1u - e));
            else val[k] = iota_top - (uint32_t)e;
        } else val[k] = 0;
    }
#pragma unroll
    for (int k = 0; k < kRadixItems; ++k) {
        const uint64_t e = run_base + (uint64_t)k * 32 + lane;
        const uint32_t d = (e < n) ? ((key[k] >> shift) & dmask) : (uint32_t)kRadix; // tail items: private bin
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t before = s_whist[warp][d];
        rank[k] = before + __popc(peers & lanemask_lt());
        __syncwarp();
        if ((peers & lanemask_lt()) == 0) s_whist[warp][d] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: exclusive scan over warps, tile totals -------------------------------------------------------
    if (tid < kRadix) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kRadixWarps; ++w) {
            const uint32_t c = s_whist[w][tid];
            s_whist[w][tid] = run;
            run += c;
        }
        s_tile_count[tid] = run;
    }
    __syncthreads();
    {
        uint32_t total;
        const uint32_t c = tid < kRadix ? s_tile_count[tid] : 0u;
        const uint32_t ex = block_exclusive_scan<kRadixThreads>(c, s_scan, total);
        if (tid < kRadix) s_tile_start[tid] = ex;
    }
    // global digit bases: exclusive scan of the pass histogram (recomputed by every tile, 256 values)
    {
        uint32_t total;
        const uint32_t c = tid < kRadix ? pass_hist[tid] : 0u;
        const uint32_t ex = block_exclusive_scan<kRadixThreads>(c, s_scan, total);
        if (tid < kRadix) s_gbase[tid] = ex; // completed below
    }
    // reorder slots for my items (must be read before the counters are overwritten by the reorder buffers)
    uint32_t slot[kRadixItems];
#pragma unroll
    for (int k = 0; k < kRadixItems; ++k) {
        const uint64_t e = run_base + (uint64_t)k * 32 + lane;
        const uint32_t d = (key[k] >> shift) & dmask;
        slot[k] = (e < n) ? (s_tile_start[d] + s_whist[warp][d] + rank[k]) : 0xffffffffu;
    }
    __syncthreads();

    // ---- decoupled look-back: one thread per digit -----------------------------------------------------------------
    if (tid < kRadix) {
        const uint32_t agg = s_tile_count[tid];
        uint32_t excl = 0;
        uint32_t *mine = lookback + (size_t)tile * kRadix + tid;
        if (tile == 0) {
            st_release_u32(mine, agg | kFlagPrefix);
        } else {
            st_release_u32(mine, agg | kFlagAggregate);
            int64_t t = (int64_t)tile - 1;
            while (true) {
                const uint32_t v = ld_acquire_u32(lookback + (size_t)t * kRadix + tid);
                const uint32_t f = v & kFlagMask;
                if (f == 0) continue;
                excl += v & ~kFlagMask;
                if (f == kFlagPrefix) break;
                --t;
            }
            st_release_u32(mine, (excl + agg) | kFlagPrefix);
        }
        s_gbase[tid] = s_gbase[tid] + excl - s_tile_start[tid];
    }
    // ---- reorder through shared memory so each digit run is written contiguously --------------------------------------
#pragma unroll
    for (int k = 0; k < kRadixItems; ++k)
        if (slot[k] != 0xffffffffu) {
            s_vals[slot[k]] = val[k];
            s_keys[slot[k]] = (KeyT)key[k];
        }
    __syncthreads();
    const uint32_t valid = (uint32_t)min((uint64_t)kRadixTile, (uint64_t)n - tile_base);
#pragma unroll
    for (int k = 0; k < kRadixItems; ++k) {
        const uint32_t q = (uint32_t)k * kRadixThreads + tid;
        if (q < valid) {
            const KeyT kk = s_keys[q];
            const uint32_t d = ((uint32_t)kk >> shift) & dmask;
            const uint32_t dst = s_gbase[d] + q;
            vals_out[dst] = s_vals[q];
            if (WRITE_KEYS) keys_out[dst] = kk;
        }
    }
}

// out[0..s0) = indexes[0..s0)   (sorter.cpp:158-160)
__global__ void k_copy_head(const uint32_t *__restrict__ indexes, uint32_t *__restrict__ out, uint32_t s0) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < s0; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = indexes ? indexes[i] : (uint32_t)i;
}

// ---- scratch reproduction for the stateless drop-in: frequencies[b] = #sorted splats in buckets < b --------------------
__global__ void k_bucket_counts(const int32_t *__restrict__ buckets, uint32_t s0, uint32_t rc, uint32_t *counts) {
    for (uint64_t i = (uint64_t)s0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rc; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&counts[buckets[i]], 1u);
}
__global__ void __launch_bounds__(1024) k_exclusive_scan_single_block(uint32_t *data, uint32_t n) {
    __shared__ uint32_t s_scan[40];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? data[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan<1024>(v, s_scan, total);
        const uint32_t carry = s_carry;
        if (i < n) data[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
}

// ---- D1: SplatMesh.computeDistancesOnGPU (SplatMesh.js:1451-1502): distances in SPLAT order ---------------------------
template <bool INTEGER, bool DYNAMIC>
__global__ void k_distances_splat_order(const void *__restrict__ centers, const uint32_t *__restrict__ scene_idx,
                                        const int32_t *__restrict__ irows /*[scenes][4]*/, const float *__restrict__ frows,
                                        uint32_t count, void *__restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        const int4 c = ld_nc_v4((const int4 *)centers + i);
        const uint32_t sc = DYNAMIC ? (scene_idx[i] & (GS_MAX_SCENES_DEV - 1)) : 0u;
        if (INTEGER) {
            const int32_t *r = irows + 4 * sc;
            uint32_t d = (uint32_t)c.x * (uint32_t)r[0] + (uint32_t)c.y * (uint32_t)r[1] + (uint32_t)c.z * (uint32_t)r[2];
            if (DYNAMIC) d += (uint32_t)r[3] * (uint32_t)c.w;
            ((int32_t *)out)[i] = (int32_t)d;
        } else {
            const float *r = frows + 4 * sc;
            float acc = __fmul_rn(__int_as_float(c.x), r[0]);
            acc = __fadd_rn(acc, __fmul_rn(__int_as_float(c.y), r[1]));
            acc = __fadd_rn(acc, __fmul_rn(__int_as_float(c.z), r[2]));
            if (DYNAMIC) acc = __fadd_rn(acc, r[3]);
            ((float *)out)[i] = acc;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Host-side launch helpers shared by the depth sort and the tile-instance sort.
static inline PassPlan make_plan_bits(int key_bits) {
    PassPlan pl{};
    int np = (key_bits + kRadixBits - 1) / kRadixBits;
    if (np < 1) np = 1;
    pl.npasses = np;
    int done = 0;
    for (int p = 0; p < np; ++p) {
        const int left = key_bits - done, b = (left + (np - p) - 1) / (np - p);
        pl.shift[p] = done;
        pl.bits[p] = b < 1 ? 1 : b;
        done += b;
    }
    return pl;
}

template <typename KeyT>
static void launch_radix_pass(uint32_t grid, bool first, bool write_keys, int valmode, const KeyT *kin, const uint32_t *vin, uint32_t iota_top,
                              KeyT *kout, uint32_t *vout, uint32_t n, const unsigned long long *n_dev, unsigned long long n_cap, int shift, int bits,
                              const uint32_t *hist, uint32_t *lookback, uint32_t *ticket, cudaStream_t st) {
#define GS_PASS(VM, WK) k_radix_pass<KeyT, VM, WK><<<grid, kRadixThreads, 0, st>>>(kin, vin, iota_top, kout, vout, n, n_dev, n_cap, shift, bits, hist, lookback, ticket)
    if (!first) valmode = kValArray;
    if (valmode == kValArray) { if (!write_keys) GS_PASS(kValArray, false); else GS_PASS(kValArray, true); }
    else if (valmode == kValArrayReversed) { if (!write_keys) GS_PASS(kValArrayReversed, false); else GS_PASS(kValArrayReversed, true); }
    else { if (!write_keys) GS_PASS(kValIotaReversed, false); else GS_PASS(kValIotaReversed, true); }
#undef GS_PASS
}

// Stable LSD radix sort of (key, value) pairs.  `first_valmode` selects how the first pass obtains its values.
// Sorted values land in vals_final; sorted keys (if write_keys_last) in keys1 when npasses is odd, else keys0.
template <typename KeyT>
static void radix_sort_pairs(KeyT *keys0, KeyT *keys1, const uint32_t *vals_src, uint32_t iota_top, int first_valmode, uint32_t *vtmp0,
                             uint32_t *vtmp1, uint32_t *vals_final, uint32_t n, const unsigned long long *n_dev, unsigned long long n_cap,
                             const PassPlan &pl, SortControl *ctl, uint32_t *lookback, bool write_keys_last, cudaStream_t st, uint32_t &launches,
                             Profiler *prof = nullptr, const char *const *pass_names = nullptr) {
    const unsigned long long n_grid = n_dev ? n_cap : n;
    const uint32_t tiles = (uint32_t)((n_grid + kRadixTile - 1) / kRadixTile);
    if (!tiles) return;
    KeyT *kin = keys0, *kout = keys1;
    const uint32_t *vin = vals_src;
    uint32_t *vt[2] = {vtmp0, vtmp1};
    for (int p = 0; p < pl.npasses; ++p) {
        const bool last = (p == pl.npasses - 1);
        uint32_t *vout = last ? vals_final : vt[p & 1];
        launch_radix_pass<KeyT>(tiles, p == 0, !last || write_keys_last, first_valmode, kin, vin, iota_top, kout, vout, n, n_dev, n_cap, pl.shift[p],
                                pl.bits[p], &ctl->hist[p][0], lookback + (size_t)p * tiles * kRadix, &ctl->ticket[p], st);
        ++launches;
        if (prof) prof->mark(pass_names ? pass_names[p] : "k_radix_pass", st);
        vin = vout;
        KeyT *t = kin; kin = kout; kout = t;
    }
}
static inline size_t radix_lookback_words(unsigned long long n, int npasses) {
    const size_t tiles = (size_t)((n + kRadixTile - 1) / kRadixTile);
    return tiles * kRadix * (size_t)npasses;
}

} // namespace gs
