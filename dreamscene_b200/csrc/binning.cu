// Tile binning and per-tile depth sort.
//
// Replaces upstream's InclusiveSum + duplicateWithKeys + global 64-bit cub::DeviceRadixSort +
// identifyTileRanges with a counting-sort by tile followed by
// an independent in-shared-memory sort of every tile's list:
//   project_sh        : per-tile pair counts (atomics)                       [project.cu]
//   scan_order_kernel : exclusive scan -> tile_start[], cursors, total pair count, and the
//                       work order (tiles bucketed by list length, longest first)
//   scatter_kernel    : every Gaussian appends (depth_bits<<32 | idx) to each tile it touches
//                       (arrival order inside a tile is arbitrary ...)
//   sort kernels      : ... and is then fixed by sorting each tile's keys on (depth bits, idx):
//                       identical to a stable sort of (tile<<32 | depth bits) over pairs emitted
//                       in Gaussian-index order, i.e. bit-exact with the oracle's lists.
//                       Epilogue gathers the 48-B record of every entry into the tile-major,
//                       depth-sorted record array the composite kernels stream with TMA.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// scan + work order: one CTA of 1024 threads (num_tiles is small: 4096 at 1024^2)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int size_bucket(uint32_t n) {   // 0 = longest lists
    return n == 0 ? 32 : __clz(n);                          // clz in [0,31] for n>0
}

__global__ void __launch_bounds__(1024)
scan_order_kernel(int ntiles, uint32_t max_pairs, const uint32_t* __restrict__ tile_count,
                  uint32_t* __restrict__ tile_start, uint32_t* __restrict__ tile_cursor,
                  uint32_t* __restrict__ work_order, uint32_t* __restrict__ header,
                  volatile uint32_t* host_notify, uint32_t notify_seq) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    __shared__ uint32_t bucket_cnt[33];
    __shared__ uint32_t bucket_pos[33];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    if (tid < 33) bucket_cnt[tid] = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        const int t = base + tid;
        const uint32_t v = (t < ntiles) ? tile_count[t] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_sums[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += n;
            }
            warp_sums[lane] = wi - w;   // exclusive
        }
        __syncthreads();
        const uint32_t excl = carry_s + warp_sums[wid] + incl - v;
        if (t < ntiles) {
            tile_start[t] = excl;
            tile_cursor[t] = excl;
            atomicAdd(&bucket_cnt[size_bucket(v)], 1u);
        }
        __syncthreads();
        if (tid == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t total = carry_s;
        tile_start[ntiles] = total;
        header[GSR_H_NUM_PAIRS] = total;
        header[GSR_H_MAX_PAIRS] = max_pairs;
        header[GSR_H_NUM_TILES] = (uint32_t)ntiles;
        header[GSR_H_OVERFLOW] = total > max_pairs ? 1u : 0u;
        uint32_t run = 0, nbig = 0;
        for (int b = 0; b < 33; ++b) {
            bucket_pos[b] = run;
            run += bucket_cnt[b];
            // lists longer than GSR_SORT_SMALL_MAX (=2^12): n >= 4097 -> clz <= 19
            if (b <= 19) nbig = run;
        }
        header[GSR_H_NUM_BIG] = nbig;   // upper bound: includes n == 4096 exactly (clz 19)
        header[GSR_H_NUM_NONEMPTY] = run - bucket_cnt[32];
        if (host_notify != nullptr) {   // mapped pinned host memory: tell the host the pair count now
            host_notify[1] = total;
            host_notify[2] = total > max_pairs ? 1u : 0u;
            host_notify[3] = (uint32_t)ntiles;
            __threadfence_system();
            host_notify[0] = notify_seq;
        }
    }
    __syncthreads();
    for (int t = tid; t < ntiles; t += 1024) {
        const uint32_t pos = atomicAdd(&bucket_pos[size_bucket(tile_count[t])], 1u);
        work_order[pos] = (uint32_t)t;
    }
}

// ---------------------------------------------------------------------------------------------
// scatter: one Gaussian per thread appends its key to every touched tile
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scatter_kernel(int P, int gx, uint32_t max_pairs, const uint4* __restrict__ rectdepth,
               uint32_t* __restrict__ tile_cursor, unsigned long long* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint4 rd = __ldg(rectdepth + i);
    if (rd.w == 0) return;
    const int minx = rd.x & 0xffff, miny = rd.x >> 16, maxx = rd.y & 0xffff, maxy = rd.y >> 16;
    const unsigned long long key = ((unsigned long long)rd.z << 32) | (uint32_t)i;
    for (int ty = miny; ty < maxy; ++ty)
        for (int tx = minx; tx < maxx; ++tx) {
            const uint32_t pos = atomicAdd(tile_cursor + ty * gx + tx, 1u);
            if (pos < max_pairs) keys[pos] = key;
        }
}

// ---------------------------------------------------------------------------------------------
// bitonic network (flip variant: every comparator leaves the minimum at the lower index, so
// virtual +inf padding above n needs no storage: comparators reaching past n are no-ops)
// ---------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void bitonic_smem(unsigned long long* s, int n, int k_begin, int k_end,
                                             int j_first_limit) {
    // runs merge sizes k = k_begin .. k_end (powers of two); for each k the steps with
    // distance < j_first_limit only (used by the out-of-core path); normally j_first_limit = inf
    for (int k = k_begin; k <= k_end; k <<= 1) {
        const int half = k >> 1;
        if (half < j_first_limit) {
            // flip step: i = b*k + off, l = b*k + k-1-off
            for (int t = threadIdx.x; ; t += NT) {
                const int b = t / half, off = t - b * half;
                const int i = b * k + off, l = b * k + k - 1 - off;
                if (i >= n) break;
                if (l < n) {
                    const unsigned long long a = s[i], c = s[l];
                    if (a > c) { s[i] = c; s[l] = a; }
                }
            }
            __syncthreads();
        }
        for (int j = half >> 1; j > 0; j >>= 1) {
            if (j >= j_first_limit) continue;
            for (int t = threadIdx.x; ; t += NT) {
                const int i = 2 * j * (t / j) + (t % j), l = i + j;
                if (i >= n) break;
                if (l < n) {
                    const unsigned long long a = s[i], c = s[l];
                    if (a > c) { s[i] = c; s[l] = a; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int next_pow2(int n) {
    return n <= 1 ? 1 : 1 << (32 - __clz(n - 1));
}

// copy geom[idx] records into the sorted array: 3 lanes per record, one 16-B part each
template <int NT, bool GLOBAL_KEYS>
__device__ __forceinline__ void gather_records(const unsigned long long* keys_sorted, int n,
                                               const GsrRec* __restrict__ geom,
                                               GsrRec* __restrict__ out) {
    const float4* g4 = reinterpret_cast<const float4*>(geom);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (int j = threadIdx.x; j < 3 * n; j += NT) {
        const int r = j / 3, part = j - 3 * r;
        const uint32_t idx = GLOBAL_KEYS ? (uint32_t)__ldcg(keys_sorted + r) : (uint32_t)keys_sorted[r];
        o4[j] = __ldg(g4 + 3 * (size_t)idx + part);
    }
}

// small tiles: n <= 4096 keys entirely in 32 KB of shared memory, 256 threads
__global__ void __launch_bounds__(256)
sort_small_kernel(const uint32_t* __restrict__ header, const uint32_t* __restrict__ work_order,
                  const uint32_t* __restrict__ tile_start, unsigned long long* __restrict__ keys,
                  const GsrRec* __restrict__ geom, GsrRec* __restrict__ records) {
    __shared__ unsigned long long s[GSR_SORT_SMALL_MAX];
    const uint32_t max_pairs = header[GSR_H_MAX_PAIRS];
    const uint32_t nonempty = header[GSR_H_NUM_NONEMPTY];
    // tiles [0, NUM_BIG) may hold > 4096 keys; those exactly at 4096 are handled here too
    for (uint32_t w = blockIdx.x; w < nonempty; w += gridDim.x) {
        const uint32_t tile = work_order[w];
        uint32_t beg = tile_start[tile], end = tile_start[tile + 1];
        if (end - beg > GSR_SORT_SMALL_MAX) continue;          // big kernel's job
        if (end > max_pairs) end = max_pairs;                  // overflow: stay in bounds
        if (beg >= end) continue;
        const int n = (int)(end - beg);
        for (int i = threadIdx.x; i < n; i += 256) s[i] = keys[beg + i];
        __syncthreads();
        bitonic_smem<256>(s, n, 2, next_pow2(n), 1 << 30);
        gather_records<256, false>(s, n, geom, records + beg);
        __syncthreads();
    }
}

// big tiles: 1024 threads; up to 16384 keys in 128 KB smem; beyond that a hierarchical
// (out-of-core) bitonic sort working in place on the global key segment.
__global__ void __launch_bounds__(1024)
sort_big_kernel(const uint32_t* __restrict__ header, const uint32_t* __restrict__ work_order,
                const uint32_t* __restrict__ tile_start, unsigned long long* __restrict__ keys,
                const GsrRec* __restrict__ geom, GsrRec* __restrict__ records) {
    extern __shared__ __align__(16) unsigned long long sb[];
    constexpr int CH = GSR_SORT_BIG_SMEM;
    const uint32_t max_pairs = header[GSR_H_MAX_PAIRS];
    const uint32_t nbig = header[GSR_H_NUM_BIG];
    for (uint32_t w = blockIdx.x; w < nbig; w += gridDim.x) {
        const uint32_t tile = work_order[w];
        uint32_t beg = tile_start[tile], end = tile_start[tile + 1];
        if (end - beg <= GSR_SORT_SMALL_MAX) continue;
        if (end > max_pairs) end = max_pairs;
        if (beg >= end) continue;
        const int n = (int)(end - beg);
        unsigned long long* gk = keys + beg;
        if (n <= CH) {
            for (int i = threadIdx.x; i < n; i += 1024) sb[i] = gk[i];
            __syncthreads();
            bitonic_smem<1024>(sb, n, 2, next_pow2(n), 1 << 30);
            gather_records<1024, false>(sb, n, geom, records + beg);
            __syncthreads();
            continue;
        }
        // ---- out-of-core: sort CH-sized chunks, then merge with global + shared steps ------
        const int N = next_pow2(n);
        for (int c0 = 0; c0 < n; c0 += CH) {
            const int m = min(CH, n - c0);
            for (int i = threadIdx.x; i < m; i += 1024) sb[i] = gk[c0 + i];
            __syncthreads();
            bitonic_smem<1024>(sb, m, 2, CH, 1 << 30);
            for (int i = threadIdx.x; i < m; i += 1024) gk[c0 + i] = sb[i];
            __syncthreads();
        }
        for (int k = 2 * CH; k <= N; k <<= 1) {
            const int half = k >> 1;
            // flip step in global memory (distance up to k-1 >= CH)
            for (int t = threadIdx.x; t < N / 2; t += 1024) {
                const int b = t / half, off = t - b * half;
                const int i = b * k + off, l = b * k + k - 1 - off;
                if (i < n && l < n) {
                    const unsigned long long a = __ldcg(gk + i), c = __ldcg(gk + l);
                    if (a > c) { __stcg(gk + i, c); __stcg(gk + l, a); }
                }
            }
            __syncthreads();
            for (int j = half >> 1; j >= CH; j >>= 1) {
                for (int t = threadIdx.x; t < N / 2; t += 1024) {
                    const int i = 2 * j * (t / j) + (t % j), l = i + j;
                    if (i < n && l < n) {
                        const unsigned long long a = __ldcg(gk + i), c = __ldcg(gk + l);
                        if (a > c) { __stcg(gk + i, c); __stcg(gk + l, a); }
                    }
                }
                __syncthreads();
            }
            // remaining distances < CH: independent inside every CH chunk -> shared memory
            for (int c0 = 0; c0 < n; c0 += CH) {
                const int m = min(CH, n - c0);
                for (int i = threadIdx.x; i < m; i += 1024) sb[i] = __ldcg(gk + c0 + i);
                __syncthreads();
                for (int j = CH >> 1; j > 0; j >>= 1) {
                    for (int t = threadIdx.x; ; t += 1024) {
                        const int i = 2 * j * (t / j) + (t % j), l = i + j;
                        if (i >= m) break;
                        if (l < m) {
                            const unsigned long long a = sb[i], c = sb[l];
                            if (a > c) { sb[i] = c; sb[l] = a; }
                        }
                    }
                    __syncthreads();
                }
                for (int i = threadIdx.x; i < m; i += 1024) __stcg(gk + c0 + i, sb[i]);
                __syncthreads();
            }
        }
        __threadfence_block();
        gather_records<1024, true>(gk, n, geom, records + beg);
        __syncthreads();
    }
}

}  // namespace

cudaError_t gsr_launch_binning(const GsrFwdArgs& a) {
    const GsrTileGrid grid = gsr_grid(a.prm.image_height, a.prm.image_width);
    uint32_t* header = reinterpret_cast<uint32_t*>(a.saved + a.vl.header);
    uint32_t* tile_start = reinterpret_cast<uint32_t*>(a.saved + a.vl.tile_start);
    uint32_t* work_order = reinterpret_cast<uint32_t*>(a.saved + a.vl.work_order);
    GsrRec* records = reinterpret_cast<GsrRec*>(a.saved + a.vl.records);
    uint32_t* tile_count = reinterpret_cast<uint32_t*>(a.scratch + a.sl.tile_count);
    uint32_t* tile_cursor = reinterpret_cast<uint32_t*>(a.scratch + a.sl.tile_cursor);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(a.scratch + a.sl.keys);
    const GsrRec* geom = reinterpret_cast<const GsrRec*>(a.scratch + a.sl.geom);
    const uint4* rectdepth = reinterpret_cast<const uint4*>(a.scratch + a.sl.rectdepth);

    scan_order_kernel<<<1, 1024, 0, a.stream>>>(grid.ntiles, a.max_pairs, tile_count, tile_start,
                                                 tile_cursor, work_order, header, a.host_notify,
                                                 a.notify_seq);
    if (a.prm.P > 0)
        scatter_kernel<<<(a.prm.P + 255) / 256, 256, 0, a.stream>>>(a.prm.P, grid.gx, a.max_pairs,
                                                                     rectdepth, tile_cursor, keys);
    const int big_smem = GSR_SORT_BIG_SMEM * 8;
    {
        cudaError_t e = cudaFuncSetAttribute(sort_big_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
        if (e != cudaSuccess) return e;
    }
    int nsm = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    }
    const int small_grid = min(grid.ntiles, nsm * 6);
    const int big_grid = min(grid.ntiles, nsm);
    sort_big_kernel<<<big_grid, 1024, big_smem, a.stream>>>(header, work_order, tile_start, keys, geom,
                                                            records);
    sort_small_kernel<<<small_grid, 256, 0, a.stream>>>(header, work_order, tile_start, keys, geom,
                                                        records);
    return cudaGetLastError();
}
