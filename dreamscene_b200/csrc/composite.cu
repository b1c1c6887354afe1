// composite_fwd / composite_bwd: per-tile front-to-back alpha blending and its backward.
//
// Replaces upstream's renderCUDA forward/backward (restated
// in oracle/splat_ref.py::composite).  B200 design:
//   * persistent CTAs (256 threads = one 16x16 tile, 8 warps x (8x4)-pixel blocks) pull tiles
//     from a queue ordered longest-list-first;
//   * a tile's depth-sorted 48-byte records are one contiguous byte range, streamed into a
//     3-stage shared-memory ring with cp.async.bulk (TMA 1-D, SASS UBLKCP) + mbarrier;
//   * each warp tests 32 records at a time (one per lane) against its own 8x4 pixel block
//     (conservative extent test, never changes which pixels blend) and only evaluates the
//     survivors, in list order, reading them back with broadcast LDS.128;
//   * forward: warp/CTA early-out when every pixel is saturated (T < 1e-4 would follow);
//   * backward: per-(warp, Gaussian) partial gradients are reduced with a halving shuffle
//     butterfly (12 SHFL for 10 values) and committed with ONE coalesced RED instruction.
#include "common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 256;   // records per pipeline stage (12 KB)
constexpr int kStages = 3;

struct __align__(128) SmemRing {
    GsrRec rec[kStages][kChunk];
    uint64_t full[kStages];
    uint32_t work;      // broadcast slot for the tile queue
    uint32_t maxlast;   // backward: max n_contrib of the tile
};

// The blending test, shared verbatim by forward and backward so both take identical decisions.
struct PairEval {
    float dx, dy, G, alpha;
    bool valid;
};
__device__ __forceinline__ PairEval eval_pair(float px, float py, float A, float B, float C,
                                              float opacity, float X, float Y) {
    PairEval e;
    e.dx = __fsub_rn(px, X);
    e.dy = __fsub_rn(py, Y);
    const float u = __fmaf_rn(A, e.dx, __fmul_rn(B, e.dy));
    const float p2 = __fmaf_rn(__fmul_rn(C, e.dy), e.dy, __fmul_rn(u, e.dx));
    e.G = ex2_approx(p2);
    e.alpha = fminf(GSR_ALPHA_MAX, __fmul_rn(opacity, e.G));
    e.valid = (p2 <= 0.0f) && (e.alpha >= GSR_ALPHA_MIN);
    return e;
}

__device__ __forceinline__ bool cull_pass(float px, float py, uint32_t ext, float X0, float Y0) {
    // block covers pixel centres [X0, X0+7] x [Y0, Y0+3]
    const __half2 eh = *reinterpret_cast<const __half2*>(&ext);
    const float2 e = __half22float2(eh);
    const float ddx = fmaxf(fmaxf(X0 - px, px - (X0 + 7.0f)), 0.0f);
    const float ddy = fmaxf(fmaxf(Y0 - py, py - (Y0 + 3.0f)), 0.0f);
    return (ddx <= e.x) && (ddy <= e.y);
}

// =============================================================================================
// Forward
// =============================================================================================
template <bool SCORE>
__global__ void __launch_bounds__(kThreads)
composite_fwd_kernel(int H, int W, int gx, int ntiles, const uint32_t* __restrict__ header,
                     const uint32_t* __restrict__ work_order,
                     const uint32_t* __restrict__ tile_start, const GsrRec* __restrict__ records,
                     const float* __restrict__ bg, uint32_t* __restrict__ queue,
                     float* __restrict__ out_color, float* __restrict__ out_depth_alpha,
                     uint32_t* __restrict__ n_contrib, float* __restrict__ score) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemRing& sm = *reinterpret_cast<SmemRing*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&sm.full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint32_t max_pairs = header[GSR_H_MAX_PAIRS];
    const float bg0 = __ldg(bg), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
    uint32_t gc = 0;   // chunks issued so far by this CTA (slot = gc % kStages, parity = (gc/kStages)&1)

    for (;;) {
        if (tid == 0) sm.work = atomicAdd(queue, 1u);
        __syncthreads();
        const uint32_t w = sm.work;
        __syncthreads();   // everyone has read the slot before thread 0 may overwrite it
        if (w >= (uint32_t)ntiles) break;
        const uint32_t tile = work_order[w];
        uint32_t beg = tile_start[tile], end = tile_start[tile + 1];
        if (end > max_pairs) end = max_pairs;
        if (beg > end) beg = end;
        const int n = (int)(end - beg);
        const int nchunks = (n + kChunk - 1) / kChunk;
        const int tyi = tile / gx, txi = tile - tyi * gx;
        const int X0i = txi * GSR_TILE + (wid & 1) * 8, Y0i = tyi * GSR_TILE + (wid >> 1) * 4;
        const int Xi = X0i + (lane & 7), Yi = Y0i + (lane >> 3);
        const float X0 = (float)X0i, Y0 = (float)Y0i, X = (float)Xi, Y = (float)Yi;
        const bool inside = Xi < W && Yi < H;
        bool done = !inside;
        float T = 1.0f, Cr = 0.f, Cg = 0.f, Cb = 0.f, Dacc = 0.f;
        uint32_t last = 0;

        // prologue: prefetch up to kStages-1 chunks
        if (tid == 0) {
            const int npre = min(nchunks, kStages - 1);
            for (int c = 0; c < npre; ++c) {
                const uint32_t slot = (gc + c) % kStages;
                const uint32_t cnt = (uint32_t)min(kChunk, n - c * kChunk);
                mbar_expect_tx(&sm.full[slot], cnt * (uint32_t)sizeof(GsrRec));
                bulk_g2s(sm.rec[slot], records + beg + (size_t)c * kChunk, cnt * (uint32_t)sizeof(GsrRec),
                         &sm.full[slot]);
            }
        }
        int c = 0;
        for (; c < nchunks; ++c) {
            // everyone has finished chunk c-1 -> its slot may be refilled; also the CTA early-out
            const int ndone = __syncthreads_count(done);
            if (ndone == kThreads) break;
            if (tid == 0 && c + kStages - 1 < nchunks) {
                const int cn = c + kStages - 1;
                const uint32_t slot = (gc + cn) % kStages;
                const uint32_t cnt = (uint32_t)min(kChunk, n - cn * kChunk);
                mbar_expect_tx(&sm.full[slot], cnt * (uint32_t)sizeof(GsrRec));
                bulk_g2s(sm.rec[slot], records + beg + (size_t)cn * kChunk, cnt * (uint32_t)sizeof(GsrRec),
                         &sm.full[slot]);
            }
            const uint32_t g = gc + c, slot = g % kStages;
            mbar_wait(&sm.full[slot], (g / kStages) & 1u);
            const GsrRec* st = sm.rec[slot];
            const int cnt = min(kChunk, n - c * kChunk);
            if (__all_sync(0xffffffffu, done)) continue;   // this warp is saturated
            for (int sub = 0; sub * 32 < cnt; ++sub) {
                const int r = sub * 32 + lane;
                bool pass = false;
                if (r < cnt) {
                    const float4 q0 = *reinterpret_cast<const float4*>(&st[r]);
                    pass = cull_pass(q0.x, q0.y, __float_as_uint(q0.z), X0, Y0);
                }
                uint32_t mask = __ballot_sync(0xffffffffu, pass);
                const float4* sp = reinterpret_cast<const float4*>(&st[sub * 32]);
                const uint32_t pos0 = (uint32_t)(c * kChunk + sub * 32 + 1);
                while (mask) {
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float4* rp = sp + 3 * b;
                    const float4 q0 = rp[0], q1 = rp[1], q2 = rp[2];
                    const PairEval e = eval_pair(q0.x, q0.y, q0.w, q1.x, q1.y, q1.z, X, Y);
                    float wgt = 0.f;
                    if (e.valid && !done) {
                        const float Tn = T * (1.0f - e.alpha);
                        if (Tn < GSR_T_STOP) {
                            done = true;
                        } else {
                            wgt = e.alpha * T;
                            Cr = fmaf(q2.x, wgt, Cr);
                            Cg = fmaf(q2.y, wgt, Cg);
                            Cb = fmaf(q2.z, wgt, Cb);
                            Dacc = fmaf(q1.w, wgt, Dacc);
                            T = Tn;
                            last = pos0 + (uint32_t)b;
                        }
                    }
                    if (SCORE) {
                        float s = wgt;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                        if (lane == 0 && s != 0.f) atomicAdd(score + __float_as_uint(q2.w), s);
                    }
                }
                if (__all_sync(0xffffffffu, done)) break;
            }
        }
        // drain chunks that were prefetched but never consumed, keep the ring position in step
        const int issued = min(nchunks, c + kStages - 1);
        if (tid == 0)
            for (int k = c; k < issued; ++k) {
                const uint32_t g = gc + k;
                mbar_wait(&sm.full[g % kStages], (g / kStages) & 1u);
            }
        gc += (uint32_t)issued;

        if (inside) {
            const size_t pix = (size_t)Yi * W + Xi, plane = (size_t)H * W;
            out_color[pix] = fmaf(T, bg0, Cr);
            out_color[plane + pix] = fmaf(T, bg1, Cg);
            out_color[2 * plane + pix] = fmaf(T, bg2, Cb);
            out_depth_alpha[pix] = Dacc;
            out_depth_alpha[plane + pix] = T;
            n_contrib[pix] = last;
        }
    }
}

// =============================================================================================
// Backward
// =============================================================================================
// Halving butterfly: after the call v[0] of lane L holds the warp-wide sum of value
// `vidx(L)` (see bwd_value_index); 5+3+2+1+1 = 12 shuffles for 10 values.
template <int N, int XOR>
__device__ __forceinline__ void halve(float (&v)[10], bool hi) {
    constexpr int Hh = (N + 1) / 2;
#pragma unroll
    for (int k = 0; k < Hh; ++k) {
        const float lo = v[k];
        const float hv = (Hh + k < N) ? v[Hh + k] : 0.0f;
        const float send = hi ? lo : hv;
        const float keep = hi ? hv : lo;
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, XOR);
    }
}
__device__ __forceinline__ int bwd_value_index(int lane) {
    // which of the 10 values this lane ends up owning (-1: a padding slot)
    int base = 0, n = 10;
    if (lane & 16) { base += 5; n = 5; } else { n = 5; }
    if (lane & 8) { base += 3; n = n - 3; } else { n = min(n, 3); }
    if (lane & 4) { base += 2; n = max(n - 2, 0); } else { n = min(n, 2); }
    if (lane & 2) { base += 1; n = max(n - 1, 0); } else { n = min(n, 1); }
    return n >= 1 ? base : -1;
}

__global__ void __launch_bounds__(kThreads)
composite_bwd_kernel(int H, int W, int gx, int ntiles, const uint32_t* __restrict__ header,
                     const uint32_t* __restrict__ work_order,
                     const uint32_t* __restrict__ tile_start, const GsrRec* __restrict__ records,
                     const float* __restrict__ bg, uint32_t* __restrict__ queue,
                     const float* __restrict__ out_depth_alpha,
                     const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dcolor,
                     const float* __restrict__ dL_ddepth_alpha, float* __restrict__ dgeom) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemRing& sm = *reinterpret_cast<SmemRing*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&sm.full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint32_t max_pairs = header[GSR_H_MAX_PAIRS];
    const uint32_t nonempty = header[GSR_H_NUM_NONEMPTY];
    const float bg0 = __ldg(bg), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
    const int vidx = bwd_value_index(lane);
    const bool commit_lane = (vidx >= 0) && !(lane & 1);
    uint32_t gc = 0;

    for (;;) {
        if (tid == 0) { sm.work = atomicAdd(queue, 1u); sm.maxlast = 0; }
        __syncthreads();
        const uint32_t w = sm.work;
        if (w >= nonempty) break;   // empty tiles have no gradient
        const uint32_t tile = work_order[w];
        uint32_t beg = tile_start[tile], end = tile_start[tile + 1];
        if (end > max_pairs) end = max_pairs;
        if (beg > end) beg = end;
        const int tyi = tile / gx, txi = tile - tyi * gx;
        const int X0i = txi * GSR_TILE + (wid & 1) * 8, Y0i = tyi * GSR_TILE + (wid >> 1) * 4;
        const int Xi = X0i + (lane & 7), Yi = Y0i + (lane >> 3);
        const float X0 = (float)X0i, Y0 = (float)Y0i, X = (float)Xi, Y = (float)Yi;
        const bool inside = Xi < W && Yi < H;
        const size_t pix = (size_t)Yi * W + Xi, plane = (size_t)H * W;
        uint32_t last = 0;
        float Tfinal = 1.f, dC0 = 0.f, dC1 = 0.f, dC2 = 0.f, dD = 0.f, dT = 0.f;
        if (inside) {
            last = n_contrib[pix];
            Tfinal = out_depth_alpha[plane + pix];
            dC0 = dL_dcolor[pix]; dC1 = dL_dcolor[plane + pix]; dC2 = dL_dcolor[2 * plane + pix];
            dD = dL_ddepth_alpha[pix]; dT = dL_ddepth_alpha[plane + pix];
        }
        if ((int)last > (int)(end - beg)) last = end - beg;   // overflow safety
        const uint32_t wmax = __reduce_max_sync(0xffffffffu, last);
        if (lane == 0 && wmax) atomicMax(&sm.maxlast, wmax);
        __syncthreads();
        const int n = (int)sm.maxlast;        // only entries [0, n) were ever blended
        __syncthreads();   // maxlast/work read by all before thread 0 resets them
        const int nchunks = (n + kChunk - 1) / kChunk;
        const float bgterm = bg0 * dC0 + bg1 * dC1 + bg2 * dC2 + dT;
        float T = Tfinal, accR = 0.f, accG = 0.f, accB = 0.f, accD = 0.f;

        // chunks are visited from the back: visit k <-> chunk index nchunks-1-k
        if (tid == 0) {
            const int npre = min(nchunks, kStages - 1);
            for (int k = 0; k < npre; ++k) {
                const int cidx = nchunks - 1 - k;
                const uint32_t slot = (gc + k) % kStages;
                const uint32_t cnt = (uint32_t)min(kChunk, n - cidx * kChunk);
                mbar_expect_tx(&sm.full[slot], cnt * (uint32_t)sizeof(GsrRec));
                bulk_g2s(sm.rec[slot], records + beg + (size_t)cidx * kChunk, cnt * (uint32_t)sizeof(GsrRec),
                         &sm.full[slot]);
            }
        }
        for (int k = 0; k < nchunks; ++k) {
            __syncthreads();   // visit k-1 fully consumed -> its slot may be refilled
            if (tid == 0 && k + kStages - 1 < nchunks) {
                const int kn = k + kStages - 1, cidx = nchunks - 1 - kn;
                const uint32_t slot = (gc + kn) % kStages;
                const uint32_t cnt = (uint32_t)min(kChunk, n - cidx * kChunk);
                mbar_expect_tx(&sm.full[slot], cnt * (uint32_t)sizeof(GsrRec));
                bulk_g2s(sm.rec[slot], records + beg + (size_t)cidx * kChunk, cnt * (uint32_t)sizeof(GsrRec),
                         &sm.full[slot]);
            }
            const uint32_t g = gc + k, slot = g % kStages;
            mbar_wait(&sm.full[slot], (g / kStages) & 1u);
            const GsrRec* st = sm.rec[slot];
            const int cidx = nchunks - 1 - k;
            const int cnt = min(kChunk, n - cidx * kChunk);
            const int cbase = cidx * kChunk;
            if ((int)wmax <= cbase) continue;   // nothing of this chunk reached this warp's pixels
            for (int sub = (cnt - 1) / 32; sub >= 0; --sub) {
                const int r = sub * 32 + lane;
                bool pass = false;
                if (r < cnt && (uint32_t)(cbase + r) < wmax) {
                    const float4 q0 = *reinterpret_cast<const float4*>(&st[r]);
                    pass = cull_pass(q0.x, q0.y, __float_as_uint(q0.z), X0, Y0);
                }
                uint32_t mask = __ballot_sync(0xffffffffu, pass);
                const float4* sp = reinterpret_cast<const float4*>(&st[sub * 32]);
                const uint32_t pos0 = (uint32_t)(cbase + sub * 32 + 1);
                while (mask) {
                    const int b = 31 - __clz(mask);
                    mask &= ~(1u << b);
                    const float4* rp = sp + 3 * b;
                    const float4 q0 = rp[0], q1 = rp[1], q2 = rp[2];
                    const uint32_t pos = pos0 + (uint32_t)b;
                    const PairEval e = eval_pair(q0.x, q0.y, q0.w, q1.x, q1.y, q1.z, X, Y);
                    const bool contrib = e.valid && pos <= last;
                    if (!__any_sync(0xffffffffu, contrib)) continue;
                    float v[10];
#pragma unroll
                    for (int j = 0; j < 10; ++j) v[j] = 0.f;
                    if (contrib) {
                        const float om = 1.0f - e.alpha;
                        const float rom = rcp_approx(om);
                        T = T * rom;                       // transmittance in front of this entry
                        const float wgt = e.alpha * T;
                        float dLda = (q2.x - accR) * dC0 + (q2.y - accG) * dC1 + (q2.z - accB) * dC2 +
                                     (q1.w - accD) * dD;
                        dLda = dLda * T - (Tfinal * rom) * bgterm;
                        accR = fmaf(e.alpha, q2.x - accR, accR);
                        accG = fmaf(e.alpha, q2.y - accG, accG);
                        accB = fmaf(e.alpha, q2.z - accB, accB);
                        accD = fmaf(e.alpha, q1.w - accD, accD);
                        const float gG = q1.z * dLda * e.G;   // dL/dG * G (no zeroing under the 0.99 clamp)
                        const float gxs = 2.0f * q0.w * e.dx + q1.x * e.dy;
                        const float gys = 2.0f * q1.y * e.dy + q1.x * e.dx;
                        v[0] = gG * gxs; v[1] = gG * gys;
                        v[2] = gG * e.dx * e.dx; v[3] = gG * e.dx * e.dy; v[4] = gG * e.dy * e.dy;
                        v[5] = e.G * dLda;
                        v[6] = wgt * dC0; v[7] = wgt * dC1; v[8] = wgt * dC2; v[9] = wgt * dD;
                    }
                    halve<10, 16>(v, lane & 16);
                    halve<5, 8>(v, lane & 8);
                    halve<3, 4>(v, lane & 4);
                    halve<2, 2>(v, lane & 2);
                    const float tot = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
                    if (commit_lane) atomicAdd(dgeom + 12 * (size_t)__float_as_uint(q2.w) + vidx, tot);
                }
            }
        }
        gc += (uint32_t)nchunks;
    }
}

}  // namespace

static int g_num_sms() {
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    return nsm;
}

cudaError_t gsr_launch_composite_fwd(const GsrFwdArgs& a) {
    const GsrTileGrid grid = gsr_grid(a.prm.image_height, a.prm.image_width);
    if (grid.ntiles == 0) return cudaSuccess;
    const uint32_t* header = reinterpret_cast<const uint32_t*>(a.saved + a.vl.header);
    const uint32_t* tile_start = reinterpret_cast<const uint32_t*>(a.saved + a.vl.tile_start);
    const uint32_t* work_order = reinterpret_cast<const uint32_t*>(a.saved + a.vl.work_order);
    const GsrRec* records = reinterpret_cast<const GsrRec*>(a.saved + a.vl.records);
    uint32_t* n_contrib = reinterpret_cast<uint32_t*>(a.saved + a.vl.n_contrib);
    uint32_t* queue = reinterpret_cast<uint32_t*>(a.scratch + a.sl.counters) + GSR_C_FWD_QUEUE;
    const int smem = (int)sizeof(SmemRing);
    const int nblocks = min(grid.ntiles, g_num_sms() * 6);
    cudaError_t e;
    if (a.prm.score_flag) {
        e = cudaFuncSetAttribute(composite_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        composite_fwd_kernel<true><<<nblocks, kThreads, smem, a.stream>>>(
            a.prm.image_height, a.prm.image_width, grid.gx, grid.ntiles, header, work_order, tile_start,
            records, a.prm.bg, queue, a.out_color, a.out_depth_alpha, n_contrib, a.score);
    } else {
        e = cudaFuncSetAttribute(composite_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        composite_fwd_kernel<false><<<nblocks, kThreads, smem, a.stream>>>(
            a.prm.image_height, a.prm.image_width, grid.gx, grid.ntiles, header, work_order, tile_start,
            records, a.prm.bg, queue, a.out_color, a.out_depth_alpha, n_contrib, a.score);
    }
    return cudaGetLastError();
}

cudaError_t gsr_launch_composite_bwd(const GsrBwdArgs& a) {
    const GsrTileGrid grid = gsr_grid(a.prm.image_height, a.prm.image_width);
    if (grid.ntiles == 0) return cudaSuccess;
    const uint32_t* header = reinterpret_cast<const uint32_t*>(a.saved + a.vl.header);
    const uint32_t* tile_start = reinterpret_cast<const uint32_t*>(a.saved + a.vl.tile_start);
    const uint32_t* work_order = reinterpret_cast<const uint32_t*>(a.saved + a.vl.work_order);
    const GsrRec* records = reinterpret_cast<const GsrRec*>(a.saved + a.vl.records);
    const uint32_t* n_contrib = reinterpret_cast<const uint32_t*>(a.saved + a.vl.n_contrib);
    uint32_t* queue = reinterpret_cast<uint32_t*>(a.scratch + a.sl.counters) + GSR_C_BWD_QUEUE;
    float* dgeom = reinterpret_cast<float*>(a.scratch + a.sl.dgeom);
    const int smem = (int)sizeof(SmemRing);
    const int nblocks = min(grid.ntiles, g_num_sms() * 4);
    cudaError_t e = cudaFuncSetAttribute(composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    composite_bwd_kernel<<<nblocks, kThreads, smem, a.stream>>>(
        a.prm.image_height, a.prm.image_width, grid.gx, grid.ntiles, header, work_order, tile_start, records,
        a.prm.bg, queue, a.out_depth_alpha, n_contrib, a.dL_dcolor, a.dL_ddepth_alpha, dgeom);
    return cudaGetLastError();
}
