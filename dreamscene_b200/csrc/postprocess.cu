// Depth/alpha -> normalised disparity (second half): the post-processing DreamScene runs
// after every render (/root/reference/scene_gaussian.py:871-881)
//     depth, alpha = chunk(depth_alpha, 2)
//     disp  = focal / (depth + alpha * 10 + 1e-5)
//     min_d = disp[alpha <= 0.1].min()        # falls back to disp.min() when no pixel is opaque
//     disp  = clamp((disp - min_d) / (disp.max() - min_d), 0, 1)
// as two small kernels forward and two backward, for a batch of views, with NO host synchronisation
// (the boolean-mask indexing above forces a device->host copy of the mask population per view).
// Gradients follow torch autograd exactly: through the quotient, through min_d and disp.max()
// (evenly distributed over ties, as torch.min()/max() do) and through the clamp (pass-through on
// the closed interval).
#include "common.cuh"

namespace {

// per-view reduction record (device memory, zero/identity-initialised by the first kernel's view-0 block? no:
// by a dedicated tiny init in the launcher via cudaMemsetAsync-free pattern: the reduce kernel for
// block 0 cannot know it runs first, so the record is initialised by reduce_init_kernel)
struct DispStats {
    unsigned int min_masked;   // float bits (disparities are positive: bit order == value order)
    unsigned int min_all;
    unsigned int max_all;
    unsigned int n_masked;
    // backward
    float sum_min, sum_max;    // sum_j dt_j (d_j - M)/(M-m)^2   and   sum_j -dt_j (d_j - m)/(M-m)^2
    unsigned int ties_min, ties_max;
};
static_assert(sizeof(DispStats) == 32, "stats record is 8 words");

__global__ void disp_init_kernel(DispStats* st, int B) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < B) {
        st[v].min_masked = 0x7f800000u; st[v].min_all = 0x7f800000u; st[v].max_all = 0u; st[v].n_masked = 0u;
        st[v].sum_min = 0.f; st[v].sum_max = 0.f; st[v].ties_min = 0u; st[v].ties_max = 0u;
    }
}

__device__ __forceinline__ float disp_raw(float depth, float alpha, float focal) {
    return __fdiv_rn(focal, __fadd_rn(__fadd_rn(depth, __fmul_rn(alpha, 10.0f)), 1e-5f));
}

// depth_alpha: [B][2][N]; grid (blocks, B)
__global__ void __launch_bounds__(256)
disp_reduce_kernel(const float* __restrict__ da, const float* __restrict__ focal, int N, DispStats* st) {
    const int v = blockIdx.y;
    const float* depth = da + (size_t)v * 2 * N;
    const float* alpha = depth + N;
    const float f = focal[v];
    unsigned int mm = 0x7f800000u, ma = 0x7f800000u, mx = 0u, nm = 0u;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256) {
        const float a = alpha[i];
        const unsigned int d = __float_as_uint(disp_raw(depth[i], a, f));
        ma = min(ma, d); mx = max(mx, d);
        if (a <= 0.1f) { mm = min(mm, d); ++nm; }
    }
    mm = __reduce_min_sync(0xffffffffu, mm); ma = __reduce_min_sync(0xffffffffu, ma);
    mx = __reduce_max_sync(0xffffffffu, mx); nm = __reduce_add_sync(0xffffffffu, nm);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&st[v].min_masked, mm); atomicMin(&st[v].min_all, ma); atomicMax(&st[v].max_all, mx);
        if (nm) atomicAdd(&st[v].n_masked, nm);
    }
}

__device__ __forceinline__ void disp_bounds(const DispStats& s, float& m, float& M) {
    m = __uint_as_float(s.n_masked > 0 ? s.min_masked : s.min_all);
    M = __uint_as_float(s.max_all);
}

__global__ void __launch_bounds__(256)
disp_normalise_kernel(const float* __restrict__ da, const float* __restrict__ focal, int N,
                      const DispStats* __restrict__ st, float* __restrict__ out) {
    const int v = blockIdx.y;
    const float* depth = da + (size_t)v * 2 * N;
    const float* alpha = depth + N;
    float m, M;
    disp_bounds(st[v], m, M);
    const float f = focal[v];
    const float den = __fsub_rn(M, m);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256) {
        const float d = disp_raw(depth[i], alpha[i], f);
        const float t = __fdiv_rn(__fsub_rn(d, m), den);
        out[(size_t)v * N + i] = fminf(fmaxf(t, 0.0f), 1.0f);       // NaN (den == 0) propagates like torch.clamp
    }
}

// backward pass 1: the two global sums feeding the min / max paths, and the tie counts
__global__ void __launch_bounds__(256)
disp_bwd_reduce_kernel(const float* __restrict__ da, const float* __restrict__ focal, int N,
                       const float* __restrict__ g_out, DispStats* st) {
    const int v = blockIdx.y;
    const float* depth = da + (size_t)v * 2 * N;
    const float* alpha = depth + N;
    float m, M;
    disp_bounds(st[v], m, M);
    const bool masked_min = st[v].n_masked > 0;
    const float f = focal[v];
    const float den = M - m, inv2 = 1.0f / (den * den);
    float sm = 0.f, sM = 0.f;
    unsigned int tm = 0u, tM = 0u;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256) {
        const float a = alpha[i];
        const float d = disp_raw(depth[i], a, f);
        const float t = (d - m) / den;
        const float dt = (t >= 0.0f && t <= 1.0f) ? g_out[(size_t)v * N + i] : 0.0f;
        sm += dt * (d - M) * inv2;
        sM -= dt * (d - m) * inv2;
        if (d == m && (!masked_min || a <= 0.1f)) ++tm;
        if (d == M) ++tM;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        sM += __shfl_xor_sync(0xffffffffu, sM, o);
    }
    tm = __reduce_add_sync(0xffffffffu, tm); tM = __reduce_add_sync(0xffffffffu, tM);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&st[v].sum_min, sm); atomicAdd(&st[v].sum_max, sM);
        if (tm) atomicAdd(&st[v].ties_min, tm);
        if (tM) atomicAdd(&st[v].ties_max, tM);
    }
}

// backward pass 2: d depth_alpha = chain through d = focal / x, x = depth + 10 alpha + 1e-5 (+ the caller's own
// gradient on the alpha channel, which is returned alongside the disparity)
__global__ void __launch_bounds__(256)
disp_bwd_apply_kernel(const float* __restrict__ da, const float* __restrict__ focal, int N,
                      const float* __restrict__ g_out, const float* __restrict__ g_alpha,
                      const DispStats* __restrict__ st, float* __restrict__ d_da) {
    const int v = blockIdx.y;
    const float* depth = da + (size_t)v * 2 * N;
    const float* alpha = depth + N;
    float m, M;
    disp_bounds(st[v], m, M);
    const bool masked_min = st[v].n_masked > 0;
    const float f = focal[v];
    const float den = M - m, inv = 1.0f / den;
    const float gmin = st[v].ties_min ? st[v].sum_min / (float)st[v].ties_min : 0.0f;
    const float gmax = st[v].ties_max ? st[v].sum_max / (float)st[v].ties_max : 0.0f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256) {
        const float a = alpha[i];
        const float d = disp_raw(depth[i], a, f);
        const float t = (d - m) * inv;
        float gd = (t >= 0.0f && t <= 1.0f) ? g_out[(size_t)v * N + i] * inv : 0.0f;
        if (d == m && (!masked_min || a <= 0.1f)) gd += gmin;
        if (d == M) gd += gmax;
        const float gx = -gd * d * d / f;                     // d(focal/x)/dx = -d^2/focal
        d_da[(size_t)v * 2 * N + i] = gx;
        d_da[(size_t)v * 2 * N + N + i] = 10.0f * gx + (g_alpha ? g_alpha[(size_t)v * N + i] : 0.0f);
    }
}

}  // namespace

cudaError_t gsr_launch_disparity_fwd(int B, int N, const float* da, const float* focal, float* out, void* stats,
                                     int num_sms, cudaStream_t s) {
    if (B <= 0 || N <= 0) return cudaSuccess;
    DispStats* st = static_cast<DispStats*>(stats);
    disp_init_kernel<<<(B + 63) / 64, 64, 0, s>>>(st, B);
    const int bx = max(1, min((N + 255) / 256, (num_sms * 8 + B - 1) / B));
    disp_reduce_kernel<<<dim3(bx, B), 256, 0, s>>>(da, focal, N, st);
    disp_normalise_kernel<<<dim3(bx, B), 256, 0, s>>>(da, focal, N, st, out);
    return cudaGetLastError();
}

cudaError_t gsr_launch_disparity_bwd(int B, int N, const float* da, const float* focal, const float* g_out,
                                     const float* g_alpha, void* stats, float* d_da, int num_sms, cudaStream_t s) {
    if (B <= 0 || N <= 0) return cudaSuccess;
    DispStats* st = static_cast<DispStats*>(stats);
    const int bx = max(1, min((N + 255) / 256, (num_sms * 8 + B - 1) / B));
    disp_bwd_reduce_kernel<<<dim3(bx, B), 256, 0, s>>>(da, focal, N, g_out, st);
    disp_bwd_apply_kernel<<<dim3(bx, B), 256, 0, s>>>(da, focal, N, g_out, g_alpha, st, d_da);
    return cudaGetLastError();
}
