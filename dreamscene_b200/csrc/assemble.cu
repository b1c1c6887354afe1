// Scene assembly: the per-render glue DreamScene runs in PyTorch before every
// rasterizer call, as ONE forward and ONE backward kernel.
//
// Reference (/root/reference/scene_gaussian.py:753-857, gs_renderer.py:464-488): for every visible
// group (environment, floor, object instances) take the raw leaf parameters, apply the activations
//     opacity = sigmoid(_opacity)   scales = exp(_scaling)   rotations = normalize(_rotation)
//     shs = cat(_features_dc, _features_rest, dim=1)         means3D = _xyz
// torch.cat the groups, then augment (scene_gaussian.py:848-856)
//     shs    = shs + randn_like(shs) * (0.2**0.5 * shs)
//     scales = clamp(scales + randn_like(scales) * (0.2**0.5 * scales / 4), 0)
// That is ~6 activation kernels + 5 cats per group list + 6 augmentation kernels per view, each a
// full pass over 2.6M-row tensors, and the same again (plus autograd's cat/split bookkeeping) in the
// backward.  Here: one pass that reads every raw parameter once and writes the five packed,
// rasterizer-ready arrays; one pass back that turns their gradients into the per-group leaf gradients.
//
// Numerics: every expression is evaluated with the reference's operation order and individually
// rounded fp32 ops (expf, IEEE division), so with the noise tensors passed in the outputs are
// bit-identical to the PyTorch expressions except `normalize` (sum-of-squares order, <= 1 ulp).
// Noise: either the caller's standard-normal tensors (z_shs[P,M,3], z_scales[P,3]: generate them with
// torch.randn in the reference's order and the random stream matches the reference's), or - z == NULL
// and seed given - a counter-based Philox4x32-10 generator evaluated in the kernel (no noise tensor
// is ever written or read; the backward regenerates the same numbers from (seed, element index)).
#include "common.cuh"
#include <cstring>

namespace {

constexpr int kMaxGroups = B200GSR_MAX_GROUPS;

struct GroupTable {
    const float* xyz[kMaxGroups];
    const float* opacity[kMaxGroups];
    const float* scaling[kMaxGroups];
    const float* rotation[kMaxGroups];
    const float* f_dc[kMaxGroups];
    const float* f_rest[kMaxGroups];
    int start[kMaxGroups + 1];     // first packed row of group g; start[num] = P
    int num;
};
struct GroupGradTable {
    float* xyz[kMaxGroups];
    float* opacity[kMaxGroups];
    float* scaling[kMaxGroups];
    float* rotation[kMaxGroups];
    float* f_dc[kMaxGroups];
    float* f_rest[kMaxGroups];
};

__device__ __forceinline__ int find_group(const GroupTable& t, int row) {
    int g = 0;
#pragma unroll 1
    while (g + 1 < t.num && row >= t.start[g + 1]) ++g;
    return g;
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (lo, hi, stream, 0), key = seed ------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
// four standard normals for quad index q of stream s (Box-Muller on the four 32-bit outputs)
__device__ __forceinline__ float4 normal4(unsigned long long seed, uint32_t stream, unsigned long long q) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), stream, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float u0 = ((float)r.x + 0.5f) * 2.3283064365386963e-10f;   // (0, 1)
    const float u1 = ((float)r.y + 0.5f) * 2.3283064365386963e-10f;
    const float u2 = ((float)r.z + 0.5f) * 2.3283064365386963e-10f;
    const float u3 = ((float)r.w + 0.5f) * 2.3283064365386963e-10f;
    const float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u1, &s0, &c0);
    __sincosf(6.283185307179586f * u3, &s1, &c1);
    return make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
}
enum { kStreamShs = 1u, kStreamScales = 2u };    // + 2 * view

// noise factor helpers: value v, standard normal z, coefficient c (0.2**0.5), divisor d (1 or 4):
// reference order  v + z * ((c * v) / d)
__device__ __forceinline__ float aug(float v, float z, float c, float d) {
    return __fadd_rn(v, __fmul_rn(z, __fdiv_rn(__fmul_rn(c, v), d)));
}

constexpr int kAsmBlock = 128;

// =============================================================================================
// forward: block = 128 packed rows; phase A one thread per Gaussian (11 floats), phase B the
// block's SH rows as a flat span of floats (coalesced stores; float4 when 3M % 4 == 0)
// =============================================================================================
// B = number of views: `scales` is [B][P,3] and `shs` [B][P,M,3] (one independently augmented copy per view, the
// raw parameters are read ONCE), means3D / opac / rots are written once; in the backward the per-view gradients of
// scales / shs are summed over the views in registers and every leaf gradient is written once.  View v uses the
// noise rows z[v] (caller's draws) or the Philox streams 2v+1 / 2v+2.
template <bool BACKWARD>
__global__ void __launch_bounds__(kAsmBlock)
assemble_kernel(GroupTable tab, GroupGradTable gtab, int P, int M, int B, float c_shs, float c_scale,
                const float* __restrict__ z_shs, const float* __restrict__ z_scales, unsigned long long seed,
                // forward outputs / backward incoming gradients (packed)
                float* __restrict__ means3D, float* __restrict__ opac, float* __restrict__ scales,
                float* __restrict__ rots, float* __restrict__ shs) {
    const int r0 = blockIdx.x * kAsmBlock;
    const int i = r0 + threadIdx.x;
    // ---- phase A ---------------------------------------------------------------------------
    if (i < P) {
        const int g = find_group(tab, i);
        const int l = i - tab.start[g];
        const float sx = tab.scaling[g][3 * (size_t)l], sy = tab.scaling[g][3 * (size_t)l + 1], sz = tab.scaling[g][3 * (size_t)l + 2];
        const float4 q = *reinterpret_cast<const float4*>(tab.rotation[g] + 4 * (size_t)l);
        const float o = tab.opacity[g][l];
        const float e[3] = {expf(sx), expf(sy), expf(sz)};
        const float sig = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-o)));
        const float nrm = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y)), __fmul_rn(q.z, q.z)), __fmul_rn(q.w, q.w))), 1e-12f);
        float dsum[3] = {0.f, 0.f, 0.f};
        for (int v = 0; v < B; ++v) {
            float zs[3] = {0.f, 0.f, 0.f};
            if (c_scale != 0.0f) {
                if (z_scales != nullptr) {
                    const float* zr = z_scales + ((size_t)v * P + i) * 3;
                    zs[0] = zr[0]; zs[1] = zr[1]; zs[2] = zr[2];
                } else {
                    const float4 n = normal4(seed, kStreamScales + 2u * (uint32_t)v, (unsigned long long)i);
                    zs[0] = n.x; zs[1] = n.y; zs[2] = n.z;
                }
            }
            float* sv = scales + ((size_t)v * P + i) * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (!BACKWARD) {
                    sv[k] = (c_scale != 0.0f) ? fmaxf(aug(e[k], zs[k], c_scale, 4.0f), 0.0f) : e[k];
                } else {
                    float d = sv[k];
                    if (c_scale != 0.0f) {
                        // y = clamp(e + z*((c*e)/4), 0): dy/de = 1 + z*c/4 where the clamp is inactive
                        const float y = aug(e[k], zs[k], c_scale, 4.0f);
                        d = (y > 0.0f) ? d * (1.0f + zs[k] * (c_scale * 0.25f)) : 0.0f;
                    }
                    dsum[k] += d;
                }
            }
        }
        if (!BACKWARD) {
            means3D[3 * (size_t)i] = tab.xyz[g][3 * (size_t)l];
            means3D[3 * (size_t)i + 1] = tab.xyz[g][3 * (size_t)l + 1];
            means3D[3 * (size_t)i + 2] = tab.xyz[g][3 * (size_t)l + 2];
            opac[i] = sig;
            *reinterpret_cast<float4*>(rots + 4 * (size_t)i) =
                make_float4(__fdiv_rn(q.x, nrm), __fdiv_rn(q.y, nrm), __fdiv_rn(q.z, nrm), __fdiv_rn(q.w, nrm));
        } else {
            // incoming gradients live in the packed arrays; outputs are the per-group leaf gradients
            gtab.xyz[g][3 * (size_t)l] = means3D[3 * (size_t)i];
            gtab.xyz[g][3 * (size_t)l + 1] = means3D[3 * (size_t)i + 1];
            gtab.xyz[g][3 * (size_t)l + 2] = means3D[3 * (size_t)i + 2];
            gtab.opacity[g][l] = opac[i] * sig * (1.0f - sig);
#pragma unroll
            for (int k = 0; k < 3; ++k) gtab.scaling[g][3 * (size_t)l + k] = dsum[k] * e[k];
            const float4 gq = *reinterpret_cast<const float4*>(rots + 4 * (size_t)i);
            const float inv = 1.0f / nrm;
            const float4 u = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
            const float dot = u.x * gq.x + u.y * gq.y + u.z * gq.z + u.w * gq.w;
            // d normalize: (g - u (u.g)) / |q|   (the eps clamp is inactive for any usable quaternion)
            *reinterpret_cast<float4*>(gtab.rotation[g] + 4 * (size_t)l) =
                make_float4((gq.x - u.x * dot) * inv, (gq.y - u.y * dot) * inv, (gq.z - u.z * dot) * inv, (gq.w - u.w * dot) * inv);
        }
    }
    // ---- phase B: SH rows of the block as a flat span ----------------------------------------
    const int rows = min(kAsmBlock, P - r0);
    if (rows <= 0) return;
    const int row_f = 3 * M;
    const size_t base = (size_t)r0 * row_f;          // first float of the block's span in one view's packed array
    const size_t view_f = (size_t)P * row_f;         // floats per view
    const int total = rows * row_f;
    const bool vec = (row_f & 3) == 0;
    const int step = vec ? 4 : 1;
    for (int f = threadIdx.x * step; f < total; f += kAsmBlock * step) {
        const int row = f / row_f, col = f - row * row_f;       // vec: all 4 elements share the row
        const int gi = r0 + row;
        const int g = find_group(tab, gi);
        const size_t l = (size_t)(gi - tab.start[g]);
        float raw[4] = {0.f, 0.f, 0.f, 0.f}, dsum[4] = {0.f, 0.f, 0.f, 0.f};
        if (!BACKWARD) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k >= step) break;
                const int c = col + k;
                raw[k] = (c < 3) ? tab.f_dc[g][3 * l + c] : tab.f_rest[g][(size_t)(row_f - 3) * l + (c - 3)];
            }
        }
        for (int v = 0; v < B; ++v) {
            float z[4] = {0.f, 0.f, 0.f, 0.f};
            const size_t e0 = base + f;                          // element index inside the view
            if (c_shs != 0.0f) {
                if (z_shs != nullptr) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (k < step) z[k] = z_shs[(size_t)v * view_f + e0 + k];
                } else {
                    const float4 n = normal4(seed, kStreamShs + 2u * (uint32_t)v, (unsigned long long)(e0 >> 2));
                    if (vec) { z[0] = n.x; z[1] = n.y; z[2] = n.z; z[3] = n.w; }
                    else z[0] = (e0 & 3) == 0 ? n.x : (e0 & 3) == 1 ? n.y : (e0 & 3) == 2 ? n.z : n.w;
                }
            }
            float* sv = shs + (size_t)v * view_f + e0;
            if (!BACKWARD) {
                float out[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) out[k] = (c_shs != 0.0f) ? aug(raw[k], z[k], c_shs, 1.0f) : raw[k];
                if (vec) *reinterpret_cast<float4*>(sv) = make_float4(out[0], out[1], out[2], out[3]);
                else sv[0] = out[0];
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k >= step) break;
                    float d = sv[k];
                    if (c_shs != 0.0f) d *= (1.0f + z[k] * c_shs);           // d(v + z*(c*v))/dv
                    dsum[k] += d;
                }
            }
        }
        if (BACKWARD) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k >= step) break;
                const int c = col + k;
                if (c < 3) gtab.f_dc[g][3 * l + c] = dsum[k];
                else gtab.f_rest[g][(size_t)(row_f - 3) * l + (c - 3)] = dsum[k];
            }
        }
    }
}

}  // namespace

cudaError_t gsr_launch_assemble(bool backward, int num_groups, const b200gsr_group* groups,
                                const b200gsr_group_grad* grads, int M, int B, float c_shs, float c_scale,
                                const float* z_shs, const float* z_scales, unsigned long long seed,
                                float* means3D, float* opac, float* scales, float* rots, float* shs, cudaStream_t s) {
    GroupTable tab;
    GroupGradTable gtab;
    memset(&tab, 0, sizeof(tab));
    memset(&gtab, 0, sizeof(gtab));
    int P = 0;
    for (int g = 0; g < num_groups; ++g) {
        tab.xyz[g] = groups[g].xyz; tab.opacity[g] = groups[g].opacity; tab.scaling[g] = groups[g].scaling;
        tab.rotation[g] = groups[g].rotation; tab.f_dc[g] = groups[g].f_dc; tab.f_rest[g] = groups[g].f_rest;
        tab.start[g] = P;
        P += groups[g].n;
        if (backward) {
            gtab.xyz[g] = grads[g].xyz; gtab.opacity[g] = grads[g].opacity; gtab.scaling[g] = grads[g].scaling;
            gtab.rotation[g] = grads[g].rotation; gtab.f_dc[g] = grads[g].f_dc; gtab.f_rest[g] = grads[g].f_rest;
        }
    }
    tab.start[num_groups] = P;
    tab.num = num_groups;
    if (P == 0) return cudaSuccess;
    const int nblocks = (P + kAsmBlock - 1) / kAsmBlock;
    if (backward)
        assemble_kernel<true><<<nblocks, kAsmBlock, 0, s>>>(tab, gtab, P, M, B, c_shs, c_scale, z_shs, z_scales, seed,
                                                            means3D, opac, scales, rots, shs);
    else
        assemble_kernel<false><<<nblocks, kAsmBlock, 0, s>>>(tab, gtab, P, M, B, c_shs, c_scale, z_shs, z_scales, seed,
                                                             means3D, opac, scales, rots, shs);
    return cudaGetLastError();
}
