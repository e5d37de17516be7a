// grad_z of the quantiser for token-major rows, as a device function working on a virtual (block, grid): it is the
// whole of k_backward_tokens, the tail blocks of k_backward_fused (one GPU) and of k_codebook_grad_sharded (N GPUs).
// Closed form (the reference has no backward code, autograd derives it from
// models/vitvqgan.py:151-171 / models/vqgan.py:148-176):
//   g_zn = G + g_loss * c1 * 2 (zn - q) / N ;  grad_z = (g_zn - zn (zn . g_zn)) / max(||z||, eps)
// kLpr lanes share a row (one float4 each, kNf4 float4 per lane when D > 128); the row dot product is an xor-shuffle
// tree inside the lane group.  Algorithmic bytes 12D + 12 per token.
#pragma once

#include "vq_common.cuh"

namespace vq {

template <int D>
__device__ __forceinline__ void backward_tokens_body(const float4* __restrict__ g, const float4* __restrict__ zn,
                                                     const float* __restrict__ denom, const int64_t* __restrict__ idx,
                                                     const float4* __restrict__ en, int64_t T, float coef_base,
                                                     const float* __restrict__ g_loss, float4* __restrict__ grad, int vblock,
                                                     int vgrid) {
    const float coef = coef_base * (g_loss ? __ldg(g_loss) : 1.f);
    constexpr int kChunks = D / 4;
    constexpr int kLpr = (kChunks < 32) ? kChunks : 32;
    constexpr int kNf4 = kChunks / kLpr;
    constexpr int kRowsPerWarp = 32 / kLpr;
    const int lane = threadIdx.x & 31;
    const int sub = lane % kLpr, grp = lane / kLpr;
    const int64_t warp = (int64_t)vblock * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)vgrid * (blockDim.x >> 5);
    for (int64_t t0 = warp * kRowsPerWarp; t0 < T; t0 += n_warps * kRowsPerWarp) {
        const int64_t t = t0 + grp;
        const bool live = t < T;
        float4 a[kNf4], gz[kNf4];
        float dot = 0.f;
        float dn = 1.f;
        if (live) {
            const int64_t k = __ldg(idx + t);
            dn = __ldg(denom + t);
#pragma unroll
            for (int f = 0; f < kNf4; ++f) {
                const int c = sub + kLpr * f;
                a[f] = __ldg(zn + t * kChunks + c);
                const float4 q = __ldg(en + k * kChunks + c);
                float4 up = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g) up = __ldcs(g + t * kChunks + c);
                // explicit roundings: the same bits from every kernel this body is inlined into (grad_z is compared
                // bit for bit between the fused, the stand-alone and the chunked host paths)
                gz[f].x = __fmaf_rn(coef, __fsub_rn(a[f].x, q.x), up.x);
                gz[f].y = __fmaf_rn(coef, __fsub_rn(a[f].y, q.y), up.y);
                gz[f].z = __fmaf_rn(coef, __fsub_rn(a[f].z, q.z), up.z);
                gz[f].w = __fmaf_rn(coef, __fsub_rn(a[f].w, q.w), up.w);
                dot = __fadd_rn(dot, __fadd_rn(__fmaf_rn(a[f].x, gz[f].x, __fmul_rn(a[f].y, gz[f].y)),
                                               __fmaf_rn(a[f].z, gz[f].z, __fmul_rn(a[f].w, gz[f].w))));
            }
        }
#pragma unroll
        for (int off = kLpr >> 1; off > 0; off >>= 1) dot = __fadd_rn(dot, __shfl_xor_sync(VQ_FULL, dot, off));
        if (live) {
            const float inv = __fdiv_rn(1.f, dn);
#pragma unroll
            for (int f = 0; f < kNf4; ++f) {
                float4 o;
                o.x = __fmul_rn(__fmaf_rn(-a[f].x, dot, gz[f].x), inv);
                o.y = __fmul_rn(__fmaf_rn(-a[f].y, dot, gz[f].y), inv);
                o.z = __fmul_rn(__fmaf_rn(-a[f].z, dot, gz[f].z), inv);
                o.w = __fmul_rn(__fmaf_rn(-a[f].w, dot, gz[f].w), inv);
                __stcs(grad + t * kChunks + sub + kLpr * f, o);
            }
        }
    }
}

}  // namespace vq
