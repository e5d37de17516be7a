// Fused gather / straight-through output / loss partial / code-usage histogram, and the decode gather.
//
// Reference call sites replaced (paths relative to /root/reference):
//   models/vitvqgan.py:162-163  z_q = l2_norm(embedding(idx))       models/vqgan.py:163-167
//   models/vitvqgan.py:166      loss (two means over all T*D elements) models/vqgan.py:169
//   models/vitvqgan.py:169      z_q = z + (z_q - z).detach()          models/vqgan.py:171
//   models/vitvqgan.py:173-176  indices_to_embeddings                 models/vqgan.py:178-182
// l2_norm(E[idx]) is bit-identical to the prepared unit code en[idx] (same row, same ATen schedule),
// so the gather reads en32 directly.  HBM-bound and purely element-wise, fully coalesced; algorithmic bytes
// 8D + 8 per token.  When the step trains the codebook the same pass also accumulates the
// codebook-gradient segment sums (they only need q - zn, which is in registers here).
#include "vq_common.cuh"
#include "vq_kernels.h"
#include "../../include/vq_b200.h"

namespace vq {

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* smem) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(VQ_FULL, v, off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    unsigned long long t = 0;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : 0ull;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(VQ_FULL, t, off);
    }
    return t;   // valid in thread 0
}

// A warp walks the (T x D) array in spans of 128 consecutive floats; lane l owns elements l, l + 32, l + 64, l + 96
// of a span, so every load / store instruction of the warp is one coalesced 128-byte access and every segment-sum
// RED instruction covers 256 contiguous bytes of int64 (whole 32-byte sectors).  kSpans spans are in flight per warp.
template <int kSpans>
__global__ void __launch_bounds__(256) k_finish(const float* __restrict__ zn, const int* __restrict__ cand,
                                                const float* __restrict__ en, int64_t T, int D, int log2d, int K,
                                                float* __restrict__ zq, void* __restrict__ idx_out, int idx_bits,
                                                int32_t* __restrict__ hist, unsigned long long* __restrict__ seg,
                                                int64_t* __restrict__ stats) {
    __shared__ unsigned long long red[8];
    long long loss_fx = 0;
    unsigned long long bad = 0;
    if (!zq) {   // indices only: one thread per row
        for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x) {
            const int k = __ldg(cand + t) & (kCandExactBit - 1);
            store_token(idx_out, t, k, idx_bits);
            if (hist) atomicAdd(hist + k, 1);
        }
        return;
    }
    const int lane = threadIdx.x & 31;
    const int64_t total = T << log2d;
    const int64_t n_spans = (total + 127) >> 7;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t sp = warp; sp < n_spans; sp += n_warps * kSpans) {
        int k[kSpans][4];
        float a[kSpans][4], q[kSpans][4];
#pragma unroll
        for (int j = 0; j < kSpans; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t e = ((sp + j * n_warps) << 7) + lane + 32 * i;
                k[j][i] = (e < total) ? (__ldg(cand + (e >> log2d)) & (kCandExactBit - 1)) : 0;
            }
#pragma unroll
        for (int j = 0; j < kSpans; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t e = ((sp + j * n_warps) << 7) + lane + 32 * i;
                a[j][i] = 0.f; q[j][i] = 0.f;
                if (e < total) {
                    a[j][i] = __ldcs(zn + e);
                    q[j][i] = __ldg(en + ((int64_t)k[j][i] << log2d) + (e & (D - 1)));
                }
            }
#pragma unroll
        for (int j = 0; j < kSpans; ++j) {
            float df[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t e = ((sp + j * n_warps) << 7) + lane + 32 * i;
                df[i] = 0.f;
                if (e < total) {
                    const int c = (int)(e & (D - 1));
                    if (c == 0) {
                        store_token(idx_out, e >> log2d, k[j][i], idx_bits);
                        if (hist) atomicAdd(hist + k[j][i], 1);
                    }
                    df[i] = __fsub_rn(q[j][i], a[j][i]);
                    __stcs(zq + e, __fadd_rn(a[j][i], df[i]));
                    if (seg) {
                        unsigned poison = 0;
                        seg_add(seg + ((int64_t)k[j][i] << log2d) + c, df[i], poison);
                        if (poison) atomicAdd(seg + ((int64_t)K << log2d) + k[j][i], 1ull);
                    }
                }
            }
            const float p = (df[0] * df[0] + df[1] * df[1]) + (df[2] * df[2] + df[3] * df[3]);
            if (is_finite(p)) loss_fx += to_fixed(p, VQ_LOSS_SHIFT);
            else bad += 1;
        }
    }
    if (stats) {
        const unsigned long long s1 = block_sum_u64((unsigned long long)loss_fx, red);
        __syncthreads();
        const unsigned long long s2 = block_sum_u64(bad, red);
        if (threadIdx.x == 0) {
            if (s1) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_LOSS_FIXED), s1);
            if (s2) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_NONFINITE), s2);
        }
    }
}

// The same pass with z_q written straight into the reference's (b, D, h, w) output (models/vqgan.py:174): 32-token x
// 32-channel tiles, reads coalesced along the channels, z_q transposed through shared memory and stored coalesced along
// the pixels -- no token-major staging buffer and no separate layout kernel.  Segment-sum REDs cover 256 contiguous bytes.
__global__ void __launch_bounds__(256) k_finish_nchw(const float* __restrict__ zn, const int* __restrict__ cand,
                                                     const float* __restrict__ en, int64_t T, int64_t hw, int D, int K,
                                                     float* __restrict__ zq_nchw, void* __restrict__ idx_out, int idx_bits,
                                                     int32_t* __restrict__ hist, unsigned long long* __restrict__ seg,
                                                     int64_t* __restrict__ stats) {
    __shared__ float tile[32][33];
    __shared__ unsigned long long red[8];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int c = c0 + x;
    float df[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int tt = y + 8 * i;
        const int64_t t = t0 + tt;
        if (t < T && c < D) {
            const int k = __ldg(cand + t) & (kCandExactBit - 1);
            const float a = __ldcs(zn + t * D + c), q = __ldg(en + (int64_t)k * D + c);
            df[i] = __fsub_rn(q, a);
            tile[tt][x] = __fadd_rn(a, df[i]);
            if (c == 0) {
                store_token(idx_out, t, k, idx_bits);
                if (hist) atomicAdd(hist + k, 1);
            }
            if (seg) {
                unsigned poison = 0;
                seg_add(seg + (int64_t)k * D + c, df[i], poison);
                if (poison) atomicAdd(seg + (int64_t)K * D + k, 1ull);
            }
        }
    }
    long long loss_fx = 0;
    unsigned long long bad = 0;
    const float p = (df[0] * df[0] + df[1] * df[1]) + (df[2] * df[2] + df[3] * df[3]);
    if (is_finite(p)) loss_fx = to_fixed(p, VQ_LOSS_SHIFT);
    else bad = 1;
    __syncthreads();
    {
        const int64_t t = t0 + x;
        if (t < T) {
            const int64_t b = t / hw, pix = t % hw;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int cc = c0 + y + 8 * i;
                if (cc < D) __stcs(zq_nchw + (b * D + cc) * hw + pix, tile[x][y + 8 * i]);
            }
        }
    }
    if (stats) {
        const unsigned long long s1 = block_sum_u64((unsigned long long)loss_fx, red);
        __syncthreads();
        const unsigned long long s2 = block_sum_u64(bad, red);
        if (threadIdx.x == 0) {
            if (s1) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_LOSS_FIXED), s1);
            if (s2) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_NONFINITE), s2);
        }
    }
}

cudaError_t launch_finish_nchw(const float* zn32, const int* cand, const CodebookView& cb, int64_t T, int64_t hw,
                               float* zq_nchw, void* idx_out, int32_t* hist, int64_t* seg_sums, int64_t* stats,
                               cudaStream_t s, int idx_bits) {
    if (T == 0) return cudaSuccess;
    dim3 grid((unsigned)((T + 31) / 32), (unsigned)((cb.D + 31) / 32));
    k_finish_nchw<<<grid, 256, 0, s>>>(zn32, cand, cb.en32, T, hw, cb.D, cb.K, zq_nchw, idx_out, idx_bits, hist,
                                       reinterpret_cast<unsigned long long*>(seg_sums), stats);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_finish(const float* zn32, const int* cand, const CodebookView& cb, int64_t T, float* zq_tok,
                          void* idx_out, int32_t* hist, int64_t* seg_sums, int64_t* stats, cudaStream_t s, int idx_bits) {
    if (T == 0) return cudaSuccess;
    constexpr int kSpans = 2;
    int log2d = 0;
    while ((1 << log2d) < cb.D) ++log2d;
    // z_q pass: a warp per kSpans spans of 128 floats; indices only: one thread per row
    const int64_t work = zq_tok ? ((T * cb.D + 127) / 128 + kSpans - 1) / kSpans : (T + 31) / 32;
    int64_t blocks = (work + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k_finish<kSpans><<<(unsigned)blocks, 256, 0, s>>>(zn32, cand, cb.en32, T, cb.D, log2d, cb.K, zq_tok, idx_out, idx_bits, hist,
                                                      zq_tok ? reinterpret_cast<unsigned long long*>(seg_sums) : nullptr, stats);
    count_launch();
    return cudaGetLastError();
}

// loss = beta*m + m (ViT) or m + beta*m (VQGAN), m = mean((q - zn)^2): both reference terms have the
// same value, only their gradients differ.  stats[NONFINITE] counts non-finite partials -> NaN.
__global__ void k_loss_finalize(const int64_t* __restrict__ stats, int64_t n_elem_total, int form, float beta,
                                float* __restrict__ loss) {
    loss[0] = loss_from_fixed(stats[VQ_STAT_LOSS_FIXED], stats[VQ_STAT_NONFINITE], n_elem_total, form, beta);
}

cudaError_t launch_loss_finalize(const int64_t* stats, int64_t n_elem_total, int form, float beta, float* loss,
                                 cudaStream_t s) {
    k_loss_finalize<<<1, 1, 0, s>>>(stats, n_elem_total, form, beta, loss);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// decode gather.  Token-major out: element-wise float4.  NCHW out: 32-token x 32-channel tiles
// through shared memory so both the row reads and the (b, D, hw) writes are coalesced.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_tok(const void* __restrict__ idx, int idx_bits, int64_t T, int chunks, int K,
                                                    const float4* __restrict__ table, float4* __restrict__ out,
                                                    int64_t* __restrict__ stats) {
    const int64_t total = T * chunks;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = e / chunks;
        const int c = (int)(e - t * chunks);
        const int64_t k = load_token(idx, t, idx_bits);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k >= 0 && k < K) v = __ldg(table + k * chunks + c);
        else if (c == 0 && stats) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_BAD_INDEX), 1ull);
        __stcs(out + e, v);
    }
}

__global__ void __launch_bounds__(256) k_gather_nchw(const void* __restrict__ idx, int idx_bits, int64_t T, int64_t hw, int D,
                                                     int K, const float* __restrict__ table, float* __restrict__ out,
                                                     int64_t* __restrict__ stats) {
    __shared__ float tile[32][33];
    __shared__ int64_t rows[32];
    const int x = threadIdx.x, y = threadIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    if (y == 0) {
        const int64_t t = t0 + x;
        int64_t k = (t < T) ? load_token(idx, t, idx_bits) : 0;
        if (k < 0 || k >= K) {
            if (blockIdx.y == 0 && stats) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_BAD_INDEX), 1ull);
            k = -1;
        }
        rows[x] = k;
    }
    __syncthreads();
    for (int tt = y; tt < 32; tt += 8) {
        const int64_t k = rows[tt];
        const int c = c0 + x;
        tile[tt][x] = (k >= 0 && c < D) ? __ldg(table + k * D + c) : 0.f;
    }
    __syncthreads();
    const int64_t t = t0 + x;
    if (t < T) {
        const int64_t b = t / hw, p = t % hw;
        for (int cc = y; cc < 32; cc += 8) {
            const int c = c0 + cc;
            if (c < D) out[(b * D + c) * hw + p] = tile[x][cc];
        }
    }
}

cudaError_t launch_gather(const void* idx, int64_t T, int64_t hw, const float* table, int K, int D,
                          int layout_out, float* out, int64_t* stats, cudaStream_t s, int idx_bits) {
    if (T == 0) return cudaSuccess;
    if (layout_out == VQ_LAYOUT_TOKEN_MAJOR) {
        const int chunks = D / 4;
        int64_t blocks = (T * chunks + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        k_gather_tok<<<(unsigned)blocks, 256, 0, s>>>(idx, idx_bits, T, chunks, K, reinterpret_cast<const float4*>(table),
                                                     reinterpret_cast<float4*>(out), stats);
    } else {
        dim3 grid((unsigned)((T + 31) / 32), (unsigned)((D + 31) / 32));
        k_gather_nchw<<<grid, dim3(32, 8), 0, s>>>(idx, idx_bits, T, hw, D, K, table, out, stats);
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace vq
