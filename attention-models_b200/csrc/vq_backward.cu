// Backward of the quantiser: grad_z (straight-through + commitment term through the normalisation)
// and the codebook gradient as a deterministic segmented sum.
//
// The reference has no backward code: autograd derives it from models/vitvqgan.py:151-171 /
// models/vqgan.py:148-176 (driven by trainers/vitgqgan.py:171-184).  Closed form:
//   g_zn      = G + g_loss * c1 * 2 (zn - q) / N
//   grad_z    = (g_zn - zn (zn . g_zn)) / max(||z||, eps)
//   S_k       = sum_{t : idx_t = k} (q_k - zn_t)
//   grad_E[k] = (g_k - en_k (en_k . g_k)) / max(||E_k||, eps),   g_k = g_loss * c2 * (2/N) * S_k
// with (c1, c2) = (beta, 1) for the ViT form and (1, beta) for the VQGAN form.
//
// Determinism: tokens are bucketed by code (count -> exclusive scan -> scatter), every bucket is cut
// into pieces of kSegPiece tokens, one warp sums a piece.  Contributions are accumulated as 2^-30
// fixed-point int64, so the result does not depend on the order tokens landed in the bucket, on how
// pieces are combined, or -- for a token-sharded job -- on the number of GPUs: the integer sums are
// simply added (all-reduce) before the final normalise-backward.  No float atomics anywhere.
// The difference form sum(q_k - zn_t) is used (not n_k q_k - sum zn_t) because the latter cancels
// on trained data (SURVEY.md section 7 "Codebook-grad cancellation").
#include "vq_common.cuh"
#include "vq_kernels.h"
#include "vq_backward_body.cuh"
#include "../../include/vq_b200.h"

namespace vq {

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct BwdWs {
    int* counts;    // K
    int* cursor;    // K
    int* offsets;   // K + 1  exclusive scan of counts
    int* pieces;    // K + 1  exclusive scan of ceil(count / kSegPiece)
    int* perm;      // T      token ids grouped by code
    size_t zero_bytes;   // counts + cursor are contiguous and zeroed per call
};

size_t backward_workspace_bytes(int64_t T, int K, int D) {
    (void)D;
    return align_up(sizeof(int) * (size_t)K * 2, 256) + 2 * align_up(sizeof(int) * ((size_t)K + 1), 256) +
           align_up(sizeof(int) * (size_t)(T > 0 ? T : 1), 256);
}

static BwdWs carve(void* ws, int64_t T, int K) {
    char* p = static_cast<char*>(ws);
    BwdWs w;
    w.counts = reinterpret_cast<int*>(p);
    w.cursor = w.counts + K;
    w.zero_bytes = sizeof(int) * (size_t)K * 2;
    p += align_up(w.zero_bytes, 256);
    w.offsets = reinterpret_cast<int*>(p); p += align_up(sizeof(int) * ((size_t)K + 1), 256);
    w.pieces = reinterpret_cast<int*>(p);  p += align_up(sizeof(int) * ((size_t)K + 1), 256);
    w.perm = reinterpret_cast<int*>(p);
    (void)T;
    return w;
}

// ---------------------------------------------------------------------------------------------
// grad_z.  kLpr lanes share a row (one float4 each, kNf4 float4 per lane when D > 128); the row dot
// product is an xor-shuffle tree inside the lane group.  Algorithmic bytes 12D + 8 per token.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_backward_tokens(const float4* __restrict__ g, const float4* __restrict__ zn,
                                                         const float* __restrict__ denom,
                                                         const int64_t* __restrict__ idx, const float4* __restrict__ en,
                                                         int64_t T, float coef_base, const float* __restrict__ g_loss,
                                                         float4* __restrict__ grad) {
    pdl_trigger();
    pdl_wait();
    backward_tokens_body<D>(g, zn, denom, idx, en, T, coef_base, g_loss, grad, blockIdx.x, gridDim.x);
}

cudaError_t launch_backward_tokens(const float* g_tok, const float* zn32, const float* denom, const int64_t* idx,
                                   const CodebookView& cb, int64_t T, float coef_commit, const float* g_loss,
                                   float* grad_tok, cudaStream_t s) {
    if (T == 0) return cudaSuccess;
    const int chunks = cb.D / 4;
    const int lpr = chunks < 32 ? chunks : 32;
    const int rows_per_block = 8 * (32 / lpr);
    int64_t blocks = (T + rows_per_block - 1) / rows_per_block;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaError_t e = cudaSuccess;
    VQ_DISPATCH_D(cb.D, (e = launch_pdl(k_backward_tokens<kD>, dim3((unsigned)blocks), dim3(256), 0, s,
                                        reinterpret_cast<const float4*>(g_tok), reinterpret_cast<const float4*>(zn32), denom,
                                        idx, reinterpret_cast<const float4*>(cb.en32), T, coef_commit, g_loss,
                                        reinterpret_cast<float4*>(grad_tok))));
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// grad_z for the VQGAN form's (b, D, hw) tensors, without token-major staging: a block owns 32 consecutive tokens; the
// upstream gradient tile is read coalesced along the pixels into shared memory, one warp per token row then computes
// g_zn = G + coef (zn - q), the row dot product and grad_z = (g_zn - zn (zn . g_zn)) / max(||z||, eps) in place, and the
// tile goes back out coalesced along the pixels.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_backward_tokens_nchw(const float* __restrict__ g, const float* __restrict__ zn,
                                                              const float* __restrict__ denom,
                                                              const int64_t* __restrict__ idx, const float* __restrict__ en,
                                                              int64_t T, int64_t hw, float coef_base,
                                                              const float* __restrict__ g_loss, float* __restrict__ grad,
                                                              int raw) {
    extern __shared__ float tile[];                   // [32][D + 1]
    constexpr int kStride = D + 1;
    const float coef = coef_base * (g_loss ? __ldg(g_loss) : 1.f);
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    {
        const int64_t t = t0 + x;
        if (t < T) {
            const int64_t b = t / hw, pix = t % hw;
            for (int c = y; c < D; c += 8) tile[x * kStride + c] = g ? __ldcs(g + (b * D + c) * hw + pix) : 0.f;
        }
    }
    __syncthreads();
    for (int tt = y; tt < 32; tt += 8) {
        const int64_t t = t0 + tt;
        if (t >= T) break;
        const int64_t k = __ldg(idx + t);
        const float inv = __fdiv_rn(1.f, __ldg(denom + t));
        float a[(D + 31) / 32], gz[(D + 31) / 32];
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < (D + 31) / 32; ++j) {
            const int c = x + 32 * j;
            a[j] = 0.f; gz[j] = 0.f;
            if (c < D) {
                a[j] = __ldg(zn + t * D + c);
                gz[j] = __fmaf_rn(coef, __fsub_rn(a[j], __ldg(en + k * D + c)), tile[tt * kStride + c]);
                dot = __fmaf_rn(a[j], gz[j], dot);
            }
        }
        dot = raw ? 0.f : warp_sum(dot);              // raw (un-normalised) form: no projection, denom = 1: grad_z = g_zn
#pragma unroll
        for (int j = 0; j < (D + 31) / 32; ++j) {
            const int c = x + 32 * j;
            if (c < D) tile[tt * kStride + c] = __fmul_rn(__fmaf_rn(-a[j], dot, gz[j]), inv);
        }
    }
    __syncthreads();
    {
        const int64_t t = t0 + x;
        if (t < T) {
            const int64_t b = t / hw, pix = t % hw;
            for (int c = y; c < D; c += 8) __stcs(grad + (b * D + c) * hw + pix, tile[x * kStride + c]);
        }
    }
}

cudaError_t launch_backward_tokens_nchw(const float* g_nchw, const float* zn32, const float* denom, const int64_t* idx,
                                        const CodebookView& cb, int64_t T, int64_t hw, float coef_commit, const float* g_loss,
                                        float* grad_nchw, cudaStream_t s, bool raw) {
    if (T == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((T + 31) / 32);
    const size_t smem = sizeof(float) * 32 * (size_t)(cb.D + 1);
    cudaError_t e = cudaSuccess;
    VQ_DISPATCH_D(cb.D, {
        if (smem > 48 * 1024) {
            static PerDeviceOnce once;
            if (once.need())
                e = cudaFuncSetAttribute(k_backward_tokens_nchw<kD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        if (e == cudaSuccess) k_backward_tokens_nchw<kD><<<blocks, 256, smem, s>>>(g_nchw, zn32, denom, idx, cb.en32, T, hw,
                                                                                   coef_commit, g_loss, grad_nchw, raw ? 1 : 0);
    });
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// bucket tokens by code
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_count(const int64_t* __restrict__ idx, int64_t T, int* __restrict__ counts) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(counts + __ldg(idx + t), 1);
}

// single block: exclusive scans of counts and of piece counts (loads first, then warp-shuffle scans)
__global__ void __launch_bounds__(1024) k_scan_segments(const int* __restrict__ counts, int K, int* __restrict__ offsets,
                                                        int* __restrict__ pieces) {
    constexpr int kMaxPer = 16;                   // K <= 16384 keeps every thread's codes in registers
    __shared__ int w_tok[32], w_pc[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (K + 1023) / 1024;
    const int lo = tid * per;
    int n[kMaxPer];
    int tok = 0, pc = 0;
    if (per <= kMaxPer) {
#pragma unroll
        for (int i = 0; i < kMaxPer; ++i) n[i] = (i < per && lo + i < K) ? __ldg(counts + lo + i) : 0;
#pragma unroll
        for (int i = 0; i < kMaxPer; ++i) { tok += n[i]; pc += (n[i] + kSegPiece - 1) / kSegPiece; }
    } else {
        for (int k = lo; k < min(K, lo + per); ++k) { const int c = counts[k]; tok += c; pc += (c + kSegPiece - 1) / kSegPiece; }
    }
    int s_tok = tok, s_pc = pc;                   // inclusive scan inside the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int a = __shfl_up_sync(VQ_FULL, s_tok, off), b = __shfl_up_sync(VQ_FULL, s_pc, off);
        if (lane >= off) { s_tok += a; s_pc += b; }
    }
    if (lane == 31) { w_tok[warp] = s_tok; w_pc[warp] = s_pc; }
    __syncthreads();
    if (warp == 0) {
        int a = w_tok[lane], b = w_pc[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int x = __shfl_up_sync(VQ_FULL, a, off), y = __shfl_up_sync(VQ_FULL, b, off);
            if (lane >= off) { a += x; b += y; }
        }
        w_tok[lane] = a; w_pc[lane] = b;          // inclusive warp totals
    }
    __syncthreads();
    int run_tok = s_tok - tok + (warp ? w_tok[warp - 1] : 0);
    int run_pc = s_pc - pc + (warp ? w_pc[warp - 1] : 0);
    if (per <= kMaxPer) {
#pragma unroll
        for (int i = 0; i < kMaxPer; ++i)
            if (i < per && lo + i < K) {
                offsets[lo + i] = run_tok; pieces[lo + i] = run_pc;
                run_tok += n[i];
                run_pc += (n[i] + kSegPiece - 1) / kSegPiece;
            }
    } else {
        for (int k = lo; k < min(K, lo + per); ++k) {
            const int c = counts[k];
            offsets[k] = run_tok; pieces[k] = run_pc;
            run_tok += c;
            run_pc += (c + kSegPiece - 1) / kSegPiece;
        }
    }
    if (tid == 1023) { offsets[K] = w_tok[31]; pieces[K] = w_pc[31]; }
}

__global__ void __launch_bounds__(256) k_scatter(const int64_t* __restrict__ idx, int64_t T,
                                                 const int* __restrict__ offsets, int* __restrict__ cursor,
                                                 int* __restrict__ perm) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)__ldg(idx + t);
        perm[offsets[k] + atomicAdd(cursor + k, 1)] = (int)t;
    }
}

// ---------------------------------------------------------------------------------------------
// one warp per piece: S_k += sum over the piece's tokens of fixed(q_k - zn_t)
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_segment_reduce(const float4* __restrict__ zn, const float4* __restrict__ en,
                                                        const int* __restrict__ offsets, const int* __restrict__ pieces,
                                                        const int* __restrict__ perm, int K,
                                                        unsigned long long* __restrict__ seg_sums) {
    constexpr int kChunks = D / 4;
    constexpr int kLpr = (kChunks < 32) ? kChunks : 32;
    constexpr int kNf4 = kChunks / kLpr;
    constexpr int kRowsPerWarp = 32 / kLpr;
    const int lane = threadIdx.x & 31;
    const int sub = lane % kLpr, grp = lane / kLpr;
    const int n_pieces = pieces[K];
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * (blockDim.x >> 5);
    for (int w = warp; w < n_pieces; w += n_warps) {
        // largest k with pieces[k] <= w (codes with no tokens own no piece and are skipped by the search)
        int lo = 0, hi = K;   // invariant: pieces[lo] <= w < pieces[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(pieces + mid) <= w) lo = mid; else hi = mid;
        }
        const int k = lo;
        const int seg_lo = __ldg(offsets + k), seg_hi = __ldg(offsets + k + 1);
        const int p_lo = seg_lo + (w - __ldg(pieces + k)) * kSegPiece;
        const int p_hi = min(seg_hi, p_lo + kSegPiece);

        float4 q[kNf4];
#pragma unroll
        for (int f = 0; f < kNf4; ++f) q[f] = __ldg(en + (int64_t)k * kChunks + sub + kLpr * f);
        long long acc[kNf4][4];
#pragma unroll
        for (int f = 0; f < kNf4; ++f)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[f][i] = 0;
        unsigned bad = 0;

        for (int j0 = p_lo; j0 < p_hi; j0 += 2 * kRowsPerWarp) {
            const int ja = j0 + grp, jb = j0 + kRowsPerWarp + grp;
            const bool la = ja < p_hi, lb = jb < p_hi;
            const int64_t ta = la ? __ldg(perm + ja) : 0, tb = lb ? __ldg(perm + jb) : 0;
            float4 a[kNf4], b[kNf4];
#pragma unroll
            for (int f = 0; f < kNf4; ++f) {
                if (la) a[f] = __ldg(zn + ta * kChunks + sub + kLpr * f);
                if (lb) b[f] = __ldg(zn + tb * kChunks + sub + kLpr * f);
            }
#pragma unroll
            for (int f = 0; f < kNf4; ++f) {
                if (la) {
                    const float d[4] = {q[f].x - a[f].x, q[f].y - a[f].y, q[f].z - a[f].z, q[f].w - a[f].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (is_finite(d[i])) acc[f][i] += to_fixed(d[i], VQ_SEG_SHIFT);
                        else bad = 1;
                    }
                }
                if (lb) {
                    const float d[4] = {q[f].x - b[f].x, q[f].y - b[f].y, q[f].z - b[f].z, q[f].w - b[f].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (is_finite(d[i])) acc[f][i] += to_fixed(d[i], VQ_SEG_SHIFT);
                        else bad = 1;
                    }
                }
            }
        }
        // fold the lane groups (rows) of the warp together; integer adds, any order
#pragma unroll
        for (int off = kLpr; off < 32; off <<= 1) {
#pragma unroll
            for (int f = 0; f < kNf4; ++f)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[f][i] += __shfl_xor_sync(VQ_FULL, acc[f][i], off);
        }
        bad = __any_sync(VQ_FULL, bad);
        if (grp == 0) {
#pragma unroll
            for (int f = 0; f < kNf4; ++f)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    atomicAdd(seg_sums + (int64_t)k * D + (sub + kLpr * f) * 4 + i, (unsigned long long)acc[f][i]);
        }
        if (bad && lane == 0) atomicAdd(seg_sums + (int64_t)K * D + k, 1ull);
    }
}

cudaError_t launch_segment_sums(const float* zn32, const int64_t* idx, const int32_t* hist, const CodebookView& cb,
                                int64_t T, int64_t* seg_sums, void* ws, size_t ws_bytes, cudaStream_t s) {
    const int K = cb.K, D = cb.D;
    if (ws_bytes < backward_workspace_bytes(T, K, D)) return cudaErrorInvalidValue;
    BwdWs w = carve(ws, T, K);
    cudaError_t e;
    if ((e = cudaMemsetAsync(seg_sums, 0, sizeof(int64_t) * ((size_t)K * D + K), s)) != cudaSuccess) return e;
    if (T == 0) return cudaSuccess;
    if ((e = cudaMemsetAsync(w.counts, 0, w.zero_bytes, s)) != cudaSuccess) return e;
    int64_t blocks = (T + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    count_launch(3);
    // the forward's code-usage histogram is exactly the per-code token count; recount only without it
    const int* counts = hist ? reinterpret_cast<const int*>(hist) : w.counts;
    if (!hist) k_count<<<(unsigned)blocks, 256, 0, s>>>(idx, T, w.counts);
    k_scan_segments<<<1, 1024, 0, s>>>(counts, K, w.offsets, w.pieces);
    k_scatter<<<(unsigned)blocks, 256, 0, s>>>(idx, T, w.offsets, w.cursor, w.perm);
    int64_t max_pieces = T / kSegPiece + (T < K ? T : K) + 1;
    int64_t rblocks = (max_pieces + 7) / 8;
    const int64_t rcap = (int64_t)sm_count() * 8;
    if (rblocks > rcap) rblocks = rcap;
    VQ_DISPATCH_D(D, (k_segment_reduce<kD><<<(unsigned)rblocks, 256, 0, s>>>(
                         reinterpret_cast<const float4*>(zn32), reinterpret_cast<const float4*>(cb.en32), w.offsets,
                         w.pieces, w.perm, K, reinterpret_cast<unsigned long long*>(seg_sums))));
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// grad_E[k] = NB(E_k, coef * S_k): one warp per code
// ---------------------------------------------------------------------------------------------
struct CodebookGradArgs {
    const long long* seg_sums; const float* en; const float* code_denom; int K; float coef_base; float* grad;
    const int64_t* stats; int64_t n_elem_total; int form; float beta; float* loss;
};
template <int D>
__device__ __forceinline__ void codebook_grad_body(const CodebookGradArgs& a, const float* __restrict__ g_loss, int vblock,
                                                   int vgrid) {
    const long long* __restrict__ seg_sums = a.seg_sums;
    const float* __restrict__ en = a.en;
    const float* __restrict__ code_denom = a.code_denom;
    float* __restrict__ grad = a.grad;
    const int K = a.K;
    if (a.loss && vblock == 0 && threadIdx.x == 0)
        a.loss[0] = loss_from_fixed(a.stats[VQ_STAT_LOSS_FIXED], a.stats[VQ_STAT_NONFINITE], a.n_elem_total, a.form, a.beta);
    const float coef = a.coef_base * (g_loss ? __ldg(g_loss) : 1.f);
    constexpr int kPer = (D + 31) / 32;
    const int lane = threadIdx.x & 31;
    const int warp = vblock * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n_warps = vgrid * (blockDim.x >> 5);
    for (int k = warp; k < K; k += n_warps) {
        float g[kPer], y[kPer];
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int d = lane + 32 * j;
            g[j] = 0.f; y[j] = 0.f;
            if (d < D) {
                g[j] = seg_to_grad(__ldg(seg_sums + (int64_t)k * D + d), coef);
                y[j] = __ldg(en + (int64_t)k * D + d);
                dot = __fmaf_rn(y[j], g[j], dot);
            }
        }
        // un-normalised form (VQ_FORM_VQGAN_L2): grad_E = coef * S_k as it is (code_denom = 1, no projection)
        dot = (a.form == VQ_FORM_VQGAN_L2) ? 0.f : warp_sum(dot);
        const float inv = __fdiv_rn(1.f, __ldg(code_denom + k));
        const bool poisoned = __ldg(seg_sums + (int64_t)K * D + k) != 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int d = lane + 32 * j;
            if (d < D) grad[(int64_t)k * D + d] = poisoned ? __int_as_float(0x7fc00000) : grad_row_element(g[j], y[j], dot, inv);
        }
    }
}

template <int D>
__global__ void __launch_bounds__(256) k_codebook_grad(CodebookGradArgs a, const float* __restrict__ g_loss) {
    pdl_trigger();
    pdl_wait();
    codebook_grad_body<D>(a, g_loss, blockIdx.x, gridDim.x);
}

// grad_z on blocks [cbg_blocks, grid), grad_E (and the loss) on blocks [0, cbg_blocks): the backward of a step whose
// segment sums came from the forward, in one launch
template <int D>
__global__ void __launch_bounds__(256) k_backward_fused(CodebookGradArgs a, int cbg_blocks, const float4* __restrict__ g,
                                                        const float4* __restrict__ zn, const float* __restrict__ denom,
                                                        const int64_t* __restrict__ idx, const float4* __restrict__ en4,
                                                        int64_t T, float coef_commit, const float* __restrict__ g_loss,
                                                        float4* __restrict__ grad_tok) {
    pdl_trigger();
    pdl_wait();
    if ((int)blockIdx.x < cbg_blocks) codebook_grad_body<D>(a, g_loss, blockIdx.x, cbg_blocks);
    else backward_tokens_body<D>(g, zn, denom, idx, en4, T, coef_commit, g_loss, grad_tok, blockIdx.x - cbg_blocks, gridDim.x - cbg_blocks);
}

cudaError_t launch_backward_fused(const float* g_tok, const float* zn32, const float* denom, const int64_t* idx,
                                  const CodebookView& cb, int64_t T, float coef_commit, const float* g_loss, float* grad_tok,
                                  const int64_t* seg_sums, float coef_codebook, float* grad_weight, const int64_t* stats,
                                  int64_t n_elem_total, int form, float beta, float* loss, cudaStream_t s) {
    const int chunks = cb.D / 4;
    const int lpr = chunks < 32 ? chunks : 32;
    const int rows_per_block = 8 * (32 / lpr);
    int64_t tok_blocks = (T + rows_per_block - 1) / rows_per_block;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (tok_blocks > cap) tok_blocks = cap;
    if (tok_blocks < 1) tok_blocks = 1;
    int cbg_blocks = (cb.K + 7) / 8;
    if (cbg_blocks > sm_count() * 2) cbg_blocks = sm_count() * 2;
    CodebookGradArgs a{reinterpret_cast<const long long*>(seg_sums), cb.en32, cb.code_denom, cb.K, coef_codebook, grad_weight,
                       stats, n_elem_total, form, beta, stats ? loss : nullptr};
    cudaError_t e = cudaSuccess;
    VQ_DISPATCH_D(cb.D, (e = launch_pdl(k_backward_fused<kD>, dim3((unsigned)(cbg_blocks + tok_blocks)), dim3(256), 0, s, a, cbg_blocks,
                                        reinterpret_cast<const float4*>(g_tok), reinterpret_cast<const float4*>(zn32), denom, idx,
                                        reinterpret_cast<const float4*>(cb.en32), T, coef_commit, g_loss,
                                        reinterpret_cast<float4*>(grad_tok))));
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_codebook_grad(const int64_t* seg_sums, const CodebookView& cb, float coef, const float* g_loss,
                                 float* grad_weight, const int64_t* stats, int64_t n_elem_total, int form, float beta,
                                 float* loss, cudaStream_t s) {
    int blocks = (cb.K + 7) / 8;
    const int cap = sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cudaError_t e = cudaSuccess;
    CodebookGradArgs a{reinterpret_cast<const long long*>(seg_sums), cb.en32, cb.code_denom, cb.K, coef, grad_weight,
                       stats, n_elem_total, form, beta, stats ? loss : nullptr};
    VQ_DISPATCH_D(cb.D, (e = launch_pdl(k_codebook_grad<kD>, dim3(blocks), dim3(256), 0, s, a, g_loss)));
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace vq
