// First consumer of the quantiser's tokens: the mask-fill and token-embedding lookup
// every generative model of the reference does right behind encode_imgs.
//
// Reference call sites replaced (paths relative to /root/reference):
//   models/muse.py:149-150     input_ids = image_tokens.masked_fill(mask, mask_token_id)
//                              labels    = image_tokens.masked_fill(~mask, ignore_index)
//   models/maskgit.py:131-132  the same (tgt / x)
//   models/muse.py:90-91       img_token_embeds = token_emb(img_token_indices); img_token_embeds += pos_enc
//   models/maskgit.py:80-81    x = input_proj(x); x += pos_enc
// One pass: ids and labels (int64) and embeds[t] = table[id_t] + pos[t mod n] in fp32 (a gather and one rounded add:
// bit-identical to the reference).  HBM-bound: 4*dim B/token written, table rows and the positional rows come from
// L2; one warp per token, float4 per lane.
#include <algorithm>

#include "vq_common.cuh"
#include "vq_kernels.h"
#include "../../include/vq_b200.h"

namespace vq {

__global__ void __launch_bounds__(256) k_token_embed(const void* __restrict__ tokens, int token_bits, const uint8_t* __restrict__ mask,
                                                     int64_t T, int64_t n_per_seq, int64_t mask_token_id,
                                                     int64_t ignore_index, const float4* __restrict__ table, int64_t V,
                                                     int chunks, const float4* __restrict__ pos, float4* __restrict__ embeds,
                                                     int64_t* __restrict__ input_ids, int64_t* __restrict__ labels,
                                                     int64_t* __restrict__ stats, const float4* __restrict__ start) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t t = warp; t < T; t += n_warps) {
        if (start) {
            // autoregressive decoder input (models/parti.py:98-106): position 0 is the start token, position i > 0 the
            // embedding of token i - 1 plus the positional row i - 1; the last token of a sequence is not consumed
            const int64_t i = t % n_per_seq;
            if (i == 0) {
                for (int c = lane; c < chunks; c += 32) __stcs(embeds + t * chunks + c, __ldg(start + c));
                continue;
            }
            const int64_t id = load_token(tokens, t - 1, token_bits);
            const bool ok = id >= 0 && id < V;
            if (!ok && lane == 0 && stats) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_BAD_INDEX), 1ull);
            const float4* row = table + id * chunks;
            const float4* prow = pos ? pos + (i - 1) * chunks : nullptr;
            for (int c = lane; c < chunks; c += 32) {
                float4 v = ok ? __ldg(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (prow) {
                    const float4 p = __ldg(prow + c);
                    v.x = __fadd_rn(v.x, p.x); v.y = __fadd_rn(v.y, p.y); v.z = __fadd_rn(v.z, p.z); v.w = __fadd_rn(v.w, p.w);
                }
                __stcs(embeds + t * chunks + c, v);
            }
            continue;
        }
        const int64_t tok = load_token(tokens, t, token_bits);
        const bool masked = mask ? (__ldg(mask + t) != 0) : false;
        const int64_t id = masked ? mask_token_id : tok;
        if (lane == 0) {
            if (input_ids) input_ids[t] = id;
            if (labels) labels[t] = masked ? tok : ignore_index;
        }
        const bool ok = id >= 0 && id < V;
        if (!ok && lane == 0 && stats) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_BAD_INDEX), 1ull);
        if (!embeds) continue;
        const float4* row = table + id * chunks;
        const float4* prow = pos ? pos + (t % n_per_seq) * chunks : nullptr;
        for (int c = lane; c < chunks; c += 32) {
            float4 v = ok ? __ldg(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (prow) {
                const float4 p = __ldg(prow + c);
                v.x = __fadd_rn(v.x, p.x); v.y = __fadd_rn(v.y, p.y); v.z = __fadd_rn(v.z, p.z); v.w = __fadd_rn(v.w, p.w);
            }
            __stcs(embeds + t * chunks + c, v);
        }
    }
}

cudaError_t launch_token_embed(const void* tokens, const uint8_t* mask, int64_t T, int64_t n_per_seq, int64_t mask_token_id,
                               int64_t ignore_index, const float* table, int64_t V, int dim, const float* pos, float* embeds,
                               int64_t* input_ids, int64_t* labels, int64_t* stats, cudaStream_t s, const float* start,
                               int token_bits) {
    if (T == 0) return cudaSuccess;
    int64_t blocks = (T + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_token_embed<<<(unsigned)blocks, 256, 0, s>>>(tokens, token_bits, mask, T, n_per_seq, mask_token_id, ignore_index,
                                                   reinterpret_cast<const float4*>(table), V, dim / 4,
                                                   reinterpret_cast<const float4*>(pos), reinterpret_cast<float4*>(embeds),
                                                   input_ids, labels, stats, reinterpret_cast<const float4*>(start));
    count_launch();
    return cudaGetLastError();
}

// token indices between the wire formats
__global__ void __launch_bounds__(256) k_tokens_convert(const void* __restrict__ in, int in_bits, void* __restrict__ out,
                                                        int out_bits, int64_t T) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x)
        store_token(out, t, (int)load_token(in, t, in_bits), out_bits);
}

cudaError_t launch_tokens_convert(const void* in, int in_bits, void* out, int out_bits, int64_t T, cudaStream_t s) {
    if (T == 0) return cudaSuccess;
    int64_t blocks = (T + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k_tokens_convert<<<(unsigned)blocks, 256, 0, s>>>(in, in_bits, out, out_bits, T);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Backward of an embedding lookup: grad_table[id] += grad_out[row] for every (id, row) pair -- what autograd derives
// for nn.Embedding (models/muse.py:90, models/maskgit.py:80, models/parti.py:100) and for the decode gather
// Codebook.indices_to_embeddings (models/vitvqgan.py:173-176, models/vqgan.py:178-182).  Deterministic the same way
// as the codebook gradient of the quantiser: the sums are accumulated as 64-bit INTEGERS (exact, order-free) and
// converted once.  Upstream gradients have no natural scale, so the fixed-point scale is chosen per call from the
// largest magnitude in grad_out: 2^(62 - e - ceil(log2(n + 1))) with 2^e > max |g| -- no sum of n terms can overflow,
// and the resolution is <= max|g| * 2^-40 for n <= 2^22 terms (finer than an fp32 accumulation of the same rows).
// Work space: (V * dim + V) int64 sums and per-row counts of non-finite terms (such rows come out NaN), then
// {scale, 2^30 / scale} as two floats.
// ---------------------------------------------------------------------------------------------------------------
struct EmbBwdMap {
    int64_t ids_per_seq, rows_per_seq, row_shift;    // grad row of ids[j] = (j / ids_per_seq) * rows_per_seq + j % ids_per_seq + row_shift
    int64_t hw;                                      // > 0: grad_out is (b, dim, hw) (NCHW), token t = b * hw + p
};
__device__ __forceinline__ int64_t emb_grad_row(const EmbBwdMap& m, int64_t j) {
    return (j / m.ids_per_seq) * m.rows_per_seq + (j % m.ids_per_seq) + m.row_shift;
}

__global__ void __launch_bounds__(256) k_absmax(const float4* __restrict__ g, int64_t n4, unsigned* __restrict__ out_bits) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(g + i);
        const float a = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));     // fmaxf drops NaN: counted later
        if (is_finite(a)) m = fmaxf(m, a);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(VQ_FULL, m, off));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));      // non-negative floats order like their bits
}

__global__ void k_emb_scale(const unsigned* __restrict__ max_bits, int64_t n_terms, float* __restrict__ scales) {
    const float m = __uint_as_float(*max_bits);
    int e = 0;
    if (m > 0.f) { frexpf(m, &e); }                  // m = f * 2^e, f in [0.5, 1): 2^e > m
    int lg = 0;
    while ((1ll << lg) < n_terms + 1) ++lg;
    int sh = 62 - e - lg;
    sh = sh > 120 ? 120 : (sh < -120 ? -120 : sh);
    scales[0] = ldexpf(1.f, sh);                     // value -> fixed point
    scales[1] = ldexpf(1.f, VQ_SEG_SHIFT - sh);      // fixed point -> value, times the 2^30 seg_to_grad() divides by
}

template <bool kNchw>
__global__ void __launch_bounds__(256) k_emb_accumulate(const int64_t* __restrict__ ids, int64_t n_ids, EmbBwdMap map,
                                                        const float* __restrict__ g, int64_t V, int dim,
                                                        const float* __restrict__ scales, unsigned long long* __restrict__ acc) {
    const float scale = __ldg(scales);
    if (!kNchw) {
        // one warp per id, lanes along the row: every RED instruction covers 32 consecutive int64 (whole sectors)
        const int lane = threadIdx.x & 31;
        const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
        for (int64_t j = warp; j < n_ids; j += n_warps) {
            const int64_t id = __ldg(ids + j);
            if (id < 0 || id >= V) continue;         // padding / ignore ids take no gradient (nn.Embedding padding_idx semantics)
            const float* row = g + emb_grad_row(map, j) * dim;
            unsigned bad = 0;
            for (int d = lane; d < dim; d += 32) {
                const float v = __ldg(row + d);
                if (is_finite(v)) atomicAdd(acc + id * dim + d, (unsigned long long)__float2ll_rn(v * scale));
                else bad = 1;
            }
            if (__any_sync(VQ_FULL, bad) && lane == 0) atomicAdd(acc + V * dim + id, 1ull);
        }
    } else {
        // (b, dim, hw): threads along the tokens of one channel plane (coalesced reads), one RED per element
        const int64_t total = n_ids * dim;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t p = i % map.hw, d = (i / map.hw) % dim, b = i / (map.hw * dim);
            const int64_t id = __ldg(ids + b * map.hw + p);
            if (id < 0 || id >= V) continue;
            const float v = __ldg(g + i);
            if (is_finite(v)) atomicAdd(acc + id * dim + d, (unsigned long long)__float2ll_rn(v * scale));
            else atomicAdd(acc + V * dim + id, 1ull);
        }
    }
}

__global__ void __launch_bounds__(256) k_emb_finish(const long long* __restrict__ acc, int64_t V, int dim,
                                                    const float* __restrict__ scales, float* __restrict__ grad_table) {
    const double inv = (double)__ldg(scales + 1) / (double)(1ll << VQ_SEG_SHIFT);
    const int64_t total = V * dim;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const bool poisoned = acc[total + i / dim] != 0;
        grad_table[i] = poisoned ? __int_as_float(0x7fc00000) : (float)((double)acc[i] * inv);
    }
}

size_t embedding_backward_bytes(int64_t V, int dim) { return sizeof(int64_t) * (size_t)(V * dim + V) + 256; }

cudaError_t launch_embedding_backward(const int64_t* ids, int64_t n_ids, int64_t ids_per_seq, int64_t rows_per_seq,
                                      int64_t row_shift, const float* grad_out, int64_t n_rows, int64_t hw, int64_t V, int dim,
                                      const CodebookView* normalised, float* grad_table, void* ws, cudaStream_t s) {
    unsigned long long* acc = static_cast<unsigned long long*>(ws);
    const size_t acc_bytes = sizeof(int64_t) * (size_t)(V * dim + V);
    float* scales = reinterpret_cast<float*>(static_cast<char*>(ws) + acc_bytes);
    unsigned* max_bits = reinterpret_cast<unsigned*>(scales + 2);
    cudaError_t e = cudaMemsetAsync(ws, 0, acc_bytes + 256, s);
    if (e != cudaSuccess) return e;
    const int cap = sm_count() * 8;
    const int64_t n4 = n_rows * dim / 4;
    if (n4 > 0) {
        k_absmax<<<(unsigned)std::min<int64_t>((n4 + 255) / 256, cap), 256, 0, s>>>(reinterpret_cast<const float4*>(grad_out), n4, max_bits);
        count_launch();
    }
    k_emb_scale<<<1, 1, 0, s>>>(max_bits, n_ids, scales);
    count_launch();
    if (n_ids > 0) {
        const EmbBwdMap map{ids_per_seq > 0 ? ids_per_seq : n_ids, rows_per_seq > 0 ? rows_per_seq : n_ids, row_shift, hw};
        if (hw > 0)
            k_emb_accumulate<true><<<(unsigned)std::min<int64_t>((n_ids * dim + 255) / 256, cap * 4), 256, 0, s>>>(
                ids, n_ids, map, grad_out, V, dim, scales, acc);
        else
            k_emb_accumulate<false><<<(unsigned)std::min<int64_t>((n_ids + 7) / 8, cap * 2), 256, 0, s>>>(
                ids, n_ids, map, grad_out, V, dim, scales, acc);
        count_launch();
    }
    if (normalised) {
        // decode gather of the ViT form: out = l2norm(E[i]) -> grad_E[k] = NB(E_k, sum of the upstream rows); the codebook
        // gradient kernel takes the integer sums as they are (its coefficient is the device-side 2^30 / scale)
        return launch_codebook_grad(reinterpret_cast<const int64_t*>(acc), *normalised, 1.f, scales + 1, grad_table, nullptr, 1, 0, 0.f,
                                    nullptr, s);
    }
    k_emb_finish<<<(unsigned)std::min<int64_t>((V * dim + 255) / 256, cap), 256, 0, s>>>(reinterpret_cast<const long long*>(acc), V, dim,
                                                                                         scales, grad_table);
    count_launch();
    return cudaGetLastError();
}

}  // namespace vq
