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
#include "vq_common.cuh"
#include "vq_kernels.h"
#include "../../include/vq_b200.h"

namespace vq {

__global__ void __launch_bounds__(256) k_token_embed(const int64_t* __restrict__ tokens, const uint8_t* __restrict__ mask,
                                                     int64_t T, int64_t n_per_seq, int64_t mask_token_id,
                                                     int64_t ignore_index, const float4* __restrict__ table, int64_t V,
                                                     int chunks, const float4* __restrict__ pos, float4* __restrict__ embeds,
                                                     int64_t* __restrict__ input_ids, int64_t* __restrict__ labels,
                                                     int64_t* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t t = warp; t < T; t += n_warps) {
        const int64_t tok = __ldg(tokens + t);
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

cudaError_t launch_token_embed(const int64_t* tokens, const uint8_t* mask, int64_t T, int64_t n_per_seq, int64_t mask_token_id,
                               int64_t ignore_index, const float* table, int64_t V, int dim, const float* pos, float* embeds,
                               int64_t* input_ids, int64_t* labels, int64_t* stats, cudaStream_t s) {
    if (T == 0) return cudaSuccess;
    int64_t blocks = (T + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_token_embed<<<(unsigned)blocks, 256, 0, s>>>(tokens, mask, T, n_per_seq, mask_token_id, ignore_index,
                                                   reinterpret_cast<const float4*>(table), V, dim / 4,
                                                   reinterpret_cast<const float4*>(pos), reinterpret_cast<float4*>(embeds),
                                                   input_ids, labels, stats);
    count_launch();
    return cudaGetLastError();
}

}  // namespace vq
