// Token-sharded job (one process per GPU of one NVSwitch box): the single exchange of the backward -- the sum over
// ranks of the integer codebook-gradient segment sums, the usage histogram and the loss partial -- fused with the
// codebook gradient it feeds, in ONE kernel over NVLink peer memory.  No collective library on this path.
//
// Reference semantics replaced: DDP's gradient all-reduce of `codebook.embedding.weight.grad`
// (trainers/vitgqgan.py:184 via accelerate, trainers/utils/base_trainer.py:29-33) + the normalise-backward of
// models/vitvqgan.py:154,162-166 that autograd derives.
//
// Every rank owns one exchange buffer (cudaMalloc, exported with CUDA IPC, mapped by all peers):
//     [ flags: 16 x 128 B ][ slot 0 ][ slot 1 ]      slot = [ seg_sums (K*D + K) int64 | stats 8 int64 | hist K int32 ]
// vq_forward writes seg_sums / stats / hist of step n straight into slot n & 1.  The kernel then
//   1. block 0 publishes "rank r finished step n" into every peer's flags[r] (st.release.sys after a system fence);
//      every block waits until its own flags[] show all peers at step n (ld.acquire.sys; bounded spin);
//   2. one warp per code pulls that code's int64 sums from every rank (256-byte coalesced NVLink reads, all ranks'
//      loads in flight before the first add), adds them -- integer adds: exact, order-free, so every rank computes
//      bit-identical totals without a broadcast -- and applies grad_E[k] = NB(E_k, coef * S_k);
//   3. the histogram and the loss partial are summed the same way.
// Two slots make a trailing barrier unnecessary: a rank overwrites slot s at step n + 2, after its step n + 1 kernel
// saw every peer publish step n + 1, i.e. after every peer's step n kernel (which read slot s) had completed.
#include <cstdio>

#include "vq_common.cuh"
#include "vq_kernels.h"
#include "../../include/vq_b200.h"

namespace vq {

namespace {
constexpr size_t kFlagStride = 128;
constexpr size_t kFlagsBytes = VQ_PEER_MAX_RANKS * kFlagStride;
inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }
}  // namespace

ExchangeLayout exchange_layout(int K, int D) {
    ExchangeLayout L;
    L.seg_bytes = sizeof(int64_t) * ((size_t)K * D + K);
    L.stats_off = align256(L.seg_bytes);
    L.hist_off = L.stats_off + align256(sizeof(int64_t) * VQ_STATS_LEN);
    L.slot_bytes = L.hist_off + align256(sizeof(int32_t) * (size_t)K);
    L.slot0_off = kFlagsBytes;
    L.total = kFlagsBytes + 2 * L.slot_bytes;
    return L;
}

struct PeerPtrs {
    const char* base[VQ_PEER_MAX_RANKS];
};

__device__ __forceinline__ long long ld_peer_s64(const long long* p) {
    long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_peer_s32(const int* p) {
    int v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int D>
__global__ void __launch_bounds__(256) k_codebook_grad_sharded(PeerPtrs peers, int world, int rank, size_t slot_off,
                                                               size_t stats_off, size_t hist_off, unsigned epoch,
                                                               const float* __restrict__ en,
                                                               const float* __restrict__ code_denom, int K, float coef_base,
                                                               const float* __restrict__ g_loss, int64_t n_elem_total,
                                                               int form, float beta, float* __restrict__ grad,
                                                               int64_t* __restrict__ hist_total, float* __restrict__ loss,
                                                               int64_t* __restrict__ stats_total) {
    const int lane = threadIdx.x & 31;
    // ---- 1. publish / wait ----
    if (world > 1) {
        if (threadIdx.x < 32) {
            const bool is_peer = lane < world && lane != rank;
            if (blockIdx.x == 0 && is_peer) {
                __threadfence_system();
                unsigned* theirs = reinterpret_cast<unsigned*>(const_cast<char*>(peers.base[lane]) + (size_t)rank * kFlagStride);
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
            }
            bool ok = true;
            if (is_peer) {
                const unsigned* mine = reinterpret_cast<const unsigned*>(peers.base[rank] + (size_t)lane * kFlagStride);
                const long long t0 = clock64();
                unsigned v;
                for (;;) {
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
                    if ((int)(v - epoch) >= 0) break;
                    if (clock64() - t0 > 6000000000ll) { ok = false; break; }   // ~3 s: a peer died; do not hang the GPU
                    __nanosleep(64);
                }
            }
            if (!__all_sync(VQ_FULL, ok) && lane == 0 && stats_total)
                atomicAdd(reinterpret_cast<unsigned long long*>(stats_total + VQ_STAT_PEER_TIMEOUT), 1ull);
        }
        __syncthreads();
    }
    // ---- 2. grad_E, one warp per code ----
    const float coef = coef_base * (g_loss ? __ldg(g_loss) : 1.f);
    constexpr int kPer = (D + 31) / 32;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * (blockDim.x >> 5);
    for (int k = warp; k < K; k += n_warps) {
        long long s[kPer], bad = 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) s[j] = 0;
        for (int r0 = 0; r0 < world; r0 += 4) {          // 4 ranks' loads in flight per batch
            long long v[4][kPer], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + u;
                b[u] = 0;
#pragma unroll
                for (int j = 0; j < kPer; ++j) v[u][j] = 0;
                if (r < world) {
                    const long long* seg = reinterpret_cast<const long long*>(peers.base[r] + slot_off);
#pragma unroll
                    for (int j = 0; j < kPer; ++j) {
                        const int d = lane + 32 * j;
                        if (d < D) v[u][j] = ld_peer_s64(seg + (int64_t)k * D + d);
                    }
                    if (lane == 0) b[u] = ld_peer_s64(seg + (int64_t)K * D + k);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                bad += b[u];
#pragma unroll
                for (int j = 0; j < kPer; ++j) s[j] += v[u][j];
            }
        }
        float g[kPer], y[kPer];
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int d = lane + 32 * j;
            g[j] = 0.f; y[j] = 0.f;
            if (d < D) {
                g[j] = coef * (float)((double)s[j] * (1.0 / (double)(1ll << VQ_SEG_SHIFT)));
                y[j] = __ldg(en + (int64_t)k * D + d);
                dot += y[j] * g[j];
            }
        }
        dot = warp_sum(dot);
        const float inv = 1.f / __ldg(code_denom + k);
        const bool poisoned = __shfl_sync(VQ_FULL, bad, 0) != 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int d = lane + 32 * j;
            if (d < D) grad[(int64_t)k * D + d] = poisoned ? __int_as_float(0x7fc00000) : (g[j] - y[j] * dot) * inv;
        }
    }
    // ---- 3. histogram and loss ----
    if (hist_total)
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
            int64_t h = 0;
            for (int r = 0; r < world; ++r)
                h += ld_peer_s32(reinterpret_cast<const int*>(peers.base[r] + slot_off + hist_off) + k);
            hist_total[k] = h;
        }
    if (blockIdx.x == 0 && threadIdx.x < VQ_STATS_LEN && (stats_total || loss)) {
        __shared__ long long tot[VQ_STATS_LEN];
        long long t = 0;
        for (int r = 0; r < world; ++r)
            t += ld_peer_s64(reinterpret_cast<const long long*>(peers.base[r] + slot_off + stats_off) + threadIdx.x);
        tot[threadIdx.x] = t;
        if (stats_total && threadIdx.x != VQ_STAT_PEER_TIMEOUT) stats_total[threadIdx.x] = t;
        __syncwarp((1u << VQ_STATS_LEN) - 1u);
        if (threadIdx.x == 0 && loss) {
            loss[0] = loss_from_fixed(tot[VQ_STAT_LOSS_FIXED], tot[VQ_STAT_NONFINITE], n_elem_total, form, beta);
        }
    }
}

cudaError_t launch_codebook_grad_sharded(const void* const* peer_bufs, int world, int rank, int slot, unsigned epoch,
                                         const CodebookView& cb, float coef, const float* g_loss, int64_t n_elem_total,
                                         int form, float beta, float* grad_weight, int64_t* hist_total, float* loss,
                                         int64_t* stats_total, cudaStream_t s) {
    const ExchangeLayout L = exchange_layout(cb.K, cb.D);
    PeerPtrs p;
    for (int r = 0; r < VQ_PEER_MAX_RANKS; ++r) p.base[r] = r < world ? static_cast<const char*>(peer_bufs[r]) : nullptr;
    int blocks = (cb.K + 7) / 8;
    const int cap = sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (stats_total) {
        cudaError_t e = cudaMemsetAsync(stats_total + VQ_STAT_PEER_TIMEOUT, 0, sizeof(int64_t), s);
        if (e != cudaSuccess) return e;
    }
    VQ_DISPATCH_D(cb.D, (k_codebook_grad_sharded<kD><<<blocks, 256, 0, s>>>(
                            p, world, rank, L.slot0_off + (size_t)slot * L.slot_bytes, L.stats_off, L.hist_off, epoch, cb.en32,
                            cb.code_denom, cb.K, coef, g_loss, n_elem_total, form, beta, grad_weight, hist_total, loss,
                            stats_total)));
    count_launch();
    return cudaGetLastError();
}

}  // namespace vq
