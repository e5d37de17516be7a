// Token-sharded job (one process per GPU of one NVSwitch box): the single exchange of the backward -- the sum over
// ranks of the integer codebook-gradient segment sums, the usage histogram and the loss partial -- fused with the
// codebook gradient it feeds, in ONE kernel over NVLink peer memory.  No collective library on this path.
//
// Reference semantics replaced: DDP's gradient all-reduce of `codebook.embedding.weight.grad`
// (trainers/vitgqgan.py:184 via accelerate, trainers/utils/base_trainer.py:29-33) + the normalise-backward of
// models/vitvqgan.py:154,162-166 that autograd derives.
//
// Every rank owns one exchange buffer (cudaMalloc, exported with CUDA IPC, mapped by all peers):
//     [ control: ready flags 16 x 128 B | done flags 16 x 128 B | counters ][ results ][ slot 0 ][ slot 1 ]
//     slot    = [ seg_sums (K*D + K) int64 | stats 8 int64 | hist K int32 ]      this rank's partials of a step
//     results = [ grad_E K*D fp32 | histogram K int64 ]                            written by the slice owners
// vq_forward writes seg_sums / stats / hist of step n straight into slot n & 1.  The kernel (one CTA per SM) is a
// reduce-scatter + all-gather in two phases, W ranks, rank r owning the codes [r K/W, (r+1) K/W):
//   1. block 0 publishes "rank r's partials of step n are ready" into every peer's ready[r] (st.release.sys after a
//      system fence); every block waits until its own ready[] shows all peers at step n (ld.acquire.sys, bounded spin);
//   2. one warp per owned code pulls that code's int64 sums from every rank (256-byte coalesced NVLink reads, all
//      ranks' loads in flight before the first add), adds them -- integer adds: exact and order-free, so the totals
//      are bit-identical to a single-GPU run on the concatenated batch -- applies grad_E[k] = NB(E_k, coef * S_k) and
//      stores the fp32 row into EVERY rank's results (128-byte NVLink writes); same for the histogram slice;
//   3. the last block of the rank to finish publishes done[r] to every peer; every block waits for all done[] flags and
//      copies the complete results from its own buffer to the caller's grad_weight / histogram.
// Per rank this moves (W-1)/W x (8D + 12) K bytes in and (W-1)/W x (4D + 8) K bytes out -- 2.8 MB at 8192 x 32 on 8
// GPUs, against 15 MB for an all-gather of the partials -- and needs two flag round trips.  A rank reuses a slot or
// the results only after its next kernel saw every peer's ready flag of the next step, i.e. after every peer's
// previous kernel -- which read them -- had completed.
#include <cstdio>
#include <cstdlib>

#include "vq_common.cuh"
#include "vq_kernels.h"
#include "vq_backward_body.cuh"
#include "../../include/vq_b200.h"

namespace vq {

namespace {
constexpr size_t kFlagStride = 128;
constexpr size_t kReadyOff = 0, kDoneOff = VQ_PEER_MAX_RANKS * kFlagStride, kCounterOff = 2 * VQ_PEER_MAX_RANKS * kFlagStride;
constexpr size_t kStepOff = kCounterOff + 128;      // device-resident count of completed exchanges of this rank
constexpr size_t kConfigOff = kCounterOff + 256;    // {timeout in ns (0 = default), device-visible address of the host abort flag}
constexpr size_t kControlBytes = kCounterOff + 512;      // (a multiple of 256: the slots behind it stay 256-byte aligned)
constexpr unsigned long long kDefaultTimeoutNs = 600ull * 1000ull * 1000ull * 1000ull;   // 10 minutes, the order of NCCL's watchdog
inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }
}  // namespace

ExchangeLayout exchange_layout(int K, int D) {
    ExchangeLayout L;
    L.seg_bytes = sizeof(int64_t) * ((size_t)K * D + K);
    L.stats_off = align256(L.seg_bytes);
    L.hist_off = L.stats_off + align256(sizeof(int64_t) * VQ_STATS_LEN);
    L.slot_bytes = L.hist_off + align256(sizeof(int32_t) * (size_t)K);
    L.results_off = kControlBytes;
    L.results_hist_off = align256(sizeof(float) * (size_t)K * D);
    L.slot0_off = L.results_off + L.results_hist_off + align256(sizeof(int64_t) * (size_t)K);
    L.total = L.slot0_off + 2 * L.slot_bytes;
    return L;
}

struct PeerPtrs {
    const char* base[VQ_PEER_MAX_RANKS];
};

// time-out (0 = default) and the device-visible address of a host flag the kernel raises when it gives up
cudaError_t peer_configure(void* own_buf, unsigned long long timeout_ns, void* abort_flag, cudaStream_t s) {
    const unsigned long long cfg[2] = {timeout_ns, reinterpret_cast<unsigned long long>(abort_flag)};
    cudaError_t e = cudaMemcpyAsync(static_cast<char*>(own_buf) + kConfigOff, cfg, sizeof(cfg), cudaMemcpyHostToDevice, s);
    return e != cudaSuccess ? e : cudaStreamSynchronize(s);      // cfg lives on this stack frame
}
// flags, block counter and step count back to zero (the configuration stays); every rank does this between two
// barriers of the job, with no exchange kernel in flight anywhere
cudaError_t peer_resync(void* own_buf, cudaStream_t s) {
    return cudaMemsetAsync(own_buf, 0, kConfigOff, s);
}

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

__device__ __forceinline__ void st_peer_f32(float* p, float v) {
    asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_peer_s64(long long* p, long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// wait until flags[r] >= epoch for every peer r (warp 0 of the block; returns false on timeout)
// (with_self: also this rank's own flag -- its last block to finish sets it)
__device__ __forceinline__ bool wait_flags(const char* own_base, size_t flags_off, int world, int rank, unsigned epoch,
                                           unsigned long long timeout_ns, bool with_self = false) {
    const int lane = threadIdx.x & 31;
    bool ok = true;
    if (lane < world && (with_self || lane != rank)) {
        const unsigned* f = reinterpret_cast<const unsigned*>(own_base + flags_off + (size_t)lane * kFlagStride);
        const unsigned long long t0 = global_ns();
        unsigned v;
        unsigned spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if ((int)(v - epoch) >= 0) break;
            // the wall clock is read every 1024 polls; a peer that never shows up must not hang the GPU for ever
            if ((++spins & 1023u) == 0 && global_ns() - t0 > timeout_ns) { ok = false; break; }
            __nanosleep(32);
        }
    }
    return __all_sync(VQ_FULL, ok);
}
__device__ __forceinline__ void publish_flags(const PeerPtrs& peers, size_t flags_off, int world, int rank, unsigned epoch,
                                              bool with_self = false) {
    const int lane = threadIdx.x & 31;
    if (lane < world && (with_self || lane != rank)) {
        __threadfence_system();
        unsigned* f = reinterpret_cast<unsigned*>(const_cast<char*>(peers.base[lane]) + flags_off + (size_t)rank * kFlagStride);
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
    }
}

// token backward riding in the same launch (blocks [exchange_blocks, grid)); exchange_blocks == 0: none
struct TokenBackwardArgs {
    const float4* g; const float4* zn; const float* denom; const int64_t* idx; const float4* en4; int64_t T; float coef_commit;
    float4* grad; int exchange_blocks;
};
// Small footprint on purpose (256 threads, <= 64 registers): the kernel spends most of its life waiting for other
// GPUs while the token backward shares the SMs with it; with 512 threads x ~100 registers it left room for one
// backward block per SM and the two kernels ran back to back in effect.
#ifndef VQ_PEER_THREADS
#define VQ_PEER_THREADS 256
#endif
template <int D>
__global__ void __launch_bounds__(VQ_PEER_THREADS, 1024 / VQ_PEER_THREADS)
k_codebook_grad_sharded(PeerPtrs peers, int world, int rank, int one_shot, ExchangeLayout L, int slot, unsigned epoch,
                        const float* __restrict__ en, const float* __restrict__ code_denom, int K, float coef_base,
                        const float* __restrict__ g_loss, int64_t n_elem_total, int form, float beta, float* __restrict__ grad,
                        int64_t* __restrict__ hist_total, float* __restrict__ loss, int64_t* __restrict__ stats_total,
                        TokenBackwardArgs tb) {
    if (tb.exchange_blocks > 0 && (int)blockIdx.x >= tb.exchange_blocks) {
        // grad_z on the blocks behind the exchange's (which are scheduled first and so all resident): no second
        // stream, no cross-stream events -- the token backward simply fills the SMs the exchange leaves idle
        if constexpr (VQ_PEER_THREADS == 256)
            backward_tokens_body<D>(tb.g, tb.zn, tb.denom, tb.idx, tb.en4, tb.T, tb.coef_commit, g_loss, tb.grad,
                                    blockIdx.x - tb.exchange_blocks, gridDim.x - tb.exchange_blocks);
        return;
    }
    __shared__ int s_last;
    const int xgrid = tb.exchange_blocks > 0 ? tb.exchange_blocks : (int)gridDim.x;     // blocks of the exchange proper
    const int lane = threadIdx.x & 31;
#ifdef VQ_PEER_TRACE
    auto now = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    unsigned long long tr[5] = {now(), 0, 0, 0, 0};
#endif
    char* own = const_cast<char*>(peers.base[rank]);
    bool timed_out = false;
    // The step number lives on the device (the last block of a launch bumps it), so a launch carries no per-step host
    // state and can sit in a CUDA graph; a host-side epoch, when given, must agree with it.
    if (world > 1) {
        const unsigned dev_epoch = *reinterpret_cast<volatile unsigned*>(own + kStepOff) + 1u;
        if ((epoch != 0u && epoch != dev_epoch) || (int)((dev_epoch - 1u) & 1u) != slot) timed_out = true;   // out of step
        epoch = dev_epoch;
    }
    const size_t slot_off = L.slot0_off + (size_t)slot * L.slot_bytes;
    unsigned long long timeout_ns = kDefaultTimeoutNs;
    int* abort_flag = nullptr;
    if (world > 1) {
        const unsigned long long* cfg = reinterpret_cast<const unsigned long long*>(own + kConfigOff);
        if (cfg[0] != 0ull) timeout_ns = cfg[0];
        abort_flag = reinterpret_cast<int*>(cfg[1]);
    }
    // A peer that never publishes its step (it died, or it is further behind than the time-out allows), or a launch that
    // is out of step with the device-side count, is FATAL for the exchange: the results are poisoned with NaN, the step
    // count is not advanced (every later launch fails the same way until vq_peer_resync), and the host abort flag is
    // raised so that the caller finds out without synchronising with the device.
    auto give_up = [&]() {
        const float nan = __int_as_float(0x7fc00000);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * D; i += xgrid * blockDim.x) grad[i] = nan;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (loss) loss[0] = nan;
            if (stats_total) atomicAdd(reinterpret_cast<unsigned long long*>(stats_total + VQ_STAT_PEER_TIMEOUT), 1ull);
            if (abort_flag) {
                *reinterpret_cast<volatile int*>(abort_flag) = 1;
                __threadfence_system();
            }
        }
    };
    // ---- 1. partials ready everywhere ----
    if (world > 1) {
        __shared__ int s_timed_out;
        if (threadIdx.x < 32) {
            if (blockIdx.x == 0 && !timed_out) publish_flags(peers, kReadyOff, world, rank, epoch);
            if (!timed_out) timed_out = !wait_flags(own, kReadyOff, world, rank, epoch, timeout_ns);
            if (threadIdx.x == 0) s_timed_out = timed_out;
        }
        __syncthreads();
        if (s_timed_out) { give_up(); return; }
    }
#ifdef VQ_PEER_TRACE
    tr[1] = now();
#endif
    // ---- 2. owned slice: totals over ranks, grad_E rows and histogram to every rank's results ----
    // A remote read is ~2 us away: a warp keeps kC codes x kR ranks of loads in flight and the grid is sized so that a
    // warp has at most a couple of iterations.
    const float coef = coef_base * (g_loss ? __ldg(g_loss) : 1.f);
    constexpr int kPer = (D + 31) / 32;
    constexpr int kC = (kPer <= 2) ? 2 : 1;
    constexpr int kR = (16 / (kPer * kC)) > 8 ? 8 : (16 / (kPer * kC));
    // one_shot (small worlds): every rank reduces ALL codes itself and writes its own outputs -- (W-1) x the partials
    // over NVLink but a single flag round trip; else the rank owns a slice and phase 3 gathers the slices
    const bool local_out = (world == 1) || one_shot;
    const int per_rank = local_out ? K : (K + world - 1) / world;
    const int k_lo = local_out ? 0 : min(K, rank * per_rank), k_hi = min(K, k_lo + per_rank);
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n_warps = xgrid * (blockDim.x >> 5);
    for (int k0 = k_lo + warp * kC; k0 < k_hi; k0 += n_warps * kC) {
        long long s[kC][kPer];
        int bad[kC];
        float y[kC][kPer], dn[kC];                       // local operands, requested before the remote ones are awaited
#pragma unroll
        for (int c = 0; c < kC; ++c) {
            bad[c] = 0;
            dn[c] = (k0 + c < k_hi) ? __ldg(code_denom + k0 + c) : 1.f;
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                s[c][j] = 0;
                const int d = lane + 32 * j;
                y[c][j] = (k0 + c < k_hi && d < D) ? __ldg(en + (int64_t)(k0 + c) * D + d) : 0.f;
            }
        }
        for (int r0 = 0; r0 < world; r0 += kR) {
            long long v[kR][kC][kPer];
            int b[kR][kC];
#pragma unroll
            for (int u = 0; u < kR; ++u) {
                const int r = r0 + u;
                const long long* seg = reinterpret_cast<const long long*>(peers.base[r < world ? r : rank] + slot_off);
#pragma unroll
                for (int c = 0; c < kC; ++c) {
                    const int k = k0 + c;
                    b[u][c] = 0;
#pragma unroll
                    for (int j = 0; j < kPer; ++j) {
                        const int d = lane + 32 * j;
                        v[u][c][j] = 0;
                        if (r < world && k < k_hi && d < D) v[u][c][j] = ld_peer_s64(seg + (int64_t)k * D + d);
                    }
                    if (r < world && k < k_hi && lane == 0)      // low word of the count: only zero / non-zero matters
                        b[u][c] = ld_peer_s32(reinterpret_cast<const int*>(seg + (int64_t)K * D + k)) |
                                  ld_peer_s32(reinterpret_cast<const int*>(seg + (int64_t)K * D + k) + 1);
                }
            }
#pragma unroll
            for (int u = 0; u < kR; ++u)
#pragma unroll
                for (int c = 0; c < kC; ++c) {
                    bad[c] |= b[u][c];
#pragma unroll
                    for (int j = 0; j < kPer; ++j) s[c][j] += v[u][c][j];
                }
        }
#pragma unroll
        for (int c = 0; c < kC; ++c) {
            const int k = k0 + c;
            if (k >= k_hi) break;
            float g[kPer];
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                const int d = lane + 32 * j;
                g[j] = 0.f;
                if (d < D) {
                    g[j] = seg_to_grad(s[c][j], coef);
                    dot = __fmaf_rn(y[c][j], g[j], dot);
                }
            }
            dot = warp_sum(dot);
            const float inv = __fdiv_rn(1.f, dn[c]);
            const bool poisoned = __shfl_sync(VQ_FULL, bad[c], 0) != 0;
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                const int d = lane + 32 * j;
                if (d < D) {
                    const float out = poisoned ? __int_as_float(0x7fc00000) : grad_row_element(g[j], y[c][j], dot, inv);
                    if (local_out) grad[(int64_t)k * D + d] = out;
                    else
                        for (int r = 0; r < world; ++r)
                            st_peer_f32(reinterpret_cast<float*>(const_cast<char*>(peers.base[r]) + L.results_off) + (int64_t)k * D + d, out);
                }
            }
        }
    }
    if (hist_total)
        // (trailing blocks first: they have no codes of the slice when K / W < warps of the grid)
        for (int k = k_lo + (xgrid - 1 - blockIdx.x) * blockDim.x + threadIdx.x; k < k_hi; k += xgrid * blockDim.x) {
            int hv[VQ_PEER_MAX_RANKS];
#pragma unroll
            for (int r = 0; r < VQ_PEER_MAX_RANKS; ++r)
                hv[r] = r < world ? ld_peer_s32(reinterpret_cast<const int*>(peers.base[r] + slot_off + L.hist_off) + k) : 0;
            long long h = 0;
#pragma unroll
            for (int r = 0; r < VQ_PEER_MAX_RANKS; ++r) h += hv[r];
            if (local_out) hist_total[k] = h;
            else
                for (int r = 0; r < world; ++r)
                    st_peer_s64(reinterpret_cast<long long*>(const_cast<char*>(peers.base[r]) + L.results_off + L.results_hist_off) + k, h);
        }
    // loss and summed statistics: 8 integers per rank, every rank adds them itself (the last warp of the last block,
    // so that no block's slice waits behind them)
    if (blockIdx.x == xgrid - 1 && threadIdx.x >= blockDim.x - 32 && (stats_total || loss)) {
        __shared__ long long tot[VQ_STATS_LEN];
        if (lane < VQ_STATS_LEN) {
            long long sv[VQ_PEER_MAX_RANKS];
#pragma unroll
            for (int r = 0; r < VQ_PEER_MAX_RANKS; ++r)
                sv[r] = r < world ? ld_peer_s64(reinterpret_cast<const long long*>(peers.base[r] + slot_off + L.stats_off) + lane) : 0;
            long long t = 0;
#pragma unroll
            for (int r = 0; r < VQ_PEER_MAX_RANKS; ++r) t += sv[r];
            tot[lane] = t;
            if (stats_total && lane != VQ_STAT_PEER_TIMEOUT) stats_total[lane] = t;
        }
        __syncwarp();
        if (lane == 0 && loss)
            loss[0] = loss_from_fixed(tot[VQ_STAT_LOSS_FIXED], tot[VQ_STAT_NONFINITE], n_elem_total, form, beta);
    }
    if (local_out) {
        if (world > 1) {                                 // the last block to get here closes the step
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned* counter = reinterpret_cast<unsigned*>(own + kCounterOff);
                if (atomicAdd(counter, 1u) == (unsigned)xgrid - 1u) {
                    *counter = 0;
                    *reinterpret_cast<volatile unsigned*>(own + kStepOff) = epoch;
                }
            }
        }
        return;
    }
    // ---- 3. slices complete everywhere, then results -> caller's tensors ----
#ifdef VQ_PEER_TRACE
    tr[2] = now();
#endif
    // every thread fences its own NVLink stores (one fence by thread 0 behind the barrier was tried: the peers then
    // read stale rows now and then)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* counter = reinterpret_cast<unsigned*>(own + kCounterOff);
        s_last = (atomicAdd(counter, 1u) == xgrid - 1);
        if (s_last) {
            *counter = 0;                                // ready for the next launch
            *reinterpret_cast<volatile unsigned*>(own + kStepOff) = epoch;   // every block has read the step number by now
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // "done" includes this rank's own flag: a block whose peers finished early must still wait for the other blocks
        // of its own rank, which write the rank's slice into the same results buffer
        if (s_last) publish_flags(peers, kDoneOff, world, rank, epoch, true);
        timed_out = !wait_flags(own, kDoneOff, world, rank, epoch, timeout_ns, true);
        if (threadIdx.x == 0) s_last = timed_out;            // (reused: "this block gave up")
    }
    __syncthreads();
    if (s_last) { give_up(); return; }
#ifdef VQ_PEER_TRACE
    tr[3] = now();
#endif
    const float4* res = reinterpret_cast<const float4*>(own + L.results_off);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * D / 4; i += xgrid * blockDim.x)
        reinterpret_cast<float4*>(grad)[i] = __ldcg(res + i);
    if (hist_total) {
        const long long* rh = reinterpret_cast<const long long*>(own + L.results_off + L.results_hist_off);
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += xgrid * blockDim.x) hist_total[k] = __ldcg(rh + k);
    }
#ifdef VQ_PEER_TRACE
    tr[4] = now();
    if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == xgrid - 1) && epoch >= 10 && epoch < 14)
        printf("rank %d epoch %u block %d: start %llu ready-wait %llu ns, slice %llu ns, done-wait %llu ns, copy %llu ns\n", rank, epoch,
               blockIdx.x, tr[0] % 1000000000ull, tr[1] - tr[0], tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3]);
#endif
}

cudaError_t launch_codebook_grad_sharded(const void* const* peer_bufs, int world, int rank, int slot, unsigned epoch,
                                         const CodebookView& cb, float coef, const float* g_loss, int64_t n_elem_total,
                                         int form, float beta, float* grad_weight, int64_t* hist_total, float* loss,
                                         int64_t* stats_total, const TokenBackward* tokens, cudaStream_t s) {
    const ExchangeLayout L = exchange_layout(cb.K, cb.D);
    PeerPtrs p;
    for (int r = 0; r < VQ_PEER_MAX_RANKS; ++r) p.base[r] = r < world ? static_cast<const char*>(peer_bufs[r]) : nullptr;
    // one CTA per SM: every block of the grid is resident, so the in-kernel waits of phase 3 cannot starve a block
    // that still has phase-2 work (they only ever wait for kernels of other GPUs)
    int blocks = sm_count();
    if (world == 1) {
        blocks = (cb.K + 7) / 8;
        if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    }
    TokenBackwardArgs tb = {};
    int64_t tok_blocks = 0;
    if (tokens && tokens->T > 0 && VQ_PEER_THREADS == 256) {
        const int chunks = cb.D / 4;
        const int lpr = chunks < 32 ? chunks : 32;
        const int rows_per_block = 8 * (32 / lpr);
        tok_blocks = (tokens->T + rows_per_block - 1) / rows_per_block;
        const int64_t cap = (int64_t)sm_count() * 16;
        if (tok_blocks > cap) tok_blocks = cap;
        tb = TokenBackwardArgs{reinterpret_cast<const float4*>(tokens->g_tok), reinterpret_cast<const float4*>(tokens->zn32),
                               tokens->denom, tokens->idx, reinterpret_cast<const float4*>(cb.en32), tokens->T,
                               tokens->coef_commit, reinterpret_cast<float4*>(tokens->grad_tok), blocks};
    }
    if (stats_total) {
        cudaError_t e = cudaMemsetAsync(stats_total + VQ_STAT_PEER_TIMEOUT, 0, sizeof(int64_t), s);
        if (e != cudaSuccess) return e;
    }
    // one flag round trip and (W-1) x 2.2 MB of reads (one-shot), or two round trips and 2 (W-1)/W x as much
    // (two-shot).  Measured per step at 8192 x 32: 2 GPUs 266 us either way, 4 GPUs 289 us one-shot / 269 us
    // two-shot -- the reads contend with the token backward that shares the launch -- so two-shot is the default;
    // VQ_EXCHANGE_ONE_SHOT=1 selects the other.
    static const int forced = getenv("VQ_EXCHANGE_ONE_SHOT") ? atoi(getenv("VQ_EXCHANGE_ONE_SHOT")) : -1;
    const int one_shot = forced >= 0 ? forced : 0;
    VQ_DISPATCH_D(cb.D, (k_codebook_grad_sharded<kD><<<(unsigned)(blocks + tok_blocks), world == 1 ? 256 : VQ_PEER_THREADS, 0, s>>>(
                            p, world, rank, one_shot, L, slot, epoch, cb.en32, cb.code_denom, cb.K, coef, g_loss, n_elem_total, form, beta,
                            grad_weight, hist_total, loss, stats_total, tb)));
    count_launch();
    return cudaGetLastError();
}

}  // namespace vq
