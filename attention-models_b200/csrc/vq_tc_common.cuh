// Device helpers shared by the tcgen05 nearest-code search kernels (vq_dist_tc.cu, vq_dist_tc16.cu):
// mbarrier, TMA, tcgen05.mma / commit / ld wrappers, shared-memory operand descriptors, setmaxnreg.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

namespace vq {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
#ifdef VQ_MBAR_TEST_WAIT
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}
// non-blocking probe of a phase (mbarrier.test_wait): issued early, its result is consumed later, so its latency hides
// behind whatever the thread waits for in between
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}
// Bounded wait: a broken pipeline traps (launch failure) instead of hanging the GPU.  try_wait itself
// suspends the thread for a hardware-defined window, so the retry loop is just try_wait + counter:
// no clock reads, (almost) no issue slots stolen from the epilogue warps sharing the scheduler.
// (tools: a trapped kernel surfaces as cudaErrorLaunchFailure on the next API call.)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try(bar, parity))
        if (++spins > (1u << 27)) __trap();      // no call here: an ABI call would pin the register budget
}
#ifdef VQ_TC_INSTRUMENT
__device__ long long g_tc_wait[16];              // [role][what] cycle totals over all CTAs (diagnostic build only)
#define VQ_TIMED_WAIT(slot, bar, parity)                                        \
    do {                                                                        \
        const long long t0__ = clock64();                                       \
        mbar_wait(bar, parity);                                                 \
        wait_acc[slot] += clock64() - t0__;                                     \
    } while (0)
__device__ long long g_tc_trace[2][96][4];          // [role][event #][timestamp kind], CTA 0 only
#define VQ_TRACE(role, idx, kind) do { if (blockIdx.x == 0 && (idx) < 96) g_tc_trace[role][idx][kind] = clock64(); } while (0)
#define VQ_COUNT(slot) wait_acc[slot] += 1000
#define VQ_TIMED_BEGIN() const long long tb__ = clock64()
#define VQ_TIMED_END(slot) wait_acc[slot] += clock64() - tb__
// role-local bookkeeping: wait_acc[] + start time, flushed to g_tc_wait[base + i] (i < n) and the total at base + n
#define VQ_INSTR_BEGIN() long long wait_acc[4] = {0, 0, 0, 0}; const long long t_begin__ = clock64()
#define VQ_INSTR_END(base, n)                                                                              \
    do {                                                                                                   \
        if ((threadIdx.x & 31) == 0) {                                                                     \
            for (int i__ = 0; i__ < (n); ++i__)                                                            \
                atomicAdd((unsigned long long*)&g_tc_wait[(base) + i__], (unsigned long long)wait_acc[i__]); \
            atomicAdd((unsigned long long*)&g_tc_wait[(base) + (n)], (unsigned long long)(clock64() - t_begin__)); \
        }                                                                                                  \
    } while (0)
#else
#define VQ_TIMED_WAIT(slot, bar, parity) mbar_wait(bar, parity)
#define VQ_INSTR_BEGIN() do { } while (0)
#define VQ_INSTR_END(base, n) do { } while (0)
#define VQ_TRACE(role, idx, kind) do { } while (0)
#define VQ_COUNT(slot) do { } while (0)
#define VQ_TIMED_BEGIN() do { } while (0)
#define VQ_TIMED_END(slot) do { } while (0)
#endif

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// the same box into every CTA of `mask` (same CTA-relative destination and barrier offsets in each)
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                      uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "h"(mask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// all threads of all CTAs of the cluster
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_64B operand tile: rows of 64 B, 8-row atoms 512 B apart (SBO), LBO = 1, version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // leading byte offset (16-byte units)
    d |= (uint64_t)(512 >> 4) << 32;        // stride byte offset
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
    return d;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ... arriving on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}

// 32 lanes x 32 columns of fp32: thread i gets columns [col, col + 32) of TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
          "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
          "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr));
}
// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// one lane of a converged warp (elect.sync): the lane that issues TMA / tcgen05.mma / commit
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// diagnostic build only: print and clear the per-role wait counters after a search launch
static inline void instrument_report(cudaStream_t s, int grid) {
#ifdef VQ_TC_INSTRUMENT
    cudaStreamSynchronize(s);
    long long w[16];
    cudaMemcpyFromSymbol(w, g_tc_wait, sizeof(w));
    const double n = grid;
    printf("[tc instrument] per-CTA mean kcycles  producer: a_empty %.0f b_empty %.0f total %.0f | mma: a_full %.0f t_empty %.0f "
           "b_full %.0f issue %.0f total %.0f | epilogue: t_full %.0f h_empty %.0f ld_wait %.0f total %.0f\n",
           w[0] / n / 1e3, w[1] / n / 1e3, w[2] / n / 1e3, w[3] / n / 1e3, w[4] / n / 1e3, w[5] / n / 1e3, w[6] / n / 1e3,
           w[7] / n / 1e3, w[8] / n / 1e3, w[9] / n / 1e3, w[10] / n / 1e3, w[11] / n / 1e3);
    printf("[tc instrument] rescoring warp 0: h_full wait %.0f steps %.0f total %.0f\n", w[12] / n / 1e3, w[13] / n / 1e3,
           w[14] / n / 1e3);
    long long z[16] = {0};
    cudaMemcpyToSymbol(g_tc_wait, z, sizeof(z));
    static int traced = 0;
    if (traced++ == 3) {
        static long long tr[2][96][4];
        cudaMemcpyFromSymbol(tr, g_tc_trace, sizeof(tr));
        const long long t0 = tr[0][8][0];
        printf("[tc trace] tile: mma(t_empty ok, committed) | epi(t_full ok, data in regs, arrived, processed)  [cycles rel.]\n");
        for (int i = 8; i < 40; ++i)
            printf("[tc trace] %2d: %6lld %6lld | %6lld %6lld %6lld %6lld\n", i, tr[0][i][0] - t0, tr[0][i][1] - t0, tr[1][i][0] - t0,
                   tr[1][i][1] - t0, tr[1][i][2] - t0, tr[1][i][3] - t0);
    }
#else
    (void)s; (void)grid;
#endif
}

// (distance, index) pairs compared as one 64-bit key: sign-corrected float bits, ties to the lower index, NaN
// below everything (torch.argmin: NaN wins).  Shared by the exact rescoring kernels behind both filters.
__device__ __forceinline__ unsigned long long dist_key(float d, int code) {
    const uint32_t b = __float_as_uint(d);
    const uint32_t o = (d != d) ? 0u : (b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u));
    return ((unsigned long long)o << 32) | (uint32_t)code;
}
__device__ __forceinline__ float key_dist(unsigned long long key) {
    const uint32_t o = (uint32_t)(key >> 32);
    return (o == 0u) ? __int_as_float(0x7fc00000) : __uint_as_float(o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}
struct Top2 {
    unsigned long long best;   // key of the best (distance, index)
    float second;              // second-smallest distance (NaN if a NaN lost)
    __device__ __forceinline__ void init() { best = ~0ull; second = INFINITY; }
    __device__ __forceinline__ void lose(unsigned long long key) {
        if (key != ~0ull) {
            const float d = key_dist(key);
            second = (d != d || second != second) ? __int_as_float(0x7fc00000) : fminf(second, d);
        }
    }
    __device__ __forceinline__ void add(unsigned long long key) {
        const bool wins = key < best;
        const unsigned long long loser = wins ? best : key;
        best = wins ? key : best;
        lose(loser);
    }
    __device__ __forceinline__ void merge(unsigned long long obest, float osecond) {
        add(obest);
        second = (osecond != osecond || second != second) ? __int_as_float(0x7fc00000) : fminf(second, osecond);
    }
};

}  // namespace tc
}  // namespace vq
