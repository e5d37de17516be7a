// Nearest-code search on the 5th-generation tensor cores (sm_100a): TMA -> shared memory ->
// tcgen05.mma (fp16 in, fp32 accumulators in TMEM) -> tcgen05.ld -> running maxima in registers,
// followed by exact fp32 rescoring of the few surviving candidates.
//
// Replaces the dense part of the reference's search (paths relative to /root/reference):
//   models/vitvqgan.py:157-161 / models/vqgan.py:157-161
//     d = sum(z^2) + sum(e^2) - 2 * einsum('bd,nd->bn', z, e);  argmin(d, dim=1)
// The T x K matrix never exists: a CTA owns 256 token rows (two M = 128 MMA row tiles), streams the
// fp16 unit codebook through a TMA ring in tiles of 128 codes, and keeps 2 x 2 accumulator tiles of
// 128 x 128 fp32 in TMEM (all 512 columns) so the tensor core fills one stage while eight epilogue
// warps drain the other.
//
// Exactness.  The tensor-core scores are only a filter.  With unit rows, |fp16 dot - exact dot| <= eps
// (eps = 1.1e-3 covers fp16 rounding of both operands, fp32 accumulation, and the |en|^2 term, which
// the filter ignores).  Any code whose approximate score is more than 2*eps below the best approximate
// score cannot be the fp32 argmin, so it is dropped; every survivor is rescored with the reference
// formula in fp32 (same fma chain as the exhaustive SIMT search => identical indices).  Rows the
// filter cannot narrow down to one 256-code group are handed to the exhaustive search instead.
//
// Epilogue arithmetic (the bound at D = 32, where a 128x128 tile is only 128 MMA cycles): one thread
// owns one row; per 32-column chunk pair it does r[j] = max3(r[j], a[j], b[j]) -- 0.5 ALU op per
// distance (FMNMX3).  Every 256 columns (a "group") the 32 slot maxima are reduced, the group enters a
// running top-3 of group maxima, and the slot maxima of the two best groups are parked in shared memory
// (predicated STS.128 on the LSU pipe, so the ALU pipe -- the bottleneck -- is not charged).  At the end
// the row knows its two best groups, which of their 32 slots can still win (8 codes each), and a bound
// on every other group; it is undecided only if a third group is within 2*eps of the best.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>

#include "../../include/vq_b200.h"
#include "vq_common.cuh"
#include "vq_kernels.h"
#include "vq_tc_common.cuh"

namespace vq {

namespace tc {

constexpr int kRowsPerCta = 256;     // two MMA row tiles of 128
constexpr int kTileN = 128;          // codes per accumulator stage
constexpr int kGroupTiles = 2;       // tiles per group (256 columns)
constexpr int kGroupCols = kTileN * kGroupTiles;
constexpr int kKBlock = 32;          // halfs per K block: 64-byte rows, SWIZZLE_64B
constexpr int kABlockBytes = kRowsPerCta * kKBlock * 2;   // 16 KiB
constexpr int kBStageBytes = kTileN * kKBlock * 2;        // 8 KiB
constexpr int kThreads = 512;        // warps 0-3 service (TMA, MMA, TMEM alloc, idle), 4-7 rescoring, 8-15 epilogue
constexpr int kRescoreThreads = 128;
// register budget after setmaxnreg (the kernel launches with 128 per thread = the whole register file):
constexpr int kRegsService = 40, kRegsRescore = 96, kRegsEpilogue = 184;
static_assert(128 * kRegsService + 128 * kRegsRescore + 256 * kRegsEpilogue <= 65536, "register file overcommitted");
constexpr float kTwoEps = 2.2e-3f;   // 2 * eps, see header comment
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
// c_format = F32 (bit 4), a/b = F16 (0), K-major both, N = 128 (bits 17..22), M = 128 (bits 24..28)

struct SmemLayout {
    uint32_t a, b, snap, hand, bars, tmem_slot, total;
};
constexpr int kHandBytes = kRowsPerCta * 32;     // per row: {g1|g2<<16 or -1, g3, mask1, mask2} {mask3, -, -, -}
__host__ __device__ constexpr int a_stages(int kb) { return kb <= 2 ? 2 : 1; }
__host__ __device__ constexpr int b_stages(int kb) { return kb == 1 ? 8 : (kb <= 4 ? 4 : 2); }
// slot-maxima snapshots of the best groups: kAreas x 256 rows x 32 floats; rows padded by 16 B
// (conflict-free STS.128) except at D = 256 where shared memory is tight (2 areas, unpadded)
__host__ __device__ constexpr int snap_areas(int kb) { return kb <= 4 ? 3 : 2; }
__host__ __device__ constexpr int snap_row_bytes(int kb) { return kb <= 4 ? 144 : 128; }
constexpr int kMaxBStages = 8;
__host__ __device__ inline SmemLayout smem_layout(int kb) {
    SmemLayout L;
    L.a = 0;
    L.b = L.a + a_stages(kb) * kb * kABlockBytes;
    L.snap = L.b + b_stages(kb) * kBStageBytes;
    L.hand = L.snap + snap_areas(kb) * kRowsPerCta * snap_row_bytes(kb);
    L.bars = L.hand + 2 * kHandBytes;
    L.tmem_slot = L.bars + 8 * (2 * kMaxBStages + 2 * 2 + 4 + 4);
    L.total = L.tmem_slot + 16;
    return L;
}

// Codes of one cell: slot s of group g covers columns 64*m + (s < 16 ? 0 : 32) + (s & 15) + {0, 16}, m = 0..3
// (an epilogue "unit" is two x16 TMEM loads, 16 columns apart, folded into 16 slots by one max3 each; even
// units feed slots 0..15, odd units slots 16..31).
__device__ __forceinline__ int cell_code(int g, int slot, int i) {
    return g * kGroupCols + 64 * (i >> 1) + ((slot & 16) << 1) + (slot & 15) + 16 * (i & 1);
}

// exact fp32 distance of one row to the 8 codes of cell (g, slot): same fma chain as the exhaustive search
template <int D, bool kRowInRegs>
__device__ __forceinline__ void rescore_cell(int g, int slot, const float4* __restrict__ z4, const float4 (&z)[kRowInRegs ? D / 4 : 1],
                                             float a_sq, const float* __restrict__ en32, const float* __restrict__ code_sq,
                                             float& best_d, int& best_i, float& second_d) {
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        const int code = cell_code(g, slot, i);
        const float4* e4 = reinterpret_cast<const float4*>(en32 + (int64_t)code * D);
        float dot = 0.f;
        if constexpr (kRowInRegs) {
            float4 ev[D / 4];
#pragma unroll
            for (int q = 0; q < D / 4; ++q) ev[q] = __ldg(e4 + q);
#pragma unroll
            for (int q = 0; q < D / 4; ++q) {
                dot = __fmaf_rn(z[q].x, ev[q].x, dot);
                dot = __fmaf_rn(z[q].y, ev[q].y, dot);
                dot = __fmaf_rn(z[q].z, ev[q].z, dot);
                dot = __fmaf_rn(z[q].w, ev[q].w, dot);
            }
        } else {
#pragma unroll 8
            for (int q = 0; q < D / 4; ++q) {
                const float4 ev = __ldg(e4 + q);
                const float4 zv = __ldg(z4 + q);
                dot = __fmaf_rn(zv.x, ev.x, dot);
                dot = __fmaf_rn(zv.y, ev.y, dot);
                dot = __fmaf_rn(zv.z, ev.z, dot);
                dot = __fmaf_rn(zv.w, ev.w, dot);
            }
        }
        const float dist = ref_distance(a_sq, __ldg(code_sq + code), dot);
        if (argmin_better(dist, code, best_d, best_i)) { second_d = best_d; best_d = dist; best_i = code; }
        else if (dist < second_d) second_d = dist;
    }
}

// One CTA per SM, persistent over row tiles.  KB = D / 32.
// warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM alloc   warps 4-7: exact fp32 rescoring of the
// previous row tile (one warp per scheduler, hidden under the ALU-bound main loop)
// warps 8-15: epilogue, one thread per row.  Registers are redistributed with setmaxnreg.
template <int KB>
__global__ void __launch_bounds__(kThreads, 1)
k_dist_tc(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int T, int K,
          const float* __restrict__ zn32, const float* __restrict__ row_sq, const float* __restrict__ en32,
          const float* __restrict__ code_sq, const int* __restrict__ cb_info, int* __restrict__ cand,
          int* __restrict__ flagged, int* __restrict__ n_flagged, int64_t* __restrict__ stats, int debug_flags) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int D = KB * kKBlock;
    constexpr int AS = a_stages(KB);
    constexpr int BS = b_stages(KB);
    constexpr int kAreas = snap_areas(KB);
    constexpr int kSnapRow = snap_row_bytes(KB);
    const SmemLayout L = smem_layout(KB);
    // swizzled TMA/UMMA tiles want a 1024-byte aligned base; the launch reserves 1 KiB of slack for this
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + L.bars;
    // barrier map (8 bytes each)
    auto b_full = [&](int s) { return bar_base + 8 * s; };
    auto b_empty = [&](int s) { return bar_base + 8 * (kMaxBStages + s); };
    auto t_full = [&](int s) { return bar_base + 8 * (2 * kMaxBStages + s); };
    auto t_empty = [&](int s) { return bar_base + 8 * (2 * kMaxBStages + 2 + s); };
    auto a_full = [&](int s) { return bar_base + 8 * (2 * kMaxBStages + 4 + s); };
    auto a_empty = [&](int s) { return bar_base + 8 * (2 * kMaxBStages + 6 + s); };
    auto h_full = [&](int s) { return bar_base + 8 * (2 * kMaxBStages + 8 + s); };
    auto h_empty = [&](int s) { return bar_base + 8 * (2 * kMaxBStages + 10 + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L.tmem_slot);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_row_tiles = (T + kRowsPerCta - 1) / kRowsPerCta;
    const int n_tiles = K / kTileN;
    const int n_groups = n_tiles / kGroupTiles;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < BS; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(t_full(s), 1); mbar_init(t_empty(s), 256); }
        for (int s = 0; s < 2; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(h_full(s), 256); mbar_init(h_empty(s), kRescoreThreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + L.tmem_slot),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        reg_dec<kRegsService>();
        if (warp == 0 && lane == 0) {
            // ===================== TMA producer =====================
            uint32_t b_cnt = 0;
            int it = 0;
#ifdef VQ_TC_INSTRUMENT
            long long wait_acc[4] = {0, 0, 0, 0};
            const long long t_begin = clock64();
#endif
            for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x, ++it) {
                const int as = it % AS;
                VQ_TIMED_WAIT(0, a_empty(as), (((uint32_t)(it / AS)) & 1u) ^ 1u);
                mbar_expect_tx(a_full(as), KB * kABlockBytes);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(smem_base + L.a + (as * KB + kb) * kABlockBytes, &tm_a, a_full(as), kb * kKBlock,
                                rt * kRowsPerCta);
                for (int n = 0; n < n_tiles; ++n) {
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb, ++b_cnt) {
                        const int s = b_cnt % BS;
                        VQ_TIMED_WAIT(1, b_empty(s), ((b_cnt / BS) & 1u) ^ 1u);
                        mbar_expect_tx(b_full(s), kBStageBytes);
                        tma_load_2d(smem_base + L.b + s * kBStageBytes, &tm_b, b_full(s), kb * kKBlock, n * kTileN);
                    }
                }
            }
#ifdef VQ_TC_INSTRUMENT
            atomicAdd((unsigned long long*)&g_tc_wait[0], (unsigned long long)wait_acc[0]);
            atomicAdd((unsigned long long*)&g_tc_wait[1], (unsigned long long)wait_acc[1]);
            atomicAdd((unsigned long long*)&g_tc_wait[2], (unsigned long long)(clock64() - t_begin));
#endif
        } else if (warp == 1 && lane == 0) {
            // ===================== MMA issuer (one thread) =====================
            uint32_t b_cnt = 0, t_cnt = 0;
            int it = 0;
#ifdef VQ_TC_INSTRUMENT
            long long wait_acc[4] = {0, 0, 0, 0};
            const long long t_begin = clock64();
#endif
            for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x, ++it) {
                const int as = it % AS;
                VQ_TIMED_WAIT(0, a_full(as), ((uint32_t)(it / AS)) & 1u);
                tc_fence_after();
                for (int n = 0; n < n_tiles; ++n, ++t_cnt) {
                    const int acc = t_cnt & 1;
                    VQ_TIMED_WAIT(1, t_empty(acc), ((t_cnt >> 1) & 1u) ^ 1u);
                    tc_fence_after();
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb, ++b_cnt) {
                        const int s = b_cnt % BS;
                        VQ_TIMED_WAIT(2, b_full(s), (b_cnt / BS) & 1u);
                        tc_fence_after();
                        const uint32_t a_addr = smem_base + L.a + (as * KB + kb) * kABlockBytes;
                        const uint32_t b_addr = smem_base + L.b + s * kBStageBytes;
#pragma unroll
                        for (int r = 0; r < 2; ++r)
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                umma_f16(tmem_base + (uint32_t)((acc * 2 + r) * kTileN),
                                         umma_desc(a_addr + r * (128 * 64) + k * 32), umma_desc(b_addr + k * 32), kIdesc,
                                         (uint32_t)((kb | k) != 0));
                        umma_commit(b_empty(s));     // smem stage free once these MMAs retire
                    }
                    umma_commit(t_full(acc));        // accumulators of this tile complete
                }
                umma_commit(a_empty(as));            // row tile's A operand no longer needed
            }
#ifdef VQ_TC_INSTRUMENT
            atomicAdd((unsigned long long*)&g_tc_wait[4], (unsigned long long)wait_acc[0]);
            atomicAdd((unsigned long long*)&g_tc_wait[5], (unsigned long long)wait_acc[1]);
            atomicAdd((unsigned long long*)&g_tc_wait[6], (unsigned long long)wait_acc[2]);
            atomicAdd((unsigned long long*)&g_tc_wait[7], (unsigned long long)(clock64() - t_begin));
#endif
        }
    } else if (warp < 8) {
        // ===================== rescoring: 128 threads, 2 rows each per row tile =====================
        reg_dec<kRegsRescore>();
        constexpr bool kRowInRegs = (D <= 32);
        const int rtid = threadIdx.x - 128;
        unsigned ties = 0, multi = 0;
        int it = 0;
        for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x, ++it) {
            const int hb = it & 1;
            mbar_wait(h_full(hb), ((uint32_t)(it >> 1)) & 1u);
            const int4* hand = reinterpret_cast<const int4*>(smem + L.hand + hb * kHandBytes);
#pragma unroll 1
            for (int rr = 0; rr < kRowsPerCta / kRescoreThreads; ++rr) {
                const int r = rtid + kRescoreThreads * rr;
                const int row = rt * kRowsPerCta + r;
                const int4 h0 = hand[2 * r], h1 = hand[2 * r + 1];
                if (row >= T || h0.x < 0) continue;                 // out of range, or left to the exhaustive search
                if (debug_flags & 1) { cand[row] = kCandExactBit; continue; }   // timing experiment only
                const int gs[3] = {h0.x & 0xFFFF, (h0.x >> 16) & 0x7FFF, h0.y};
                const uint32_t ms[3] = {(uint32_t)h0.z, (uint32_t)h0.w, (uint32_t)h1.x};
                float4 z[kRowInRegs ? D / 4 : 1];
                const float4* z4 = reinterpret_cast<const float4*>(zn32 + (int64_t)row * D);
                if (kRowInRegs) {
#pragma unroll
                    for (int q = 0; q < (kRowInRegs ? D / 4 : 1); ++q) z[q] = __ldg(z4 + q);
                }
                const float a_sq = __ldg(row_sq + row);
                float best_d = INFINITY, second_d = INFINITY;
                int best_i = 0x7fffffff, n_cells = 0;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    uint32_t m = ms[a];
                    while (m) {
                        const int slot = __ffs(m) - 1;
                        m &= m - 1;
                        ++n_cells;
                        rescore_cell<D, kRowInRegs>(gs[a], slot, z4, z, a_sq, en32, code_sq, best_d, best_i, second_d);
                    }
                }
                cand[row] = best_i | kCandExactBit;
                if (second_d - best_d < VQ_NEAR_TIE_REL * fabsf(best_d)) ++ties;
                if (n_cells > 1) ++multi;
            }
            mbar_arrive(h_empty(hb));
        }
        if (stats) {
            ties = __reduce_add_sync(VQ_FULL, ties);
            multi = __reduce_add_sync(VQ_FULL, multi);
            if (lane == 0) {
                if (ties) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_NEAR_TIE_ROWS), (unsigned long long)ties);
                if (multi) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_AMBIGUOUS_ROWS), (unsigned long long)multi);
            }
        }
    } else {
        // ===================== epilogue: 8 warps, one thread per row =====================
        reg_inc<kRegsEpilogue>();
        const int e = warp - 8;
        const int r_sub = e >> 2;                    // which 128-row MMA tile
        const int quarter = warp & 3;                // TMEM lane quarter this warp may read
        const int row_in_cta = r_sub * 128 + quarter * 32 + lane;
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(r_sub * kTileN);
        const uint32_t snap0 = smem_base + L.snap + (uint32_t)row_in_cta * kSnapRow;
        constexpr uint32_t kSnapArea = kRowsPerCta * kSnapRow;
        const bool force_exhaustive = codebook_degenerate(cb_info);
        uint32_t phase = 0;                          // parity of t_full: flips once per group (2 tiles, 2 stages)
        int it = 0;
#ifdef VQ_TC_INSTRUMENT
        long long wait_acc[4] = {0, 0, 0, 0};
        const long long t_begin = clock64();
#endif
        // a "batch" is 64 accumulator columns = 4 x16 TMEM loads; batch b+1 is in flight while batch b is
        // folded into the 32 slots (columns 0-31 of the batch -> slots 0-15, columns 32-63 -> slots 16-31)
        auto load_batch = [&](uint32_t ta, float* v) {
            tmem_ld16(ta, v);
            tmem_ld16(ta + 16, v + 16);
            tmem_ld16(ta + 32, v + 32);
            tmem_ld16(ta + 48, v + 48);
        };
        for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x, ++it) {
            float slot[32];
            float m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY, m4 = -INFINITY;
            int g1 = 0, g2 = 0, g3 = 0;
            uint32_t a1 = 0, a2 = 1, a3 = (kAreas == 3) ? 2 : 1;   // snapshot areas of the best / second / third group
            float buf[2][64];
            VQ_TIMED_WAIT(0, t_full(0), phase);
            tc_fence_after();
            load_batch(tbase, buf[0]);
            for (int g = 0; g < n_groups; ++g) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    tmem_ld_wait();                                   // batch b has landed in buf[b & 1]
                    if (b == 1) { tc_fence_before(); mbar_arrive(t_empty(0)); }   // tile 0 fully in registers
                    if (b == 3) { tc_fence_before(); mbar_arrive(t_empty(1)); }
                    if (b < 3) {
                        if (b == 1) { VQ_TIMED_WAIT(0, t_full(1), phase); tc_fence_after(); }
                        load_batch(tbase + (uint32_t)(((b + 1) >> 1) * 2 * kTileN + ((b + 1) & 1) * 64), buf[(b + 1) & 1]);
                    } else if (g + 1 < n_groups) {
                        VQ_TIMED_WAIT(0, t_full(0), phase ^ 1u);
                        tc_fence_after();
                        load_batch(tbase, buf[0]);
                    }
                    const float* v = buf[b & 1];
                    if (b == 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) { slot[j] = fmaxf(v[j], v[16 + j]); slot[16 + j] = fmaxf(v[32 + j], v[48 + j]); }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            slot[j] = max3(slot[j], v[j], v[16 + j]);
                            slot[16 + j] = max3(slot[16 + j], v[32 + j], v[48 + j]);
                        }
                    }
                }
                phase ^= 1u;
                // group maximum: 3-input tree over the 32 slots
                float t[11];
#pragma unroll
                for (int j = 0; j < 10; ++j) t[j] = max3(slot[3 * j], slot[3 * j + 1], slot[3 * j + 2]);
                t[10] = fmaxf(slot[30], slot[31]);
                const float u0 = max3(t[0], t[1], t[2]), u1 = max3(t[3], t[4], t[5]), u2 = max3(t[6], t[7], t[8]);
                const float c1 = max3(max3(u0, u1, u2), t[9], t[10]);
                const bool is1 = c1 > m1, is2 = c1 > m2, is3 = (kAreas == 3) ? (c1 > m3) : is2;
                if (is3) {
                    // whichever rank the group takes, the group that drops out is the current last one: reuse its area
                    const uint32_t dst = snap0 + a3 * kSnapArea;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16 * q), "f"(slot[4 * q]),
                                     "f"(slot[4 * q + 1]), "f"(slot[4 * q + 2]), "f"(slot[4 * q + 3])
                                     : "memory");
                }
                // sorted insert of c1 into (m1 >= m2 >= m3 >= m4); identities and areas follow
                const float lo1 = fminf(c1, m1);
                m1 = fmaxf(c1, m1);
                const float lo2 = fminf(lo1, m2);
                m2 = fmaxf(lo1, m2);
                if (kAreas == 3) {
                    const float lo3 = fminf(lo2, m3);
                    m3 = fmaxf(lo2, m3);
                    m4 = fmaxf(lo3, m4);
                    const int ng3 = is2 ? g2 : (is3 ? g : g3);
                    const uint32_t na3 = is2 ? a2 : a3;
                    const int ng2 = is1 ? g1 : (is2 ? g : g2);
                    const uint32_t na2 = is1 ? a1 : (is2 ? a3 : a2);
                    const uint32_t na1 = is1 ? a3 : a1;
                    g1 = is1 ? g : g1; g2 = ng2; g3 = ng3;
                    a1 = na1; a2 = na2; a3 = na3;
                } else {
                    m3 = fmaxf(lo2, m3);             // with two areas m3 is the bound on everything else
                    const int ng2 = is1 ? g1 : (is2 ? g : g2);
                    const uint32_t na1 = is1 ? a2 : a1, na2 = is1 ? a1 : a2;
                    g1 = is1 ? g : g1; g2 = ng2;
                    a1 = na1; a2 = na2; a3 = na2;
                }
            }
            // ---- row verdict ----
            const float thr = m1 - kTwoEps;
            float kept[32];
            uint32_t mask[3] = {0u, 0u, 0u};
            const float mv[3] = {m1, m2, m3};
            const uint32_t av[3] = {a1, a2, a3};
#pragma unroll
            for (int a = 0; a < kAreas; ++a) {
                if (a == 0 || mv[a] >= thr) {
                    const uint32_t src = snap0 + av[a] * kSnapArea;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(kept[4 * q]), "=f"(kept[4 * q + 1]), "=f"(kept[4 * q + 2]), "=f"(kept[4 * q + 3])
                                     : "r"(src + 16 * q));
#pragma unroll
                    for (int j = 0; j < 32; ++j) mask[a] |= (kept[j] >= thr) ? (1u << j) : 0u;
                }
            }
            // decided iff no further group can hold the winner (NaN / -inf rows fail the comparisons)
            const float bound = (kAreas == 3) ? m4 : m3;
            const bool decided = (bound < thr) && (mask[0] != 0) && !force_exhaustive;
            const int row = rt * kRowsPerCta + row_in_cta;
            const bool in_range = row < T;
            const bool flag = in_range && !decided;
            // hand the verdict to the rescoring warps (double-buffered)
            const int hb = it & 1;
            VQ_TIMED_WAIT(1, h_empty(hb), (((uint32_t)(it >> 1)) & 1u) ^ 1u);
            int4* hand = reinterpret_cast<int4*>(smem + L.hand + hb * kHandBytes);
            hand[2 * row_in_cta] = make_int4(decided ? (g1 | (g2 << 16)) : -1, g3, (int)mask[0], (int)mask[1]);
            hand[2 * row_in_cta + 1] = make_int4((int)mask[2], 0, 0, 0);
            mbar_arrive(h_full(hb));
            if (flag) cand[row] = -1;
            const uint32_t ballot = __ballot_sync(VQ_FULL, flag);
            if (ballot) {
                int base = 0;
                if (lane == 0) base = atomicAdd(n_flagged, __popc(ballot));
                base = __shfl_sync(VQ_FULL, base, 0);
                if (flag) flagged[base + __popc(ballot & ((1u << lane) - 1))] = row;
                if (lane == 0 && stats)
                    atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_FALLBACK_ROWS),
                              (unsigned long long)__popc(ballot));
            }
        }
#ifdef VQ_TC_INSTRUMENT
        if (threadIdx.x == 256) {   // one epilogue thread per CTA
            atomicAdd((unsigned long long*)&g_tc_wait[8], (unsigned long long)wait_acc[0]);
            atomicAdd((unsigned long long*)&g_tc_wait[9], (unsigned long long)wait_acc[1]);
            atomicAdd((unsigned long long*)&g_tc_wait[10], (unsigned long long)(clock64() - t_begin));
        }
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace tc

namespace tc {
// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// (rows, D) fp16 row-major -> boxes of box_rows x 32 halfs, 64-byte swizzle, OOB rows read as zero
static bool make_map(CUtensorMap* map, const void* base, uint64_t rows, int D, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)D * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kKBlock, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc

// vq_dist_tc16.cu: D = 32 with fp16 accumulators and packed 16-bit maxima
bool tc16_supported(int64_t T, int K, int D);
cudaError_t launch_dist_tc16(const CUtensorMap& ma, const CUtensorMap& mb, int T, const float* zn32, const float* row_sq,
                             const CodebookView& cb, int* cand, int* flagged, int* n_flagged, int64_t* stats,
                             void* records, cudaStream_t s);
size_t tc16_workspace_bytes(int64_t T);

bool tc_supported(int64_t T, int K, int D) {
    const bool d_ok = (D == 32 || D == 64 || D == 128 || D == 256);
    // group ids travel as 15/16-bit fields
    return d_ok && K >= tc::kGroupCols && (K % tc::kGroupCols) == 0 && K / tc::kGroupCols <= 32767 && T >= 256;
}

size_t tc_workspace_bytes(int64_t T, int K, int D) {
    return tc16_supported(T, K, D) ? tc16_workspace_bytes(T) : 0;
}

template <int KB>
static cudaError_t launch_tc_kernel(const CUtensorMap& ma, const CUtensorMap& mb, int T, const float* zn32,
                                    const float* row_sq, const CodebookView& cb, int* cand, int* flagged,
                                    int* n_flagged, int64_t* stats, cudaStream_t s) {
    const tc::SmemLayout L = tc::smem_layout(KB);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc::k_dist_tc<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total + 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int n_row_tiles = (T + tc::kRowsPerCta - 1) / tc::kRowsPerCta;
    const int grid = n_row_tiles < sm_count() ? n_row_tiles : sm_count();
    static const int debug_flags = getenv("VQ_TC_DEBUG") ? atoi(getenv("VQ_TC_DEBUG")) : 0;
    tc::k_dist_tc<KB><<<grid, tc::kThreads, L.total + 1024, s>>>(ma, mb, T, cb.K, zn32, row_sq, cb.en32, cb.code_sq, cb.info,
                                                                 cand, flagged, n_flagged, stats, debug_flags);
    count_launch();
#ifdef VQ_TC_INSTRUMENT
    {
        cudaStreamSynchronize(s);
        long long w[16];
        cudaMemcpyFromSymbol(w, tc::g_tc_wait, sizeof(w));
        const double n = grid;
        printf("[tc instrument] per-CTA mean kcycles  producer: a_empty %.0f b_empty %.0f total %.0f | mma: a_full %.0f t_empty %.0f "
               "b_full %.0f total %.0f | epilogue: t_full %.0f h_empty %.0f total %.0f\n",
               w[0] / n / 1e3, w[1] / n / 1e3, w[2] / n / 1e3, w[4] / n / 1e3, w[5] / n / 1e3, w[6] / n / 1e3, w[7] / n / 1e3,
               w[8] / n / 1e3, w[9] / n / 1e3, w[10] / n / 1e3);
        long long z[16] = {0};
        cudaMemcpyToSymbol(tc::g_tc_wait, z, sizeof(z));
    }
#endif
    return cudaGetLastError();
}

cudaError_t launch_dist_tc(const __half* zn16, const float* zn32, const float* row_sq, const CodebookView& cb,
                           int64_t T, int* cand, int* flagged, int* n_flagged, int64_t* stats, void* tc_ws,
                           cudaStream_t s) {
    if (T == 0) return cudaSuccess;
    CUtensorMap ma, mb;
    if (tc16_supported(T, cb.K, cb.D)) {
        // one 256-row box per row tile, one 128-code box per n-tile
        if (!tc::make_map(&ma, zn16, (uint64_t)T, cb.D, 256) || !tc::make_map(&mb, cb.en16, (uint64_t)cb.K, cb.D, 128))
            return cudaErrorInvalidValue;
        return launch_dist_tc16(ma, mb, (int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, tc_ws, s);
    }
    if (!tc::make_map(&ma, zn16, (uint64_t)T, cb.D, tc::kRowsPerCta) ||
        !tc::make_map(&mb, cb.en16, (uint64_t)cb.K, cb.D, tc::kTileN))
        return cudaErrorInvalidValue;
    switch (cb.D / tc::kKBlock) {
        case 1: return launch_tc_kernel<1>(ma, mb, (int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, s);
        case 2: return launch_tc_kernel<2>(ma, mb, (int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, s);
        case 4: return launch_tc_kernel<4>(ma, mb, (int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, s);
        case 8: return launch_tc_kernel<8>(ma, mb, (int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace vq
