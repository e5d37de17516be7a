// Nearest-code search on the 5th-generation tensor cores (sm_100a): TMA -> shared memory ->
// tcgen05.mma (fp16 in, fp32 accumulators in TMEM) -> tcgen05.ld -> running maxima in registers,
// followed by exact fp32 rescoring of the few surviving candidates (k_rescore_g, a second launch).
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
__host__ __device__ constexpr int a_block_bytes(int halves) { return halves * 128 * kKBlock * 2; }
constexpr int kBStageBytes = kTileN * kKBlock * 2;        // 8 KiB
constexpr int kThreads = 512;        // warps 0-3 service (TMA, MMA, TMEM alloc, idle), 4-7 spare, 8-15 epilogue
constexpr int kRecordBytes = 32;     // verdict record per row
// register budget after setmaxnreg (the kernel launches with 128 per thread = the whole register file):
constexpr int kRegsService = 56, kRegsSpare = 24, kRegsEpilogue = 216;    // warps 0-3, 4-7 (and idle epilogue warps), 8-15
static_assert(128 * kRegsService + 128 * kRegsSpare + 256 * kRegsEpilogue <= 65536, "register file overcommitted");
constexpr float kTwoEps = 2.2e-3f;   // 2 * eps, see header comment
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
// c_format = F32 (bit 4), a/b = F16 (0), K-major both, N = 128 (bits 17..22), M = 128 (bits 24..28)

struct SmemLayout {
    uint32_t a, b, snap, hand, bars, tmem_slot, total;
};
__host__ __device__ constexpr int a_stages(int kb) { return kb <= 2 ? 2 : 1; }
// codebook ring.  At D = 256 the A operand of 256 rows takes 128 KB and leaves four 8 KB stages: each is consumed by 4 MMAs
// (~400 cycles) against a TMA round trip of ~700, so the issuers starve for tiles (2 -> 4 stages: 147 -> 126 us for cfg 2's
// search).  128-row CTAs (A = 64 KB, 16 stages) on their own are no way out: every CTA then streams the whole codebook
// for half the rows, 512 MB of L2 -> SM traffic for cfg 2, and the kernel sits at the ~6.2 TB/s the L2 delivers (82.8 vs
// 66.3 us, same box).  `halves` = 1 therefore runs as CLUSTERS OF TWO CTAs that share the codebook stream: each CTA
// fetches one 64-code half of every stage and TMA-multicasts it into both, a stage is released by both CTAs' issuers
// (multicast tcgen05.commit).  There are 16 stages: two rings of 8, one per issuer (the issuers take alternate code
// tiles; a ring of its own lets each issuer see every phase of its barriers - a parity wait by a thread that skipped
// the previous phase returns early).
#ifndef VQ_TC_BS256
#define VQ_TC_BS256 4      // 2 -> 4: 147 -> 126 us (cfg 2), 5.38 -> 4.31 ms (1 M tokens), same-box A/B with two issuers
#endif
__host__ __device__ constexpr int b_stages(int kb, int halves = 2) { return kb == 1 ? 8 : (kb <= 4 ? 4 : (halves == 1 ? 16 : VQ_TC_BS256)); }
// slot-maxima snapshots of the best three groups: 3 areas x 256 rows x 32 slots; rows padded by 16 B (conflict-free
// STS.128).  At D = 256 the A operand alone is 128 KB, so the snapshots are kept as fp16 (80-byte rows): the verdict
// then compares against thr - kSnapSlack, which covers the rounding of the stored maxima (|score| <= 1).
__host__ __device__ constexpr int snap_areas(int kb) { return 3; }
__host__ __device__ constexpr bool snap_half(int kb) { return kb > 4; }
__host__ __device__ constexpr int snap_row_bytes(int kb) { return kb <= 4 ? 144 : 80; }
constexpr float kSnapSlack = 5.0e-4f;   // >= 2^-11: fp16 rounding of a slot maximum of magnitude <= 1
constexpr int kMaxBStages = 16;
__host__ __device__ inline SmemLayout smem_layout(int kb, int halves = 2) {
    SmemLayout L;
    L.a = 0;
    L.b = L.a + a_stages(kb) * kb * a_block_bytes(halves);
    L.snap = L.b + b_stages(kb, halves) * kBStageBytes;
    L.hand = L.snap + snap_areas(kb) * (128 * halves) * snap_row_bytes(kb);
    L.bars = L.hand;
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

// One CTA per SM, persistent over row tiles.  KB = D / 32.
// warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM alloc   warps 8-15: epilogue, one thread per row, which
// ends with the row's verdict record in global memory.  Registers are redistributed with setmaxnreg.
// (Rescoring used to run on warps 4-7 of this kernel: at D = 256 its dependent 1 KB gathers held the tile hand-off
// and the tensor pipe sat at 20 %; it is now k_rescore_g, which reads whole 128-byte lines of cell copies.)
template <int KB, int HALVES>
__global__ void __launch_bounds__(kThreads, 1)
k_dist_tc(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int T, int K,
          const int* __restrict__ cb_info, int4* __restrict__ records, int* __restrict__ cand,
          int* __restrict__ flagged, int* __restrict__ n_flagged, int64_t* __restrict__ stats, int splits) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int AS = a_stages(KB);
    constexpr int BS = b_stages(KB, HALVES);
    constexpr int kRows = 128 * HALVES;               // token rows per CTA: two MMA row halves, or one
    constexpr int CL = HALVES == 1 ? 2 : 1;           // CTAs per cluster (launch attribute): they share the codebook stream
    // HALVES = 1 (see b_stages): two issuers on alternate code tiles, each with its own producer thread, ring of RING
    // stages and pair of accumulator stages; a full / empty barrier covers KS consecutive stages (k blocks) - a wait on
    // an mbarrier costs ~90 cycles even when its phase completed long ago, and at one wait per stage the producer and
    // the issuers spent more time on barriers than the two MMAs of a stage take (first cluster version: 91.5 us for cfg 2
    // against 66.4 with 256-row CTAs; tensor pipe 43 % busy, L2 18 %).
    constexpr int KS = HALVES == 1 ? 2 : 1;
    constexpr int RING = HALVES == 1 ? BS / 2 : BS;
    constexpr int NACC = HALVES == 1 ? 4 : 2;         // accumulator stages of HALVES x 128 columns
    static_assert(KB % KS == 0 && RING % KS == 0 && NACC * HALVES * kTileN <= 512, "stage geometry");
    constexpr int kABlock = a_block_bytes(HALVES);
    constexpr int kAreas = snap_areas(KB);
    constexpr int kSnapRow = snap_row_bytes(KB);
    constexpr bool kSnapHalf = snap_half(KB);
    const SmemLayout L = smem_layout(KB, HALVES);
    // swizzled TMA/UMMA tiles want a 1024-byte aligned base; the launch reserves 1 KiB of slack for this
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + L.bars;
    // barrier map (8 bytes each)
    auto b_full = [&](int s) { return bar_base + 8 * s; };
    auto b_empty = [&](int s) { return bar_base + 8 * (kMaxBStages + s); };
    auto t_full = [&](int s) { return bar_base + 8 * (2 * kMaxBStages + s); };
    auto t_empty = [&](int s) { return bar_base + 8 * (2 * kMaxBStages + 4 + s); };
    auto a_full = [&](int s) { return bar_base + 8 * (2 * kMaxBStages + 8 + s); };
    auto a_empty = [&](int s) { return bar_base + 8 * (2 * kMaxBStages + 10 + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L.tmem_slot);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // A work item is (row tile, code split): with few row tiles (a 16 384-token encode is 64 of them for 148 SMs) the
    // codebook is cut into `splits` ranges of whole groups and every range gets its own CTA, record and verdict.  A
    // verdict taken against the split's own best score keeps a superset of what the global best would keep, so the
    // union of the splits' surviving cells contains the winner whenever every split is decided.
    // With clusters a work item is taken by a cluster: its CTAs own consecutive row tiles (a tile past the end reads
    // zeros and writes nothing) and walk the same code tiles in step.
    const uint32_t cta_rank = CL == 2 ? cluster_cta_rank() : 0u;
    const int unit0 = blockIdx.x / CL, n_units = gridDim.x / CL;
    const int n_row_tiles = ((T + kRows - 1) / kRows + CL - 1) / CL;
    const int n_items = n_row_tiles * splits;
    const int n_tiles = K / kTileN / splits;            // per item
    const int n_groups = n_tiles / kGroupTiles;         // per item

    if (warp == 1 && lane == 0) {
        // two MMA issuers (one per row half) commit to the stage / tile / row-tile barriers
        // a stage is read by both row halves, or by its ring's issuer in either CTA of the cluster
        for (int s = 0; s < BS; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 2); }
        for (int s = 0; s < NACC; ++s) { mbar_init(t_full(s), HALVES); mbar_init(t_empty(s), 128 * HALVES); }
        for (int s = 0; s < 2; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 2); }
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
    if constexpr (CL == 2) cluster_sync_all();        // the peer's barriers are initialised before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // (with one row half only four epilogue warps take registers, so the service warps can keep more)
        reg_dec<HALVES == 1 ? 80 : kRegsService>();
        if ((warp == 0 || (HALVES == 1 && warp == 2)) && lane == 0) {
            // ===================== TMA producers: one, or one per issuer ring =====================
            const int pw = warp == 2 ? 1 : 0;
            const int n0 = HALVES == 2 ? 0 : pw, n_step = HALVES == 2 ? 1 : 2;
            int it = 0;
#ifdef VQ_TC_INSTRUMENT
            long long wait_acc[4] = {0, 0, 0, 0};
            const long long t_begin = clock64();
#endif
            for (int item = unit0; item < n_items; item += n_units, ++it) {
                const int rt = (item / splits) * CL + (int)cta_rank, tile0 = (item % splits) * n_tiles;
                if (pw == 0) {
                    const int as = it % AS;
                    VQ_TIMED_WAIT(0, a_empty(as), (((uint32_t)(it / AS)) & 1u) ^ 1u);
                    mbar_expect_tx(a_full(as), KB * kABlock);
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb)
                        tma_load_2d(smem_base + L.a + (as * KB + kb) * kABlock, &tm_a, a_full(as), kb * kKBlock,
                                    rt * kRows);
                }
                for (int n = n0; n < n_tiles; n += n_step) {
                    const uint32_t t_cnt = (uint32_t)(it * n_tiles + n);          // n_tiles is even: ring = n & 1
                    uint32_t ring_cnt = (HALVES == 2 ? t_cnt : (t_cnt >> 1)) * KB;  // stages this ring has carried so far
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb, ++ring_cnt) {
                        const int s = pw * RING + (int)(ring_cnt % RING);          // stage; its barrier is that of stage s - s % KS
                        const int sb = s - (kb % KS);
                        if (kb % KS == 0) {
                            VQ_TIMED_WAIT(1, b_empty(sb), ((ring_cnt / RING) & 1u) ^ 1u);
                            mbar_expect_tx(b_full(sb), KS * kBStageBytes);
                        }
                        if constexpr (CL == 2)      // this CTA's half of the stage (tm_b boxes are 64 codes), into both CTAs
                            tma_load_2d_multicast(smem_base + L.b + s * kBStageBytes + cta_rank * (kBStageBytes / 2), &tm_b, b_full(sb),
                                                  kb * kKBlock, (tile0 + n) * kTileN + (int)cta_rank * (kTileN / 2), (uint16_t)3);
                        else
                            tma_load_2d(smem_base + L.b + s * kBStageBytes, &tm_b, b_full(sb), kb * kKBlock, (tile0 + n) * kTileN);
                    }
                }
            }
#ifdef VQ_TC_INSTRUMENT
            if (pw == 0) {
                atomicAdd((unsigned long long*)&g_tc_wait[0], (unsigned long long)wait_acc[0]);
                atomicAdd((unsigned long long*)&g_tc_wait[1], (unsigned long long)wait_acc[1]);
                atomicAdd((unsigned long long*)&g_tc_wait[2], (unsigned long long)(clock64() - t_begin));
            }
#endif
        } else if ((warp == 1 || warp == 3) && lane == 0) {
            // ===================== MMA issuers: two threads =====================
            // A thread gets one tcgen05.mma out per ~105 cycles whatever its size, and issuers overlap
            // (tools/ubench_mma.cu): with a single issuer an M = N = 128, K = 16 MMA "retired" at half rate here.
            // HALVES = 2: one issuer per row half - the halves write different accumulators, both read every stage.
            // HALVES = 1: the issuers take alternate code tiles (= alternate accumulator stages) of the one row half.
            const int w = warp == 3 ? 1 : 0;
            const int r = HALVES == 2 ? w : 0;
            const int n0 = HALVES == 2 ? 0 : w, n_step = HALVES == 2 ? 1 : 2;
            int it = 0;
#ifdef VQ_TC_INSTRUMENT
            long long wait_acc[4] = {0, 0, 0, 0};
            const long long t_begin = clock64();
#endif
            for (int item = unit0; item < n_items; item += n_units, ++it) {
                const int as = it % AS;
                VQ_TIMED_WAIT(0, a_full(as), ((uint32_t)(it / AS)) & 1u);
                tc_fence_after();
                for (int n = n0; n < n_tiles; n += n_step) {
                    const uint32_t t_cnt = (uint32_t)(it * n_tiles + n);      // n_tiles is even: t_cnt & 1 = n & 1
                    const int acc = (int)(t_cnt % NACC);                       // HALVES = 1: issuer w owns stages w, w + 2
                    VQ_TIMED_WAIT(1, t_empty(acc), ((t_cnt / NACC) & 1u) ^ 1u);
                    tc_fence_after();
                    // ring position of the tile's first stage: HALVES = 2 - one ring, every tile; 1 - own ring, own tiles
                    uint32_t b_cnt = (HALVES == 2 ? t_cnt : (t_cnt >> 1)) * KB;
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb, ++b_cnt) {
                        const int s = (HALVES == 2 ? 0 : w * RING) + (int)(b_cnt % RING);
                        const int sb = s - (kb % KS);
                        if (kb % KS == 0) {
                            VQ_TIMED_WAIT(2, b_full(sb), (b_cnt / RING) & 1u);
                            tc_fence_after();
                        }
                        const uint32_t a_addr = smem_base + L.a + (as * KB + kb) * kABlock;
                        const uint32_t b_addr = smem_base + L.b + s * kBStageBytes;
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            umma_f16(tmem_base + (uint32_t)((acc * HALVES + r) * kTileN),
                                     umma_desc(a_addr + r * (128 * 64) + k * 32), umma_desc(b_addr + k * 32), kIdesc,
                                     (uint32_t)((kb | k) != 0));
                        // smem stages free once their readers' MMAs retire
                        if (kb % KS == KS - 1) {
                            if constexpr (CL == 2) umma_commit_multicast(b_empty(sb), (uint16_t)3);
                            else umma_commit(b_empty(sb));
                        }
                    }
                    umma_commit(t_full(acc));        // this issuer's accumulator of the tile complete
                }
                umma_commit(a_empty(as));            // row tile's A operand no longer needed by this issuer
            }
#ifdef VQ_TC_INSTRUMENT
            if (w == 0) {
                atomicAdd((unsigned long long*)&g_tc_wait[4], (unsigned long long)wait_acc[0]);
                atomicAdd((unsigned long long*)&g_tc_wait[5], (unsigned long long)wait_acc[1]);
                atomicAdd((unsigned long long*)&g_tc_wait[6], (unsigned long long)wait_acc[2]);
                atomicAdd((unsigned long long*)&g_tc_wait[7], (unsigned long long)(clock64() - t_begin));
            }
#endif
        }
    } else if (warp < 8 || (HALVES == 1 && warp >= 12)) {
        // spare warps: hand their registers to the epilogue (setmaxnreg.inc waits for them) and leave
        reg_dec<kRegsSpare>();
    } else {
        // ===================== epilogue: 8 warps, one thread per row =====================
        reg_inc<kRegsEpilogue>();
        const int e = warp - 8;
        const int r_sub = e >> 2;                    // which 128-row MMA tile
        const int quarter = warp & 3;                // TMEM lane quarter this warp may read
        const int row_in_cta = r_sub * 128 + quarter * 32 + lane;
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(r_sub * kTileN);
        constexpr uint32_t kStage = HALVES * kTileN;     // TMEM columns of an accumulator stage
        const uint32_t snap0 = smem_base + L.snap + (uint32_t)row_in_cta * kSnapRow;
        constexpr uint32_t kSnapArea = kRows * kSnapRow;
        const bool force_exhaustive = codebook_degenerate(cb_info);
        // a group (2 tiles) drains accumulator stages sp, sp + 1: group count gc -> sp = 2 gc mod NACC, parity of its use
        uint32_t gc = 0;
        auto sp_of = [&](uint32_t c) { return (int)((2u * c) % NACC); };
        auto ph_of = [&](uint32_t c) { return ((2u * c) / NACC) & 1u; };
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
        for (int item = unit0; item < n_items; item += n_units, ++it) {
            const int rt = (item / splits) * CL + (int)cta_rank, split = item % splits;
            const int group0 = split * n_groups;        // global id of the item's first group
            float slot[32];
            float m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY, m4 = -INFINITY;
            int g1 = 0, g2 = 0, g3 = 0;
            uint32_t a1 = 0, a2 = 1, a3 = (kAreas == 3) ? 2 : 1;   // snapshot areas of the best / second / third group
            float buf[2][64];
            VQ_TIMED_WAIT(0, t_full(sp_of(gc)), ph_of(gc));
            tc_fence_after();
            load_batch(tbase + (uint32_t)sp_of(gc) * kStage, buf[0]);
            for (int g = 0; g < n_groups; ++g, ++gc) {
                const int sp = sp_of(gc);
                const uint32_t phase = ph_of(gc);
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    tmem_ld_wait();                                   // batch b has landed in buf[b & 1]
                    if (b == 1) { tc_fence_before(); mbar_arrive(t_empty(sp)); }   // tile 0 fully in registers
                    if (b == 3) { tc_fence_before(); mbar_arrive(t_empty(sp + 1)); }
                    if (b < 3) {
                        if (b == 1) { VQ_TIMED_WAIT(0, t_full(sp + 1), phase); tc_fence_after(); }
                        load_batch(tbase + (uint32_t)((sp + ((b + 1) >> 1)) * kStage + ((b + 1) & 1) * 64), buf[(b + 1) & 1]);
                    } else if (g + 1 < n_groups) {
                        VQ_TIMED_WAIT(0, t_full(sp_of(gc + 1)), ph_of(gc + 1));
                        tc_fence_after();
                        load_batch(tbase + (uint32_t)sp_of(gc + 1) * kStage, buf[0]);
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
                    if constexpr (kSnapHalf) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint32_t h[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const __half2 p = __floats2half2_rn(slot[8 * q + 2 * j], slot[8 * q + 2 * j + 1]);
                                h[j] = *reinterpret_cast<const uint32_t*>(&p);
                            }
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16 * q), "r"(h[0]), "r"(h[1]),
                                         "r"(h[2]), "r"(h[3])
                                         : "memory");
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16 * q), "f"(slot[4 * q]),
                                         "f"(slot[4 * q + 1]), "f"(slot[4 * q + 2]), "f"(slot[4 * q + 3])
                                         : "memory");
                    }
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
                    if constexpr (kSnapHalf) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint32_t h[4];
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3])
                                         : "r"(src + 16 * q));
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
                                kept[8 * q + 2 * j] = f.x; kept[8 * q + 2 * j + 1] = f.y;
                            }
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(kept[4 * q]), "=f"(kept[4 * q + 1]), "=f"(kept[4 * q + 2]), "=f"(kept[4 * q + 3])
                                         : "r"(src + 16 * q));
                    }
                    const float thr_slot = kSnapHalf ? thr - kSnapSlack : thr;
#pragma unroll
                    for (int j = 0; j < 32; ++j) mask[a] |= (kept[j] >= thr_slot) ? (1u << j) : 0u;
                }
            }
            // decided iff no further group can hold the winner (NaN / -inf rows fail the comparisons)
            const float bound = (kAreas == 3) ? m4 : m3;
            const bool decided = (bound < thr) && (mask[0] != 0) && !force_exhaustive;
            const int row = rt * kRows + row_in_cta;
            const bool in_range = row < T;
            const bool flag = in_range && !decided;
            // the verdict record of the row, for the exact rescoring kernel: {g1 | g2 << 16 (or -1: undecided), g3,
            // mask1, mask2} {mask3, -, -, -}
            if (in_range) {
                int4* rec = records + 2 * ((int64_t)split * T + row);
                rec[0] = make_int4(decided ? ((group0 + g1) | ((group0 + g2) << 16)) : -1, group0 + g3, (int)mask[0], (int)mask[1]);
                rec[1] = make_int4((int)mask[2], __float_as_int(m1), 0, 0);     // m1: the split's best approximate score
            }
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
    if constexpr (CL == 2) cluster_sync_all();        // nothing of the peer's (multicast tiles, stage releases) is still on its way
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Exact fp32 rescoring behind the filter, any D: cand[row] = argmin over the codes of the row's surviving cells.
// Same shape as k_exact_finish16's phase A: a warp owns 4 rows per iteration, staged in shared memory; the 8 lanes of
// a group each own one member of a cell and read whole 128-byte lines of the generic cell copies (CodebookView::en32c,
// kind 2) in batches of 8 chunks; every lane runs the sequential fma chain over d = 0..D-1 of the exhaustive search,
// so all paths return identical indices.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRescoreThreads = 128;
constexpr int kMaxSplits = 4;
constexpr int kFlaggedSlices = kFewFlaggedSlices;
struct __align__(16) FlaggedPartial {
    unsigned long long best; float second; float pad;
};
// exact distance of the staged row zs to member m of generic cell ci: sequential fma chain over d = 0..D-1
template <int D>
__device__ __forceinline__ float cell_distance_g(const float4* __restrict__ en32c, const float* __restrict__ csq_cell, int ci,
                                                 int m, const float4* zs, float a_sq) {
    constexpr int kChunks = D / 4;
    const float4* e4 = en32c + (int64_t)ci * kChunks * 8 + m;
    const float csq = __ldg(csq_cell + ci * 8 + m);
    float dot = 0.f;
#ifndef VQ_RESCORE_UNROLL
#define VQ_RESCORE_UNROLL 4   // 1: 64 us, 2: 60 us, 4: 50 us for the rescoring of a 16 384-token 8192 x 256 encode
#endif
    constexpr int kBatchUnroll = VQ_RESCORE_UNROLL;    // batches of 8 chunks the compiler may overlap
#pragma unroll kBatchUnroll
    for (int q0 = 0; q0 < kChunks; q0 += 8) {
        float4 ev[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) ev[q] = __ldg(e4 + (q0 + q) * 8);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 z = zs[q0 + q];
            dot = __fmaf_rn(z.x, ev[q].x, dot);
            dot = __fmaf_rn(z.y, ev[q].y, dot);
            dot = __fmaf_rn(z.z, ev[q].z, dot);
            dot = __fmaf_rn(z.w, ev[q].w, dot);
        }
    }
    return ref_distance(a_sq, csq, dot);
}

// (Measured and rejected, three times: screening a cell's 8 members with their fp16 values first (512 B per code instead
// of 1 KB, exact chain only for members within 2 eps of the row's best approximate score).  From the row-major fp16
// codes a lane walks its own 512-byte row - 32 lines per warp-level load instead of 4: 55 vs 31 us for cfg 2, 1.83 vs
// 0.66 ms for 1 M tokens.  From coalesced fp16 cell copies (whole lines, like the fp32 ones): 46 vs 31 us, 1.07 vs 0.66
// ms - the screen and the exact chain are two dependent load -> fma-chain phases per cell, and the kernel is bound by
// that per-warp latency, not by bytes.)
// Without a minimum-blocks bound ptxas settles on 48-64 registers for this kernel and serialises the 32 loads a batch
// wants in flight; 4 blocks per SM = up to 128 registers keeps them all outstanding.
#ifndef VQ_RESCORE_MINBLOCKS
#define VQ_RESCORE_MINBLOCKS 4
#endif
template <int D, int kS>      // kS: code splits the records may come from (1, or kMaxSplits)
__global__ void __launch_bounds__(kRescoreThreads, VQ_RESCORE_MINBLOCKS)
k_rescore_g(const int4* __restrict__ rec, const float* __restrict__ zn32, const float* __restrict__ row_sq,
            const float4* __restrict__ en32c, const float* __restrict__ csq_cell, int T, int K, int splits,
            const int* __restrict__ flagged, const int* __restrict__ n_flagged, FlaggedPartial* __restrict__ partial,
            int* __restrict__ done, int* __restrict__ cand, int64_t* __restrict__ stats) {
    constexpr int kChunks = D / 4;
    __shared__ __align__(16) float4 s_z[kRescoreThreads / 32][4][kChunks + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = lane & 7, grp = lane >> 3;
    const float4* zn4 = reinterpret_cast<const float4*>(zn32);
    unsigned ties = 0, multi = 0;
    const int groups = gridDim.x * (kRescoreThreads / 8);
    for (int row0 = (blockIdx.x * kRescoreThreads + threadIdx.x - lane) >> 3; row0 < T; row0 += groups) {
        const int row = row0 + grp;
        // the row's verdict records, one per code split (splits <= kMaxSplits); undecided in any split: not ours
        int gs[3 * kS];
        uint32_t ms[3 * kS];
        float best_of[kS], gmax = -INFINITY;
        bool valid = row < T;
        float a_sq = 0.f;
        if (row < T) a_sq = __ldg(row_sq + row);
#pragma unroll
        for (int sp = 0; sp < kS; ++sp) {
            int4 h0 = make_int4(-1, 0, 0, 0), h1 = make_int4(0, 0, 0, 0);
            if (sp < splits && row < T) {
                h0 = __ldg(rec + 2 * ((int64_t)sp * T + row));
                h1 = __ldg(rec + 2 * ((int64_t)sp * T + row) + 1);
            }
            if (sp < splits) valid = valid && (h0.x >= 0);
            best_of[sp] = (sp < splits) ? __int_as_float(h1.y) : -INFINITY;
            gmax = fmaxf(gmax, best_of[sp]);
            gs[3 * sp] = h0.x & 0xFFFF; gs[3 * sp + 1] = (h0.x >> 16) & 0x7FFF; gs[3 * sp + 2] = h0.y;
            ms[3 * sp] = (sp < splits) ? (uint32_t)h0.z : 0u;
            ms[3 * sp + 1] = (sp < splits) ? (uint32_t)h0.w : 0u;
            ms[3 * sp + 2] = (sp < splits) ? (uint32_t)h1.x : 0u;
        }
        __syncwarp();
        // the warp's 4 rows are 4 * kChunks contiguous float4
#pragma unroll
        for (int j = 0; j < kChunks / 8; ++j) {
            const int e = lane + 32 * j;                    // float4 index within the 4 rows
            const int r = e / kChunks, c = e % kChunks;
            if (row0 + r < T) s_z[warp][r][c] = __ldg(zn4 + (int64_t)row0 * kChunks + e);
        }
        __syncwarp();
        const float4* zs = s_z[warp][grp];
        // a split whose best approximate score is more than 2 eps below the best of all splits cannot hold the winner
        int n_cells = 0;
#pragma unroll
        for (int i = 0; i < 3 * kS; ++i) {
            if (!valid || best_of[i / 3] < gmax - kTwoEps) ms[i] = 0u;
            n_cells += __popc(ms[i]);
        }
        const int n_iter = __reduce_max_sync(VQ_FULL, n_cells);
        Top2 top;
        top.init();
#pragma unroll 1
        for (int it = 0; it < n_iter; ++it) {
            // first non-empty mask (static indexing keeps gs / ms in registers)
            int g = -1, slot = 0;
#pragma unroll
            for (int i = 0; i < 3 * kS; ++i)
                if (g < 0 && ms[i] != 0u) {
                    slot = __ffs(ms[i]) - 1;
                    ms[i] &= ms[i] - 1;
                    g = gs[i];
                }
            if (g >= 0) {
                const int ci = g * 32 + slot;
                const int code = g * kGroupCols + 64 * (m >> 1) + ((slot & 16) << 1) + (slot & 15) + 16 * (m & 1);
                top.add(dist_key(cell_distance_g<D>(en32c, csq_cell, ci, m, zs, a_sq), code));
            }
        }
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) {
            const unsigned long long ob = __shfl_xor_sync(VQ_FULL, top.best, off);
            const float os = __shfl_xor_sync(VQ_FULL, top.second, off);
            top.merge(ob, os);
        }
        if (valid && m == 0) {
            const float bd = key_dist(top.best);
            cand[row] = (int)(uint32_t)top.best | kCandExactBit;
            if (top.second - bd < VQ_NEAR_TIE_REL * fabsf(bd)) ++ties;
            if (n_cells > 1) ++multi;
        }
    }
    // ---- the rows the filter could not decide, when they are few (latency matters, not throughput).  A work item is
    // (batch of 16 listed rows, slice of ALL cells): the block stages the 16 rows, group gi of 8 lanes scores row gi
    // against every cell of the slice - the 16 groups walk the same cells together, so a cell comes out of L2 once per
    // batch (one row per item, as it was, read the whole 8 MB of cell copies per listed row: 224 MB for the 28 rows a
    // cfg 2 encode leaves undecided, 60 % of this kernel's L2 traffic).  The last item of a batch to finish folds the
    // partial (best, second) pairs of its rows.  Longer lists are left to the tiled exhaustive kernel (k_scan_exact).
    {
        const int n = *n_flagged;
        if (n > 0 && n <= kFewFlagged) {
            constexpr int kBatchRows = kRescoreThreads / 8;
            __shared__ int s_last;
            float4 (*zrow)[kChunks + 1] = reinterpret_cast<float4 (*)[kChunks + 1]>(&s_z[0][0][0]);
            const int bgrp = threadIdx.x >> 3;
            const int n_batches = (n + kBatchRows - 1) / kBatchRows;
            const int n_cells = K / 8;
            const int per_slice = (n_cells + kFlaggedSlices - 1) / kFlaggedSlices;
            for (int item = blockIdx.x; item < n_batches * kFlaggedSlices; item += gridDim.x) {
                const int batch = item / kFlaggedSlices, slice = item % kFlaggedSlices;
                __syncthreads();                       // the previous item's shared values are consumed
                for (int e = threadIdx.x; e < kBatchRows * kChunks; e += kRescoreThreads) {
                    const int r = e / kChunks, c = e % kChunks;
                    const int i = batch * kBatchRows + r;
                    if (i < n) zrow[r][c] = __ldg(zn4 + (int64_t)flagged[i] * kChunks + c);
                }
                __syncthreads();
                const int i = batch * kBatchRows + bgrp;
                const bool have = i < n;
                Top2 top;
                top.init();
                if (have) {
                    const float a_sq = __ldg(row_sq + flagged[i]);
                    const int c_end = min(n_cells, (slice + 1) * per_slice);
                    for (int ci = slice * per_slice; ci < c_end; ++ci) {
                        const int g = ci >> 5, slot = ci & 31;
                        const int code = g * kGroupCols + 64 * (m >> 1) + ((slot & 16) << 1) + (slot & 15) + 16 * (m & 1);
                        top.add(dist_key(cell_distance_g<D>(en32c, csq_cell, ci, m, zrow[bgrp], a_sq), code));
                    }
                }
#pragma unroll
                for (int off = 4; off > 0; off >>= 1) {
                    const unsigned long long ob = __shfl_xor_sync(VQ_FULL, top.best, off);
                    const float os = __shfl_xor_sync(VQ_FULL, top.second, off);
                    top.merge(ob, os);
                }
                if (have && m == 0) {
                    FlaggedPartial pp;
                    pp.best = top.best; pp.second = top.second; pp.pad = 0.f;
                    partial[(int64_t)i * kFlaggedSlices + slice] = pp;
                }
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0) s_last = (atomicAdd(done + batch, 1) == kFlaggedSlices - 1);
                __syncthreads();
                if (s_last) {
                    __threadfence();
                    // warp w folds rows 4w .. 4w + 3 of the batch: lane = slices lane, lane + 32, ...
                    for (int rr = 0; rr < kBatchRows / (kRescoreThreads / 32); ++rr) {
                        const int fi = batch * kBatchRows + warp * (kBatchRows / (kRescoreThreads / 32)) + rr;
                        if (fi >= n) break;
                        Top2 fin;
                        fin.init();
                        for (int q = lane; q < kFlaggedSlices; q += 32) {
                            const FlaggedPartial* pq = partial + (int64_t)fi * kFlaggedSlices + q;
                            fin.merge(__ldcg(&pq->best), __ldcg(&pq->second));
                        }
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) {
                            const unsigned long long ob = __shfl_xor_sync(VQ_FULL, fin.best, off);
                            const float os = __shfl_xor_sync(VQ_FULL, fin.second, off);
                            fin.merge(ob, os);
                        }
                        if (lane == 0) {
                            const float bd = key_dist(fin.best);
                            cand[flagged[fi]] = (int)(uint32_t)fin.best | kCandExactBit;
                            if (fin.second - bd < VQ_NEAR_TIE_REL * fabsf(bd)) ++ties;
                        }
                    }
                    if (threadIdx.x == 0) done[batch] = 0;     // ready for the next call
                }
            }
        }
    }
    if (stats) {
        ties = __reduce_add_sync(VQ_FULL, ties);
        multi = __reduce_add_sync(VQ_FULL, multi);
        if (lane == 0) {
            if (ties) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_NEAR_TIE_ROWS), (unsigned long long)ties);
            if (multi) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_AMBIGUOUS_ROWS), (unsigned long long)multi);
        }
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
int tc16_max_clusters();
cudaError_t launch_dist_tc16(const CUtensorMap& ma, const CUtensorMap& mb, int T, const __half* zn16, const float* zn32,
                             const float* row_sq, const CodebookView& cb, int* cand, int* flagged, int* n_flagged, int64_t* stats,
                             void* records, cudaStream_t s);
size_t tc16_workspace_bytes(int64_t T);

bool tc_supported(int64_t T, int K, int D) {
    const bool d_ok = (D == 32 || D == 64 || D == 128 || D == 256);
    // group ids travel as 15/16-bit fields
    return d_ok && K >= tc::kGroupCols && (K % tc::kGroupCols) == 0 && K / tc::kGroupCols <= 32767 && T >= 256;
}

// D = 256: clusters of two 128-row CTAs sharing the codebook stream (k_dist_tc<8, 1>), when the device can co-schedule
// them.  Returns how many such clusters fit the device at once (0: use the 256-row kernel).  VQ_TC_CLUSTER=0 disables.
static int tc_max_clusters() {
    static PerDeviceOnce once;
    static int n_of[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (once.need()) {
        const char* e = getenv("VQ_TC_CLUSTER");
        int n = 0;
        if (!(e && e[0] == '0')) {
            const tc::SmemLayout L = tc::smem_layout(8, 1);
            if (cudaFuncSetAttribute(tc::k_dist_tc<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total + 1024) == cudaSuccess) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(2 * sm_count());
                cfg.blockDim = dim3(tc::kThreads);
                cfg.dynamicSmemBytes = L.total + 1024;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                if (cudaOccupancyMaxActiveClusters(&n, tc::k_dist_tc<8, 1>, &cfg) != cudaSuccess) n = 0;
            }
            cudaGetLastError();
        }
        n_of[dev & 63] = n;
    }
    return n_of[dev & 63];
}

// geometry of a generic-filter call: rows per CTA, CTAs per cluster, code splits (as many as keep row units x splits
// within one wave), grid
struct TcGeom {
    int rows, cl, splits, grid;
};
static TcGeom tc_geom(int64_t T, int K, int D) {
    TcGeom g;
    const int clusters = (D == 256) ? tc_max_clusters() : 0;
    g.cl = clusters > 0 ? 2 : 1;
    g.rows = g.cl == 2 ? 128 : tc::kRowsPerCta;
    const int64_t units = ((T + g.rows - 1) / g.rows + g.cl - 1) / g.cl;
    const int wave = g.cl == 2 ? clusters : sm_count();
    const int n_groups = K / tc::kGroupCols;
    g.splits = tc::kMaxSplits;
    while (g.splits > 1 && (n_groups % g.splits != 0 || units * g.splits > wave)) g.splits >>= 1;
    const int64_t items = units * g.splits;
    g.grid = (int)(items < wave ? items : wave) * g.cl;
    return g;
}
static int tc_splits(int64_t T, int K, int D) { return tc_geom(T, K, D).splits; }

// a row can be listed once per split
int tc_flag_multiplier(int64_t T, int K, int D) {
    return (tc_supported(T, K, D) && !tc16_supported(T, K, D)) ? tc_splits(T, K, D) : 1;
}

size_t tc_workspace_bytes(int64_t T, int K, int D) {
    if (tc16_supported(T, K, D)) return tc16_workspace_bytes(T);
    return tc_supported(T, K, D) ? (size_t)(T > 0 ? T : 1) * tc::kRecordBytes * tc_splits(T, K, D) : 0;
}

template <int D>
static cudaError_t launch_rescore_d(int T, const float* zn32, const float* row_sq, const CodebookView& cb, int* cand,
                                    const int* flagged, int* n_flagged, int64_t* stats, const void* records, void* partial_ws,
                                    cudaStream_t s) {
    const int splits = tc_splits(T, cb.K, D);
    const int rows_per_block = tc::kRescoreThreads / 8;
    int64_t blocks = ((int64_t)T + rows_per_block - 1) / rows_per_block;
    const int64_t cap = (int64_t)sm_count() * 16 * 2;
    if (blocks > cap) blocks = cap;
    if (splits == 1)
        tc::k_rescore_g<D, 1><<<(unsigned)blocks, tc::kRescoreThreads, 0, s>>>(
            static_cast<const int4*>(records), zn32, row_sq, reinterpret_cast<const float4*>(cb.en32c), cb.csq_cell, T, cb.K,
            splits, flagged, n_flagged, static_cast<tc::FlaggedPartial*>(partial_ws), n_flagged + 64, cand, stats);
    else
        tc::k_rescore_g<D, tc::kMaxSplits><<<(unsigned)blocks, tc::kRescoreThreads, 0, s>>>(
            static_cast<const int4*>(records), zn32, row_sq, reinterpret_cast<const float4*>(cb.en32c), cb.csq_cell, T, cb.K,
            splits, flagged, n_flagged, static_cast<tc::FlaggedPartial*>(partial_ws), n_flagged + 64, cand, stats);
    count_launch();
    return cudaGetLastError();
}

// the exact rescoring behind the generic filter (launch_dist_tc with !tc16_supported): cand[] for every decided row,
// and for the undecided ones when at most kFewFlagged are listed
cudaError_t launch_rescore_generic(const float* zn32, const float* row_sq, const CodebookView& cb, int64_t T, int* cand,
                                   const int* flagged, int* n_flagged, int64_t* stats, const void* tc_ws, void* partial_ws,
                                   cudaStream_t s) {
    if (T == 0) return cudaSuccess;
    switch (cb.D) {
        case 32: return launch_rescore_d<32>((int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, tc_ws, partial_ws, s);
        case 64: return launch_rescore_d<64>((int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, tc_ws, partial_ws, s);
        case 128: return launch_rescore_d<128>((int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, tc_ws, partial_ws, s);
        case 256: return launch_rescore_d<256>((int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, tc_ws, partial_ws, s);
        default: return cudaErrorInvalidValue;
    }
}

template <int KB, int HALVES>
static cudaError_t launch_tc_kernel(const CUtensorMap& ma, const CUtensorMap& mb, int T, const float* zn32,
                                    const float* row_sq, const CodebookView& cb, int* cand, int* flagged,
                                    int* n_flagged, int64_t* stats, void* records, void* partial_ws, cudaStream_t s) {
    const tc::SmemLayout L = tc::smem_layout(KB, HALVES);
    static PerDeviceOnce once;
    if (once.need()) {
        cudaError_t e = cudaFuncSetAttribute(tc::k_dist_tc<KB, HALVES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total + 1024);
        if (e != cudaSuccess) return e;
    }
    const TcGeom geom = tc_geom(T, cb.K, KB * tc::kKBlock);
    const int splits = geom.splits, grid = geom.grid;
    if (geom.cl != (HALVES == 1 ? 2 : 1)) return cudaErrorInvalidValue;
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(tc::kThreads);
        cfg.dynamicSmemBytes = L.total + 1024;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)geom.cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, tc::k_dist_tc<KB, HALVES>, ma, mb, T, cb.K, (const int*)cb.info,
                                           static_cast<int4*>(records), cand, flagged, n_flagged, stats, splits);
        if (e != cudaSuccess) return e;
    }
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
                           void* partial_ws, cudaStream_t s) {
    if (T == 0) return cudaSuccess;
    CUtensorMap ma, mb;
    if (tc16_supported(T, cb.K, cb.D)) {
        // one 256-row box per row tile, one 128-code box per n-tile; clusters of two CTAs: 128 rows, and each CTA fetches
        // (and multicasts) one 64-code half of every n-tile
        const bool cl = tc16_max_clusters() > 0;
        if (!tc::make_map(&ma, zn16, (uint64_t)T, cb.D, cl ? 128 : 256) ||
            !tc::make_map(&mb, cb.en16, (uint64_t)cb.K, cb.D, cl ? 64 : 128))
            return cudaErrorInvalidValue;
        return launch_dist_tc16(ma, mb, (int)T, zn16, zn32, row_sq, cb, cand, flagged, n_flagged, stats, tc_ws, s);
    }
    // clusters: a CTA fetches (and multicasts) one 64-code half of every codebook stage
    const TcGeom geom = tc_geom(T, cb.K, cb.D);
    const int rows = geom.rows;
    if (!tc::make_map(&ma, zn16, (uint64_t)T, cb.D, rows) ||
        !tc::make_map(&mb, cb.en16, (uint64_t)cb.K, cb.D, geom.cl == 2 ? tc::kTileN / 2 : tc::kTileN))
        return cudaErrorInvalidValue;
    switch (cb.D / tc::kKBlock) {
        case 1: return launch_tc_kernel<1, 2>(ma, mb, (int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, tc_ws, partial_ws, s);
        case 2: return launch_tc_kernel<2, 2>(ma, mb, (int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, tc_ws, partial_ws, s);
        case 4: return launch_tc_kernel<4, 2>(ma, mb, (int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, tc_ws, partial_ws, s);
        case 8:
            if (rows == 128)
                return launch_tc_kernel<8, 1>(ma, mb, (int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, tc_ws, partial_ws, s);
            return launch_tc_kernel<8, 2>(ma, mb, (int)T, zn32, row_sq, cb, cand, flagged, n_flagged, stats, tc_ws, partial_ws, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace vq
