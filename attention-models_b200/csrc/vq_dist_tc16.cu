// Nearest-code search for codebook_dim = 32 on the 5th-generation tensor cores with **fp16 accumulators**.
//
// Replaces the dense part of the reference's search (paths relative to /root/reference):
//   models/vitvqgan.py:157-161   d = sum(z^2) + sum(e^2) - 2 * einsum('bd,nd->bn', z, e);  argmin(d, dim=1)
//
// Why a second kernel.  At D = 32 a 128 x 128 accumulator tile costs the tensor core only 128 cycles but
// holds 16 384 scores that have to be drained from TMEM and compared.  FMNMX / VIMNMX issue at 64 lanes per
// clock per SM (profiles/r01_ubench_alu_tmem.txt), so the fp32 epilogue of vq_dist_tc.cu (FMNMX3, 2 scores per
// ALU op) needs as many ALU cycles as the MMA needs tensor cycles and ends up the longer stage.  Here the MMA
// writes fp16 accumulators (instruction descriptor c_format = F16), tcgen05.ld ... .pack::16b delivers two
// scores per register, and VIMNMX3.S16x2 -- a packed signed-16-bit three-input max -- folds FOUR scores per
// ALU op: non-negative fp16 bit patterns order like integers, negative ones sort below every non-negative
// one, which is all a running maximum needs as long as the row's best score is positive (rows where it is
// not are handed to the exhaustive search).
//
// Exactness is unchanged: the tensor-core scores are a filter, every survivor is rescored with the
// reference's fp32 formula (same fma chain as vq_dist_simt.cu), rows the filter cannot decide go to the
// exhaustive search.  eps covers fp16 rounding of both operands plus the two fp16 roundings of the
// accumulator (one per K = 16 MMA).
//
// Shape of the computation: one persistent CTA per SM, 256 token rows per CTA (two M = 128 MMA row tiles),
// codebook streamed by TMA in 256-code stages (two n-tiles of 128 codes), 2 x 2 accumulator tiles in TMEM.
// A "group" is 4 n-tiles = 512 codes; a row keeps 64 slot maxima per group (32 packed registers; slot hs
// covers codes g*512 + hs + 64*m, m = 0..7 -- a "cell"), the maxima of the best three groups are parked in
// shared memory, and at the end of the row tile every cell within 2*eps of the row's best score is rescored.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>

#include "../../include/vq_b200.h"
#include "vq_common.cuh"
#include "vq_kernels.h"
#include "vq_tc_common.cuh"

namespace vq {
namespace tc16 {

using namespace tc;

constexpr int kD = 32;
constexpr int kRowsPerCta = 256;                    // two MMA row tiles of 128
constexpr int kTileN = 128;                         // codes per accumulator stage
constexpr int kGroupTiles = 4;
constexpr int kGroupCols = kTileN * kGroupTiles;    // 512 codes per group
constexpr int kCellCodes = 8;                       // codes per (group, slot) cell
constexpr int kABytes = kRowsPerCta * kD * 2;       // 16 KiB: one row tile of fp16 unit rows
constexpr int kBStageCodes = 2 * kTileN;            // one TMA = two n-tiles
constexpr int kBStageBytes = kBStageCodes * kD * 2; // 16 KiB
constexpr int kBStages = 4;
constexpr int kAStages = 2;
constexpr int kThreads = 512;
constexpr int kRescoreThreads = 128;
constexpr int kRegsService = 40, kRegsRescore = 104, kRegsEpilogue = 184;
static_assert(128 * kRegsService + 128 * kRegsRescore + 256 * kRegsEpilogue <= 65536, "register file overcommitted");
// |fp16-pipeline score - exact dot| <= eps: 1.1e-3 for the fp16 operands (as in vq_dist_tc.cu) + one fp16 ulp
// per accumulator rounding (2^-11 below 1, 2^-10 in [1, 2)); the threshold is additionally rounded down to fp16
constexpr float kTwoEps = 2.f * (1.1e-3f + 4.8829e-4f + 9.7657e-4f);
constexpr float kMinThreshold = 1e-3f;              // the filter only trusts rows whose threshold is safely positive
// c_format = F16 (0), a/b = F16 (0), K-major both, N = 128, M = 128
constexpr uint32_t kIdesc = ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

constexpr int kSnapRow = 144;                       // 32 packed slot registers + 16 B pad (conflict-free STS.128)
constexpr int kSnapArea = kRowsPerCta * kSnapRow;
constexpr int kHandBytes = kRowsPerCta * 32;        // per row: {g1|g2<<16 or -1, g3, mask0 lo, hi} {mask1 lo, hi, mask2 lo, hi}

struct SmemLayout {
    uint32_t a, b, snap, hand, bars, tmem_slot, total;
};
__host__ __device__ inline SmemLayout smem_layout() {
    SmemLayout L;
    L.a = 0;
    L.b = L.a + kAStages * kABytes;
    L.snap = L.b + kBStages * kBStageBytes;
    L.hand = L.snap + 3 * kSnapArea;
    L.bars = L.hand + 2 * kHandBytes;
    L.tmem_slot = L.bars + 8 * 32;
    L.total = L.tmem_slot + 16;
    return L;
}

// 32 lanes x 128 columns of fp16 accumulators, two adjacent columns per register (low half = even column)
__device__ __forceinline__ void tmem_ld_tile(uint32_t taddr, uint32_t (&v)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
          "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
          "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
          "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
          "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr));
}

__device__ __forceinline__ int lo16(uint32_t p) { return (int)(p << 16) >> 16; }
__device__ __forceinline__ int hi16(uint32_t p) { return (int)p >> 16; }

// exact fp32 distances of one row to the 8 codes of cell (g, hs).  Per code the dot product is the same
// sequential fma chain over d = 0..31 as the exhaustive search; the 8 chains are interleaved so that the 8
// code rows (8 different 128-byte lines) are fetched with one L2 round trip instead of eight.
__device__ __forceinline__ void rescore_cell(int g, int hs, const float4 (&z)[kD / 4], float a_sq,
                                             const float* __restrict__ en32, const float* __restrict__ code_sq,
                                             float& best_d, int& best_i, float& second_d) {
    const int code0 = g * kGroupCols + hs;
    const float4* e4 = reinterpret_cast<const float4*>(en32 + (int64_t)code0 * kD);
    float csq[kCellCodes], dot[kCellCodes];
#pragma unroll
    for (int m = 0; m < kCellCodes; ++m) { csq[m] = __ldg(code_sq + code0 + 64 * m); dot[m] = 0.f; }
#pragma unroll
    for (int q = 0; q < kD / 4; ++q) {
        float4 ev[kCellCodes];
#pragma unroll
        for (int m = 0; m < kCellCodes; ++m) ev[m] = __ldg(e4 + m * (64 * kD / 4) + q);
#pragma unroll
        for (int m = 0; m < kCellCodes; ++m) {
            dot[m] = __fmaf_rn(z[q].x, ev[m].x, dot[m]);
            dot[m] = __fmaf_rn(z[q].y, ev[m].y, dot[m]);
            dot[m] = __fmaf_rn(z[q].z, ev[m].z, dot[m]);
            dot[m] = __fmaf_rn(z[q].w, ev[m].w, dot[m]);
        }
    }
#pragma unroll
    for (int m = 0; m < kCellCodes; ++m) {
        const int code = code0 + 64 * m;
        const float dist = ref_distance(a_sq, csq[m], dot[m]);
        if (argmin_better(dist, code, best_d, best_i)) { second_d = best_d; best_d = dist; best_i = code; }
        else if (dist < second_d) second_d = dist;
    }
}

// kServiceHigh: the TMA / MMA / TMEM warps take the highest warp ids (the issue arbiter favours high ids).
template <bool kServiceHigh>
__global__ void __launch_bounds__(kThreads, 1)
k_dist_tc16(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int T, int K,
            const float* __restrict__ zn32, const float* __restrict__ row_sq, const float* __restrict__ en32,
            const float* __restrict__ code_sq, const int* __restrict__ cb_info, int* __restrict__ cand,
            int* __restrict__ flagged, int* __restrict__ n_flagged, int64_t* __restrict__ stats, int debug_flags) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const SmemLayout L = smem_layout();
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + L.bars;
    auto b_full = [&](int s) { return bar_base + 8 * s; };
    auto b_empty = [&](int s) { return bar_base + 8 * (4 + s); };
    // accumulator stage q = 2 * (n-tile parity) + (row half): 128 rows x 128 codes, i.e. 128 TMEM columns
    auto t_full = [&](int q) { return bar_base + 8 * (8 + q); };
    auto t_empty = [&](int q) { return bar_base + 8 * (12 + q); };
    auto a_full = [&](int s) { return bar_base + 8 * (16 + s); };
    auto a_empty = [&](int s) { return bar_base + 8 * (18 + s); };
    auto h_full = [&](int s) { return bar_base + 8 * (20 + s); };
    auto h_empty = [&](int s) { return bar_base + 8 * (22 + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L.tmem_slot);

    // shuffled so the compiler knows the warp index is warp-uniform: role branches and everything the
    // service warps derive from it then live in the uniform datapath (no R2UR / elect loops around tcgen05 ops)
    const int warp = __shfl_sync(VQ_FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    constexpr int kEpiWarp0 = kServiceHigh ? 0 : 8;
    constexpr int kRescoreWarp0 = kServiceHigh ? 8 : 4;
    constexpr int kTmaWarp = kServiceHigh ? 12 : 0, kMmaWarp = kTmaWarp + 1, kAllocWarp = kTmaWarp + 2;   // second MMA warp: kTmaWarp + 3
    const int n_row_tiles = (T + kRowsPerCta - 1) / kRowsPerCta;
    const int n_tiles = K / kTileN;
    const int n_groups = n_tiles / kGroupTiles;

    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < kBStages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 2); }   // both MMA warps commit
        for (int q = 0; q < 4; ++q) { mbar_init(t_full(q), 1); mbar_init(t_empty(q), 4); }     // one arrive per epilogue warp of the half
        for (int s = 0; s < 2; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 2); }
        for (int s = 0; s < 2; ++s) { mbar_init(h_full(s), 256); mbar_init(h_empty(s), kRescoreThreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kAllocWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_base + L.tmem_slot),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // every role reads the TMEM base address itself, after its setmaxnreg: a value kept live across the role
    // split ends up in a local-memory spill slot that the epilogue would reload once per tile

    if (warp >= kTmaWarp && warp < kTmaWarp + 4) {
        reg_dec<kRegsService>();
        if (warp == kTmaWarp) {
            // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
            uint32_t b_cnt = 0;
            int it = 0;
            VQ_INSTR_BEGIN();
            for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x, ++it) {
                const int as = it & 1;
                VQ_TIMED_WAIT(0, a_empty(as), (((uint32_t)(it >> 1)) & 1u) ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(a_full(as), kABytes);
                    tma_load_2d(smem_base + L.a + as * kABytes, &tm_a, a_full(as), 0, rt * kRowsPerCta);
                }
                __syncwarp();
                for (int nb = 0; nb < n_tiles / 2; ++nb, ++b_cnt) {
                    const int s = b_cnt % kBStages;
                    VQ_TIMED_WAIT(1, b_empty(s), ((b_cnt / kBStages) & 1u) ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(b_full(s), kBStageBytes);
                        tma_load_2d(smem_base + L.b + s * kBStageBytes, &tm_b, b_full(s), 0, nb * kBStageCodes);
                    }
                    __syncwarp();
                }
            }
            VQ_INSTR_END(0, 2);
        } else if (warp == kMmaWarp || warp == kMmaWarp + 2) {
            // ===================== MMA issuers: one warp per row half (whole warp loops, one elected lane issues) ==========
            // At D = 32 a 128 x 128 x 32 unit is only 128 tensor cycles, so the issue path (barrier wait, descriptors,
            // 2 x tcgen05.mma, commit) of a single thread would be the bottleneck; the two halves are independent.
            const int r = (warp == kMmaWarp) ? 0 : 1;
            const uint32_t tmem_base = *tmem_slot;
            uint32_t b_cnt = 0, t_cnt = 0;
            int it = 0;
            VQ_INSTR_BEGIN();
            for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x, ++it) {
                const int as = it & 1;
                VQ_TIMED_WAIT(0, a_full(as), ((uint32_t)(it >> 1)) & 1u);
                tc_fence_after();
                const uint32_t a_addr = smem_base + L.a + as * kABytes + r * (128 * 64);
                for (int nb = 0; nb < n_tiles / 2; ++nb, ++b_cnt) {
                    const int s = b_cnt % kBStages;
                    VQ_TIMED_WAIT(2, b_full(s), (b_cnt / kBStages) & 1u);
                    tc_fence_after();
                    const uint32_t b_addr = smem_base + L.b + s * kBStageBytes;
#pragma unroll
                    for (int p = 0; p < 2; ++p, ++t_cnt) {      // n-tile p of this B stage -> accumulator stage 2p + r
                        VQ_TIMED_WAIT(1, t_empty(2 * p + r), ((t_cnt >> 1) & 1u) ^ 1u);
                        tc_fence_after();
                        VQ_TIMED_BEGIN();
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                umma_f16(tmem_base + (uint32_t)((p * 2 + r) * kTileN), umma_desc(a_addr + k * 32),
                                         umma_desc(b_addr + p * (kTileN * 64) + k * 32), kIdesc, (uint32_t)k);
                            umma_commit(t_full(2 * p + r));
                        }
                        __syncwarp();
                        VQ_TIMED_END(3);
                    }
                    if (elect_one()) umma_commit(b_empty(s));
                    __syncwarp();
                }
                if (elect_one()) umma_commit(a_empty(as));
                __syncwarp();
            }
            if (r == 0) { VQ_INSTR_END(3, 4); }
        }
    } else if (warp >= kRescoreWarp0 && warp < kRescoreWarp0 + 4) {
        // ===================== rescoring: 128 threads, 2 rows each per row tile =====================
        reg_dec<kRegsRescore>();
        const int rtid = threadIdx.x - kRescoreWarp0 * 32;
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
                const int gs1 = (h0.x >> 16) & 0x7FFF, gs2 = h0.y;
                const unsigned long long m1 = (unsigned long long)(uint32_t)h1.x | ((unsigned long long)(uint32_t)h1.y << 32);
                const unsigned long long m2 = (unsigned long long)(uint32_t)h1.z | ((unsigned long long)(uint32_t)h1.w << 32);
                float4 z[kD / 4];
                const float4* z4 = reinterpret_cast<const float4*>(zn32 + (int64_t)row * kD);
#pragma unroll
                for (int q = 0; q < kD / 4; ++q) z[q] = __ldg(z4 + q);
                const float a_sq = __ldg(row_sq + row);
                float best_d = INFINITY, second_d = INFINITY;
                int best_i = 0x7fffffff, n_cells = 0;
                // ONE loop over the row's cells, whichever area they come from: lanes of a warp that sit in
                // different areas / mask halves still share every iteration (a loop per area would serialise them)
                unsigned long long cur = (unsigned long long)(uint32_t)h0.z | ((unsigned long long)(uint32_t)h0.w << 32);
                int a = 0, g = h0.x & 0xFFFF;
                for (;;) {
                    if (cur == 0ull) {
                        if (a == 2) break;
                        ++a;
                        cur = (a == 1) ? m1 : m2;
                        g = (a == 1) ? gs1 : gs2;
                        continue;
                    }
                    const int hs = __ffsll((long long)cur) - 1;
                    cur &= cur - 1;
                    ++n_cells;
                    rescore_cell(g, hs, z, a_sq, en32, code_sq, best_d, best_i, second_d);
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
        const int e = warp - kEpiWarp0;
        const int r_sub = e >> 2;                    // which 128-row MMA tile
        const int quarter = warp & 3;                // TMEM lane quarter this warp may read
        const int row_in_cta = r_sub * 128 + quarter * 32 + lane;
        const uint32_t tbase = *tmem_slot + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(r_sub * kTileN);
        const uint32_t snap0 = smem_base + L.snap + (uint32_t)row_in_cta * kSnapRow;
        const bool force_exhaustive = (cb_info[0] != 0);
        uint32_t t_cnt = 0;                          // n-tiles drained so far (same sequence as the MMA warp)
        int it = 0;
        uint32_t bufA[64], bufB[64];
        VQ_INSTR_BEGIN();
        for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x, ++it) {
            uint32_t slot[32];
            int m1 = -32768, m2 = -32768, m3 = -32768, m4 = -32768;
            int g1 = 0, g2 = 0, g3 = 0;
            uint32_t a1 = 0, a2 = 1, a3 = 2;         // snapshot areas of the best / second / third group
            VQ_TIMED_WAIT(0, t_full(r_sub), (t_cnt >> 1) & 1u);
            tc_fence_after();
            tmem_ld_tile(tbase, bufA);
            for (int g = 0; g < n_groups; ++g) {
#pragma unroll
                for (int b = 0; b < kGroupTiles; ++b) {
                    uint32_t (&cur)[64] = (b & 1) ? bufB : bufA;
                    uint32_t (&nxt)[64] = (b & 1) ? bufA : bufB;
                    VQ_TIMED_BEGIN();
                    tmem_ld_wait();                                   // tile b is in registers: its TMEM stage is free
                    VQ_TIMED_END(2);
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(t_empty(2 * (b & 1) + r_sub));
                    ++t_cnt;
                    if (b < kGroupTiles - 1 || g + 1 < n_groups) {
                        VQ_TIMED_WAIT(0, t_full(2 * ((b + 1) & 1) + r_sub), (t_cnt >> 1) & 1u);
                        tc_fence_after();
                        tmem_ld_tile(tbase + (uint32_t)(((b + 1) & 1) * 2 * kTileN), nxt);
                    }
                    if (b == 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) slot[j] = __vmaxs2(cur[j], cur[j + 32]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) slot[j] = __vimax3_s16x2(slot[j], cur[j], cur[j + 32]);
                    }
                }
                // group maximum: 3-input tree over the 32 packed registers, then the two halves
                uint32_t t[11];
#pragma unroll
                for (int j = 0; j < 10; ++j) t[j] = __vimax3_s16x2(slot[3 * j], slot[3 * j + 1], slot[3 * j + 2]);
                t[10] = __vmaxs2(slot[30], slot[31]);
                const uint32_t u0 = __vimax3_s16x2(t[0], t[1], t[2]), u1 = __vimax3_s16x2(t[3], t[4], t[5]),
                               u2 = __vimax3_s16x2(t[6], t[7], t[8]);
                const uint32_t pk = __vimax3_s16x2(__vimax3_s16x2(u0, u1, u2), t[9], t[10]);
                const int c1 = max(lo16(pk), hi16(pk));
                const bool is1 = c1 > m1, is2 = c1 > m2, is3 = c1 > m3;
                if (is3) {
                    // whichever rank the group takes, the group that drops out is the current last one: reuse its area
                    const uint32_t dst = snap0 + a3 * kSnapArea;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16 * q), "r"(slot[4 * q]),
                                     "r"(slot[4 * q + 1]), "r"(slot[4 * q + 2]), "r"(slot[4 * q + 3])
                                     : "memory");
                }
                // sorted insert of c1 into (m1 >= m2 >= m3 >= m4); identities and areas follow
                const int lo1 = min(c1, m1);
                m1 = max(c1, m1);
                const int lo2 = min(lo1, m2);
                m2 = max(lo1, m2);
                const int lo3 = min(lo2, m3);
                m3 = max(lo2, m3);
                m4 = max(lo3, m4);
                const int ng3 = is2 ? g2 : (is3 ? g : g3);
                const uint32_t na3 = is2 ? a2 : a3;
                const int ng2 = is1 ? g1 : (is2 ? g : g2);
                const uint32_t na2 = is1 ? a1 : (is2 ? a3 : a2);
                const uint32_t na1 = is1 ? a3 : a1;
                g1 = is1 ? g : g1; g2 = ng2; g3 = ng3;
                a1 = na1; a2 = na2; a3 = na3;
            }
            // ---- row verdict ----
            // m1 is a non-negative finite fp16 pattern for every row the filter may decide; NaN / Inf patterns,
            // negative best scores and thresholds near zero leave the row to the exhaustive search
            const float m1f = __half2float(__ushort_as_half((unsigned short)(m1 & 0xFFFF)));
            const float thr_f = m1f - kTwoEps;
            const bool thr_ok = (m1 >= 0) && (m1 < 0x7C00) && (thr_f >= kMinThreshold);
            const int thr = thr_ok ? (int)__half_as_ushort(__float2half_rd(thr_f)) : 0x7BFF;
            const uint32_t thr2 = (uint32_t)thr * 0x10001u;
            uint32_t mask[6] = {0u, 0u, 0u, 0u, 0u, 0u};
            const int mv[3] = {m1, m2, m3};
            const uint32_t av[3] = {a1, a2, a3};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (a == 0 || mv[a] >= thr) {
                    const uint32_t src = snap0 + av[a] * kSnapArea;
                    uint32_t kept[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(kept[4 * q]), "=r"(kept[4 * q + 1]), "=r"(kept[4 * q + 2]), "=r"(kept[4 * q + 3])
                                     : "r"(src + 16 * q));
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        bool ph, pl;
                        (void)__vibmax_s16x2(kept[j], thr2, &ph, &pl);      // per half: kept >= thr
                        const uint32_t bits = (pl ? (1u << ((2 * j) & 31)) : 0u) | (ph ? (1u << ((2 * j + 1) & 31)) : 0u);
                        mask[2 * a + (j >> 4)] |= bits;
                    }
                }
            }
            // decided iff no further group can hold the winner
            const bool decided = thr_ok && (m4 < thr) && ((mask[0] | mask[1]) != 0) && !force_exhaustive;
            const int row = rt * kRowsPerCta + row_in_cta;
            const bool in_range = row < T;
            const bool flag = in_range && !decided;
            // hand the verdict to the rescoring warps (double-buffered)
            const int hb = it & 1;
            VQ_TIMED_WAIT(1, h_empty(hb), (((uint32_t)(it >> 1)) & 1u) ^ 1u);
            int4* hand = reinterpret_cast<int4*>(smem + L.hand + hb * kHandBytes);
            hand[2 * row_in_cta] = make_int4(decided ? (g1 | (g2 << 16)) : -1, g3, (int)mask[0], (int)mask[1]);
            hand[2 * row_in_cta + 1] = make_int4((int)mask[2], (int)mask[3], (int)mask[4], (int)mask[5]);
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
        if (threadIdx.x == kEpiWarp0 * 32) { VQ_INSTR_END(8, 3); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kAllocWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(512u) : "memory");
    }
}

}  // namespace tc16

bool tc16_supported(int64_t T, int K, int D) {
    static const bool disabled = getenv("VQ_TC16_DISABLE") && atoi(getenv("VQ_TC16_DISABLE")) != 0;
    // group ids travel as 15/16-bit fields; the threshold logic needs whole 512-code groups
    return !disabled && D == tc16::kD && K >= tc16::kGroupCols && (K % tc16::kGroupCols) == 0 &&
           K / tc16::kGroupCols <= 32767 && T >= 256;
}

cudaError_t launch_dist_tc16(const CUtensorMap& ma, const CUtensorMap& mb, int T, const float* zn32, const float* row_sq,
                             const CodebookView& cb, int* cand, int* flagged, int* n_flagged, int64_t* stats,
                             cudaStream_t s) {
    const tc16::SmemLayout L = tc16::smem_layout();
    static const bool service_low = getenv("VQ_TC16_SERVICE_LOW") && atoi(getenv("VQ_TC16_SERVICE_LOW")) != 0;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc16::k_dist_tc16<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total + 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tc16::k_dist_tc16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total + 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int n_row_tiles = (T + tc16::kRowsPerCta - 1) / tc16::kRowsPerCta;
    const int grid = n_row_tiles < sm_count() ? n_row_tiles : sm_count();
    static const int debug_flags = getenv("VQ_TC_DEBUG") ? atoi(getenv("VQ_TC_DEBUG")) : 0;
    if (service_low)
        tc16::k_dist_tc16<false><<<grid, tc16::kThreads, L.total + 1024, s>>>(ma, mb, T, cb.K, zn32, row_sq, cb.en32, cb.code_sq,
                                                                             cb.info, cand, flagged, n_flagged, stats, debug_flags);
    else
        tc16::k_dist_tc16<true><<<grid, tc16::kThreads, L.total + 1024, s>>>(ma, mb, T, cb.K, zn32, row_sq, cb.en32, cb.code_sq,
                                                                            cb.info, cand, flagged, n_flagged, stats, debug_flags);
    count_launch();
    tc::instrument_report(s, grid);
    return cudaGetLastError();
}

}  // namespace vq
