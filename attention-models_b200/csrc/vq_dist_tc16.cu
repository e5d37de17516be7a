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
// Two kernels.  k_dist_tc16 is the filter: it writes one 32-byte verdict record per row (the best three groups
// and, per group, a 64-bit mask of the cells that can still win).  k_rescore16 turns records into indices with
// the exact fp32 formula.  Rescoring inside the tensor-core kernel was tried first and cost more than it hid: four
// rescoring warps per SM are instruction-latency bound (~300 dependent instructions per item at IPC 0.2) and
// their scattered 16-byte loads saturate the L1 wavefront pipe under the epilogue; as its own kernel the same
// work runs at full occupancy.  Rescoring is warp-cooperative: 8 lanes share one row, each load instruction
// fetches whole 128-byte code rows (the cell's 8 rows are transposed through shared memory), so the L1 sees one
// wavefront per code row instead of eight.
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
constexpr int kBStageBytes = kTileN * kD * 2;       // 8 KiB: one n-tile of fp16 unit codes per TMA
constexpr int kBStages = 6;
constexpr int kAStages = 2;
constexpr int kThreads = 384;          // warps 0-7 epilogue, 8 TMA, 9 + 11 MMA (one per row half), 10 TMEM allocator
constexpr int kRegsService = 40, kRegsEpilogue = 232;
static_assert(128 * kRegsService + 256 * kRegsEpilogue <= 65536, "register file overcommitted");
// |fp16-pipeline score - exact dot| <= eps: 1.1e-3 for the fp16 operands (as in vq_dist_tc.cu) + one fp16 ulp
// per accumulator rounding (2^-11 below 1, 2^-10 in [1, 2)); the threshold is additionally rounded down to fp16
constexpr float kTwoEps = 2.f * (1.1e-3f + 4.8829e-4f + 4.8829e-4f);
constexpr float kTwoEpsNearOne = 2.f * (1.1e-3f + 9.7657e-4f + 9.7657e-4f);
constexpr float kMinThreshold = 1e-3f;              // the filter only trusts rows whose threshold is safely positive
// c_format = F16 (0), a/b = F16 (0), K-major both, N = 128, M = 128
constexpr uint32_t kIdesc = ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

constexpr int kSnapRow = 144;                       // 32 packed slot registers + 16 B pad (conflict-free STS.128)
constexpr int kSnapArea = kRowsPerCta * kSnapRow;
// verdict record, 48 bytes per row: {g1 | g2 << 16 or -1 (undecided), g3 | g4 << 16, mask0 lo, hi}
// {mask1 lo, hi, mask2 lo, hi} {mask3 lo, hi, -, -}: the best four groups and, per group, the cells that can still win
constexpr int kRecordBytes = 48;
constexpr int kAreas = 4;

struct SmemLayout {
    uint32_t a, b, snap, bars, tmem_slot, total;
};
__host__ __device__ inline SmemLayout smem_layout() {
    SmemLayout L;
    L.a = 0;
    L.b = L.a + kAStages * kABytes;
    L.snap = L.b + kBStages * kBStageBytes;
    L.bars = L.snap + kAreas * kSnapArea;
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

// running (best distance, its index, second-best distance) of one row, torch.argmin ordering
struct Best3 {
    float d1; int i1; float d2;
};
// fold another lane's triple in (codes are disjoint between lanes): same rules as vq_dist_simt.cu's merge
__device__ __forceinline__ void best3_merge(Best3& a, float d1, int i1, float d2) {
    const float nan = __int_as_float(0x7fc00000);
    if (argmin_better(d1, i1, a.d1, a.i1)) {
        const float loser = a.d1;
        a.d1 = d1; a.i1 = i1;
        a.d2 = (loser != loser || d2 != d2) ? nan : fminf(loser, d2);
    } else {
        a.d2 = (d1 != d1 || a.d2 != a.d2) ? nan : fminf(a.d2, d1);
    }
}

// kServiceHigh: the TMA / MMA / TMEM warps take the highest warp ids (the issue arbiter favours high ids).
template <bool kServiceHigh>
__global__ void __launch_bounds__(kThreads, 1)
k_dist_tc16(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int T, int K,
            const int* __restrict__ cb_info, int4* __restrict__ rec, int* __restrict__ cand,
            int* __restrict__ flagged, int* __restrict__ n_flagged, int64_t* __restrict__ stats) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const SmemLayout L = smem_layout();
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + L.bars;
    auto b_full = [&](int s) { return bar_base + 8 * s; };
    auto b_empty = [&](int s) { return bar_base + 8 * (8 + s); };
    // accumulator stage q = 2 * (n-tile parity) + (row half): 128 rows x 128 codes, i.e. 128 TMEM columns
    auto t_full = [&](int q) { return bar_base + 8 * (16 + q); };
    auto t_empty = [&](int q) { return bar_base + 8 * (20 + q); };
    auto a_full = [&](int s) { return bar_base + 8 * (24 + s); };
    auto a_empty = [&](int s) { return bar_base + 8 * (26 + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L.tmem_slot);

    // shuffled so the compiler knows the warp index is warp-uniform: role branches and everything the
    // service warps derive from it then live in the uniform datapath (no R2UR / elect loops around tcgen05 ops)
    const int warp = __shfl_sync(VQ_FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    constexpr int kEpiWarp0 = kServiceHigh ? 0 : 4;
    constexpr int kTmaWarp = kServiceHigh ? 8 : 0, kMmaWarp = kTmaWarp + 1, kAllocWarp = kTmaWarp + 2;   // second MMA warp: kTmaWarp + 3
    const int n_row_tiles = (T + kRowsPerCta - 1) / kRowsPerCta;
    const int n_tiles = K / kTileN;
    const int n_groups = n_tiles / kGroupTiles;

    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < kBStages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 2); }   // both MMA warps commit
        for (int q = 0; q < 4; ++q) { mbar_init(t_full(q), 1); mbar_init(t_empty(q), 4); }     // one arrive per epilogue warp of the half
        for (int s = 0; s < 2; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 2); }
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
    // barrier init and the TMEM allocation above overlap the tail of the token prep kernel (programmatic dependent
    // launch); nothing before this line reads what that kernel writes
    pdl_trigger();
    pdl_wait();
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
                for (int n = 0; n < n_tiles; ++n, ++b_cnt) {
                    const int s = b_cnt % kBStages;
                    VQ_TIMED_WAIT(1, b_empty(s), ((b_cnt / kBStages) & 1u) ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(b_full(s), kBStageBytes);
                        tma_load_2d(smem_base + L.b + s * kBStageBytes, &tm_b, b_full(s), 0, n * kTileN);
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
                for (int n = 0; n < n_tiles; n += 2) {
#pragma unroll
                    for (int p = 0; p < 2; ++p, ++t_cnt, ++b_cnt) {      // n-tile n + p -> accumulator stage 2p + r
                        const int s = b_cnt % kBStages;
                        VQ_TIMED_WAIT(2, b_full(s), (b_cnt / kBStages) & 1u);
                        VQ_TIMED_WAIT(1, t_empty(2 * p + r), ((t_cnt >> 1) & 1u) ^ 1u);
                        tc_fence_after();
                        if (r == 0 && lane == 0) VQ_TRACE(0, (int)t_cnt, 0);
                        const uint32_t b_addr = smem_base + L.b + s * kBStageBytes;
                        VQ_TIMED_BEGIN();
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                umma_f16(tmem_base + (uint32_t)((p * 2 + r) * kTileN), umma_desc(a_addr + k * 32),
                                         umma_desc(b_addr + k * 32), kIdesc, (uint32_t)k);
                            umma_commit(t_full(2 * p + r));
                            umma_commit(b_empty(s));
                        }
                        __syncwarp();
                        if (r == 0 && lane == 0) VQ_TRACE(0, (int)t_cnt, 1);
                        VQ_TIMED_END(3);
                    }
                }
                if (elect_one()) umma_commit(a_empty(as));
                __syncwarp();
            }
            if (r == 0) { VQ_INSTR_END(3, 4); }
        }
    } else {
        // ===================== epilogue: 8 warps, one thread per row =====================
        reg_inc<kRegsEpilogue>();
        const int e = warp - kEpiWarp0;
        const int r_sub = e >> 2;                    // which 128-row MMA tile
        const int quarter = warp & 3;                // TMEM lane quarter this warp may read (warp id mod 4)
        const int row_in_cta = r_sub * 128 + quarter * 32 + lane;
        const uint32_t tbase = *tmem_slot + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(r_sub * kTileN);
        const uint32_t snap0 = smem_base + L.snap + (uint32_t)row_in_cta * kSnapRow;
        const bool force_exhaustive = codebook_degenerate(cb_info);
        uint32_t t_cnt = 0;                          // n-tiles drained so far (same sequence as the MMA warp)
        int it = 0;
        uint32_t bufA[64], bufB[64];
        VQ_INSTR_BEGIN();
        for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x, ++it) {
            uint32_t slot[32];
            int m1 = -32768, m2 = -32768, m3 = -32768, m4 = -32768, m5 = -32768;
            int g1 = 0, g2 = 0, g3 = 0, g4 = 0;
            uint32_t a1 = 0, a2 = 1, a3 = 2, a4 = 3;  // snapshot areas of the best four groups
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
#ifdef VQ_TC_INSTRUMENT
                    if (e == 0 && lane == 0) { asm volatile("" ::"r"(cur[0]), "r"(cur[63])); VQ_TRACE(1, (int)t_cnt, 1); }
#endif
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(t_empty(2 * (b & 1) + r_sub));
                    if (e == 0 && lane == 0) VQ_TRACE(1, (int)t_cnt, 2);
                    ++t_cnt;
                    if (b < kGroupTiles - 1 || g + 1 < n_groups) {
                        VQ_TIMED_WAIT(0, t_full(2 * ((b + 1) & 1) + r_sub), (t_cnt >> 1) & 1u);
                        if (e == 0 && lane == 0) VQ_TRACE(1, (int)t_cnt, 0);
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
#ifdef VQ_TC_INSTRUMENT
                    if (e == 0 && lane == 0) { asm volatile("" ::"r"(slot[0]), "r"(slot[31])); VQ_TRACE(1, (int)t_cnt - 1, 3); }
#endif
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
                const bool is1 = c1 > m1, is2 = c1 > m2, is3 = c1 > m3, is4 = c1 > m4;
                if (is4) {
                    // whichever rank the group takes, the group that drops out is the current last one: reuse its area
                    const uint32_t dst = snap0 + a4 * kSnapArea;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16 * q), "r"(slot[4 * q]),
                                     "r"(slot[4 * q + 1]), "r"(slot[4 * q + 2]), "r"(slot[4 * q + 3])
                                     : "memory");
                }
                // sorted insert of c1 into (m1 >= m2 >= m3 >= m4 >= m5); identities and areas follow
                const int lo1 = min(c1, m1);
                m1 = max(c1, m1);
                const int lo2 = min(lo1, m2);
                m2 = max(lo1, m2);
                const int lo3 = min(lo2, m3);
                m3 = max(lo2, m3);
                const int lo4 = min(lo3, m4);
                m4 = max(lo3, m4);
                m5 = max(lo4, m5);
                const int ng4 = is3 ? g3 : (is4 ? g : g4);
                const uint32_t na4 = is3 ? a3 : a4;
                const int ng3 = is2 ? g2 : (is3 ? g : g3);
                const uint32_t na3 = is2 ? a2 : (is3 ? a4 : a3);
                const int ng2 = is1 ? g1 : (is2 ? g : g2);
                const uint32_t na2 = is1 ? a1 : (is2 ? a4 : a2);
                const uint32_t na1 = is1 ? a4 : a1;
                g1 = is1 ? g : g1; g2 = ng2; g3 = ng3; g4 = ng4;
                a1 = na1; a2 = na2; a3 = na3; a4 = na4;
            }
            // ---- row verdict ----
            // m1 is a non-negative finite fp16 pattern for every row the filter may decide; NaN / Inf patterns,
            // negative best scores and thresholds near zero leave the row to the exhaustive search
            const float m1f = __half2float(__ushort_as_half((unsigned short)(m1 & 0xFFFF)));
            // a code whose K = 16 partial sum reaches 1 has a final score >= 0.95, so below 0.9 both accumulator
            // roundings are at most one ulp of [0.5, 1)
            const float thr_f = m1f - (m1f < 0.9f ? kTwoEps : kTwoEpsNearOne);
            const bool thr_ok = (m1 >= 0) && (m1 < 0x7C00) && (thr_f >= kMinThreshold);
            const int thr = thr_ok ? (int)__half_as_ushort(__float2half_rd(thr_f)) : 0x7BFF;
            const uint32_t thr2 = (uint32_t)thr * 0x10001u;
            uint32_t mask[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            const int mv[4] = {m1, m2, m3, m4};
            const uint32_t av[4] = {a1, a2, a3, a4};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
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
            const int n_cand = __popc(mask[0]) + __popc(mask[1]) + __popc(mask[2]) + __popc(mask[3]) + __popc(mask[4]) + __popc(mask[5]) +
                               __popc(mask[6]) + __popc(mask[7]);
            const bool decided = thr_ok && (m5 < thr) && ((mask[0] | mask[1]) != 0) && (n_cand <= 16) && !force_exhaustive;
            const int row = rt * kRowsPerCta + row_in_cta;
            const bool in_range = row < T;
            const bool flag = in_range && !decided;
            if (in_range) {
                rec[3 * (int64_t)row] = make_int4(decided ? (g1 | (g2 << 16)) : -1, g3 | (g4 << 16), (int)mask[0], (int)mask[1]);
                rec[3 * (int64_t)row + 1] = make_int4((int)mask[2], (int)mask[3], (int)mask[4], (int)mask[5]);
                rec[3 * (int64_t)row + 2] = make_int4((int)mask[6], (int)mask[7], 0, 0);
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
        if (threadIdx.x == kEpiWarp0 * 32) { VQ_INSTR_END(8, 3); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kAllocWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(512u) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Exact fp32 kernels behind the filter.  Both read the cell copies of the codebook (CodebookView::en32c): the 8
// lanes of a group each own one code of a cell and read whole 128-byte lines together, and every lane runs the
// same sequential fma chain over d = 0..31 as the exhaustive search of vq_dist_simt.cu, so all paths return
// identical indices.  (distance, index) pairs are compared as one 64-bit key: sign-corrected float bits, ties to
// the lower index, NaN below everything (torch.argmin: NaN wins).
// ---------------------------------------------------------------------------------------------------------------
using tc::dist_key;
using tc::key_dist;
using tc::Top2;

// distances of one row (z in registers) to the 8 codes of cell ci: this lane's code is m
__device__ __forceinline__ float cell_distance(const float4* __restrict__ en32c, const float* __restrict__ csq_cell, int ci,
                                               int m, const float4 (&z)[kD / 4], float a_sq) {
    const float4* e4 = en32c + (int64_t)ci * 64 + m;
    float4 ev[kD / 4];
#pragma unroll
    for (int q = 0; q < kD / 4; ++q) ev[q] = __ldg(e4 + 8 * q);
    const float csq = __ldg(csq_cell + ci * 8 + m);
    float dot = 0.f;
#pragma unroll
    for (int q = 0; q < kD / 4; ++q) {
        dot = __fmaf_rn(z[q].x, ev[q].x, dot);
        dot = __fmaf_rn(z[q].y, ev[q].y, dot);
        dot = __fmaf_rn(z[q].z, ev[q].z, dot);
        dot = __fmaf_rn(z[q].w, ev[q].w, dot);
    }
    return ref_distance(a_sq, csq, dot);
}

// the same with the row staged in shared memory (zs: the row's 8 float4; the 8 lanes of a group read one address)
__device__ __forceinline__ float cell_distance_staged(const float4* __restrict__ en32c, const float* __restrict__ csq_cell,
                                                      int ci, int m, const float4* zs, float a_sq) {
    const float4* e4 = en32c + (int64_t)ci * 64 + m;
    float4 ev[kD / 4];
#pragma unroll
    for (int q = 0; q < kD / 4; ++q) ev[q] = __ldg(e4 + 8 * q);
    const float csq = __ldg(csq_cell + ci * 8 + m);
    float dot = 0.f;
#pragma unroll
    for (int q = 0; q < kD / 4; ++q) {
        const float4 z = zs[q];
        dot = __fmaf_rn(z.x, ev[q].x, dot);
        dot = __fmaf_rn(z.y, ev[q].y, dot);
        dot = __fmaf_rn(z.z, ev[q].z, dot);
        dot = __fmaf_rn(z.w, ev[q].w, dot);
    }
    return ref_distance(a_sq, csq, dot);
}

// What vq_finish.cu does for one row: z_q = zn + (q - zn), loss partial in fixed point and -- when the step trains
// the codebook -- the row's terms of the segment sums S_code += fixed(q - zn).
struct FinishOut {
    float* zq; int64_t* idx; int32_t* hist; unsigned long long* seg;
};
// chunk `c` (4 elements) of one row: returns q - zn of the chunk
__device__ __forceinline__ float4 finish_chunk(const float4 a, const float4* __restrict__ en4, const FinishOut& out, int row,
                                               int code, int c, long long& loss_fx, unsigned& bad) {
    const float4 q = __ldg(en4 + (int64_t)code * (kD / 4) + c);
    float4 df, o;
    df.x = __fsub_rn(q.x, a.x); df.y = __fsub_rn(q.y, a.y); df.z = __fsub_rn(q.z, a.z); df.w = __fsub_rn(q.w, a.w);
    o.x = __fadd_rn(a.x, df.x); o.y = __fadd_rn(a.y, df.y); o.z = __fadd_rn(a.z, df.z); o.w = __fadd_rn(a.w, df.w);
    __stcs(reinterpret_cast<float4*>(out.zq) + (int64_t)row * (kD / 4) + c, o);
    const float p = (df.x * df.x + df.y * df.y) + (df.z * df.z + df.w * df.w);
    if (is_finite(p)) loss_fx += to_fixed(p, VQ_LOSS_SHIFT);
    else bad += 1;
    return df;
}
// a whole row by one thread (the few rows of phase B)
__device__ __forceinline__ void finish_row_serial(const float4* __restrict__ zn4, const float4* __restrict__ en4,
                                                  const FinishOut& out, int K, int row, int code, long long& loss_fx,
                                                  unsigned& bad) {
    unsigned poison = 0;
    for (int c = 0; c < kD / 4; ++c) {
        const float4 df = finish_chunk(__ldg(zn4 + (int64_t)row * (kD / 4) + c), en4, out, row, code, c, loss_fx, bad);
        if (out.seg) {
            unsigned long long* slot = out.seg + (int64_t)code * kD + 4 * c;
            seg_add(slot + 0, df.x, poison); seg_add(slot + 1, df.y, poison);
            seg_add(slot + 2, df.z, poison); seg_add(slot + 3, df.w, poison);
        }
    }
    if (poison) atomicAdd(out.seg + (int64_t)K * kD + code, 1ull);
}

// Everything behind the filter in one launch:
//   phase A  rescoring of the verdict records, one 8-lane group per row: index = argmin over the row's surviving
//            cells; the same lanes then write idx / hist / z_q / loss partial (the former k_finish pass);
//   phase B  exhaustive search of the rows the filter could not decide (a few dozen per 262 144).  Latency
//            matters there, not throughput: a listed row is split over kFlaggedSlices items, an item leaves its
//            (best, second) in `partial`, the last item of a row to finish folds them and finishes the row.
//            Rows [0, min(*n_rows, cap)) of the list; `done` holds one zeroed counter per listed row.
constexpr int kExactThreads = 128;
#ifndef VQ_EXACT_MIN_BLOCKS
#define VQ_EXACT_MIN_BLOCKS 8   // 64 registers (24 B of spills): 2% faster than 6 blocks of 80 registers
#endif
constexpr int kFlaggedSlices = 32;
struct __align__(16) FlaggedPartial {
    unsigned long long best; float second; float pad;
};
__global__ void __launch_bounds__(kExactThreads, VQ_EXACT_MIN_BLOCKS)
k_exact_finish16(const int4* __restrict__ rec, const float* __restrict__ zn32, const float* __restrict__ row_sq,
                 const float* __restrict__ en32, const float4* __restrict__ en32c, const float* __restrict__ csq_cell, int T,
                 int K, const int* __restrict__ flagged, const int* __restrict__ n_flagged, int flagged_cap,
                 FlaggedPartial* __restrict__ partial, int* __restrict__ done, int* __restrict__ cand, FinishOut out,
                 int64_t* __restrict__ stats) {
    __shared__ unsigned long long s_best[kExactThreads / 32];
    __shared__ float s_second[kExactThreads / 32];
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = lane & 7;
    const float4* zn4 = reinterpret_cast<const float4*>(zn32);
    // per-thread counters share one register (a thread sees at most a few dozen rows): ties in bits 0-9, rows
    // rescored over several cells in bits 10-19, non-finite loss partials above
    unsigned counts = 0, bad = 0;
    constexpr unsigned kTie = 1u, kMulti = 1u << 10;
    long long loss_fx = 0;
    // ---------------- phase A ----------------
    // A warp owns 4 rows per iteration.  Their unit rows (512 contiguous bytes) arrive with ONE coalesced 16-byte load
    // per lane and are staged in shared memory (the 8 lanes of a row then read them as broadcasts: 4 registers instead
    // of 32 for the row), in the same round trip as the row's record.
    // Finish: lane m owns chunk m of the row (one 16-byte access per lane for zn, the code row and z_q); for the
    // segment sums the differences are transposed through shared memory so that lane m adds elements m, m + 8, m + 16,
    // m + 24: every RED instruction of a group then covers 8 consecutive int64 (two whole 32-byte sectors; 4x fewer
    // L2 reduction transactions than 4 consecutive elements per lane, tools/ubench_red.cu).
    // Row strides of 9 / 10 float4 put the 4 rows of a warp in different banks: the broadcast LDS.128 of the rescoring
    // (one address per 8-lane group) and the strided LDS.32 of the segment-sum transposition are conflict-free.
    __shared__ __align__(16) float4 s_z[kExactThreads / 32][4][kD / 4 + 1];
    __shared__ __align__(16) float4 s_df[kExactThreads / 32][4][kD / 4 + 2];
    const float4* en4 = reinterpret_cast<const float4*>(en32);
    const int groups = gridDim.x * (kExactThreads / 8);
    const int grp = lane >> 3;
    int row0 = (blockIdx.x * kExactThreads + threadIdx.x - lane) >> 3;
    int4 n0 = make_int4(-1, 0, 0, 0), n1 = make_int4(0, 0, 0, 0), n2 = make_int4(0, 0, 0, 0);
    float4 nz = make_float4(0.f, 0.f, 0.f, 0.f);
    float n_sq = 0.f;
    auto fetch = [&](int r0) {
        const int r = r0 + grp;
        n0 = make_int4(-1, 0, 0, 0);
        if (r < T) {
            n0 = __ldg(rec + 3 * (int64_t)r); n1 = __ldg(rec + 3 * (int64_t)r + 1); n2 = __ldg(rec + 3 * (int64_t)r + 2);
            nz = __ldg(zn4 + (int64_t)r0 * (kD / 4) + lane);      // row r0 + (lane >> 3), chunk lane & 7
            n_sq = __ldg(row_sq + r);
        }
    };
#ifndef VQ_EXACT_PREFETCH
#define VQ_EXACT_PREFETCH 0     // 1: records / rows of the next iteration prefetched into registers -- measured 3 us slower
#endif                          // (17 more live registers spill under the 64-register cap; the data sits in L2 anyway)
    if (VQ_EXACT_PREFETCH && row0 < T) fetch(row0);
    for (; row0 < T; row0 += groups) {
        const int row = row0 + grp;
        if (!VQ_EXACT_PREFETCH) fetch(row0);
        const int4 h0 = n0, h1 = n1, h2 = n2;
        const float a_sq = n_sq;
        __syncwarp();                                   // the previous iteration's reads of s_z are done
        s_z[warp][grp][m] = nz;
        __syncwarp();
        if (VQ_EXACT_PREFETCH && row0 + groups < T) fetch(row0 + groups);
        const float4* zs = s_z[warp][grp];
        const bool valid = (row < T) && (h0.x >= 0);
        auto u64 = [](int lo, int hi) { return (unsigned long long)(uint32_t)lo | ((unsigned long long)(uint32_t)hi << 32); };
        unsigned long long cur = valid ? u64(h0.z, h0.w) : 0ull;
        const unsigned long long m1 = valid ? u64(h1.x, h1.y) : 0ull, m2 = valid ? u64(h1.z, h1.w) : 0ull,
                                 m3 = valid ? u64(h2.x, h2.y) : 0ull;
        const int gs1 = (h0.x >> 16) & 0x7FFF, gs2 = h0.y & 0xFFFF, gs3 = (h0.y >> 16) & 0x7FFF;
        int a = 0, g = h0.x & 0xFFFF;
        const int n_cells = __popcll(cur) + __popcll(m1) + __popcll(m2) + __popcll(m3);
        const int n_iter = __reduce_max_sync(VQ_FULL, n_cells);
        Top2 top;
        top.init();
#pragma unroll 1
        for (int c = 0; c < n_iter; ++c) {
#pragma unroll
            for (int skip = 0; skip < 3; ++skip)
                if (cur == 0ull && a < 3) { ++a; cur = (a == 1) ? m1 : (a == 2 ? m2 : m3); g = (a == 1) ? gs1 : (a == 2 ? gs2 : gs3); }
            if (cur != 0ull) {
                const int hs = __ffsll((long long)cur) - 1;
                cur &= cur - 1;
                const float dist = cell_distance_staged(en32c, csq_cell, g * 64 + hs, m, zs, a_sq);
                top.add(dist_key(dist, g * kGroupCols + hs + 64 * m));
            }
        }
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) {
            const unsigned long long ob = __shfl_xor_sync(VQ_FULL, top.best, off);
            const float os = __shfl_xor_sync(VQ_FULL, top.second, off);
            top.merge(ob, os);
        }
        const int code = (int)(uint32_t)top.best;
        // rows of the warp that chose the same code (collapsed codebooks): their histogram count and segment-sum terms
        // are added once, by the first of them -- the L2 serialises reductions per address
        const unsigned same = __match_any_sync(VQ_FULL, valid ? code : -1 - grp);
        const bool dup = __any_sync(VQ_FULL, __popc(same) > 8);        // uniform; false on all but collapsed usage
        const bool lead = valid && (!dup || ((__ffs(same) - 1) >> 3) == grp);
        if (valid) {
            if (m == 0) {
                const float bd = key_dist(top.best);
                if (cand) cand[row] = code | kCandExactBit;
                out.idx[row] = code;
                if (out.hist && lead) atomicAdd(out.hist + code, dup ? __popc(same & 0x01010101u) : 1);
                if (top.second - bd < VQ_NEAR_TIE_REL * fabsf(bd)) counts += kTie;
                if (n_cells > 1) counts += kMulti;
            }
        }
        if (out.zq) {                                   // uniform
            float4 df = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) df = finish_chunk(zs[m], en4, out, row, code, m, loss_fx, bad);
            if (out.seg) {                              // uniform
                s_df[warp][grp][m] = df;
                __syncwarp();
                if (!dup) {
                    if (valid) {
                        const float* d = reinterpret_cast<const float*>(s_df[warp][grp]);
                        unsigned poison = 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) seg_add(out.seg + (int64_t)code * kD + m + 8 * i, d[m + 8 * i], poison);
                        if (poison) atomicAdd(out.seg + (int64_t)K * kD + code, 1ull);
                    }
                } else if (lead) {
                    long long acc[4] = {0, 0, 0, 0};
                    unsigned poison = 0;
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                        if ((same >> (8 * p)) & 1u) {       // group p chose this code too (its lane 0 is in `same`)
                            const float* d = reinterpret_cast<const float*>(s_df[warp][p]);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float v = d[m + 8 * i];
                                if (is_finite(v)) acc[i] += to_fixed(v, VQ_SEG_SHIFT);   // per term, as every other path
                                else poison = 1;
                            }
                        }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        atomicAdd(out.seg + (int64_t)code * kD + m + 8 * i, (unsigned long long)acc[i]);
                    if (poison) atomicAdd(out.seg + (int64_t)K * kD + code, 1ull);
                }
            }
        }
    }
    // ---------------- phase B ----------------
    {
        const int grp = threadIdx.x >> 3;
        int n = *n_flagged;
        if (n > flagged_cap) n = flagged_cap;
        const int n_cells = K / kCellCodes;
        const int per_slice = (n_cells + kFlaggedSlices - 1) / kFlaggedSlices;
        for (int item = blockIdx.x; item < n * kFlaggedSlices; item += gridDim.x) {
            const int i = item / kFlaggedSlices, slice = item % kFlaggedSlices;
            const int row = flagged[i];
            __syncthreads();                       // the previous item's shared values are consumed
            if (threadIdx.x < kD / 4) s_z[0][0][threadIdx.x] = __ldg(zn4 + (int64_t)row * (kD / 4) + threadIdx.x);
            __syncthreads();
            const float4* zs = s_z[0][0];
            const float a_sq = __ldg(row_sq + row);
            Top2 top;
            top.init();
            const int c_end = min(n_cells, (slice + 1) * per_slice);
            for (int ci = slice * per_slice + grp; ci < c_end; ci += kExactThreads / 8) {
                const float dist = cell_distance_staged(en32c, csq_cell, ci, m, zs, a_sq);
                top.add(dist_key(dist, (ci >> 6) * kGroupCols + (ci & 63) + 64 * m));
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long ob = __shfl_xor_sync(VQ_FULL, top.best, off);
                const float os = __shfl_xor_sync(VQ_FULL, top.second, off);
                top.merge(ob, os);
            }
            if (lane == 0) { s_best[warp] = top.best; s_second[warp] = top.second; }
            __syncthreads();
            if (threadIdx.x == 0) {
                Top2 all;
                all.init();
                for (int w = 0; w < kExactThreads / 32; ++w) all.merge(s_best[w], s_second[w]);
                FlaggedPartial pp;
                pp.best = all.best; pp.second = all.second; pp.pad = 0.f;
                partial[(int64_t)i * kFlaggedSlices + slice] = pp;
                __threadfence();
                if (atomicAdd(done + i, 1) == kFlaggedSlices - 1) {
                    __threadfence();
                    Top2 fin;
                    fin.init();
                    for (int w = 0; w < kFlaggedSlices; ++w) {
                        const FlaggedPartial* q = partial + (int64_t)i * kFlaggedSlices + w;
                        fin.merge(__ldcg(&q->best), __ldcg(&q->second));
                    }
                    const float bd = key_dist(fin.best);
                    const int code = (int)(uint32_t)fin.best;
                    if (cand) cand[row] = code | kCandExactBit;
                    out.idx[row] = code;
                    if (out.hist) atomicAdd(out.hist + code, 1);
                    if (fin.second - bd < VQ_NEAR_TIE_REL * fabsf(bd)) counts += kTie;
                    if (out.zq)
                        finish_row_serial(zn4, reinterpret_cast<const float4*>(en32), out, K, row, code, loss_fx, bad);
                    done[i] = 0;                   // ready for the next call
                }
            }
        }
    }
    // ---------------- statistics ----------------
    if (stats) {
        const unsigned ties = __reduce_add_sync(VQ_FULL, counts & 1023u);
        const unsigned multi = __reduce_add_sync(VQ_FULL, (counts >> 10) & 1023u);
        bad = __reduce_add_sync(VQ_FULL, bad);
        unsigned long long lf = (unsigned long long)loss_fx;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) lf += __shfl_xor_sync(VQ_FULL, lf, off);
        if (lane == 0) {
            if (ties) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_NEAR_TIE_ROWS), (unsigned long long)ties);
            if (multi) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_AMBIGUOUS_ROWS), (unsigned long long)multi);
            if (lf) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_LOSS_FIXED), lf);
            if (bad) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_NONFINITE), (unsigned long long)bad);
        }
    }
}

}  // namespace tc16

bool tc16_supported(int64_t T, int K, int D) {
    static const bool disabled = getenv("VQ_TC16_DISABLE") && atoi(getenv("VQ_TC16_DISABLE")) != 0;
    // code ids travel as 16-bit fields of the rescoring items; the threshold logic needs whole 512-code groups
    return !disabled && D == tc16::kD && K >= tc16::kGroupCols && (K % tc16::kGroupCols) == 0 && K <= 65536 && T >= 256;
}

size_t tc16_workspace_bytes(int64_t T) { return (size_t)(T > 0 ? T : 1) * tc16::kRecordBytes; }

cudaError_t launch_dist_tc16(const CUtensorMap& ma, const CUtensorMap& mb, int T, const float* zn32, const float* row_sq,
                             const CodebookView& cb, int* cand, int* flagged, int* n_flagged, int64_t* stats,
                             void* records, cudaStream_t s) {
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
    int4* rec = static_cast<int4*>(records);
    const int n_row_tiles = (T + tc16::kRowsPerCta - 1) / tc16::kRowsPerCta;
    const int grid = n_row_tiles < sm_count() ? n_row_tiles : sm_count();
    cudaError_t e;
    if (service_low)
        e = launch_pdl(tc16::k_dist_tc16<false>, dim3(grid), dim3(tc16::kThreads), L.total + 1024, s, ma, mb, T, cb.K, cb.info, rec,
                       cand, flagged, n_flagged, stats);
    else
        e = launch_pdl(tc16::k_dist_tc16<true>, dim3(grid), dim3(tc16::kThreads), L.total + 1024, s, ma, mb, T, cb.K, cb.info, rec,
                       cand, flagged, n_flagged, stats);
    count_launch();
    tc::instrument_report(s, grid);
    return e != cudaSuccess ? e : cudaGetLastError();
}

// Exact indices and the finish pass behind the filter (see k_exact_finish16).  zq_tok / hist may be null.
// done_counters: kFlaggedCap zeroed ints; partial_ws: kFlaggedCap * 32 * 16 bytes.  Listed rows beyond kFlaggedCap
// (degenerate inputs only) are left in cand[] = -1 for the caller's overflow path.
cudaError_t launch_exact_finish16(const void* records, const float* zn32, const float* row_sq, const CodebookView& cb, int64_t T,
                                  const int* flagged, const int* n_flagged, int* done_counters, void* partial_ws, int* cand,
                                  float* zq_tok, int64_t* idx_out, int32_t* hist, int64_t* seg_sums, int64_t* stats,
                                  cudaStream_t s) {
    const int cap = (int)(T < kFlaggedCap ? T : kFlaggedCap);
    const int rows_per_block = tc16::kExactThreads / 8;
    int64_t blocks = (T + rows_per_block - 1) / rows_per_block;
    static const int grid_mult = getenv("VQ_EXACT_GRID_MULT") ? atoi(getenv("VQ_EXACT_GRID_MULT")) : 32;
    const int64_t grid_cap = (int64_t)sm_count() * grid_mult;    // 32 = 8 resident blocks per SM x 4 waves
    if (blocks > grid_cap) blocks = grid_cap;
    tc16::FinishOut out;
    out.zq = zq_tok; out.idx = idx_out; out.hist = hist;
    out.seg = zq_tok ? reinterpret_cast<unsigned long long*>(seg_sums) : nullptr;
    cudaError_t e = launch_pdl(tc16::k_exact_finish16, dim3((unsigned)blocks), dim3(tc16::kExactThreads), 0, s,
                               static_cast<const int4*>(records), zn32, row_sq, cb.en32, reinterpret_cast<const float4*>(cb.en32c),
                               cb.csq_cell, (int)T, cb.K, flagged, n_flagged, cap, static_cast<tc16::FlaggedPartial*>(partial_ws),
                               done_counters, cand, out, stats);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace vq
