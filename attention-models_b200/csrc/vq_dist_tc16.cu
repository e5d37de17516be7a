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
// Two kernels.  k_dist_tc16 is the filter: it writes one 16-byte verdict record per row -- the ids of the up to 8
// cells (8 codes each) that can still hold the winner.  k_exact_finish16 turns records into indices with the exact
// fp32 formula and does everything behind the search in the same pass (idx, histogram, z_q, loss partial, segment
// sums), plus the exhaustive search of the few rows the filter could not decide.  Rescoring inside the tensor-core
// kernel was tried first and cost more than it hid: four rescoring warps per SM are instruction-latency bound and
// their scattered loads saturate the L1 wavefront pipe under the epilogue; as its own kernel the same work runs at
// full occupancy.  Rescoring is warp-cooperative: 8 lanes share one row, lane m owns member m of every cell and each
// load instruction fetches whole 128-byte lines of the cell-interleaved codebook copy.
//
// Shape of the computation: one persistent CTA per SM, 256 token rows per CTA (two M = 128 MMA row tiles),
// codebook streamed by TMA in 128-code stages, 2 x 2 accumulator tiles of 128 columns in TMEM.
// A "group" is 4 n-tiles = 512 codes; a row keeps 64 slot maxima per group (32 packed registers; slot hs
// covers codes g*512 + hs + 64*m, m = 0..7 -- a "cell"), the slot maxima of the best four groups are parked in
// shared memory, and at the end of the row tile every cell within 2*eps of the row's best score goes into the record.
//
// What bounds the filter (tools/ubench_mma.cu, profiles/r02_ubench_mma.txt): an M = 128, K = 16 tcgen05.mma with both
// operands in shared memory costs ~60 cycles + 0.42 cycles per column -- 107 cycles at N = 128 against the nominal 64
// -- and a D = 32 tile is just two of them: 2 x 107 cycles per 128 x 128 tile is exactly the ~58 % tensor-pipe
// activity ncu reports.  k_dist_tc16_ts below keeps the token rows in tensor memory instead (83 cycles per MMA).
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
// MMA issuer warps.  A thread issues one tcgen05.mma per ~110-140 cycles however small the instruction is, and several
// issuer warps overlap perfectly (tools/ubench_mma.cu: 137 / 69 / 34 cycles per N = 128 MMA aggregated over 1 / 2 / 4
// issuers): at D = 32 a tile is two MMAs, so it is the ISSUE rate of a warp, not the tensor pipe, that paces the kernel.
// 4 issuers = one per accumulator stage (row half x n-tile parity); 2 = one per row half (the first-generation layout).
#ifndef VQ_TC16_ISSUERS
#define VQ_TC16_ISSUERS 4
#endif
constexpr int kIssuers = VQ_TC16_ISSUERS;
static_assert(kIssuers == 2 || kIssuers == 4, "2 or 4 MMA issuer warps");
// warps 0-7 epilogue; 8 TMA; then the issuers and the TMEM allocator: 384 threads (9, 11 issue; 10 allocates) or 512
// (9-12 issue; 13 allocates).  setmaxnreg moves registers inside the pool the block was launched with.
constexpr int kThreads = kIssuers == 2 ? 384 : 512;
constexpr int kRegsService = 40, kRegsEpilogue = kIssuers == 2 ? 232 : 216;
static_assert((kThreads - 256) * kRegsService + 256 * kRegsEpilogue <= (65536 / kThreads / 8 * 8) * kThreads, "register pool overcommitted");
// |fp16-pipeline score - exact dot| <= eps: 1.1e-3 for the fp16 operands (as in vq_dist_tc.cu) + one fp16 ulp
// per accumulator rounding (2^-11 below 1, 2^-10 in [1, 2)); the threshold is additionally rounded down to fp16
constexpr float kTwoEps = 2.f * (1.1e-3f + 4.8829e-4f + 4.8829e-4f);
constexpr float kTwoEpsNearOne = 2.f * (1.1e-3f + 9.7657e-4f + 9.7657e-4f);
constexpr float kMinThreshold = 1e-3f;              // the filter only trusts rows whose threshold is safely positive
// c_format = F16 (0), a/b = F16 (0), K-major both, N = 128, M = 128
constexpr uint32_t kIdesc = ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

constexpr int kSnapRow = 144;                       // 32 packed slot registers + 16 B pad (conflict-free STS.128)
constexpr int kSnapArea = kRowsPerCta * kSnapRow;
// verdict record, 16 bytes per row: up to 8 surviving cells as 16-bit ids (group * 64 + slot), 0xFFFF = none, filled
// from the low end; all-ones = the filter could not decide the row (it is on the list for the exhaustive search)
constexpr int kRecordBytes = 16;
constexpr int kRecCells = 8;
constexpr int kAreas = 4;

struct SmemLayout {
    uint32_t a, b, snap, bars, tmem_slot, xch, total;
};
constexpr int kSnapRowH = 128;                      // halves = 1: unpadded snapshot rows, 16-byte chunks swizzled by the row
constexpr int kXchBytes = 128 * (2 * 4 + 2 * 4 + 8);  // per row: two best scores, two "decided" flags, one record half
// halves = 2: a CTA owns 256 rows (two MMA row halves).  halves = 1: 128 rows, CTAs run in clusters of two that share the
// codebook stream (see k_dist_tc16); the ring then has 8 stages: with four issuers taking the tiles in turn a stage must
// always come back to the same issuer (a parity wait by a thread that skipped a phase returns early).
__host__ __device__ constexpr int b_stages16(int halves) { return halves == 1 ? 8 : kBStages; }
__host__ __device__ inline SmemLayout smem_layout(int halves = 2) {
    SmemLayout L;
    L.a = 0;
    L.b = L.a + kAStages * (kABytes / 2 * halves);
    L.snap = L.b + b_stages16(halves) * kBStageBytes;
    L.bars = L.snap + (halves == 1 ? kAreas * 256 * kSnapRowH : kAreas * kSnapArea);
    L.tmem_slot = L.bars + 8 * 32;
    L.xch = L.tmem_slot + 16;
    L.total = L.xch + (halves == 1 ? kXchBytes : 0);
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

__device__ __forceinline__ void pair_barrier(int id) {      // the two warps that own the same 32 rows (64 threads)
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
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
//
// HALVES = 1 (VQ_TC16_CLUSTER=1; MEASURED SLOWER, kept as the measured alternative): a CTA owns ONE 128-row half and all
// four accumulator stages belong to it; CTAs run in clusters of two that share the codebook stream (each fetches one
// 64-code half of every tile and TMA-multicasts it into both; a stage is released by both CTAs' issuers through a
// multicast tcgen05.commit).  The idea: with 256 rows per CTA a row half has only two accumulator stages, and the hand-off
// chain of a stage (MMAs retire -> mbarrier -> the epilogue warp wakes -> tcgen05.ld of the tile -> release -> the issuer
// wakes) is ~570 cycles for ~128 cycles of tensor work; four stages per half let the MMAs run up to four tiles ahead of
// the drain.  Issuer q owns stage q and the tiles n = q mod 4; eight epilogue warps drain them, two per TMEM lane
// quarter taking alternate tiles (see the epilogue).  Correct on the first run both times, and: 176.9 us with four
// epilogue warps (issuers waiting for free stages 83 % of the time: one warp per scheduler cannot drain a tile in less
// than ~400 cycles), 173.9 us with eight (two threads per row on alternate tiles), 178.2 us with eight and one hand-off
// per 256-column stage (the form below: the hand-off is not what costs) against 124.1 us for the 256-row kernel on the
// same box.  What the instrumented build shows: an epilogue warp is busy ~500-620 cycles per 32 x 128 tile it drains
// however the tiles are handed over (510 in the 256-row kernel; here every thread also runs the per-group bookkeeping,
// ~350 cycles, for its half of the tiles), two such warps per scheduler, so the drain - not the accumulator ring -
// paces the kernel at D = 32, where a tile is only two MMAs; and
// the MMAs themselves cannot go below ~77-84 cycles each per SM (tools/ubench_mma.cu; the D = 256 kernel, which has 16
// MMAs per tile to hide its drain behind, sits at 77), i.e. ~75 us for this shape.  Getting past both needs
// cta_group::2 MMAs (one instruction per CTA pair: half the issue cost per SM) and a cheaper drain, not more stages.
template <bool kServiceHigh, int HALVES>
__global__ void __launch_bounds__(kThreads, 1)
k_dist_tc16(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int T, int K,
            const int* __restrict__ cb_info, int4* __restrict__ rec, int* __restrict__ cand,
            int* __restrict__ flagged, int* __restrict__ n_flagged, int64_t* __restrict__ stats) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    static_assert(HALVES == 2 || kIssuers == 4, "the one-half form has one issuer per accumulator stage");
    constexpr int CL = HALVES == 1 ? 2 : 1;                   // CTAs per cluster (launch attribute)
    constexpr int kRows = 128 * HALVES;                       // token rows of a CTA
    constexpr int kABytesH = kRows * kD * 2;
    constexpr int kBS = b_stages16(HALVES);
    constexpr uint32_t kSnapAreaH = kRows * kSnapRow;
    const SmemLayout L = smem_layout(HALVES);
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
    constexpr int kTmaWarp = kServiceHigh ? 8 : 0, kMmaWarp = kTmaWarp + 1;
    constexpr int kAllocWarp = kIssuers == 2 ? kTmaWarp + 2 : kTmaWarp + 5;     // (2 issuers: warps kMmaWarp and kMmaWarp + 2)
    // a work unit is the rows of one cluster step: CL consecutive row tiles, one per CTA (a tile past the end reads zeros
    // and writes nothing)
    const uint32_t cta_rank = CL == 2 ? cluster_cta_rank() : 0u;
    const int unit0 = blockIdx.x / CL, unit_step = gridDim.x / CL;
    const int n_units = ((T + kRows - 1) / kRows + CL - 1) / CL;
    const int n_tiles = K / kTileN;
    const int n_groups = n_tiles / kGroupTiles;

    if (warp == kMmaWarp && lane == 0) {
        // a stage is read by both row halves' issuers, or by its issuer in either CTA of the cluster
        for (int s = 0; s < kBS; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 2); }
        // accumulator hand-off.  256-row CTAs: four stages of 128 columns (tile parity x row half), one issuer each, drained
        // by the half's four warps.  One row half: TWO stages of 256 columns (= two code tiles, one issuer each) - a
        // hand-off (wait for the stage, tcgen05.ld, release) costs an epilogue warp ~350 cycles whatever the stage holds
        // against ~110 for folding a tile, so it is paid once per 256 codes.
        for (int q = 0; q < 4; ++q) { mbar_init(t_full(q), HALVES == 1 ? 2 : 1); mbar_init(t_empty(q), 4); }
        for (int s = 0; s < 2; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), kIssuers); }
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
    if constexpr (CL == 2) cluster_sync_all();      // the peer's barriers are initialised before anything is sent to them
    tc_fence_after();
    // barrier init and the TMEM allocation above overlap the tail of the token prep kernel (programmatic dependent
    // launch); nothing before this line reads what that kernel writes
    pdl_trigger();
    pdl_wait();
    // every role reads the TMEM base address itself, after its setmaxnreg: a value kept live across the role
    // split ends up in a local-memory spill slot that the epilogue would reload once per tile

    if (warp >= kTmaWarp && warp < kTmaWarp + (kThreads - 256) / 32) {
        reg_dec<kRegsService>();
        if (warp == kTmaWarp) {
            // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
            uint32_t b_cnt = 0;
            int it = 0;
            VQ_INSTR_BEGIN();
            for (int u = unit0; u < n_units; u += unit_step, ++it) {
                const int rt = u * CL + (int)cta_rank;
                const int as = it & 1;
                VQ_TIMED_WAIT(0, a_empty(as), (((uint32_t)(it >> 1)) & 1u) ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(a_full(as), kABytesH);
                    tma_load_2d(smem_base + L.a + as * kABytesH, &tm_a, a_full(as), 0, rt * kRows);
                }
                __syncwarp();
                for (int n = 0; n < n_tiles; ++n, ++b_cnt) {
                    const int s = b_cnt % kBS;
                    VQ_TIMED_WAIT(1, b_empty(s), ((b_cnt / kBS) & 1u) ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(b_full(s), kBStageBytes);
                        if constexpr (CL == 2)      // this CTA's 64-code half of the tile (tm_b boxes are 64 codes), into both CTAs
                            tma_load_2d_multicast(smem_base + L.b + s * kBStageBytes + cta_rank * (kBStageBytes / 2), &tm_b, b_full(s),
                                                  0, n * kTileN + (int)cta_rank * (kTileN / 2), (uint16_t)3);
                        else
                            tma_load_2d(smem_base + L.b + s * kBStageBytes, &tm_b, b_full(s), 0, n * kTileN);
                    }
                    __syncwarp();
                }
            }
            VQ_INSTR_END(0, 2);
        } else if (kIssuers == 2 ? (warp == kMmaWarp || warp == kMmaWarp + 2) : (warp >= kMmaWarp && warp < kMmaWarp + 4)) {
            // ===================== MMA issuers (whole warp loops, one elected lane issues) ==========
            // r: row half; with 4 issuers also p0: the n-tile parity (= accumulator stage 2 p0 + r) this warp serves
            const int iw = warp - kMmaWarp;
            const int r = HALVES == 1 ? 0 : (kIssuers == 2 ? (iw >> 1) : (iw & 1));
            const int p0 = HALVES == 1 ? iw : (kIssuers == 2 ? 0 : (iw >> 1));
            constexpr int kStep = HALVES == 1 ? 4 : (kIssuers == 2 ? 1 : 2);   // n-tiles between two tiles of this issuer
            const uint32_t tmem_base = *tmem_slot;
            uint32_t t_cnt = 0;                                        // tiles issued by this warp
            int it = 0;
            VQ_INSTR_BEGIN();
            for (int u = unit0; u < n_units; u += unit_step, ++it) {
                const int as = it & 1;
                VQ_TIMED_WAIT(0, a_full(as), ((uint32_t)(it >> 1)) & 1u);
                tc_fence_after();
                const uint32_t a_addr = smem_base + L.a + as * kABytesH + r * (128 * 64);
                for (int n = p0; n < n_tiles; n += kStep, ++t_cnt) {
                    const int q = HALVES == 1 ? iw : 2 * (n & 1) + r;    // accumulator stage: own, or 2 (tile parity) + row half
                    const uint32_t b_cnt = (uint32_t)it * (uint32_t)n_tiles + (uint32_t)n;
                    const int s = b_cnt % kBS;
                    VQ_TIMED_WAIT(2, b_full(s), (b_cnt / kBS) & 1u);
                    // uses of the stage so far: every tile this warp issued into it
                    const uint32_t use = (HALVES == 2 && kIssuers == 2) ? (t_cnt >> 1) : t_cnt;
                    const int qb = HALVES == 1 ? (q >> 1) : q;      // hand-off barriers: per 256-column stage, or per tile
                    VQ_TIMED_WAIT(1, t_empty(qb), (use & 1u) ^ 1u);
                    tc_fence_after();
                    if (iw == 0 && lane == 0) VQ_TRACE(0, (int)t_cnt, 0);
                    const uint32_t b_addr = smem_base + L.b + s * kBStageBytes;
                    VQ_TIMED_BEGIN();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            umma_f16(tmem_base + (uint32_t)(q * kTileN), umma_desc(a_addr + k * 32),
                                     umma_desc(b_addr + k * 32), kIdesc, (uint32_t)k);
                        umma_commit(t_full(qb));
                        if constexpr (CL == 2) umma_commit_multicast(b_empty(s), (uint16_t)3);
                        else umma_commit(b_empty(s));
                    }
                    __syncwarp();
                    if (iw == 0 && lane == 0) VQ_TRACE(0, (int)t_cnt, 1);
                    VQ_TIMED_END(3);
                }
                if (elect_one()) umma_commit(a_empty(as));
                __syncwarp();
            }
            const int r_report = iw;
            (void)r_report;
            if (iw == 0) { VQ_INSTR_END(3, 4); }
        }
    } else if (HALVES == 1) {
        // ===================== epilogue, one row half: TWO threads per row =====================
        // An epilogue warp's per-tile chain (wait for the tile, tcgen05.ld of 64 registers, release, fold) takes ~400
        // cycles however idle its scheduler is (measured: four warps draining every tile of their rows made the kernel
        // 177 us, the issuers waiting for free stages 83 % of the time), so the drain needs two warps per scheduler, as
        // the 256-row kernel has.  Warp (quarter, par) drains accumulator stage `par` of its 32 rows: of a 512-code group
        // it sees tiles 2 par and 2 par + 1, and its 32 packed slot registers hold the maxima of 64 sub-cells of 4 codes -
        // sub-cell (j, half) of thread par and of thread par ^ 1 together are cell (j, half) of the usual layout (2
        // columns x 4 tiles), so the verdict records and everything behind them are unchanged.  Each thread ranks the
        // groups by its own sub-cell maxima and snapshots its own best four; the row's two threads meet once per row
        // tile: the row's best score is the larger of theirs, each marks its sub-cells within 2 eps of it, the row is
        // decided iff both could (fifth-best own group below the threshold, at most four cells each), and each writes one
        // half of the record (thread 1 drops the cell ids thread 0 already lists).
        reg_inc<kRegsEpilogue>();
        const int e = warp - kEpiWarp0;
        const int par = e >> 2;
        const int quarter = warp & 3;                // TMEM lane quarter this warp may read (warp id mod 4)
        const int row_in_cta = quarter * 32 + lane;
        const int pair_id = 1 + quarter;
        const uint32_t tbase = *tmem_slot + ((uint32_t)(quarter * 32) << 16);
        constexpr uint32_t kAreaStride = 256 * kSnapRowH;
        const uint32_t snap0 = smem_base + L.snap + (uint32_t)(par * 128 + row_in_cta) * kSnapRowH;
        const uint32_t swz = (uint32_t)(lane & 7);
        volatile int* xch = reinterpret_cast<volatile int*>(smem + L.xch);
        volatile unsigned long long* xrec = reinterpret_cast<volatile unsigned long long*>(smem + L.xch + 128 * 16);
        const bool force_exhaustive = codebook_degenerate(cb_info);
        uint32_t uses = 0;                           // hand-offs of this warp's stage so far (one per group)
        int it = 0;
        uint32_t bufA[64], bufB[64];
        VQ_INSTR_BEGIN();
        for (int u = unit0; u < n_units; u += unit_step, ++it) {
            const int rt = u * CL + (int)cta_rank;
            uint32_t slot[32];
            int m1 = -32768, m2 = -32768, m3 = -32768, m4 = -32768, m5 = -32768;
            int g1 = 0, g2 = 0, g3 = 0, g4 = 0;
            uint32_t a1 = 0, a2 = 1, a3 = 2, a4 = 3;  // snapshot areas of this thread's best four groups
            auto group_end = [&](const int g) {
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
                    const uint32_t dst = snap0 + a4 * kAreaStride;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16 * ((uint32_t)q ^ swz)), "r"(slot[4 * q]),
                                     "r"(slot[4 * q + 1]), "r"(slot[4 * q + 2]), "r"(slot[4 * q + 3])
                                     : "memory");
                }
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
            };
            for (int g = 0; g < n_groups; ++g) {
                // stage `par` holds the tiles 2 par and 2 par + 1 of the group
                VQ_TIMED_WAIT(0, t_full(par), uses & 1u);
                tc_fence_after();
                tmem_ld_tile(tbase + (uint32_t)((2 * par) * kTileN), bufA);
                tmem_ld_tile(tbase + (uint32_t)((2 * par + 1) * kTileN), bufB);
                if (g > 0) group_end(g - 1);                      // the previous group's bookkeeping hides in the loads' latency
                VQ_TIMED_BEGIN();
                tmem_ld_wait();                                   // both tiles are in registers: the stage is free
                VQ_TIMED_END(2);
                tc_fence_before();
                if (lane == 0) mbar_arrive(t_empty(par));
                ++uses;
#pragma unroll
                for (int j = 0; j < 32; ++j) slot[j] = __vimax3_s16x2(__vmaxs2(bufA[j], bufA[j + 32]), bufB[j], bufB[j + 32]);
            }
            group_end(n_groups - 1);
            // ---- row verdict, taken by the row's two threads ----
            xch[row_in_cta * 2 + par] = m1;
            pair_barrier(pair_id);
            const int m1_row = max(m1, xch[row_in_cta * 2 + (par ^ 1)]);
            const float m1f = __half2float(__ushort_as_half((unsigned short)(m1_row & 0xFFFF)));
            const float thr_f = m1f - (m1f < 0.9f ? kTwoEps : kTwoEpsNearOne);
            const bool thr_ok = (m1_row >= 0) && (m1_row < 0x7C00) && (thr_f >= kMinThreshold);
            const int thr = thr_ok ? (int)__half_as_ushort(__float2half_rd(thr_f)) : 0x7BFF;
            const uint32_t thr2 = (uint32_t)thr * 0x10001u;
            uint32_t mask[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            const int mv[4] = {m1, m2, m3, m4};
            const uint32_t av[4] = {a1, a2, a3, a4};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if (mv[a] >= thr) {
                    const uint32_t src = snap0 + av[a] * kAreaStride;
                    uint32_t kept[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(kept[4 * q]), "=r"(kept[4 * q + 1]), "=r"(kept[4 * q + 2]), "=r"(kept[4 * q + 3])
                                     : "r"(src + 16 * ((uint32_t)q ^ swz)));
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        bool ph, pl;
                        (void)__vibmax_s16x2(kept[j], thr2, &ph, &pl);      // per half: kept >= thr
                        const uint32_t bits = (pl ? (1u << ((2 * j) & 31)) : 0u) | (ph ? (1u << ((2 * j + 1) & 31)) : 0u);
                        mask[2 * a + (j >> 4)] |= bits;
                    }
                }
            }
            const int n_cand = __popc(mask[0]) + __popc(mask[1]) + __popc(mask[2]) + __popc(mask[3]) + __popc(mask[4]) + __popc(mask[5]) +
                               __popc(mask[6]) + __popc(mask[7]);
            const bool mine_ok = thr_ok && (m5 < thr) && (n_cand <= kRecCells / 2) && !force_exhaustive;
            // this thread's half of the record: its surviving cells as 16-bit ids (group * 64 + slot), 0xFFFF = none
            unsigned long long half_rec = ~0ull;
            if (mine_ok) {
                const int gv[4] = {g1, g2, g3, g4};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t w = mask[2 * a + h];
                        while (w) {
                            const uint32_t id = (uint32_t)gv[a] * 64u + 32u * h + (uint32_t)(__ffs((int)w) - 1);
                            w &= w - 1;
                            half_rec = (half_rec << 16) | id;
                        }
                    }
            }
            xch[256 + row_in_cta * 2 + par] = mine_ok ? 1 : 0;
            if (par == 0) xrec[row_in_cta] = half_rec;
            pair_barrier(pair_id);
            const bool decided = mine_ok && (xch[256 + row_in_cta * 2 + (par ^ 1)] != 0);
            if (par == 1 && decided) {
                // cells both threads list (both halves of the cell near the best score) stay in thread 0's half only
                const unsigned long long other = xrec[row_in_cta];
                unsigned long long mine = ~0ull;
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    const uint32_t id = (uint32_t)(half_rec >> (16 * f)) & 0xFFFFu;
                    bool dup = id == 0xFFFFu;
#pragma unroll
                    for (int f2 = 0; f2 < 4; ++f2) dup = dup || (id == ((uint32_t)(other >> (16 * f2)) & 0xFFFFu));
                    if (!dup) mine = (mine << 16) | id;
                }
                half_rec = mine;
            }
            if (!decided) half_rec = ~0ull;
            const int row = rt * kRows + row_in_cta;
            const bool in_range = row < T;
            if (in_range) reinterpret_cast<uint2*>(rec)[2 * (int64_t)row + par] = make_uint2((uint32_t)half_rec, (uint32_t)(half_rec >> 32));
            if (par == 0) {                          // warp-uniform: one of the two warps lists the undecided rows
                const bool flag = in_range && !decided;
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
            pair_barrier(pair_id);                   // the exchange words are free for the next row tile
        }
        if (threadIdx.x == kEpiWarp0 * 32) { VQ_INSTR_END(8, 3); }
    } else {
        // ===================== epilogue: one thread per row (8 warps, or 4 with one row half) =====================
        reg_inc<kRegsEpilogue>();
        const int e = warp - kEpiWarp0;
        const int r_sub = HALVES == 2 ? (e >> 2) : 0;   // which 128-row MMA tile
        const int quarter = warp & 3;                // TMEM lane quarter this warp may read (warp id mod 4)
        const int row_in_cta = r_sub * 128 + quarter * 32 + lane;
        const uint32_t tbase = *tmem_slot + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(r_sub * kTileN);
        // tile b of a group (n_tiles is a multiple of 4, so b = tile count mod 4): its accumulator stage, the stage's
        // TMEM column, and the parity of the stage's use the running tile count t stands for
        auto stage_of = [&](int b) { return HALVES == 2 ? 2 * (b & 1) + r_sub : b; };
        auto col_of = [&](int b) { return (uint32_t)((HALVES == 2 ? 2 * (b & 1) : b) * kTileN); };
        auto phase_of = [&](uint32_t t) { return HALVES == 2 ? (t >> 1) & 1u : (t >> 2) & 1u; };
        const uint32_t snap0 = smem_base + L.snap + (uint32_t)row_in_cta * kSnapRow;
        const bool force_exhaustive = codebook_degenerate(cb_info);
        uint32_t t_cnt = 0;                          // n-tiles drained so far (same sequence as the MMA warp)
        int it = 0;
        uint32_t bufA[64], bufB[64];
        VQ_INSTR_BEGIN();
        for (int u = unit0; u < n_units; u += unit_step, ++it) {
            const int rt = u * CL + (int)cta_rank;
            uint32_t slot[32];
            int m1 = -32768, m2 = -32768, m3 = -32768, m4 = -32768, m5 = -32768;
            int g1 = 0, g2 = 0, g3 = 0, g4 = 0;
            uint32_t a1 = 0, a2 = 1, a3 = 2, a4 = 3;  // snapshot areas of the best four groups
            auto group_end = [&](const int g) {
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
                    const uint32_t dst = snap0 + a4 * kSnapAreaH;
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
            };
            VQ_TIMED_WAIT(0, t_full(stage_of(0)), phase_of(t_cnt));
            tc_fence_after();
            tmem_ld_tile(tbase + col_of(0), bufA);
            for (int g = 0; g < n_groups; ++g) {
#pragma unroll
                for (int b = 0; b < kGroupTiles; ++b) {
                    uint32_t (&cur)[64] = (b & 1) ? bufB : bufA;
                    uint32_t (&nxt)[64] = (b & 1) ? bufA : bufB;
                    const bool has_next = b < kGroupTiles - 1 || g + 1 < n_groups;
                    VQ_TIMED_BEGIN();
                    tmem_ld_wait();                                   // tile b is in registers: its TMEM stage is free
                    VQ_TIMED_END(2);
#ifdef VQ_TC_INSTRUMENT
                    if (e == 0 && lane == 0) { asm volatile("" ::"r"(cur[0]), "r"(cur[63])); VQ_TRACE(1, (int)t_cnt, 1); }
#endif
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(t_empty(stage_of(b)));
                    if (e == 0 && lane == 0) VQ_TRACE(1, (int)t_cnt, 2);
                    ++t_cnt;
                    if (has_next) {
                        // (probing this barrier with mbarrier.test_wait ahead of the tcgen05.wait::ld above was measured: no gain)
                        VQ_TIMED_WAIT(0, t_full(stage_of((b + 1) & 3)), phase_of(t_cnt));
                        if (e == 0 && lane == 0) VQ_TRACE(1, (int)t_cnt, 0);
                        tc_fence_after();
                        tmem_ld_tile(tbase + col_of((b + 1) & 3), nxt);
                    }
                    if (b == 0) {
                        // the previous group's bookkeeping runs here, behind the hand-off of this tile's TMEM stage: the
                        // tensor pipe refills that stage meanwhile instead of idling with both stages full
                        if (g > 0) group_end(g - 1);
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
            }
            group_end(n_groups - 1);
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
                    const uint32_t src = snap0 + av[a] * kSnapAreaH;
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
            // decided iff no further group can hold the winner and the survivors fit a record
            const int n_cand = __popc(mask[0]) + __popc(mask[1]) + __popc(mask[2]) + __popc(mask[3]) + __popc(mask[4]) + __popc(mask[5]) +
                               __popc(mask[6]) + __popc(mask[7]);
            const bool decided = thr_ok && (m5 < thr) && ((mask[0] | mask[1]) != 0) && (n_cand <= kRecCells) && !force_exhaustive;
            const int row = rt * kRows + row_in_cta;
            const bool in_range = row < T;
            const bool flag = in_range && !decided;
            // verdict record: the surviving cells as 16-bit cell ids (group * 64 + slot), pushed into a 128-bit shift register
            // from the low end; 0xFFFF = none.  An undecided row keeps all-ones (its first field says "listed").
            unsigned long long lo = ~0ull, hi = ~0ull;
            if (decided) {
                const int gv[4] = {g1, g2, g3, g4};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t w = mask[2 * a + h];
                        while (w) {
                            const uint32_t id = (uint32_t)gv[a] * 64u + 32u * h + (uint32_t)(__ffs((int)w) - 1);
                            w &= w - 1;
                            hi = (hi << 16) | (lo >> 48);
                            lo = (lo << 16) | id;
                        }
                    }
            }
            if (in_range) rec[row] = make_int4((int)(uint32_t)lo, (int)(uint32_t)(lo >> 32), (int)(uint32_t)hi, (int)(uint32_t)(hi >> 32));
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
    if constexpr (CL == 2) cluster_sync_all();      // nothing of the peer's (multicast tiles, stage releases) is still on its way
    if (warp == kAllocWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(512u) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------------------------
// The filter with a WIDE DRAIN: 16 epilogue warps instead of 8.
//
// The idea: with four MMA issuers the first kernel above looked bound by the serial latency of an epilogue warp's
// per-tile loop (wait for the finished tile -> tcgen05.ld -> release -> fold, ~500 cycles) with only two such warps per
// scheduler.  MEASURED: no faster (see tc16_wide_drain below); kept as the measured alternative.  Here every 128 x 128 accumulator tile is drained by EIGHT warps -- two per TMEM lane
// quarter, each taking one 64-column half (tcgen05.ld.x32.pack::16b: 32 registers) -- so four epilogue warps per scheduler
// hide each other's hand-off waits, and a thread needs 16 slot registers + one 32-register load buffer (96-register
// budget, 768 threads).  A thread folds column c with c + 32 of its half over the 4 tiles of a group: a slot is again a
// cell of 8 codes (cell layout 3 in vq_kernels.h), 32 cells per thread and group, 64-byte snapshots (swizzled chunks
// instead of row padding: 4 areas x 512 threads x 64 B = 128 KB).  A row's verdict is taken by its TWO threads: they
// exchange their best scores through shared memory, each marks the cells of its half within 2 eps of the row's best
// and fills its own half of the 16-byte record (up to 4 cell ids each); the row is decided iff both halves are.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kThreadsW = 768;         // warps 0-15 epilogue, 16 TMA, 17-20 MMA issuers (one per accumulator stage), 21 TMEM allocator
constexpr int kRegsServiceW = 40, kRegsEpilogueW = 96;
static_assert(256 * kRegsServiceW + 512 * kRegsEpilogueW <= (65536 / kThreadsW / 8 * 8) * kThreadsW, "register pool overcommitted");
constexpr int kSnapRowW = 64;                              // 16 packed slot registers, chunk order swizzled by the lane
constexpr int kSnapAreaW = 2 * kRowsPerCta * kSnapRowW;    // 32 KB: both column halves of 256 rows
struct SmemLayoutW {
    uint32_t a, b, snap, xch, bars, tmem_slot, total;
};
__host__ __device__ inline SmemLayoutW smem_layout_w() {
    SmemLayoutW L;
    L.a = 0;
    L.b = L.a + kAStages * kABytes;
    L.snap = L.b + kBStages * kBStageBytes;
    L.xch = L.snap + kAreas * kSnapAreaW;                  // [2 rounds][256 rows][2 halves] ints: best score / decided
    L.bars = L.xch + 2 * kRowsPerCta * 2 * 4;
    L.tmem_slot = L.bars + 8 * 32;
    L.total = L.tmem_slot + 16;
    return L;
}
// 32 lanes x 64 columns of fp16 accumulators, two adjacent columns per register
__device__ __forceinline__ void tmem_ld_half_tile(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

__global__ void __launch_bounds__(kThreadsW, 1)
k_dist_tc16_w16(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int T, int K,
                const int* __restrict__ cb_info, uint2* __restrict__ rec, int* __restrict__ flagged, int* __restrict__ n_flagged,
                int64_t* __restrict__ stats) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const SmemLayoutW L = smem_layout_w();
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

    const int warp = __shfl_sync(VQ_FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    constexpr int kTmaWarp = 16, kMmaWarp = 17, kAllocWarp = 21;
    const int n_row_tiles = (T + kRowsPerCta - 1) / kRowsPerCta;
    const int n_tiles = K / kTileN;
    const int n_groups = n_tiles / kGroupTiles;

    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < kBStages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 2); }   // both halves' issuers commit
        for (int q = 0; q < 4; ++q) { mbar_init(t_full(q), 1); mbar_init(t_empty(q), 8); }          // 8 warps drain a stage
        for (int s = 0; s < 2; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 4); }          // 4 issuers
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
    pdl_trigger();
    pdl_wait();

    if (warp >= kTmaWarp) {
        reg_dec<kRegsServiceW>();
        if (warp == kTmaWarp) {
            // ===================== TMA producer =====================
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
        } else if (warp >= kMmaWarp && warp < kMmaWarp + 4) {
            // ===================== MMA issuers: one warp per accumulator stage (row half r, n-tile parity p0) =====================
            const int iw = warp - kMmaWarp;
            const int r = iw & 1, p0 = iw >> 1;
            const uint32_t tmem_base = *tmem_slot;
            uint32_t t_cnt = 0;
            int it = 0;
            VQ_INSTR_BEGIN();
            for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x, ++it) {
                const int as = it & 1;
                VQ_TIMED_WAIT(0, a_full(as), ((uint32_t)(it >> 1)) & 1u);
                tc_fence_after();
                const uint32_t a_addr = smem_base + L.a + as * kABytes + r * (128 * 64);
                for (int n = p0; n < n_tiles; n += 2, ++t_cnt) {
                    const uint32_t b_cnt = (uint32_t)it * (uint32_t)n_tiles + (uint32_t)n;
                    const int s = b_cnt % kBStages;
                    VQ_TIMED_WAIT(2, b_full(s), (b_cnt / kBStages) & 1u);
                    VQ_TIMED_WAIT(1, t_empty(2 * p0 + r), (t_cnt & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t b_addr = smem_base + L.b + s * kBStageBytes;
                    VQ_TIMED_BEGIN();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            umma_f16(tmem_base + (uint32_t)((p0 * 2 + r) * kTileN), umma_desc(a_addr + k * 32),
                                     umma_desc(b_addr + k * 32), kIdesc, (uint32_t)k);
                        umma_commit(t_full(2 * p0 + r));
                        umma_commit(b_empty(s));
                    }
                    __syncwarp();
                    VQ_TIMED_END(3);
                }
                if (elect_one()) umma_commit(a_empty(as));
                __syncwarp();
            }
            if (iw == 0) { VQ_INSTR_END(3, 4); }
        }
    } else {
        // ===================== epilogue: 16 warps; thread = (row, column half) =====================
        reg_inc<kRegsEpilogueW>();
        VQ_INSTR_BEGIN();
        const int quarter = warp & 3;                // TMEM lane quarter this warp may read (warp id mod 4)
        const int r_sub = (warp >> 2) & 1;           // which 128-row MMA tile
        const int h2 = warp >> 3;                    // which 64-column half of every tile
        const int row_in_cta = r_sub * 128 + quarter * 32 + lane;
        const uint32_t tbase = *tmem_slot + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(r_sub * kTileN + h2 * 64);
        const uint32_t snap0 = smem_base + L.snap + (uint32_t)(h2 * kRowsPerCta + row_in_cta) * kSnapRowW;
        const uint32_t swz = (uint32_t)((lane >> 1) & 3);          // chunk swizzle: conflict-free 16-byte accesses at a 64-byte stride
        volatile int* xch = reinterpret_cast<volatile int*>(smem + L.xch);
        const int pair_id = 1 + r_sub * 4 + quarter;               // named barrier of the two warps that share these 32 rows
        const bool force_exhaustive = codebook_degenerate(cb_info);
        uint32_t t_cnt = 0;                          // n-tiles drained so far (same sequence as the issuers)
        uint32_t bufA[32], bufB[32];
        for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x) {
            uint32_t slot[16];
            // the best four groups of this thread's column half as sorted keys (score << 16 | group << 2 | snapshot area):
            // one register each, and the sorted insert is 9 min / max
            // (k4 is always the smallest KEY, low bits included: it is the entry a better group pushes out)
            int k1 = (int)0x80000000 | 3, k2 = (int)0x80000000 | 2, k3 = (int)0x80000000 | 1, k4 = (int)0x80000000 | 0;
            int m5 = -32768;
            auto group_end = [&](const int g) {
                uint32_t t[6];
#pragma unroll
                for (int j = 0; j < 5; ++j) t[j] = __vimax3_s16x2(slot[3 * j], slot[3 * j + 1], slot[3 * j + 2]);
                t[5] = slot[15];
                const uint32_t pk = __vimax3_s16x2(__vimax3_s16x2(t[0], t[1], t[2]), __vimax3_s16x2(t[3], t[4], t[5]), t[5]);
                const int c1 = max(lo16(pk), hi16(pk));
                const uint32_t area = (uint32_t)k4 & 3u;            // the entry that drops out frees its area
                const int nk = (c1 << 16) | (g << 2) | (int)area;
                if (nk > k4) {
                    const uint32_t dst = snap0 + area * kSnapAreaW;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16 * ((uint32_t)q ^ swz)), "r"(slot[4 * q]),
                                     "r"(slot[4 * q + 1]), "r"(slot[4 * q + 2]), "r"(slot[4 * q + 3])
                                     : "memory");
                }
                const int lo1 = min(nk, k1);
                k1 = max(nk, k1);
                const int lo2 = min(lo1, k2);
                k2 = max(lo1, k2);
                const int lo3 = min(lo2, k3);
                k3 = max(lo2, k3);
                const int lo4 = min(lo3, k4);
                k4 = max(lo3, k4);
                m5 = max(m5, lo4 >> 16);
            };
            {
                VQ_TIMED_WAIT(0, t_full(r_sub), (t_cnt >> 1) & 1u);
                tc_fence_after();
                tmem_ld_half_tile(tbase, bufA);
            }
            for (int g = 0; g < n_groups; ++g) {
#pragma unroll
                for (int b = 0; b < kGroupTiles; ++b) {
                    uint32_t (&cur)[32] = (b & 1) ? bufB : bufA;
                    uint32_t (&nxt)[32] = (b & 1) ? bufA : bufB;
                    VQ_TIMED_BEGIN();
                    tmem_ld_wait();                                   // this half of tile b is in registers
                    VQ_TIMED_END(2);
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(t_empty(2 * (b & 1) + r_sub));
                    ++t_cnt;
                    if (b < kGroupTiles - 1 || g + 1 < n_groups) {
                        VQ_TIMED_WAIT(0, t_full(2 * ((b + 1) & 1) + r_sub), (t_cnt >> 1) & 1u);
                        tc_fence_after();
#if !defined(VQ_EXP) || VQ_EXP < 2
                        tmem_ld_half_tile(tbase + (uint32_t)(((b + 1) & 1) * 2 * kTileN), nxt);
#endif
                    }
#if defined(VQ_EXP) && VQ_EXP >= 1
                    // timing experiment only (results are garbage): no fold / bookkeeping
                    if (b == 0 && g == 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) slot[j] = cur[j] ^ cur[j + 16];
                    } else { slot[0] ^= cur[0] ^ cur[31]; }
#else
                    if (b == 0) {
                        if (g > 0) group_end(g - 1);                  // behind the hand-off of this tile's stage
#pragma unroll
                        for (int j = 0; j < 16; ++j) slot[j] = __vmaxs2(cur[j], cur[j + 16]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) slot[j] = __vimax3_s16x2(slot[j], cur[j], cur[j + 16]);
                    }
#endif
                }
            }
            group_end(n_groups - 1);
            const int m1 = k1 >> 16;
            // ---- row verdict, taken by the row's two threads ----
            xch[row_in_cta * 2 + h2] = m1;
            pair_barrier(pair_id);
            const int m1_row = max(m1, xch[row_in_cta * 2 + (h2 ^ 1)]);
            const float m1f = __half2float(__ushort_as_half((unsigned short)(m1_row & 0xFFFF)));
            const float thr_f = m1f - (m1f < 0.9f ? kTwoEps : kTwoEpsNearOne);
            const bool thr_ok = (m1_row >= 0) && (m1_row < 0x7C00) && (thr_f >= kMinThreshold);
            const int thr = thr_ok ? (int)__half_as_ushort(__float2half_rd(thr_f)) : 0x7BFF;
            const uint32_t thr2 = (uint32_t)thr * 0x10001u;
            uint32_t mask[4] = {0u, 0u, 0u, 0u};
            const int kv[4] = {k1, k2, k3, k4};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if ((kv[a] >> 16) >= thr) {
                    const uint32_t src = snap0 + ((uint32_t)kv[a] & 3u) * kSnapAreaW;
                    uint32_t kept[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(kept[4 * q]), "=r"(kept[4 * q + 1]), "=r"(kept[4 * q + 2]), "=r"(kept[4 * q + 3])
                                     : "r"(src + 16 * ((uint32_t)q ^ swz)));
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        bool ph, pl;
                        (void)__vibmax_s16x2(kept[j], thr2, &ph, &pl);      // per half: kept >= thr
                        mask[a] |= (pl ? (1u << (2 * j)) : 0u) | (ph ? (1u << (2 * j + 1)) : 0u);
                    }
                }
            }
            const int n_cand = __popc(mask[0]) + __popc(mask[1]) + __popc(mask[2]) + __popc(mask[3]);
            const bool mine_ok = thr_ok && (m5 < thr) && (n_cand <= kRecCells / 2) && !force_exhaustive;
            xch[2 * kRowsPerCta + row_in_cta * 2 + h2] = mine_ok ? 1 : 0;
            pair_barrier(pair_id);
            const bool decided = mine_ok && (xch[2 * kRowsPerCta + row_in_cta * 2 + (h2 ^ 1)] != 0);
            const int row = rt * kRowsPerCta + row_in_cta;
            const bool in_range = row < T;
            // this thread's half of the record: its surviving cells as 16-bit ids (group * 64 + half * 32 + slot), 0xFFFF = none
            unsigned long long half_rec = ~0ull;
            if (decided) {
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    uint32_t w = mask[a];
                    while (w) {
                        const uint32_t id = (uint32_t)((kv[a] >> 2) & 0x3FFF) * 64u + 32u * (uint32_t)h2 + (uint32_t)(__ffs((int)w) - 1);
                        w &= w - 1;
                        half_rec = (half_rec << 16) | id;
                    }
                }
            }
            if (in_range) rec[2 * (int64_t)row + h2] = make_uint2((uint32_t)half_rec, (uint32_t)(half_rec >> 32));
            if (h2 == 0) {                           // warp-uniform: one of the two warps lists the undecided rows
                const bool flag = in_range && !decided;
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
            pair_barrier(pair_id);                   // the exchange words are free for the next row tile
        }
        if (threadIdx.x == 0) { VQ_INSTR_END(8, 3); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kAllocWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The filter with the token rows in TENSOR MEMORY (tcgen05.mma "TS" form: A from TMEM, B from shared memory).
//
// Why: an M = 128, K = 16 MMA costs a fixed ~60 cycles on top of ~0.42 cycles per column when A comes from shared
// memory and ~30 when it comes from tensor memory (tools/ubench_mma.cu, profiles/r02_ubench_mma.txt: N = 128 takes
// 107 / 83 cycles against the nominal 64).  At codebook_dim = 32 a 128 x 128 tile is just two such instructions, so
// the SS kernel above is bound by that fixed cost (2 x 107 cycles per tile = the 58 % tensor-pipe activity ncu
// shows), not by its accumulator ring.  Here the row tile's fp16 unit rows are written once per row tile into 16
// TMEM columns per row half by the epilogue threads themselves (tcgen05.st; thread = row = TMEM lane) -- no A tile in
// shared memory, no TMA for it -- and every MMA reads them from there.
// TMEM: 3 accumulator stages of 128 columns shared by BOTH row halves (stage of the c-th (n-tile, half) pair:
// c mod 3), A at columns 384 + 32 * (row-tile parity) + 16 * half, three MMA issuer warps (one per stage: a wait on
// an mbarrier costs ~90 cycles even when its phase completed long ago, so one issuer walking every pair is too slow).
// MEASURED (B200, 262 144 x 8192): 142 us against 124 us for the SS kernel -- the A columns leave room for only three
// accumulator stages, and three stages cannot cover the issue -> commit -> tcgen05.ld -> release chain (~420 cycles
// before the next MMA of a stage can start): the epilogue waits for finished tiles 35 % of its time.  Kept behind
// VQ_TC16_TS=1 as the measured alternative; the default stays the SS kernel above.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kAccStages = 3;
constexpr int kACol = kAccStages * kTileN;
constexpr int kBStagesTs = 8;
constexpr int kThreadsTs = 384;        // warps 0-7 epilogue, 8 TMA, 9-11 MMA issuers (10 also allocates TMEM)
// setmaxnreg moves registers inside the pool the block was launched with (384 threads x 168): the epilogue's increase
// blocks forever if the two budgets add up to more than that
constexpr int kRegsServiceTs = 56, kRegsEpilogueTs = 224;
static_assert(128 * kRegsServiceTs + 256 * kRegsEpilogueTs <= 384 * 168, "register pool overcommitted");
struct SmemLayoutTs {
    uint32_t b, snap, bars, tmem_slot, total;
};
__host__ __device__ inline SmemLayoutTs smem_layout_ts() {
    SmemLayoutTs L;
    L.b = 0;
    L.snap = L.b + kBStagesTs * kBStageBytes;
    L.bars = L.snap + kAreas * kSnapArea;
    L.tmem_slot = L.bars + 8 * 32;
    L.total = L.tmem_slot + 16;
    return L;
}
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 16 columns: thread i writes its 16 registers to columns [col, col + 16) of TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint4 (&v)[4]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0].x), "r"(v[0].y), "r"(v[0].z), "r"(v[0].w), "r"(v[1].x), "r"(v[1].y), "r"(v[1].z), "r"(v[1].w),
          "r"(v[2].x), "r"(v[2].y), "r"(v[2].z), "r"(v[2].w), "r"(v[3].x), "r"(v[3].y), "r"(v[3].z), "r"(v[3].w)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kThreadsTs, 1)
k_dist_tc16_ts(const __grid_constant__ CUtensorMap tm_b, const uint4* __restrict__ zn16, int T, int K,
               const int* __restrict__ cb_info, int4* __restrict__ rec, int* __restrict__ flagged, int* __restrict__ n_flagged,
               int64_t* __restrict__ stats) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const SmemLayoutTs L = smem_layout_ts();
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + L.bars;
    auto b_full = [&](int s) { return bar_base + 8 * s; };
    auto b_empty = [&](int s) { return bar_base + 8 * (8 + s); };
    // "stage q holds a finished tile of row half h": one barrier per (stage, half).  A stage serves the two halves in turn
    // (3 is odd), and a parity wait is only sound when the waiter also consumed the barrier's previous phase -- a half
    // that ran ahead would otherwise see the other half's still-pending phase as "mine, completed"
    auto t_full = [&](int q, int h) { return bar_base + 8 * (16 + 2 * q + h); };
    auto t_empty = [&](int q) { return bar_base + 8 * (24 + q); };
    auto a_full = [&](int s) { return bar_base + 8 * (27 + s); };
    auto a_empty = [&](int s) { return bar_base + 8 * (29 + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L.tmem_slot);

    const int warp = __shfl_sync(VQ_FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    constexpr int kEpiWarp0 = 0, kTmaWarp = 8, kMmaWarp = 9, kAllocWarp = 10;
    const int n_row_tiles = (T + kRowsPerCta - 1) / kRowsPerCta;
    const int n_tiles = K / kTileN;
    const int n_groups = n_tiles / kGroupTiles;

    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < kBStagesTs; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 2); }       // both halves' MMAs done
        for (int q = 0; q < kAccStages; ++q) { mbar_init(t_full(q, 0), 1); mbar_init(t_full(q, 1), 1); mbar_init(t_empty(q), 4); }   // 4 warps drain
        for (int s = 0; s < 2; ++s) { mbar_init(a_full(s), 8); mbar_init(a_empty(s), kAccStages); } // 8 epilogue warps wrote / 3 issuers done
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
    pdl_trigger();
    pdl_wait();

    if (warp >= kTmaWarp) {
        reg_dec<kRegsServiceTs>();
        if (warp == kTmaWarp) {
            // ===================== TMA producer: codebook tiles only =====================
            uint32_t b_cnt = 0;
            VQ_INSTR_BEGIN();
            for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x) {
                for (int n = 0; n < n_tiles; ++n, ++b_cnt) {
                    const int s = b_cnt % kBStagesTs;
                    VQ_TIMED_WAIT(1, b_empty(s), ((b_cnt / kBStagesTs) & 1u) ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(b_full(s), kBStageBytes);
                        tma_load_2d(smem_base + L.b + s * kBStageBytes, &tm_b, b_full(s), 0, n * kTileN);
                    }
                    __syncwarp();
                }
            }
            VQ_INSTR_END(0, 2);
        } else {
            // ===================== MMA issuers: one warp per accumulator stage =====================
            // A wait on an mbarrier costs ~90 cycles even when its phase is long complete, so one thread walking
            // t_empty / b_full / issue / commit for every (n-tile, half) pair needs ~350 cycles per pair against the
            // 167 the tensor pipe takes; three issuers (pairs c = q, q + 3, ...; stage q) have 500 each.
            const int q = warp - kMmaWarp;                         // accumulator stage served by this warp
            const uint32_t tmem_base = *tmem_slot;
            const int pairs_per_rt = 2 * n_tiles;
            const int my_row_tiles = ((int)blockIdx.x < n_row_tiles) ? (n_row_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
            int it = 0, c_in = q;                                  // row-tile count and pair index inside the row tile
            uint32_t tile_cnt = (uint32_t)(q >> 1);                // n-tiles of this CTA so far (b ring position)
            uint32_t ph = 0;
            VQ_INSTR_BEGIN();
            if (my_row_tiles > 0) {
                VQ_TIMED_WAIT(0, a_full(0), 0u);
                tc_fence_after();
            }
            while (it < my_row_tiles) {
                const int h = c_in & 1;
                const int sb = tile_cnt % kBStagesTs;
                VQ_TIMED_WAIT(2, b_full(sb), (tile_cnt / kBStagesTs) & 1u);
                VQ_TIMED_WAIT(1, t_empty(q), ph ^ 1u);
                tc_fence_after();
                const uint32_t b_addr = smem_base + L.b + sb * kBStageBytes;
                const uint32_t a_col = tmem_base + (uint32_t)(kACol + (it & 1) * 32 + h * 16);
                VQ_TIMED_BEGIN();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        umma_f16_ts(tmem_base + (uint32_t)(q * kTileN), a_col + (uint32_t)(k * 8), umma_desc(b_addr + k * 32), kIdesc,
                                    (uint32_t)k);
                    umma_commit(t_full(q, h));
                    umma_commit(b_empty(sb));
                }
                __syncwarp();
                VQ_TIMED_END(3);
                ph ^= 1u;
                // next pair of this issuer: c + 3
                const int c_next = c_in + kAccStages;
                tile_cnt += (uint32_t)((c_next >> 1) - (c_in >> 1));
                c_in = c_next;
                if (c_in >= pairs_per_rt) {
                    c_in -= pairs_per_rt;
                    if (elect_one()) umma_commit(a_empty(it & 1));      // this issuer's MMAs of the row tile are all queued
                    __syncwarp();
                    ++it;
                    if (it < my_row_tiles) {
                        VQ_TIMED_WAIT(0, a_full(it & 1), ((uint32_t)(it >> 1)) & 1u);
                        tc_fence_after();
                    }
                }
            }
            if (q == 0) { VQ_INSTR_END(3, 4); }
        }
    } else {
        // ===================== epilogue: 8 warps, one thread per row =====================
        reg_inc<kRegsEpilogueTs>();
        VQ_INSTR_BEGIN();
        const int e = warp - kEpiWarp0;
        const int r_sub = e >> 2;                    // which 128-row half
        const int quarter = warp & 3;                // TMEM lane quarter this warp may access (warp id mod 4)
        const int row_in_cta = r_sub * 128 + quarter * 32 + lane;
        const uint32_t tlane = *tmem_slot + ((uint32_t)(quarter * 32) << 16);
        const uint32_t snap0 = smem_base + L.snap + (uint32_t)row_in_cta * kSnapRow;
        const bool force_exhaustive = codebook_degenerate(cb_info);
        // this half's (n-tile, half) pairs are c = 2 j + r_sub: stage c mod 3; the half meets each stage every third tile,
        // so the phase parity of its barriers flips every three tiles
        uint32_t st = (uint32_t)r_sub, ph = 0, t3 = 0;
        auto advance = [&]() {
            st += 2;
            if (st >= (uint32_t)kAccStages) st -= (uint32_t)kAccStages;
            if (++t3 == 3u) { t3 = 0; ph ^= 1u; }
        };
        // the thread's own fp16 unit row of a row tile (64 bytes; rows past the end read as zero)
        auto load_row = [&](int rt, uint4 (&v)[4]) {
            const int64_t row = (int64_t)rt * kRowsPerCta + row_in_cta;
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = (row < T) ? __ldg(zn16 + row * 4 + i) : make_uint4(0u, 0u, 0u, 0u);
        };
        // ... into the A columns of buffer `buf` (its w-th use): wait until the MMAs of the row tile that used it completed
        auto store_row = [&](int buf, int w, const uint4 (&v)[4]) {
            VQ_TIMED_WAIT(1, a_empty(buf), ((uint32_t)w & 1u) ^ 1u);
            tc_fence_after();
            tmem_st16(tlane + (uint32_t)(kACol + buf * 32 + r_sub * 16), v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full(buf));
        };
        int it = 0;
        uint32_t bufA[64], bufB[64];
        {
            uint4 first[4];
            if ((int)blockIdx.x < n_row_tiles) { load_row(blockIdx.x, first); store_row(0, 0, first); }
        }
        for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x, ++it) {
            uint32_t slot[32];
            int m1 = -32768, m2 = -32768, m3 = -32768, m4 = -32768, m5 = -32768;
            int g1 = 0, g2 = 0, g3 = 0, g4 = 0;
            uint32_t a1 = 0, a2 = 1, a3 = 2, a4 = 3;  // snapshot areas of the best four groups
            const bool has_next = rt + (int)gridDim.x < n_row_tiles;
            uint4 a_next[4];
            if (has_next) load_row(rt + gridDim.x, a_next);     // lands while the first group is folded
            auto group_end = [&](const int g) {
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
            };

            VQ_TIMED_WAIT(0, t_full(st, r_sub), ph);
            tc_fence_after();
            tmem_ld_tile(tlane + st * (uint32_t)kTileN, bufA);
            for (int g = 0; g < n_groups; ++g) {
#pragma unroll
                for (int b = 0; b < kGroupTiles; ++b) {
                    uint32_t (&cur)[64] = (b & 1) ? bufB : bufA;
                    uint32_t (&nxt)[64] = (b & 1) ? bufA : bufB;
                    VQ_TIMED_BEGIN();
                    tmem_ld_wait();                                   // tile b is in registers: its TMEM stage is free
                    VQ_TIMED_END(2);
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(t_empty(st));
                    advance();
                    if (b < kGroupTiles - 1 || g + 1 < n_groups) {
                        VQ_TIMED_WAIT(0, t_full(st, r_sub), ph);
                        tc_fence_after();
                        tmem_ld_tile(tlane + st * (uint32_t)kTileN, nxt);
                    }
                    if (b == 0) {
                        if (g > 0) group_end(g - 1);      // behind the hand-off of this tile's stage
                        if (g == 1 && has_next) store_row((it + 1) & 1, (it + 1) >> 1, a_next);
#pragma unroll
                        for (int j = 0; j < 32; ++j) slot[j] = __vmaxs2(cur[j], cur[j + 32]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) slot[j] = __vimax3_s16x2(slot[j], cur[j], cur[j + 32]);
                    }
                }
            }
            group_end(n_groups - 1);
            if (n_groups == 1 && has_next) store_row((it + 1) & 1, (it + 1) >> 1, a_next);
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
            // decided iff no further group can hold the winner and the survivors fit a record
            const int n_cand = __popc(mask[0]) + __popc(mask[1]) + __popc(mask[2]) + __popc(mask[3]) + __popc(mask[4]) + __popc(mask[5]) +
                               __popc(mask[6]) + __popc(mask[7]);
            const bool decided = thr_ok && (m5 < thr) && ((mask[0] | mask[1]) != 0) && (n_cand <= kRecCells) && !force_exhaustive;
            const int row = rt * kRowsPerCta + row_in_cta;
            const bool in_range = row < T;
            const bool flag = in_range && !decided;
            // verdict record: the surviving cells as 16-bit cell ids (group * 64 + slot), pushed into a 128-bit shift register
            // from the low end; 0xFFFF = none.  An undecided row keeps all-ones (its first field says "listed").
            unsigned long long lo = ~0ull, hi = ~0ull;
            if (decided) {
                const int gv[4] = {g1, g2, g3, g4};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t w = mask[2 * a + h];
                        while (w) {
                            const uint32_t id = (uint32_t)gv[a] * 64u + 32u * h + (uint32_t)(__ffs((int)w) - 1);
                            w &= w - 1;
                            hi = (hi << 16) | (lo >> 48);
                            lo = (lo << 16) | id;
                        }
                    }
            }
            if (in_range) rec[row] = make_int4((int)(uint32_t)lo, (int)(uint32_t)(lo >> 32), (int)(uint32_t)hi, (int)(uint32_t)(hi >> 32));
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
    float* zq; void* idx; int32_t* hist; unsigned long long* seg; int idx_bits;
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
//            matters there, not throughput: each of the first `flagged_cap` listed rows is split over kFlaggedSlices
//            items, an item leaves its (best, second) in `partial`, the last item of a row to finish folds them and
//            finishes the row (`done` holds one zeroed counter per such row); listed rows beyond the cap (degenerate
//            inputs only) are searched by one block each.
// The instruction stream of phase A is what bounds the kernel (the loads are L2 hits and 32 warps per SM hide them), so
// it is kept lean: 16-byte records holding cell ids, 32-bit ordered distance keys, one pass of 8-lane minimum shuffles.
constexpr int kExactThreads = 128;
#ifndef VQ_EXACT_MIN_BLOCKS
#define VQ_EXACT_MIN_BLOCKS 8
#endif
constexpr int kFlaggedSlices = 32;
struct __align__(16) FlaggedPartial {
    unsigned long long best; float second; float pad;
};
// distance -> unsigned key with torch.argmin's order: NaN (wins) -> 0, else sign-corrected float bits
__device__ __forceinline__ uint32_t dist_ord(float d) {
    const uint32_t b = __float_as_uint(d);
    return (d != d) ? 0u : (b ^ ((uint32_t)((int)b >> 31) | 0x80000000u));
}
__device__ __forceinline__ float ord_dist(uint32_t o) {
    return (o == 0u) ? __int_as_float(0x7fc00000) : __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}

// kCellKind: the cell layout of the records (vq_kernels.h) as a compile-time constant - the kernel is bound by its
// instruction stream, and with the default layout the 8 members of a cell are 8 KB apart from ONE base address
// (immediate offsets instead of eight 64-bit address computations behind a runtime switch).
template <int kCellKind>
__global__ void __launch_bounds__(kExactThreads, VQ_EXACT_MIN_BLOCKS)
k_exact_finish16(const int4* __restrict__ rec, const float* __restrict__ zn32, const float* __restrict__ row_sq,
                 const float* __restrict__ en32, const float4* __restrict__ en32c, const float* __restrict__ csq_cell, int T,
                 int K, const int* __restrict__ flagged, const int* __restrict__ n_flagged, int flagged_cap,
                 FlaggedPartial* __restrict__ partial, int* __restrict__ done, FinishOut out, int64_t* __restrict__ stats) {
    constexpr int cell_kind = kCellKind;
    __shared__ unsigned long long s_best[kExactThreads / 32];
    __shared__ float s_second[kExactThreads / 32];
    // Row strides of 9 / 10 float4 put the 4 rows of a warp in different banks: the broadcast LDS.128 of the rescoring
    // (one address per 8-lane group) and the strided LDS.32 of the segment-sum transposition are conflict-free.
    __shared__ __align__(16) float4 s_z[kExactThreads / 32][4][kD / 4 + 1];
    __shared__ __align__(16) float4 s_df[kExactThreads / 32][4][kD / 4 + 2];
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = lane & 7, grp = lane >> 3;
    const float4* zn4 = reinterpret_cast<const float4*>(zn32);
    const float4* en4 = reinterpret_cast<const float4*>(en32);
    // per-thread counters share one register (a thread sees at most a few dozen rows): ties in bits 0-9, rows
    // rescored over several cells in bits 10-19
    unsigned counts = 0, bad = 0;
    constexpr unsigned kTie = 1u, kMulti = 1u << 10;
    long long loss_fx = 0;
    // ---------------- phase A ----------------
    // A warp owns 4 rows per iteration; lane m of a row's 8-lane group owns member m of every cell (whole 128-byte lines
    // of the cell copies) and chunk m of the row in the finish.  The unit rows (512 contiguous bytes) arrive with ONE
    // coalesced 16-byte load per lane and are staged in shared memory for the broadcast reads of the dot products.
    const int groups = gridDim.x * (kExactThreads / 8);
    const uint32_t gmask = 0xFFu << (8 * grp);
    for (int row0 = (blockIdx.x * kExactThreads + threadIdx.x - lane) >> 3; row0 < T; row0 += groups) {
        const int row = row0 + grp;
        const bool in_range = row < T;
        int4 r = make_int4(-1, -1, -1, -1);
        float4 nz = make_float4(0.f, 0.f, 0.f, 0.f);
        float a_sq = 0.f;
        if (in_range) {
            r = __ldg(rec + row);
            nz = __ldg(zn4 + (int64_t)row0 * (kD / 4) + lane);      // row row0 + (lane >> 3), chunk lane & 7
            a_sq = __ldg(row_sq + row);
        }
        // (fetching the next iteration's record / row with cp.async into a second buffer was measured: no gain, the kernel
        // is bound by L1 wavefronts -- 70 % of peak in ncu -- not by the latency of these first loads)
        const bool valid = in_range && ((r.x & r.y & r.z & r.w) != -1);      // all-ones: the filter left the row undecided
        const uint32_t w0i = (uint32_t)r.x, w1i = (uint32_t)r.y, w2i = (uint32_t)r.z, w3i = (uint32_t)r.w;
        // this lane's best over the row's cells: (ordered distance, code), and the ordered distance of its runner-up
        uint32_t bo, so;
        int bc, n_cells;
        auto take = [&](float dist, int code) {
            const uint32_t o = dist_ord(dist);
            const bool wins = (o < bo) || (o == bo && code < bc);
            so = wins ? bo : min(so, o);
            bc = wins ? code : bc;
            bo = wins ? o : bo;
        };
        // Fast pass.  Lane m keeps chunk m of its row in registers and reads chunk m of each of the cell's 8 members (one
        // whole 128-byte line of the row-major unit codes per member and row); the 8 x 8 partial sums are transposed and
        // added with 7 shuffles, after which lane m holds the dot product of member m.  No shared-memory staging, no
        // broadcast reads of the row: a third fewer L1 wavefronts than a full dot product per lane.  The summation
        // order differs from the library's reference chain (one sequential fma chain over d = 0..31, vq_dist_simt.cu), by
        // at most ~1.5e-6 in the distance of unit vectors: rows whose two best candidates are closer than kSafeGap fall
        // through to the chain below, everything else is decided here with the same index the chain would give.
        constexpr float kSafeGapAbs = 1e-5f, kSafeGapRel = 2e-5f;
        {
            bo = so = 0xFFFFFFFFu; bc = 0x7FFFFFFF; n_cells = 0;
            uint32_t w0 = w0i, w1 = w1i, w2 = w2i, w3 = w3i;
            // cells of this group: the record fields that are not 0xFFFF (some filters fill the two halves of a record
            // separately, so empty fields may sit between full ones).  The loop diverges per 8-lane group - a group's
            // lanes share the record - and its shuffles name only the group: groups with fewer cells skip the body.
            int mine = 0;
            if (valid) {
                const uint32_t t0 = ~w0i, t1 = ~w1i, t2 = ~w2i, t3 = ~w3i;      // a half is zero iff its field is empty
                mine = ((t0 & 0xFFFFu) != 0u) + ((t0 >> 16) != 0u) + ((t1 & 0xFFFFu) != 0u) + ((t1 >> 16) != 0u) +
                       ((t2 & 0xFFFFu) != 0u) + ((t2 >> 16) != 0u) + ((t3 & 0xFFFFu) != 0u) + ((t3 >> 16) != 0u);
            }
            n_cells = mine;
            for (int c = 0; c < mine; ++c) {
                // next non-empty field
                while ((w0 & 0xFFFFu) == 0xFFFFu) {
                    w0 = __funnelshift_r(w0, w1, 16); w1 = __funnelshift_r(w1, w2, 16); w2 = __funnelshift_r(w2, w3, 16);
                    w3 = (w3 >> 16) | 0xFFFF0000u;
                }
                const int ci = (int)(w0 & 0xFFFFu);
                w0 = __funnelshift_r(w0, w1, 16); w1 = __funnelshift_r(w1, w2, 16); w2 = __funnelshift_r(w2, w3, 16);
                w3 = (w3 >> 16) | 0xFFFF0000u;
                float p[8];
                const float4* e_base = en4 + (int64_t)tc16_code_of(ci, 0, cell_kind) * (kD / 4) + m;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 e;
                    if constexpr (kCellKind == 1) e = __ldg(e_base + j * (64 * (kD / 4)));      // member j: 64 codes further
                    else e = __ldg(en4 + (int64_t)tc16_code_of(ci, j, cell_kind) * (kD / 4) + m);
                    p[j] = __fmaf_rn(nz.w, e.w, __fmaf_rn(nz.z, e.z, __fmaf_rn(nz.y, e.y, __fmul_rn(nz.x, e.x))));
                }
                const float csq = __ldg(csq_cell + ci * 8 + m);
                float q4[4], q2[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float recv = __shfl_xor_sync(gmask, (m & 4) ? p[i] : p[i + 4], 4);
                    q4[i] = ((m & 4) ? p[i + 4] : p[i]) + recv;
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float recv = __shfl_xor_sync(gmask, (m & 2) ? q4[i] : q4[i + 2], 2);
                    q2[i] = ((m & 2) ? q4[i + 2] : q4[i]) + recv;
                }
                const float recv = __shfl_xor_sync(gmask, (m & 1) ? q2[0] : q2[1], 1);
                const float dot = ((m & 1) ? q2[1] : q2[0]) + recv;
                take(ref_distance(a_sq, csq, dot), tc16_code_of(ci, m, cell_kind));
            }
            __syncwarp();
        }
        uint32_t omin;
        int code;
        float bd, runner;
        auto group_argmin = [&]() {
            // minimum over the 8 lanes of the group, ties to the lowest code
            omin = bo;
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) omin = min(omin, __shfl_xor_sync(VQ_FULL, omin, off));
            code = (bo == omin) ? bc : 0x7FFFFFFF;
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) code = min(code, __shfl_xor_sync(VQ_FULL, code, off));
            bd = ord_dist(omin);
            runner = ord_dist((bc == code) ? so : bo);      // (a lane without a second candidate reports NaN: no candidate)
        };
        group_argmin();
        // a lane vetoes the fast verdict if one of ITS other candidates is not safely behind the best (NaN anywhere: veto)
        const bool has_runner = ((bc == code) ? so : bo) != 0xFFFFFFFFu;
        const bool unsafe = valid && ((bd != bd) || (has_runner && !(runner - bd > kSafeGapAbs + kSafeGapRel * fabsf(bd))));
        bool close = false;
        if (__any_sync(VQ_FULL, unsafe)) {
            // the library's reference chain for the whole warp iteration (a few rows per 10 000 get here): rows staged
            // in shared memory, lane m owns member m of every cell and runs the sequential fma chain over d = 0..31
            __syncwarp();                               // the previous iteration's reads of s_z / s_df are done
            s_z[warp][grp][m] = nz;
            __syncwarp();
            const float4* zs = s_z[warp][grp];
            bo = so = 0xFFFFFFFFu; bc = 0x7FFFFFFF;
            uint32_t w0 = w0i, w1 = w1i, w2 = w2i, w3 = w3i;
            for (int f = 0; valid && f < 8; ++f) {
                const int ci = (int)(w0 & 0xFFFFu);
                w0 = __funnelshift_r(w0, w1, 16); w1 = __funnelshift_r(w1, w2, 16); w2 = __funnelshift_r(w2, w3, 16);
                w3 = (w3 >> 16) | 0xFFFF0000u;
                if (ci != 0xFFFF) take(cell_distance_staged(en32c, csq_cell, ci, m, zs, a_sq), tc16_code_of(ci, m, cell_kind));
            }
            group_argmin();
            // near tie: some other candidate of the row within 1e-6 relative of the best distance
            close = valid && (runner - bd < VQ_NEAR_TIE_REL * fabsf(bd));
        }
        const uint32_t close_ballot = __ballot_sync(VQ_FULL, close);
        // rows of the warp that chose the same code (collapsed codebooks): their histogram count and segment-sum terms
        // are added once, by the first of them -- the L2 serialises reductions per address
        const unsigned same = __match_any_sync(VQ_FULL, valid ? code : -1 - grp);
        const bool dup = __any_sync(VQ_FULL, __popc(same) > 8);        // uniform; false on all but collapsed usage
        const bool lead = valid && (!dup || ((__ffs(same) - 1) >> 3) == grp);
        if (valid && m == 0) {
            store_token(out.idx, row, code, out.idx_bits);
            if (out.hist && lead) atomicAdd(out.hist + code, dup ? __popc(same & 0x01010101u) : 1);
            if (close_ballot & gmask) counts += kTie;
            if (n_cells > 1) counts += kMulti;
        }
        if (out.zq) {                                   // uniform
            float4 df = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) df = finish_chunk(nz, en4, out, row, code, m, loss_fx, bad);
            if (out.seg) {                              // uniform
                __syncwarp();                           // the previous iteration's reads of s_df are done
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
        const int grp16 = threadIdx.x >> 3;
        const int n = *n_flagged;
        const int n_sliced = min(n, flagged_cap);
        const int n_cells = K / kCellCodes;
        const int per_slice = (n_cells + kFlaggedSlices - 1) / kFlaggedSlices;
        const int n_items = n_sliced * kFlaggedSlices + (n - n_sliced);
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const bool sliced = item < n_sliced * kFlaggedSlices;
            const int i = sliced ? item / kFlaggedSlices : n_sliced + (item - n_sliced * kFlaggedSlices);
            const int slice = sliced ? item % kFlaggedSlices : 0;
            const int row = flagged[i];
            __syncthreads();                       // the previous item's shared values are consumed
            if (threadIdx.x < kD / 4) s_z[0][0][threadIdx.x] = __ldg(zn4 + (int64_t)row * (kD / 4) + threadIdx.x);
            __syncthreads();
            const float4* zs = s_z[0][0];
            const float a_sq = __ldg(row_sq + row);
            Top2 top;
            top.init();
            const int c_begin = sliced ? slice * per_slice : 0;
            const int c_end = sliced ? min(n_cells, (slice + 1) * per_slice) : n_cells;
            for (int ci = c_begin + grp16; ci < c_end; ci += kExactThreads / 8) {
                const float dist = cell_distance_staged(en32c, csq_cell, ci, m, zs, a_sq);
                top.add(dist_key(dist, tc16_code_of(ci, m, cell_kind)));
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
                Top2 fin;
                fin.init();
                for (int w = 0; w < kExactThreads / 32; ++w) fin.merge(s_best[w], s_second[w]);
                bool last = true;
                if (sliced) {
                    FlaggedPartial pp;
                    pp.best = fin.best; pp.second = fin.second; pp.pad = 0.f;
                    partial[(int64_t)i * kFlaggedSlices + slice] = pp;
                    __threadfence();
                    last = atomicAdd(done + i, 1) == kFlaggedSlices - 1;
                    if (last) {
                        __threadfence();
                        fin.init();
                        for (int w = 0; w < kFlaggedSlices; ++w) {
                            const FlaggedPartial* q = partial + (int64_t)i * kFlaggedSlices + w;
                            fin.merge(__ldcg(&q->best), __ldcg(&q->second));
                        }
                        done[i] = 0;               // ready for the next call
                    }
                }
                if (last) {
                    const float bd = key_dist(fin.best);
                    const int code = (int)(uint32_t)fin.best;
                    store_token(out.idx, row, code, out.idx_bits);
                    if (out.hist) atomicAdd(out.hist + code, 1);
                    if (fin.second - bd < VQ_NEAR_TIE_REL * fabsf(bd)) counts += kTie;
                    if (out.zq) finish_row_serial(zn4, en4, out, K, row, code, loss_fx, bad);
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

// VQ_TC16_W16=1 selects the 16-epilogue-warp kernel (k_dist_tc16_w16, cell layout 3).  Measured equal to the 8-warp
// kernel (123.0 vs 122.3 us, same box): with the epilogue's work removed altogether the kernel still takes 96 us (the
// MMA / hand-off side), so neither drain is what bounds it; the default stays the kernel with fewer moving parts.
bool tc16_wide_drain() {
    static const bool on = getenv("VQ_TC16_W16") && atoi(getenv("VQ_TC16_W16")) != 0;
    return on;
}

bool tc16_disabled() {
    static const bool disabled = getenv("VQ_TC16_DISABLE") && atoi(getenv("VQ_TC16_DISABLE")) != 0;
    return disabled;
}

bool tc16_supported(int64_t T, int K, int D) {
    const bool disabled = tc16_disabled();
    // code ids travel as 16-bit fields of the rescoring items; the threshold logic needs whole 512-code groups
    return !disabled && D == tc16::kD && K >= tc16::kGroupCols && (K % tc16::kGroupCols) == 0 && K <= 65536 && T >= 256;
}

size_t tc16_workspace_bytes(int64_t T) { return (size_t)(T > 0 ? T : 1) * tc16::kRecordBytes; }

// VQ_TC16_CLUSTER=1: the filter as clusters of two 128-row CTAs (k_dist_tc16<true, 1>; measured slower than the 256-row
// default, see the kernel's header).  Returns how many such clusters the device can hold at once; 0 = use the 256-row
// kernel (the default, a build with two issuers, the wide-drain / rows-in-TMEM variants, or a device that cannot
// co-schedule the pairs).
int tc16_max_clusters() {
    if constexpr (tc16::kIssuers != 4) {
        return 0;
    } else {
        static PerDeviceOnce once;
        static int n_of[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (once.need()) {
            const char* e = getenv("VQ_TC16_CLUSTER");
            const bool alt = (getenv("VQ_TC16_TS") && atoi(getenv("VQ_TC16_TS")) != 0) || tc16_wide_drain();
            int n = 0;
            if (e && e[0] == '1' && !alt) {
                const tc16::SmemLayout L = tc16::smem_layout(1);
                if (cudaFuncSetAttribute(tc16::k_dist_tc16<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)L.total + 1024) == cudaSuccess) {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(2 * sm_count());
                    cfg.blockDim = dim3(tc16::kThreads);
                    cfg.dynamicSmemBytes = L.total + 1024;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at;
                    cfg.numAttrs = 1;
                    if (cudaOccupancyMaxActiveClusters(&n, tc16::k_dist_tc16<true, 1>, &cfg) != cudaSuccess) n = 0;
                }
                cudaGetLastError();
            }
            n_of[dev & 63] = n;
        }
        return n_of[dev & 63];
    }
}

cudaError_t launch_dist_tc16(const CUtensorMap& ma, const CUtensorMap& mb, int T, const __half* zn16, const float* zn32,
                             const float* row_sq, const CodebookView& cb, int* cand, int* flagged, int* n_flagged, int64_t* stats,
                             void* records, cudaStream_t s) {
    // VQ_TC16_TS=1: the variant with the token rows in tensor memory (measured slower, see k_dist_tc16_ts)
    static const bool ss_form = !(getenv("VQ_TC16_TS") && atoi(getenv("VQ_TC16_TS")) != 0);
    int4* rec = static_cast<int4*>(records);
    const int n_row_tiles = (T + tc16::kRowsPerCta - 1) / tc16::kRowsPerCta;
    int grid = n_row_tiles < sm_count() ? n_row_tiles : sm_count();
    cudaError_t e;
    if (cb.cell_kind == 3) {
        const tc16::SmemLayoutW L = tc16::smem_layout_w();
        static PerDeviceOnce once;
        if (once.need()) {
            e = cudaFuncSetAttribute(tc16::k_dist_tc16_w16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total + 1024);
            if (e != cudaSuccess) return e;
        }
        e = launch_pdl(tc16::k_dist_tc16_w16, dim3(grid), dim3(tc16::kThreadsW), L.total + 1024, s, ma, mb, T, cb.K, cb.info,
                       reinterpret_cast<uint2*>(rec), flagged, n_flagged, stats);
    } else if (ss_form && tc16_max_clusters() > 0) {
        // (the caller built the tensor maps for this form: 128-row token boxes, 64-code codebook boxes)
        if constexpr (tc16::kIssuers == 4) {
            const tc16::SmemLayout L = tc16::smem_layout(1);
            const int n_units = ((T + 127) / 128 + 1) / 2;
            const int clusters = n_units < tc16_max_clusters() ? n_units : tc16_max_clusters();
            grid = 2 * clusters;
            e = launch_pdl_cluster(tc16::k_dist_tc16<true, 1>, dim3(grid), dim3(tc16::kThreads), L.total + 1024, s, 2, ma, mb, T, cb.K,
                                   cb.info, rec, cand, flagged, n_flagged, stats);
        } else {
            e = cudaErrorInvalidValue;
        }
    } else if (ss_form) {
        const tc16::SmemLayout L = tc16::smem_layout(2);
        static PerDeviceOnce once;
        if (once.need()) {
            e = cudaFuncSetAttribute(tc16::k_dist_tc16<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total + 1024);
            if (e != cudaSuccess) return e;
        }
        e = launch_pdl(tc16::k_dist_tc16<true, 2>, dim3(grid), dim3(tc16::kThreads), L.total + 1024, s, ma, mb, T, cb.K, cb.info, rec,
                       cand, flagged, n_flagged, stats);
    } else {
        const tc16::SmemLayoutTs L = tc16::smem_layout_ts();
        static PerDeviceOnce once;
        if (once.need()) {
            e = cudaFuncSetAttribute(tc16::k_dist_tc16_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total + 1024);
            if (e != cudaSuccess) return e;
        }
        e = launch_pdl(tc16::k_dist_tc16_ts, dim3(grid), dim3(tc16::kThreadsTs), L.total + 1024, s, mb,
                       reinterpret_cast<const uint4*>(zn16), T, cb.K, cb.info, rec, flagged, n_flagged, stats);
    }
    count_launch();
    tc::instrument_report(s, grid);
    return e != cudaSuccess ? e : cudaGetLastError();
}

// Exact indices and the finish pass behind the filter (see k_exact_finish16).  zq_tok / hist may be null.
// done_counters: kFlaggedCap zeroed ints; partial_ws: kFlaggedCap * 32 * 16 bytes.  Listed rows beyond kFlaggedCap
// (degenerate inputs only) are left in cand[] = -1 for the caller's overflow path.
cudaError_t launch_exact_finish16(const void* records, const float* zn32, const float* row_sq, const CodebookView& cb, int64_t T,
                                  const int* flagged, const int* n_flagged, int* done_counters, void* partial_ws,
                                  float* zq_tok, void* idx_out, int32_t* hist, int64_t* seg_sums, int64_t* stats,
                                  cudaStream_t s, int idx_bits) {
    const int cap = (int)(T < kFlaggedCap ? T : kFlaggedCap);
    const int rows_per_block = tc16::kExactThreads / 8;
    int64_t blocks = (T + rows_per_block - 1) / rows_per_block;
    static const int grid_mult = getenv("VQ_EXACT_GRID_MULT") ? atoi(getenv("VQ_EXACT_GRID_MULT")) : 32;
    const int64_t grid_cap = (int64_t)sm_count() * grid_mult;    // 32 = 8 resident blocks per SM x 4 waves
    if (blocks > grid_cap) blocks = grid_cap;
    tc16::FinishOut out;
    out.zq = zq_tok; out.idx = idx_out; out.idx_bits = idx_bits; out.hist = hist;
    out.seg = zq_tok ? reinterpret_cast<unsigned long long*>(seg_sums) : nullptr;
    auto* kernel = cb.cell_kind == 3 ? tc16::k_exact_finish16<3> : tc16::k_exact_finish16<1>;
    cudaError_t e = launch_pdl(kernel, dim3((unsigned)blocks), dim3(tc16::kExactThreads), 0, s,
                               static_cast<const int4*>(records), zn32, row_sq, cb.en32, reinterpret_cast<const float4*>(cb.en32c),
                               cb.csq_cell, (int)T, cb.K, flagged, n_flagged, cap, static_cast<tc16::FlaggedPartial*>(partial_ws),
                               done_counters, out, stats);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace vq
