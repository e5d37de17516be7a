// Exhaustive fp32 nearest-code search (SIMT).
//
// Replaces, for the rows it is given, the reference's materialised distance matrix + argmin:
//   models/vitvqgan.py:157-161   d = sum(z^2) + sum(e^2) - 2*einsum('bd,nd->bn'); argmin(d, dim=1)
//   models/vqgan.py:157-161
// without ever storing the T x K matrix.  It is (a) the parity baseline, (b) the path for shapes
// the tcgen05 kernel does not take, and (c) the fallback for the few rows the fp16 tensor-core pass
// cannot decide rigorously.  The dot product is a sequential fp32 fma chain over d = 0..D-1 --
// the same chain the rescoring step of the tensor-core path uses, so both paths return identical
// indices.  Ties go to the lowest index and NaN wins, as torch.argmin does.
//
// Tiling: 256 threads compute a 64-row x 128-code tile; each thread owns 4 rows x 8 codes in
// registers and walks D in chunks of 32 staged d-major in shared memory (conflict-free LDS.128).
#include "vq_common.cuh"
#include "vq_kernels.h"
#include "../../include/vq_b200.h"

namespace vq {

constexpr int kTM = 64, kTN = 128, kDK = 32;
constexpr int kScanSplits = 32;          // code ranges per listed row tile
constexpr int64_t kScanSplitCap = 8192;  // listed rows that get the split treatment

struct Best {
    float d1; int i1; float d2;   // best distance, its index, second-best distance
};

__device__ __forceinline__ void best_insert(Best& b, float d, int i) {
    if (argmin_better(d, i, b.d1, b.i1)) { b.d2 = b.d1; b.d1 = d; b.i1 = i; }
    else if (d < b.d2 || d != d) b.d2 = d;
}

// Finish pass of one listed row by one thread (the overflow rows of the D = 32 fallback list: degenerate inputs
// only): idx / hist / z_q / loss partial / segment sums, same expressions as vq_finish.cu.
__device__ __forceinline__ void finish_row_serial(const float* __restrict__ zn32, const float* __restrict__ en32, int D, int K,
                                                  int row, int code, const ListedFinish& fin, long long& loss_fx,
                                                  unsigned long long& bad) {
    store_token(fin.idx, row, code, fin.idx_bits);
    if (fin.hist) atomicAdd(fin.hist + code, 1);
    if (!fin.zq) return;
    unsigned poison = 0;
    for (int d0 = 0; d0 < D; d0 += 4) {
        float df[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a = __ldg(zn32 + (int64_t)row * D + d0 + i), q = __ldg(en32 + (int64_t)code * D + d0 + i);
            df[i] = __fsub_rn(q, a);
            fin.zq[(int64_t)row * D + d0 + i] = __fadd_rn(a, df[i]);
            if (fin.seg) seg_add(fin.seg + (int64_t)code * D + d0 + i, df[i], poison);
        }
        const float p = (df[0] * df[0] + df[1] * df[1]) + (df[2] * df[2] + df[3] * df[3]);
        if (is_finite(p)) loss_fx += to_fixed(p, VQ_LOSS_SHIFT);
        else bad += 1;
    }
    if (poison) atomicAdd(fin.seg + (int64_t)K * D + code, 1ull);
}

__device__ __forceinline__ void best_merge(Best& a, float d1, int i1, float d2) {
    if (argmin_better(d1, i1, a.d1, a.i1)) {
        const float loser = a.d1;
        a.d1 = d1; a.i1 = i1;
        a.d2 = fminf(loser, d2) ;
        if (loser != loser || d2 != d2) a.d2 = __int_as_float(0x7fc00000);
    } else {
        const float cand = d1;
        if (cand != cand || a.d2 != a.d2) a.d2 = __int_as_float(0x7fc00000);
        else a.d2 = fminf(a.d2, cand);
    }
}

// rows == nullptr: every row 0..T-1.  Otherwise the listed rows [row_begin, min(*n_rows_ptr, row_end)).
// partial == nullptr: the block scans all K codes and writes cand[] itself.  Otherwise blockIdx.y selects
// one of gridDim.y code ranges and the (best, index, second) triple goes to partial[split][list slot];
// the last range to finish a row tile folds them (order-independent argmin => deterministic).
__global__ void __launch_bounds__(256) k_scan_exact(const float* __restrict__ zn32, const float* __restrict__ row_sq,
                                                    const float* __restrict__ en32, const float* __restrict__ code_sq,
                                                    int64_t T, int K, int D, const int* __restrict__ rows,
                                                    const int* __restrict__ n_rows_ptr, int64_t row_begin,
                                                    int64_t row_end, int* __restrict__ cand,
                                                    float4* __restrict__ partial, int partial_cap,
                                                    int* __restrict__ tile_done, int64_t* __restrict__ stats,
                                                    ListedFinish fin, int min_rows, int64_t tail_end) {
    __shared__ __align__(16) float zs[kDK][kTM + 4];
    __shared__ __align__(16) float es[kDK][kTN + 4];
    __shared__ int row_id[kTM];
    __shared__ int s_last;

    pdl_trigger();
    pdl_wait();
    const int64_t n_listed = rows ? (int64_t)(*n_rows_ptr) : T;
    if (rows && n_listed <= min_rows) return;       // a short list was handled by the rescoring kernel
    long long loss_fx = 0;
    unsigned long long bad = 0;
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;     // rows ty*4..+3 ; codes tx*4..+3 and 64+tx*4..+3
    // pass 0: rows [row_begin, row_end), split over blockIdx.y when `partial` is given.  pass 1 (tail_end > row_end):
    // the listed rows beyond the split treatment, [row_end, tail_end), unsplit, over all blocks of the grid.
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1 && (tail_end <= row_end || n_listed <= row_end)) break;
        float4* part = pass == 0 ? partial : nullptr;
        const int64_t n_rows = pass == 0 ? (n_listed < row_end ? n_listed : row_end) : (n_listed < tail_end ? n_listed : tail_end);
        const int64_t first = pass == 0 ? row_begin : row_end;
        const int64_t blk = pass == 0 ? (int64_t)blockIdx.x : (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
        const int64_t n_blk = pass == 0 ? (int64_t)gridDim.x : (int64_t)gridDim.x * gridDim.y;
        int k_lo = 0, k_hi = K;
        if (part) {
            const int per = ((K + (int)gridDim.y - 1) / (int)gridDim.y + kTN - 1) / kTN * kTN;
            k_lo = min(K, (int)blockIdx.y * per);
            k_hi = min(K, k_lo + per);
        }
        for (int64_t tile0 = first + blk * kTM; tile0 < n_rows; tile0 += n_blk * kTM) {
            __syncthreads();
            if (tid < kTM) {
                const int64_t i = tile0 + tid;
                row_id[tid] = (i < n_rows) ? (rows ? rows[i] : (int)i) : -1;
            }
            __syncthreads();

            Best best[4];
    #pragma unroll
            for (int r = 0; r < 4; ++r) { best[r].d1 = INFINITY; best[r].i1 = 0x7fffffff; best[r].d2 = INFINITY; }
            float a_sq[4];
    #pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int rid = row_id[ty * 4 + r];
                a_sq[r] = (rid >= 0) ? row_sq[rid] : 0.f;
            }

            for (int k0 = k_lo; k0 < k_hi; k0 += kTN) {
                float acc[4][8];
    #pragma unroll
                for (int r = 0; r < 4; ++r)
    #pragma unroll
                    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

                for (int d0 = 0; d0 < D; d0 += kDK) {
                    __syncthreads();
                    // stage zs[d][row] and es[d][code] (global reads coalesced along d)
                    for (int e = tid; e < kTM * kDK; e += 256) {
                        const int r = e / kDK, d = e % kDK;
                        const int rid = row_id[r];
                        zs[d][r] = (rid >= 0 && d0 + d < D) ? __ldg(zn32 + (int64_t)rid * D + d0 + d) : 0.f;
                    }
                    for (int e = tid; e < kTN * kDK; e += 256) {
                        const int c = e / kDK, d = e % kDK;
                        const int k = k0 + c;
                        es[d][c] = (k < K && d0 + d < D) ? __ldg(en32 + (int64_t)k * D + d0 + d) : 0.f;
                    }
                    __syncthreads();
    #pragma unroll 8
                    for (int d = 0; d < kDK; ++d) {
                        const float4 zv = *reinterpret_cast<const float4*>(&zs[d][ty * 4]);
                        const float4 e0 = *reinterpret_cast<const float4*>(&es[d][tx * 4]);
                        const float4 e1 = *reinterpret_cast<const float4*>(&es[d][64 + tx * 4]);
                        const float zr[4] = {zv.x, zv.y, zv.z, zv.w};
                        const float ec[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    #pragma unroll
                        for (int r = 0; r < 4; ++r)
    #pragma unroll
                            for (int c = 0; c < 8; ++c) acc[r][c] = __fmaf_rn(zr[r], ec[c], acc[r][c]);
                    }
                }
                // distances of this code tile, in increasing code order per thread
    #pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int k = k0 + ((c < 4) ? (tx * 4 + c) : (64 + tx * 4 + c - 4));
                    if (k < K) {
                        const float b_sq = __ldg(code_sq + k);
    #pragma unroll
                        for (int r = 0; r < 4; ++r) best_insert(best[r], ref_distance(a_sq[r], b_sq, acc[r][c]), k);
                    }
                }
            }
            // merge the 16 threads (tx) that share each row: xor-shuffles stay inside 16-lane groups
    #pragma unroll
            for (int r = 0; r < 4; ++r) {
    #pragma unroll
                for (int off = 8; off > 0; off >>= 1) {
                    const float od1 = __shfl_xor_sync(VQ_FULL, best[r].d1, off);
                    const int oi1 = __shfl_xor_sync(VQ_FULL, best[r].i1, off);
                    const float od2 = __shfl_xor_sync(VQ_FULL, best[r].d2, off);
                    best_merge(best[r], od1, oi1, od2);
                }
                const int rid = row_id[ty * 4 + r];
                if (tx == 0 && rid >= 0) {
                    if (part) {
                        const int64_t slot = tile0 + ty * 4 + r;
                        part[(int64_t)blockIdx.y * partial_cap + slot] =
                            make_float4(best[r].d1, __int_as_float(best[r].i1), best[r].d2, 0.f);
                    } else {
                        cand[rid] = best[r].i1 | kCandExactBit;
                        const float gap = best[r].d2 - best[r].d1;
                        if (stats && gap < VQ_NEAR_TIE_REL * fabsf(best[r].d1))
                            atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_NEAR_TIE_ROWS), 1ull);
                        if (fin.idx) finish_row_serial(zn32, en32, D, K, rid, best[r].i1, fin, loss_fx, bad);
                    }
                }
            }
            if (part) {
                // the last code range to finish a row tile folds all ranges: lane = range, shuffle tree.
                // argmin with index tie-break and "second smallest" do not depend on the fold order.
                __threadfence();
                __syncthreads();
                const int tile_idx = (int)(tile0 / kTM);
                if (tid == 0) s_last = (atomicAdd(tile_done + tile_idx, 1) == (int)gridDim.y - 1);
                __syncthreads();
                if (s_last) {
                    __threadfence();
                    const int lane = tid & 31, w = tid >> 5;
                    for (int rr = w * 8; rr < w * 8 + 8; ++rr) {
                        const int64_t slot = tile0 + rr;
                        if (slot >= n_rows) break;
                        Best b;
                        b.d1 = INFINITY; b.i1 = 0x7fffffff; b.d2 = INFINITY;
                        if (lane < (int)gridDim.y) {
                            const float4 p = __ldcg(part + (int64_t)lane * partial_cap + slot);
                            b.d1 = p.x; b.i1 = __float_as_int(p.y); b.d2 = p.z;
                        }
    #pragma unroll
                        for (int off = 16; off > 0; off >>= 1) {
                            const float od1 = __shfl_xor_sync(VQ_FULL, b.d1, off);
                            const int oi1 = __shfl_xor_sync(VQ_FULL, b.i1, off);
                            const float od2 = __shfl_xor_sync(VQ_FULL, b.d2, off);
                            best_merge(b, od1, oi1, od2);
                        }
                        if (lane == 0) {
                            cand[row_id[rr]] = b.i1 | kCandExactBit;
                            if (stats && (b.d2 - b.d1) < VQ_NEAR_TIE_REL * fabsf(b.d1))
                                atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_NEAR_TIE_ROWS), 1ull);
                        }
                    }
                    if (tid == 0) tile_done[tile_idx] = 0;      // ready for the next call
                }
            }
        }
    }
    if (fin.idx && fin.zq && stats) {
        if (loss_fx) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_LOSS_FIXED), (unsigned long long)loss_fx);
        if (bad) atomicAdd(reinterpret_cast<unsigned long long*>(stats + VQ_STAT_NONFINITE), bad);
    }
}

// [tile-done counters (one int per 64-row tile, 1 KiB)] [partials: splits x cap float4]
constexpr size_t kTileDoneBytes = 1024;
static_assert(kScanSplitCap / kTM * sizeof(int) <= kTileDoneBytes && kScanTileCounters * sizeof(int) == kTileDoneBytes,
              "tile counters do not fit");
size_t scan_partial_bytes(int64_t T) {
    const int64_t cap = T < kScanSplitCap ? T : kScanSplitCap;
    const size_t tiled = kTileDoneBytes + sizeof(float4) * (size_t)kScanSplits * (size_t)(cap > 0 ? cap : 1);
    const size_t few = (size_t)kFewFlagged * kFewFlaggedSlices * 16;      // k_rescore_g's partials: rows x slices
    return tiled > few ? tiled : few;
}

cudaError_t launch_scan_exact(const float* zn32, const float* row_sq, const CodebookView& cb, int64_t T,
                              const int* rows, const int* n_rows, int64_t max_rows, int* cand, int64_t* stats,
                              void* partial_ws, cudaStream_t s, int min_rows, int* tile_done_zeroed) {
    const int64_t n = rows ? max_rows : T;
    if (n == 0) return cudaSuccess;
    const int64_t cap_blocks = (int64_t)sm_count() * 16;
    if (!rows || !partial_ws) {
        int64_t blocks = (n + kTM - 1) / kTM;
        if (blocks > cap_blocks) blocks = cap_blocks;
        k_scan_exact<<<(unsigned)blocks, 256, 0, s>>>(zn32, row_sq, cb.en32, cb.code_sq, T, cb.K, cb.D, rows, n_rows, 0,
                                                     n, cand, nullptr, 0, nullptr, stats, ListedFinish{}, min_rows, (int64_t)0);
        count_launch();
        return cudaGetLastError();
    }
    // listed rows: the first `cap` of them with the codebook split over blockIdx.y, the rest unsplit
    const int cap = (int)(n < kScanSplitCap ? n : kScanSplitCap);
    int splits = cb.K / kTN;
    if (splits > kScanSplits) splits = kScanSplits;
    if (splits < 1) splits = 1;
    // tile-done counters: the caller's (cleared with the other per-call counters by the forward's first kernel), else
    // the head of partial_ws, cleared here
    int* tile_done = tile_done_zeroed;
    float4* partial = reinterpret_cast<float4*>(static_cast<char*>(partial_ws) + kTileDoneBytes);
    if (!tile_done) {
        tile_done = static_cast<int*>(partial_ws);
        cudaError_t e = cudaMemsetAsync(tile_done, 0, kTileDoneBytes, s);
        if (e != cudaSuccess) return e;
    }
    // the list length is only known on the device: the grid is one wave of (row tiles x code splits) blocks, which
    // read the row count and leave when the list is short - enough parallelism for the ~0.5 % of rows the D = 256
    // filter leaves undecided.  Listed rows beyond `cap` are scanned unsplit by the same
    // blocks afterwards (pass 1 of the kernel).
    // (two resident blocks per SM is all the kernel's registers allow: more blocks than that only make the usual
    // case - nothing listed, every block reads the count and leaves - cost 6 us instead of 2.5; blocks stride over tiles)
    int tiles = (cap + kTM - 1) / kTM;
    const int resident = (2 * sm_count() + splits - 1) / splits;
    if (tiles > resident) tiles = resident;
    dim3 grid((unsigned)tiles, (unsigned)splits);
    k_scan_exact<<<grid, 256, 0, s>>>(zn32, row_sq, cb.en32, cb.code_sq, T, cb.K, cb.D, rows, n_rows, 0, cap, cand,
                                      partial, cap, tile_done, stats, ListedFinish{}, min_rows, (int64_t)n);
    count_launch();
    return cudaGetLastError();
}

// listed rows [row_begin, *n_rows) without the code split, finished in the same launch (behind the D = 32 per-row
// fallback: its overflow beyond kFlaggedCap; the blocks read the device-side row count and leave when there is none)
cudaError_t launch_scan_listed_tail(const float* zn32, const float* row_sq, const CodebookView& cb, int64_t T,
                                    const int* rows, const int* n_rows, int64_t row_begin, int* cand, int64_t* stats,
                                    const ListedFinish& fin, cudaStream_t s) {
    int64_t blocks = (T - row_begin + kTM - 1) / kTM;
    if (blocks <= 0) return cudaSuccess;
    if (blocks > sm_count()) blocks = sm_count();
    cudaError_t e = launch_pdl(k_scan_exact, dim3((unsigned)blocks), dim3(256), 0, s, zn32, row_sq, cb.en32, cb.code_sq, T, cb.K,
                               cb.D, rows, n_rows, row_begin, T, cand, nullptr, 0, nullptr, stats, fin, 0, (int64_t)0);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace vq
