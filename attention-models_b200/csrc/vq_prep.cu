// Normalisation kernels: codebook preparation and token preparation.
//
// Reference call sites replaced (paths relative to /root/reference):
//   models/vitvqgan.py:152  z = l2_norm(z)                        models/vqgan.py:151-152
//   models/vitvqgan.py:154  embedd_norm = l2_norm(weight)         models/vqgan.py:155
//   models/vitvqgan.py:157  sum(z_flattened**2, dim=1)            models/vqgan.py:157
//   models/vitvqgan.py:158  sum(embedd_norm**2, dim=1)            models/vqgan.py:158
// All HBM-bound: one pass over the rows, coalesced 128 B (or float4) accesses, grid sized in
// multiples of the SM count.
#include <cstdlib>

#include "vq_common.cuh"
#include "vq_kernels.h"

namespace vq {

long long g_kernel_launches = 0;
bool pdl_enabled() {
    static const bool on = getenv("VQ_PDL") && atoi(getenv("VQ_PDL")) != 0;
    return on;
}
static int g_sm_count[64] = {0};
int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && g_sm_count[dev] > 0) return g_sm_count[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    if (dev >= 0 && dev < 64) g_sm_count[dev] = n;
    return n;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t codebook_bytes(int K, int D) {
    size_t n = 0;
    n += align_up(sizeof(float) * (size_t)K * D, 256);    // en32
    n += align_up(sizeof(float) * (size_t)K, 256);        // code_sq
    n += align_up(sizeof(float) * (size_t)K, 256);        // code_denom
    n += align_up(sizeof(__half) * (size_t)K * D, 256);   // en16
    n += 256;                                             // info
    if (has_cell_layout(K, D)) {
        n += align_up(sizeof(float) * (size_t)K * D, 256);   // en32c
        n += align_up(sizeof(float) * (size_t)K, 256);       // csq_cell
    }
    return n;
}

CodebookView codebook_view(void* cb, int K, int D) {
    char* p = static_cast<char*>(cb);
    CodebookView v;
    v.K = K; v.D = D;
    v.en32 = reinterpret_cast<float*>(p);        p += align_up(sizeof(float) * (size_t)K * D, 256);
    v.code_sq = reinterpret_cast<float*>(p);     p += align_up(sizeof(float) * (size_t)K, 256);
    v.code_denom = reinterpret_cast<float*>(p);  p += align_up(sizeof(float) * (size_t)K, 256);
    v.en16 = reinterpret_cast<__half*>(p);       p += align_up(sizeof(__half) * (size_t)K * D, 256);
    v.info = reinterpret_cast<int*>(p);          p += 256;
    v.en32c = nullptr; v.csq_cell = nullptr;
    v.cell_kind = cell_layout_kind(K, D);
    if (has_cell_layout(K, D)) {
        v.en32c = reinterpret_cast<float*>(p);   p += align_up(sizeof(float) * (size_t)K * D, 256);
        v.csq_cell = reinterpret_cast<float*>(p);
    }
    return v;
}

// Codebook preparation: block b leaves its count of non-unit codes in info[b] (grid <= kInfoSlots; block 0 clears
// the unused slots), so the count needs neither atomics nor a zeroed buffer.
__device__ __forceinline__ void write_info(int* __restrict__ info, int n_bad, int vblock, int vgrid) {
    const int total = __syncthreads_count(n_bad != 0);
    if (threadIdx.x == 0) info[vblock] = total;
    if (vblock == 0)
        for (int i = vgrid + threadIdx.x; i < kInfoSlots; i += blockDim.x) info[i] = 0;
}

// One warp per row, kRows rows in flight per warp for memory-level parallelism.
template <int D, bool kIsCodebook>
__global__ void __launch_bounds__(256) k_prep_rows(const float* __restrict__ in, int64_t rows,
                                                   float* __restrict__ unit32, float* __restrict__ sq,
                                                   float* __restrict__ denom, __half* __restrict__ unit16,
                                                   int* __restrict__ info, float4* __restrict__ en32c,
                                                   float* __restrict__ csq_cell, ZeroList zl, int raw) {
    using M = RowMap<D>;
    constexpr int kRows = (M::kPerLane <= 4) ? 4 : 2;
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();
    if (!kIsCodebook) zero_ranges(zl, blockIdx.x, gridDim.x);
    int n_bad = 0;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r0 = warp * kRows; r0 < rows; r0 += n_warps * kRows) {
        float x[kRows][M::kPerLane];
#pragma unroll
        for (int i = 0; i < kRows; ++i)
            if (r0 + i < rows) M::load(in + (r0 + i) * D, lane, x[i]);
#pragma unroll
        for (int i = 0; i < kRows; ++i) {
            const int64_t r = r0 + i;
            if (r >= rows) break;
            // raw: the un-normalised (plain squared-L2) form keeps the rows as they are (x / 1 is exact)
            const float den = raw ? 1.f : norm_denominator(M::template sumsq<true>(x[i]));
#pragma unroll
            for (int j = 0; j < M::kPerLane; ++j) x[i][j] = __fdiv_rn(x[i][j], den);
            const float s2 = M::template sumsq<false>(x[i]);
            if (unit32) M::store(unit32 + r * D, lane, x[i]);
            if (unit16) M::store_half(unit16 + r * D, lane, x[i]);
            if (kIsCodebook && en32c) {
                // generic cell copies (cell_layout_kind 2): RowMap's vector mapping holds float4 #(lane + 32 m) of the row
                if constexpr (M::kVec) {
                    int ci, mem;
                    generic_cell_of((int)r, ci, mem);
#pragma unroll
                    for (int m = 0; m < M::kPerLane / 4; ++m)
                        en32c[((int64_t)ci * (D / 4) + lane + 32 * m) * 8 + mem] =
                            make_float4(x[i][4 * m], x[i][4 * m + 1], x[i][4 * m + 2], x[i][4 * m + 3]);
                    if (lane == 0) csq_cell[ci * 8 + mem] = s2;
                }
            }
            if (lane == 0) {
                if (sq) sq[r] = s2;
                if (denom) denom[r] = den;
                if (kIsCodebook && !(fabsf(s2 - 1.f) < 1e-4f)) ++n_bad;
            }
        }
    }
    if (kIsCodebook) write_info(info, n_bad, blockIdx.x, gridDim.x);
}

// D < 128 (ATen's non-vectorised schedule): D/4 lanes share a row, one float4 each, 32/(D/4) rows per
// warp instruction.  ATen gives element e to lane e % W (W = min(D, 32)), folds the D/W elements of a lane
// first and then runs the shuffle-down tree with element offsets W/2 ... 1.  With 4 consecutive elements
// per thread those offsets become: thread offsets (D/4)/2 ... 1 (which include the per-lane fold when
// D = 64), then the two in-thread steps (x0+x2, x1+x3) and their sum -- bit-identical to ATen.
// The body works on a virtual (block, grid) so that one launch can prepare the codebook on its first blocks and the
// token rows on the others (k_prep_rows_fused).
template <int D, bool kIsCodebook>
__device__ __forceinline__ void prep_rows_small_body(const float4* __restrict__ in, int64_t rows,
                                                     float4* __restrict__ unit32, float* __restrict__ sq,
                                                     float* __restrict__ denom, uint2* __restrict__ unit16,
                                                     int* __restrict__ info, float4* __restrict__ en32c,
                                                     float* __restrict__ csq_cell, int cell_kind, int vblock, int vgrid,
                                                     int raw = 0) {
    static_assert(D == 16 || D == 32 || D == 64, "small-row prep covers D < 128");
    int n_bad = 0;
    constexpr int kLpr = D / 4;
    constexpr int kRpw = 32 / kLpr;
    constexpr int kUnroll = 4;
    const int lane = threadIdx.x & 31;
    const int sub = lane % kLpr, grp = lane / kLpr;
    const int64_t warp = (int64_t)vblock * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)vgrid * (blockDim.x >> 5);
    auto tree = [&](float4 v) {
#pragma unroll
        for (int off = kLpr >> 1; off > 0; off >>= 1) {
            v.x = __fadd_rn(v.x, __shfl_down_sync(VQ_FULL, v.x, off));
            v.y = __fadd_rn(v.y, __shfl_down_sync(VQ_FULL, v.y, off));
            v.z = __fadd_rn(v.z, __shfl_down_sync(VQ_FULL, v.z, off));
            v.w = __fadd_rn(v.w, __shfl_down_sync(VQ_FULL, v.w, off));
        }
        const float total = __fadd_rn(__fadd_rn(v.x, v.z), __fadd_rn(v.y, v.w));
        return __shfl_sync(VQ_FULL, total, grp * kLpr);
    };
    for (int64_t r0 = warp * kRpw * kUnroll; r0 < rows; r0 += n_warps * kRpw * kUnroll) {
        float4 x[kUnroll];
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) {
            const int64_t r = r0 + i * kRpw + grp;
            x[i] = (r < rows) ? __ldg(in + r * kLpr + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) {
            const int64_t r = r0 + i * kRpw + grp;
            float4 v = x[i];
            const float den = raw ? 1.f : norm_denominator(
                tree(make_float4(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y), __fmul_rn(v.z, v.z), __fmul_rn(v.w, v.w))));
            v.x = __fdiv_rn(v.x, den); v.y = __fdiv_rn(v.y, den); v.z = __fdiv_rn(v.z, den); v.w = __fdiv_rn(v.w, den);
            const float s2 =
                tree(make_float4(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y), __fmul_rn(v.z, v.z), __fmul_rn(v.w, v.w)));
            if (r < rows) {
                if (unit32) unit32[r * kLpr + sub] = v;
                if (unit16) {
                    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
                    uint2 u;
                    u.x = *reinterpret_cast<const uint32_t*>(&a);
                    u.y = *reinterpret_cast<const uint32_t*>(&b);
                    unit16[r * kLpr + sub] = u;
                }
                if (sub == 0) {
                    if (sq) sq[r] = s2;
                    if (denom) denom[r] = den;
                    if (kIsCodebook && !(fabsf(s2 - 1.f) < 1e-4f)) ++n_bad;
                }
                if (kIsCodebook && en32c) {
                    // cell copies of the unit codes (see cell_layout_kind): chunk `sub` of code r
                    const int code = (int)r;
                    int ci, m;
                    if (cell_kind == 1 || cell_kind == 3) tc16_cell_of(code, cell_kind, ci, m);
                    else generic_cell_of(code, ci, m);
                    en32c[((int64_t)ci * kLpr + sub) * 8 + m] = v;
                    if (sub == 0) csq_cell[ci * 8 + m] = s2;
                }
            }
        }
    }
    if (kIsCodebook) write_info(info, n_bad, vblock, vgrid);
}

template <int D, bool kIsCodebook>
__global__ void __launch_bounds__(256) k_prep_rows_small(const float4* __restrict__ in, int64_t rows,
                                                         float4* __restrict__ unit32, float* __restrict__ sq,
                                                         float* __restrict__ denom, uint2* __restrict__ unit16,
                                                         int* __restrict__ info, float4* __restrict__ en32c,
                                                         float* __restrict__ csq_cell, int cell_kind, ZeroList zl, int raw) {
    pdl_trigger();
    pdl_wait();
    if (!kIsCodebook) zero_ranges(zl, blockIdx.x, gridDim.x);
    prep_rows_small_body<D, kIsCodebook>(in, rows, unit32, sq, denom, unit16, info, en32c, csq_cell, cell_kind, blockIdx.x,
                                         gridDim.x, raw);
}

// codebook rows on blocks [0, cb_blocks), token rows on the rest: the two preparations of a training step in one launch
struct PrepRowsArgs {
    const float4* in; int64_t rows; float4* unit32; float* sq; float* denom; uint2* unit16;
};
template <int D>
__global__ void __launch_bounds__(256) k_prep_rows_fused(PrepRowsArgs cbk, int* __restrict__ info, float4* __restrict__ en32c,
                                                         float* __restrict__ csq_cell, int cell_kind, int cb_blocks,
                                                         PrepRowsArgs tok, ZeroList zl) {
    pdl_trigger();
    pdl_wait();
    if ((int)blockIdx.x < cb_blocks) {
        prep_rows_small_body<D, true>(cbk.in, cbk.rows, cbk.unit32, cbk.sq, cbk.denom, cbk.unit16, info, en32c, csq_cell,
                                      cell_kind, blockIdx.x, cb_blocks);
    } else {
        const int vblock = blockIdx.x - cb_blocks, vgrid = gridDim.x - cb_blocks;
        zero_ranges(zl, vblock, vgrid);
        prep_rows_small_body<D, false>(tok.in, tok.rows, tok.unit32, tok.sq, tok.denom, tok.unit16, nullptr, nullptr, nullptr, 0,
                                       vblock, vgrid);
    }
}

template <int D, bool kIsCodebook>
static cudaError_t prep_rows(const float* in, int64_t rows, float* unit32, float* sq, float* denom, __half* unit16,
                             int* info, float* en32c, float* csq_cell, int cell_kind, const ZeroList& zl, cudaStream_t s,
                             int raw = 0) {
    if (rows == 0 && kIsCodebook) return cudaSuccess;
    const int64_t cap = kIsCodebook ? kInfoSlots : (int64_t)sm_count() * 8;
    if constexpr (D < 128) {
        constexpr int rows_per_block = 8 * (32 / (D / 4)) * 4;
        int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        cudaError_t e = launch_pdl(k_prep_rows_small<D, kIsCodebook>, dim3((unsigned)blocks), dim3(256), 0, s,
                                   reinterpret_cast<const float4*>(in), rows, reinterpret_cast<float4*>(unit32), sq, denom,
                                   reinterpret_cast<uint2*>(unit16), info, reinterpret_cast<float4*>(en32c), csq_cell, cell_kind, zl, raw);
        if (e != cudaSuccess) return e;
    } else {
        constexpr int kRows = (RowMap<D>::kPerLane <= 4) ? 4 : 2;
        const int warps_per_block = 8;
        int64_t blocks = (rows + (int64_t)warps_per_block * kRows - 1) / (warps_per_block * kRows);
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        cudaError_t e = launch_pdl(k_prep_rows<D, kIsCodebook>, dim3((unsigned)blocks), dim3(warps_per_block * 32), 0, s, in,
                                   rows, unit32, sq, denom, unit16, info, reinterpret_cast<float4*>(en32c), csq_cell, zl, raw);
        if (e != cudaSuccess) return e;
    }
    count_launch();
    return cudaGetLastError();
}

static cudaError_t prep_codebook_rows(const float* weight, const CodebookView& cb, cudaStream_t s, int raw) {
    const ZeroList none = {};
    VQ_DISPATCH_D(cb.D, return (prep_rows<kD, true>(weight, cb.K, cb.en32, cb.code_sq, cb.code_denom, cb.en16,
                                                     cb.info, cb.en32c, cb.csq_cell, cb.cell_kind, none, s, raw)));
    return cudaSuccess;
}

// One launch: unit codes (fp32 + fp16), squared norms, row norms, the degenerate-code counts and, at D = 32, the
// cell copies the exact rescoring reads.
// raw: the plain squared-L2 form -- en32 = the weights themselves, code_sq = sum(E^2), code_denom = 1 (the info block then
// reports every code as "not unit": the tensor-core filters, whose bounds assume unit rows, stand aside)
cudaError_t launch_prep_codebook(const float* weight, const CodebookView& cb, cudaStream_t s, bool raw) {
    return prep_codebook_rows(weight, cb, s, raw ? 1 : 0);
}

cudaError_t launch_prep_tokens(const float* z, int64_t T, int D, float* zn32, float* row_sq, float* denom,
                               __half* zn16, const ZeroList& zl, cudaStream_t s, bool raw) {
    VQ_DISPATCH_D(D, return (prep_rows<kD, false>(z, T, zn32, row_sq, denom, zn16, nullptr, nullptr, nullptr, 0, zl, s, raw ? 1 : 0)));
    return cudaSuccess;
}

// Codebook and token preparation in one launch (token-major rows, D < 128); false if the shape needs two launches.
bool prep_fusable(int D) { return D < 128; }
cudaError_t launch_prep_fused(const float* weight, const CodebookView& cb, const float* z, int64_t T, float* zn32, float* row_sq,
                              float* denom, __half* zn16, const ZeroList& zl, cudaStream_t s) {
    const int D = cb.D;
    const int rows_per_block = 8 * (32 / (D / 4)) * 4;
    int cb_blocks = (cb.K + rows_per_block - 1) / rows_per_block;
    if (cb_blocks > kInfoSlots) cb_blocks = kInfoSlots;
    int64_t tok_blocks = (T + rows_per_block - 1) / rows_per_block;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (tok_blocks > cap) tok_blocks = cap;
    if (tok_blocks < 1) tok_blocks = 1;
    PrepRowsArgs c{reinterpret_cast<const float4*>(weight), cb.K, reinterpret_cast<float4*>(cb.en32), cb.code_sq, cb.code_denom,
                   reinterpret_cast<uint2*>(cb.en16)};
    PrepRowsArgs t{reinterpret_cast<const float4*>(z), T, reinterpret_cast<float4*>(zn32), row_sq, denom,
                   reinterpret_cast<uint2*>(zn16)};
    cudaError_t e = cudaErrorInvalidValue;
    switch (D) {
        case 16: e = launch_pdl(k_prep_rows_fused<16>, dim3((unsigned)(cb_blocks + tok_blocks)), dim3(256), 0, s, c, cb.info,
                                reinterpret_cast<float4*>(cb.en32c), cb.csq_cell, cb.cell_kind, cb_blocks, t, zl); break;
        case 32: e = launch_pdl(k_prep_rows_fused<32>, dim3((unsigned)(cb_blocks + tok_blocks)), dim3(256), 0, s, c, cb.info,
                                reinterpret_cast<float4*>(cb.en32c), cb.csq_cell, cb.cell_kind, cb_blocks, t, zl); break;
        case 64: e = launch_pdl(k_prep_rows_fused<64>, dim3((unsigned)(cb_blocks + tok_blocks)), dim3(256), 0, s, c, cb.info,
                                reinterpret_cast<float4*>(cb.en32c), cb.csq_cell, cb.cell_kind, cb_blocks, t, zl); break;
        default: break;
    }
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_zero_ranges(ZeroList zl) { zero_ranges(zl, blockIdx.x, gridDim.x); }

// the same zeroing as a launch of its own (layouts whose first kernel is not a token-major prep)
cudaError_t launch_zero_ranges(const ZeroList& zl, cudaStream_t s) {
    k_zero_ranges<<<sm_count() * 2, 256, 0, s>>>(zl);
    count_launch();
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_fill_ones(float* __restrict__ p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 1.f;
}
cudaError_t launch_fill_ones(float* p, int64_t n, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    k_fill_ones<<<(unsigned)blocks, 256, 0, s>>>(p, n);
    count_launch();
    return cudaGetLastError();
}

// row_sq only, from already-normalised contiguous rows
template <int D>
__global__ void __launch_bounds__(256) k_row_sumsq(const float* __restrict__ zn, int64_t rows, float* __restrict__ sq) {
    using M = RowMap<D>;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < rows; r += n_warps) {
        float x[M::kPerLane];
        M::load(zn + r * D, lane, x);
        const float s2 = M::template sumsq<false>(x);
        if (lane == 0) sq[r] = s2;
    }
}

cudaError_t launch_row_sumsq(const float* zn32, int64_t T, int D, float* row_sq, cudaStream_t s) {
    if (T == 0) return cudaSuccess;
    int64_t blocks = (T + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    VQ_DISPATCH_D(D, (k_row_sumsq<kD><<<(unsigned)blocks, 256, 0, s>>>(zn32, T, row_sq)));
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// NCHW view normalised over the channel dim (reference models/vqgan.py:151-152).  ATen reduces a
// strided dim with block (32, 4): lane x owns 4 consecutive hw positions (float4), threadIdx.y
// strides the channels; thread y accumulates channels y + 4i + 16m in accumulator i (fma chain),
// combines ((a0+a1)+a2)+a3, then a shared-memory tree over y (offsets 2, 1).  With D < 64 ATen
// does not split the channels across y: one thread walks all channels, accumulator i <- c = i+4m.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128) k_norm_nchw(const float* __restrict__ z, int64_t T, int64_t hw,
                                                   float* __restrict__ denom, ZeroList zl) {
    zero_ranges(zl, blockIdx.x, gridDim.x);           // first kernel of an NCHW forward: clears what the call accumulates into
    constexpr int kSplit = (D >= 64) ? 4 : 1;
    __shared__ float4 part[4][32];
    const int x = threadIdx.x, y = threadIdx.y;
    const int64_t t = ((int64_t)blockIdx.x * 32 + x) * 4;          // first of this thread's 4 tokens
    float4 total = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < T) {
        const int64_t b = t / hw, p = t % hw;                      // hw % 4 == 0: the 4 tokens share b
        const float* base = z + (b * D) * hw + p;
        float4 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int start = (kSplit == 4) ? y : 0;
        for (int c0 = start; c0 < D; c0 += 4 * kSplit) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(c0 + i * kSplit) * hw));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i].x = __fmaf_rn(v[i].x, v[i].x, acc[i].x);
                acc[i].y = __fmaf_rn(v[i].y, v[i].y, acc[i].y);
                acc[i].z = __fmaf_rn(v[i].z, v[i].z, acc[i].z);
                acc[i].w = __fmaf_rn(v[i].w, v[i].w, acc[i].w);
            }
        }
        total.x = __fadd_rn(__fadd_rn(__fadd_rn(acc[0].x, acc[1].x), acc[2].x), acc[3].x);
        total.y = __fadd_rn(__fadd_rn(__fadd_rn(acc[0].y, acc[1].y), acc[2].y), acc[3].y);
        total.z = __fadd_rn(__fadd_rn(__fadd_rn(acc[0].z, acc[1].z), acc[2].z), acc[3].z);
        total.w = __fadd_rn(__fadd_rn(__fadd_rn(acc[0].w, acc[1].w), acc[2].w), acc[3].w);
    }
    if (kSplit == 4) {
        part[y][x] = total;
        __syncthreads();
        if (y < 2) {
            const float4 o = part[y + 2][x];
            total.x = __fadd_rn(total.x, o.x); total.y = __fadd_rn(total.y, o.y);
            total.z = __fadd_rn(total.z, o.z); total.w = __fadd_rn(total.w, o.w);
            part[y][x] = total;
        }
        __syncthreads();
        if (y == 0) {
            const float4 o = part[1][x];
            total.x = __fadd_rn(total.x, o.x); total.y = __fadd_rn(total.y, o.y);
            total.z = __fadd_rn(total.z, o.z); total.w = __fadd_rn(total.w, o.w);
        }
    }
    if (y == 0 && t < T) {
        float4 d;
        d.x = norm_denominator(total.x); d.y = norm_denominator(total.y);
        d.z = norm_denominator(total.z); d.w = norm_denominator(total.w);
        *reinterpret_cast<float4*>(denom + t) = d;
    }
}

// Any other ATen schedule (few tokens, hw % 4 != 0): one thread per token walks the S channel
// stripes in turn (stripe y: accumulator i <- channels y + S*(i + 4m)) and folds them with the same
// tree (offsets S/2 ... 1).  Only small or odd-shaped inputs come here, so simplicity wins.
__global__ void __launch_bounds__(128) k_norm_nchw_generic(const float* __restrict__ z, int64_t T, int64_t hw, int D,
                                                           int S, float* __restrict__ denom, ZeroList zl) {
    zero_ranges(zl, blockIdx.x, gridDim.x);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int64_t b = t / hw, p = t % hw;
    const float* base = z + (b * D) * hw + p;
    float part[128];
    for (int y = 0; y < S; ++y) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int c = y;
        for (; c + 3 * S < D; c += 4 * S)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float v = __ldg(base + (int64_t)(c + i * S) * hw);
                acc[i] = __fmaf_rn(v, v, acc[i]);
            }
        for (int i = 0; i < 4 && c < D; ++i, c += S) {
            const float v = __ldg(base + (int64_t)c * hw);
            acc[i] = __fmaf_rn(v, v, acc[i]);
        }
        part[y] = __fadd_rn(__fadd_rn(__fadd_rn(acc[0], acc[1]), acc[2]), acc[3]);
    }
    for (int off = S >> 1; off > 0; off >>= 1)
        for (int y = 0; y < off; ++y) part[y] = __fadd_rn(part[y], part[y + off]);
    denom[t] = norm_denominator(part[0]);
}

// ATen's launch shape for a reduction over a non-fastest dim (Reduce.cuh:1034-1180, :99-108):
// returns S, the number of channel stripes combined by block_y_reduce (1 = no split).
static int aten_strided_stripes(int64_t T, int64_t hw, int D) {
    const int vec = (hw % 4 == 0) ? 4 : ((hw % 2 == 0) ? 2 : 1);
    const int64_t max_threads = 512 / vec;
    auto last_pow2 = [](int64_t n) { int64_t p = 1; while (p * 2 <= n) p *= 2; return p; };
    const int64_t dim0 = T / vec > 0 ? T / vec : 1;
    const int64_t dim0_pow2 = dim0 < max_threads ? last_pow2(dim0) : max_threads;
    const int64_t dim1_pow2 = D < max_threads ? last_pow2(D) : max_threads;
    int64_t bw = dim0_pow2 < 32 ? dim0_pow2 : 32;
    const int64_t bh = dim1_pow2 < max_threads / bw ? dim1_pow2 : max_threads / bw;
    const int64_t thresh = bh * 16 < 256 ? bh * 16 : 256;
    return (D >= thresh) ? (int)bh : 1;
}

cudaError_t launch_norm_nchw(const float* z, int64_t T, int64_t hw, int D, float* denom, const ZeroList& zl, cudaStream_t s) {
    if (T == 0) return cudaSuccess;
    const int S = aten_strided_stripes(T, hw, D);
    const bool fast = (hw % 4 == 0) && ((D >= 64 && S == 4) || (D < 64 && S == 1));
    if (fast) {
        const int64_t blocks = (T / 4 + 31) / 32;
        VQ_DISPATCH_D(D, (k_norm_nchw<kD><<<(unsigned)blocks, dim3(32, 4), 0, s>>>(z, T, hw, denom, zl)));
    } else {
        if (S > 128) return cudaErrorInvalidValue;
        k_norm_nchw_generic<<<(unsigned)((T + 127) / 128), 128, 0, s>>>(z, T, hw, D, S, denom, zl);
    }
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// The NCHW prep in ONE launch (it was three: k_norm_nchw, k_nchw_to_tok, k_row_sumsq = 29 us of a 190 us cfg 2 encode,
// the input read twice and the unit rows written and read back).  A CTA of 128 threads owns 32 tokens x D channels:
// thread (y = warp, i = lane >> 3, tq = lane & 7) is k_norm_nchw's accumulator i of channel stripe y for token quad tq -
// the same fma chain over channels y + 4i + 16m, the same ((a0+a1)+a2)+a3 and (s0+s2)+(s1+s3) - and keeps the D/16
// float4 it loaded in registers.  It then divides them by the denominators, transposes through shared memory, and each
// warp writes 8 token rows: unit rows fp32 + fp16 and, from the row as RowMap holds it, row_sq exactly as k_row_sumsq.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128) k_prep_nchw_fused(const float* __restrict__ z, int64_t T, int64_t hw,
                                                         float* __restrict__ denom, float* __restrict__ zn32,
                                                         __half* __restrict__ zn16, float* __restrict__ row_sq, ZeroList zl,
                                                         int raw) {
    using M = RowMap<D>;
    constexpr int kM = D / 16;                  // channels per accumulator chain
    constexpr int kStride = D + 4;              // floats per staged token row (16-byte aligned rows)
    extern __shared__ __align__(16) float s_tile[];          // [32][kStride]
    __shared__ float4 part[4][8];
    zero_ranges(zl, blockIdx.x, gridDim.x);
    const int lane = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int tq = lane & 7, i = lane >> 3;
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    const int64_t t = t0 + 4 * tq;                           // first of this thread's 4 tokens (hw % 4 == 0: same image)
    float4 v[kM];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < T) {
        const int64_t b = t / hw, p = t % hw;
        const float* base = z + (b * D) * hw + p;
#pragma unroll
        for (int m = 0; m < kM; ++m) v[m] = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(y + 4 * i + 16 * m) * hw));
#pragma unroll
        for (int m = 0; m < kM; ++m) {
            acc.x = __fmaf_rn(v[m].x, v[m].x, acc.x);
            acc.y = __fmaf_rn(v[m].y, v[m].y, acc.y);
            acc.z = __fmaf_rn(v[m].z, v[m].z, acc.z);
            acc.w = __fmaf_rn(v[m].w, v[m].w, acc.w);
        }
    } else {
#pragma unroll
        for (int m = 0; m < kM; ++m) v[m] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 dn = make_float4(1.f, 1.f, 1.f, 1.f);
    if (!raw) {
        // stripe sum ((a0 + a1) + a2) + a3 on every lane of the token quad
        float4 sy;
        {
            float4 a[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                a[k].x = __shfl_sync(VQ_FULL, acc.x, tq + 8 * k);
                a[k].y = __shfl_sync(VQ_FULL, acc.y, tq + 8 * k);
                a[k].z = __shfl_sync(VQ_FULL, acc.z, tq + 8 * k);
                a[k].w = __shfl_sync(VQ_FULL, acc.w, tq + 8 * k);
            }
            sy.x = __fadd_rn(__fadd_rn(__fadd_rn(a[0].x, a[1].x), a[2].x), a[3].x);
            sy.y = __fadd_rn(__fadd_rn(__fadd_rn(a[0].y, a[1].y), a[2].y), a[3].y);
            sy.z = __fadd_rn(__fadd_rn(__fadd_rn(a[0].z, a[1].z), a[2].z), a[3].z);
            sy.w = __fadd_rn(__fadd_rn(__fadd_rn(a[0].w, a[1].w), a[2].w), a[3].w);
        }
        if (i == 0) part[y][tq] = sy;
        __syncthreads();
        const float4 s0 = part[0][tq], s1 = part[1][tq], s2 = part[2][tq], s3 = part[3][tq];
        float4 tot;
        tot.x = __fadd_rn(__fadd_rn(s0.x, s2.x), __fadd_rn(s1.x, s3.x));
        tot.y = __fadd_rn(__fadd_rn(s0.y, s2.y), __fadd_rn(s1.y, s3.y));
        tot.z = __fadd_rn(__fadd_rn(s0.z, s2.z), __fadd_rn(s1.z, s3.z));
        tot.w = __fadd_rn(__fadd_rn(s0.w, s2.w), __fadd_rn(s1.w, s3.w));
        dn.x = norm_denominator(tot.x); dn.y = norm_denominator(tot.y);
        dn.z = norm_denominator(tot.z); dn.w = norm_denominator(tot.w);
    }
    if (y == 0 && i == 0 && t < T) *reinterpret_cast<float4*>(denom + t) = dn;
    // transpose: token 4 tq + c, channel y + 4 i + 16 m
    {
        float* col = s_tile + (4 * tq) * kStride + y + 4 * i;
#pragma unroll
        for (int m = 0; m < kM; ++m) {
            col[0 * kStride + 16 * m] = raw ? v[m].x : __fdiv_rn(v[m].x, dn.x);
            col[1 * kStride + 16 * m] = raw ? v[m].y : __fdiv_rn(v[m].y, dn.y);
            col[2 * kStride + 16 * m] = raw ? v[m].z : __fdiv_rn(v[m].z, dn.z);
            col[3 * kStride + 16 * m] = raw ? v[m].w : __fdiv_rn(v[m].w, dn.w);
        }
    }
    __syncthreads();
    for (int rr = 0; rr < 8; ++rr) {
        const int r = 8 * y + rr;
        const int64_t row = t0 + r;
        if (row >= T) break;
        const float* src = s_tile + r * kStride;
        float x[M::kPerLane];
        if constexpr (M::kVec) {
#pragma unroll
            for (int m = 0; m < M::kPerLane / 4; ++m) {
                const float4 q = *reinterpret_cast<const float4*>(src + (lane + 32 * m) * 4);
                x[4 * m + 0] = q.x; x[4 * m + 1] = q.y; x[4 * m + 2] = q.z; x[4 * m + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < M::kPerLane; ++j) x[j] = src[lane + M::kWidth * j];
        }
        M::store(zn32 + row * D, lane, x);
        if (zn16) M::store_half(zn16 + row * D, lane, x);
        const float s2 = M::template sumsq<false>(x);
        if (lane == 0) row_sq[row] = s2;
    }
}

bool prep_nchw_fused_supported(int64_t T, int64_t hw, int D) {
    static const bool off = [] { const char* e = getenv("VQ_PREP_NCHW_FUSED"); return e && e[0] == '0'; }();
    return !off && T > 0 && hw % 4 == 0 && (D == 64 || D == 128 || D == 256) && aten_strided_stripes(T, hw, D) == 4;
}

cudaError_t launch_prep_nchw_fused(const float* z, int64_t T, int64_t hw, int D, float* denom, float* zn32, __half* zn16,
                                   float* row_sq, const ZeroList& zl, bool raw, cudaStream_t s) {
    const unsigned blocks = (unsigned)((T + 31) / 32);
    const size_t smem = (size_t)32 * (D + 4) * sizeof(float);
#define VQ_PREP_NCHW_CASE(kD)                                                                                           \
    case kD: {                                                                                                          \
        static PerDeviceOnce once;                                                                                      \
        if (once.need()) {                                                                                              \
            cudaError_t e = cudaFuncSetAttribute(k_prep_nchw_fused<kD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                             \
        }                                                                                                               \
        k_prep_nchw_fused<kD><<<blocks, 128, smem, s>>>(z, T, hw, denom, zn32, zn16, row_sq, zl, raw ? 1 : 0);           \
        break;                                                                                                          \
    }
    switch (D) {
        VQ_PREP_NCHW_CASE(64)
        VQ_PREP_NCHW_CASE(128)
        VQ_PREP_NCHW_CASE(256)
        default: return cudaErrorInvalidValue;
    }
#undef VQ_PREP_NCHW_CASE
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Layout changes between (b, D, hw) and (T, D): 32 x 32 tiles through padded shared memory, both
// sides coalesced.  nchw_to_tok optionally divides by denom[t] (true division, as F.normalize).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_nchw_to_tok(const float* __restrict__ in, int64_t T, int64_t hw, int D,
                                                     const float* __restrict__ denom, float* __restrict__ out32,
                                                     __half* __restrict__ out16) {
    __shared__ float tile[32][33];
    const int x = threadIdx.x, y = threadIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    {
        const int64_t t = t0 + x;
        if (t < T) {
            const int64_t b = t / hw, p = t % hw;
            const float dn = denom ? denom[t] : 1.f;
            for (int cc = y; cc < 32; cc += 8) {
                const int c = c0 + cc;
                if (c < D) {
                    const float v = __ldg(in + (b * D + c) * hw + p);
                    tile[cc][x] = denom ? __fdiv_rn(v, dn) : v;
                }
            }
        }
    }
    __syncthreads();
    for (int tt = y; tt < 32; tt += 8) {
        const int64_t t = t0 + tt;
        const int c = c0 + x;
        if (t < T && c < D) {
            const float v = tile[x][tt];
            if (out32) out32[t * D + c] = v;
            if (out16) out16[t * D + c] = __float2half_rn(v);
        }
    }
}

__global__ void __launch_bounds__(256) k_tok_to_nchw(const float* __restrict__ in, int64_t T, int64_t hw, int D,
                                                     float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int x = threadIdx.x, y = threadIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int tt = y; tt < 32; tt += 8) {
        const int64_t t = t0 + tt;
        const int c = c0 + x;
        if (t < T && c < D) tile[tt][x] = __ldg(in + t * D + c);
    }
    __syncthreads();
    const int64_t t = t0 + x;
    if (t < T) {
        const int64_t b = t / hw, p = t % hw;
        for (int cc = y; cc < 32; cc += 8) {
            const int c = c0 + cc;
            if (c < D) out[(b * D + c) * hw + p] = tile[x][cc];
        }
    }
}

cudaError_t launch_nchw_to_tok(const float* in, int64_t T, int64_t hw, int D, const float* denom, float* out32,
                               __half* out16, cudaStream_t s) {
    if (T == 0) return cudaSuccess;
    dim3 grid((unsigned)((T + 31) / 32), (unsigned)((D + 31) / 32));
    k_nchw_to_tok<<<grid, dim3(32, 8), 0, s>>>(in, T, hw, D, denom, out32, out16);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_tok_to_nchw(const float* in, int64_t T, int64_t hw, int D, float* out, cudaStream_t s) {
    if (T == 0) return cudaSuccess;
    dim3 grid((unsigned)((T + 31) / 32), (unsigned)((D + 31) / 32));
    k_tok_to_nchw<<<grid, dim3(32, 8), 0, s>>>(in, T, hw, D, out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace vq
