// pre_quant / post_quant projection fusion.
//
// Reference call sites replaced (paths relative to /root/reference):
//   models/vitvqgan.py:185,192-193,207-208   enc = self.pre_quant(enc)  ->  self.codebook(enc)   (nn.Linear(dim, 32))
//   models/vitvqgan.py:187,199-200           self.post_quant(self.codebook.indices_to_embeddings(indices))
//   models/vqgan.py:228,241-242              the same with nn.Conv2d(dim, dim, 1) on (b, D, h, w)
//
// k_prequant_prep: the Linear C -> 32 and the token preparation (K1, vq_prep.cu) in one pass.  The encoder rows x (T, C)
// are read once (2 KB per token at C = 512: the kernel is HBM-bound on them), z = x W^T + b lives only in registers /
// shared memory, and what reaches HBM is exactly what K1 writes: unit rows in fp32 and fp16, sum(zn^2), max(|z|, eps).
// The 2 x 4D bytes per token of the z round trip and one launch disappear.
//   * GEMM: warp-level tensor-core MMAs (mma.sync m16n8k8, tf32 operands, fp32 accumulators) in the 3xTF32 split
//     x = xh + xl, W = Wh + Wl, z += xl Wh + xh Wl + xh Wh: every product of two tf32 numbers is exact in fp32, the dropped
//     xl Wl term is below 2^-22 of a product, so z carries fp32-GEMM-level rounding (a different summation order than
//     cuBLAS, like any two fp32 GEMMs).  N = 32 leaves a tcgen05 tile 7/8 empty and its fixed issue cost unamortised
//     (DESIGN section 7), and the op is bound by the 2 KB per token it reads: the legacy warp MMA is the right tool here.
//     The tensor core's accumulate truncates (round toward zero), which would bias a 192-MMA chain by ~1e-5 of |z|; so
//     each k8 step is a 3-MMA chain from a zero accumulator and the running sum over k is added outside with
//     round-to-nearest FADDs (they issue in the shadow of the MMAs).  Measured error: tests/test_gpu_projected.py prints
//     it next to cuBLAS's fp32 error on the same inputs.
//   * the contraction index is permuted so that a lane's A fragments are 8 consecutive floats of its row (two 16-byte
//     shared-memory loads): fragment "k" slots (t, t + 4) of step s hold k = 32 c + 8 t + 2 s, + 1.  W sits in shared memory
//     once per block in that fragment order ({b0, b1} per lane, one conflict-free 8-byte load per (step, n-block)) and is
//     split into its tf32 halves where it is used.
//   * the encoder rows reach the MMAs through a per-warp ring of three 32-row x 32-column tiles filled by cp.async
//     (16-byte copies, L2-only, 144-byte row pitch: conflict-free fragment loads): two chunks are always in flight per warp
//     -- 64 KB per SM -- and nothing about it is up to the compiler.  The first version prefetched into registers; at 255
//     registers ptxas sank those loads to the end of the loop body, and 29 % of all stall samples sat on the first use of
//     the loaded value (profiles/r02b_prequant_ncu_variantB.md: 226 us, tensor pipe 44 %).
//   * normalisation: the accumulators go through a per-warp shared-memory tile into K1's lane mapping and then through
//     K1's own instruction sequence (ATen's summation order, explicit _rn intrinsics): given the same z the outputs are
//     bit-identical to vq_prep.cu's.
// Persistent blocks (one per SM, 8 warps, 32 rows per warp and iteration; a warp's (tile, chunk) pairs form one stream, so
// the ring stays full across tile boundaries).
// Measured (cfg3 rows x 512 features, profiles/r02b_*): 228 us = 0.35 of the HBM roofline, 8 % faster than cuBLAS's fp32
// Linear alone (249 us) and 1.11x over Linear + K1 + the quantiser for the whole encode.  The floors are 82 us of HBM time
// and 86 us of tensor time (12.6 M warp MMAs at the 8 cycles per m16n8k8 and sub-partition the legacy path sustains,
// profiles/r02_ubench_hmma.txt); what keeps the kernel 2.7x above them is instruction issue: 1007 instructions per warp
// and chunk around 96 MMAs -- cvt.rna.tf32.f32 is a 4-instruction sequence on sm_100a (FSETP / add / LOP3 / SEL), 128 of
// them per chunk -- with two warps per scheduler at 230 registers (issue slots 57 % busy, tensor pipe 43 %, the rest
// fixed-latency dependencies; profiles/r02b_prequant_ncu.md).  Next: the split as two integer ops (rna = add 0x1000 to the bit
// pattern, the MMA truncates the operand itself; NaN payloads need one guard per fragment, not per value), W split once
// into shared memory again for C <= 512, the tile/chunk division hoisted, and 16 rows per warp at <= 128 registers so that
// four warps per scheduler cover the dependency latency.
//
// k_project_codebook: table[k] = W_post y_k + b_post over the K codes (y = unit code or raw code): with it
// decode_indices' lookup + projection is ONE gather from a (K, C) table (16 MB at K = 8192, C = 512: L2-resident)
// instead of a gather, a (T, D) round trip and a (T x D x C) GEMM -- the projection is done K times, not T times.
#include "vq_common.cuh"
#include "vq_kernels.h"

namespace vq {

namespace {

constexpr int kPqD = 32;              // codebook_dim this kernel is built for (4 n-blocks of 8)
constexpr int kPqWarps = 8;
constexpr int kPqThreads = kPqWarps * 32;
constexpr int kPqRows = 32;           // rows per warp and iteration: two m16 tiles
constexpr int kPqChunk = 32;          // columns of x per pipeline stage (8 per lane of a quad: one 128-byte line per row)
constexpr int kPqSteps = kPqChunk / 8;      // k8 MMA steps per chunk
constexpr int kPqColMultiple = 64;    // in_features the ABI accepts: multiples of 64
constexpr int kPqZStride = 40;        // floats per row of the staging tile: conflict-free float2 stores / float4 loads
constexpr int kPqZTile = 16 * kPqZStride;   // one m16 tile per warp
constexpr int kPqStages = 3;          // x tiles per warp: one being multiplied, two in flight
constexpr int kPqXStride = 36;        // floats per row of an x tile (144 B): conflict-free 16-byte fragment loads
constexpr int kPqXTile = kPqRows * kPqXStride;

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

__device__ __forceinline__ uint32_t tf32_of(float v) {
#ifdef VQ_PQ_INTSPLIT
    // Experiment build (build.py -DVQ_PQ_INTSPLIT --suffix=intsplit; NOT the default, not yet run on a GPU): round to nearest,
    // ties away, as two integer ops on the bit pattern -- what cvt.rna.tf32.f32 computes for every finite value; ptxas
    // expands the cvt itself into a 4-instruction sequence on sm_100a, and 128 of them per chunk are what makes the
    // kernel issue-bound (profiles/r02b_prequant_ncu.md).  A NaN still reaches z through the low part (x - hi).
    return (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
#else
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return u;
#endif
}
// the low part of the split: x - hi is exact in fp32; the default rounds it to tf32 as well, the experiment build hands the
// MMA the fp32 bits (the tensor core reads the top 19 bits: a truncation of the low part, below 2^-21 of x)
__device__ __forceinline__ uint32_t tf32_low_of(float v, uint32_t hi) {
#ifdef VQ_PQ_INTSPLIT
    return __float_as_uint(__fsub_rn(v, __uint_as_float(hi)));
#else
    return tf32_of(__fsub_rn(v, __uint_as_float(hi)));
#endif
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// d = a b (accumulator operand zero)
__device__ __forceinline__ void mma_tf32_zero(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}

__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// K1's per-row arithmetic for D = 32 (prep_rows_small_body<32, false>, vq_prep.cu): 8 lanes per row, one float4 each.
__device__ __forceinline__ float pq_tree(float4 v, int grp) {
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
        v.x = __fadd_rn(v.x, __shfl_down_sync(VQ_FULL, v.x, off));
        v.y = __fadd_rn(v.y, __shfl_down_sync(VQ_FULL, v.y, off));
        v.z = __fadd_rn(v.z, __shfl_down_sync(VQ_FULL, v.z, off));
        v.w = __fadd_rn(v.w, __shfl_down_sync(VQ_FULL, v.w, off));
    }
    const float total = __fadd_rn(__fadd_rn(v.x, v.z), __fadd_rn(v.y, v.w));
    return __shfl_sync(VQ_FULL, total, grp * 8);
}

__device__ __forceinline__ float4 sq4(const float4& v) {
    return make_float4(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y), __fmul_rn(v.z, v.z), __fmul_rn(v.w, v.w));
}

}  // namespace

__global__ void __launch_bounds__(kPqThreads, 1)
k_prequant_prep(const float* __restrict__ x, int64_t T, int C, const float* __restrict__ w, const float* __restrict__ bias,
                float4* __restrict__ unit32, float* __restrict__ sq, float* __restrict__ denom, uint2* __restrict__ unit16,
                float4* __restrict__ z_out, ZeroList zl) {
    extern __shared__ __align__(16) unsigned char pq_smem[];
    float2* wraw = reinterpret_cast<float2*>(pq_smem);                                  // [C/32][4 n-blocks][4 steps][32 lanes]
    float* xring = reinterpret_cast<float*>(pq_smem + (size_t)C * kPqD * 4);             // [warps][stages][32][kPqXStride]
    float* ztile = xring + kPqWarps * kPqStages * kPqXTile;                               // [warps][16][kPqZStride]

    pdl_trigger();
    pdl_wait();
    zero_ranges(zl, blockIdx.x, gridDim.x);

    // ---- W (32, C) -> fragment order ---------------------------------------------------------------------------
    {
        float* wf = reinterpret_cast<float*>(wraw);
        for (int e = threadIdx.x; e < kPqD * C; e += kPqThreads) {
            const int n = e / C, k = e - n * C;
            const int chunk = k >> 5, within = k & 31;
            const int t = within >> 3, s = (within & 7) >> 1, which = within & 1;
            const int j = n >> 3, g = n & 7;
            const int slot = ((chunk * 4 + j) * kPqSteps + s) * 32 + g * 4 + t;
            wf[slot * 2 + which] = __ldg(w + e);
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;        // MMA fragment coordinates
    const int grp = lane >> 3, sub = lane & 7;     // K1's row mapping: 8 lanes per row
    float* zt = ztile + warp * kPqZTile;
    float* xr = xring + warp * (kPqStages * kPqXTile);
    float bias_r[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        bias_r[j][0] = bias ? __ldg(bias + 8 * j + 2 * t) : 0.f;
        bias_r[j][1] = bias ? __ldg(bias + 8 * j + 2 * t + 1) : 0.f;
    }
    const int n_chunks = C / kPqChunk;
    const int64_t n_tiles = (T + kPqRows - 1) / kPqRows;
    // this warp's tiles: first, first + stride, ...; its (tile, chunk) pairs are one stream of n_items work items
    const int64_t first = (int64_t)warp * gridDim.x + blockIdx.x, stride = (int64_t)kPqWarps * gridDim.x;
    const int my_tiles = first < n_tiles ? (int)((n_tiles - first + stride - 1) / stride) : 0;
    const int n_items = my_tiles * n_chunks;

    // item n -> stage n % 3: 32 rows x 32 columns, lane l copies the 16-byte pieces l, l + 32, ... (8 lanes per row)
    auto issue = [&](int n) {
        if (n < n_items) {
            const int ti = n / n_chunks, c = n - ti * n_chunks;
            const int64_t r0 = (first + (int64_t)ti * stride) * kPqRows;
            float* dst = xr + (n % kPqStages) * kPqXTile;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int id = q * 32 + lane, row = id >> 3, piece = id & 7;
                int64_t r = r0 + row;
                if (r >= T) r = T - 1;          // clamped: the stores of such rows are guarded
                cp_async16(dst + row * kPqXStride + piece * 4, x + r * C + c * kPqChunk + piece * 4);
            }
        }
        cp_async_commit();                      // (an empty group keeps the wait count uniform)
    };
    issue(0);
    issue(1);

    float acc[2][4][4];
    for (int n = 0; n < n_items; ++n) {
        const int ti = n / n_chunks, c = n - ti * n_chunks;
        const int64_t r0 = (first + (int64_t)ti * stride) * kPqRows;
        if (c == 0) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.f;
        }
        issue(n + 2);                 // refills the stage item n - 1 was read from (all lanes left it: __syncwarp below)
        cp_async_wait<2>();           // this lane's copies of item n have landed ...
        __syncwarp();                 // ... and so have the other lanes'
        const float* xs = xr + (n % kPqStages) * kPqXTile;
        // rows of this lane's A fragments: 8 i + g (i = 0, 1: first m16 tile; 2, 3: second), columns 8 t .. 8 t + 7
        float4 cur[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < 2; ++q)
                cur[i][q] = *reinterpret_cast<const float4*>(xs + (8 * i + g) * kPqXStride + 8 * t + 4 * q);
        const float2* wb = wraw + (size_t)c * (4 * kPqSteps * 32) + lane;
#pragma unroll
        for (int s = 0; s < kPqSteps; ++s) {
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float e0 = comp(cur[i][s >> 1], (s & 1) * 2), e1 = comp(cur[i][s >> 1], (s & 1) * 2 + 1);
                const uint32_t h0 = tf32_of(e0), h1 = tf32_of(e1);
                const uint32_t l0 = tf32_low_of(e0, h0);
                const uint32_t l1 = tf32_low_of(e1, h1);
                // i even: rows g (a0, a2); i odd: rows g + 8 (a1, a3)
                ah[i >> 1][(i & 1)] = h0;     ah[i >> 1][(i & 1) + 2] = h1;
                al[i >> 1][(i & 1)] = l0;     al[i >> 1][(i & 1) + 2] = l1;
            }
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 b = wb[(j * kPqSteps + s) * 32];
                bh[j][0] = tf32_of(b.x);
                bh[j][1] = tf32_of(b.y);
                bl[j][0] = tf32_low_of(b.x, bh[j][0]);
                bl[j][1] = tf32_low_of(b.y, bh[j][1]);
            }
            // One k8 step inside the tensor core per (n-block, row tile): a 3-MMA chain from a zero accumulator, small terms
            // first, the eight chains phase by phase.  The running sum over k stays outside the tensor core: round-to-nearest
            // adds (the MMA's own accumulate truncates).
            float d[2][4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int m = 0; m < 2; ++m) mma_tf32_zero(d[m][j], al[m], bh[j][0], bh[j][1]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int m = 0; m < 2; ++m) mma_tf32(d[m][j], ah[m], bl[j][0], bl[j][1]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int m = 0; m < 2; ++m) mma_tf32(d[m][j], ah[m], bh[j][0], bh[j][1]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[m][j][q] = __fadd_rn(acc[m][j][q], d[m][j][q]);
        }
        __syncwarp();                 // every lane has read its fragments: the stage may be refilled
        if (c + 1 < n_chunks) continue;

        // ---- z = acc + bias -> K1's mapping -> unit rows ---------------------------------------------------------
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // c0, c1: row g, columns 8 j + 2 t, + 1;  c2, c3: row g + 8
                *reinterpret_cast<float2*>(zt + g * kPqZStride + 8 * j + 2 * t) =
                    make_float2(__fadd_rn(acc[m][j][0], bias_r[j][0]), __fadd_rn(acc[m][j][1], bias_r[j][1]));
                *reinterpret_cast<float2*>(zt + (g + 8) * kPqZStride + 8 * j + 2 * t) =
                    make_float2(__fadd_rn(acc[m][j][2], bias_r[j][0]), __fadd_rn(acc[m][j][3], bias_r[j][1]));
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int rl = it * 4 + grp;
                const int64_t r = r0 + 16 * m + rl;
                float4 v = *reinterpret_cast<const float4*>(zt + rl * kPqZStride + 4 * sub);
                const float4 zraw = v;
                const float den = norm_denominator(pq_tree(sq4(v), grp));
                v.x = __fdiv_rn(v.x, den); v.y = __fdiv_rn(v.y, den); v.z = __fdiv_rn(v.z, den); v.w = __fdiv_rn(v.w, den);
                const float s2 = pq_tree(sq4(v), grp);
                if (r < T) {
                    if (z_out) z_out[r * 8 + sub] = zraw;
                    if (unit32) unit32[r * 8 + sub] = v;
                    if (unit16) {
                        const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
                        uint2 u;
                        u.x = *reinterpret_cast<const uint32_t*>(&a);
                        u.y = *reinterpret_cast<const uint32_t*>(&b);
                        unit16[r * 8 + sub] = u;
                    }
                    if (sub == 0) {
                        if (sq) sq[r] = s2;
                        if (denom) denom[r] = den;
                    }
                }
            }
        }
    }
}

size_t prequant_smem_bytes(int C) {
    return (size_t)C * kPqD * 4 + sizeof(float) * (size_t)kPqWarps * (kPqStages * kPqXTile + kPqZTile);
}

bool prequant_supported(int C, int D) {
    return D == kPqD && C >= kPqColMultiple && C % kPqColMultiple == 0 && prequant_smem_bytes(C) <= 227 * 1024;
}

cudaError_t launch_prequant_prep(const float* x, int64_t T, int C, const float* w, const float* bias, int D, float* zn32,
                                 float* row_sq, float* denom, __half* zn16, float* z_out, const ZeroList& zl, cudaStream_t s) {
    if (!prequant_supported(C, D)) return cudaErrorInvalidValue;
    const size_t smem = prequant_smem_bytes(C);
    // per device, and cheap: set on every launch
    cudaError_t e = cudaFuncSetAttribute(k_prequant_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t n_tiles = (T + kPqRows - 1) / kPqRows;
    int64_t blocks = n_tiles < sm_count() ? n_tiles : sm_count();
    if (blocks < 1) blocks = 1;
    e = launch_pdl(k_prequant_prep, dim3((unsigned)blocks), dim3(kPqThreads), smem, s, x, T, C, w, bias,
                   reinterpret_cast<float4*>(zn32), row_sq, denom, reinterpret_cast<uint2*>(zn16),
                   reinterpret_cast<float4*>(z_out), zl);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// table[k][c] = b[c] + sum_d y[k][d] * W[c][d]   (fp32 fma chain over d = 0 .. D - 1, then the bias: one rounding each)
// 8 codes per block in shared memory; thread c walks row c of W (D floats, read as float4) once for all 8.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kPcCodes = 8;
__global__ void __launch_bounds__(256) k_project_codebook(const float* __restrict__ y, int K, int D, const float* __restrict__ w,
                                                          const float* __restrict__ bias, int C, float* __restrict__ table) {
    extern __shared__ __align__(16) float pc_rows[];      // [kPcCodes][D]
    const int k0 = blockIdx.x * kPcCodes;
    for (int e = threadIdx.x; e < kPcCodes * D; e += blockDim.x) {
        const int k = k0 + e / D;
        pc_rows[e] = (k < K) ? __ldg(y + (int64_t)k * D + (e % D)) : 0.f;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc[kPcCodes];
#pragma unroll
        for (int i = 0; i < kPcCodes; ++i) acc[i] = 0.f;
        const float4* wr = reinterpret_cast<const float4*>(w + (int64_t)c * D);
        for (int d4 = 0; d4 < D / 4; ++d4) {
            const float4 wv = __ldg(wr + d4);
#pragma unroll
            for (int i = 0; i < kPcCodes; ++i) {
                const float4 yv = *reinterpret_cast<const float4*>(pc_rows + i * D + 4 * d4);
                acc[i] = __fmaf_rn(yv.x, wv.x, acc[i]);
                acc[i] = __fmaf_rn(yv.y, wv.y, acc[i]);
                acc[i] = __fmaf_rn(yv.z, wv.z, acc[i]);
                acc[i] = __fmaf_rn(yv.w, wv.w, acc[i]);
            }
        }
        const float bv = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
        for (int i = 0; i < kPcCodes; ++i)
            if (k0 + i < K) table[(int64_t)(k0 + i) * C + c] = __fadd_rn(acc[i], bv);
    }
}

cudaError_t launch_project_codebook(const float* y, int K, int D, const float* w, const float* bias, int C, float* table,
                                    cudaStream_t s) {
    if (K <= 0 || C <= 0) return cudaSuccess;
    const int blocks = (K + kPcCodes - 1) / kPcCodes;
    k_project_codebook<<<blocks, 256, sizeof(float) * kPcCodes * D, s>>>(y, K, D, w, bias, C, table);
    count_launch();
    return cudaGetLastError();
}

}  // namespace vq
