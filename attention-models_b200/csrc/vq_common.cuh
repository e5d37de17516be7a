// Shared device helpers for the VQ quantiser kernels (sm_100a).
//
// Arithmetic rules that make z_q bit-equal to what the reference computes on a GPU
// (reference models/vitvqgan.py:16-17,151-171; models/vqgan.py:7-8,148-176; ATen order from
// torch/include/ATen/native/cuda/Reduce.cuh, restated in oracle/aten_order.py):
//   * every rounded operation of the reference is one explicit __f*_rn intrinsic here, so the
//     compiler can neither contract nor reassociate it;
//   * row sums of squares follow ATen's lane mapping, 4-accumulator combine and shuffle-down tree.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "vq_kernels.h"

#define VQ_NORM_EPS 1e-12f          // F.normalize eps (torch/nn/functional.py:5707-5708)
#define VQ_SEG_SHIFT 30              // fixed-point scale of the codebook-gradient segment sums
#define VQ_LOSS_SHIFT 24             // fixed-point scale of the per-row loss sums
#define VQ_NEAR_TIE_REL 1e-6f        // north_star: top-2 distances closer than this are "near ties"
#define VQ_FULL 0xffffffffu

namespace vq {

// programmatic dependent launch (see launch_pdl in vq_kernels.h); both are no-ops in a classic launch
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Row <-> lane mapping.  One warp owns one row of D floats.
//   D <  128 : lane l holds elements l + W*j           (W = min(D, 32); ATen: lanes stride the row)
//   D >= 128 : lane l holds float4 #(l + 32*m), m < D/128  (ATen: "vectorize along input")
// D must be a power of two in [16, 512] (dispatch in vq_abi.cu).
// ---------------------------------------------------------------------------------------------
template <int D>
struct RowMap {
    static constexpr bool kVec = (D >= 128);
    static constexpr int kWidth = (D < 32) ? D : 32;                   // lanes that hold data
    static constexpr int kPerLane = kVec ? (D / 32) : (D / kWidth);     // floats per lane
    static_assert(D >= 16 && D <= 512 && (D & (D - 1)) == 0, "unsupported codebook_dim");

    __device__ static __forceinline__ bool active(int lane) { return lane < kWidth; }
    // element index of register slot j in lane `lane`
    __device__ static __forceinline__ int elem(int lane, int j) {
        if constexpr (kVec) return (lane + 32 * (j >> 2)) * 4 + (j & 3);
        else return lane + kWidth * j;
    }

    __device__ static __forceinline__ void load(const float* __restrict__ row, int lane, float (&x)[kPerLane]) {
        if constexpr (kVec) {
#pragma unroll
            for (int m = 0; m < kPerLane / 4; ++m) {
                float4 v = __ldg(reinterpret_cast<const float4*>(row) + lane + 32 * m);
                x[4 * m + 0] = v.x; x[4 * m + 1] = v.y; x[4 * m + 2] = v.z; x[4 * m + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kPerLane; ++j) x[j] = active(lane) ? __ldg(row + lane + kWidth * j) : 0.f;
        }
    }

    __device__ static __forceinline__ void store(float* __restrict__ row, int lane, const float (&x)[kPerLane]) {
        if constexpr (kVec) {
#pragma unroll
            for (int m = 0; m < kPerLane / 4; ++m)
                reinterpret_cast<float4*>(row)[lane + 32 * m] =
                    make_float4(x[4 * m + 0], x[4 * m + 1], x[4 * m + 2], x[4 * m + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < kPerLane; ++j)
                if (active(lane)) row[lane + kWidth * j] = x[j];
        }
    }

    __device__ static __forceinline__ void store_half(__half* __restrict__ row, int lane, const float (&x)[kPerLane]) {
        if constexpr (kVec) {
#pragma unroll
            for (int m = 0; m < kPerLane / 4; ++m) {
                __half2 a = __floats2half2_rn(x[4 * m + 0], x[4 * m + 1]);
                __half2 b = __floats2half2_rn(x[4 * m + 2], x[4 * m + 3]);
                uint2 u;
                u.x = *reinterpret_cast<uint32_t*>(&a);
                u.y = *reinterpret_cast<uint32_t*>(&b);
                reinterpret_cast<uint2*>(row)[lane + 32 * m] = u;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kPerLane; ++j)
                if (active(lane)) row[lane + kWidth * j] = __float2half_rn(x[j]);
        }
    }

    // Sum over the row in ATen's CUDA order; the result is broadcast to every lane.
    // kFused: each step is fma(v, v, acc) on raw values (NormTwoOps, vector norm).
    // !kFused: v*v is rounded first, then added (torch.sum(t**2, dim=1)).
    template <bool kFused>
    __device__ static __forceinline__ float sumsq(const float (&x)[kPerLane]) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if constexpr (kVec) {
#pragma unroll
            for (int m = 0; m < kPerLane / 4; ++m)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float v = x[4 * m + i];
                    acc[i] = kFused ? __fmaf_rn(v, v, acc[i]) : __fadd_rn(acc[i], __fmul_rn(v, v));
                }
        } else {
            // <= 4 elements per lane, one per accumulator (thread_reduce_impl's tail path)
#pragma unroll
            for (int j = 0; j < kPerLane; ++j) acc[j] = __fmul_rn(x[j], x[j]);
        }
        float v = __fadd_rn(__fadd_rn(__fadd_rn(acc[0], acc[1]), acc[2]), acc[3]);
#pragma unroll
        for (int off = kWidth >> 1; off > 0; off >>= 1) v = __fadd_rn(v, __shfl_down_sync(VQ_FULL, v, off));
        return __shfl_sync(VQ_FULL, v, 0);
    }
};

// Token indices on the wire: int64 (what the reference returns), or int32 / uint16 where the caller asked for a narrow
// format (encode-only calls and the consumers behind them; K <= 65536 for 16 bits).  `bits` is warp-uniform.
__device__ __forceinline__ void store_token(void* base, int64_t t, int code, int bits) {
    if (bits == 64) static_cast<int64_t*>(base)[t] = code;
    else if (bits == 32) static_cast<int32_t*>(base)[t] = code;
    else static_cast<uint16_t*>(base)[t] = (uint16_t)code;
}
__device__ __forceinline__ int64_t load_token(const void* base, int64_t t, int bits) {
    if (bits == 64) return __ldg(static_cast<const int64_t*>(base) + t);
    if (bits == 32) return (int64_t)__ldg(static_cast<const int32_t*>(base) + t);
    return (int64_t)__ldg(static_cast<const uint16_t*>(base) + t);
}

// clamp_min(eps) keeps NaN (ATen semantics); fmaxf would drop it.
__device__ __forceinline__ float clamp_min_keep_nan(float v, float lo) { return (v < lo) ? lo : v; }

__device__ __forceinline__ float norm_denominator(float sumsq) {
    return clamp_min_keep_nan(__fsqrt_rn(sumsq), VQ_NORM_EPS);
}

// torch.argmin ordering (LessOrNan, torch/include/ATen/native/SharedReduceOps.h:435-446):
// NaN beats everything, ties go to the lower index.
__device__ __forceinline__ bool argmin_better(float d_new, int i_new, float d_old, int i_old) {
    const bool nan_new = (d_new != d_new), nan_old = (d_old != d_old);
    if (nan_new || nan_old) {
        if (nan_new && nan_old) return i_new < i_old;
        return nan_new;
    }
    return (d_new < d_old) || (d_new == d_old && i_new < i_old);
}

// d = (|zn|^2 + |en_k|^2) - 2 * dot  with the reference's association (models/vitvqgan.py:157-159)
__device__ __forceinline__ float ref_distance(float row_sq, float code_sq, float dot) {
    return __fsub_rn(__fadd_rn(row_sq, code_sq), __fmul_rn(2.f, dot));
}

__device__ __forceinline__ bool is_finite(float v) { return fabsf(v) <= 3.402823466e38f; }

__device__ __forceinline__ long long to_fixed(float v, int shift) {
    return __float2ll_rn(v * static_cast<float>(1ll << shift));
}

// One term of the codebook-gradient segment sum S_k += fixed(q_k - zn_t) as a 64-bit integer reduction (RED.ADD.64):
// integer sums are exact and order-free, so accumulating straight from the token pass is as deterministic as the
// bucketed sum (vq_backward.cu).  Non-finite terms are not added; the caller counts them per code instead.
__device__ __forceinline__ void seg_add(unsigned long long* __restrict__ slot, float d, unsigned& poison) {
    if (is_finite(d)) atomicAdd(slot, (unsigned long long)to_fixed(d, VQ_SEG_SHIFT));
    else poison = 1;
}

// Buffers a forward zeroes before use (stats, histogram, segment sums, fallback-list counters), cleared by the
// first kernel of the call instead of one memset node each.  Sizes in bytes, multiples of 4.
__device__ __forceinline__ void zero_ranges(const ZeroList& zl, int vblock, int vgrid) {
    const size_t per_block = (size_t)blockDim.x * blockDim.y;
    const size_t tid = (size_t)vblock * per_block + threadIdx.y * blockDim.x + threadIdx.x, n_thr = (size_t)vgrid * per_block;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!zl.ptr[i]) continue;
        if (((reinterpret_cast<size_t>(zl.ptr[i]) | zl.bytes[i]) & 15) == 0) {
            uint4* p = static_cast<uint4*>(zl.ptr[i]);
            for (size_t j = tid; j < zl.bytes[i] / 16; j += n_thr) p[j] = make_uint4(0u, 0u, 0u, 0u);
        } else {
            uint32_t* p = static_cast<uint32_t*>(zl.ptr[i]);
            for (size_t j = tid; j < zl.bytes[i] / 4; j += n_thr) p[j] = 0u;
        }
    }
}

// cb.info: kInfoSlots per-block counts of codes whose |en|^2 is not ~1 (zero / non-finite rows), written by the
// codebook preparation (every slot, no atomics, nothing to zero).  Any non-zero slot = degenerate codebook.
constexpr int kInfoSlots = 64;
__device__ __forceinline__ bool codebook_degenerate(const int* __restrict__ info) {
    int any = 0;
#pragma unroll
    for (int i = 0; i < kInfoSlots / 4; ++i) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(info) + i);
        any |= v.x | v.y | v.z | v.w;
    }
    return any != 0;
}

// loss = beta*m + m (ViT, models/vitvqgan.py:166) or m + beta*m (VQGAN, models/vqgan.py:169), m = mean((q - zn)^2)
// from the fixed-point sum: both reference terms have the same value, only their gradients differ.  Non-finite
// partials make the loss NaN.  `form`: 0 = ViT, 1 = VQGAN (VQ_FORM_*).
__device__ __forceinline__ float loss_from_fixed(long long loss_fixed, long long nonfinite, long long n_elem_total, int form,
                                                 float beta) {
    const double sum = (double)loss_fixed / (double)(1ll << VQ_LOSS_SHIFT);
    float m = (float)(sum / (double)n_elem_total);
    if (nonfinite != 0) m = __int_as_float(0x7fc00000);
    const float bm = __fmul_rn(beta, m);
    return (form == 0) ? __fadd_rn(bm, m) : __fadd_rn(m, bm);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = __fadd_rn(v, __shfl_xor_sync(VQ_FULL, v, off));
    return v;
}

// Pieces of grad_E[k] = (g - y (y . g)) / max(||E_k||, eps), g = coef * S_k, with every rounding explicit: the
// single-GPU kernel, the fused backward and the sharded exchange kernel must give the same bits.
__device__ __forceinline__ float seg_to_grad(long long seg_sum, float coef) {
    return __fmul_rn(coef, (float)((double)seg_sum * (1.0 / (double)(1ll << VQ_SEG_SHIFT))));
}
__device__ __forceinline__ float grad_row_element(float g, float y, float dot, float inv) {
    return __fmul_rn(__fmaf_rn(-y, dot, g), inv);
}

}  // namespace vq
