"""numpy emulation of the summation order ATen's CUDA reduce kernel uses (test infrastructure only).

"Bit-exact z_q" means reproducing the reference *as it runs on the GPU*: the
reference calls ``F.normalize`` / ``torch.sum`` (models/vitvqgan.py:16-17,157-158;
models/vqgan.py:7-8,157-158) and ATen's ``gpu_reduce_kernel`` decides the order in
which the D squares of a row are added.  This file restates that order from the
installed torch 2.11.0+cu128 headers so the CUDA kernels can be checked bit for
bit on any machine:

* ``torch/include/ATen/native/cuda/Reduce.cuh:99-108``  set_block_dimension
* ``...Reduce.cuh:1034-1180`` setReduceConfig (vectorise-input iff the reduced dim is the
  fastest one and has >= 128 elements; otherwise lanes stride the row)
* ``...Reduce.cuh:500-560`` input_vectorized_thread_reduce_impl (4 accumulators from float4 loads)
* ``...Reduce.cuh:562-631`` thread_reduce_impl (vt0 = 4 strided accumulators, combined ((a0+a1)+a2)+a3)
* ``...Reduce.cuh:633-671`` block_x_reduce (warp shuffle-down, offsets W/2 ... 1)
* ``...Reduce.cuh:673-690`` block_y_reduce (shared-memory tree over threadIdx.y)
* ``torch/include/ATen/native/SharedReduceOps.h:378-394`` NormTwoOps: acc + x*x (an FMA when
  nvcc contracts it), project = sqrt

Whether these orders really are what the installed binary does is confirmed on
the B200 by ``tests/test_gpu_aten_order.py`` (torch CUDA vs this emulation).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
VT0 = 4  # accumulators per thread in ATen's reduce kernel


def fma32(a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """Correctly rounded float32 fma(a, b, c) for float32 inputs (normal range).

    a*b is exact in float64.  TwoSum gives p + c == s + e exactly with s = fl64(p + c);
    rounding s to float32 is already the right answer unless s lies exactly half-way
    between two float32 values (low 29 mantissa bits == 0x10000000), in which case the
    discarded e breaks the tie.
    """
    a64, b64, c64 = a.astype(np.float64), b.astype(np.float64), c.astype(np.float64)
    p = a64 * b64
    s = p + c64
    bb = s - p
    e = (p - (s - bb)) + (c64 - bb)
    r = s.astype(F32)
    bits = s.view(np.uint64)
    tie = ((bits & np.uint64(0x1FFFFFFF)) == np.uint64(0x10000000)) & (e != 0)
    if tie.any():
        toward_zero = (bits & ~np.uint64(0x1FFFFFFF)).view(np.float64).astype(F32)
        away = np.nextafter(toward_zero, np.copysign(F32(np.inf), toward_zero).astype(F32))
        outward = (e * np.sign(s)) > 0
        r = np.where(tie, np.where(outward, away, toward_zero), r).astype(F32)
    return r


def _accumulate(acc: np.ndarray, col: np.ndarray, fused: bool) -> np.ndarray:
    if fused:                      # NormTwoOps::reduce  acc + x*x  -> FFMA
        return fma32(col, col, acc)
    return (acc + col).astype(F32)  # plain sum of an already materialised x**2


def _thread_reduce(x: np.ndarray, start: int, stride: int, fused: bool) -> np.ndarray:
    """thread_reduce_impl for one thread position (vectorised over rows of x)."""
    rows, n = x.shape
    acc = [np.zeros(rows, F32) for _ in range(VT0)]
    idx = start
    while idx + (VT0 - 1) * stride < n:
        for i in range(VT0):
            acc[i] = _accumulate(acc[i], x[:, idx + i * stride], fused)
        idx += VT0 * stride
    for i in range(VT0):
        if idx >= n:
            break
        acc[i] = _accumulate(acc[i], x[:, idx], fused)
        idx += stride
    out = acc[0]
    for i in range(1, VT0):
        out = (out + acc[i]).astype(F32)
    return out


def _tree(vals: np.ndarray) -> np.ndarray:
    """shuffle-down / shared-memory tree: offsets W/2 ... 1, result in slot 0.  vals: (rows, W)."""
    v = vals.copy()
    off = v.shape[1] // 2
    while off > 0:
        v[:, :off] = (v[:, :off] + v[:, off:2 * off]).astype(F32)
        off //= 2
    return v[:, 0]


def _last_pow2(n: int) -> int:
    p = 1
    while p * 2 <= n:
        p *= 2
    return p


def rowsum_contiguous(x: np.ndarray, fused: bool) -> np.ndarray:
    """Sum over the last (contiguous) dim of a (rows, D) float32 array, ATen CUDA order.

    ``fused=True``: x holds the raw values and each step is fma(v, v, acc) (vector norm).
    ``fused=False``: x already holds the squares and each step is acc + v (torch.sum(t**2, 1)).
    """
    x = np.ascontiguousarray(x, dtype=F32)
    rows, D = x.shape
    if D >= 128 and D % 4 == 0:
        # vectorise-input: lane L accumulates float4 #L, #L+32, ...; component i -> accumulator i
        lanes = 32
        per_lane = np.zeros((rows, lanes), F32)
        for lane in range(lanes):
            acc = [np.zeros(rows, F32) for _ in range(VT0)]
            v = lane
            while v * 4 + 3 < D:
                for i in range(VT0):
                    acc[i] = _accumulate(acc[i], x[:, v * 4 + i], fused)
                v += lanes
            out = acc[0]
            for i in range(1, VT0):
                out = (out + acc[i]).astype(F32)
            per_lane[:, lane] = out
        return _tree(per_lane)
    if D >= 128:
        raise NotImplementedError("D >= 128 with D % 4 != 0 (unaligned vectorised head/tail) is not mirrored")
    lanes = min(_last_pow2(D), 32)
    per_lane = np.stack([_thread_reduce(x, lane, lanes, fused) for lane in range(lanes)], axis=1)
    return _tree(per_lane)


def strided_stripes(T: int, hw: int, D: int) -> int:
    """Number of channel stripes ATen's block_y_reduce combines for a (b, D, hw) view reduced over D
    (Reduce.cuh:1034-1180 with :99-108): depends on the vector width (hw % 4 / % 2) and on how many
    outputs there are, because few outputs shrink block.x and grow block.y."""
    vec = 4 if hw % 4 == 0 else (2 if hw % 2 == 0 else 1)
    max_threads = 512 // vec
    dim0 = max(T // vec, 1)
    dim0_pow2 = _last_pow2(dim0) if dim0 < max_threads else max_threads
    dim1_pow2 = _last_pow2(D) if D < max_threads else max_threads
    bw = min(dim0_pow2, 32)
    bh = min(dim1_pow2, max_threads // bw)
    return bh if D >= min(bh * 16, 256) else 1


def rowsum_channel_strided(x: np.ndarray, fused: bool) -> np.ndarray:
    """Sum over dim 1 of a (b, D, hw) float32 array (the NCHW view the VQGAN form normalises),
    ATen CUDA order: thread y accumulates channels y + S*(i + 4m) in accumulator i, then a tree over y."""
    x = np.ascontiguousarray(x, dtype=F32)
    b, D, hw = x.shape
    rows = np.ascontiguousarray(x.transpose(0, 2, 1)).reshape(b * hw, D)
    split = strided_stripes(b * hw, hw, D)
    parts = np.stack([_thread_reduce(rows, y, split, fused) for y in range(split)], axis=1)
    return _tree(parts)


def normalise_contiguous(x: np.ndarray, eps: float = 1e-12):
    """F.normalize over the last dim of (rows, D): returns (unit rows, denominators)."""
    x = np.ascontiguousarray(x, dtype=F32)
    denom = np.maximum(np.sqrt(rowsum_contiguous(x, fused=True)), F32(eps)).astype(F32)
    return (x / denom[:, None]).astype(F32), denom


def normalise_nchw(x: np.ndarray, eps: float = 1e-12):
    """F.normalize over the channel dim of a (b, D, hw) view: returns (token-major unit rows, denominators)."""
    b, D, hw = x.shape
    denom = np.maximum(np.sqrt(rowsum_channel_strided(x, fused=True)), F32(eps)).astype(F32)
    rows = np.ascontiguousarray(x.transpose(0, 2, 1)).reshape(b * hw, D).astype(F32)
    return (rows / denom[:, None]).astype(F32), denom
