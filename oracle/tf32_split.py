"""Numerical model of the fused pre_quant GEMM (attention-models_b200/csrc/vq_prequant.cu) -- test infrastructure only.

The reference computes ``self.pre_quant(enc_imgs)`` (/root/reference/models/vitvqgan.py:185,192) with an fp32 GEMM.  The
kernel computes it on the warp-level tensor cores in the 3xTF32 split; this file restates that arithmetic in numpy so that the
tolerance the GPU tests use (tests/test_gpu_projected.py: Z_TOL) is checked on the CPU against a float64 product and against
a plain float32 GEMM:

    xh = rna_tf32(x), xl = rna_tf32(x - xh)          (cvt.rna.tf32.f32: round to nearest, ties away, 10 explicit mantissa bits)
    per k8 step:  d = xl.wh + xh.wl + xh.wh           (three MMAs from a zero accumulator; every product of two tf32 numbers
                                                        is exact, the 8-term sums are formed inside the tensor core)
    z += d                                            (fp32 round-to-nearest adds outside the tensor core), then + bias

What the tensor core does with the low bits of its internal sums is not documented, so two models bracket it: ``exact`` keeps
each 3-MMA chain exact and rounds once to fp32 (optimistic), ``truncate`` chops every MMA's result to fp32 toward zero
(pessimistic).  Both must stay at fp32-GEMM level.
"""
from __future__ import annotations

import numpy as np


def rna_tf32(x: np.ndarray) -> np.ndarray:
    """cvt.rna.tf32.f32 on finite values: add half an ulp of the 10-bit mantissa to the bit pattern, clear the low 13 bits."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def split_tf32(x: np.ndarray):
    hi = rna_tf32(x)
    lo = rna_tf32((x.astype(np.float32) - hi).astype(np.float32))
    return hi, lo


def _chop_to_f32_toward_zero(v: np.ndarray) -> np.ndarray:
    r = v.astype(np.float32)
    over = np.abs(r.astype(np.float64)) > np.abs(v)
    return np.where(over, np.nextafter(r, np.float32(0)), r).astype(np.float32)


def linear_3xtf32(x: np.ndarray, w: np.ndarray, bias=None, model: str = "exact") -> np.ndarray:
    """z = x w^T + bias as the kernel forms it.  x: (T, C), w: (D, C), C a multiple of 8."""
    xh, xl = split_tf32(x)
    wh, wl = split_tf32(w)
    T, C = x.shape
    z = np.zeros((T, w.shape[0]), np.float32)
    for k in range(0, C, 8):
        s = slice(k, k + 8)
        a_h, a_l = xh[:, s].astype(np.float64), xl[:, s].astype(np.float64)
        b_h, b_l = wh[:, s].astype(np.float64).T, wl[:, s].astype(np.float64).T
        if model == "exact":
            d = (a_l @ b_h + a_h @ b_l + a_h @ b_h).astype(np.float32)
        else:
            d = _chop_to_f32_toward_zero(a_l @ b_h)
            d = _chop_to_f32_toward_zero(d.astype(np.float64) + a_h @ b_l)
            d = _chop_to_f32_toward_zero(d.astype(np.float64) + a_h @ b_h)
        z = (z + d).astype(np.float32)          # fp32 round-to-nearest add
    if bias is not None:
        z = (z + bias.astype(np.float32)).astype(np.float32)
    return z


def error_over_sum_abs_terms(z: np.ndarray, x: np.ndarray, w: np.ndarray, bias=None) -> float:
    """max |z - z64| / (sum_c |x_c w_dc| + |b_d|): the measure the GPU tests and bench.py --config cfg3pre print."""
    z64 = x.astype(np.float64) @ w.astype(np.float64).T + (0 if bias is None else bias.astype(np.float64))
    scale = np.abs(x).astype(np.float64) @ np.abs(w).astype(np.float64).T + (0 if bias is None else np.abs(bias).astype(np.float64))
    return float(np.max(np.abs(z.astype(np.float64) - z64) / np.maximum(scale, 1e-300)))
