"""Seeded synthetic inputs of the bench workloads (neutral ground: bench.py's product arm must not need oracle/).

Same draws as SURVEY.md section 8(d) names: codebooks initialised like the reference's two ``Codebook`` constructors
(/root/reference/models/vitvqgan.py:149 ``normal_()``, /root/reference/models/vqgan.py:146 ``uniform_(-1/K, 1/K)``),
N(0, 1) latents and upstream gradients.  CPU generators, so that every rank / device sees identical codebooks.
"""
from __future__ import annotations

import torch


def make_codebook(form: str, K: int, D: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    if form == "vit":
        return torch.randn(K, D, generator=g)
    return (torch.rand(K, D, generator=g) * 2 - 1) / K


def make_latents(shape, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


# BASELINE.json configs (per-GPU shape under weak scaling; the leading dim is what strong scaling divides)
CONFIGS = {
    "cfg1": dict(form="vit", K=8192, D=32, shape=(2, 1024, 32), mode="step",
                 desc="cfg1: ViT-VQGAN quantiser 8192x32, batch 2 x 1024 tokens (the reference's CPU-runnable case), fwd+bwd"),
    "cfg2": dict(form="vqgan", K=8192, D=256, shape=(64, 256, 16, 16), mode="encode",
                 desc="cfg2: VQGAN quantiser 8192x256, encode_imgs batch 64 x 16x16 latents (NCHW), indices only"),
    "cfg2fwd": dict(form="vqgan", K=8192, D=256, shape=(64, 256, 16, 16), mode="step",
                    desc="cfg2 (full): VQGAN quantiser 8192x256, batch 64 x 16x16 latents (NCHW), fwd+bwd"),
    "cfg3": dict(form="vit", K=8192, D=32, shape=(256, 1024, 32), mode="step",
                 desc="cfg3: ViT-VQGAN quantiser fwd+bwd (STE + codebook grad), codebook 8192x32, 256 img x 1024 tok"),
    "cfg4": dict(form="vit", K=8192, D=32, shape=(512, 1024, 32), mode="roundtrip",
                 desc="cfg4: MaskGIT/Muse tokenisation, encode_imgs + decode_indices round trip, 8192x32, 512 img x 1024 tok"),
}


def resolve_config(name: str):
    """'cfgN' or 'sweep:T,K,D' (token-major fwd+bwd step over T tokens, BASELINE.json configs[4])."""
    if name in CONFIGS:
        return dict(CONFIGS[name], name=name)
    if name.startswith("sweep:"):
        T, K, D = (int(v) for v in name[6:].split(","))
        if T % 1024:
            raise ValueError("sweep T must be a multiple of 1024")
        return dict(form="vit", K=K, D=D, shape=(T // 1024, 1024, D), mode="step", name=name,
                    desc=f"cfg5 sweep point: token-major fwd+bwd, {T} tokens x codebook {K}x{D}")
    raise ValueError(f"unknown config {name!r}: cfg1, cfg2, cfg2fwd, cfg3, cfg4 or sweep:T,K,D")


def local_shape(cfg, world: int, scaling: str):
    """Per-rank input shape: weak = the config's shape on every rank, strong = its batch split over the ranks."""
    shape = tuple(cfg["shape"])
    if scaling == "weak" or world == 1:
        return shape
    if shape[0] % world:
        raise ValueError(f"batch {shape[0]} of {cfg['name']} does not split over {world} ranks")
    return (shape[0] // world,) + shape[1:]


def tokens_of(shape, D: int) -> int:
    n = 1
    for s in shape:
        n *= s
    return n // D
