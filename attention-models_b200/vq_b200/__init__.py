"""vq_b200 -- B200-native VQ codebook quantiser, drop-in for the two ``Codebook`` modules of
pranoyr/attention-models (models/vitvqgan.py:140-176 and models/vqgan.py:138-182).

    from vq_b200.vitvqgan import Codebook      # ViT-VQGAN form  (token-major input)
    from vq_b200.vqgan import Codebook         # CNN-VQGAN form  (NCHW input)

All arithmetic runs in hand-written sm_100a CUDA kernels behind a C ABI (include/vq_b200.h);
there is no CPU path.
"""
from . import _lib
from .functional import (PreparedCodebook, encode_indices, indices_to_embeddings, prepare_codebook, quantise)
from .integration import patch_reference_model

__all__ = ["PreparedCodebook", "encode_indices", "indices_to_embeddings", "prepare_codebook", "quantise",
           "patch_reference_model", "_lib"]
