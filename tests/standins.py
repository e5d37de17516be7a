"""Stand-ins of the reference's wrapper models for tests that cannot import /root/reference (the GPU box does not have
it): the same call sites around the quantiser, small encoders / decoders in place of the real ones.

  ViTVQGANStandIn  /root/reference/models/vitvqgan.py:180-210  (pre_quant / post_quant are nn.Linear)
  VQGANStandIn     /root/reference/models/vqgan.py:221-251      (pre_quant / post_quant are 1x1 nn.Conv2d,
                                                                 encode_imgs reshapes the flat indices to (b, n))
  OracleCodebook   the oracle restatement (oracle/vq_oracle.py) behind the reference Codebook's module surface
"""
import torch
import torch.nn as nn

from oracle import vq_oracle as vo


class OracleCodebook(nn.Module):
    def __init__(self, form, codebook_size, codebook_dim, beta=0.25):
        super().__init__()
        self.form, self.codebook_size, self.codebook_dim, self.beta = form, codebook_size, codebook_dim, beta
        self.embedding = nn.Embedding(codebook_size, codebook_dim)

    def forward(self, z):
        o = vo.quantise(self.form, z, self.embedding.weight, self.beta)
        return o.z_q, o.indices, o.loss

    def indices_to_embeddings(self, indices):
        return vo.indices_to_embeddings(self.form, indices, self.embedding.weight)


class ViTVQGANStandIn(nn.Module):
    def __init__(self, patch_dim, dim, codebook_size, codebook_dim):
        super().__init__()
        self.encoder = nn.Sequential(nn.Linear(patch_dim, dim), nn.GELU(), nn.Linear(dim, dim))
        self.pre_quant = nn.Linear(dim, codebook_dim)
        self.codebook = OracleCodebook("vit", codebook_size, codebook_dim)
        self.post_quant = nn.Linear(codebook_dim, dim)
        self.decoder = nn.Sequential(nn.Linear(dim, dim), nn.GELU(), nn.Linear(dim, patch_dim))

    def forward(self, imgs):                          # vitvqgan.py:190-196
        enc_imgs = self.encoder(imgs)
        enc_imgs = self.pre_quant(enc_imgs)
        embeds, indices, loss = self.codebook(enc_imgs)
        embeds = self.post_quant(embeds)
        out = self.decoder(embeds)
        return out, loss

    def decode_indices(self, indices):                # vitvqgan.py:198-202
        embeds = self.codebook.indices_to_embeddings(indices)
        embeds = self.post_quant(embeds)
        return self.decoder(embeds)

    def encode_imgs(self, imgs):                      # vitvqgan.py:204-210
        enc_imgs = self.encoder(imgs)
        enc_imgs = self.pre_quant(enc_imgs)
        _, indices, _ = self.codebook(enc_imgs)
        return indices


class VQGANStandIn(nn.Module):
    def __init__(self, in_ch, dim, codebook_size):
        super().__init__()
        self.encoder = nn.Sequential(nn.Conv2d(in_ch, dim, 3, stride=2, padding=1), nn.SiLU(), nn.Conv2d(dim, dim, 3, padding=1))
        self.pre_quant = nn.Conv2d(dim, dim, 1)
        self.codebook = OracleCodebook("vqgan", codebook_size, dim)
        self.post_quant = nn.Conv2d(dim, dim, 1)
        self.decoder = nn.Sequential(nn.Conv2d(dim, dim, 3, padding=1), nn.SiLU(), nn.ConvTranspose2d(dim, in_ch, 4, stride=2, padding=1))

    def forward(self, imgs):                          # vqgan.py:231-237
        enc_imgs = self.encoder(imgs)
        enc_imgs = self.pre_quant(enc_imgs)
        embeds, indices, loss = self.codebook(enc_imgs)
        embeds = self.post_quant(embeds)
        out = self.decoder(embeds)
        return out, loss

    def decode_indices(self, indices):                # vqgan.py:239-243
        embeds = self.codebook.indices_to_embeddings(indices)
        embeds = self.post_quant(embeds)
        return self.decoder(embeds)

    def encode_imgs(self, imgs):                      # vqgan.py:245-251
        b = imgs.shape[0]
        enc_imgs = self.encoder(imgs)
        enc_imgs = self.pre_quant(enc_imgs)
        _, indices, _ = self.codebook(enc_imgs)
        return indices.reshape(b, -1)                 # rearrange(indices, '(b i) -> b i', b=b)
