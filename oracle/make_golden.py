"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python oracle/make_golden.py

Imports ``Codebook`` from ``/root/reference/models/vitvqgan.py:140-176`` and
``/root/reference/models/vqgan.py:138-182`` (read-only mount, nothing is copied),
feeds them the seeded inputs of SURVEY.md section 8(d) and stores what they return.
Inputs are NOT stored: they are regenerated bit-identically from the seeds by
``oracle.vq_oracle.make_codebook / make_latents`` (CPU ``torch.Generator``).
The reference never runs on the GPU box; these files are how it travels.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vq_oracle as vo  # noqa: E402  (input generators only)

REF = "/root/reference"
if REF not in sys.path:
    sys.path.insert(0, REF)
from models.vitvqgan import Codebook as RefVitCodebook    # noqa: E402
from models.vqgan import Codebook as RefVqganCodebook      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(1)            # fixed CPU reduction order for the fixtures
torch.use_deterministic_algorithms(True)


def _module(form, K, D, beta, weight):
    cls = RefVitCodebook if form == vo.VIT else RefVqganCodebook
    m = cls(K, D, beta)
    with torch.no_grad():
        m.embedding.weight.copy_(weight)
    return m


def forward_case(name, form, K, D, shape, w_seed, z_seed, beta=0.25, edit=None):
    w = vo.make_codebook(form, K, D, w_seed)
    z = vo.make_latents(shape, z_seed)
    if edit is not None:
        edit(z, w)
    m = _module(form, K, D, beta, w)
    with torch.no_grad():
        z_q, idx, loss = m(z)
    idx_dtype = np.uint16 if K <= 65536 else np.int64
    np.savez_compressed(os.path.join(OUT, name + ".npz"), form=form, K=K, D=D, shape=np.array(shape),
                        w_seed=w_seed, z_seed=z_seed, beta=beta, edit=(edit.__name__ if edit else ""),
                        z_q=z_q.numpy(), indices=idx.numpy().astype(idx_dtype), idx_shape=np.array(idx.shape),
                        loss=loss.numpy())
    print(name, "loss", float(loss), "idx[:4]", idx.reshape(-1)[:4].tolist())


def step_case(name, form, K, D, shape, w_seed, z_seed, g_seed, beta=0.25):
    w = vo.make_codebook(form, K, D, w_seed)
    z = vo.make_latents(shape, z_seed).requires_grad_(True)
    up = vo.make_latents(shape, g_seed)
    m = _module(form, K, D, beta, w)
    z_q, idx, loss = m(z)
    ((z_q * up).sum() + loss).backward()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), form=form, K=K, D=D, shape=np.array(shape),
                        w_seed=w_seed, z_seed=z_seed, g_seed=g_seed, beta=beta,
                        z_q=z_q.detach().numpy(), indices=idx.numpy().astype(np.uint16),
                        idx_shape=np.array(idx.shape), loss=loss.detach().numpy(),
                        grad_z=z.grad.numpy(), grad_weight=m.embedding.weight.grad.numpy())
    print(name, "loss", float(loss), "|grad_w|", float(m.embedding.weight.grad.norm()))


def decode_case(name, form, K, D, b, n, w_seed, i_seed):
    w = vo.make_codebook(form, K, D, w_seed)
    g = torch.Generator().manual_seed(i_seed)
    idx = torch.randint(0, K, (b, n), generator=g)
    m = _module(form, K, D, 0.25, w)
    with torch.no_grad():
        e = m.indices_to_embeddings(idx)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), form=form, K=K, D=D, b=b, n=n, w_seed=w_seed,
                        i_seed=i_seed, embeds=e.contiguous().numpy(), embeds_shape=np.array(e.shape))
    print(name, tuple(e.shape))


def token_embed_case(name, vocab, dim, b, n, seed):
    """MaskGIT's token consumer on the reference's own modules: `fill_mask`'s two masked_fill lines
    (/root/reference/models/maskgit.py:131-132) on a seeded mask, then `input_proj(x)`; `x += pos_enc`
    (models/maskgit.py:80-81) of an unmodified BiDirectionalTransformer."""
    from models.maskgit import BiDirectionalTransformer
    g = torch.Generator().manual_seed(seed)
    m = BiDirectionalTransformer(dim=dim, vocab_size=vocab, num_patches=n, n_heads=2, d_head=16, dec_depth=1)
    with torch.no_grad():
        m.input_proj.weight.copy_(torch.randn(vocab + 1, dim, generator=g))
        m.pos_enc.copy_(0.02 * torch.randn(1, n, dim, generator=g))
        tokens = torch.randint(0, vocab, (b, n), generator=g)
        mask = torch.rand(b, n, generator=g) < 0.4
        tgt = tokens.masked_fill(~mask, -1)                       # maskgit.py:131
        x = tokens.masked_fill(mask, m.mask_token_id)             # maskgit.py:132
        e = m.input_proj(x)                                       # maskgit.py:80
        e += m.pos_enc                                            # maskgit.py:81
    np.savez_compressed(os.path.join(OUT, name + ".npz"), vocab=vocab, dim=dim, b=b, n=n, seed=seed,
                        input_ids=x.numpy(), labels=tgt.numpy(), embeds=e.numpy())
    print(name, "masked", int(mask.sum()), "of", b * n)


def token_grad_case(name, vocab, dim, b, n, seed):
    """Backward of the two token consumers through the reference's own modules and autograd:
    MaskGIT (models/maskgit.py:80-81: input_proj lookup, `+= pos_enc` with pos_enc an nn.Parameter) and Parti's shifted
    decoder input (models/parti.py:98-106: token_emb(tokens[:, :-1]), PositionalEncoding, start token; eval mode so the
    dropout of models/positional_encoding.py:42 is the identity)."""
    from einops import repeat
    from models.maskgit import BiDirectionalTransformer
    from models.positional_encoding import PositionalEncoding
    g = torch.Generator().manual_seed(seed)
    m = BiDirectionalTransformer(dim=dim, vocab_size=vocab, num_patches=n, n_heads=2, d_head=16, dec_depth=1)
    with torch.no_grad():
        m.input_proj.weight.copy_(torch.randn(vocab + 1, dim, generator=g))
        m.pos_enc.copy_(0.02 * torch.randn(1, n, dim, generator=g))
    tokens = torch.randint(0, vocab, (b, n), generator=g)
    mask = torch.rand(b, n, generator=g) < 0.4
    up = torch.randn(b, n, dim, generator=g) * 1e-3
    x = tokens.masked_fill(mask, m.mask_token_id)
    e = m.input_proj(x)
    e = e + m.pos_enc            # (out-of-place form of maskgit.py:81 so that autograd can run on the leaf)
    (e * up).sum().backward()
    # Parti
    token_emb = torch.nn.Embedding(vocab, dim)
    pos = PositionalEncoding(dim).eval()
    start = torch.nn.Parameter(torch.randn(dim, generator=g))
    with torch.no_grad():
        token_emb.weight.copy_(torch.randn(vocab, dim, generator=g))
    inp, labels = tokens[:, :-1], tokens                                   # parti.py:98
    pe = token_emb(inp)                                                    # parti.py:100
    pe = pos(pe)                                                           # parti.py:102
    pe = torch.cat((repeat(start, 'd -> b 1 d', b=b), pe), dim=1)          # parti.py:104-105
    (pe * up).sum().backward()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), vocab=vocab, dim=dim, b=b, n=n, seed=seed, tokens=tokens.numpy(),
                        mask=mask.numpy(), upstream=up.numpy(), maskgit_table=m.input_proj.weight.detach().numpy(),
                        maskgit_pos=m.pos_enc.detach().numpy(), maskgit_embeds=e.detach().numpy(),
                        maskgit_grad_table=m.input_proj.weight.grad.numpy(), maskgit_grad_pos=m.pos_enc.grad.numpy(),
                        parti_table=token_emb.weight.detach().numpy(), parti_pe=pos.pe[:n].numpy(), parti_start=start.detach().numpy(),
                        parti_embeds=pe.detach().numpy(), parti_labels=labels.numpy(),
                        parti_grad_table=token_emb.weight.grad.numpy(), parti_grad_start=start.grad.numpy())
    print(name, "ok")


def decode_grad_case(name, form, K, D, b, n, seed):
    """indices_to_embeddings is differentiable w.r.t. the codebook in the reference (nn.Embedding lookup; the ViT form
    normalises behind it): gradients of the unmodified Codebook classes through autograd."""
    cb = _module(form, K, D, 0.25, vo.make_codebook(form, K, D, seed))
    g = torch.Generator().manual_seed(seed + 1)
    idx = torch.randint(0, K, (b, n), generator=g)
    out = cb.indices_to_embeddings(idx)
    up = torch.randn(out.shape, generator=g)
    (out * up).sum().backward()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), form=form, K=K, D=D, weight=cb.embedding.weight.detach().numpy(),
                        indices=idx.numpy(), upstream=up.numpy(), out=out.detach().numpy(),
                        grad_weight=cb.embedding.weight.grad.numpy())
    print(name, "ok")


def projected_case(name, K, D, C, b, n, seed, beta=0.25):
    """The ViT wrapper's call sites around the quantiser on the reference's own modules (models/vitvqgan.py:185-187
    wiring, :192-194 forward, :199-200 decode_indices): nn.Linear pre_quant -> reference Codebook -> nn.Linear post_quant,
    forward + autograd, and post_quant(indices_to_embeddings(indices)) for seeded tokens."""
    w = vo.make_codebook(vo.VIT, K, D, seed)
    w_pre, b_pre = vo.projection_inputs(C, D, seed + 1)
    w_post, b_post = vo.projection_inputs(D, C, seed + 2)
    x = vo.make_latents((b, n, C), seed + 3).requires_grad_(True)
    up = vo.make_latents((b, n, D), seed + 4)
    pre, post = torch.nn.Linear(C, D), torch.nn.Linear(D, C)
    with torch.no_grad():
        pre.weight.copy_(w_pre); pre.bias.copy_(b_pre); post.weight.copy_(w_post); post.bias.copy_(b_post)
    cb = _module(vo.VIT, K, D, beta, w)
    enc = pre(x)                                   # vitvqgan.py:192
    z_q, idx, loss = cb(enc)                       # vitvqgan.py:193
    ((z_q * up).sum() + loss).backward()
    g = torch.Generator().manual_seed(seed + 5)
    tokens = torch.randint(0, K, (b, n), generator=g)
    with torch.no_grad():
        dec = post(cb.indices_to_embeddings(tokens))      # vitvqgan.py:199-200
    np.savez_compressed(os.path.join(OUT, name + ".npz"), K=K, D=D, C=C, b=b, n=n, seed=seed, beta=beta,
                        z=enc.detach().numpy(), z_q=z_q.detach().numpy(), indices=idx.numpy().astype(np.uint16),
                        loss=loss.detach().numpy(), grad_x=x.grad.numpy(), grad_w_pre=pre.weight.grad.numpy(),
                        grad_b_pre=pre.bias.grad.numpy(), grad_weight=cb.embedding.weight.grad.numpy(),
                        tokens=tokens.numpy().astype(np.uint16), decoded=dec.numpy())
    print(name, "loss", float(loss), "idx[:4]", idx.reshape(-1)[:4].tolist())


def projected_conv_case(name, K, D, b, side, seed):
    """models/vqgan.py:241-242 on the reference's own Codebook: post_quant (nn.Conv2d(dim, dim, 1), :228) of
    indices_to_embeddings(indices)."""
    w = vo.make_codebook(vo.VQGAN, K, D, seed)
    w_post, b_post = vo.projection_inputs(D, D, seed + 2, conv=True)
    post = torch.nn.Conv2d(D, D, 1)
    with torch.no_grad():
        post.weight.copy_(w_post); post.bias.copy_(b_post)
    cb = _module(vo.VQGAN, K, D, 0.25, w)
    g = torch.Generator().manual_seed(seed + 5)
    tokens = torch.randint(0, K, (b, side * side), generator=g)
    with torch.no_grad():
        dec = post(cb.indices_to_embeddings(tokens))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), K=K, D=D, b=b, side=side, seed=seed,
                        tokens=tokens.numpy().astype(np.uint16), decoded=dec.numpy())
    print(name, tuple(dec.shape))


def degenerate_rows(z, w):
    """zero row, NaN row, a row equal to a code, a zero code (SURVEY.md section 7 'Degenerate rows')."""
    z[0, 0] = 0.0
    z[0, 1, 3] = float("nan")
    z[0, 2] = w[5] * 3.0
    z[1, 0] = -w[7]


def _selected(fn):
    """`python oracle/make_golden.py name ...` regenerates only the named fixtures."""
    def run(name, *a, **k):
        if len(sys.argv) == 1 or name in sys.argv[1:]:
            fn(name, *a, **k)
    return run


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    forward_case, step_case, decode_case = _selected(forward_case), _selected(step_case), _selected(decode_case)
    token_embed_case, token_grad_case, decode_grad_case = _selected(token_embed_case), _selected(token_grad_case), _selected(decode_grad_case)
    # BASELINE.json configs[0]: ViT form K=8192 D=32, 2 x 1024 tokens
    forward_case("vit_cfg1_fwd", vo.VIT, 8192, 32, (2, 1024, 32), 0, 1)
    # a slice of configs[1]: VQGAN form K=8192 D=256, NCHW 16x16 latents
    forward_case("vqgan_cfg2_slice_fwd", vo.VQGAN, 8192, 256, (4, 256, 16, 16), 0, 2)
    # fwd + bwd (configs[2] objective) at sizes whose gradients stay small on disk
    step_case("vit_step_small", vo.VIT, 1024, 32, (4, 256, 32), 10, 11, 12)
    step_case("vit_step_beta", vo.VIT, 512, 32, (2, 128, 32), 13, 14, 15, beta=0.7)
    step_case("vqgan_step_small", vo.VQGAN, 512, 256, (2, 256, 8, 8), 20, 21, 22)
    step_case("vqgan_step_d64", vo.VQGAN, 256, 64, (3, 64, 4, 4), 23, 24, 25, beta=0.4)
    # decode_indices side
    decode_case("vit_decode", vo.VIT, 8192, 32, 2, 1024, 0, 30)
    decode_case("vqgan_decode", vo.VQGAN, 1024, 256, 2, 64, 31, 32)
    # edge rows
    forward_case("vit_degenerate_fwd", vo.VIT, 64, 32, (2, 8, 32), 40, 41, edit=degenerate_rows)
    # first consumer of the tokens
    token_embed_case("maskgit_token_embed", 512, 64, 3, 16, 50)
    token_grad_case("token_consumers_grad", 512, 64, 3, 16, 51)
    decode_grad_case("vit_decode_grad", "vit", 512, 32, 3, 16, 52)
    decode_grad_case("vqgan_decode_grad", "vqgan", 256, 64, 2, 16, 53)
    # pre_quant / post_quant around the quantiser
    _selected(projected_case)("vit_projected_step", 1024, 32, 128, 3, 96, 60)
    _selected(projected_conv_case)("vqgan_projected_decode", 512, 64, 2, 4, 61)
