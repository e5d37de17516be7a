"""First consumer of the tokens: vq_token_embed against the oracle and the reference fixture."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "attention-models_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from oracle import vq_oracle as vo  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_token_embed_matches_reference_fixture(dev):
    from vq_b200.consumers import masked_token_embeddings
    g = np.load(os.path.join(GOLDEN, "maskgit_token_embed.npz"))
    vocab, dim, b, n, seed = (int(g[k]) for k in ("vocab", "dim", "b", "n", "seed"))
    table, pos, tokens, mask = vo.token_embed_inputs(vocab, dim, b, n, seed)
    e, ids, labels = masked_token_embeddings(tokens.to(dev), mask.to(dev), vocab, table.to(dev), pos.to(dev), -1)
    assert np.array_equal(ids.cpu().numpy(), g["input_ids"]) and np.array_equal(labels.cpu().numpy(), g["labels"])
    assert np.array_equal(e.cpu().numpy(), g["embeds"])


@pytest.mark.parametrize("b,n,vocab,dim", [(64, 256, 8192, 512), (3, 1024, 8192, 768), (1, 1, 16, 4), (512, 1024, 8192, 64)])
def test_token_embed_matches_oracle_and_consumes_the_quantisers_indices(dev, b, n, vocab, dim):
    """Sizes of MaskGIT / Muse (256 or 1024 tokens per image, dim 512 / 768) up to configs[3]'s 512 x 1024 tokens; the
    tokens are what the drop-in Codebook.encode returns; bit-exact against the oracle; no-mask and no-pos variants;
    out-of-range ids raise like nn.Embedding."""
    from vq_b200.consumers import masked_token_embeddings
    from vq_b200.vitvqgan import Codebook
    g = torch.Generator().manual_seed(5)
    table = torch.randn(vocab + 1, dim, generator=g).to(dev)
    pos = (0.02 * torch.randn(1, n, dim, generator=g)).to(dev)
    if vocab == 8192 and b * n >= 256:
        cb = Codebook(8192, 32).to(dev)
        tokens = cb.encode(torch.randn(b, n, 32, generator=g).to(dev)).view(b, n)
    else:
        tokens = torch.randint(0, vocab, (b, n), generator=g).to(dev)
    mask = (torch.rand(b, n, generator=g) < 0.55).to(dev)
    for mk, ps in ((mask, pos), (None, pos), (mask, None)):
        e, ids, labels = masked_token_embeddings(tokens, mk, vocab, table, ps, -1)
        re, rids, rlabels = vo.masked_token_embeddings(tokens, mk, vocab, table, ps, -1)
        assert torch.equal(ids, rids) and torch.equal(labels, rlabels) and torch.equal(e, re)
    bad = tokens.clone()
    bad[0, 0] = vocab + 1
    with pytest.raises(IndexError):
        masked_token_embeddings(bad, None, vocab, table, pos)
    with pytest.raises(RuntimeError, match="no CPU path"):
        masked_token_embeddings(tokens.cpu(), None, vocab, table, pos)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_consumer_gradients_match_reference_fixture(dev):
    """Training-side of the consumers: gradients of MaskGIT's lookup (table, trainable pos_enc) and of Parti's shifted
    decoder input (table, start token) against the reference's modules + autograd (tests/golden/token_consumers_grad.npz);
    the forward of the causal form bit-exact."""
    from vq_b200.consumers import causal_token_embeddings, masked_token_embeddings
    g = np.load(os.path.join(GOLDEN, "token_consumers_grad.npz"))
    vocab = int(g["vocab"])
    tokens, mask = torch.from_numpy(g["tokens"]).to(dev), torch.from_numpy(g["mask"]).to(dev)
    up = torch.from_numpy(g["upstream"]).to(dev)
    table = torch.from_numpy(g["maskgit_table"]).to(dev).requires_grad_(True)
    pos = torch.from_numpy(g["maskgit_pos"]).to(dev).requires_grad_(True)
    e, _, _ = masked_token_embeddings(tokens, mask, vocab, table, pos, -1)
    (e * up).sum().backward()
    assert np.array_equal(e.detach().cpu().numpy(), g["maskgit_embeds"])
    assert _rel(table.grad.cpu().numpy(), g["maskgit_grad_table"]) < 1e-6
    assert _rel(pos.grad.cpu().numpy(), g["maskgit_grad_pos"]) < 1e-6
    ptable = torch.from_numpy(g["parti_table"]).to(dev).requires_grad_(True)
    start = torch.from_numpy(g["parti_start"]).to(dev).requires_grad_(True)
    pe, labels = causal_token_embeddings(tokens, ptable, torch.from_numpy(g["parti_pe"]).to(dev), start)
    (pe * up).sum().backward()
    assert np.array_equal(pe.detach().cpu().numpy(), g["parti_embeds"]) and np.array_equal(labels.cpu().numpy(), g["parti_labels"])
    assert _rel(ptable.grad.cpu().numpy(), g["parti_grad_table"]) < 1e-6
    assert _rel(start.grad.cpu().numpy(), g["parti_grad_start"]) < 1e-6


@pytest.mark.parametrize("b,n,vocab,dim", [(16, 256, 8192, 512), (2, 1024, 8192, 64), (3, 7, 16, 4)])
def test_causal_embeddings_and_embedding_backward_match_oracle(dev, b, n, vocab, dim):
    """Parti-sized shapes against the oracle on the same device: forward bit-exact; the integer-accumulated backward within
    1e-6 of autograd's and bitwise repeatable (also under heavy id collisions: tiny vocabularies, tiny gradients)."""
    from vq_b200.consumers import causal_token_embeddings
    gen = torch.Generator().manual_seed(9)
    tokens = torch.randint(0, vocab, (b, n), generator=gen).to(dev)
    pe = torch.randn(n, dim, generator=gen).to(dev)
    up = (torch.randn(b, n, dim, generator=gen) * 1e-4).to(dev)
    grads = []
    for fn in (causal_token_embeddings, vo.causal_token_embeddings, causal_token_embeddings):
        table = torch.randn(vocab, dim, generator=torch.Generator().manual_seed(10)).to(dev).requires_grad_(True)
        start = torch.randn(dim, generator=torch.Generator().manual_seed(11)).to(dev).requires_grad_(True)
        e, labels = fn(tokens, table, pe, start)
        (e * up).sum().backward()
        grads.append((e.detach(), table.grad.clone(), start.grad.clone()))
        assert torch.equal(labels, tokens)
    assert torch.equal(grads[0][0], grads[1][0])
    assert _rel(grads[0][1].cpu().numpy(), grads[1][1].cpu().numpy()) < 1e-6
    assert _rel(grads[0][2].cpu().numpy(), grads[1][2].cpu().numpy()) < 1e-6
    assert torch.equal(grads[0][1], grads[2][1])                      # deterministic: integer sums


@pytest.mark.parametrize("name", ["vit_decode_grad", "vqgan_decode_grad"])
def test_decode_gradient_matches_reference_fixture(dev, name):
    """Codebook.indices_to_embeddings is differentiable w.r.t. the weights like the reference's (models/vitvqgan.py:173-176,
    models/vqgan.py:178-182): gradient of the drop-in modules against the unmodified reference + autograd."""
    from vq_b200 import vitvqgan, vqgan
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    form, K, D = str(g["form"]), int(g["K"]), int(g["D"])
    m = (vitvqgan.Codebook if form == "vit" else vqgan.Codebook)(K, D).to(dev)
    with torch.no_grad():
        m.embedding.weight.copy_(torch.from_numpy(g["weight"]))
    out = m.indices_to_embeddings(torch.from_numpy(g["indices"]).to(dev))
    assert out.requires_grad
    (out * torch.from_numpy(g["upstream"]).to(dev)).sum().backward()
    # (the fixture was minted on the CPU: the normalised form agrees within 2 ulp(1.0), the raw gather exactly)
    assert np.abs(out.detach().cpu().numpy() - g["out"]).max() <= (2 * 2.0 ** -23 if form == "vit" else 0.0)
    assert _rel(m.embedding.weight.grad.cpu().numpy(), g["grad_weight"]) < 1e-5
    with torch.no_grad():
        assert not m.indices_to_embeddings(torch.from_numpy(g["indices"]).to(dev)).requires_grad


@pytest.mark.parametrize("index_dtype", [torch.int32, torch.uint16])
@pytest.mark.parametrize("form", ["vit", "vqgan"])
def test_narrow_token_wire_format_end_to_end(dev, form, index_dtype):
    """encode -> tokens as int32 / uint16 (VQ_FLAG_IDX32 / IDX16) -> decode, masked and causal consumers reading them
    directly: everything equal to the int64 path (the reference's dtype), including the gradients behind the consumers.
    Codes up to 65535 must survive the uint16 format; the format is refused where it cannot hold the codes."""
    from vq_b200 import functional as F_vq
    from vq_b200.consumers import causal_token_embeddings, masked_token_embeddings
    g = torch.Generator().manual_seed(11)
    if form == "vit":
        K, D, shape = 65536, 32, (4, 256, 32)
    else:
        K, D, shape = 1024, 64, (4, 64, 16, 16)
    w = vo.make_codebook(form, K, D, 3).to(dev)
    z = vo.make_latents(shape, 4).to(dev)
    if form == "vit":
        z[0, :8] = w[K - 8:] * 1.5           # rows that land on the highest codes: 65528 .. 65535
    idx64 = F_vq.encode_indices(z, w, form)
    idx = F_vq.encode_indices(z, w, form, index_dtype=index_dtype)
    assert idx.dtype == index_dtype and idx64.dtype == torch.int64
    assert torch.equal(F_vq.convert_tokens(idx, torch.int64), idx64)
    assert torch.equal(F_vq.convert_tokens(idx64, index_dtype).view(torch.uint8), idx.view(torch.uint8))
    if form == "vit":
        assert int(idx64[:8].min()) >= K - 8
    b = shape[0]
    tok64, tok = idx64.view(b, -1), idx.view(b, -1)
    # decode
    wg = w.clone().requires_grad_(True)
    d64 = F_vq.indices_to_embeddings(tok64, wg, form)
    up = torch.randn(d64.shape, generator=g).to(dev)
    (d64 * up).sum().backward()
    g64 = wg.grad.clone()
    wg.grad = None
    dn = F_vq.indices_to_embeddings(tok, wg, form)
    (dn * up).sum().backward()
    assert torch.equal(dn, d64) and torch.equal(wg.grad, g64)
    # consumers
    n = tok.shape[1]
    table = torch.randn(K + 1, 64, generator=g).to(dev).requires_grad_(True)
    pos = (0.02 * torch.randn(n, 64, generator=g)).to(dev)
    mask = (torch.rand(b, n, generator=g) < 0.5).to(dev)
    outs = []
    for t in (tok64, tok):
        table.grad = None
        e, ids, labels = masked_token_embeddings(t, mask, K, table, pos, -100)
        e.square().sum().backward()
        outs.append((e.detach(), ids, labels, table.grad.clone()))
    for a, c in zip(*outs):
        assert torch.equal(a, c)
    assert outs[1][1].dtype == torch.int64 and outs[1][2].dtype == torch.int64       # what cross_entropy wants
    start = torch.randn(64, generator=g).to(dev)
    e64, _ = causal_token_embeddings(tok64, table.detach(), pos, start)
    en, lab = causal_token_embeddings(tok, table.detach(), pos, start)
    assert torch.equal(en, e64) and lab.dtype == index_dtype
    # a format that cannot hold the codes is refused; narrow indices only without z_q / backward state
    if index_dtype == torch.uint16:
        from vq_b200._lib import VQLibraryError
        big = vo.make_codebook("vit", 65536 + 512, 32, 5).to(dev)
        with pytest.raises(VQLibraryError, match="IDX16"):
            F_vq.encode_indices(vo.make_latents((1, 256, 32), 6).to(dev), big, "vit", index_dtype=torch.uint16)
    with pytest.raises(TypeError):
        F_vq.encode_indices(z, w, form, index_dtype=torch.int16)
