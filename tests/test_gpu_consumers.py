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
