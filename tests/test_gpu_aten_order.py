"""Confirm on the GPU that ATen's CUDA reductions add in the order oracle/aten_order.py restates, and that
the prep kernels reproduce torch's F.normalize bit for bit (SURVEY.md Appendix B caveat: the schedules
were read from headers in a GPU-less container and must be confirmed empirically)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import aten_order as ao

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.mark.parametrize("D", [16, 32, 64, 128, 256, 512])
def test_contiguous_norm_and_sum_order(dev, D):
    g = torch.Generator().manual_seed(D)
    x = torch.randn(4096, D, generator=g)
    xd = x.to(dev)
    norm_gpu = xd.norm(p=2, dim=-1).cpu().numpy()
    emu = np.sqrt(ao.rowsum_contiguous(x.numpy(), fused=True)).astype(np.float32)
    assert np.array_equal(norm_gpu, emu), f"norm order differs on {(norm_gpu != emu).mean():.3%} of rows"
    sq_gpu = torch.sum(xd ** 2, dim=1).cpu().numpy()
    emu_sq = ao.rowsum_contiguous((x.numpy() ** 2).astype(np.float32), fused=False)
    assert np.array_equal(sq_gpu, emu_sq), f"sum order differs on {(sq_gpu != emu_sq).mean():.3%} of rows"
    zn_gpu = F.normalize(xd, p=2, dim=-1).cpu().numpy()
    zn_emu, _ = ao.normalise_contiguous(x.numpy())
    assert np.array_equal(zn_gpu, zn_emu)


@pytest.mark.parametrize("b,D,h,w", [(4, 256, 16, 16), (2, 64, 8, 8), (3, 32, 4, 4), (2, 128, 8, 8), (3, 64, 4, 4),
                                     (1, 256, 8, 8), (1, 256, 4, 4), (2, 256, 5, 6), (2, 256, 5, 5), (1, 64, 3, 3),
                                     (5, 512, 2, 2), (2, 16, 8, 8)])
def test_channel_strided_norm_order(dev, b, D, h, w):
    g = torch.Generator().manual_seed(D + h)
    x = torch.randn(b, D, h, w, generator=g)
    view = x.to(dev).permute(0, 2, 3, 1)                     # what models/vqgan.py:151 normalises
    norm_gpu = view.norm(p=2, dim=-1).reshape(-1).cpu().numpy()
    emu = np.sqrt(ao.rowsum_channel_strided(x.numpy().reshape(b, D, h * w), fused=True)).astype(np.float32)
    assert np.array_equal(norm_gpu, emu), f"strided norm order differs on {(norm_gpu != emu).mean():.3%} of rows"
    zn_gpu = F.normalize(view, p=2, dim=-1).reshape(-1, D).cpu().numpy()
    zn_emu, _ = ao.normalise_nchw(x.numpy().reshape(b, D, h * w))
    assert np.array_equal(zn_gpu, zn_emu)


@pytest.mark.parametrize("D", [16, 32, 64, 128, 256, 512])
def test_prepared_codebook_bits(dev, D):
    """vq_codebook_prepare's unit codes == F.normalize(weight) on the same GPU, bit for bit."""
    import vq_b200
    g = torch.Generator().manual_seed(100 + D)
    w = torch.randn(1000, D, generator=g).to(dev)
    prepared = vq_b200.prepare_codebook(w)
    en = prepared.blob[:1000 * D * 4].view(torch.float32).view(1000, D)
    assert torch.equal(en, F.normalize(w, p=2, dim=-1))
