"""pre_quant / post_quant fused with the quantiser against the oracle on the same GPU
and against fixtures minted from the reference's own modules (tests/golden/vit_projected_step.npz,
vqgan_projected_decode.npz; oracle/make_golden.py::projected_case).

Tolerance decision (include/vq_b200.h, vq_forward_projected): the fused GEMM (3xTF32 on the tensor cores) sums in an order
of its own, like any two fp32 GEMMs.  So
  * z is compared with a float64 product: |z - z64| <= Z_TOL * sum_c |x_c W_dc| (+ the same for the bias), and its error is
    reported next to cuBLAS's own fp32 error on the same inputs;
  * GIVEN the kernel's z, every other output is bit-identical to vq_forward on that z (asserted);
  * against F.linear + quantise, an index may differ only on a row whose two best codes are closer than the GEMM's
    rounding (top-2 relative gap < GAP_TOL); such rows are counted and reported like near-ties, all others must agree and
    their z_q agree to a few ulps.
"""
import copy

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import vq_oracle as vo
from standins import ViTVQGANStandIn, VQGANStandIn

pytestmark = pytest.mark.gpu

Z_TOL = 2e-5      # of sum |x_c W_dc|: fp32 accumulation over C terms inside the tensor-core MMAs
GAP_TOL = 1e-4    # top-2 relative distance gap under which an index may follow the GEMM's rounding
GRAD_TOL = 2e-5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


def _case(T, C, K, seed, dev, D=32):
    w = vo.make_codebook(vo.VIT, K, D, seed).to(dev)
    w_pre, b_pre = (t.to(dev) for t in vo.projection_inputs(C, D, seed + 1))
    x = vo.make_latents((T, C), seed + 3).to(dev)
    return x, w_pre, b_pre, w


def _z_error(z, x, w_pre, b_pre):
    """max over elements of |z - z64| / (sum_c |x_c W_dc| + |b_d|)"""
    z64 = x.double() @ w_pre.double().t() + (0 if b_pre is None else b_pre.double())
    scale = x.abs().double() @ w_pre.abs().double().t() + (0 if b_pre is None else b_pre.abs().double())
    return float(((z.double() - z64).abs() / scale.clamp_min(1e-30)).max())


@pytest.mark.parametrize("T,C,K", [(288, 128, 1024), (4096, 512, 8192), (1000, 64, 512), (33, 768, 8192), (1, 256, 512),
                                   (20000, 512, 8192)])
@pytest.mark.parametrize("bias", [True, False])
def test_fused_pre_quant_forward(dev, T, C, K, bias):
    from vq_b200 import functional as F_vq, projected
    x, w_pre, b_pre, w = _case(T, C, K, 70 + T % 7, dev)
    if not bias:
        b_pre = None
    assert projected.prequant_supported(C, 32)
    z_q, idx, loss, hist, stats, z = projected.quantise_projected(x, w_pre, b_pre, w, 0.25, return_z=True)
    assert z.shape == (T, 32) and z_q.shape == (T, 32) and idx.shape == (T,) and idx.dtype == torch.int64

    # 1. the GEMM: fp32-GEMM-level rounding
    err = _z_error(z, x, w_pre, b_pre)
    err_cublas = _z_error(torch.nn.functional.linear(x, w_pre, b_pre), x, w_pre, b_pre)
    print(f"\nT={T} C={C}: max |z - z64| / sum|terms| = {err:.2e} (cuBLAS fp32 on the same inputs: {err_cublas:.2e})")
    assert err <= Z_TOL

    # 2. given the kernel's z, everything behind it is vq_forward's, bit for bit
    z_q2, idx2, loss2, hist2, _ = F_vq.quantise(z, w, "vit", 0.25)
    assert torch.equal(idx, idx2) and torch.equal(z_q, z_q2) and torch.equal(hist, hist2)
    assert torch.equal(loss, loss2)
    assert int(hist.sum()) == T

    # 3. against the reference's composition (oracle on the same GPU): indices may move only on near-tie rows
    z_ref, o = vo.quantise_projected(x, w_pre, b_pre, w, 0.25)
    bad = idx != o.indices.reshape(-1)
    n_bad = int(bad.sum())
    if n_bad:
        gap = vo.top2_relative_gap(vo.unit_rows(z_ref)[bad], vo.unit_rows(w))
        print(f"T={T} C={C}: {n_bad} rows follow the GEMM's rounding (max top-2 gap {float(gap.max()):.2e})")
        assert float(gap.max()) < GAP_TOL
    assert n_bad <= max(2, T // 2000)
    assert rel_err(z_q[~bad].cpu().numpy(), o.z_q.reshape(T, 32)[~bad].cpu().numpy()) < 1e-5
    if n_bad == 0:
        assert rel_err(loss.cpu().numpy(), o.loss.cpu().numpy()) < 1e-5

    # 4. the indices-only path and the narrow token formats give the same tokens
    for dt in (torch.int64, torch.int32, torch.uint16):
        t = projected.encode_indices_projected(x, w_pre, b_pre, w, index_dtype=dt)
        assert t.dtype == dt and torch.equal(t.to(torch.int64), idx)


def test_fused_pre_quant_exhaustive_scan_agrees(dev):
    from vq_b200 import projected
    x, w_pre, b_pre, w = _case(8192, 512, 8192, 91, dev)
    a = projected.quantise_projected(x, w_pre, b_pre, w, 0.25)
    b = projected.quantise_projected(x, w_pre, b_pre, w, 0.25, exact_scan=True)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])


def test_fused_pre_quant_unsupported_shapes(dev):
    from vq_b200 import projected
    _, _, _, w = _case(1, 128, 512, 5, dev)
    assert not projected.prequant_supported(96, 32) and not projected.prequant_supported(1024, 32)
    assert not projected.prequant_supported(512, 64)
    with pytest.raises(ValueError):
        projected.quantise_projected(torch.zeros(4, 96, device=dev), torch.zeros(32, 96, device=dev), None, w, 0.25)


def test_fused_pre_quant_gradients_match_autograd_of_the_composition(dev):
    from vq_b200 import projected
    T, C, K = 4096, 256, 2048
    x, w_pre, b_pre, w = _case(T, C, K, 33, dev)
    up = vo.make_latents((T, 32), 37).to(dev)
    leaves_ref = [t.clone().requires_grad_(True) for t in (x, w_pre, b_pre, w)]
    _, o = vo.quantise_projected(*leaves_ref, 0.25)
    ((o.z_q * up).sum() + 3.0 * o.loss).backward()
    leaves = [t.clone().requires_grad_(True) for t in (x, w_pre, b_pre, w)]
    z_q, idx, loss, hist, stats = projected.quantise_projected(*leaves, 0.25)
    ((z_q * up).sum() + 3.0 * loss).backward()
    if not torch.equal(idx, o.indices.reshape(-1)):
        pytest.skip("seed produced a row that follows the GEMM's rounding; covered by test_fused_pre_quant_forward")
    for name, a, b in zip(("x", "w_pre", "b_pre", "weight"), leaves, leaves_ref):
        assert a.grad is not None, name
        assert rel_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < GRAD_TOL, name
    # only some inputs need gradients
    xg = x.clone().requires_grad_(True)
    z_q, _, loss, _, _ = projected.quantise_projected(xg, w_pre, None, w, 0.25)
    (z_q.sum() + loss).backward()
    assert xg.grad is not None and xg.grad.shape == x.shape


def test_fused_pre_quant_matches_reference_fixture(dev):
    from vq_b200 import projected
    g = load_golden("vit_projected_step")
    K, D, C, b, n, seed = (int(g[k]) for k in ("K", "D", "C", "b", "n", "seed"))
    w = vo.make_codebook(vo.VIT, K, D, seed).to(dev).requires_grad_(True)
    w_pre, b_pre = (t.to(dev).requires_grad_(True) for t in vo.projection_inputs(C, D, seed + 1))
    x = vo.make_latents((b, n, C), seed + 3).to(dev).requires_grad_(True)
    up = vo.make_latents((b, n, D), seed + 4).to(dev)
    z_q, idx, loss, hist, stats, z = projected.quantise_projected(x, w_pre, b_pre, w, float(g["beta"]), return_z=True)
    assert z_q.shape == (b, n, D) and z.shape == (b, n, D)
    assert rel_err(z.detach().cpu().numpy(), g["z"]) < 2e-6
    ref_idx = torch.from_numpy(g["indices"].astype(np.int64)).reshape(-1)
    bad = idx.cpu() != ref_idx
    if int(bad.sum()):
        zn = vo.unit_rows(torch.from_numpy(g["z"]).reshape(-1, D))
        assert float(vo.top2_relative_gap(zn[bad], vo.unit_rows(w.detach().cpu())).max()) < GAP_TOL
    assert int(bad.sum()) <= 1
    assert rel_err(z_q.detach().cpu().numpy().reshape(-1, D)[~bad.numpy()], g["z_q"].reshape(-1, D)[~bad.numpy()]) < 1e-6
    if int(bad.sum()) == 0:
        assert rel_err(loss.detach().cpu().numpy(), g["loss"]) < 1e-5
        ((z_q * up).sum() + loss).backward()
        for name, t in (("grad_x", x), ("grad_w_pre", w_pre), ("grad_b_pre", b_pre), ("grad_weight", w)):
            assert rel_err(t.grad.cpu().numpy(), g[name]) < GRAD_TOL, name


@pytest.mark.parametrize("form", ["vit", "vqgan"])
def test_projected_decode_matches_reference_fixture_and_oracle(dev, form):
    from vq_b200 import projected
    if form == "vit":
        g = load_golden("vit_projected_step")
        K, D, C, seed = (int(g[k]) for k in ("K", "D", "C", "seed"))
        w_post, b_post = (t.to(dev) for t in vo.projection_inputs(D, C, seed + 2))
        shape = (int(g["b"]), int(g["n"]), C)
    else:
        g = load_golden("vqgan_projected_decode")
        K, D, seed = (int(g[k]) for k in ("K", "D", "seed"))
        C = D
        w_post, b_post = (t.to(dev) for t in vo.projection_inputs(D, D, seed + 2, conv=True))
        shape = (int(g["b"]), D, int(g["side"]), int(g["side"]))
    w = vo.make_codebook(form, K, D, seed).to(dev)
    table = projected.ProjectedTable(w, w_post, b_post, form)
    assert table.table.shape == (K, C)
    for dt in (torch.int64, torch.int32, torch.uint16):
        tokens = torch.from_numpy(g["tokens"].astype(np.int64)).to(dev).to(dt)
        out = table.gather(tokens)
        assert tuple(out.shape) == shape
        assert rel_err(out.cpu().numpy(), g["decoded"]) < 1e-6
        ref = vo.decode_projected(form, tokens.to(torch.int64), w, w_post, b_post)
        assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < 1e-6
    with pytest.raises(IndexError):
        table.gather(torch.full((1, 4), K, dtype=torch.int64, device=dev))
    # no bias; larger table (cfg 1 shapes: K = 8192, C = 512 / the CNN form's 256 x 256 conv)
    K2, D2, C2 = (8192, 32, 512) if form == "vit" else (8192, 256, 256)
    w2 = vo.make_codebook(form, K2, D2, 3).to(dev)
    wp2, _ = vo.projection_inputs(D2, C2, 4, conv=(form == "vqgan"))
    wp2 = wp2.to(dev)
    tok2 = torch.randint(0, K2, (4, 256), device=dev)
    out2 = projected.ProjectedTable(w2, wp2, None, form).gather(tok2)
    ref2 = vo.decode_projected(form, tok2, w2, wp2, None)
    assert out2.shape == ref2.shape and rel_err(out2.cpu().numpy(), ref2.cpu().numpy()) < 1e-6


@pytest.mark.parametrize("form", ["vit", "vqgan"])
def test_wrapper_call_sites_with_fused_projections(dev, form):
    """models/vitvqgan.py:190-210 / models/vqgan.py:239-243 around the patched codebook with fuse_projections=True."""
    from vq_b200.integration import patch_reference_model
    torch.manual_seed(23)
    if form == "vit":
        ref = ViTVQGANStandIn(patch_dim=48, dim=128, codebook_size=4096, codebook_dim=32).to(dev)
        ref.codebook.embedding.weight.data.normal_()
        imgs = torch.randn(5, 200, 48, device=dev)
    else:
        ref = VQGANStandIn(in_ch=3, dim=64, codebook_size=1024).to(dev)
        ref.codebook.embedding.weight.data.uniform_(-1.0 / 1024, 1.0 / 1024)
        imgs = torch.randn(4, 3, 32, 32, device=dev)
    new = patch_reference_model(copy.deepcopy(ref), form=form, fuse_projections=True)
    assert list(new.state_dict().keys()) == list(ref.state_dict().keys())
    with torch.no_grad():
        t_ref, t_new = ref.encode_imgs(imgs), new.encode_imgs(imgs)
        assert t_new.shape == t_ref.shape and t_new.dtype == torch.int64
        same = t_ref == t_new
        if form == "vqgan":
            assert bool(same.all())
        else:
            z = ref.pre_quant(ref.encoder(imgs)).reshape(-1, 32)
            if not bool(same.all()):
                gap = vo.top2_relative_gap(vo.unit_rows(z)[~same.reshape(-1)], vo.unit_rows(ref.codebook.embedding.weight))
                assert float(gap.max()) < GAP_TOL and int((~same).sum()) <= 2
        # decode_indices: one gather from the projected codes, then the decoder
        assert rel_err(new.decode_indices(t_ref).cpu().numpy(), ref.decode_indices(t_ref).cpu().numpy()) < 1e-5
    if form == "vit" and bool(same.all()):
        with torch.no_grad():       # evaluation forward: post_quant comes from the projected table
            out_ref, loss_ref = ref(imgs)
            out_new, loss_new = new(imgs)
        assert rel_err(out_new.cpu().numpy(), out_ref.cpu().numpy()) < 1e-5
        assert rel_err(loss_new.cpu().numpy(), loss_ref.cpu().numpy()) < 1e-5
        target = torch.randn_like(imgs)
        outs = []
        for m in (ref, new):
            m.zero_grad(set_to_none=True)
            out, loss = m(imgs)
            ((out - target).square().mean() + loss).backward()
            outs.append((out.detach(), loss.detach()))
        assert rel_err(outs[1][0].cpu().numpy(), outs[0][0].cpu().numpy()) < 1e-5
        assert rel_err(outs[1][1].cpu().numpy(), outs[0][1].cpu().numpy()) < 1e-5
        for (name, p_ref), (_, p_new) in zip(ref.named_parameters(), new.named_parameters()):
            assert p_new.grad is not None, name
            assert rel_err(p_new.grad.cpu().numpy(), p_ref.grad.cpu().numpy()) < 5e-5, name
    # with gradients enabled and trainable weights decode_indices keeps the differentiable composition
    new.zero_grad(set_to_none=True)
    new.decode_indices(t_ref).sum().backward()
    assert new.codebook.embedding.weight.grad is not None


def test_full_size_cfg3_fused_pre_quant(dev):
    """BASELINE cfg3 behind the ViT encoder's 512 features: 262 144 rows x 512 -> 32, K = 8192."""
    from vq_b200 import functional as F_vq, projected
    T, C, K = 262144, 512, 8192
    x, w_pre, b_pre, w = _case(T, C, K, 101, dev)
    prepared = F_vq.prepare_codebook(w)
    idx = projected.encode_indices_projected(x, w_pre, b_pre, w, prepared=prepared)
    z_ref = torch.nn.functional.linear(x, w_pre, b_pre)
    idx_ref = F_vq.encode_indices(z_ref, w, "vit", prepared=prepared)
    bad = idx != idx_ref
    n_bad = int(bad.sum())
    print(f"\ncfg3 + pre_quant(512 -> 32): {n_bad} of {T} rows follow the GEMM's rounding")
    if n_bad:
        gap = vo.top2_relative_gap(vo.unit_rows(z_ref)[bad], vo.unit_rows(w))
        assert float(gap.max()) < GAP_TOL
    assert n_bad <= T // 2000
    hist = torch.bincount(idx, minlength=K)
    assert int(hist.sum()) == T


@pytest.mark.parametrize("form", ["vit", "vqgan"])
def test_empty_batch_matches_the_reference_convention(dev, form):
    """No tokens: empty z_q / indices, NaN loss like the reference's ``torch.mean`` of nothing; nothing is launched."""
    from vq_b200 import projected, vitvqgan, vqgan
    if form == "vit":
        m, z = vitvqgan.Codebook(512, 32).to(dev), torch.empty(0, 4, 32, device=dev)
    else:
        m, z = vqgan.Codebook(512, 64).to(dev), torch.empty(0, 64, 2, 2, device=dev)
    o = vo.quantise(form, z.cpu(), m.embedding.weight.detach().cpu(), m.beta)
    z_q, idx, loss = m(z)
    assert z_q.shape == o.z_q.shape and idx.shape == o.indices.shape and idx.dtype == torch.int64
    assert bool(torch.isnan(loss)) and bool(torch.isnan(o.loss))
    assert m.encode(z).numel() == 0 and int(m.last_histogram.sum()) == 0
    if form == "vit":
        x, w_pre, b_pre, w = _case(0, 128, 512, 5, dev)
        z_q, idx, loss, hist, stats = projected.quantise_projected(x, w_pre, b_pre, w, 0.25)
        assert z_q.shape == (0, 32) and idx.numel() == 0 and bool(torch.isnan(loss)) and int(hist.sum()) == 0
        assert projected.encode_indices_projected(x, w_pre, b_pre, w).numel() == 0
