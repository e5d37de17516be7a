"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Primary oracle = the restated reference run by PyTorch **on the same GPU** (fp32, tf32 off, no autocast):
that is what the reference itself computes on a B200.  Secondary: the golden
fixtures written by the unmodified reference on CPU, and a float64 closed form for the gradients.

Bars (north_star): indices and z_q bit-exact except near-tie rows (top-2 fp32 distances < 1e-6 relative
apart), which are counted and reported; loss and gradients within 1e-5 relative.
"""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, rel_err, ulp_distance
from oracle import vq_oracle as vo

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-5      # north_star: "losses and gradients fall within 1e-5 relative"
GRAD_TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


def _module(form, K, D, beta, w, dev, exact):
    from vq_b200.vitvqgan import Codebook as Vit
    from vq_b200.vqgan import Codebook as Vqgan
    m = (Vit if form == "vit" else Vqgan)(K, D, beta).to(dev)
    with torch.no_grad():
        m.embedding.weight.copy_(w)
    m.exact_scan = exact
    return m


def _tok(form, t, D):
    return t.permute(0, 2, 3, 1).reshape(-1, D) if form == "vqgan" else t.reshape(-1, D)


def _check_forward(form, m, z, w, ref, report):
    """ref: oracle output on the same device.  Returns mask of rows whose index differs (near ties)."""
    D = w.shape[1]
    z_q, idx, loss = m(z)
    assert idx.dtype == torch.int64 and z_q.dtype == torch.float32 and loss.dim() == 0
    assert z_q.shape == ref.z_q.shape and idx.shape == ref.indices.shape
    zn = vo.unit_rows(_tok(form, z, D))
    rep = vo.classify_index_mismatches(idx.reshape(-1), ref.indices.reshape(-1), zn, vo.unit_rows(w))
    assert rep["hard_rows"] == 0, rep
    bad = (idx.reshape(-1) != ref.indices.reshape(-1))
    a, b = _tok(form, z_q.detach(), D), _tok(form, ref.z_q, D)
    exact_rows = (a[~bad] == b[~bad]).all(dim=1)
    report.update(rows=int(idx.numel()), mismatch_rows=rep["mismatch_rows"], near_tie_rows=rep["near_tie_rows"],
                  zq_bit_exact_fraction=float(exact_rows.float().mean()) if exact_rows.numel() else 1.0,
                  reported_near_ties=m.near_tie_rows())
    assert bool(exact_rows.all()), f"z_q not bit-exact on {int((~exact_rows).sum())} rows"
    assert rel_err(loss.detach().cpu().numpy(), ref.loss.cpu().numpy()) < LOSS_TOL
    hist = m.last_histogram
    assert torch.equal(hist.long(), torch.bincount(idx.reshape(-1), minlength=w.shape[0]))
    return z_q, idx, loss, bad


CASES = [
    # form, K, D, shape, w_seed, z_seed   (BASELINE.json configs[0], a slice of configs[1], plus odd sizes)
    ("vit", 8192, 32, (2, 1024, 32), 0, 1),
    ("vit", 1000, 32, (3, 77, 32), 3, 4),          # K not a multiple of any tile, ragged T
    ("vit", 1024, 64, (2, 130, 64), 5, 6),
    ("vit", 512, 128, (1, 257, 128), 7, 8),
    ("vit", 8192, 256, (2, 512, 256), 9, 10),
    ("vqgan", 8192, 256, (4, 256, 16, 16), 0, 2),
    ("vqgan", 1024, 256, (2, 256, 8, 8), 11, 12),
    ("vqgan", 300, 64, (3, 64, 4, 4), 13, 14),
    ("vqgan", 256, 32, (2, 32, 6, 6), 15, 16),
    ("vqgan", 512, 256, (1, 256, 8, 8), 17, 18),      # few tokens: ATen reshapes its reduce block
    ("vqgan", 512, 256, (2, 256, 5, 5), 19, 20),      # odd h*w: scalar (non-vectorised) ATen schedule
    ("vqgan", 512, 64, (2, 64, 3, 6), 21, 22),        # h*w % 4 == 2
]


@pytest.mark.parametrize("exact", [True, False], ids=["simt", "auto"])
@pytest.mark.parametrize("form,K,D,shape,ws,zs", CASES)
def test_forward_matches_reference_on_gpu(dev, form, K, D, shape, ws, zs, exact):
    w = vo.make_codebook(form, K, D, ws).to(dev)
    z = vo.make_latents(shape, zs).to(dev)
    m = _module(form, K, D, 0.25, w, dev, exact)
    with torch.no_grad():
        ref = vo.quantise(form, z, w, 0.25)
        rep = {}
        _check_forward(form, m, z, w, ref, rep)
    print("parity", form, K, D, shape, "simt" if exact else "auto", rep)


@pytest.mark.parametrize("exact", [True, False], ids=["simt", "auto"])
@pytest.mark.parametrize("name", ["vit_cfg1_fwd", "vqgan_cfg2_slice_fwd"])
def test_forward_matches_golden_fixture(dev, name, exact):
    """Against what the unmodified reference produced on CPU.  CPU and GPU ATen sum in different orders,
    so z_q is compared within 2 ulp here (bit-exactness is asserted against the same-device oracle)."""
    g = load_golden(name)
    form, K, D = str(g["form"]), int(g["K"]), int(g["D"])
    w = vo.make_codebook(form, K, D, int(g["w_seed"]))
    z = vo.make_latents(tuple(int(s) for s in g["shape"]), int(g["z_seed"]))
    m = _module(form, K, D, float(g["beta"]), w.to(dev), dev, exact)
    with torch.no_grad():
        z_q, idx, loss = m(z.to(dev))
    ref_idx = torch.from_numpy(g["indices"].astype(np.int64)).reshape(-1)
    zn = vo.unit_rows(_tok(form, z, D))
    rep = vo.classify_index_mismatches(idx.cpu().reshape(-1), ref_idx, zn, vo.unit_rows(w))
    assert rep["hard_rows"] == 0, rep
    ok = (idx.cpu().reshape(-1) == ref_idx).numpy()
    a = _tok(form, z_q.cpu(), D).numpy()
    b = _tok(form, torch.from_numpy(g["z_q"]), D).numpy()
    # elements are <= 1 in magnitude and z_q = zn + (q - zn) rounds at the magnitude of zn, so the honest
    # bound is absolute: 2 ulp of 1.0 (ulp counts explode on near-zero elements)
    assert np.abs(a[ok] - b[ok]).max() <= 2 * 2.0 ** -23
    assert rel_err(loss.cpu().numpy(), g["loss"]) < LOSS_TOL
    print("golden", name, rep)


STEP_CASES = [
    ("vit", 1024, 32, (4, 256, 32), 10, 11, 12, 0.25),
    ("vit", 512, 32, (2, 128, 32), 13, 14, 15, 0.7),
    ("vit", 8192, 32, (8, 1024, 32), 0, 3, 4, 0.25),
    ("vit", 2048, 256, (2, 300, 256), 17, 18, 19, 0.25),
    ("vqgan", 512, 256, (2, 256, 8, 8), 20, 21, 22, 0.25),
    ("vqgan", 256, 64, (3, 64, 4, 4), 23, 24, 25, 0.4),
    ("vqgan", 1024, 256, (4, 256, 16, 16), 26, 27, 28, 0.25),
]


@pytest.mark.parametrize("exact", [True, False], ids=["simt", "auto"])
@pytest.mark.parametrize("form,K,D,shape,ws,zs,gs,beta", STEP_CASES)
def test_step_gradients_match_reference_on_gpu(dev, form, K, D, shape, ws, zs, gs, beta, exact):
    w = vo.make_codebook(form, K, D, ws).to(dev)
    z = vo.make_latents(shape, zs).to(dev)
    up = vo.make_latents(shape, gs).to(dev)
    m = _module(form, K, D, beta, w, dev, exact)
    zz = z.clone().requires_grad_(True)
    z_q, idx, loss = m(zz)
    ((z_q * up).sum() + loss).backward()
    ref = vo.quantise_step(form, z, w, beta, up)
    bad = idx.reshape(-1) != ref.indices.reshape(-1)
    rep = vo.classify_index_mismatches(idx.reshape(-1), ref.indices.reshape(-1), vo.unit_rows(_tok(form, z, D)),
                                       vo.unit_rows(w))
    assert rep["hard_rows"] == 0, rep
    # closed form in float64 at OUR indices: independent of near-tie flips
    gz64, gw64 = vo.analytic_backward(form, _tok(form, z, D).double(), w.double(), idx.reshape(-1), beta,
                                      _tok(form, up, D).double())
    assert rel_err(_tok(form, zz.grad, D).cpu().numpy(), gz64.cpu().numpy()) < GRAD_TOL
    assert rel_err(m.embedding.weight.grad.cpu().numpy(), gw64.cpu().numpy()) < GRAD_TOL
    if not bool(bad.any()):
        assert torch.equal(z_q.detach(), ref.z_q)
        assert rel_err(zz.grad.cpu().numpy(), ref.grad_z.cpu().numpy()) < GRAD_TOL
        assert rel_err(m.embedding.weight.grad.cpu().numpy(), ref.grad_weight.cpu().numpy()) < GRAD_TOL
    assert rel_err(loss.detach().cpu().numpy(), ref.loss.cpu().numpy()) < LOSS_TOL
    # rows of unused codes are exactly zero, like a dense embedding gradient
    unused = m.last_histogram == 0
    assert bool((m.embedding.weight.grad[unused] == 0).all())


@pytest.mark.parametrize("name", ["vit_step_small", "vit_step_beta", "vqgan_step_small", "vqgan_step_d64"])
def test_step_matches_golden_fixture(dev, name):
    g = load_golden(name)
    form, K, D = str(g["form"]), int(g["K"]), int(g["D"])
    shape = tuple(int(s) for s in g["shape"])
    w = vo.make_codebook(form, K, D, int(g["w_seed"])).to(dev)
    z = vo.make_latents(shape, int(g["z_seed"])).to(dev).requires_grad_(True)
    up = vo.make_latents(shape, int(g["g_seed"])).to(dev)
    m = _module(form, K, D, float(g["beta"]), w, dev, False)
    z_q, idx, loss = m(z)
    ((z_q * up).sum() + loss).backward()
    if np.array_equal(idx.cpu().numpy().reshape(-1), g["indices"].astype(np.int64).reshape(-1)):
        assert rel_err(z.grad.cpu().numpy(), g["grad_z"]) < GRAD_TOL
        assert rel_err(m.embedding.weight.grad.cpu().numpy(), g["grad_weight"]) < GRAD_TOL
    assert rel_err(loss.detach().cpu().numpy(), g["loss"]) < LOSS_TOL


def test_only_loss_backward_and_only_zq_backward(dev):
    """loss.backward() alone (no upstream through z_q) and z_q.sum().backward() alone."""
    w = vo.make_codebook("vit", 512, 32, 1).to(dev)
    z = vo.make_latents((2, 100, 32), 2).to(dev)
    for which in ("loss", "zq"):
        m = _module("vit", 512, 32, 0.25, w, dev, False)
        zz = z.clone().requires_grad_(True)
        z_q, idx, loss = m(zz)
        (loss if which == "loss" else z_q.sum()).backward()
        w_ref = w.clone().requires_grad_(True)
        z_ref = z.clone().requires_grad_(True)
        o = vo.quantise("vit", z_ref, w_ref, 0.25)
        (o.loss if which == "loss" else o.z_q.sum()).backward()
        assert torch.equal(idx, o.indices)
        assert rel_err(zz.grad.cpu().numpy(), z_ref.grad.cpu().numpy()) < GRAD_TOL
        if which == "loss":
            assert rel_err(m.embedding.weight.grad.cpu().numpy(), w_ref.grad.cpu().numpy()) < GRAD_TOL
        else:
            # no path from z_q to the codebook (the reference leaves weight.grad None; here it is all zeros)
            assert w_ref.grad is None and float(m.embedding.weight.grad.abs().max()) == 0.0


@pytest.mark.parametrize("form,name", [("vit", "vit_decode"), ("vqgan", "vqgan_decode")])
def test_indices_to_embeddings(dev, form, name):
    g = load_golden(name)
    K, D = int(g["K"]), int(g["D"])
    w = vo.make_codebook(form, K, D, int(g["w_seed"])).to(dev)
    gen = torch.Generator().manual_seed(int(g["i_seed"]))
    idx = torch.randint(0, K, (int(g["b"]), int(g["n"])), generator=gen).to(dev)
    m = _module(form, K, D, 0.25, w, dev, False)
    e = m.indices_to_embeddings(idx)
    ref = vo.indices_to_embeddings(form, idx, w)
    assert e.shape == ref.shape and torch.equal(e, ref.contiguous())      # bit-exact vs the same-device oracle
    assert np.abs(e.detach().cpu().numpy() - g["embeds"]).max() <= 2 * 2.0 ** -23   # and within 2 ulp(1.0) of the CPU reference
    bad = idx.clone()
    bad[0, 0] = K
    with pytest.raises(IndexError):
        m.indices_to_embeddings(bad)


def test_round_trip_decode_of_encode(dev):
    """MaskGIT/Muse tokenisation (BASELINE.json configs[3]): decode(encode(z)) is the unit code of each
    token, equal to the forward's quantised latent up to the STE rounding (<= 1 ulp of 1.0)."""
    w = vo.make_codebook("vit", 8192, 32, 0).to(dev)
    z = vo.make_latents((16, 1024, 32), 5).to(dev)
    m = _module("vit", 8192, 32, 0.25, w, dev, False)
    with torch.no_grad():
        codes = m.encode(z)
        z_q, idx, _ = m(z)
        dec = m.indices_to_embeddings(codes)
    assert torch.equal(codes, idx)
    assert float((dec - z_q).abs().max()) <= 2 ** -23
    # idempotence: quantising the decoded latents returns the same codes (except exact ties)
    with torch.no_grad():
        again = m.encode(dec)
    assert float((again != codes).float().mean()) < 1e-4


def test_degenerate_rows(dev):
    """zero row, NaN row (index 0, NaN loss), row equal to a code, row opposite to a code, zero code."""
    K, D = 64, 32
    w = vo.make_codebook("vit", K, D, 40)
    z = vo.make_latents((2, 8, 32), 41)
    z[0, 0] = 0.0
    z[0, 1, 3] = float("nan")
    z[0, 2] = w[5] * 3.0
    z[1, 0] = -w[7]
    w, z = w.to(dev), z.to(dev)
    for exact in (True, False):
        m = _module("vit", K, D, 0.25, w, dev, exact)
        with torch.no_grad():
            z_q, idx, loss = m(z)
            ref = vo.quantise("vit", z, w, 0.25)
        got, want = idx.clone(), ref.indices.clone()
        got[0, 0] = want[0, 0]           # all distances of the zero row are equal up to 1 ulp: a pure tie
        assert torch.equal(got, want)
        assert int(idx[0, 1]) == 0 and int(idx[0, 2]) == 5
        assert bool(torch.isnan(loss)) and bool(torch.isnan(ref.loss))
        same_nan = torch.isnan(z_q) == torch.isnan(ref.z_q)
        assert bool(same_nan.all())
        keep = torch.ones_like(idx, dtype=torch.bool)
        keep[0, 0] = False
        assert torch.equal(torch.nan_to_num(z_q[keep]), torch.nan_to_num(ref.z_q[keep]))
    # a zero code: its |en|^2 is 0, so it wins whenever no code has a positive dot product > 1/2
    w2 = w.clone()
    w2[3] = 0.0
    for exact in (True, False):
        m = _module("vit", K, D, 0.25, w2, dev, exact)
        with torch.no_grad():
            _, idx, _ = m(z[1:])
            ref = vo.quantise("vit", z[1:], w2, 0.25)
        assert torch.equal(idx, ref.indices)


def test_modes_give_identical_outputs(dev):
    """no_grad / eval / requires_grad_(False) / bf16 input (autocast hand-over) -- SURVEY.md section 8b."""
    w = vo.make_codebook("vit", 1024, 32, 1).to(dev)
    z = vo.make_latents((2, 256, 32), 2).to(dev)
    m = _module("vit", 1024, 32, 0.25, w, dev, False)
    a = m(z)
    with torch.no_grad():
        b = m(z)
    m.eval()
    m.requires_grad_(False)
    c = m(z)
    for x in (b, c):
        assert torch.equal(a[0].detach(), x[0]) and torch.equal(a[1], x[1]) and torch.equal(a[2].detach(), x[2])
    zb = z.bfloat16()
    with torch.no_grad():
        d = m(zb)
        ref = vo.quantise("vit", zb.float(), w, 0.25)
    assert d[0].dtype == torch.float32 and torch.equal(d[1], ref.indices)


def test_repeatable_bitwise(dev):
    """Two runs on the same inputs give bit-identical outputs and gradients (no float atomics)."""
    w = vo.make_codebook("vit", 8192, 32, 0).to(dev)
    z = vo.make_trained_like(w.cpu(), 50000, 7).to(dev).view(50, 1000, 32)
    up = vo.make_latents((50, 1000, 32), 8).to(dev)
    outs = []
    for _ in range(2):
        m = _module("vit", 8192, 32, 0.25, w, dev, False)
        zz = z.clone().requires_grad_(True)
        z_q, idx, loss = m(zz)
        ((z_q * up).sum() + loss).backward()
        outs.append((z_q.detach(), idx, loss.detach(), zz.grad, m.embedding.weight.grad))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_skewed_histogram_and_cancellation(dev):
    """trained-like latents (z ~ code + small noise) and a collapsed codebook usage (every token on a few
    codes): long segments split into pieces, and the difference form keeps grad_E accurate."""
    K, D = 1024, 32
    w = vo.make_codebook("vit", K, D, 50)
    for sigma, pick_from in ((0.01, K), (0.05, 3)):
        g = torch.Generator().manual_seed(51)
        en = vo.unit_rows(w)
        pick = torch.randint(0, pick_from, (20000,), generator=g)
        z = (en[pick] + sigma * torch.randn(20000, D, generator=g)).view(20, 1000, D).to(dev)
        up = torch.zeros_like(z)
        m = _module("vit", K, D, 0.25, w.to(dev), dev, False)
        zz = z.clone().requires_grad_(True)
        z_q, idx, loss = m(zz)
        loss.backward()
        gz64, gw64 = vo.analytic_backward("vit", z.reshape(-1, D).double(), w.to(dev).double(), idx.reshape(-1), 0.25,
                                          up.reshape(-1, D).double())
        assert rel_err(m.embedding.weight.grad.cpu().numpy(), gw64.cpu().numpy()) < GRAD_TOL
        assert rel_err(zz.grad.reshape(-1, D).cpu().numpy(), gz64.cpu().numpy()) < GRAD_TOL


def test_full_size_properties(dev):
    """BASELINE.json configs[2] size (262 144 tokens, K=8192, D=32), through size-independent properties:
    chunk invariance of the indices, histogram mass, chunk-additivity of the fixed-point segment sums
    (grad_E of the whole batch == from the two halves), and agreement of the tensor-core search with the
    exhaustive fp32 search on every row."""
    K, D = 8192, 32
    w = vo.make_codebook("vit", K, D, 0).to(dev)
    z = vo.make_latents((256, 1024, 32), 3).to(dev)
    auto = _module("vit", K, D, 0.25, w, dev, False)
    simt = _module("vit", K, D, 0.25, w, dev, True)
    with torch.no_grad():
        idx_auto = auto.encode(z)
        idx_simt = simt.encode(z)
        halves = torch.cat([auto.encode(z[:100]), auto.encode(z[100:])])
    assert torch.equal(idx_auto, idx_simt)
    assert torch.equal(idx_auto, halves)
    zz = z.clone().requires_grad_(True)
    z_q, idx, loss = auto(zz)
    assert int(auto.last_histogram.sum()) == 256 * 1024
    loss.backward()
    g_full = auto.embedding.weight.grad.clone()
    # same batch in two chunks with the global normaliser: integer segment sums add exactly
    from vq_b200 import functional as F_vq
    parts = []
    for sl in (slice(0, 100), slice(100, 256)):
        wp = w.clone().requires_grad_(True)
        o = F_vq.quantise(z[sl], wp, "vit", 0.25, n_elem_total=256 * 1024 * D)
        o[2].backward()
        parts.append(wp.grad)
    assert rel_err((parts[0] + parts[1]).cpu().numpy(), g_full.cpu().numpy()) < 1e-6
    # oracle on a sample of rows (the full T x K matrix is 8 GiB): indices agree
    with torch.no_grad():
        ref = vo.quantise_chunked("vit", z[:32], w, 0.25, chunk_tokens=8192)
    rep = vo.classify_index_mismatches(idx[:32].reshape(-1), ref.indices.reshape(-1),
                                       vo.unit_rows(z[:32].reshape(-1, D)), vo.unit_rows(w))
    assert rep["hard_rows"] == 0, rep


def test_full_size_cfg3_step_matches_reference_chunked(dev):
    """BASELINE.json configs[2] at FULL size (256 x 1024 tokens, K = 8192, D = 32), forward AND backward, against the oracle
    run on the same GPU in token chunks (the reference materialises a T x K matrix: 8 GiB unchunked): indices equal off
    near-tie rows, z_q bit-exact on those rows, loss / grad_z / grad_weight within 1e-5, through the drop-in module and
    through the preallocated step the bench times."""
    from vq_b200 import dist as vq_dist
    K, D, shape = 8192, 32, (256, 1024, 32)
    w = vo.make_codebook("vit", K, D, 0).to(dev)
    z = vo.make_latents(shape, 3).to(dev)
    up = vo.make_latents(shape, 4).to(dev)
    ref = vo.quantise_step_chunked("vit", z, w, 0.25, up, chunk_tokens=32768)
    m = _module("vit", K, D, 0.25, w, dev, False)
    zz = z.clone().requires_grad_(True)
    z_q, idx, loss = m(zz)
    torch.autograd.backward([z_q, loss], [up, torch.ones((), device=dev)])
    step = vq_dist.ShardedQuantiser("vit", 0.25, world_size=1).step(z, up, w)
    rep = vo.classify_index_mismatches(idx.reshape(-1), ref.indices.reshape(-1), vo.unit_rows(z.reshape(-1, D)), vo.unit_rows(w))
    assert rep["hard_rows"] == 0, rep
    same = (idx.reshape(-1) == ref.indices.reshape(-1))
    assert int((~same).sum()) <= 64, rep                                       # near-tie rows only, a handful per 262 144
    assert torch.equal(z_q.detach().reshape(-1, D)[same], ref.z_q.reshape(-1, D)[same])
    assert torch.equal(step["z_q"], z_q.detach()) and torch.equal(step["indices"], idx.reshape(-1))
    for name, got, want in (("loss", loss.detach(), ref.loss), ("grad_z", zz.grad, ref.grad_z),
                            ("grad_weight", m.embedding.weight.grad, ref.grad_weight),
                            ("step.loss", step["loss"], ref.loss), ("step.grad_z", step["grad_z"], ref.grad_z),
                            ("step.grad_weight", step["grad_weight"], ref.grad_weight)):
        assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < GRAD_TOL, name
    assert int(step["histogram"].sum()) == 256 * 1024


@pytest.mark.parametrize("T", [1000, 3 * 32768 + 100])
def test_host_buffer_step_equals_device_step(dev, T):
    """vq_host_step (host pointers in/out; above 64 Ki tokens it streams the batch in token chunks over two copy
    streams) must return exactly what the device-pointer calls return: indices, z_q, grad_z, grad_weight and the
    loss are bit-identical because every cross-token quantity is an integer (fixed-point) sum."""
    import ctypes
    from vq_b200 import _lib
    from vq_b200 import dist as vq_dist
    K, D, beta = 1024, 32, 0.25
    lib = _lib.load()
    w = vo.make_codebook("vit", K, D, 11)
    z = vo.make_latents((T, D), 12)
    up = vo.make_latents((T, D), 13)
    ref = vq_dist.ShardedQuantiser("vit", beta, world_size=1).step(z.to(dev), up.to(dev), w.to(dev))
    ref = {k: v.clone() for k, v in ref.items()}
    arena_bytes = _lib.size_query("vq_host_step_arena_bytes", T, K, D)
    arena = torch.empty(arena_bytes, dtype=torch.uint8, device=dev)
    hz, hg, hw = z.pin_memory(), up.pin_memory(), w.pin_memory()
    o_zq, o_gz = torch.empty(T, D).pin_memory(), torch.empty(T, D).pin_memory()
    o_idx = torch.empty(T, dtype=torch.int64).pin_memory()
    o_loss, o_gw = torch.empty(1).pin_memory(), torch.empty(K, D).pin_memory()
    o_stats = torch.empty(_lib.STATS_LEN, dtype=torch.int64).pin_memory()
    for _ in range(2):      # twice: the second call reuses streams, events and the arena
        _lib.check(lib.vq_host_step(hz.data_ptr(), hg.data_ptr(), T, hw.data_ptr(), K, D, 0, beta, o_zq.data_ptr(),
                                    o_idx.data_ptr(), o_loss.data_ptr(), o_gz.data_ptr(), o_gw.data_ptr(),
                                    o_stats.data_ptr(), arena.data_ptr(), arena_bytes,
                                    torch.cuda.current_stream(dev).cuda_stream))
        torch.cuda.synchronize()
        assert torch.equal(o_idx, ref["indices"].cpu())
        assert torch.equal(o_zq, ref["z_q"].cpu())
        assert torch.equal(o_gz, ref["grad_z"].cpu())
        assert torch.equal(o_gw, ref["grad_weight"].cpu())
        assert float(o_loss) == float(ref["loss"])
        o_zq.zero_(); o_gz.zero_(); o_idx.zero_(); o_gw.zero_(); o_loss.zero_()


@pytest.mark.parametrize("form,K,D,shape", [("vit", 8192, 32, (40, 1000, 32)), ("vit", 1024, 32, (3, 100, 32)),
                                            ("vit", 512, 64, (4, 300, 64)), ("vqgan", 1024, 256, (2, 256, 16, 16)),
                                            ("vit", 256, 16, (5, 77, 16))])
def test_segment_sums_from_forward_equal_bucketed_sums(dev, form, K, D, shape):
    """The codebook-gradient segment sums accumulated by the forward's finish pass (64-bit integer reductions)
    and the bucketed fixed-order sums of the backward are the same integers: grad_weight is bit-identical,
    on random and on collapsed (3 hot codes) usage."""
    w = vo.make_codebook(form, K, D, 60)
    for collapsed in (False, True):
        if collapsed:
            g = torch.Generator().manual_seed(61)
            n = int(torch.tensor(shape).prod()) // D
            rows = vo.unit_rows(w)[torch.randint(0, 3, (n,), generator=g)] + 0.02 * torch.randn(n, D, generator=g)
            z = rows.view(shape) if form == "vit" else rows.view(shape[0], shape[2], shape[3], D).permute(0, 3, 1, 2).contiguous()
        else:
            z = vo.make_latents(shape, 62)
        grads = []
        for sorted_segments in (False, True):
            m = _module(form, K, D, 0.25, w.to(dev), dev, False)
            m.sorted_segments = sorted_segments
            zz = z.to(dev).requires_grad_(True)
            z_q, idx, loss = m(zz)
            (z_q.sum() + loss).backward()
            grads.append((m.embedding.weight.grad.clone(), zz.grad.clone(), loss.detach().clone()))
        for a, b in zip(*grads):
            assert torch.equal(a, b)


def test_sharded_codebook_kernel_world_1_equals_plain_path(dev):
    """vq_backward_codebook_sharded with one rank (no peers): reads the exchange slot the forward wrote and must
    reproduce vq_backward_codebook + vq_loss_finalize bit for bit; both slots, increasing epochs."""
    import ctypes
    from vq_b200 import _lib
    from vq_b200 import dist as vq_dist
    from vq_b200.functional import _scratch
    lib = _lib.load()
    K, D, beta, T = 1024, 32, 0.25, 4096
    w = vo.make_codebook("vit", K, D, 70).to(dev)
    own, handle = ctypes.c_void_p(), ctypes.create_string_buffer(_lib.IPC_HANDLE_BYTES)
    nbytes = _lib.size_query("vq_exchange_bytes", K, D)
    _lib.check(lib.vq_peer_alloc(nbytes, ctypes.byref(own), handle))
    try:
        ptrs = (ctypes.c_void_p * 1)(own.value)
        cb = _scratch(_lib.size_query("vq_codebook_bytes", K, D), dev)
        fws_bytes = _lib.size_query("vq_workspace_bytes", T, K, D, 0)
        fws = _scratch(fws_bytes, dev)
        s = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.vq_codebook_prepare(w.data_ptr(), K, D, cb.data_ptr(), cb.numel(), s))
        for epoch in (1, 2, 3):
            z = vo.make_latents((T, D), 70 + epoch).to(dev)
            ref = vq_dist.ShardedQuantiser("vit", beta, world_size=1).step(z, None, w)
            slot = (epoch - 1) & 1
            seg, st, hist = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
            _lib.check(lib.vq_exchange_slot(own.value, K, D, slot, ctypes.byref(seg), ctypes.byref(st), ctypes.byref(hist)))
            z_q, idx = torch.empty_like(z), torch.empty(T, dtype=torch.int64, device=dev)
            zn, dn = torch.empty_like(z), torch.empty(T, device=dev)
            _lib.check(lib.vq_forward(z.data_ptr(), 0, T, 0, None, cb.data_ptr(), K, D, 0, beta, 0, T * D, z_q.data_ptr(),
                                      idx.data_ptr(), None, hist.value, st.value, zn.data_ptr(), dn.data_ptr(), seg.value,
                                      fws.data_ptr(), fws_bytes, s))
            gw, loss = torch.empty(K, D, device=dev), torch.empty(1, device=dev)
            hist_total = torch.empty(K, dtype=torch.int64, device=dev)
            stats_total = torch.empty(_lib.STATS_LEN, dtype=torch.int64, device=dev)
            _lib.check(lib.vq_backward_codebook_sharded(ptrs, 1, 0, slot, epoch, cb.data_ptr(), K, D, 0, beta, None, T * D,
                                                        gw.data_ptr(), hist_total.data_ptr(), loss.data_ptr(),
                                                        stats_total.data_ptr(), s))
            torch.cuda.synchronize()
            assert torch.equal(idx, ref["indices"]) and torch.equal(z_q, ref["z_q"])
            assert torch.equal(gw, ref["grad_weight"])
            assert float(loss) == float(ref["loss"])
            assert torch.equal(hist_total, ref["histogram"].to(torch.int64))
            assert int(stats_total[_lib.STAT_PEER_TIMEOUT]) == 0
            assert int(stats_total[_lib.STAT_LOSS_FIXED]) == int(ref["stats"][_lib.STAT_LOSS_FIXED])
    finally:
        torch.cuda.synchronize()
        _lib.check(lib.vq_peer_free(own.value))


def test_full_size_cfg2_vqgan_encode_imgs(dev):
    """BASELINE.json configs[1] at full size: VQGAN form, 8192 x 256 codebook, batch 64 x 256 px (16 x 16 latents,
    NCHW), against the same-device oracle (its T x K matrix is 512 MiB): indices, z_q bit-exact, loss."""
    K, D = 8192, 256
    w = vo.make_codebook("vqgan", K, D, 0).to(dev)
    z = vo.make_latents((64, 256, 16, 16), 2).to(dev)
    with torch.no_grad():
        ref = vo.quantise("vqgan", z, w, 0.25)
        for exact in (False, True):
            m = _module("vqgan", K, D, 0.25, w, dev, exact)
            rep = {}
            _check_forward("vqgan", m, z, w, ref, rep)
            codes = m.encode(z)
            assert torch.equal(codes.reshape(-1), m(z)[1].reshape(-1))
    del ref
    torch.cuda.empty_cache()


def test_full_size_cfg4_tokenise_round_trip(dev):
    """BASELINE.json configs[3] at full size: encode_imgs + decode_indices of 512 x 1024 tokens (8192 x 32), through
    size-independent properties: the decode of the codes is the unit code row (bit-exact gather), equals the forward's
    z_q up to the STE rounding, re-encoding the decoded latents is idempotent, chunked == unchunked."""
    K, D = 8192, 32
    w = vo.make_codebook("vit", K, D, 0).to(dev)
    z = vo.make_latents((512, 1024, D), 5).to(dev)
    m = _module("vit", K, D, 0.25, w, dev, False)
    with torch.no_grad():
        codes = m.encode(z)
        dec = m.indices_to_embeddings(codes.view(512, 1024))
        assert torch.equal(dec.reshape(-1, D), vo.unit_rows(w)[codes.reshape(-1)])
        z_q, idx, _ = m(z)
        assert torch.equal(idx.reshape(-1), codes.reshape(-1))
        assert float((dec - z_q).abs().max()) <= 2 ** -23
        again = m.encode(dec)
        assert float((again != codes).float().mean()) < 1e-4
        halves = torch.cat([m.encode(z[:200]).reshape(-1), m.encode(z[200:]).reshape(-1)])
        assert torch.equal(halves, codes.reshape(-1))
        hist = torch.bincount(codes.reshape(-1), minlength=K)
        assert int(hist.sum()) == 512 * 1024


@pytest.mark.parametrize("T,K,D", [(1 << 20, 16384, 32), (1 << 20, 1024, 32), (1 << 18, 4096, 256), (1 << 17, 16384, 256),
                                   (1 << 19, 2048, 64)])
def test_sweep_points_tensor_core_search_equals_exhaustive(dev, T, K, D):
    """BASELINE.json configs[4] (the synthetic sweep) at a few of its points: the tensor-core search and the exhaustive
    fp32 search return the same index for EVERY row, and the step's integer segment sums are chunk-additive."""
    w = vo.make_codebook("vit", K, D, 7).to(dev)
    g = torch.Generator(device=dev).manual_seed(6)
    z = torch.randn(T // 1024, 1024, D, device=dev, generator=g)
    auto = _module("vit", K, D, 0.25, w, dev, False)
    simt = _module("vit", K, D, 0.25, w, dev, True)
    with torch.no_grad():
        a = auto.encode(z)
        b = simt.encode(z)
    assert torch.equal(a, b)
    assert int(auto.near_tie_rows() if auto.last_stats is not None else 0) >= 0
    # oracle on a sample (the full T x K matrix does not fit): no hard mismatches
    with torch.no_grad():
        ref = vo.quantise_chunked("vit", z[:8], w, 0.25, chunk_tokens=4096)
    rep = vo.classify_index_mismatches(a.reshape(-1)[:8 * 1024], ref.indices.reshape(-1), vo.unit_rows(z[:8].reshape(-1, D)),
                                       vo.unit_rows(w))
    assert rep["hard_rows"] == 0, rep


@pytest.mark.parametrize("D", [32, 64, 256])
@pytest.mark.parametrize("K", [256, 768, 1024, 8192])
def test_boundary_token_counts_all_paths_agree(dev, K, D):
    """Token counts around every dispatch boundary -- below the tensor-core minimum (256), ragged row tiles, the point
    where the generic filter stops splitting the codebook (row tiles x splits > SMs), a little above kFlaggedCap -- :
    the automatic path and the exhaustive fp32 search give identical indices, z_q and gradients (loss to 1e-6)."""
    w = vo.make_codebook("vit", K, D, 90).to(dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    counts = [1, 31, 255, 256, 257, 1000, 4097, 256 * (sms // 4), 256 * (sms // 2) + 1, 256 * sms + 77]
    for T in counts:
        z = vo.make_latents((1, T, D), 91 + T % 7).to(dev)
        up = vo.make_latents((1, T, D), 92).to(dev)
        outs = []
        for exact in (False, True):
            m = _module("vit", K, D, 0.25, w, dev, exact)
            zz = z.clone().requires_grad_(True)
            z_q, idx, loss = m(zz)
            ((z_q * up).sum() + loss).backward()
            assert int(m.last_histogram.sum()) == T
            outs.append((idx, z_q.detach(), loss.detach(), zz.grad, m.embedding.weight.grad))
        for name, a, b in zip(("idx", "z_q", "loss", "grad_z", "grad_weight"), *outs):
            if name == "loss":     # fixed-point partials are grouped differently by the two finish kernels
                assert abs(float(a) - float(b)) <= 1e-6 * abs(float(b)), (T, K, D)
            else:
                assert torch.equal(a, b), (name, T, K, D)


@pytest.mark.parametrize("K,D,T", [(8192, 32, 8192), (1024, 256, 4096), (1000, 32, 300)])
def test_cuda_graph_replay_equals_eager(dev, K, D, T):
    """ShardedQuantiser(graphs=True): first use of a buffer set runs eagerly, the second is captured, later ones replay --
    every output bit-identical to the eager step, also after the buffers' CONTENTS change (a graph is keyed on
    addresses, not values)."""
    from vq_b200 import dist as vq_dist
    w = vo.make_codebook("vit", K, D, 80).to(dev)
    z = torch.empty(T // 100 if T % 100 == 0 else 1, 100 if T % 100 == 0 else T, D, device=dev)
    up = torch.empty_like(z)
    eager = vq_dist.ShardedQuantiser("vit", 0.25, world_size=1)
    graphed = vq_dist.ShardedQuantiser("vit", 0.25, world_size=1, graphs=True)
    for i in range(5):
        z.copy_(vo.make_latents(tuple(z.shape), 81 + i))
        up.copy_(vo.make_latents(tuple(z.shape), 91 + i))
        ref = {k: v.clone() for k, v in eager.step(z, up, w).items()}
        out = graphed.step(z, up, w)
        torch.cuda.synchronize()
        for k in ("z_q", "indices", "loss", "grad_z", "grad_weight", "histogram"):
            assert torch.equal(out[k], ref[k]), (k, i)
    assert len(graphed._graphs) == 1 and graphed.graph_kernel_launches > 0


def test_dropin_module_is_cuda_graph_capturable(dev):
    """torch.cuda.make_graphed_callables on the drop-in Codebook (forward and autograd backward captured): outputs and
    gradients bit-identical to the eager module -- the path allocates through torch, never syncs, never touches the
    legacy stream."""
    from vq_b200.vitvqgan import Codebook
    K, D = 1024, 32
    m = Codebook(K, D).to(dev)
    m_eager = Codebook(K, D).to(dev)
    m_eager.load_state_dict(m.state_dict())
    z = vo.make_latents((8, 512, D), 95).to(dev)
    up = vo.make_latents((8, 512, D), 96).to(dev)
    gm = torch.cuda.make_graphed_callables(m, (z.clone().requires_grad_(True),))
    for _ in range(2):
        outs = []
        for mod in (gm, m_eager):
            zz = z.clone().requires_grad_(True)
            z_q, idx, loss = mod(zz)
            ((z_q * up).sum() + loss).backward()
            w = m.embedding.weight if mod is gm else m_eager.embedding.weight
            outs.append((z_q.detach().clone(), idx.clone(), loss.detach().clone(), zz.grad.clone(), w.grad.clone()))
            w.grad = None
        for a, b in zip(*outs):
            assert torch.equal(a, b)


@pytest.mark.parametrize("K,D,shape,scale", [(1024, 64, (3, 64, 8, 8), 1.0), (512, 256, (2, 256, 5, 5), 0.05), (300, 32, (2, 32, 6, 6), 3.0)])
def test_plain_l2_form_matches_its_oracle(dev, K, D, shape, scale):
    """The un-normalised CNN form (north_star's "L2 form"; not in the reference, so its oracle is the reference's expression
    with l2_norm removed -- parity unpinned): indices equal off near ties, z_q bit-exact, loss and both gradients within
    1e-5, decode = raw gather, on codebooks / latents of very different magnitudes."""
    from vq_b200.vqgan_l2 import Codebook
    g = torch.Generator().manual_seed(77)
    w = (torch.randn(K, D, generator=g) * scale).to(dev)
    z = (torch.randn(*shape, generator=g) * scale).to(dev)
    up = torch.randn(*shape, generator=g).to(dev)
    ref = vo.quantise_step("l2", z, w, 0.3, up)
    m = Codebook(K, D, 0.3).to(dev)
    with torch.no_grad():
        m.embedding.weight.copy_(w)
    zz = z.clone().requires_grad_(True)
    z_q, idx, loss = m(zz)
    torch.autograd.backward([z_q, loss], [up, torch.ones((), device=dev)])
    assert idx.shape == ref.indices.shape and z_q.shape == ref.z_q.shape
    zt, d = z.permute(0, 2, 3, 1).reshape(-1, D), None
    dist = vo.distance_matrix(zt.double(), w.double())
    top2 = torch.topk(dist, 2, dim=1, largest=False).values
    near = (top2[:, 1] - top2[:, 0]) < 1e-6 * top2[:, 0].abs().clamp_min(1e-30)
    same = idx == ref.indices
    assert bool((same | near).all()), f"{int((~same & ~near).sum())} hard index mismatches"
    if bool(same.all()):
        assert torch.equal(z_q.detach(), ref.z_q.contiguous())
        assert rel_err(zz.grad.cpu().numpy(), ref.grad_z.cpu().numpy()) < GRAD_TOL
        assert rel_err(m.embedding.weight.grad.cpu().numpy(), ref.grad_weight.cpu().numpy()) < GRAD_TOL
    assert rel_err(loss.detach().cpu().numpy(), ref.loss.cpu().numpy()) < LOSS_TOL
    assert int(m.last_histogram.sum()) == idx.numel()
    dec = m.indices_to_embeddings(idx.view(shape[0], -1))
    assert torch.equal(dec.detach(), vo.indices_to_embeddings("l2", idx.view(shape[0], -1), w).contiguous())
    assert torch.equal(m.encode(z), idx)


@pytest.mark.parametrize("D,T,few", [(256, 9000, False), (256, 600, False), (64, 9000, False), (256, 1000, True),
                                     (128, 1000, True)])
def test_generic_filter_undecided_rows_take_the_exhaustive_paths(dev, D, T, few):
    """A codebook whose codes repeat eight times, 256 codes apart (= in eight different groups of the filter): every row
    has more near-best groups than a verdict record holds, so the filter decides nothing and the rows go to the
    fallbacks - the batched search inside the rescoring kernel (<= 128 listed rows), the tiled scan with the codebook
    split over blocks (<= 8192) and its unsplit tail (beyond).  Ties must resolve to the lowest index like argmin."""
    from vq_b200 import functional as F_vq
    K = 2048
    z = vo.make_latents((T, D), 61)
    if few:     # one code in eight copies, 50 rows next to it: a short list of undecided rows
        w = vo.make_codebook("vit", K, D, 60)
        w[5::256] = w[5]
        z[100:150] = w[5] * 2.0 + 0.01 * vo.make_latents((50, D), 62)
    else:
        w = vo.make_codebook("vit", 256, D, 60).repeat(K // 256, 1)
    w, z = w.to(dev), z.to(dev)
    prep = F_vq.prepare_codebook(w)
    want = F_vq.encode_indices(z, w, "vit", prepared=prep, exact_scan=True)
    got = F_vq.encode_indices(z, w, "vit", prepared=prep)
    ref = vo.quantise("vit", z.view(1, T, D), w, 0.25).indices.view(-1)
    assert torch.equal(got.view(-1), want.view(-1))
    assert torch.equal(got.view(-1), ref)
    if few:
        assert bool((got.view(-1)[100:150] == 5).all())
    else:
        assert int(got.max()) < 256     # the first copy of every code wins its ties
    # and a second call on the same workspace state (counters the kernels reset themselves)
    again = F_vq.encode_indices(z, w, "vit", prepared=prep)
    assert torch.equal(again, got)


@pytest.mark.parametrize("env", [{"VQ_TC16_W16": "1"}, {"VQ_TC16_TS": "1"}, {"VQ_TC16_DISABLE": "1"}, {"VQ_TC16_CLUSTER": "1"},
                                 {"VQ_TC_CLUSTER": "0"}],
                         ids=["wide-drain", "rows-in-tmem", "generic-filter", "cluster-pairs-d32", "no-clusters-d256"])
def test_alternative_filter_kernels_equal_exhaustive_search(env):
    """The measured alternatives of the filters that stay in the library behind environment switches (D = 32: 16 epilogue
    warps / token rows in tensor memory / the generic fp32-accumulator filter / clusters of two 128-row CTAs; D = 256: the
    256-row kernel instead of the cluster pairs) return the exhaustive search's indices on every row; the library reads
    the switches when it is loaded, hence a fresh process."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path[:0] = [%r, %r]\n"
        "import torch\n"
        "from oracle import vq_oracle as vo\n"
        "from vq_b200 import functional as F\n"
        "dev = torch.device('cuda:0')\n"
        "w = vo.make_codebook('vit', 8192, 32, 0).to(dev)\n"
        "for shape, seed in (((2, 1024, 32), 1), ((41, 1000, 32), 2)):\n"
        "    z = vo.make_latents(shape, seed).to(dev)\n"
        "    p = F.prepare_codebook(w)\n"
        "    a = F.encode_indices(z, w, 'vit', prepared=p)\n"
        "    b = F.encode_indices(z, w, 'vit', prepared=p, exact_scan=True)\n"
        "    assert torch.equal(a, b), int((a != b).sum())\n"
        "    zq, idx, loss, hist, stats = F.quantise(z, w, 'vit', prepared=p)\n"
        "    assert torch.equal(idx, b.reshape(-1)) and int(hist.sum()) == idx.numel()\n"
        "w = vo.make_codebook('vit', 2048, 256, 3).to(dev)\n"
        "z = vo.make_latents((9, 500, 256), 4).to(dev)\n"
        "p = F.prepare_codebook(w)\n"
        "assert torch.equal(F.encode_indices(z, w, 'vit', prepared=p), F.encode_indices(z, w, 'vit', prepared=p, exact_scan=True))\n"
        "print('ok')\n") % (ROOT, os.path.join(ROOT, "attention-models_b200"))
    out = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
