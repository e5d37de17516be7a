"""Pin the oracle restatement against outputs of the unmodified reference (tests/golden/*.npz,
written by oracle/make_golden.py).  CPU only.

The reference has no tests or golden vectors of its own, so these fixtures --
the reference's own Codebook.forward / indices_to_embeddings / autograd run in the build
container -- are what "parity pinned" means for this repo.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err, ulp_distance
from oracle import vq_oracle as vo

torch.set_num_threads(1)


def _assert_indices(form, idx, ref_idx, z, w):
    """Exact, except near-tie rows (top-2 fp32 distances < 1e-6 relative apart) on a host whose BLAS
    sums in another order than the container that wrote the fixtures."""
    if np.array_equal(idx.numpy().reshape(-1), ref_idx.reshape(-1)):
        return np.zeros(ref_idx.size, bool)
    D = w.shape[1]
    zt = z.permute(0, 2, 3, 1).reshape(-1, D) if form == vo.VQGAN else z.reshape(-1, D)
    r = vo.classify_index_mismatches(idx.reshape(-1), torch.from_numpy(ref_idx.reshape(-1)),
                                     vo.unit_rows(zt), vo.unit_rows(w))
    assert r["hard_rows"] == 0, r
    return idx.numpy().reshape(-1) != ref_idx.reshape(-1)


def _assert_zq(form, z_q, ref, bad_rows):
    D = ref.shape[1] if form == vo.VQGAN else ref.shape[-1]
    a = z_q.permute(0, 2, 3, 1).reshape(-1, D).numpy() if form == vo.VQGAN else z_q.reshape(-1, D).numpy()
    b = np.transpose(ref, (0, 2, 3, 1)).reshape(-1, D) if form == vo.VQGAN else ref.reshape(-1, D)
    assert ulp_distance(a[~bad_rows], b[~bad_rows]).max() <= 2


def _inputs(g):
    form = str(g["form"])
    w = vo.make_codebook(form, int(g["K"]), int(g["D"]), int(g["w_seed"]))
    z = vo.make_latents(tuple(int(s) for s in g["shape"]), int(g["z_seed"]))
    return form, w, z


@pytest.mark.parametrize("name", ["vit_cfg1_fwd", "vqgan_cfg2_slice_fwd"])
def test_forward_matches_reference(name):
    g = load_golden(name)
    form, w, z = _inputs(g)
    out = vo.quantise(form, z, w, float(g["beta"]))
    ref_idx = g["indices"].astype(np.int64).reshape(tuple(g["idx_shape"]))
    # same ops, same library, same thread count -> bit-identical on CPU
    assert out.indices.dtype == torch.int64
    assert tuple(out.indices.shape) == ref_idx.shape
    bad = _assert_indices(form, out.indices, ref_idx, z, w)
    assert out.z_q.shape == torch.Size(g["z_q"].shape)
    _assert_zq(form, out.z_q, g["z_q"], bad)
    assert rel_err(out.loss.numpy(), g["loss"]) < 1e-6


def test_vqgan_output_layout_is_nchw_and_indices_flat():
    g = load_golden("vqgan_cfg2_slice_fwd")
    form, w, z = _inputs(g)
    out = vo.quantise(form, z, w)
    assert out.z_q.shape == z.shape                       # (b, D, h, w)
    assert out.indices.dim() == 1 and out.indices.numel() == z.shape[0] * z.shape[2] * z.shape[3]


@pytest.mark.parametrize("name", ["vit_step_small", "vit_step_beta", "vqgan_step_small", "vqgan_step_d64"])
def test_step_matches_reference(name):
    g = load_golden(name)
    form, w, z = _inputs(g)
    up = vo.make_latents(tuple(int(s) for s in g["shape"]), int(g["g_seed"]))
    o = vo.quantise_step(form, z, w, float(g["beta"]), up)
    bad = _assert_indices(form, o.indices, g["indices"].astype(np.int64), z, w)
    assert not bad.any()          # gradients below are only comparable when every row agrees
    _assert_zq(form, o.z_q, g["z_q"], bad)
    assert rel_err(o.loss.numpy(), g["loss"]) < 1e-6
    assert rel_err(o.grad_z.numpy(), g["grad_z"]) < 1e-6
    assert rel_err(o.grad_weight.numpy(), g["grad_weight"]) < 1e-6


@pytest.mark.parametrize("name", ["vit_step_small", "vit_step_beta", "vqgan_step_small", "vqgan_step_d64"])
def test_analytic_backward_matches_reference_autograd(name):
    """SURVEY.md Appendix A closed form vs the reference's autograd gradients (tolerance 1e-5, north_star)."""
    g = load_golden(name)
    form, w, z = _inputs(g)
    up = vo.make_latents(tuple(int(s) for s in g["shape"]), int(g["g_seed"]))
    D = int(g["D"])
    if form == vo.VQGAN:
        z_tok = z.permute(0, 2, 3, 1).reshape(-1, D)
        up_tok = up.permute(0, 2, 3, 1).reshape(-1, D)
    else:
        z_tok, up_tok = z.reshape(-1, D), up.reshape(-1, D)
    idx = torch.from_numpy(g["indices"].astype(np.int64).reshape(-1))
    gz, gw = vo.analytic_backward(form, z_tok.double(), w.double(), idx, float(g["beta"]), up_tok.double())
    ref_gz = torch.from_numpy(g["grad_z"])
    if form == vo.VQGAN:
        ref_gz = ref_gz.permute(0, 2, 3, 1).reshape(-1, D)
    assert rel_err(gz.numpy(), ref_gz.reshape(-1, D).numpy()) < 1e-5
    assert rel_err(gw.numpy(), g["grad_weight"]) < 1e-5


@pytest.mark.parametrize("name", ["vit_decode", "vqgan_decode"])
def test_indices_to_embeddings_matches_reference(name):
    g = load_golden(name)
    form = str(g["form"])
    w = vo.make_codebook(form, int(g["K"]), int(g["D"]), int(g["w_seed"]))
    gen = torch.Generator().manual_seed(int(g["i_seed"]))
    idx = torch.randint(0, int(g["K"]), (int(g["b"]), int(g["n"])), generator=gen)
    e = vo.indices_to_embeddings(form, idx, w)
    assert tuple(e.shape) == tuple(g["embeds_shape"])
    assert ulp_distance(e.contiguous().numpy(), g["embeds"]).max() <= 2


def test_degenerate_rows_match_reference():
    """zero row, NaN row (-> index 0, NaN loss), exact-code row, anti-code row."""
    g = load_golden("vit_degenerate_fwd")
    form, w, z = _inputs(g)
    z[0, 0] = 0.0
    z[0, 1, 3] = float("nan")
    z[0, 2] = w[5] * 3.0
    z[1, 0] = -w[7]
    out = vo.quantise(form, z, w)
    ref_idx = g["indices"].astype(np.int64).reshape(tuple(g["idx_shape"]))
    got = out.indices.numpy()
    got[0, 0] = ref_idx[0, 0]     # the all-zero row is decided by 1-ulp differences of ||en||^2
    assert np.array_equal(got, ref_idx)
    assert ref_idx[0, 1] == 0 and ref_idx[0, 2] == 5
    assert np.isnan(float(g["loss"])) and np.isnan(float(out.loss))
    a, b = out.z_q.numpy().copy(), g["z_q"].copy()
    a[0, 0] = b[0, 0]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert ulp_distance(np.nan_to_num(a), np.nan_to_num(b)).max() <= 2


def test_chunked_equals_whole():
    w = vo.make_codebook(vo.VIT, 512, 32, 1)
    z = vo.make_latents((8, 64, 32), 2)
    up = vo.make_latents((8, 64, 32), 3)
    a = vo.quantise_step(vo.VIT, z, w, 0.25, up)
    b = vo.quantise_step_chunked(vo.VIT, z, w, 0.25, up, chunk_tokens=128)
    assert torch.equal(a.indices, b.indices)
    assert rel_err(b.grad_weight.numpy(), a.grad_weight.numpy()) < 1e-6
    assert rel_err(b.grad_z.numpy(), a.grad_z.numpy()) < 1e-6
    assert rel_err(b.loss.numpy(), a.loss.numpy()) < 1e-6
    c = vo.quantise_chunked(vo.VIT, z, w, 0.25, chunk_tokens=128)
    assert torch.equal(a.indices, c.indices)


def test_histogram_is_bincount():
    idx = torch.tensor([[0, 3, 3], [7, 0, 3]])
    h = vo.code_histogram(idx, 8)
    assert h.tolist() == [2, 0, 0, 3, 0, 0, 0, 1]


def test_near_tie_classifier():
    w = vo.make_codebook(vo.VIT, 256, 32, 5)
    z = vo.make_latents((64, 32), 6)
    zn, en = vo.unit_rows(z), vo.unit_rows(w)
    idx = torch.argmin(vo.distance_matrix(zn, en), dim=1)
    wrong = idx.clone()
    wrong[3] = (idx[3] + 1) % 256
    r = vo.classify_index_mismatches(wrong, idx, zn, en)
    assert r == {"mismatch_rows": 1, "near_tie_rows": 0, "hard_rows": 1}
    assert torch.equal(vo.argmin_fp64(zn, en), idx)
    assert (vo.top2_relative_gap(zn, en) >= 0).all()


def test_token_consumer_oracle_matches_reference_fixture():
    """the masked-fill + embedding + pos_enc restatement against what the reference's own
    BiDirectionalTransformer modules produced (tests/golden/maskgit_token_embed.npz)."""
    g = load_golden("maskgit_token_embed")
    vocab, dim, b, n, seed = (int(g[k]) for k in ("vocab", "dim", "b", "n", "seed"))
    table, pos, tokens, mask = vo.token_embed_inputs(vocab, dim, b, n, seed)
    embeds, ids, labels = vo.masked_token_embeddings(tokens, mask, vocab, table, pos, -1)
    assert np.array_equal(ids.numpy(), g["input_ids"]) and np.array_equal(labels.numpy(), g["labels"])
    assert np.array_equal(embeds.numpy(), g["embeds"])          # a gather and one fp32 add: bit-exact


def test_token_consumer_gradients_and_causal_form_match_reference_fixture():
    """The oracle's two token consumers, forward and autograd backward, against what the reference's own modules gave
    (tests/golden/token_consumers_grad.npz: MaskGIT's BiDirectionalTransformer lookup and Parti's shifted decoder input)."""
    g = load_golden("token_consumers_grad")
    vocab = int(g["vocab"])
    tokens, mask, up = torch.from_numpy(g["tokens"]), torch.from_numpy(g["mask"]), torch.from_numpy(g["upstream"])
    table = torch.from_numpy(g["maskgit_table"]).requires_grad_(True)
    pos = torch.from_numpy(g["maskgit_pos"]).requires_grad_(True)
    e, _, _ = vo.masked_token_embeddings(tokens, mask, vocab, table, pos, -1)
    (e * up).sum().backward()
    assert np.array_equal(e.detach().numpy(), g["maskgit_embeds"])
    assert rel_err(table.grad.numpy(), g["maskgit_grad_table"]) < 1e-6 and rel_err(pos.grad.numpy(), g["maskgit_grad_pos"]) < 1e-6
    ptable = torch.from_numpy(g["parti_table"]).requires_grad_(True)
    start = torch.from_numpy(g["parti_start"]).requires_grad_(True)
    pe, labels = vo.causal_token_embeddings(tokens, ptable, torch.from_numpy(g["parti_pe"]), start)
    (pe * up).sum().backward()
    assert np.array_equal(pe.detach().numpy(), g["parti_embeds"]) and np.array_equal(labels.numpy(), g["parti_labels"])
    assert rel_err(ptable.grad.numpy(), g["parti_grad_table"]) < 1e-6 and rel_err(start.grad.numpy(), g["parti_grad_start"]) < 1e-6


@pytest.mark.parametrize("name", ["vit_decode_grad", "vqgan_decode_grad"])
def test_decode_is_differentiable_like_the_reference(name):
    """indices_to_embeddings carries a gradient to the codebook in the reference (nn.Embedding lookup, normalised in the
    ViT form): the oracle's restatement gives the fixture's gradient."""
    g = load_golden(name)
    w = torch.from_numpy(g["weight"]).requires_grad_(True)
    out = vo.indices_to_embeddings(str(g["form"]), torch.from_numpy(g["indices"]), w)
    (out * torch.from_numpy(g["upstream"])).sum().backward()
    assert np.array_equal(out.detach().numpy(), g["out"])
    assert rel_err(w.grad.numpy(), g["grad_weight"]) < 1e-6


def test_projected_quantiser_oracle_matches_reference_fixture():
    """pre_quant -> Codebook -> autograd and post_quant(indices_to_embeddings) of the ViT wrapper
    as the reference's own modules computed them (oracle/make_golden.py::projected_case)."""
    g = load_golden("vit_projected_step")
    K, D, C, b, n, seed = (int(g[k]) for k in ("K", "D", "C", "b", "n", "seed"))
    w = vo.make_codebook(vo.VIT, K, D, seed).requires_grad_(True)
    w_pre, b_pre = (t.requires_grad_(True) for t in vo.projection_inputs(C, D, seed + 1))
    w_post, b_post = vo.projection_inputs(D, C, seed + 2)
    x = vo.make_latents((b, n, C), seed + 3).requires_grad_(True)
    up = vo.make_latents((b, n, D), seed + 4)
    z, o = vo.quantise_projected(x, w_pre, b_pre, w, float(g["beta"]))
    assert ulp_distance(z.detach().numpy(), g["z"]).max() <= 4
    bad = _assert_indices(vo.VIT, o.indices, g["indices"].astype(np.int64), z.detach(), w.detach())
    _assert_zq(vo.VIT, o.z_q.detach(), g["z_q"], bad)
    assert rel_err(o.loss.detach().numpy(), g["loss"]) < 1e-6
    ((o.z_q * up).sum() + o.loss).backward()
    for name, t in (("grad_x", x), ("grad_w_pre", w_pre), ("grad_b_pre", b_pre), ("grad_weight", w)):
        assert rel_err(t.grad.numpy(), g[name]) < 1e-5, name
    tokens = torch.from_numpy(g["tokens"].astype(np.int64))
    dec = vo.decode_projected(vo.VIT, tokens, w.detach(), w_post, b_post)
    assert dec.shape == (b, n, C) and rel_err(dec.numpy(), g["decoded"]) < 1e-6


def test_projected_decode_oracle_matches_reference_fixture_cnn_form():
    g = load_golden("vqgan_projected_decode")
    K, D, b, side, seed = (int(g[k]) for k in ("K", "D", "b", "side", "seed"))
    w = vo.make_codebook(vo.VQGAN, K, D, seed)
    w_post, b_post = vo.projection_inputs(D, D, seed + 2, conv=True)
    tokens = torch.from_numpy(g["tokens"].astype(np.int64))
    dec = vo.decode_projected(vo.VQGAN, tokens, w, w_post, b_post)
    assert dec.shape == (b, D, side, side) and rel_err(dec.numpy(), g["decoded"]) < 1e-6


def test_3xtf32_model_of_the_fused_pre_quant_gemm_is_fp32_level():
    """The arithmetic of k_prequant_prep restated in numpy (oracle/tf32_split.py): its error against a float64 product is at
    the level of a plain fp32 GEMM under both models of the tensor core's internal sums, far inside the GPU tests' Z_TOL
    (2e-5 of sum |x_c W_dc|); measured on a B200: 1.1e-7, cuBLAS fp32 2.6e-7 (profiles/r02b_bench_cfg3pre.json)."""
    from oracle import tf32_split as ts
    C, D, T = 512, 32, 256
    w, b = (t.numpy() for t in vo.projection_inputs(C, D, 1))
    x = vo.make_latents((T, C), 3).numpy()
    hi, lo = ts.split_tf32(x)
    assert np.all((hi.view(np.uint32) & 0x1FFF) == 0) and np.all((lo.view(np.uint32) & 0x1FFF) == 0)
    assert np.max(np.abs((x - hi - lo) / x)) < 2.0 ** -21          # what the dropped xl.wl term and lo's rounding leave
    e_fp32 = ts.error_over_sum_abs_terms(torch.nn.functional.linear(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b)).numpy(), x, w, b)
    e_exact = ts.error_over_sum_abs_terms(ts.linear_3xtf32(x, w, b, "exact"), x, w, b)
    e_trunc = ts.error_over_sum_abs_terms(ts.linear_3xtf32(x, w, b, "truncate"), x, w, b)
    assert e_exact < 5e-7 and e_trunc < 1e-6 and e_fp32 < 1e-6
    assert e_trunc < 20 * max(e_fp32, 5e-8)
    # a tf32-only GEMM (no split) would NOT be: that is what the split buys
    e_tf32 = ts.error_over_sum_abs_terms((ts.rna_tf32(x).astype(np.float64) @ ts.rna_tf32(w).astype(np.float64).T + b).astype(np.float32), x, w, b)
    assert e_tf32 > 20 * e_exact


@pytest.mark.parametrize("form,shape", [(vo.VIT, (0, 4, 32)), (vo.VQGAN, (0, 32, 2, 2))])
def test_empty_batch_convention(form, shape):
    """A batch without tokens: the reference (and so the oracle) returns empty z_q / indices and a NaN loss
    (``torch.mean`` of nothing).  The product mirrors that without launching anything (functional._empty_result)."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "attention-models_b200"))
    from vq_b200 import functional as F_vq
    o = vo.quantise(form, torch.empty(*shape), vo.make_codebook(form, 64, 32, 0), 0.25)
    z_q, idx, loss, hist, stats = F_vq._empty_result(shape, 64, torch.device("cpu"))
    assert z_q.shape == o.z_q.shape and z_q.dtype == o.z_q.dtype
    assert idx.numel() == o.indices.numel() == 0 and idx.dtype == o.indices.dtype
    assert torch.isnan(loss) and torch.isnan(o.loss) and loss.dim() == 0
    assert int(hist.sum()) == 0 and hist.shape == (64,)
