"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/vq_b200.h
declares, answers its size queries, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vq_b200.h")


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, os.path.join(ROOT, "attention-models_b200", "csrc"))
    import build as vq_build
    vq_build.build()
    from vq_b200 import _lib
    return _lib


def _declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"VQ_API\s+[\w\s\*]+?\b(vq_\w+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound(lib):
    declared = _declared_symbols()
    assert len(declared) >= 14
    cdll = lib.load()
    for name in declared:
        assert hasattr(cdll, name), f"{name} declared in include/vq_b200.h but not exported"
    assert sorted(lib.SIGNATURES) == declared, "ctypes table and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (vq_\w+)", out)))
    assert exported == declared, "library exports symbols the header does not declare (or vice versa)"


def test_abi_version_and_constants_match_header(lib):
    text = open(HEADER).read()
    consts = dict(re.findall(r"#define\s+(VQ_\w+)\s+(\d+)", text))
    assert lib.load().vq_abi_version() == int(consts["VQ_ABI_VERSION"]) == lib.ABI_VERSION
    assert (lib.FORM_VIT, lib.FORM_VQGAN) == (int(consts["VQ_FORM_VIT"]), int(consts["VQ_FORM_VQGAN"]))
    assert (lib.LAYOUT_TOKEN_MAJOR, lib.LAYOUT_NCHW) == (int(consts["VQ_LAYOUT_TOKEN_MAJOR"]), int(consts["VQ_LAYOUT_NCHW"]))
    assert (lib.FLAG_INDICES_ONLY, lib.FLAG_EXACT_SCAN) == (int(consts["VQ_FLAG_INDICES_ONLY"]), int(consts["VQ_FLAG_EXACT_SCAN"]))
    assert lib.FLAG_KEEP_STATS == int(consts["VQ_FLAG_KEEP_STATS"])
    assert lib.STATS_LEN == int(consts["VQ_STATS_LEN"])
    for name in ("NEAR_TIE_ROWS", "AMBIGUOUS_ROWS", "FALLBACK_ROWS", "LOSS_FIXED", "BAD_INDEX", "NONFINITE"):
        assert getattr(lib, "STAT_" + name) == int(consts["VQ_STAT_" + name])


def test_size_queries_and_argument_errors(lib):
    cb = lib.size_query("vq_codebook_bytes", 8192, 32)
    assert cb >= 8192 * 32 * 6 + 8192 * 8          # fp32 + fp16 unit codes, squared norms, row norms
    ws = lib.size_query("vq_workspace_bytes", 262144, 8192, 32, 0)
    assert ws >= 262144 * 32 * 4
    assert lib.size_query("vq_backward_workspace_bytes", 1024, 512, 32) > 0
    assert lib.size_query("vq_host_step_arena_bytes", 1024, 512, 32) > 0
    out = ctypes.c_size_t(0)
    cdll = lib.load()
    assert cdll.vq_codebook_bytes(8192, 48, ctypes.byref(out)) != 0          # unsupported codebook_dim
    assert b"codebook_dim" in cdll.vq_last_error()
    assert cdll.vq_workspace_bytes(-1, 8192, 32, 0, ctypes.byref(out)) != 0
    with pytest.raises(lib.VQLibraryError):
        lib.size_query("vq_codebook_bytes", 0, 32)


def test_no_cpu_fallback():
    """Product modules refuse CPU tensors instead of silently computing elsewhere."""
    from vq_b200.vitvqgan import Codebook as Vit
    from vq_b200.vqgan import Codebook as Vqgan
    with pytest.raises(RuntimeError, match="no CPU path"):
        Vit(64, 32)(torch.randn(1, 4, 32))
    with pytest.raises(RuntimeError, match="no CPU path"):
        Vqgan(64, 32)(torch.randn(1, 32, 2, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        Vit(64, 32).indices_to_embeddings(torch.zeros(1, 4, dtype=torch.long))


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_compute_call_without_device_reports_cuda_error(lib):
    cdll = lib.load()
    assert cdll.vq_device_info(None, None, None) == 3          # VQ_ERR_CUDA


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under attention-models_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "attention-models_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(base, f)


def test_module_surface_matches_reference():
    """ctor signature, attributes and state_dict key the reference's callers rely on."""
    import inspect
    from vq_b200.vitvqgan import Codebook as Vit
    from vq_b200.vqgan import Codebook as Vqgan
    sv = inspect.signature(Vit.__init__).parameters
    assert [sv[k].default for k in ("codebook_size", "codebook_dim", "beta")] == [8192, 32, 0.25]
    sq = inspect.signature(Vqgan.__init__).parameters
    assert [sq[k].default for k in ("codebook_size", "codebook_dim", "beta")] == [1024, 256, 0.25]
    m = Vit(codebook_size=128, codebook_dim=32)
    assert (m.codebook_size, m.codebook_dim, m.beta) == (128, 32, 0.25)
    assert list(m.state_dict().keys()) == ["embedding.weight"]
    assert isinstance(m.embedding, torch.nn.Embedding)
    q = Vqgan(64, 32)
    assert float(q.embedding.weight.abs().max()) <= 1.0 / 64          # uniform_(-1/K, 1/K), vqgan.py:146
    assert hasattr(m, "indices_to_embeddings") and hasattr(m, "forward")


def test_exchange_layout_queries_and_argument_errors(lib):
    """Host-side logic of the token-sharded entry points: sizes, slot pointers, argument checks (no device needed)."""
    cdll = lib.load()
    K, D = 8192, 32
    n = lib.size_query("vq_exchange_bytes", K, D)
    seg_bytes = (K * D + K) * 8
    assert n >= 2 * (seg_bytes + K * 4 + 64) + K * D * 4 + K * 8          # two slots + results + control
    base = 1 << 20                                                         # any aligned fake address: only arithmetic happens
    ptrs = []
    for slot in (0, 1):
        seg, st, hist = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        assert cdll.vq_exchange_slot(base, K, D, slot, ctypes.byref(seg), ctypes.byref(st), ctypes.byref(hist)) == 0
        assert base < seg.value < st.value < hist.value < base + n
        assert st.value - seg.value >= seg_bytes and seg.value % 256 == 0
        ptrs.append(seg.value)
    assert ptrs[1] - ptrs[0] >= seg_bytes + K * 4 + 64
    assert cdll.vq_exchange_slot(base, K, D, 2, None, None, None) != 0      # slot out of range
    one = (ctypes.c_void_p * 1)(base)
    # world / rank / slot validation happens before anything touches the device
    assert cdll.vq_backward_codebook_sharded(one, 0, 0, 0, 1, base, K, D, 0, 0.25, None, K * D, base, None, None, None, None) != 0
    assert cdll.vq_backward_codebook_sharded(one, 1, 1, 0, 1, base, K, D, 0, 0.25, None, K * D, base, None, None, None, None) != 0
    assert cdll.vq_backward_codebook_sharded(one, 17, 0, 0, 1, base, K, D, 0, 0.25, None, K * D, base, None, None, None, None) != 0
    assert b"world" in cdll.vq_last_error()
    assert cdll.vq_backward_sharded(one, 1, 0, 3, 1, None, 1024, base, base, base, base, K, D, 0, 0.25, None, 1024 * D, base, base,
                                    None, None, None, None) != 0
    # the one-GPU fused backward needs the forward's segment sums
    assert cdll.vq_backward(None, 0, 1024, 0, base, base, base, base, K, D, 0, 0.25, None, 1024 * D, None, None, base, base, None,
                            base, 1 << 20, None) != 0
    assert b"seg_sums" in cdll.vq_last_error()
    assert lib.PEER_MAX_RANKS == 16 and lib.IPC_HANDLE_BYTES == 64


def test_workspace_grows_with_code_splits(lib):
    """Few row tiles: the generic filter cuts the codebook into ranges, one record per (row, range)."""
    small = lib.size_query("vq_workspace_bytes", 16384, 8192, 256, 0)      # 64 row tiles -> 2 ranges
    large = lib.size_query("vq_workspace_bytes", 16384 * 64, 8192, 256, 0)
    assert small / 16384 > large / (16384 * 64)                            # more record bytes per row when split


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="the reference is only mounted in the build container")
def test_patch_reference_model_swaps_the_quantiser_in_place():
    """INTEGRATION.md route 1 on the unmodified reference classes: same attributes, same state_dict keys and weights,
    and the swapped module refuses CPU tensors (there is no CPU path)."""
    import importlib
    import torch.nn as nn
    sys.path.insert(0, "/root/reference")
    sys.dont_write_bytecode = True
    from vq_b200.integration import patch_reference_model
    from vq_b200 import vitvqgan as b_vit, vqgan as b_vqgan
    for mod_name, ours in (("models.vitvqgan", b_vit.Codebook), ("models.vqgan", b_vqgan.Codebook)):
        ref_cls = importlib.import_module(mod_name).Codebook

        class Holder(nn.Module):                     # stands in for VQGAN / ViTVQGAN: only `.codebook` matters here
            def __init__(self):
                super().__init__()
                self.codebook = ref_cls(512, 32, 0.4)

        model = Holder()
        before = {k: v.clone() for k, v in model.state_dict().items()}
        patch_reference_model(model)
        assert type(model.codebook) is ours
        assert (model.codebook.codebook_size, model.codebook.codebook_dim, model.codebook.beta) == (512, 32, 0.4)
        after = model.state_dict()
        assert list(after) == list(before) == ["codebook.embedding.weight"]
        assert torch.equal(after["codebook.embedding.weight"], before["codebook.embedding.weight"])
        z = torch.randn(1, 4, 32) if ours is b_vit.Codebook else torch.randn(1, 32, 2, 2)
        with pytest.raises(RuntimeError, match="no CPU path"):
            model.codebook(z)


def test_projection_fusion_host_logic(lib):
    """SURVEY.md 8(f) rank 1, host side only: which Linear shapes the fused pre_quant covers, argument errors of the new
    entry points (they are refused before any CUDA call), and which wrapper methods patch_reference_model rebinds."""
    import torch.nn as nn
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from standins import ViTVQGANStandIn, VQGANStandIn
    from vq_b200 import projected
    from vq_b200.integration import patch_reference_model
    from vq_b200.vitvqgan import Codebook as Vit
    from vq_b200.vqgan import Codebook as Vqgan
    cdll = lib.load()
    assert [c for c in (32, 64, 96, 128, 256, 512, 640, 768, 832, 1024) if projected.prequant_supported(c, 32)] == \
        [64, 128, 256, 512, 640, 768]
    assert not projected.prequant_supported(512, 64) and not projected.prequant_supported(512, 16)
    # unsupported shape / wrong form / misaligned rows: VQ_ERR_ARG with a message, no device needed
    # (weight, cb, K, D, form, beta, flags, n_elem_total, z_q, idx, loss, hist, stats, saved_zn, saved_denom, seg_sums, z_out,
    #  ws, ws_bytes, stream)
    args_tail = (None, 1, 512, 32, 0, ctypes.c_float(0.25), 0, 32, None, 1, None, None, None, None, None, None, None, 1, 1, None)
    assert len(args_tail) + 5 == len(lib.SIGNATURES["vq_forward_projected"][1])
    assert cdll.vq_forward_projected(16, 96, 16, None, 1, *args_tail) == 1
    assert b"unsupported" in cdll.vq_last_error()
    bad_form = list(args_tail); bad_form[4] = 1
    assert cdll.vq_forward_projected(16, 128, 16, None, 1, *bad_form) == 1 and b"ViT form" in cdll.vq_last_error()
    assert cdll.vq_forward_projected(20, 128, 16, None, 1, *args_tail) == 1 and b"aligned" in cdll.vq_last_error()
    assert cdll.vq_project_codebook(None, None, 512, 32, 1, 16, None, 128, 16, None) == 1       # normalise=1 without cb
    assert cdll.vq_gather_projected(16, 64, 4, 0, 16, 512, 130, 0, 16, None, None) == 1         # token-major needs C % 4 == 0
    assert cdll.vq_gather_projected(16, 8, 4, 0, 16, 512, 128, 0, 16, None, None) == 1          # token_bits
    # the modules know what they cover
    assert Vit(512, 32).supports_fused_pre_quant(nn.Linear(128, 32))
    assert not Vit(512, 32).supports_fused_pre_quant(nn.Linear(100, 32))
    assert not Vit(512, 64).supports_fused_pre_quant(nn.Linear(128, 64))
    assert not Vqgan(512, 64).supports_fused_pre_quant(nn.Conv2d(64, 64, 1))
    with pytest.raises(RuntimeError, match="no CPU path"):
        Vit(512, 32).forward_projected(torch.randn(2, 4, 128), nn.Linear(128, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        Vit(512, 32).decode_projected(torch.zeros(1, 4, dtype=torch.long), nn.Linear(32, 128))
    # under autocast (the reference's pre_quant then runs in bf16) or for non-fp32 rows the fused wrappers keep the two calls
    from vq_b200 import integration
    assert integration._fusable_now(torch.zeros(2, 4)) and not integration._fusable_now(torch.zeros(2, 4, dtype=torch.bfloat16))
    with torch.autocast("cpu", dtype=torch.bfloat16):
        assert not integration._fusable_now(torch.zeros(2, 4))
    # wrapper rebinding
    names = lambda m: tuple(getattr(getattr(m, a), "__func__", getattr(m, a)).__name__ for a in ("forward", "encode_imgs", "decode_indices"))
    assert names(patch_reference_model(ViTVQGANStandIn(48, 128, 512, 32), form="vit", fuse_projections=True)) == \
        ("_forward_vit_fused", "_encode_imgs_vit_fused", "_decode_indices_fused")
    assert names(patch_reference_model(ViTVQGANStandIn(48, 100, 512, 32), form="vit", fuse_projections=True)) == \
        ("forward", "_encode_imgs_vit", "_decode_indices_fused")          # Linear(100, 32): the reference's two calls stay
    assert names(patch_reference_model(VQGANStandIn(3, 64, 256), form="vqgan", fuse_projections=True)) == \
        ("forward", "_encode_imgs_vqgan", "_decode_indices_fused")
    assert names(patch_reference_model(ViTVQGANStandIn(48, 128, 512, 32), form="vit")) == ("forward", "_encode_imgs_vit", "decode_indices")


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="the reference is only present in the build container")
def test_patch_reference_model_on_the_unmodified_reference_vqgan():
    """INTEGRATION.md route 1 on the reference's own wrapper class (models/vqgan.py:221-251; ViTVQGAN cannot be constructed at
    all: its FeedForward is broken): the codebook is swapped, checkpoints keep their keys, encode_imgs /
    decode_indices are rebound (fast encode, projected decode), forward stays the reference's."""
    sys.path.insert(0, "/root/reference")
    sys.dont_write_bytecode = True
    from models.vqgan import VQGAN
    from vq_b200 import vqgan as b_vqgan
    from vq_b200.integration import patch_reference_model
    model = VQGAN(64, 512)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    reference_forward = model.forward.__func__
    patch_reference_model(model, fuse_projections=True)
    assert type(model.codebook) is b_vqgan.Codebook and model.codebook.codebook_dim == 64
    after = model.state_dict()
    assert list(after) == list(before) and all(torch.equal(after[k], before[k]) for k in before)
    assert model.forward.__func__ is reference_forward
    assert model.encode_imgs.__func__.__name__ == "_encode_imgs_vqgan"
    assert model.decode_indices.__func__.__name__ == "_decode_indices_fused"
    with pytest.raises(RuntimeError, match="CUDA"):
        with torch.no_grad():
            model.decode_indices(torch.zeros(1, 16, dtype=torch.long))


def test_header_is_plain_c(tmp_path):
    """include/vq_b200.h is a C ABI: it must compile as C99 (no C++ types, no torch types in any signature)."""
    src = tmp_path / "h.c"
    src.write_text('#include "vq_b200.h"\nint main(void) { return VQ_ABI_VERSION > 0 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
