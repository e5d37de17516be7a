"""N > 1 host logic on CPU: two gloo ranks exchange the packed int64 buffer of vq_b200.dist.

The CUDA kernels cannot run here, so each rank produces what vq_backward_tokens would hand over for its
shard (fixed-point segment sums, histogram, loss partial) from the oracle, the ranks all-reduce the packed
buffer exactly as ShardedQuantiser.step does, and the result must equal the single-process reference on the
concatenated batch (multi-GPU oracle = single-process reference on the global batch).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SEG_SHIFT, LOSS_SHIFT = 30, 24     # VQ_SEG_SHIFT / VQ_LOSS_SHIFT of csrc/vq_common.cuh
K, D, B, N_TOK, BETA = 64, 16, 8, 32, 0.25


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_contribution(form, z, weight):
    """What one rank's kernels produce for its shard: (seg_sums int64 K*D+K, hist int32, stats)."""
    from oracle import vq_oracle as vo
    from vq_b200._lib import STAT_LOSS_FIXED, STAT_NONFINITE, STATS_LEN
    zn = vo.unit_rows(z).reshape(-1, D)
    out = vo.quantise(form, z, weight, BETA)
    idx = out.indices.reshape(-1)
    q = vo.unit_rows(weight[idx])
    contrib = torch.round((q - zn).double() * float(1 << SEG_SHIFT)).to(torch.int64)
    seg = torch.zeros(K * D + K, dtype=torch.int64)
    seg[: K * D] = torch.zeros(K, D, dtype=torch.int64).index_add_(0, idx, contrib).reshape(-1)
    hist = torch.bincount(idx, minlength=K).to(torch.int32)
    stats = torch.zeros(STATS_LEN, dtype=torch.int64)
    row_loss = ((q - zn) ** 2).sum(-1)
    stats[STAT_LOSS_FIXED] = torch.round(row_loss.double() * float(1 << LOSS_SHIFT)).to(torch.int64).sum()
    stats[STAT_NONFINITE] = 0
    return seg, hist, stats, idx


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "attention-models_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import vq_oracle as vo
        from vq_b200 import dist as vq_dist
        torch.set_num_threads(1)
        weight = vo.make_codebook("vit", K, D, 0)
        z_global = vo.make_latents((B, N_TOK, D), 1)
        z_local = vq_dist.shard_batch(z_global, rank, world)
        assert z_local.shape[0] == B // world
        seg, hist, stats, idx = _shard_contribution("vit", z_local, weight)
        pack = vq_dist.PackedReduce(K, D)
        buf = pack.allocate("cpu")
        pack.seg(buf).copy_(seg)
        pack.fill_side_channels(buf, hist, stats)
        pack.all_reduce(buf)
        q.put((rank, buf.numpy().copy(), idx.numpy().copy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_packed_all_reduce_equals_single_process_on_global_batch():
    from oracle import vq_oracle as vo
    from vq_b200 import dist as vq_dist
    from vq_b200._lib import STAT_LOSS_FIXED
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank holds the same reduced buffer
    assert np.array_equal(results[0][1], results[1][1])
    buf = torch.from_numpy(results[0][1])
    pack = vq_dist.PackedReduce(K, D)

    weight = vo.make_codebook("vit", K, D, 0)
    z_global = vo.make_latents((B, N_TOK, D), 1)
    seg_g, hist_g, stats_g, idx_g = _shard_contribution("vit", z_global, weight)
    # integer sums are exact and order-free: bit-identical to the unsharded run
    assert torch.equal(pack.seg(buf), seg_g)
    assert torch.equal(pack.hist(buf), hist_g.to(torch.int64))
    assert int(pack.stats_from(buf)[STAT_LOSS_FIXED]) == int(stats_g[STAT_LOSS_FIXED])
    # shards are contiguous image ranges in rank order
    assert np.array_equal(np.concatenate([results[0][2], results[1][2]]), idx_g.numpy())

    # the reduced sums give the reference's global-batch gradient and loss (1e-5 relative)
    up = vo.make_latents((B, N_TOK, D), 2)
    ref = vo.quantise_step("vit", z_global, weight, BETA, up)
    n_elem = B * N_TOK * D
    S = pack.seg(buf)[: K * D].double().reshape(K, D) / float(1 << SEG_SHIFT)
    nrm = weight.double().norm(dim=-1, keepdim=True).clamp_min(1e-12)
    y = weight.double() / nrm
    g = (2.0 / n_elem) * S
    grad_w = (g - y * (y * g).sum(-1, keepdim=True)) / nrm
    err = float((grad_w - ref.grad_weight.double()).norm() / ref.grad_weight.double().norm())
    assert err < 1e-5, err
    loss = (1.0 + BETA) * (int(stats_g[STAT_LOSS_FIXED]) / float(1 << LOSS_SHIFT)) / n_elem
    assert abs(loss - float(ref.loss)) / float(ref.loss) < 1e-5


def test_shard_batch_rejects_ragged_split():
    from vq_b200 import dist as vq_dist
    with pytest.raises(ValueError):
        vq_dist.shard_batch(torch.zeros(5, 4, 8), 0, 2)
    x = torch.arange(24.).reshape(4, 3, 2)
    assert torch.equal(torch.cat([vq_dist.shard_batch(x, r, 2) for r in range(2)]), x)
