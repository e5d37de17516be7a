"""Two ranks on two GPUs (NCCL only carries the IPC handles and the test's own gathers): the token-sharded step
with the fused peer-memory exchange must give, on every rank, exactly the single-GPU result on the concatenated
batch -- grad_weight, histogram and loss bit-identical (integer sums), z_q / indices / grad_z equal to the matching
rows.  Skipped on a box with fewer than two GPUs (run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

K, D, B, N_TOK, BETA = 8192, 32, 64, 1024, 0.25


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, exchange, q, form="vit"):
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "attention-models_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    graphs = exchange == "peer-graphs"
    if exchange.startswith("peer-"):       # the library reads the mode once, when it is loaded (fresh process here)
        os.environ["VQ_EXCHANGE_ONE_SHOT"] = "1" if exchange == "peer-one-shot" else "0"
        exchange = "peer"
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import vq_oracle as vo
        from vq_b200 import _lib
        from vq_b200 import dist as vq_dist
        kk, dd = (K, D) if form == "vit" else (1024, 64)
        full = (B, N_TOK, dd) if form == "vit" else (B, dd, 16, 16)          # NCHW latents for the CNN form
        local = (B // world,) + full[1:]
        w = vo.make_codebook(form, kk, dd, 0).to(dev)
        sharded = vq_dist.ShardedQuantiser(form, BETA, world_size=world, exchange=exchange, graphs=graphs)
        single = vq_dist.ShardedQuantiser(form, BETA, world_size=1)
        ok = True
        msgs = []
        zbuf = [torch.empty(*local, device=dev) for _ in range(2)]       # fixed buffers: graphs replay on them
        ubuf = [torch.empty(*local, device=dev) for _ in range(2)]
        for step in range(8 if graphs else 4):          # both slots twice (graphs: eager, capture, then replays)
            zg = vo.make_latents(full, 100 + step).to(dev)
            ug = vo.make_latents(full, 200 + step).to(dev)
            zbuf[step & 1].copy_(vq_dist.shard_batch(zg, rank, world))
            ubuf[step & 1].copy_(vq_dist.shard_batch(ug, rank, world))
            out = sharded.step(zbuf[step & 1], ubuf[step & 1], w)
            out = {k: v.clone() for k, v in out.items()}
            ref = single.step(zg, ug, w)
            per = out["indices"].numel()
            rows = slice(rank * per, (rank + 1) * per)
            items = slice(rank * (B // world), (rank + 1) * (B // world))
            checks = {
                "grad_weight": torch.equal(out["grad_weight"], ref["grad_weight"]),
                "loss": float(out["loss"]) == float(ref["loss"]),
                "histogram": torch.equal(out["histogram"].to(torch.int64), ref["histogram"].to(torch.int64)),
                "indices": torch.equal(out["indices"], ref["indices"][rows]),
                "z_q": torch.equal(out["z_q"], ref["z_q"][items]),
                "grad_z": torch.equal(out["grad_z"], ref["grad_z"][items]),
                "no_timeout": int(out["stats"][_lib.STAT_PEER_TIMEOUT]) == 0,
            }
            for name, good in checks.items():
                if not good:
                    ok = False
                    msgs.append(f"step {step}: {name} differs")
        torch.cuda.synchronize()
        sharded.close()
        q.put((rank, ok, msgs))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("exchange", ["peer-one-shot", "peer-two-shot", "peer-graphs", "collective"])
def test_sharded_step_equals_single_gpu_on_global_batch(exchange, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    if world > 2 and exchange in ("peer-one-shot", "collective"):
        pytest.skip("covered at world 2")
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, exchange, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, msgs in results:
        assert ok, (rank, msgs)


def _run(world, target, *extra):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, world, port, *extra, q) if target is not _worker else (r, world, port, extra[0], q, *extra[1:]))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, msgs in results:
        assert ok, (rank, msgs)


@pytest.mark.parametrize("exchange", ["peer-two-shot", "collective"])
def test_sharded_step_vqgan_form_two_gpus(exchange):
    """The CNN form (NCHW in / out): exchange kernel on the caller's stream, token backward on a side stream."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, _worker, exchange, "vqgan")


def _timeout_worker(rank, world, port, q):
    import time

    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "attention-models_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import vq_oracle as vo
        from vq_b200 import _lib
        from vq_b200 import dist as vq_dist
        w = vo.make_codebook("vit", K, D, 0).to(dev)
        sharded = vq_dist.ShardedQuantiser("vit", BETA, world_size=world, exchange="peer", peer_timeout_s=1.0)
        single = vq_dist.ShardedQuantiser("vit", BETA, world_size=1)
        msgs = []

        def one(step, delay):
            zg = vo.make_latents((8, N_TOK, D), 300 + step).to(dev)
            ug = vo.make_latents((8, N_TOK, D), 400 + step).to(dev)
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 1 and delay:
                time.sleep(delay)                # a late rank: dataloader stall, checkpointing, ...
            out = sharded.step(vq_dist.shard_batch(zg, rank, world).contiguous(), vq_dist.shard_batch(ug, rank, world).contiguous(), w)
            out = {k: v.clone() for k, v in out.items()}
            torch.cuda.synchronize()
            return out, single.step(zg, ug, w)

        # 1. a delay well inside the time-out: the early rank waits, the result is exact
        out, ref = one(0, 0.2)
        if not torch.equal(out["grad_weight"], ref["grad_weight"]) or int(out["stats"][_lib.STAT_PEER_TIMEOUT]) != 0:
            msgs.append("short delay: wrong result or spurious time-out")
        # 2. a delay beyond the time-out: FATAL and loud on both ranks -- NaN results, counter, host flag, next step raises
        out, _ = one(1, 3.0)
        if not (torch.isnan(out["grad_weight"]).all() and torch.isnan(out["loss"]).all()):
            msgs.append("time-out: grad_weight / loss not poisoned")
        if int(out["stats"][_lib.STAT_PEER_TIMEOUT]) == 0:
            msgs.append("time-out: stats counter not bumped")
        raised = False
        try:
            one(2, 0.0)
        except vq_dist.PeerTimeoutError:
            raised = True
        if not raised:
            msgs.append("time-out: the next step did not raise PeerTimeoutError")
        # 3. collective resync: the exchange works again and is exact
        sharded.resync()
        out, ref = one(3, 0.0)
        if not torch.equal(out["grad_weight"], ref["grad_weight"]) or float(out["loss"]) != float(ref["loss"]):
            msgs.append("after resync: wrong result")
        sharded.close()
        q.put((rank, not msgs, msgs))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_peer_timeout_is_fatal_and_recoverable():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, _timeout_worker)
