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


def _worker(rank, world, port, exchange, q):
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
        w = vo.make_codebook("vit", K, D, 0).to(dev)
        sharded = vq_dist.ShardedQuantiser("vit", BETA, world_size=world, exchange=exchange, graphs=graphs)
        single = vq_dist.ShardedQuantiser("vit", BETA, world_size=1)
        ok = True
        msgs = []
        zbuf = [torch.empty(B // world, N_TOK, D, device=dev) for _ in range(2)]       # fixed buffers: graphs replay on them
        ubuf = [torch.empty(B // world, N_TOK, D, device=dev) for _ in range(2)]
        for step in range(8 if graphs else 4):          # both slots twice (graphs: eager, capture, then replays)
            zg = vo.make_latents((B, N_TOK, D), 100 + step).to(dev)
            ug = vo.make_latents((B, N_TOK, D), 200 + step).to(dev)
            zbuf[step & 1].copy_(vq_dist.shard_batch(zg, rank, world))
            ubuf[step & 1].copy_(vq_dist.shard_batch(ug, rank, world))
            out = sharded.step(zbuf[step & 1], ubuf[step & 1], w)
            out = {k: v.clone() for k, v in out.items()}
            ref = single.step(zg, ug, w)
            per = B // world * N_TOK
            rows = slice(rank * per, (rank + 1) * per)
            checks = {
                "grad_weight": torch.equal(out["grad_weight"], ref["grad_weight"]),
                "loss": float(out["loss"]) == float(ref["loss"]),
                "histogram": torch.equal(out["histogram"].to(torch.int64), ref["histogram"].to(torch.int64)),
                "indices": torch.equal(out["indices"], ref["indices"][rows]),
                "z_q": torch.equal(out["z_q"].reshape(-1, D), ref["z_q"].reshape(-1, D)[rows]),
                "grad_z": torch.equal(out["grad_z"].reshape(-1, D), ref["grad_z"].reshape(-1, D)[rows]),
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
