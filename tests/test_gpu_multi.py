"""Index-sharded multi-GPU solve through the public API with NCCL (needs >= 2 GPUs; skipped otherwise).
The partition/gather logic itself is covered on CPU with gloo in test_host_api.py."""
import os

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, B, q):
    import torch.distributed as dist
    import lunar_module_ascent_trajectory_optimiser_b200 as lm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    p = lm.dispersed_params(B, seed=11)
    sol = lm.optimise_batch(p, device=rank, group=True)           # every rank gets the full result
    lo, hi = lm.shard_bounds(B, world, rank)
    local = lm.optimise_batch(lm.AscentParams(**{k: (v[lo:hi] if isinstance(v, torch.Tensor) else v)
                                                 for k, v in p.__dict__.items()}), device=rank)
    ok = (len(sol) == B and int((sol.status != 0).sum()) == 0 and
          torch.allclose(sol.tf[lo:hi].cpu(), local.tf.cpu(), rtol=1e-10, atol=0) and
          sol.states["y"].shape == (B, 200))
    q.put((rank, bool(ok), float(sol.tf_seconds[0])))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_solve_nccl_two_gpus(built_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    B = 700          # not divisible by 32 or by the world size times 32
    procs = [ctx.Process(target=_worker, args=(r, 2, 29533, B, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=300) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert [r[:2] for r in res] == [(0, True), (1, True)]
    assert abs(res[0][2] - 434.0276531) < 1e-5 and res[0][2] == res[1][2]
