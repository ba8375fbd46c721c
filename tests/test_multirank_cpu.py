"""world_size-2 gloo test of the N>1 path's host logic: per-rank seeding/sharding of the image batch and the
max-over-ranks reduction bench.py uses (no GPU; the kernels themselves are covered by -m gpu tests)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cips3d_b200
    from oracle import cips3d_oracle as O
    # every rank builds the same generator from the same synthetic weights ...
    G = cips3d_b200.GeneratorNerfINR(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in O.G_CFG.items()}, device="cpu")
    G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234))
    digest = torch.stack([p.detach().double().sum() for p in G.parameters()]).sum().reshape(1)
    gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, digest)
    same_weights = all(torch.equal(gathered[0], g) for g in gathered)
    # ... and draws ITS OWN latents (bench.py: torch.manual_seed(1000 + rank)); shards must differ
    torch.manual_seed(1000 + rank)
    zs = G.get_zs(4)
    zsum = zs["z_nerf"].sum().reshape(1).double()
    zg = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(zg, zsum)
    # weak scaling bookkeeping: images = per-rank batch * world; time = MAX over ranks
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        out.put(dict(same_weights=same_weights, distinct_latents=not torch.equal(zg[0], zg[1]), tmax=t.item(),
                     global_batch=4 * world))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = out.get()
    assert res["same_weights"] and res["distinct_latents"]
    assert res["tmax"] == 11.0 and res["global_batch"] == 8
