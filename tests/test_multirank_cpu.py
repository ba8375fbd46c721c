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


class _StubG:                       # eval / get_zs / __call__ -> (imgs, pitch_yaw), like GeneratorNerfINR in gen_images
    def eval(self):
        return self

    def get_zs(self, b):
        return {"z": torch.randn(b, 4)}

    def __call__(self, zs, forward_points=None, **kw):
        b, r = zs["z"].shape[0], kw["img_size"]
        return torch.tanh(torch.randn(b, r, r, 3)).permute(0, 3, 1, 2), torch.zeros(b, 2)


def _gen_worker(rank, world, port, fake_dir, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _emu import emulated
    import cips3d_b200.inference as inf
    torch.manual_seed(50 + rank)
    with emulated(async_mode=0):    # the uint8 conversion kernel runs on the CPU emulation; the sharding logic is what is tested
        n = inf.gen_images(rank, world, _StubG(), {"fov": 12}, fake_dir, num_imgs=10, img_size=8, batch_size=4, ext="png")
    out.put((rank, n))
    dist.destroy_process_group()


def test_two_rank_gen_images_shards_the_dump(tmp_path):
    """gen_images.py:30-73 under world_size 2: rank 0 creates the directory, both ranks pass the barrier, every rank writes
    its interleaved share (idx_b * batch_size + idx_i * world_size + rank) and together they cover 0..N-1 exactly once."""
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    fake = str(tmp_path / "fake")
    procs = [ctx.Process(target=_gen_worker, args=(r, 2, port, fake, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    counts = dict(out.get() for _ in range(2))
    assert counts == {0: 6, 1: 6}                       # ceil(10 / 4) = 3 iterations x (4 // 2) images per rank
    assert sorted(os.listdir(fake)) == [f"{i:0>5}.png" for i in range(12)]
