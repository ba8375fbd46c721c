"""One optimisation step of the reference's training loop (exp/cips3d/scripts/train.py:326-491) on the cips3d_b200 modules,
timed with CUDA events: the train-step configurations of BASELINE.json (configs 3-5).

    python tools/bench_train_step.py --config 3            # r128, 16 img/GPU, GeneratorNerfINR + aux D, train_aux_img
    python tools/bench_train_step.py --config 4            # r256, 16 img/GPU, freeze-NeRF recipe (train_ffhq_high) [--aux: BASELINE's literal "incl. aux discriminator"]
    python tools/bench_train_step.py --config 5            # r256, 8 img/GPU, finetune recipe (freeze-NeRF G, diffaug D)
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train_step.py --config 3     # DDP over NCCL

The step is the reference's, line for line in structure:
  D step: G under no_grad (fused kernels) -> D(real) with real.requires_grad -> R1 penalty through autograd.grad(create_graph)
          (double backward through the native bias_act / upfirdn2d ops) -> D(fake) -> softplus losses -> clip + Adam
  G step: G with grad_points (part_grad_forward) -> D -> softplus -> clip + Adam -> EMA of the state_dict
`--optim fused` (default) uses cips3d_b200.FusedAdam / EMA (two launches per optimiser step), `--optim torch` the reference's
torch.optim.Adam + clip_grad_norm_ + comm_model_utils-style EMA loop, for an A/B on the optimiser tail.
Real images are synthetic (U(-1,1)); weights are the synthetic init-like set.  One JSON line per run."""
import argparse
import copy
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

import cips3d_b200
from oracle import cips3d_oracle as O       # weights recipe only (synthetic_state_dict); nothing of the oracle is timed

CONFIGS = {
    # BASELINE.json configs[2..4]; per-GPU batch, resolution, generator class, aux images, diffaug, grad/forward points
    3: dict(res=128, batch=16, frozen=False, aux=True, diffaug=False, grad_points=256, forward_points=256, warmup_D=False),
    4: dict(res=256, batch=16, frozen=True, aux=False, diffaug=True, grad_points=256, forward_points=256, warmup_D=True),
    5: dict(res=256, batch=8, frozen=True, aux=False, diffaug=True, grad_points=256, forward_points=256, warmup_D=True),
}
G_KWARGS = dict(O.G_KWARGS)
GRAD_CLIP, R1_LAMBDA, D_REG_EVERY, BETAS = 10, 10.0, 1, (0.0, 0.999)        # ffhq_exp.yaml:159-171


def requires_grad(model, flag):
    for p in model.parameters():
        p.requires_grad_(flag)


def build_step(cfg, dev, optim="fused", cips_backend="torch", ddp=False, local=0, g_cfg=None, d_kwargs=None, film_backend="torch",
               integrate_backend="torch", linear_backend="torch"):
    """Modules, optimisers and the step closure of one configuration on device `dev`.  Separate from main() so that
    tests/test_train_step_cpu.py can execute the very same step on the CPU emulation of the kernels (tiny sizes)."""
    class _A:        # the two switches the step reads
        pass
    args = _A()
    args.optim, args.cips_backend = optim, cips_backend
    G_cls = cips3d_b200.GeneratorNerfINR_freeze_NeRF if cfg["frozen"] else cips3d_b200.GeneratorNerfINR
    gcfg = {k: (dict(v) if isinstance(v, dict) else v) for k, v in (g_cfg or O.G_CFG).items()}
    G = G_cls(**gcfg, device=dev).to(dev)
    G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234, sigma_bias=0.3))
    D = cips3d_b200.Discriminator_MultiScale_Aux(**dict(dict(diffaug=cfg["diffaug"], max_size=1024, channel_multiplier=2,
                                                             first_downsample=False, stddev_group=0), **(d_kwargs or {}))).to(dev)
    G.inr_net.train_backend = args.cips_backend
    G.train_integrate = integrate_backend      # volume integration of the autograd graph as the native op (csrc/integrate_ops.cu)
    for m in G.modules():           # FiLM + sine of the NeRF branch's autograd graph as the native op (csrc/film_ops.cu)
        if isinstance(m, cips3d_b200.FiLMLayer):
            m.fused_film = film_backend == "fused"
            m.fused_linear = linear_backend == "fused"      # needs fused_film (the z -> sin step is the native op then)
    G.siren.fused_linear = linear_backend == "fused"
    G_ema = copy.deepcopy(G)
    G_run, D_run = G, D
    if ddp:
        from torch.nn.parallel import DistributedDataParallel as DDP
        G_run = DDP(G, device_ids=[local], find_unused_parameters=True, broadcast_buffers=False)     # train.py:235-236
        D_run = DDP(D, device_ids=[local], find_unused_parameters=True, broadcast_buffers=False)
    lr_g, lr_d = (1e-4, 5e-4) if cfg["frozen"] else (2e-4, 2e-3)
    if args.optim == "fused":
        opt_G = cips3d_b200.FusedAdam([{"params": G_run.parameters(), "initial_lr": lr_g}], lr=lr_g, betas=BETAS)
        opt_D = cips3d_b200.FusedAdam([{"params": D_run.parameters(), "initial_lr": lr_d}], lr=lr_d, betas=BETAS)
        ema = cips3d_b200.EMA(source=G, target=G_ema, decay=0.999, start_itr=0)
    else:
        opt_G = torch.optim.Adam([{"params": G_run.parameters(), "initial_lr": lr_g}], lr=lr_g, betas=BETAS, weight_decay=0)
        opt_D = torch.optim.Adam([{"params": D_run.parameters(), "initial_lr": lr_d}], lr=lr_d, betas=BETAS, weight_decay=0)
        ema = None
    R, B = cfg["res"], cfg["batch"]
    dummy = torch.tensor([0.0], device=dev)

    def step(it):
        alpha = min(1, it / 10000) if cfg["warmup_D"] else 1.0                         # train.py:329-332
        nerf_noise = 0.0 if cfg["frozen"] else max(0.0, 1.0 - it / 5000.0)             # nerf_noise_disable (train_ffhq_high)
        aux_reg = cfg["aux"]
        real = torch.rand(B, 3, R, R, device=dev) * 2 - 1
        # ---- TRAIN DISCRIMINATOR (train.py:334-438)
        requires_grad(G_run, False)
        requires_grad(D_run, True)
        with torch.no_grad():
            zs = G.get_zs(B)
            fp = cfg["forward_points"] ** 2 if R >= 256 and cfg["forward_points"] else None
            gen_imgs, _ = G_run(zs, img_size=R, nerf_noise=nerf_noise, return_aux_img=aux_reg, forward_points=fp, grad_points=None,
                                **G_KWARGS)
        if aux_reg:
            real = torch.cat([real, real], dim=0)
        real.requires_grad_()
        r_preds, _, _ = D_run(real, alpha=alpha, use_aux_disc=aux_reg)
        if R1_LAMBDA > 0 and it % D_REG_EVERY == 0:
            grad_real = torch.autograd.grad(outputs=r_preds.sum(), inputs=real, create_graph=True)[0]
            grad_penalty = grad_real.flatten(start_dim=1).square().sum(dim=1, keepdim=True)
            grad_penalty = 0.5 * R1_LAMBDA * grad_penalty * D_REG_EVERY + 0.0 * r_preds
        else:
            grad_penalty = dummy
        g_preds, _, _ = D_run(gen_imgs, alpha=alpha, use_aux_disc=aux_reg)
        d_loss = (F.softplus(g_preds) + F.softplus(-r_preds) + grad_penalty).mean()
        opt_D.zero_grad()
        d_loss.backward()
        if args.optim == "fused":
            opt_D.step(max_norm=GRAD_CLIP)
        else:
            torch.nn.utils.clip_grad_norm_(D_run.parameters(), GRAD_CLIP)
            opt_D.step()
        # ---- TRAIN GENERATOR (train.py:440-491)
        requires_grad(G_run, True)
        requires_grad(D_run, False)
        zs = G.get_zs(B)
        gp = cfg["grad_points"] ** 2 if cfg["grad_points"] else None
        gen_imgs, _ = G_run(zs, img_size=R, nerf_noise=nerf_noise, return_aux_img=aux_reg, grad_points=gp, forward_points=None, **G_KWARGS)
        g_preds, _, _ = D_run(gen_imgs.to(torch.float32), alpha=alpha, use_aux_disc=aux_reg)
        g_loss = F.softplus(-g_preds).mean()
        g_loss.backward()
        if args.optim == "fused":
            opt_G.step(max_norm=GRAD_CLIP, ema=ema, itr=it, zero_grad=True)
        else:
            torch.nn.utils.clip_grad_norm_(G_run.parameters(), GRAD_CLIP)
            opt_G.step()
            opt_G.zero_grad()
            sd, td = G.state_dict(), G_ema.state_dict()
            with torch.no_grad():
                for k in sd:
                    td[k].data.copy_(td[k].data * 0.999 + sd[k].data * (1 - 0.999))
        return d_loss.detach(), g_loss.detach()

    return step, dict(G=G, D=D, G_ema=G_ema, opt_G=opt_G, opt_D=opt_D, G_cls=G_cls)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--aux", action="store_true", help="config 4 with train_aux_img / aux discriminator on (BASELINE's wording)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--optim", default="fused", choices=["fused", "torch"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--cips-backend", default="torch", choices=["torch", "fused"],
                    help="fused: CIPSNet.train_backend = 'fused' (native forward + backward chain, fp16 library GEMMs for dW)")
    ap.add_argument("--film-backend", default="torch", choices=["torch", "fused"],
                    help="fused: FiLMLayer.fused_film = True (native FiLM+sine forward/backward in the NeRF autograd graph; configs with NeRF gradients)")
    ap.add_argument("--integrate-backend", default="torch", choices=["torch", "fused"],
                    help="fused: GeneratorNerfINR.train_integrate = 'fused' (native fancy_integration forward/backward in the NeRF autograd graph)")
    ap.add_argument("--linear-backend", default="torch", choices=["torch", "fused"],
                    help="fused: the NeRF field's per-point linears (forward + data gradient) on the tcgen05 split-fp16 GEMM (ops.points_linear)")
    ap.add_argument("--tf32", action="store_true", help="allow TF32 in the torch autograd GEMMs / cuDNN convs of the training graph "
                    "(NOT the reference's numerics: torch defaults to fp32 matmuls); measures what the library path can give")
    ap.add_argument("--no-cudnn-tf32", action="store_true", help="force fp32 cuDNN convolutions (torch's default -- and the reference's, "
                    "torch >= 1.7 -- is cudnn.allow_tf32 = True, which this tool keeps)")
    ap.add_argument("--profile", default=None, help="write a torch.profiler kernel table of 2 steps to this file (after the timing)")
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
    torch.backends.cudnn.allow_tf32 = not args.no_cudnn_tf32
    cfg = dict(CONFIGS[args.config])
    if args.aux:
        cfg["aux"] = True
    if args.batch:
        cfg["batch"] = args.batch
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    ddp = world > 1
    if ddp:
        torch.distributed.init_process_group("nccl")
    torch.manual_seed(1234 + rank)
    step, mods = build_step(cfg, dev, args.optim, args.cips_backend, ddp, local, film_backend=args.film_backend, integrate_backend=args.integrate_backend,
                             linear_backend=args.linear_backend)
    G_cls = mods["G_cls"]
    R, B = cfg["res"], cfg["batch"]

    for it in range(args.warmup):
        step(it)
    if ddp:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(args.steps):
        dl, gl = step(args.warmup + it)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if ddp:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    nccl = None
    if args.profile:          # every rank steps (DDP collectives); rank 0 records
        from torch.profiler import profile, ProfilerActivity
        import contextlib
        ctx = profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) if rank == 0 else contextlib.nullcontext()
        with ctx as prof:
            for it in range(2):
                step(args.warmup + args.steps + it)
            torch.cuda.synchronize()
        if rank == 0:
            ka = prof.key_averages()
            with open(args.profile, "w") as f:
                f.write(ka.table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
            dt = lambda e: getattr(e, "self_device_time_total", getattr(e, "self_cuda_time_total", 0.0))     # noqa: E731
            nccl = dict(nccl_kernel_ms_per_step=sum(dt(e) for e in ka if "nccl" in e.key.lower()) / 2e3,
                        all_gpu_kernel_ms_per_step=sum(dt(e) for e in ka) / 2e3,
                        note="sum of kernel durations on rank 0 over 2 profiled steps / 2; the all-reduce kernels run on DDP's "
                             "communication stream concurrently with backward, so their sum is exposed only where it exceeds the "
                             "step's compute (compare ms_per_step across n_gpus)")
    if rank == 0:
        print(json.dumps(dict(
            metric="train step (D step + G step) images/s", value=B * world / ms.item() * 1e3, unit="images/s", ms_per_step=ms.item(),
            n_gpus=world, steps=args.steps, warmup=args.warmup,
            config=dict(baseline_config=args.config, resolution=R, batch_per_gpu=B, generator=G_cls.__name__, train_aux_img=cfg["aux"],
                        diffaug=cfg["diffaug"], grad_points=cfg["grad_points"], optim=args.optim, tf32_autograd=bool(args.tf32),
                        cudnn_tf32=bool(torch.backends.cudnn.allow_tf32), cips_backend=args.cips_backend, film_backend=args.film_backend, integrate_backend=args.integrate_backend, linear_backend=args.linear_backend,
                        note="G forward under no_grad runs the fused kernels; the G step's graph uses the native training ops the "
                             "*_backend flags select (DESIGN.md 4.10-4.14), torch CUDA ops otherwise; D convolutions are the library's implicit "
                             "GEMM behind ops.conv2d's autograd structure; D's bias_act / blur are native"),
            d_loss=float(dl), g_loss=float(gl), comm=nccl, finite=bool(math.isfinite(float(dl)) and math.isfinite(float(gl))))))
    if ddp:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
