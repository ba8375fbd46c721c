"""HBM roofline of the fused optimiser tail (csrc/optim_ops.cu) on the reference's two models, next to the reference's
own sequence (clip_grad_norm_ + torch.optim.Adam.step + zero_grad + EMA.update over the state_dict) on the same GPU.
Algorithmic bytes per parameter: norm pass 4 (read g) + update 20 read (g, p, m, v, ema) + 16 written (p, m, v, ema)
[+ 4 written when zero_grad writes zeros]; without EMA 16 + 12.  CUDA events, L2 flushed between repetitions."""
import copy, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cips3d_b200
from oracle import cips3d_oracle as O
from _util import build_generator

dev = "cuda:0"
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
flush = torch.empty(64 * 1024 * 1024, device=dev)


def timeit(fn, reps=20):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def bench(name, model, with_ema):
    n = sum(p.numel() for p in model.parameters())
    for p in model.parameters():
        p.grad = torch.randn_like(p) * 0.01
    m_ema = copy.deepcopy(model) if with_ema else None
    fused = cips3d_b200.FusedAdam(model.parameters(), lr=2e-4, betas=(0.0, 0.999))
    ema = cips3d_b200.EMA(model, m_ema, decay=0.999, start_itr=0) if with_ema else None
    ms_f = timeit(lambda: fused.step(max_norm=10.0, ema=ema, itr=1, zero_grad=False))
    ref = torch.optim.Adam(model.parameters(), lr=2e-4, betas=(0.0, 0.999))

    def ref_step():
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        ref.step()
        if with_ema:
            sd, td = model.state_dict(), m_ema.state_dict()
            with torch.no_grad():
                for k in sd:
                    td[k].data.copy_(td[k].data * 0.999 + sd[k].data * (1 - 0.999))
    ms_r = timeit(ref_step)
    bytes_ = n * (4 + (36 if with_ema else 28))
    print(json.dumps(dict(op=f"clip+adam{'+ema' if with_ema else ''} ({name})", params=n, tensors=len(list(model.parameters())),
                          fused_ms=ms_f, fused_gbs=bytes_ / ms_f / 1e6, frac=bytes_ / ms_f / 1e6 / peak,
                          torch_ms=ms_r, speedup=ms_r / ms_f)))


G = build_generator(dev, O.synthetic_state_dict(O.generator_template(), seed=1)).train()
bench("GeneratorNerfINR", G, True)
D = cips3d_b200.Discriminator_MultiScale_Aux(diffaug=False, max_size=256, channel_multiplier=2, first_downsample=False, stddev_group=0).to(dev)
bench("Discriminator_MultiScale_Aux", D, False)
print(json.dumps(dict(hbm_peak_gbs=peak, note="algorithmic bytes / CUDA-event median; L2 flushed between reps; launch overhead of the Python table build included")))
