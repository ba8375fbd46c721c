"""HBM roofline of the discriminator's two native ops at FFHQ r256 shapes (CUDA events, L2 flushed
between iterations by a 256 MB write)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cips3d_b200
from cips3d_b200 import ops
dev = "cuda:0"
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
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


out = []
k = cips3d_b200.discriminator.make_kernel([1, 3, 3, 1]).to(dev)
VARIANTS = os.environ.get("C3D_BLUR", "stream")
for (B, C, H) in [(16, 128, 256), (16, 256, 128), (16, 512, 64)]:
    x = torch.randn(B, C, H, H, device=dev)
    b = torch.randn(C, device=dev)
    ms = timeit(lambda: ops.bias_act(x, b))
    gb = 2 * x.numel() * 4 / 1e9
    out.append(dict(op="bias_act fwd", shape=[B, C, H, H], ms=ms, gbs=gb / ms * 1e3, frac=gb / ms * 1e3 / peak))
    y = ops.bias_act(x, b)
    ms = timeit(lambda: ops.bias_act(x, None, y, 3, 1))
    gb = 3 * x.numel() * 4 / 1e9
    out.append(dict(op="bias_act bwd(grad=1)", shape=[B, C, H, H], ms=ms, gbs=gb / ms * 1e3, frac=gb / ms * 1e3 / peak))
    for pad in ((2, 2), (1, 1)):
        ms = timeit(lambda: ops._upfirdn2d_raw(x, k, (1, 1), (1, 1), (pad[0], pad[1], pad[0], pad[1])))
        Ho = H + pad[0] + pad[1] - 3
        gb = (x.numel() + B * C * Ho * Ho) * 4 / 1e9
        out.append(dict(op=f"upfirdn2d blur pad{pad} (C3D_BLUR={VARIANTS})", shape=[B, C, H, H], ms=ms, gbs=gb / ms * 1e3, frac=gb / ms * 1e3 / peak))
# image export: 4*C bytes read + C written per pixel; the torch-op chain the reference runs beside it
for (B, H) in [(16, 256), (64, 256), (16, 512)]:
    nhwc = torch.tanh(torch.randn(B, H, H, 3, device=dev))
    for name, x in (("generator layout (NCHW view of NHWC)", nhwc.permute(0, 3, 1, 2)), ("NCHW", nhwc.permute(0, 3, 1, 2).contiguous())):
        ms = timeit(lambda: ops.image_to_u8(x))
        gb = x.numel() * 5 / 1e9
        out.append(dict(op=f"image_to_u8 save_image, {name}", shape=[B, 3, H, H], ms=ms, gbs=gb / ms * 1e3, frac=gb / ms * 1e3 / peak))

    def torch_chain():      # make_grid.norm_ip + save_image's conversion, batched (the reference does it image by image)
        t = x.clone()
        t.clamp_(min=-1, max=1)
        t.sub_(-1).div_(2)
        return t.mul(255).add_(0.5).clamp_(0, 255).permute(0, 2, 3, 1).to(torch.uint8)
    ms = timeit(torch_chain)
    out.append(dict(op="image_to_u8: torch op chain (clone, clamp_, sub_, div_, mul, add_, clamp_, permute+to)", shape=[B, 3, H, H], ms=ms,
                    gbs=x.numel() * 5 / 1e9 / ms * 1e3, frac=x.numel() * 5 / 1e9 / ms * 1e3 / peak))
# FiLM + sine of the NeRF autograd graph (csrc/film_ops.cu): forward 8 B/element, backward 12 B/element, against the torch ops
for (B, P, Cn) in [(16, 16384 * 24, 128), (16, 16384 * 24, 64)]:
    z = torch.randn(B, P, Cn, device=dev)
    gain, bias = torch.randn(B, 1, Cn, device=dev) * 5 + 30, torch.randn(B, 1, Cn, device=dev)
    dy = torch.randn(B, P, Cn, device=dev)
    ms = timeit(lambda: ops.film_sin(z, gain, bias), reps=10)
    out.append(dict(op="film_sin fwd", shape=[B, P, Cn], ms=ms, gbs=z.numel() * 8 / 1e9 / ms * 1e3, frac=z.numel() * 8 / 1e9 / ms * 1e3 / peak))
    ms_t = timeit(lambda: torch.sin(gain * z + bias), reps=10)
    out.append(dict(op="film_sin fwd: torch (mul, add, sin)", shape=[B, P, Cn], ms=ms_t, gbs=z.numel() * 8 / 1e9 / ms_t * 1e3, frac=z.numel() * 8 / 1e9 / ms_t * 1e3 / peak))
    zr, gr, br = z.clone().requires_grad_(), gain.clone().requires_grad_(), bias.clone().requires_grad_()
    y = ops.film_sin(zr, gr, br)
    ms = timeit(lambda: torch.autograd.grad(y, (zr, gr, br), dy, retain_graph=True), reps=10)
    out.append(dict(op="film_sin bwd (dz + dgain + dbias)", shape=[B, P, Cn], ms=ms, gbs=z.numel() * 12 / 1e9 / ms * 1e3, frac=z.numel() * 12 / 1e9 / ms * 1e3 / peak))
    yt = torch.sin(gr * zr + br)
    ms_t = timeit(lambda: torch.autograd.grad(yt, (zr, gr, br), dy, retain_graph=True), reps=10)
    out.append(dict(op="film_sin bwd: torch autograd", shape=[B, P, Cn], ms=ms_t, gbs=z.numel() * 12 / 1e9 / ms_t * 1e3, frac=z.numel() * 12 / 1e9 / ms_t * 1e3 / peak))
    del z, dy, zr, y, yt
    torch.cuda.empty_cache()
# volume integration of the NeRF autograd graph (csrc/integrate_ops.cu): per sample forward reads C + 2 floats (+ C / T written per
# ray), backward reads the same + d_fea and writes C + 1 -- against the torch ops of generator._torch_integrate
for (B, N, T, Cn) in [(16, 16384, 24, 32), (8, 65536, 24, 32)]:
    rs = torch.randn(B, N, T, Cn + 1, device=dev)
    zs = torch.sort(0.88 + 0.24 * torch.rand(B, N, T, device=dev), -1).values
    dfe = torch.randn(B, N, Cn, device=dev)
    fwd_bytes = 4 * B * N * (T * (Cn + 2) + Cn + T)
    bwd_bytes = 4 * B * N * (T * (Cn + 2) + Cn + T * (Cn + 1))
    ms = timeit(lambda: ops.integrate(rs, zs, None, "relu", True, False), reps=10)
    out.append(dict(op="integrate fwd", shape=[B, N, T, Cn + 1], ms=ms, gbs=fwd_bytes / 1e9 / ms * 1e3, frac=fwd_bytes / 1e9 / ms * 1e3 / peak))
    ms_t = timeit(lambda: cips3d_b200.generator._torch_integrate(rs, zs, None, "relu", True, False, Cn), reps=10)
    out.append(dict(op="integrate fwd: torch ops", shape=[B, N, T, Cn + 1], ms=ms_t, gbs=fwd_bytes / 1e9 / ms_t * 1e3, frac=fwd_bytes / 1e9 / ms_t * 1e3 / peak))
    rr = rs.clone().requires_grad_()
    fea, _ = ops.integrate(rr, zs, None, "relu", True, False)
    ms = timeit(lambda: torch.autograd.grad(fea, rr, dfe, retain_graph=True), reps=10)
    out.append(dict(op="integrate bwd", shape=[B, N, T, Cn + 1], ms=ms, gbs=bwd_bytes / 1e9 / ms * 1e3, frac=bwd_bytes / 1e9 / ms * 1e3 / peak))
    fea_t, _ = cips3d_b200.generator._torch_integrate(rr, zs, None, "relu", True, False, Cn)
    ms_t = timeit(lambda: torch.autograd.grad(fea_t, rr, dfe, retain_graph=True), reps=10)
    out.append(dict(op="integrate bwd: torch autograd", shape=[B, N, T, Cn + 1], ms=ms_t, gbs=bwd_bytes / 1e9 / ms_t * 1e3, frac=bwd_bytes / 1e9 / ms_t * 1e3 / peak))
    # merged form against torch's cat + sort + gather + integrate graph (what the hierarchical recipes run)
    S = T // 2
    fi, co = torch.randn(B, N, S, Cn + 1, device=dev).requires_grad_(), torch.randn(B, N, S, Cn + 1, device=dev).requires_grad_()
    zf, zc = 0.88 + 0.24 * torch.rand(B, N, S, device=dev), torch.sort(0.88 + 0.24 * torch.rand(B, N, S, device=dev), -1).values

    def torch_graph():
        all_z, ind = torch.sort(torch.cat([zf, zc], -1), dim=-1)
        all_out = torch.gather(torch.cat([fi, co], -2), -2, ind[..., None].expand(-1, -1, -1, Cn + 1))
        return cips3d_b200.generator._torch_integrate(all_out, all_z, None, "relu", True, False, Cn)[0]
    ms = timeit(lambda: ops.integrate_merged(fi, zf, co, zc, None, "relu", True, False), reps=10)
    out.append(dict(op="integrate_merged fwd", shape=[B, N, T, Cn + 1], ms=ms, gbs=fwd_bytes / 1e9 / ms * 1e3, frac=fwd_bytes / 1e9 / ms * 1e3 / peak))
    ms_t = timeit(torch_graph, reps=10)
    out.append(dict(op="integrate_merged fwd: torch cat + sort + gather + integrate", shape=[B, N, T, Cn + 1], ms=ms_t, gbs=fwd_bytes / 1e9 / ms_t * 1e3, frac=fwd_bytes / 1e9 / ms_t * 1e3 / peak))
    fm = ops.integrate_merged(fi, zf, co, zc, None, "relu", True, False)[0]
    ms = timeit(lambda: torch.autograd.grad(fm, (fi, co), dfe, retain_graph=True), reps=10)
    out.append(dict(op="integrate_merged bwd", shape=[B, N, T, Cn + 1], ms=ms, gbs=bwd_bytes / 1e9 / ms * 1e3, frac=bwd_bytes / 1e9 / ms * 1e3 / peak))
    ft = torch_graph()
    ms_t = timeit(lambda: torch.autograd.grad(ft, (fi, co), dfe, retain_graph=True), reps=10)
    out.append(dict(op="integrate_merged bwd: torch autograd", shape=[B, N, T, Cn + 1], ms=ms_t, gbs=bwd_bytes / 1e9 / ms_t * 1e3, frac=bwd_bytes / 1e9 / ms_t * 1e3 / peak))
    del rs, zs, dfe, rr, fea, fea_t, fi, co, zf, zc, fm, ft
    torch.cuda.empty_cache()
for r in out:
    print(json.dumps(r))
print(json.dumps(dict(hbm_peak_gbs=peak, note="algorithmic bytes (read x + write y [+ read ref]) / CUDA-event median; L2 flushed between reps")))
