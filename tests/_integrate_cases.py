"""Shared inputs / checker of the volume-integration op tests (emulation on the CPU, hardware on the GPU)."""
import torch

from oracle import cips3d_oracle as O

# (rays shape, samples, channels, clamp, last_back, white_back, noise)
CASES = [((2, 37), 24, 32, "relu", True, False, False),        # FFHQ recipe: 12 + 12 samples, last_back
         ((3, 11), 24, 32, "softplus", False, True, True),
         ((1, 65), 12, 32, "relu", False, False, True),         # coarse pass of get_fine_points (weights only matter)
         ((2, 9), 32, 32, "softplus", True, True, False),       # widest ray the kernel takes
         ((1, 5), 2, 32, "relu", True, False, False),           # two samples (the reference needs >= 2: deltas[..., :1])
         ((2, 19), 24, 3, "relu", True, False, False),          # pi-GAN rgb_dim 3 (one partial channel group)
         ((1, 13), 7, 40, "softplus", False, False, True),      # two channel groups, second partial
         ((1, 6), 5, 128, "relu", True, True, False)]           # four channel groups


def make(case, seed, device="cpu"):
    lead, T, Cn, clamp, lb, wb, noisy = case
    g = torch.Generator().manual_seed(seed)
    rs = torch.randn(*lead, T, Cn + 1, generator=g)
    opaque = torch.rand(*lead, 1, generator=g) < 0.3                      # a third of the rays saturate (alpha -> 1 early)
    rs[..., Cn] = (rs[..., Cn] + 0.3) * torch.where(opaque, 400.0, 8.0)
    z = torch.sort(0.88 + 0.24 * torch.rand(*lead, T, generator=g), -1).values
    noise = torch.randn(*lead, T, generator=g) * 0.5 if noisy else None
    d_fea = torch.randn(*lead, Cn, generator=g)
    mv = lambda t: None if t is None else t.to(device)                   # noqa: E731
    return mv(rs), mv(z), mv(noise), mv(d_fea)


def reference(case, rs, z, noise, d_fea, dtype=torch.float64):
    """fp64 (or same-dtype) oracle integrate + torch autograd."""
    _, T, Cn, clamp, lb, wb, _ = case
    r = rs.detach().to(dtype).requires_grad_()
    fea, _, w = O.integrate(r, z.to(dtype), None if noise is None else noise.to(dtype), clamp, lb, wb, dim_rgb=Cn)
    (dr,) = torch.autograd.grad(fea, r, d_fea.to(dtype), retain_graph=True)
    return fea.detach(), w.detach(), dr


def reference_depth_grad(case, rs, z, noise, dtype=torch.float64):
    """gradient of a depth loss sum(weights * z) -- it reaches rgb_sigma only through the returned weights"""
    _, T, Cn, clamp, lb, wb, _ = case
    r = rs.detach().to(dtype).requires_grad_()
    _, _, w = O.integrate(r, z.to(dtype), None if noise is None else noise.to(dtype), clamp, lb, wb, dim_rgb=Cn)
    (dr,) = torch.autograd.grad((w * z.to(dtype)).sum(), r)
    return dr


def check(case, pkg, rs, z, noise, d_fea):
    _, T, Cn, clamp, lb, wb, _ = case
    r = rs.clone().requires_grad_()
    assert pkg.ops.integrate_supported(r, z, noise)
    fea, w = pkg.ops.integrate(r, z, noise, clamp, lb, wb)
    assert fea.shape == rs.shape[:-2] + (Cn,) and w.shape == z.shape
    (dr,) = torch.autograd.grad(fea, r, d_fea, retain_graph=True)
    # a loss on the returned weights (depth = sum w z, as piGAN_lib's depth maps / inverse rendering would use): the reference's
    # fancy_integration propagates it; here a torch-op recompute inside IntegrateFunction.backward
    (dw,) = torch.autograd.grad((w * z).sum(), r)
    dw64 = reference_depth_grad(case, rs.cpu(), z.cpu(), None if noise is None else noise.cpu())
    assert (dw.cpu().double()[..., :Cn]).abs().max().item() == 0.0
    assert (dw.cpu().double()[..., Cn] - dw64[..., Cn]).abs().max().item() < 2e-4 * dw64[..., Cn].abs().max().item() + 1e-6
    f64, w64, d64 = reference(case, rs.cpu(), z.cpu(), None if noise is None else noise.cpu(), d_fea.cpu())
    fea, w, dr = fea.detach().cpu().double(), w.cpu().double(), dr.cpu().double()
    # weights are in [0, 1]; a product of up to 32 factors 1 - alpha_j, each carrying expf's rounding (<= 2 ulp on the GPU)
    assert (w - w64).abs().max().item() < 5e-6
    assert (fea - f64).abs().max().item() < 1e-5 * (1 + f64.abs().max().item())
    dc, dc64 = dr[..., :Cn], d64[..., :Cn]
    assert (dc - dc64).abs().max().item() < 1e-5 * (1 + dc64.abs().max().item())
    ds, ds64 = dr[..., Cn], d64[..., Cn]
    # d(sigma): fp32 alphas next to saturation carry ~1e-7 absolute error that the 1 / t_i of the cumprod backward scales up
    assert (ds - ds64).abs().max().item() < 2e-4 * ds64.abs().max().item() + 1e-6
    return fea, dr


# merged form: (rays shape, samples per half, channels, clamp, last_back, white_back, noise)
MERGED_CASES = [((2, 33), 12, 32, "relu", True, False, False),     # FFHQ recipe: 12 fine + 12 coarse
                ((2, 17), 12, 32, "softplus", False, True, True),
                ((1, 9), 16, 32, "relu", True, True, True),         # widest: 16 + 16
                ((1, 7), 1, 32, "relu", True, False, False),        # one sample per half
                ((2, 21), 12, 3, "relu", True, False, True),        # pi-GAN rgb_dim 3
                ((1, 5), 5, 70, "softplus", False, False, False)]   # three channel groups (kernel instantiation 4)


def make_merged(case, seed, device="cpu"):
    lead, S, Cn, clamp, lb, wb, noisy = case
    g = torch.Generator().manual_seed(seed)
    halves = []
    for _ in range(2):
        rs = torch.randn(*lead, S, Cn + 1, generator=g)
        opaque = torch.rand(*lead, 1, generator=g) < 0.3
        rs[..., Cn] = (rs[..., Cn] + 0.3) * torch.where(opaque, 400.0, 8.0)
        halves.append(rs)
    z_coarse = torch.sort(0.88 + 0.24 * torch.rand(*lead, S, generator=g), -1).values
    z_fine = 0.88 + 0.24 * torch.rand(*lead, S, generator=g)               # fine depths arrive unsorted relative to the coarse ones
    z_fine[..., 0] = z_coarse[..., S // 2]                                  # a tie across the halves: cat order (fine first) wins
    noise = torch.randn(*lead, 2 * S, generator=g) * 0.5 if noisy else None
    d_fea = torch.randn(*lead, Cn, generator=g)
    mv = lambda t: None if t is None else t.to(device)                     # noqa: E731
    return mv(halves[0]), mv(z_fine), mv(halves[1]), mv(z_coarse), mv(noise), mv(d_fea)


def reference_merged(case, fine, z_fine, coarse, z_coarse, noise, d_fea, dtype=torch.float64):
    """generator.py:1733-1752 in fp64: cat, sort, gather, fancy_integration; autograd to both halves."""
    _, S, Cn, clamp, lb, wb, _ = case
    f, c = fine.detach().to(dtype).requires_grad_(), coarse.detach().to(dtype).requires_grad_()
    all_out = torch.cat([f, c], -2)
    all_z, ind = torch.sort(torch.cat([z_fine, z_coarse], -1), dim=-1, stable=True)  # fp32 depths: same order as the kernel sees
    all_out = torch.gather(all_out, -2, ind[..., None].expand(*([-1] * (ind.dim())), Cn + 1))
    fea, _, w = O.integrate(all_out, all_z.to(dtype), None if noise is None else noise.to(dtype), clamp, lb, wb, dim_rgb=Cn)
    df, dc = torch.autograd.grad(fea, (f, c), d_fea.to(dtype))
    return fea.detach(), w.detach(), all_z, df, dc


def check_merged(case, pkg, fine, z_fine, coarse, z_coarse, noise, d_fea):
    _, S, Cn, clamp, lb, wb, _ = case
    f, c = fine.clone().requires_grad_(), coarse.clone().requires_grad_()
    assert pkg.ops.integrate_merged_supported(f, z_fine, c, z_coarse, noise)
    fea, w, zs = pkg.ops.integrate_merged(f, z_fine, c, z_coarse, noise, clamp, lb, wb)
    assert not w.requires_grad and not zs.requires_grad and w.shape == zs.shape == z_fine.shape[:-1] + (2 * S,)
    df, dc = torch.autograd.grad(fea, (f, c), d_fea)
    cpu = lambda t: None if t is None else t.cpu()                         # noqa: E731
    f64, w64, z64, df64, dc64 = reference_merged(case, *(cpu(t) for t in (fine, z_fine, coarse, z_coarse, noise, d_fea)))
    assert torch.equal(zs.cpu(), z64)                                       # the sort itself is exact
    assert (w.cpu().double() - w64).abs().max().item() < 5e-6
    assert (fea.detach().cpu().double() - f64).abs().max().item() < 1e-5 * (1 + f64.abs().max().item())
    for got, want in ((df, df64), (dc, dc64)):
        got = got.cpu().double()
        assert (got[..., :Cn] - want[..., :Cn]).abs().max().item() < 1e-5 * (1 + want[..., :Cn].abs().max().item())
    smax = max(df64[..., Cn].abs().max().item(), dc64[..., Cn].abs().max().item())
    for got, want in ((df, df64), (dc, dc64)):
        assert (got.cpu().double()[..., Cn] - want[..., Cn]).abs().max().item() < 2e-4 * smax + 1e-6


# ------------------------------------------------------------------ goldens of the REAL pigan_utils.fancy_integration
# (tests/golden/fancy_integration.npz, written by tools/make_golden_integrate.py from the unmodified reference)
GOLDEN_CASES = ("relu_lastback", "softplus_white_noise", "coarse_relu_noise", "pigan_rgb3_backs", "merged_relu_lastback",
                "merged_softplus_noise")


def load_golden(name):
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fancy_integration.npz"))
    d = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(name + "/")}
    cfg = [int(v) for v in d.pop("cfg")]
    d = {k: torch.from_numpy(v) if getattr(v, "ndim", 0) else float(v) for k, v in d.items()}
    return cfg, d


def check_golden(name, pkg, device="cpu"):
    """The native op on the stored inputs and the stored noise draw vs what the real function returned (values and gradient)."""
    (B, N, T, Cn, softplus, lb, wb, merged), d = load_golden(name)
    clamp = "softplus" if softplus else "relu"
    mv = lambda t: t.to(device)                                            # noqa: E731
    noise = mv(d["noise"][..., 0]) if d["noise_std"] != 0 else None
    leaf = mv(d["rgb_sigma"]).requires_grad_()
    if merged:
        S = T // 2
        fea, w, zs = pkg.ops.integrate_merged(leaf[:, :, :S], mv(d["z_fine"][..., 0]), leaf[:, :, S:], mv(d["z_coarse"][..., 0]),
                                              noise, clamp, bool(lb), bool(wb))
        assert torch.equal(zs.cpu(), d["z"][..., 0])                       # the reference's sorted depths, bit for bit
    else:
        zs = mv(d["z"][..., 0])
        fea, w = pkg.ops.integrate(leaf, zs, noise, clamp, bool(lb), bool(wb))
    (grad,) = torch.autograd.grad(fea, leaf, mv(d["d_rgb"]))
    depth = torch.sum(w * zs, -1, keepdim=True)
    assert (w.cpu() - d["weights"][..., 0]).abs().max().item() < 5e-6
    assert (fea.detach().cpu() - d["rgb"]).abs().max().item() < 1e-5 * (1 + d["rgb"].abs().max().item())
    assert (depth.cpu() - d["depth"]).abs().max().item() < 1e-5
    g_ref = d["grad"]
    assert (grad.cpu()[..., :Cn] - g_ref[..., :Cn]).abs().max().item() < 1e-5 * (1 + g_ref[..., :Cn].abs().max().item())
    # fp32 on both sides: the reference's own d(sigma) carries the rounding of 1 / t_i next to saturated samples
    assert (grad.cpu()[..., Cn] - g_ref[..., Cn]).abs().max().item() < 1e-3 * g_ref[..., Cn].abs().max().item() + 1e-6


PDF_GOLDEN_CASES = ("pdf_s12", "pdf_n32_k40", "pdf_det", "pdf_peaky")


def check_pdf_golden(name, pkg, device="cpu"):
    """ops.sample_pdf_from_u on the stored bins / weights / uniforms vs what the real pigan_utils.sample_pdf returned."""
    (rays, n, k, det), d = load_golden(name)
    got = pkg.ops.sample_pdf_from_u(d["bins"].to(device), d["weights"].to(device), d["u"].to(device))
    assert got.shape == (rays, k)
    # an index that flips on a 1-ulp difference of the cdf moves the sample continuously; the bins span 0.24
    assert (got.cpu() - d["samples"]).abs().max().item() < 2e-6


# ------------------------------------------------------------------ goldens of the REAL explicit-points path
# (tests/golden/points_forward.npz, tools/make_golden_points.py: get_world_points_and_direction -> points_forward)
POINTS_GOLDEN_CASES = ("hier_noise_lastback", "flat_softplus_white")


def check_points_golden(name, pkg, G, device, backend):
    """Replay the reference's draws through this package's get_world_points_and_direction + points_forward; compare with what
    the unmodified functions returned."""
    import os
    import numpy as np
    from _util import replay_draws
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "points_forward.npz"))
    d = {k.split("/", 1)[1]: torch.from_numpy(np.asarray(g[k])) for k in g.files if k.startswith(name + "/")}
    R, hier, softplus, wb, lb, subset, n_world = (int(v) for v in d["cfg"])
    draws = [d[k] for k in sorted(k for k in d if k.startswith("draw"))]
    zs = {k: d[k].to(device) for k in ("z_nerf", "z_inr")}
    G.train_integrate = backend
    with torch.no_grad(), replay_draws(draws, device):
        world = pkg.comm_utils.get_world_points_and_direction(
            batch_size=2, num_steps=12, img_size=R, fov=12, ray_start=0.88, ray_end=1.12, h_stddev=0.3, v_stddev=0.155,
            h_mean=np.pi * 0.5, v_mean=np.pi * 0.5, sample_dist="gaussian", lock_view_dependence=False, device=device)
        for got, key in zip(world, ("points", "dirs_exp", "origins", "dirs", "z_vals", "pitch", "yaw")):
            assert (got.cpu() - d["world_" + key]).abs().max().item() < 2e-6, key
        pts, dirs_exp, origins, dirs, z_vals, _, _ = world
        inr, aux = G.points_forward(
            style_dict=G.mapping_network(**zs), transformed_points=pts.view(2, R * R, 12, 3),
            transformed_ray_directions_expanded=dirs_exp.view(2, R * R, 12, 3), num_steps=12, hierarchical_sample=bool(hier),
            z_vals=z_vals, clamp_mode="softplus" if softplus else "relu", nerf_noise=float(d["nerf_noise"]),
            transformed_ray_origins=origins, transformed_ray_directions=dirs, white_back=bool(wb), last_back=bool(lb),
            return_aux_img=True, idx_grad=torch.arange(0, R * R, 2, device=device) if subset else None)
    # rendered RGB within 1e-3 relative (north_star); the fp32 field through torch ops is far tighter than that
    assert (inr.cpu() - d["inr"]).abs().max().item() < 1e-3 * d["inr"].abs().max().item()
    assert (aux.cpu() - d["aux"]).abs().max().item() < 1e-3 * d["aux"].abs().max().item()
    return (inr.cpu() - d["inr"]).abs().max().item()
