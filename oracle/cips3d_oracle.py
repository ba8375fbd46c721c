"""CPU oracle for the CIPS-3D hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement (torch CPU tensors, fp32 or fp64) of the reference's generator
forward (rays -> FiLM-SIREN -> hierarchical resampling -> volume integration -> per-pixel
CIPS MLP) and discriminator forward.  Every function cites the reference file:line it
follows (paths relative to the reference checkout).  It works from a plain ``state_dict``
with the reference's key names and takes every random draw as an explicit argument, so the
same inputs can be replayed through the reference (tools/make_golden.py), this oracle and
the CUDA path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl
reference legs may import this module.  The product package (cips-3d_b200/) never does;
it fails loudly when the CUDA library is missing.

Parity pin: tests/test_oracle_vs_reference.py runs the real reference (imported through
tools/ref_shim.py) against this file whenever /root/reference exists, and
tests/golden/*.npz hold outputs of the real reference that this file must reproduce
everywhere (tests/test_oracle_golden.py).  The reference itself ships no golden vectors
for this path.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# configuration of the shipping model: exp/cips3d/configs/ffhq_exp.yaml:43-81, 86-96, 117-126
# --------------------------------------------------------------------------------------
G_CFG = dict(
    z_dim=256,
    nerf_cfg=dict(in_dim=3, hidden_dim=128, hidden_layers=2, rgb_dim=32, style_dim=128),
    mapping_nerf_cfg=dict(z_dim=256, hidden_dim=128, base_layers=4, head_layers=0),
    inr_cfg=dict(input_dim=32, style_dim=512, hidden_dim=512, pre_rgb_dim=3),
    mapping_inr_cfg=dict(z_dim=512, hidden_dim=512, base_layers=8, head_layers=0,
                         add_norm=True, norm_out=True),
)
G_KWARGS = dict(fov=12, ray_start=0.88, ray_end=1.12, num_steps=12, h_stddev=0.3,
                v_stddev=0.155, hierarchical_sample=True, psi=1., sample_dist="gaussian")
CIPS_BLOCK_NAMES = ("4", "8", "16", "32", "64", "128", "256", "512", "1024")  # generator.py:1041-1051


# --------------------------------------------------------------------------------------
# R2  get_initial_rays_trig            exp/comm/comm_utils.py:365-412
# --------------------------------------------------------------------------------------
def initial_rays(img_size, fov, ray_start, ray_end, num_steps, dtype=torch.float32):
    """-> dirs_cam (HW,3) unit, z_vals (S,).  Ray index = h*W + w (comm_utils.py:392-395)."""
    W = H = img_size
    xs = torch.linspace(-1, 1, W, dtype=dtype)
    ys = torch.linspace(1, -1, H, dtype=dtype)
    x = xs[None, :].expand(H, W).reshape(-1)          # x varies fastest (w)
    y = ys[:, None].expand(H, W).reshape(-1)          # y varies with h
    z = -torch.ones_like(x) / np.tan((2 * math.pi * fov / 360) / 2)
    d = torch.stack([x, y, z], -1)
    d = d / torch.norm(d, dim=-1, keepdim=True)        # normalize_vecs, comm_utils.py:353-362
    z_vals = torch.linspace(ray_start, ray_end, num_steps, dtype=dtype)
    return d, z_vals


# --------------------------------------------------------------------------------------
# R4  sample_camera_positions ('gaussian'/'normal' branch + clamp)   comm_utils.py:451-535
# --------------------------------------------------------------------------------------
def camera_origin(yaw_n, pitch_n, h_stddev, v_stddev, h_mean=math.pi * 0.5, v_mean=math.pi * 0.5, r=1):
    """yaw_n, pitch_n: the N(0,1) draws (B,1) in the order the reference draws them."""
    theta = yaw_n * h_stddev + h_mean
    phi = pitch_n * v_stddev + v_mean
    phi = torch.clamp(phi, 1e-5, math.pi - 1e-5)
    o = torch.zeros((theta.shape[0], 3), dtype=theta.dtype)
    o[:, 0:1] = r * torch.sin(phi) * torch.cos(theta)
    o[:, 2:3] = r * torch.sin(phi) * torch.sin(theta)
    o[:, 1:2] = r * torch.cos(phi)
    return o, phi, theta


# --------------------------------------------------------------------------------------
# R5  create_cam2world_matrix          comm_utils.py:538-581
# --------------------------------------------------------------------------------------
def _nrm(v):
    return v / torch.norm(v, dim=-1, keepdim=True)


def cam2world(forward, origin, up=None):
    forward = _nrm(forward)
    if up is None:
        up = torch.tensor([0, 1, 0], dtype=forward.dtype).expand_as(forward)
    left = _nrm(torch.cross(up, forward, dim=-1))
    up = _nrm(torch.cross(forward, left, dim=-1))
    B = forward.shape[0]
    rot = torch.eye(4, dtype=forward.dtype).unsqueeze(0).repeat(B, 1, 1)
    rot[:, :3, :3] = torch.stack((-left, up, -forward), dim=-1)
    tr = torch.eye(4, dtype=forward.dtype).unsqueeze(0).repeat(B, 1, 1)
    tr[:, :3, 3] = origin
    return tr @ rot


# --------------------------------------------------------------------------------------
# R3+R6  perturb_points + transform_sampled_points      comm_utils.py:416-438, 584-679
# --------------------------------------------------------------------------------------
def world_points(dirs_cam, z_vals, jitter_u, c2w):
    """dirs_cam (HW,3), z_vals (S,), jitter_u (B,HW,S) U[0,1), c2w (B,4,4)
    -> points (B,HW,S,3), z (B,HW,S), dirs_w (B,HW,3), origins (B,3)"""
    B = c2w.shape[0]
    dist = z_vals[1] - z_vals[0]
    off = (jitter_u - 0.5) * dist                                   # comm_utils.py:431-433
    z = z_vals[None, None, :] + off                                 # (B,HW,S)
    p_cam = dirs_cam[None, :, None, :] * z_vals[None, None, :, None]  # points before jitter
    p_cam = p_cam + off[..., None] * dirs_cam[None, :, None, :]     # comm_utils.py:436-437
    Rm = c2w[:, :3, :3]
    t = c2w[:, :3, 3]
    pts = torch.einsum("bij,bnsj->bnsi", Rm, p_cam) + t[:, None, None, :]
    dirs_w = torch.einsum("bij,nj->bni", Rm, dirs_cam)
    return pts, z, dirs_w, t


# --------------------------------------------------------------------------------------
# R1  MultiHeadMappingNetwork          exp/cips3d/models/multi_head_mapping.py:13-19,130-153
# --------------------------------------------------------------------------------------
def mapping_network(sd, prefix, z, base_layers, add_norm=False, norm_out=False, head_layers=0, **_):
    assert head_layers == 0, "shipping configs use Identity heads (ffhq_exp.yaml:61,75)"
    x = z * torch.rsqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)     # PixelNorm
    idx = 0
    for i in range(base_layers):
        x = F.linear(x, sd[f"{prefix}.base_net.{idx}.weight"], sd[f"{prefix}.base_net.{idx}.bias"])
        idx += 1
        if i != base_layers - 1:
            if add_norm:
                x = F.layer_norm(x, x.shape[-1:], sd[f"{prefix}.base_net.{idx}.weight"],
                                 sd[f"{prefix}.base_net.{idx}.bias"])
                idx += 1
            x = F.leaky_relu(x, 0.2)
            idx += 1
    if norm_out:
        x = F.layer_norm(x, x.shape[-1:], sd[f"{prefix}.base_net.{idx}.weight"],
                         sd[f"{prefix}.base_net.{idx}.bias"])
    return x


# --------------------------------------------------------------------------------------
# R7-R9  UniformBoxWarp + FiLMLayer + NeRFNetwork     nerf_network.py:39-45,
#        film_layer.py:78-107, exp/cips3d/models/generator.py:260-317
# --------------------------------------------------------------------------------------
def film_params(sd, prefix, w):
    """gamma = 15*gain_fc(w)+30, beta = bias_fc(w)   (film_layer.py:59,89-91)"""
    g = F.linear(w, sd[f"{prefix}.gain_fc.weight"], sd[f"{prefix}.gain_fc.bias"]) * 15 + 30
    b = F.linear(w, sd[f"{prefix}.bias_fc.weight"], sd[f"{prefix}.bias_fc.bias"])
    return g, b


def film_layer(sd, prefix, x, w):
    g, b = film_params(sd, prefix, w)
    x = F.linear(x, sd[f"{prefix}.linear.weight"], sd[f"{prefix}.linear.bias"])
    return torch.sin(g[:, None, :] * x + b[:, None, :])


def nerf_network(sd, points, w_nerf, prefix="siren", hidden_layers=2):
    """points (B,P,3), w_nerf (B,128) -> (B,P,33) = [feature(32), sigma(1)]"""
    x = points * (2 / 0.24)                                            # gridwarper
    for i in range(hidden_layers):
        x = film_layer(sd, f"{prefix}.network.{i}", x, w_nerf)
    sigma = F.linear(x, sd[f"{prefix}.final_layer.weight"], sd[f"{prefix}.final_layer.bias"])
    x = film_layer(sd, f"{prefix}.color_layer_sine", x, w_nerf)
    rgb = F.linear(x, sd[f"{prefix}.color_layer_linear.0.weight"], sd[f"{prefix}.color_layer_linear.0.bias"])
    return torch.cat([rgb, sigma], dim=-1)


# --------------------------------------------------------------------------------------
# R10  fancy_integration               exp/pigan/pigan_utils.py:212-273
# --------------------------------------------------------------------------------------
def integrate(rgb_sigma, z, noise, clamp_mode="relu", last_back=False, white_back=False, dim_rgb=32):
    """rgb_sigma (B,N,n,33), z (B,N,n), noise (B,N,n) already scaled by noise_std (or None)
    -> rgb (B,N,32), depth (B,N), weights (B,N,n)"""
    rgbs = rgb_sigma[..., :dim_rgb]
    sig = rgb_sigma[..., dim_rgb]
    deltas = z[..., 1:] - z[..., :-1]
    deltas = torch.cat([deltas, 1e10 * torch.ones_like(deltas[..., :1])], -1)
    if noise is not None:
        sig = sig + noise
    if clamp_mode == "softplus":
        alphas = 1 - torch.exp(-deltas * F.softplus(sig))
    elif clamp_mode == "relu":
        alphas = 1 - torch.exp(-deltas * F.relu(sig))
    else:
        raise AssertionError("Need to choose clamp mode")
    shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-10], -1)
    weights = alphas * torch.cumprod(shifted, -1)[..., :-1]
    wsum = weights.sum(-1)
    if last_back:
        weights = weights.clone()
        weights[..., -1] += 1 - wsum
    rgb = torch.sum(weights[..., None] * rgbs, -2)
    depth = torch.sum(weights * z, -1)
    if white_back:
        rgb = rgb + 1 - wsum[..., None]
    return rgb, depth, weights


# --------------------------------------------------------------------------------------
# R11  sample_pdf                      exp/pigan/pigan_utils.py:164-209
# --------------------------------------------------------------------------------------
def sample_pdf(bins, weights, u, eps=1e-5):
    """bins (R,S-1), weights (R,S-2), u (R,Nimp) -> samples (R,Nimp)"""
    n_s = weights.shape[1]
    weights = weights + eps
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u)
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, n_s)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_b, bin_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    return bin_b + (u - cdf_b) / denom * (bin_a - bin_b)


def fine_z_vals(coarse, z, noise_c, pdf_u, clamp_mode="relu", dim_rgb=32):
    """get_fine_points_and_direction  exp/dev/nerf_inr/models/generator_nerf_inr.py:537-598
    coarse (B,N,S,33), z (B,N,S), noise_c (B,N,S)|None, pdf_u (B*N,S) -> fine z (B,N,S)"""
    B, N, S = z.shape
    _, _, w = integrate(coarse, z, noise_c, clamp_mode=clamp_mode, dim_rgb=dim_rgb)
    w = w.reshape(B * N, S) + 1e-5
    zz = z.reshape(B * N, S)
    mid = 0.5 * (zz[:, :-1] + zz[:, 1:])
    return sample_pdf(mid, w[:, 1:-1], pdf_u).reshape(B, N, S)


# --------------------------------------------------------------------------------------
# R13-R16  SinStyleMod / SinBlock / ToRGB / CIPSNet / aux_to_rbg
#   exp/comm/models/mod_conv_fc.py:452-496, exp/cips3d/models/generator.py:949-974,
#   983-1006, 1107-1154, 1204-1208
# --------------------------------------------------------------------------------------
def sin_style_mod(sd, prefix, x, w, eps=1e-8):
    s = F.linear(w, sd[f"{prefix}.modulation.weight"], sd[f"{prefix}.modulation.bias"])
    weight = sd[f"{prefix}.weight"] * (s[:, :, None] + 1)                # (B,in,out)
    demod = torch.rsqrt(weight.pow(2).sum(1) + eps)                     # (B,out)
    weight = weight * demod[:, None, :]
    return torch.bmm(x, weight)


def cips_net(sd, x, w_inr, prefix="inr_net", img_size=1024, return_hidden=False):
    """x (B,N,32), w_inr (B,512) -> rgb (B,N,3) after tanh. All 9 blocks run unless
    img_size names an earlier block (points_forward never passes it: generator.py:1754)."""
    stop = str(2 ** int(np.log2(img_size)))
    rgb = 0
    for idx, name in enumerate(CIPS_BLOCK_NAMES):
        x0 = x
        x = F.leaky_relu(sin_style_mod(sd, f"{prefix}.network.{name}.mod1", x, w_inr), 0.2)
        x = F.leaky_relu(sin_style_mod(sd, f"{prefix}.network.{name}.mod2", x, w_inr), 0.2)
        if idx >= 4 and x.shape[-1] == x0.shape[-1]:
            x = x + x0
        if idx >= 3:
            rgb = F.linear(x, sd[f"{prefix}.to_rgbs.{name}.linear.weight"],
                           sd[f"{prefix}.to_rgbs.{name}.linear.bias"]) + rgb
        if name == stop:
            break
    out = torch.tanh(rgb)          # pre_rgb_dim == 3 -> tanh only (generator.py:1089-1095)
    return (out, x) if return_hidden else out


def aux_to_rgb(sd, fea):
    return torch.tanh(F.linear(fea, sd["aux_to_rbg.0.weight"], sd["aux_to_rbg.0.bias"]))


# --------------------------------------------------------------------------------------
# R17  points_forward / whole_grad_forward / forward    generator.py:1256-1370,1378-1534,1659-1762
# --------------------------------------------------------------------------------------
def render_features(sd, w_nerf, c2w, jitter_u, pdf_u, noise_c=None, noise_f=None, *, img_size,
                    fov, ray_start, ray_end, num_steps, hierarchical_sample=True,
                    clamp_mode="relu", white_back=False, last_back=False, ray_idx=None):
    """The volumetric renderer (rows R2-R12).  Random draws:
         jitter_u (B,HW,S)  U[0,1)     comm_utils.py:432
         noise_c  (B,N,S)   N(0,1)*noise_std or None      pigan_utils.py:246 (coarse pass)
         pdf_u    (B*N,S)   U[0,1)     pigan_utils.py:192
         noise_f  (B,N,2S)  N(0,1)*noise_std or None      pigan_utils.py:246 (final pass)
       ray_idx: optional LongTensor selecting a ray subset (gather_points, comm_utils.py:264-282).
       -> dict(pixels_fea (B,N,32), depth, weights, coarse (B,N,S,33), fine (B,N,S,33), fine_z, z, all_z)"""
    dt = c2w.dtype
    dirs_cam, z_vals = initial_rays(img_size, fov, ray_start, ray_end, num_steps, dt)
    pts, z, dirs_w, origins = world_points(dirs_cam, z_vals, jitter_u, c2w)
    if ray_idx is not None:
        pts, z, dirs_w = pts[:, ray_idx], z[:, ray_idx], dirs_w[:, ray_idx]
    B, N, S = z.shape
    coarse = nerf_network(sd, pts.reshape(B, N * S, 3), w_nerf).reshape(B, N, S, -1)
    out = dict(coarse=coarse, z=z)
    if hierarchical_sample:
        fz = fine_z_vals(coarse, z, noise_c, pdf_u, clamp_mode)
        fpts = origins[:, None, None, :] + dirs_w[:, :, None, :] * fz[..., None]
        fine = nerf_network(sd, fpts.reshape(B, N * S, 3), w_nerf).reshape(B, N, S, -1)
        all_out = torch.cat([fine, coarse], dim=-2)
        all_z = torch.cat([fz, z], dim=-1)
        all_z, ind = torch.sort(all_z, dim=-1)                 # generator.py:1735-1738
        all_out = torch.gather(all_out, -2, ind[..., None].expand(-1, -1, -1, all_out.shape[-1]))
        out.update(fine=fine, fine_z=fz)
    else:
        all_out, all_z = coarse, z
    fea, depth, wts = integrate(all_out, all_z, noise_f, clamp_mode, last_back, white_back)
    out.update(pixels_fea=fea, depth=depth, weights=wts, all_z=all_z)
    return out


def generator_forward(sd, zs, draws, *, img_size, fov, ray_start, ray_end, num_steps,
                      h_stddev, v_stddev, hierarchical_sample=True, h_mean=math.pi * 0.5,
                      v_mean=math.pi * 0.5, clamp_mode="relu", nerf_noise=0., white_back=False,
                      last_back=False, return_aux_img=False, cfg=G_CFG, return_all=False, **_):
    """GeneratorNerfINR.forward (whole_grad_forward branch, psi == 1).
    draws: dict with jitter_u (B,HW,S), yaw_n (B,1), pitch_n (B,1), noise_c (B,HW,S),
    pdf_u (B*HW,S), noise_f (B,HW,2S) -- the raw rand/randn tensors in reference order."""
    w_nerf = mapping_network(sd, "mapping_network_nerf", zs["z_nerf"], **cfg["mapping_nerf_cfg"])
    w_inr = mapping_network(sd, "mapping_network_inr", zs["z_inr"], **cfg["mapping_inr_cfg"])
    origin, pitch, yaw = camera_origin(draws["yaw_n"], draws["pitch_n"], h_stddev, v_stddev, h_mean, v_mean)
    c2w = cam2world(-origin, origin)
    nc = draws["noise_c"] * nerf_noise if nerf_noise else None
    nf = draws["noise_f"] * nerf_noise if nerf_noise else None
    r = render_features(sd, w_nerf, c2w, draws["jitter_u"], draws["pdf_u"], nc, nf,
                        img_size=img_size, fov=fov, ray_start=ray_start, ray_end=ray_end,
                        num_steps=num_steps, hierarchical_sample=hierarchical_sample,
                        clamp_mode=clamp_mode, white_back=white_back, last_back=last_back)
    rgb = cips_net(sd, r["pixels_fea"], w_inr)
    B = rgb.shape[0]
    img = rgb.reshape(B, img_size, img_size, 3).permute(0, 3, 1, 2)
    pitch_yaw = torch.cat([pitch, yaw], -1)
    if return_aux_img:
        aux = aux_to_rgb(sd, r["pixels_fea"]).reshape(B, img_size, img_size, 3).permute(0, 3, 1, 2)
        img = torch.cat([img, aux])
        pitch_yaw = torch.cat([pitch_yaw, pitch_yaw])
    if return_all:
        r.update(w_nerf=w_nerf, w_inr=w_inr, c2w=c2w)
        return img, pitch_yaw, r
    return img, pitch_yaw


def draw_randoms(B, img_size, num_steps, generator=None, dtype=torch.float32):
    """Draw the per-forward random tensors in the reference's call order."""
    HW, S = img_size * img_size, num_steps
    g = generator
    d = OrderedDict()
    d["jitter_u"] = torch.rand((B, HW, S, 1), generator=g, dtype=dtype)[..., 0]
    d["yaw_n"] = torch.randn((B, 1), generator=g, dtype=dtype)
    d["pitch_n"] = torch.randn((B, 1), generator=g, dtype=dtype)
    d["noise_c"] = torch.randn((B, HW, S, 1), generator=g, dtype=dtype)[..., 0]
    d["pdf_u"] = torch.rand((B * HW, S), generator=g, dtype=dtype)
    d["noise_f"] = torch.randn((B, HW, 2 * S, 1), generator=g, dtype=dtype)[..., 0]
    return d


# --------------------------------------------------------------------------------------
# D3  fused bias + leaky-ReLU          exp/comm/op/fused_bias_act_kernel.cu:19-50
# --------------------------------------------------------------------------------------
def bias_act(x, bias=None, ref=None, act=3, grad=0, alpha=0.2, scale=2 ** 0.5):
    """y = f(x + b[(i / step_b) % size_b]) * scale, literal restatement of the switch at
    fused_bias_act_kernel.cu:34-46.  x any shape (N,C,...)."""
    x = x.contiguous()
    if bias is not None and bias.numel():
        shape = [1, -1] + [1] * (x.dim() - 2)
        x = x + bias.view(*shape)
    if act == 1:
        y = x if grad < 2 else torch.zeros_like(x)
    elif act == 3:
        if grad == 0:
            y = torch.where(x > 0, x, x * alpha)
        elif grad == 1:
            y = torch.where(ref > 0, x, x * alpha)
        else:
            y = torch.zeros_like(x)
    else:
        raise ValueError(act)
    return y * scale


# --------------------------------------------------------------------------------------
# D2  upfirdn2d                        exp/comm/op/upfirdn2d_kernel.cu:52-139 (generic form L17-50)
# --------------------------------------------------------------------------------------
def upfirdn2d(x, kernel, up=(1, 1), down=(1, 1), pad=(0, 0, 0, 0)):
    """x (B,C,H,W); kernel (kh,kw); pad = (x0,x1,y0,y1).  Zero-upsample, pad/crop,
    correlate with the flipped kernel, decimate."""
    up_x, up_y = up
    down_x, down_y = down
    px0, px1, py0, py1 = pad
    B, C, H, W = x.shape
    kh, kw = kernel.shape
    y = x.reshape(B * C, 1, H, 1, W, 1)
    y = F.pad(y, [0, up_x - 1, 0, 0, 0, up_y - 1])
    y = y.reshape(B * C, 1, H * up_y, W * up_x)
    y = F.pad(y, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    y = y[:, :, max(-py0, 0): y.shape[2] - max(-py1, 0), max(-px0, 0): y.shape[3] - max(-px1, 0)]
    w = torch.flip(kernel, [0, 1]).view(1, 1, kh, kw).to(y.dtype)
    y = F.conv2d(y, w)
    y = y[:, :, ::down_y, ::down_x]
    return y.reshape(B, C, y.shape[2], y.shape[3])


def upfirdn2d_loops(x, kernel, up=(1, 1), down=(1, 1), pad=(0, 0, 0, 0)):
    """Pure-python/numpy loop form of the same op (small cases only) -- an independent
    check of the conv2d formulation above, following upfirdn2d_kernel.cu:17-50."""
    x = np.asarray(x, dtype=np.float64)
    k = np.asarray(kernel, dtype=np.float64)
    up_x, up_y = up
    down_x, down_y = down
    px0, px1, py0, py1 = pad
    B, C, H, W = x.shape
    kh, kw = k.shape
    out_h = (H * up_y + py0 + py1 - kh) // down_y + 1
    out_w = (W * up_x + px0 + px1 - kw) // down_x + 1
    out = np.zeros((B, C, out_h, out_w))
    for oy in range(out_h):
        for ox in range(out_w):
            acc = np.zeros((B, C))
            for ky in range(kh):
                for kx in range(kw):
                    iy = oy * down_y + ky - py0
                    ix = ox * down_x + kx - px0
                    if iy < 0 or ix < 0 or iy % up_y or ix % up_x:
                        continue
                    iy //= up_y
                    ix //= up_x
                    if iy >= H or ix >= W:
                        continue
                    acc += x[:, :, iy, ix] * k[kh - 1 - ky, kw - 1 - kx]
            out[:, :, oy, ox] = acc
    return out


# --------------------------------------------------------------------------------------
# D1,D4,D5  EqualConv2d / ConvLayer / ResBlock / EqualLinear / Discriminator_MultiScale(_Aux)
#   exp/cips3d/models/discriminator.py:20-54, 134-222, 224-252, 254-288, 405-585, 588-664
# --------------------------------------------------------------------------------------
def _blur_kernel(dtype):
    k = torch.tensor([1., 3., 3., 1.], dtype=dtype)
    k = k[None, :] * k[:, None]
    return k / k.sum()                                                      # make_kernel L57-65


def conv_layer(sd, prefix, x, kernel_size, downsample=False, activate=True, bias=True):
    """ConvLayer L134-222 (no upsample / reflect paths: unused by the shipping D)."""
    w = sd[f"{prefix}.equal_conv.weight"]
    cin = w.shape[1]
    scale = 1 / math.sqrt(cin * kernel_size ** 2)
    stride, padding = 1, 0
    if downsample:
        p = (4 - 2) + (kernel_size - 1)
        pad0, pad1 = (p + 1) // 2, p // 2
        x = upfirdn2d(x, _blur_kernel(x.dtype), pad=(pad0, pad1, pad0, pad1))
        stride = 2
    else:
        padding = (kernel_size - 1) // 2
    b = sd.get(f"{prefix}.equal_conv.bias") if (bias and not activate) else None
    x = F.conv2d(x, w * scale, bias=b, stride=stride, padding=padding)
    if activate:
        if bias:
            x = bias_act(x, sd[f"{prefix}.flrelu.bias"])
        else:
            x = F.leaky_relu(x, 0.2) * math.sqrt(2)
    return x


def res_block(sd, prefix, x, first_downsample=False):
    if first_downsample:
        out = conv_layer(sd, f"{prefix}.conv1", x, 3, downsample=True)
        out = conv_layer(sd, f"{prefix}.conv2", out, 3)
    else:
        out = conv_layer(sd, f"{prefix}.conv1", x, 3)
        out = conv_layer(sd, f"{prefix}.conv2", out, 3, downsample=True)
    skip = conv_layer(sd, f"{prefix}.skip", x, 1, downsample=True, activate=False, bias=False)
    return (out + skip) / math.sqrt(2)


def equal_linear(sd, prefix, x, activation=False, lr_mul=1):
    w = sd[f"{prefix}.weight"]
    scale = (1 / math.sqrt(w.shape[1])) * lr_mul
    if activation:
        return bias_act(F.linear(x, w * scale), sd[f"{prefix}.bias"] * lr_mul)
    return F.linear(x, w * scale, bias=sd[f"{prefix}.bias"] * lr_mul)


def discriminator_multiscale(sd, prefix, x, alpha=1., first_downsample=False):
    """Discriminator_MultiScale.forward L502-585, stddev_group == 0, diffaug False."""
    size = x.shape[-1]
    log_size = int(math.log(size, 2))
    cur = conv_layer(sd, f"{prefix}.conv_in.{2 ** log_size}", x, 1)
    cur = res_block(sd, f"{prefix}.convs.{2 ** log_size}", cur, first_downsample)
    if alpha < 1:
        down = F.interpolate(x, scale_factor=0.5, mode="bilinear")
        down = conv_layer(sd, f"{prefix}.conv_in.{2 ** (log_size - 1)}", down, 1)
        out = alpha * cur + (1 - alpha) * down
    else:
        out = cur
    for i in range(log_size - 1, 2, -1):
        out = res_block(sd, f"{prefix}.convs.{2 ** i}", out, first_downsample)
    out = conv_layer(sd, f"{prefix}.final_conv", out, 3)
    out = out.reshape(out.shape[0], -1)
    out = equal_linear(sd, f"{prefix}.space_linear", out, activation=True)
    return equal_linear(sd, f"{prefix}.out_linear", out)


def discriminator_forward(sd, x, use_aux_disc=False, alpha=1.):
    """Discriminator_MultiScale_Aux.forward L647-664 (aux D: first_downsample=True, L633-638)."""
    if use_aux_disc:
        b = x.shape[0] // 2
        main = discriminator_multiscale(sd, "main_disc", x[:b], alpha, False)
        aux = discriminator_multiscale(sd, "aux_disc", x[b:], alpha, True)
        return torch.cat([main, aux], 0)
    return discriminator_multiscale(sd, "main_disc", x, alpha, False)


# --------------------------------------------------------------------------------------
# Synthetic weights shared by the reference, the oracle and the CUDA path
# --------------------------------------------------------------------------------------
def synthetic_state_dict(template, seed=1234, sigma_bias=0.0, dtype=torch.float32):
    """Deterministic, platform-independent weights for a model whose ``state_dict()`` has the
    given {key: shape} template.  numpy PCG64 (not torch's RNG) so that the GPU box regenerates
    identical bytes.  Scales follow the reference's init distributions so activations are
    realistic: frequency_init(25) for FiLM linears (film_layer.py:16-22), kaiming for the rest.
    sigma_bias shifts siren.final_layer.bias (the "trained-like" variant of SURVEY.md §7.3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = OrderedDict()
    for key in sorted(template.keys()):
        shape = tuple(template[key])
        n = int(np.prod(shape)) if len(shape) else 1
        v = rng.standard_normal(n, dtype=np.float64).reshape(shape)
        u = rng.random(n, dtype=np.float64).reshape(shape) * 2 - 1
        if key.endswith("kernel"):                                   # Blur buffers
            k = np.array([1., 3., 3., 1.])
            k = np.outer(k, k)
            val = k / k.sum()
        elif ".norm." in key or ("base_net" in key and len(shape) == 1 and _is_layernorm(key, template)):
            val = (1.0 + 0.1 * v) if key.endswith("weight") else 0.1 * v
        elif key.endswith("bias"):
            if key.startswith("siren.final_layer"):
                val = 0.1 * u + sigma_bias
            elif "bias_fc" in key or "gain_fc" in key:
                val = 0.1 * u
            elif "siren" in key:
                val = u / math.sqrt(max(_fan_in_of_bias(key, template), 1))
            else:
                val = 0.05 * v
        elif key.endswith("weight"):
            if "siren.network" in key and ".linear." in key or "color_layer_sine.linear" in key:
                fan = shape[-1]
                val = u * math.sqrt(6 / fan) / 25
            elif key == "siren.network.0.layer.weight":              # piGAN first_layer_film_sine_init (siren.py:36-40)
                val = u / shape[-1]
            elif ".layer.weight" in key or (key.startswith("siren.color_layer_linear") and shape[0] == 3):
                val = u * math.sqrt(6 / shape[-1]) / 25              # piGAN frequency_init(25) (siren.py:75-81)
            elif key.startswith("siren.mapping_network.network.6"):   # last mapping layer * 0.25 (siren.py:66-67)
                val = 0.25 * v * math.sqrt(2 / (1 + 0.2 ** 2) / shape[-1])
            elif "gain_fc" in key or "bias_fc" in key:
                val = 0.25 * u / math.sqrt(shape[-1])
            elif "inr_net.network" in key and len(shape) == 3:       # (1,in,out) SinStyleMod
                val = v * math.sqrt(2 / (1 + 0.2 ** 2) / shape[2])
            elif "to_rgbs" in key or key.startswith("aux_to_rbg"):
                val = u * math.sqrt(6 / shape[-1]) / 25
            elif len(shape) == 4:                                    # EqualConv2d: N(0,1)
                val = v
            elif "space_linear" in key or "out_linear" in key:        # EqualLinear: N(0,1)
                val = v
            else:                                                     # nn.Linear kaiming(a=0.2)
                val = v * math.sqrt(2 / (1 + 0.2 ** 2) / shape[-1])
        else:
            val = 0.1 * v
        out[key] = torch.from_numpy(np.ascontiguousarray(val)).to(dtype).reshape(shape)
    return out


def _is_layernorm(key, template):
    stem = key.rsplit(".", 1)[0]
    w = template.get(stem + ".weight")
    return w is not None and len(tuple(w)) == 1


def _fan_in_of_bias(key, template):
    w = template.get(key.rsplit(".", 1)[0] + ".weight")
    return tuple(w)[-1] if w is not None and len(tuple(w)) >= 2 else 1


def generator_template(cfg=G_CFG):
    """{key: shape} of GeneratorNerfINR.state_dict(), derived from cfg."""
    t = OrderedDict()
    n, m = cfg["nerf_cfg"], cfg["mapping_nerf_cfg"]
    hd, sdim = n["hidden_dim"], n["style_dim"]

    def film(p, i, o):
        t[f"{p}.linear.weight"], t[f"{p}.linear.bias"] = (o, i), (o,)
        t[f"{p}.gain_fc.weight"], t[f"{p}.gain_fc.bias"] = (o, sdim), (o,)
        t[f"{p}.bias_fc.weight"], t[f"{p}.bias_fc.bias"] = (o, sdim), (o,)
    i = n["in_dim"]
    for l in range(n["hidden_layers"]):
        film(f"siren.network.{l}", i, hd)
        i = hd
    t["siren.final_layer.weight"], t["siren.final_layer.bias"] = (1, hd), (1,)
    film("siren.color_layer_sine", hd, hd // 2)
    t["siren.color_layer_linear.0.weight"], t["siren.color_layer_linear.0.bias"] = (n["rgb_dim"], hd // 2), (n["rgb_dim"],)

    def mapping(p, c):
        idx, d = 0, c["z_dim"]
        for l in range(c["base_layers"]):
            t[f"{p}.base_net.{idx}.weight"], t[f"{p}.base_net.{idx}.bias"] = (c["hidden_dim"], d), (c["hidden_dim"],)
            d = c["hidden_dim"]
            idx += 1
            if l != c["base_layers"] - 1:
                if c.get("add_norm"):
                    t[f"{p}.base_net.{idx}.weight"], t[f"{p}.base_net.{idx}.bias"] = (d,), (d,)
                    idx += 1
                idx += 1
        if c.get("norm_out"):
            t[f"{p}.base_net.{idx}.weight"], t[f"{p}.base_net.{idx}.bias"] = (d,), (d,)
    mapping("mapping_network_nerf", m)
    c = cfg["inr_cfg"]
    i = c["input_dim"]
    for name in CIPS_BLOCK_NAMES:
        o = c["hidden_dim"]
        for mod, (a, b) in (("mod1", (i, o)), ("mod2", (o, o))):
            p = f"inr_net.network.{name}.{mod}"
            t[f"{p}.weight"] = (1, a, b)
            t[f"{p}.modulation.weight"], t[f"{p}.modulation.bias"] = (a, c["style_dim"]), (a,)
            t[f"{p}.norm.weight"], t[f"{p}.norm.bias"] = (a,), (a,)
        t[f"inr_net.to_rgbs.{name}.linear.weight"], t[f"inr_net.to_rgbs.{name}.linear.bias"] = (c["pre_rgb_dim"], o), (c["pre_rgb_dim"],)
        i = o
    mapping("mapping_network_inr", cfg["mapping_inr_cfg"])
    t["aux_to_rbg.0.weight"], t["aux_to_rbg.0.bias"] = (3, n["rgb_dim"]), (3,)
    return t


# --------------------------------------------------------------------------------------
# Optimiser tail of the training step:
#   exp/cips3d/scripts/train.py:417-438 (D) / :468-491 (G)
#     clip_grad_norm_(parameters, grad_clip)  -- torch/nn/utils/clip_grad.py: norm of the per-tensor L2 norms,
#                                                clip_coef = max_norm / (total_norm + 1e-6), clamped to 1, grads *= coef
#     Adam.step()                             -- torch/optim/adam.py::_single_tensor_adam (weight_decay 0, amsgrad off)
#   exp/comm/comm_model_utils.py:99-121 EMA.update: target = target * decay + source * (1 - decay), skipped while
#     itr < start_itr.
# In place on the given lists of tensors; `step` is the count AFTER the increment (1 for the first update).
# Pinned against torch.optim.Adam + torch.nn.utils.clip_grad_norm_ + the reference's EMA class in
# tests/test_oracle_vs_reference.py::test_optimiser_tail_vs_torch_and_reference_ema.
# --------------------------------------------------------------------------------------
def clip_adam_ema_step(params, grads, exp_avg, exp_avg_sq, ema=None, *, step, lr, betas, eps=1e-8,
                       max_norm=None, ema_decay=None):
    total_norm = None
    if max_norm is not None:
        norms = [torch.linalg.vector_norm(g, 2.0) for g in grads]
        total_norm = torch.linalg.vector_norm(torch.stack(norms), 2.0)
        clip_coef = max_norm / (total_norm + 1e-6)
        clip_coef_clamped = torch.clamp(clip_coef, max=1.0)
        for g in grads:
            g.mul_(clip_coef_clamped)
    beta1, beta2 = betas
    bias_correction1 = 1 - beta1 ** step
    bias_correction2 = 1 - beta2 ** step
    step_size = lr / bias_correction1
    bias_correction2_sqrt = bias_correction2 ** 0.5
    for i, (p, g, m, v) in enumerate(zip(params, grads, exp_avg, exp_avg_sq)):
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / bias_correction2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)
        if ema is not None and ema_decay is not None:
            ema[i].copy_(ema[i] * ema_decay + p * (1 - ema_decay))
    return total_norm


# --------------------------------------------------------------------------------------
# pi-GAN surface: piGAN_lib/siren/siren.py:47-73 CustomMappingNetwork, :83-94 FiLMLayer,
# :97-152 TALLSIREN, :160-215 SPATIALSIRENBASELINE; piGAN_lib/generators/generators.py:12-96 ImplicitGenerator3d.forward,
# :110-204 staged_forward; piGAN_lib/generators/volumetric_rendering.py (same ray / camera / integration math as
# exp/comm/comm_utils.py + exp/pigan/pigan_utils.py restated above, with 3 colour channels).
# state_dict keys: siren.network.<i>.layer.{weight,bias}, siren.final_layer.*, siren.color_layer_sine.layer.*,
# siren.color_layer_linear.0.*, siren.mapping_network.network.{0,2,4,6}.*
# --------------------------------------------------------------------------------------
PIGAN_KWARGS = dict(fov=12, ray_start=0.88, ray_end=1.12, num_steps=12, h_stddev=0.3, v_stddev=0.155,
                    h_mean=math.pi * 0.5, v_mean=math.pi * 0.5, hierarchical_sample=True, sample_dist="gaussian",
                    clamp_mode="relu", last_back=False, white_back=False)        # curriculums.py CelebA, stage 0


def pigan_template(z_dim=256, hidden_dim=256, n_layers=8, map_hidden=256):
    t = OrderedDict()
    i = 3
    for l in range(n_layers):
        t[f"siren.network.{l}.layer.weight"], t[f"siren.network.{l}.layer.bias"] = (hidden_dim, i), (hidden_dim,)
        i = hidden_dim
    t["siren.final_layer.weight"], t["siren.final_layer.bias"] = (1, hidden_dim), (1,)
    t["siren.color_layer_sine.layer.weight"], t["siren.color_layer_sine.layer.bias"] = (hidden_dim, hidden_dim + 3), (hidden_dim,)
    t["siren.color_layer_linear.0.weight"], t["siren.color_layer_linear.0.bias"] = (3, hidden_dim), (3,)
    d = z_dim
    for j, idx in enumerate((0, 2, 4, 6)):
        o = map_hidden if j < 3 else (n_layers + 1) * hidden_dim * 2
        t[f"siren.mapping_network.network.{idx}.weight"], t[f"siren.mapping_network.network.{idx}.bias"] = (o, d), (o,)
        d = map_hidden
    return t


def pigan_mapping(sd, z):
    """CustomMappingNetwork.forward (siren.py:69-73): -> frequencies, phase_shifts (B, 9*256) each (raw, before *15+30)"""
    x = z
    for j, idx in enumerate((0, 2, 4, 6)):
        x = F.linear(x, sd[f"siren.mapping_network.network.{idx}.weight"], sd[f"siren.mapping_network.network.{idx}.bias"])
        if j < 3:
            x = F.leaky_relu(x, 0.2)
    h = x.shape[-1] // 2
    return x[..., :h], x[..., h:]


def pigan_siren(sd, points, ray_directions, frequencies, phase_shifts, gridwarp=True):
    """SPATIALSIRENBASELINE / TALLSIREN .forward_with_frequencies_phase_shifts (siren.py:133-152, 196-215):
    points, ray_directions (B,P,3), raw frequencies / phase shifts (B, 9*H) -> (B,P,4) = [sigmoid rgb(3), sigma]"""
    H = sd["siren.final_layer.weight"].shape[1]
    n_layers = sum(1 for k in sd if k.startswith("siren.network.") and k.endswith(".layer.weight"))
    freq = frequencies * 15 + 30
    x = points * (2 / 0.24) if gridwarp else points

    def film(prefix, x, f, ph):
        y = F.linear(x, sd[f"{prefix}.layer.weight"], sd[f"{prefix}.layer.bias"])
        return torch.sin(f.unsqueeze(1) * y + ph.unsqueeze(1))
    for i in range(n_layers):
        x = film(f"siren.network.{i}", x, freq[..., i * H:(i + 1) * H], phase_shifts[..., i * H:(i + 1) * H])
    sigma = F.linear(x, sd["siren.final_layer.weight"], sd["siren.final_layer.bias"])
    rgb = film("siren.color_layer_sine", torch.cat([ray_directions, x], -1), freq[..., -H:], phase_shifts[..., -H:])
    rgb = torch.sigmoid(F.linear(rgb, sd["siren.color_layer_linear.0.weight"], sd["siren.color_layer_linear.0.bias"]))
    return torch.cat([rgb, sigma], -1)


def pigan_forward(sd, z, draws, *, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev,
                  h_mean=math.pi * 0.5, v_mean=math.pi * 0.5, hierarchical_sample=True, sample_dist="gaussian",
                  lock_view_dependence=False, clamp_mode="relu", nerf_noise=0.0, white_back=False, last_back=False,
                  gridwarp=True, freq_phase=None, return_all=False, **_):
    """ImplicitGenerator3d.forward (generators.py:26-96); freq_phase=(frequencies, phase_shifts) overrides the mapping
    network (staged_forward's truncation, generators.py:121-126).  -> pixels (B,3,R,R) in [-1,1], pitch_yaw (B,2)"""
    assert sample_dist in ("gaussian", "normal")
    B, R, S = z.shape[0], img_size, num_steps
    dt = z.dtype
    dirs_cam, z_vals = initial_rays(R, fov, ray_start, ray_end, S, dtype=dt)
    origin, pitch, yaw = camera_origin(draws["yaw_n"].to(dt), draws["pitch_n"].to(dt), h_stddev, v_stddev, h_mean, v_mean)
    c2w = cam2world(-origin, origin)
    pts, zc, dirs_w, t = world_points(dirs_cam, z_vals, draws["jitter_u"].to(dt), c2w)   # (B,HW,S,3), (B,HW,S), (B,HW,3), (B,3)
    origins = t[:, None, :].expand(-1, R * R, -1)
    dirs_exp = dirs_w[:, :, None, :].expand(-1, -1, S, -1).reshape(B, R * R * S, 3)
    if lock_view_dependence:
        dirs_exp = torch.zeros_like(dirs_exp)
        dirs_exp[..., -1] = -1
    fr, ph = freq_phase if freq_phase is not None else pigan_mapping(sd, z)
    coarse = pigan_siren(sd, pts.reshape(B, -1, 3), dirs_exp, fr, ph, gridwarp).reshape(B, R * R, S, 4)
    noise_c = draws["noise_c"].to(dt) * nerf_noise if (draws.get("noise_c") is not None) else None
    if hierarchical_sample:
        fz = fine_z_vals(coarse, zc, noise_c, draws["pdf_u"].to(dt), clamp_mode=clamp_mode, dim_rgb=3)
        fpts = origins[:, :, None, :] + dirs_w[:, :, None, :] * fz[..., None]
        fine = pigan_siren(sd, fpts.reshape(B, -1, 3), dirs_exp, fr, ph, gridwarp).reshape(B, R * R, S, 4)
        all_out = torch.cat([fine, coarse], -2)
        all_z = torch.cat([fz, zc], -1)
        all_z, idx = torch.sort(all_z, dim=-1)
        all_out = torch.gather(all_out, -2, idx[..., None].expand(-1, -1, -1, 4))
    else:
        all_out, all_z = coarse, zc
    noise_f = draws["noise_f"].to(dt) * nerf_noise
    rgb, depth, w = integrate(all_out, all_z, noise_f, clamp_mode=clamp_mode, last_back=last_back, white_back=white_back,
                              dim_rgb=3)
    pixels = rgb.reshape(B, R, R, 3).permute(0, 3, 1, 2).contiguous() * 2 - 1
    py = torch.cat([pitch, yaw], -1)
    if return_all:
        return pixels, py, dict(coarse=coarse, all_z=all_z, rgb=rgb, depth=depth, weights=w)
    return pixels, py


# --------------------------------------------------------------------------------------
# Image export of the inference paths: fp32 image(s) in [-1, 1] -> uint8, channels last.
#   'save_image'     exp/cips3d/scripts/gen_images.py:64, sample_images.py:73:
#                    torchvision.utils.save_image(img, path, normalize=True, value_range=(-1, 1)) on one image:
#                    make_grid.norm_ip  img.clamp_(low, high); img.sub_(low).div_(max(high - low, 1e-5))
#                    save_image         grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to("cpu", torch.uint8)
#                    (torchvision is a dependency of the reference, not vendored in it; pinned to the installed
#                    torchvision 0.26 through a lossless PNG round trip in tests/test_image_export_cpu.py)
#   'tensor_to_pil'  exp/cips3d/models/st_web.py:44-46:  img * 0.5 + 0.5, then the same mul / add / clamp / cast
#   'to_pil'         exp/comm/comm_utils.py:21-24: (frame + 1) * 0.5, then torchvision to_pil_image:
#                    (npimg * 255).astype(np.uint8)  -- a truncation without rounding or clamping
# Accepts (C, H, W) or (B, C, H, W); returns (H, W, C) or (B, H, W, C).
# --------------------------------------------------------------------------------------
def image_to_u8(img, mode="save_image", value_range=(-1, 1)):
    x = img.detach().to("cpu", torch.float32).clone()
    if mode == "save_image":
        low, high = value_range
        x.clamp_(min=low, max=high)
        x.sub_(low).div_(max(high - low, 1e-5))
        y = x.mul(255).add_(0.5).clamp_(0, 255)
    elif mode == "tensor_to_pil":
        x = x * 0.5 + 0.5
        y = x.mul(255).add_(0.5).clamp_(0, 255)
    elif mode == "to_pil":
        x = (x + 1) * 0.5
        y = torch.from_numpy(x.numpy() * 255).clamp_(0, 255)      # in range for tanh outputs; the cast alone is undefined outside
    else:
        raise ValueError(mode)
    perm = (1, 2, 0) if y.dim() == 3 else (0, 2, 3, 1)
    return y.permute(*perm).to(torch.uint8).contiguous()
