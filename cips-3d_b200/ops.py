"""Tensor-level entry points over the C-ABI.  Everything here launches hand-written CUDA
kernels from libcips3d_b200.so on the current torch stream; torch only provides device
memory.  No CPU path, no PyTorch fallback for the forward ops."""
import ctypes as C
import math

import numpy as np
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import CipsParams, CipsWeights, PiganWeights, RayIO, RayParams, SirenWeights, check, load, ptr, stream_ptr

CLAMP_MODES = {"relu": 0, "softplus": 1}
# bench.py sets this to a dict to collect CUDA-event pairs around the two hot entry points
PROFILE = None


def _prof_begin(key):
    if PROFILE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _prof_end(key, e0):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        PROFILE.setdefault(key, []).append((e0, e1))


def _f32c(t, name):
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise _lib.C3dError(f"{name}: expected float32, got {t.dtype}")
    return t.contiguous()


def _is_cl(t):
    """4-D tensor stored channels-last (N, H, W, C in memory) and not at the same time NCHW-contiguous"""
    return t is not None and t.dim() == 4 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last)


def _like(t, ref):
    """t in the memory format of ref (channels-last or NCHW-contiguous): the discriminator's native ops keep whatever layout
    their input has, so a channels-last discriminator never transposes"""
    if t is None:
        return None
    return t.contiguous(memory_format=torch.channels_last) if _is_cl(ref) else t.contiguous()


def z_cam_from_fov(fov):
    """-1 / tan(fov*pi/360) evaluated like exp/comm/comm_utils.py:396 (fp32 division)."""
    return float(np.float32(-1.0) / np.float32(np.tan((2 * math.pi * fov / 360) / 2)))


# --------------------------------------------------------------------------------------
# volumetric renderer
# --------------------------------------------------------------------------------------
def render_features(siren, film, cam2world, jitter_u, pdf_u=None, noise_c=None, noise_f=None, *,
                    img_size, fov, ray_start, ray_end, num_steps, hierarchical_sample=True,
                    clamp_mode="relu", noise_std=0.0, white_back=False, last_back=False,
                    ray_idx=None, ray_offset=0, n_rays=None, impl=None, debug=False,
                    want_depth=False, want_weights=False):
    """Fused rays -> FiLM-SIREN -> resample -> composite (c3d_ray_siren_fwd).

    siren: dict of the NeRFNetwork parameters {w0,b0,w1,b1,w_sigma,b_sigma,wc,bc,wl,bl}
    film:  dict {gamma0,beta0,gamma1,beta1,gammac,betac}, each (B,C)
    cam2world (B,4,4); jitter_u (B,R*R,S); pdf_u (B*N,S); noise_c (B,N,S); noise_f (B,N,nS)
    Returns dict(pixels_fea (B,N,32)[, depth, weights, coarse, fine, all_z])."""
    lib = load()
    if clamp_mode not in CLAMP_MODES:
        raise AssertionError("Need to choose clamp mode")       # pigan_utils.py:252-253
    B = cam2world.shape[0]
    S = int(num_steps)
    HW = img_size * img_size
    if ray_idx is not None:
        ray_idx = ray_idx.to(torch.int32).contiguous()
        N = ray_idx.numel()
    else:
        N = HW - ray_offset if n_rays is None else int(n_rays)
    nS = 2 * S if hierarchical_sample else S
    dev = cam2world.device
    p = RayParams(batch=B, img_size=img_size, num_steps=S, n_rays=N, ray_offset=int(ray_offset),
                  hierarchical=int(bool(hierarchical_sample)), clamp_mode=CLAMP_MODES[clamp_mode],
                  white_back=int(bool(white_back)), last_back=int(bool(last_back)),
                  impl=_lib.default_impl("ray") if impl is None else impl,
                  z_cam=z_cam_from_fov(fov), ray_start=float(ray_start), ray_end=float(ray_end),
                  noise_std=float(noise_std))
    keep = []  # keep contiguous copies alive until launch

    def dp(t, name, shape=None):
        t = _f32c(t, name)
        if t is not None:
            if shape is not None and tuple(t.shape) != tuple(shape):
                raise _lib.C3dError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
            keep.append(t)
        return ptr(t)

    w = SirenWeights(
        w0=dp(siren["w0"], "w0", (128, 3)), b0=dp(siren["b0"], "b0", (128,)),
        w1=dp(siren["w1"], "w1", (128, 128)), b1=dp(siren["b1"], "b1", (128,)),
        w_sigma=dp(siren["w_sigma"], "w_sigma", (1, 128)), b_sigma=dp(siren["b_sigma"], "b_sigma", (1,)),
        wc=dp(siren["wc"], "wc", (64, 128)), bc=dp(siren["bc"], "bc", (64,)),
        wl=dp(siren["wl"], "wl", (32, 64)), bl=dp(siren["bl"], "bl", (32,)),
        gamma0=dp(film["gamma0"], "gamma0", (B, 128)), beta0=dp(film["beta0"], "beta0", (B, 128)),
        gamma1=dp(film["gamma1"], "gamma1", (B, 128)), beta1=dp(film["beta1"], "beta1", (B, 128)),
        gammac=dp(film["gammac"], "gammac", (B, 64)), betac=dp(film["betac"], "betac", (B, 64)))
    out = {"pixels_fea": torch.empty((B, N, 32), device=dev, dtype=torch.float32)}
    if want_depth or debug:
        out["depth"] = torch.empty((B, N), device=dev, dtype=torch.float32)
    if want_weights or debug:
        out["weights"] = torch.empty((B, N, nS), device=dev, dtype=torch.float32)
    if debug:
        out["coarse"] = torch.empty((B, N, S, 33), device=dev, dtype=torch.float32)
        out["all_z"] = torch.empty((B, N, nS), device=dev, dtype=torch.float32)
        if hierarchical_sample:
            out["fine"] = torch.empty((B, N, S, 33), device=dev, dtype=torch.float32)
    if ray_idx is not None:
        keep.append(ray_idx)
    io = RayIO(
        cam2world=dp(cam2world, "cam2world", (B, 4, 4)),
        ray_idx=ptr(ray_idx) if ray_idx is not None else None,
        jitter_u=dp(jitter_u, "jitter_u", (B, HW, S)),
        noise_c=dp(noise_c, "noise_c", (B, N, S)) if (noise_c is not None and noise_std != 0) else None,
        pdf_u=dp(pdf_u, "pdf_u", (B * N, S)) if hierarchical_sample else None,
        noise_f=dp(noise_f, "noise_f", (B, N, nS)) if (noise_f is not None and noise_std != 0) else None,
        pixels_fea=ptr(out["pixels_fea"]), depth=ptr(out.get("depth")), weights=ptr(out.get("weights")),
        dbg_coarse=ptr(out.get("coarse")), dbg_fine=ptr(out.get("fine")), dbg_all_z=ptr(out.get("all_z")))
    wsb = lib.c3d_ray_siren_workspace_bytes(C.byref(p))
    ws = torch.empty((max(wsb, 4) + 3) // 4, device=dev, dtype=torch.float32)
    ev = _prof_begin("ray")
    check(lib.c3d_ray_siren_fwd(C.byref(p), C.byref(w), C.byref(io), ptr(ws), wsb, stream_ptr()),
          "c3d_ray_siren_fwd")
    _prof_end("ray", ev)
    return out


# --------------------------------------------------------------------------------------
# per-pixel CIPS MLP
# --------------------------------------------------------------------------------------
def cips_forward(x, weights, style1p, demod, rgb_w, rgb_b, *, n_blocks=9, skip_from=4, rgb_from=3,
                 impl=None, return_hidden=False):
    """x (B,N,in) -> tanh(rgb) (B,N,3).  weights[l] (in_l,out), style1p[l] (B,in_l),
    demod[l] (B,out), rgb_w[b] (3,hidden) / rgb_b[b] (3,) (None for blocks < rgb_from).
    return_hidden=False (the image alone) lets the tensor-core kernel keep the skip connections' stream as fp16 (csrc/cips_tc.cu ResT;
    C3D_CIPS_RES16=0 turns that off); with return_hidden=True the stream is fp32 and the image is bit-identical to that setting's."""
    lib = load()
    x = _f32c(x, "x")
    B, N, in_dim = x.shape
    hidden = weights[0].shape[1]
    p = CipsParams(batch=B, n_pix=N, in_dim=in_dim, hidden=hidden, n_blocks=n_blocks, skip_from=skip_from,
                   rgb_from=rgb_from, impl=_lib.default_impl("cips") if impl is None else impl)
    keep = []
    cw = CipsWeights()
    for l in range(2 * n_blocks):
        for arr, src, nm in ((cw.w, weights, "w"), (cw.style1p, style1p, "style1p"), (cw.demod, demod, "demod")):
            t = _f32c(src[l], f"{nm}[{l}]")
            keep.append(t)
            arr[l] = ptr(t)
    for b in range(n_blocks):
        if b >= rgb_from:
            tw, tb = _f32c(rgb_w[b], "rgb_w"), _f32c(rgb_b[b], "rgb_b")
            keep += [tw, tb]
            cw.rgb_w[b], cw.rgb_b[b] = ptr(tw), ptr(tb)
    rgb = torch.empty((B, N, 3), device=x.device, dtype=torch.float32)
    hid = torch.empty((B, N, hidden), device=x.device, dtype=torch.float32) if return_hidden else None
    wsb = lib.c3d_cips_workspace_bytes(C.byref(p))
    ws = torch.empty((max(wsb, 4) + 3) // 4, device=x.device, dtype=torch.float32)
    ev = _prof_begin("cips")
    check(lib.c3d_cips_fwd(C.byref(p), C.byref(cw), ptr(x), ptr(rgb), ptr(hid), ptr(ws), wsb, stream_ptr()),
          "c3d_cips_fwd")
    _prof_end("cips", ev)
    return (rgb, hid) if return_hidden else rgb


# --------------------------------------------------------------------------------------
# discriminator ops (same call surface as exp/comm/op/{fused_act,upfirdn2d}.py)
# --------------------------------------------------------------------------------------
def bias_act(x, bias=None, ref=None, act=3, grad=0, alpha=0.2, scale=2 ** 0.5):
    """fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale) -- fused_bias_act.cpp:11-20.
    A channels-last input (N, H, W, C in memory) is processed in place of layout: the bias index is then i % C (step 1)."""
    lib = load()
    if x.dtype != torch.float32:
        raise _lib.C3dError(f"x: expected float32, got {x.dtype}")
    cl = _is_cl(x)
    if not cl:
        x = x.contiguous()
    y = torch.empty_like(x)                       # keeps x's strides
    if x.numel() == 0:
        return y
    b = _f32c(bias, "bias") if bias is not None and bias.numel() else None
    r = None
    if ref is not None and ref.numel():
        if ref.dtype != torch.float32:
            raise _lib.C3dError(f"ref: expected float32, got {ref.dtype}")
        r = _like(ref, x)
    step_b = 1
    if not cl:
        for i in range(2, x.dim()):
            step_b *= x.size(i)                                 # fused_bias_act_kernel.cu:67-69
    check(lib.c3d_bias_act(ptr(x), ptr(b), ptr(r), ptr(y), x.numel(), step_b,
                           b.numel() if b is not None else 1, act, grad, alpha, scale, stream_ptr()),
          "c3d_bias_act")
    return y


class _FusedLeakyReLUBackward(Function):            # fused_act.py:19-49
    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        grad_input = bias_act(grad_output, None, out, 3, 1, negative_slope, scale)
        dim = [0] + list(range(2, grad_input.ndim))
        return grad_input, grad_input.sum(dim).detach()

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        return bias_act(_like(gradgrad_input, out), gradgrad_bias, out, 3, 1, ctx.negative_slope,
                        ctx.scale), None, None, None


class _FusedLeakyReLU(Function):                    # fused_act.py:52-70
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = bias_act(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        gi, gb = _FusedLeakyReLUBackward.apply(_like(grad_output, out), out, ctx.negative_slope, ctx.scale)
        return gi, gb, None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    return _FusedLeakyReLU.apply(input, bias, negative_slope, scale)


def _upfirdn2d_raw(x, kernel, up, down, pad):
    lib = load()
    kernel = _f32c(kernel, "kernel")
    kh, kw = kernel.shape
    if _is_cl(x) and x.dtype == torch.float32 and up == (1, 1) and down == (1, 1) and (kh, kw) == (4, 4):
        # channels-last blur (every Blur of the discriminator and its backward): c3d_blur_nhwc, output channels-last too
        B, Cc, H, W = x.shape
        out_h, out_w = H + pad[2] + pad[3] - 3, W + pad[0] + pad[1] - 3
        y = torch.empty((B, Cc, out_h, out_w), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        check(lib.c3d_blur_nhwc(ptr(x), ptr(kernel), ptr(y), B, H, W, Cc, pad[0], pad[1], pad[2], pad[3], stream_ptr()), "c3d_blur_nhwc")
        return y
    x = _f32c(x, "x")
    B, Cc, H, W = x.shape
    out_h = (H * up[1] + pad[2] + pad[3] - kh) // down[1] + 1
    out_w = (W * up[0] + pad[0] + pad[1] - kw) // down[0] + 1
    y = torch.empty((B, Cc, out_h, out_w), device=x.device, dtype=torch.float32)
    check(lib.c3d_upfirdn2d(ptr(x), ptr(kernel), ptr(y), B * Cc, H, W, kh, kw, up[0], up[1], down[0], down[1],
                            pad[0], pad[1], pad[2], pad[3], stream_ptr()), "c3d_upfirdn2d")
    return y


# ---------------------------------------------------------------------------------------------------------------------
# Discriminator convolution (row D1: exp/cips3d/models/discriminator.py:20-54 calls F.conv2d).  The arithmetic is the library's
# (cuDNN / CUTLASS implicit GEMM, TF32 as torch and the reference default to), but the AUTOGRAD STRUCTURE is this repo's: the
# R1 penalty (train.py:379-386) differentiates ||dD/dx||^2, i.e. runs the DOUBLE backward of every convolution, and torch's
# generic `_convolution_double_backward` evaluates the weight gradient of that pass as a convolution with batch and channel
# dimensions swapped -- a "kernel" the size of the feature map, which cuDNN serves with a legacy indexed implicit-GEMM:
# 16 calls of 9.2 ms = 40 % of a whole config-5 train step (profiles/r02b_train_profiles.md).  Here the three primitive
# products (fprop, dgrad, wgrad) are autograd Functions whose backwards are expressed in terms of each other, so every pass of
# every order runs one of the three purpose-built kernels.
def _conv_args(stride, padding):
    return [stride, stride], [padding, padding], [1, 1], False, [0, 0], 1


def _conv_fprop(x, w, b, stride, padding):
    return torch.ops.aten.convolution(x, w, b, *_conv_args(stride, padding))


def _conv_dgrad(gy, w, x_shape, stride, padding):
    x_like = gy.new_empty(1).expand(x_shape)          # only its sizes are read (torch.nn.grad.conv2d_input does the same)
    return torch.ops.aten.convolution_backward(gy, x_like, w, None, *_conv_args(stride, padding), (True, False, False))[0]


def _conv_wgrad(gy, x, w_shape, stride, padding):
    w_like = gy.new_empty(1).expand(w_shape)
    return torch.ops.aten.convolution_backward(gy, x, w_like, None, *_conv_args(stride, padding), (False, True, False))[1]


class _Conv2dDgrad(Function):
    """gx = dgrad(gy, w); its backward (the double backward of the convolution): d gy = fprop(ggx, w), d w = wgrad(gy, ggx)."""

    @staticmethod
    def forward(ctx, gy, w, x_shape, stride, padding):
        ctx.save_for_backward(gy, w)
        ctx.cfg = (tuple(x_shape), stride, padding)
        return _conv_dgrad(gy, w, x_shape, stride, padding)

    @staticmethod
    @once_differentiable
    def backward(ctx, ggx):
        gy, w = ctx.saved_tensors
        x_shape, stride, padding = ctx.cfg
        ggx = _like(ggx, gy)
        d_gy = _conv_fprop(ggx, w, None, stride, padding) if ctx.needs_input_grad[0] else None
        d_w = _conv_wgrad(gy, ggx, w.shape, stride, padding) if ctx.needs_input_grad[1] else None
        return d_gy, d_w, None, None, None


class _Conv2dWgrad(Function):
    """gw = wgrad(gy, x); its backward: d gy = fprop(x, ggw), d x = dgrad(gy, ggw)."""

    @staticmethod
    def forward(ctx, gy, x, w_shape, stride, padding):
        ctx.save_for_backward(gy, x)
        ctx.cfg = (tuple(w_shape), stride, padding)
        return _conv_wgrad(gy, x, w_shape, stride, padding)

    @staticmethod
    @once_differentiable
    def backward(ctx, ggw):
        gy, x = ctx.saved_tensors
        _, stride, padding = ctx.cfg
        ggw = _like(ggw, x)
        d_gy = _conv_fprop(x, ggw, None, stride, padding) if ctx.needs_input_grad[0] else None
        d_x = _conv_dgrad(gy, ggw, x.shape, stride, padding) if ctx.needs_input_grad[1] else None
        return d_gy, d_x, None, None, None


class _Conv2d(Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, padding):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, b is not None)
        return _conv_fprop(x, w, b, stride, padding)

    @staticmethod
    def backward(ctx, gy):                     # differentiable again: R1 runs autograd.grad(..., create_graph=True) through it
        x, w = ctx.saved_tensors
        stride, padding, has_bias = ctx.cfg
        gy = _like(gy, x)
        gx = _Conv2dDgrad.apply(gy, w, x.shape, stride, padding) if ctx.needs_input_grad[0] else None
        gw = _Conv2dWgrad.apply(gy, x, w.shape, stride, padding) if ctx.needs_input_grad[1] else None
        gb = gy.sum((0, 2, 3)) if has_bias and ctx.needs_input_grad[2] else None
        return gx, gw, gb, None, None


def conv2d(input, weight, bias=None, stride=1, padding=0):
    """F.conv2d(input, weight, bias, stride, padding) for the discriminator's EqualConv2d (square stride / padding, groups 1,
    dilation 1), twice differentiable, with every backward pass of every order on a purpose-built fprop / dgrad / wgrad kernel."""
    if _is_cl(input):      # channels-last activations: channels-last weights too, and cuDNN's NHWC kernels run without transposes
        return _Conv2d.apply(input, weight.contiguous(memory_format=torch.channels_last), bias, int(stride), int(padding))
    return _Conv2d.apply(input.contiguous(), weight.contiguous(), bias, int(stride), int(padding))


class _UpFirDn2dBackward(Function):                 # upfirdn2d.py:18-85
    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size):
        grad_input = _upfirdn2d_raw(grad_output, grad_kernel, down, up, g_pad)
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad, ctx.cl = up, down, pad, _is_cl(grad_output)
        assert tuple(grad_input.shape) == tuple(in_size)
        return grad_input

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        gg = gradgrad_input.contiguous(memory_format=torch.channels_last) if ctx.cl else gradgrad_input.contiguous()
        return _upfirdn2d_raw(gg, kernel, ctx.up, ctx.down, ctx.pad), None, None, None, None, None, None, None


class _UpFirDn2d(Function):                         # upfirdn2d.py:88-141
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kh, kw = kernel.shape
        _, _, in_h, in_w = input.shape
        out = _upfirdn2d_raw(input, kernel, up, down, pad)
        out_h, out_w = out.shape[2:]
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        ctx.in_size = tuple(input.shape)
        ctx.cl = _is_cl(input)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        ctx.g_pad = (kw - pad_x0 - 1, in_w * up_x - out_w * down_x + pad_x0 - up_x + 1,
                     kh - pad_y0 - 1, in_h * up_y - out_h * down_y + pad_y0 - up_y + 1)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        go = grad_output.contiguous(memory_format=torch.channels_last) if ctx.cl else grad_output.contiguous()
        gi = _UpFirDn2dBackward.apply(go, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad, ctx.g_pad, ctx.in_size)
        return gi, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """exp/comm/op/upfirdn2d.py:144-149."""
    return _UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))


def selftest_umma(a, b, a_in_tmem=False):
    """D = A @ B^T through one tcgen05 tile (fp16 operands, fp32 accumulate)."""
    lib = load()
    a, b = _f32c(a, "a"), _f32c(b, "b")
    assert a.shape[0] == 128 and a.shape[1] == b.shape[1]
    d = torch.empty((128, b.shape[0]), device=a.device, dtype=torch.float32)
    check(lib.c3d_selftest_umma(ptr(a), ptr(b), ptr(d), b.shape[0], a.shape[1], int(a_in_tmem), stream_ptr()),
          "c3d_selftest_umma")
    return d


def selftest_umma_pair(a, b):
    """D = A @ B^T through one tcgen05 cta_group::2 tile (cluster of two CTAs, M = 256; fp16 operands, fp32 accumulate)."""
    lib = load()
    a, b = _f32c(a, "a"), _f32c(b, "b")
    assert a.shape[0] == 256 and a.shape[1] == b.shape[1]
    d = torch.empty((256, b.shape[0]), device=a.device, dtype=torch.float32)
    check(lib.c3d_selftest_umma_pair(ptr(a), ptr(b), ptr(d), b.shape[0], a.shape[1], stream_ptr()), "c3d_selftest_umma_pair")
    return d


# --------------------------------------------------------------------------------------
# pi-GAN renderer (piGAN_lib ImplicitGenerator3d + TALLSIREN / SPATIALSIRENBASELINE)
# --------------------------------------------------------------------------------------
def pigan_render(siren, cam2world, jitter_u, pdf_u=None, noise_c=None, noise_f=None, *, img_size, fov, ray_start, ray_end,
                 num_steps, hierarchical_sample=True, clamp_mode="relu", noise_std=0.0, white_back=False, last_back=False,
                 lock_view=False, ray_idx=None, ray_offset=0, n_rays=None, debug=False, want_depth=False, want_weights=False):
    """rays -> 8-layer FiLM-SIREN (hidden 256, view-dependent colour) -> resample -> composite (c3d_pigan_render_fwd).
    siren: dict from pigan._Siren.kernel_weights (w, b: lists per layer; freq, phase: lists of (B,hidden) incl. the colour
    layer; w_sigma, b_sigma, wc, bc, wl, bl; hidden; gridwarp).  Returns dict(rgb (B,N,3) in [0,1][, depth, weights, coarse,
    fine, all_z])."""
    lib = load()
    if clamp_mode not in CLAMP_MODES:
        raise AssertionError("Need to choose clamp mode")
    B = cam2world.shape[0]
    S = int(num_steps)
    HW = img_size * img_size
    if ray_idx is not None:
        ray_idx = ray_idx.to(torch.int32).contiguous()
        N = ray_idx.numel()
    else:
        N = HW - ray_offset if n_rays is None else int(n_rays)
    nS = 2 * S if hierarchical_sample else S
    dev = cam2world.device
    H, L = int(siren["hidden"]), len(siren["w"])
    p = RayParams(batch=B, img_size=img_size, num_steps=S, n_rays=N, ray_offset=int(ray_offset),
                  hierarchical=int(bool(hierarchical_sample)), clamp_mode=CLAMP_MODES[clamp_mode],
                  white_back=int(bool(white_back)), last_back=int(bool(last_back)), impl=_lib.IMPL_SIMT,
                  z_cam=z_cam_from_fov(fov), ray_start=float(ray_start), ray_end=float(ray_end), noise_std=float(noise_std))
    keep = []

    def dp(t, name, shape=None):
        t = _f32c(t, name)
        if t is not None:
            if shape is not None and tuple(t.shape) != tuple(shape):
                raise _lib.C3dError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
            keep.append(t)
        return ptr(t)

    w = PiganWeights(n_layers=L, hidden=H, gridwarp=int(bool(siren["gridwarp"])))
    for l in range(L):
        w.w[l] = dp(siren["w"][l], f"w[{l}]", (H, 3 if l == 0 else H))
        w.b[l] = dp(siren["b"][l], f"b[{l}]", (H,))
    for l in range(L + 1):
        w.freq[l] = dp(siren["freq"][l], f"freq[{l}]", (B, H))
        w.phase[l] = dp(siren["phase"][l], f"phase[{l}]", (B, H))
    w.w_sigma, w.b_sigma = dp(siren["w_sigma"], "w_sigma", (1, H)), dp(siren["b_sigma"], "b_sigma", (1,))
    w.wc, w.bc = dp(siren["wc"], "wc", (H, H + 3)), dp(siren["bc"], "bc", (H,))
    w.wl, w.bl = dp(siren["wl"], "wl", (3, H)), dp(siren["bl"], "bl", (3,))
    out = {"rgb": torch.empty((B, N, 3), device=dev, dtype=torch.float32)}
    if want_depth or debug:
        out["depth"] = torch.empty((B, N), device=dev, dtype=torch.float32)
    if want_weights or debug:
        out["weights"] = torch.empty((B, N, nS), device=dev, dtype=torch.float32)
    if debug:
        out["coarse"] = torch.empty((B, N, S, 4), device=dev, dtype=torch.float32)
        out["all_z"] = torch.empty((B, N, nS), device=dev, dtype=torch.float32)
        if hierarchical_sample:
            out["fine"] = torch.empty((B, N, S, 4), device=dev, dtype=torch.float32)
    if ray_idx is not None:
        keep.append(ray_idx)
    io = RayIO(
        cam2world=dp(cam2world, "cam2world", (B, 4, 4)), ray_idx=ptr(ray_idx) if ray_idx is not None else None,
        jitter_u=dp(jitter_u, "jitter_u", (B, HW, S)),
        noise_c=dp(noise_c, "noise_c", (B, N, S)) if (noise_c is not None and noise_std != 0) else None,
        pdf_u=dp(pdf_u, "pdf_u", (B * N, S)) if hierarchical_sample else None,
        noise_f=dp(noise_f, "noise_f", (B, N, nS)) if (noise_f is not None and noise_std != 0) else None,
        pixels_fea=ptr(out["rgb"]), depth=ptr(out.get("depth")), weights=ptr(out.get("weights")),
        dbg_coarse=ptr(out.get("coarse")), dbg_fine=ptr(out.get("fine")), dbg_all_z=ptr(out.get("all_z")))
    wsb = lib.c3d_pigan_workspace_bytes(C.byref(p))
    ws = torch.empty((max(wsb, 4) + 3) // 4, device=dev, dtype=torch.float32)
    ev = _prof_begin("pigan")
    check(lib.c3d_pigan_render_fwd(C.byref(p), C.byref(w), C.byref(io), int(bool(lock_view)), ptr(ws), wsb, stream_ptr()),
          "c3d_pigan_render_fwd")
    _prof_end("pigan", ev)
    return out


# --------------------------------------------------------------------------------------
# CIPS MLP with a native backward (first part): fused forward that stashes the activations,
# gradient chain on tcgen05 (c3d_cips_bwd), weight gradients as library GEMMs over the two stashes
# --------------------------------------------------------------------------------------
def _cips_structs(x, weights, style1p, demod, rgb_w, rgb_b, n_blocks, skip_from, rgb_from):
    x = _f32c(x, "x")
    B, N, in_dim = x.shape
    hidden = weights[0].shape[1]
    p = CipsParams(batch=B, n_pix=N, in_dim=in_dim, hidden=hidden, n_blocks=n_blocks, skip_from=skip_from, rgb_from=rgb_from,
                   impl=_lib.IMPL_TC)
    keep = [x]
    cw = CipsWeights()
    for l in range(2 * n_blocks):
        for arr, src, nm in ((cw.w, weights, "w"), (cw.style1p, style1p, "style1p"), (cw.demod, demod, "demod")):
            t = _f32c(src[l].detach(), f"{nm}[{l}]")
            keep.append(t)
            arr[l] = ptr(t)
    for b in range(n_blocks):
        if b >= rgb_from:
            tw = _f32c(rgb_w[b].detach(), "rgb_w")
            keep.append(tw)
            cw.rgb_w[b] = ptr(tw)
            if rgb_b is not None:
                tb = _f32c(rgb_b[b].detach(), "rgb_b")
                keep.append(tb)
                cw.rgb_b[b] = ptr(tb)
    return x, p, cw, keep


def cips_forward_train(x, weights, style1p, demod, rgb_w, rgb_b, *, n_blocks=9, skip_from=4, rgb_from=3):
    """c3d_cips_fwd_train: -> rgb (B,N,3), acts (2*n_blocks, B, N, hidden) fp16 (every layer's output as the next layer consumed
    it), zsign (2*n_blocks, B, N, hidden/16) int16 (sign bits of z_l on the residual layers)"""
    lib = load()
    x, p, cw, keep = _cips_structs(x, weights, style1p, demod, rgb_w, rgb_b, n_blocks, skip_from, rgb_from)
    B, N, _ = x.shape
    rgb = torch.empty((B, N, 3), device=x.device, dtype=torch.float32)
    acts = torch.empty((2 * n_blocks, B, N, p.hidden), device=x.device, dtype=torch.float16)
    zsign = torch.empty((2 * n_blocks, B, N, p.hidden // 16), device=x.device, dtype=torch.int16)
    wsb = lib.c3d_cips_workspace_bytes(C.byref(p))
    ws = torch.empty((max(wsb, 4) + 3) // 4, device=x.device, dtype=torch.float32)
    check(lib.c3d_cips_fwd_train(C.byref(p), C.byref(cw), ptr(x), ptr(rgb), ptr(acts), ptr(zsign), ptr(ws), wsb, stream_ptr()),
          "c3d_cips_fwd_train")
    return rgb, acts, zsign


def cips_backward_chain(x, acts, zsign, g_pre_scaled, weights, style1p, demod, rgb_w, *, n_blocks=9, skip_from=4, rgb_from=3, want_dx=True):
    """c3d_cips_bwd: g_pre_scaled (B,N,3) = S * dL/d(rgb before tanh) -> dz (2*n_blocks,B,N,hidden) fp16 = S*dZ_l, dx (B,N,in) = S*dL/dx"""
    lib = load()
    x, p, cw, keep = _cips_structs(x, weights, style1p, demod, rgb_w, None, n_blocks, skip_from, rgb_from)
    B, N, in_dim = x.shape
    g = _f32c(g_pre_scaled, "g")
    dz = torch.empty_like(acts)
    dx = torch.empty((B, N, in_dim), device=x.device, dtype=torch.float32) if want_dx else None
    wsb = lib.c3d_cips_bwd_workspace_bytes(C.byref(p))
    ws = torch.empty((max(wsb, 4) + 3) // 4, device=x.device, dtype=torch.float32)
    check(lib.c3d_cips_bwd(C.byref(p), C.byref(cw), ptr(acts), ptr(zsign), ptr(g), ptr(dz), ptr(dx), ptr(ws), wsb, stream_ptr()), "c3d_cips_bwd")
    return dz, dx


def _bmm_f32(a16, b16):
    """fp16 x fp16 -> fp32 batched GEMM (tensor cores on the GPU: plain library GEMM)"""
    if a16.is_cuda:
        return torch.bmm(a16, b16, out_dtype=torch.float32)
    return torch.bmm(a16.float(), b16.float())


class CipsMLPFunction(Function):
    """y = CIPSNet core (generator.py:1107-1154) with a native backward.  Inputs: x (B,N,in); per layer W (in,out), s1p (B,in),
    demod (B,out); per ToRGB block w (3,512), b (3).  forward = c3d_cips_fwd_train, backward = c3d_cips_bwd + library GEMMs:
        T_l[b]  = X_l[b]^T dZ_l[b]                      (dL/dW''_l, W''[k][n] = s1p[k] W[k][n] d[n])
        dW_l    = sum_b s1p[b,k] T_l[b,k,n] d[b,n],   d s1p[b,k] = sum_n T W d,   d demod[b,n] = sum_k T W s1p
    (the dependence of demod on W and s1p is differentiated by torch outside, where demod is computed)."""

    @staticmethod
    def forward(ctx, x, n_blocks, skip_from, rgb_from, *tensors):
        L = 2 * n_blocks
        W, S1, D = list(tensors[:L]), list(tensors[L:2 * L]), list(tensors[2 * L:3 * L])
        n_rgb = n_blocks - rgb_from
        RW = [None] * rgb_from + list(tensors[3 * L:3 * L + n_rgb])
        RB = [None] * rgb_from + list(tensors[3 * L + n_rgb:3 * L + 2 * n_rgb])
        rgb, acts, zsign = cips_forward_train(x, W, S1, D, RW, RB, n_blocks=n_blocks, skip_from=skip_from, rgb_from=rgb_from)
        ctx.cfg = (n_blocks, skip_from, rgb_from)
        ctx.save_for_backward(x, rgb, acts, zsign, *tensors)
        return rgb

    @staticmethod
    def backward(ctx, grad_out):
        n_blocks, skip_from, rgb_from = ctx.cfg
        L = 2 * n_blocks
        x, rgb, acts, zsign, *tensors = ctx.saved_tensors
        W, S1, D = list(tensors[:L]), list(tensors[L:2 * L]), list(tensors[2 * L:3 * L])
        n_rgb = n_blocks - rgb_from
        RW = [None] * rgb_from + list(tensors[3 * L:3 * L + n_rgb])
        g_pre = grad_out * (1 - rgb * rgb)                                   # through tanh
        scale = 1024.0 / g_pre.abs().max().clamp_min(1e-30)                  # device scalar: fp16 range management, no sync
        g_s = (g_pre * scale).contiguous()
        dz, dx = cips_backward_chain(x, acts, zsign, g_s, W, S1, D, RW, n_blocks=n_blocks, skip_from=skip_from, rgb_from=rgb_from,
                                     want_dx=ctx.needs_input_grad[0])
        inv = 1.0 / scale
        gW, gS, gD = [], [], []
        x16 = x.to(torch.float16)
        for l in range(L):
            X = x16 if l == 0 else acts[l - 1]
            T = _bmm_f32(X.transpose(1, 2), dz[l]) * inv                       # (B, in_l, out)
            gW.append(torch.einsum("bk,bkn,bn->kn", S1[l], T, D[l]))
            TW = T * W[l].unsqueeze(0)
            gS.append(torch.einsum("bkn,bn->bk", TW, D[l]))
            gD.append(torch.einsum("bkn,bk->bn", TW, S1[l]))
        g16 = g_s.to(torch.float16)
        gRW, gRB = [], []
        for b in range(rgb_from, n_blocks):
            gRW.append(_bmm_f32(g16.transpose(1, 2), acts[2 * b + 1]).sum(0) * inv)      # (3, 512)
            gRB.append(g_pre.sum((0, 1)))
        return (dx * inv if dx is not None else None, None, None, None, *gW, *gS, *gD, *gRW, *gRB)


def cips_style_prep(styles, mod_w, mod_b, weights, eps=1e-8):
    """c3d_cips_style_prep: per layer l style (B,style_dim), modulation weight (in_l,style_dim) / bias (in_l), W (in_l,512)
    -> lists s1p[l] (B,in_l), demod[l] (B,512) in ONE launch."""
    lib = load()
    L = len(weights)
    B = styles[0].shape[0]
    dev = styles[0].device
    a = _lib.StylePrep(n_layers=L, style_dim=styles[0].shape[1], eps=float(eps))
    keep, s1p, demod = [], [], []
    for l in range(L):
        ts = [_f32c(styles[l], "style"), _f32c(mod_w[l].detach(), "mod_w"), _f32c(mod_b[l].detach(), "mod_b"), _f32c(weights[l].detach(), "w")]
        keep += ts
        in_l = ts[3].shape[0]
        s1p.append(torch.empty((B, in_l), device=dev, dtype=torch.float32))
        demod.append(torch.empty((B, ts[3].shape[1]), device=dev, dtype=torch.float32))
        a.style[l], a.mod_w[l], a.mod_b[l], a.w[l] = (ptr(t) for t in ts)
        a.s1p[l], a.demod[l], a.in_dim[l] = ptr(s1p[l]), ptr(demod[l]), in_l
    check(lib.c3d_cips_style_prep(C.byref(a), B, stream_ptr()), "c3d_cips_style_prep")
    return s1p, demod


_U8_MODES = {"save_image": 0, "tensor_to_pil": 1, "to_pil": 2}


def image_to_u8(img, mode="save_image", value_range=(-1, 1)):
    """(B, C, H, W) or (C, H, W) fp32 generator output -> (B, H, W, C) / (H, W, C) uint8 on the device, bit-identical to
    what the reference's inference scripts hand to PIL (include/cips3d_b200.h, c3d_image_to_u8):
      'save_image'     torchvision save_image(img, path, normalize=True, value_range=value_range)  gen_images.py:64
      'tensor_to_pil'  st_web.py:44-46          'to_pil'  comm_utils.py:21-24
    The result is what `Image.fromarray(...)` takes after a `.cpu().numpy()` of 1 byte per sample."""
    lib = load()
    if mode not in _U8_MODES:
        raise ValueError(f"image_to_u8: unknown mode {mode!r} (one of {sorted(_U8_MODES)})")
    if img.dtype != torch.float32:
        raise _lib.C3dError(f"img: expected float32, got {img.dtype}")
    if img.dim() not in (3, 4):
        raise ValueError(f"image_to_u8: expected (B, C, H, W) or (C, H, W), got {tuple(img.shape)}")
    x4 = img if img.dim() == 4 else img[None]
    B, Cn, H, W = x4.shape
    # the generator returns an NCHW *view* of the CIPS kernel's (B, H*W, 3) output: convert it in place of a re-layout
    cl = x4.permute(0, 2, 3, 1).is_contiguous() and not x4.is_contiguous()
    src = x4.permute(0, 2, 3, 1) if cl else x4.contiguous()
    out = torch.empty((B, H, W, Cn), device=img.device, dtype=torch.uint8)
    lo, hi = value_range
    check(lib.c3d_image_to_u8(ptr(src), ptr(out), B, Cn, H, W, int(cl), _U8_MODES[mode], float(lo), float(hi),
                              stream_ptr()), "c3d_image_to_u8")
    return out if img.dim() == 4 else out[0]


def film_sin_supported(z, gain, bias):
    """Shapes the native FiLM-sine takes: z (B, P, C) fp32, gain / bias (B, C) or (B, 1, C), C % 4 == 0 and 256 % (C / 4) == 0."""
    if z.dim() != 3 or z.dtype != torch.float32 or gain.dtype != torch.float32 or bias.dtype != torch.float32:
        return False
    B, _, Cn = z.shape
    ok = lambda t: t.numel() == B * Cn and t.shape[0] == B and t.shape[-1] == Cn       # noqa: E731
    return Cn >= 4 and Cn % 4 == 0 and 256 % (Cn // 4) == 0 and ok(gain) and ok(bias)


# ---------------------------------------------------------------------------------------------------------------------
# Per-point linear layer of the NeRF training graph (csrc/points_linear_tc.cu)
# ---------------------------------------------------------------------------------------------------------------------
def points_linear_supported(x, weight):
    """x (..., K) fp32 with K in {32, 64, 128}; weight (N, K) with N in {32, 64, 128}"""
    return (x.dtype == torch.float32 and weight.dtype == torch.float32 and weight.dim() == 2 and x.shape[-1] == weight.shape[1]
            and weight.shape[0] in (32, 64, 128) and weight.shape[1] in (32, 64, 128) and x.numel() > 0)


def _points_linear_raw(x2, w, bias, scale, transposed):
    """x2 (rows, K) contiguous; w (N, K) [transposed False] or (K, N) [True]; -> (rows, N)"""
    lib = load()
    rows, K = x2.shape
    N = w.shape[1] if transposed else w.shape[0]
    y = torch.empty((rows, N), device=x2.device, dtype=torch.float32)
    nbytes = lib.c3d_points_linear_workspace_bytes(N, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x2.device)
    check(lib.c3d_points_linear(ptr(x2), ptr(w), ptr(bias) if bias is not None else None, ptr(scale) if scale is not None else None,
                                ptr(y), rows, K, N, int(bool(transposed)), ptr(ws), nbytes, stream_ptr()), "c3d_points_linear")
    return y


class PointsLinearFunction(Function):
    """F.linear(x, weight, bias) for the per-point layers of the NeRF field (film_layer.py:78-107, generator.py:236-243) with the
    forward GEMM and the data-gradient GEMM on the tcgen05 split-fp16 kernel (fp32-equivalent products); the weight gradient
    dW = dZ^T X (a points-long reduction) and db stay library reductions over the same tensors, as for the CIPS MLP.
    The gradient operand is pre-scaled to 1024 / max|dZ| on the device (no host sync) so that its fp16 halves stay normal."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        w = weight.contiguous()
        y = _points_linear_raw(x2, w, bias.contiguous() if bias is not None else None, None, False)
        ctx.save_for_backward(x2, w)
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            scale = (1024.0 / dy2.abs().amax().clamp_min(1e-30)).reshape(1).to(torch.float32)
            dx = _points_linear_raw(dy2, w, None, scale, True).view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            dw = dy2.t().mm(x2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy2.sum(0)
        return dx, dw, db


def points_linear(x, weight, bias=None):
    return PointsLinearFunction.apply(x, weight, bias)


class FilmSinFunction(Function):
    """y = sin(gain * z + bias) with per-image gain / bias (film_layer.py:94-107): one native pass forward, one backward
    (dz and the per-image reductions dgain, dbias together); saves z only.  No double backward (the generator's graph
    needs none: R1 differentiates the discriminator)."""

    @staticmethod
    def forward(ctx, z, gain, bias):
        lib = load()
        z = z.contiguous()
        B, P, Cn = z.shape
        g2, b2 = gain.reshape(B, Cn).contiguous(), bias.reshape(B, Cn).contiguous()
        y = torch.empty_like(z)
        check(lib.c3d_film_sin_fwd(ptr(z), ptr(g2), ptr(b2), ptr(y), B, P, Cn, stream_ptr()), "c3d_film_sin_fwd")
        ctx.save_for_backward(z, g2, b2)
        ctx.shapes = (gain.shape, bias.shape)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        lib = load()
        z, g2, b2 = ctx.saved_tensors
        B, P, Cn = z.shape
        dy = dy.contiguous()
        dz = torch.empty_like(z)
        dg, db = torch.empty_like(g2), torch.empty_like(b2)
        nbytes = lib.c3d_film_sin_bwd_workspace_bytes(B, P, Cn)
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=z.device)
        check(lib.c3d_film_sin_bwd(ptr(z), ptr(g2), ptr(b2), ptr(dy), ptr(dz), ptr(dg), ptr(db), B, P, Cn, ptr(ws), nbytes,
                                   stream_ptr()), "c3d_film_sin_bwd")
        return dz, dg.view(ctx.shapes[0]), db.view(ctx.shapes[1])


def film_sin(z, gain, bias):
    return FilmSinFunction.apply(z, gain, bias)


def integrate_supported(rgb_sigma, z, noise=None):
    """Shapes the native volume integration takes: rgb_sigma (..., T, C + 1) fp32 with T <= 32 sorted samples and C <= 128
    channels, z (..., T), noise None or (..., T) / (..., T, 1)."""
    if rgb_sigma.dim() < 2 or rgb_sigma.dtype != torch.float32 or z.dtype != torch.float32:
        return False
    T, C1 = rgb_sigma.shape[-2:]
    if not (1 <= T <= 32 and 2 <= C1 <= 129) or z.shape != rgb_sigma.shape[:-1]:
        return False
    return noise is None or (noise.dtype == torch.float32 and noise.numel() == z.numel())


def _torch_weights_grad(rs, z, noise, clamp, last_back, d_weights):
    """d(loss)/d(rgb_sigma) through the compositing WEIGHTS (pigan_utils.py:241-262), for the rare caller that differentiates the
    returned weights / depth (e.g. a depth loss in piGAN_lib/inverse_render-style use).  The training loop never does -- weights feed
    sample_pdf under no_grad -- so this is a torch-op recompute from the saved inputs, not a kernel."""
    with torch.enable_grad():
        rs_ = rs.detach().requires_grad_(True)
        sig = rs_[..., -1]
        deltas = torch.cat([z[..., 1:] - z[..., :-1], torch.full_like(z[..., :1], 1e10)], -1)
        if noise is not None:
            sig = sig + noise
        act = torch.nn.functional.softplus(sig) if clamp == CLAMP_MODES["softplus"] else torch.relu(sig)
        alphas = 1 - torch.exp(-deltas * act)
        shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-10], -1)
        w = alphas * torch.cumprod(shifted, -1)[..., :-1]
        if last_back:
            w = torch.cat([w[..., :-1], w[..., -1:] + (1 - w.sum(-1))[..., None]], -1)
        (g,) = torch.autograd.grad(w, rs_, d_weights)
    return g


class IntegrateFunction(Function):
    """fancy_integration (pigan_utils.py:212-273) on sorted samples as one native pass forward and one backward
    (csrc/integrate_ops.cu).  Returns (fea, weights); gradients flow to rgb_sigma -- z and noise are constants of the
    reference's graph too (fine depths are drawn under no_grad, generator_nerf_inr.py:537).  Saves only its inputs.
    A gradient arriving through `weights` (depth / weight losses) is honoured by a torch-op recompute (_torch_weights_grad).
    No double backward (the generator's graph needs none)."""

    @staticmethod
    def forward(ctx, rgb_sigma, z, noise, clamp_mode, last_back, white_back):
        lib = load()
        if clamp_mode not in CLAMP_MODES:
            raise AssertionError("Need to choose clamp mode")             # pigan_utils.py:252-253
        lead, (T, C1) = rgb_sigma.shape[:-2], rgb_sigma.shape[-2:]
        rs = rgb_sigma.contiguous()
        zc = z.contiguous()
        nz = None if noise is None else noise.reshape(z.shape).contiguous()
        rays = zc.numel() // T
        fea = torch.empty(*lead, C1 - 1, dtype=torch.float32, device=rs.device)
        weights = torch.empty(*lead, T, dtype=torch.float32, device=rs.device)
        check(lib.c3d_integrate_fwd(ptr(rs), ptr(zc), ptr(nz) if nz is not None else None, ptr(fea), ptr(weights), rays, T,
                                    C1 - 1, CLAMP_MODES[clamp_mode], int(bool(last_back)), int(bool(white_back)), stream_ptr()),
              "c3d_integrate_fwd")
        ctx.save_for_backward(rs, zc, nz if nz is not None else torch.empty(0, device=rs.device))
        ctx.cfg = (nz is not None, CLAMP_MODES[clamp_mode], int(bool(last_back)), int(bool(white_back)))
        ctx.set_materialize_grads(False)          # an unused output arrives as None in backward, not as a zero tensor
        return fea, weights

    @staticmethod
    @once_differentiable
    def backward(ctx, d_fea, d_weights):
        lib = load()
        rs, zc, nz = ctx.saved_tensors
        has_noise, clamp, last_back, white_back = ctx.cfg
        T, C1 = rs.shape[-2:]
        d_rs = None
        if d_fea is not None:
            d_fea = d_fea.contiguous()
            d_rs = torch.empty_like(rs)
            check(lib.c3d_integrate_bwd(ptr(rs), ptr(zc), ptr(nz) if has_noise else None, ptr(d_fea), ptr(d_rs), zc.numel() // T, T,
                                        C1 - 1, clamp, last_back, white_back, stream_ptr()), "c3d_integrate_bwd")
        if d_weights is not None:
            g = _torch_weights_grad(rs, zc, nz if has_noise else None, clamp, last_back, d_weights)
            d_rs = g if d_rs is None else d_rs + g
        return d_rs, None, None, None, None, None


def integrate(rgb_sigma, z, noise=None, clamp_mode="relu", last_back=False, white_back=False):
    """(fea, weights) of fancy_integration; differentiable in rgb_sigma."""
    return IntegrateFunction.apply(rgb_sigma, z, noise, clamp_mode, last_back, white_back)


def integrate_merged_supported(fine, z_fine, coarse, z_coarse, noise=None):
    """Both halves (..., S, C + 1) fp32 with S <= 16 and C <= 128, depths (..., S); noise None or 2S values per ray."""
    if fine.shape != coarse.shape or z_fine.shape != z_coarse.shape or fine.dim() < 2:
        return False
    if any(t.dtype != torch.float32 for t in (fine, coarse, z_fine, z_coarse)):
        return False
    S, C1 = fine.shape[-2:]
    if not (1 <= S <= 16 and 2 <= C1 <= 129) or z_fine.shape != fine.shape[:-1]:
        return False
    return noise is None or (noise.dtype == torch.float32 and noise.numel() == 2 * z_fine.numel())


class IntegrateMergedFunction(Function):
    """torch.cat([fine, coarse]) + torch.sort of the depths + torch.gather + fancy_integration (generator.py:1733-1752) as one
    native pass forward and one backward: the 2S samples of a ray are rank-sorted in registers, colour rows are read from and
    gradients written to their source rows (csrc/integrate_ops.cu, merged form).  Returns (fea, weights, z_sorted); gradients
    flow to fine and coarse only.  Equal depths keep cat order (a stable sort)."""

    @staticmethod
    def forward(ctx, fine, z_fine, coarse, z_coarse, noise, clamp_mode, last_back, white_back):
        lib = load()
        if clamp_mode not in CLAMP_MODES:
            raise AssertionError("Need to choose clamp mode")             # pigan_utils.py:252-253
        lead, (S, C1) = fine.shape[:-2], fine.shape[-2:]
        f, c, zf, zc = fine.contiguous(), coarse.contiguous(), z_fine.contiguous(), z_coarse.contiguous()
        nz = None if noise is None else noise.reshape(*lead, 2 * S).contiguous()
        rays = zf.numel() // S
        fea = torch.empty(*lead, C1 - 1, dtype=torch.float32, device=f.device)
        weights = torch.empty(*lead, 2 * S, dtype=torch.float32, device=f.device)
        z_sorted = torch.empty(*lead, 2 * S, dtype=torch.float32, device=f.device)
        cfg = (CLAMP_MODES[clamp_mode], int(bool(last_back)), int(bool(white_back)))
        check(lib.c3d_integrate_merge_fwd(ptr(f), ptr(zf), ptr(c), ptr(zc), ptr(nz), ptr(fea), ptr(weights), ptr(z_sorted), rays, S,
                                          C1 - 1, *cfg, stream_ptr()), "c3d_integrate_merge_fwd")
        ctx.save_for_backward(f, zf, c, zc, nz if nz is not None else torch.empty(0, device=f.device))
        ctx.cfg = (nz is not None,) + cfg
        ctx.mark_non_differentiable(weights, z_sorted)
        return fea, weights, z_sorted

    @staticmethod
    @once_differentiable
    def backward(ctx, d_fea, _d_weights, _d_z):
        lib = load()
        f, zf, c, zc, nz = ctx.saved_tensors
        has_noise, clamp, last_back, white_back = ctx.cfg
        S, C1 = f.shape[-2:]
        d_fea = d_fea.contiguous()
        d_f, d_c = torch.empty_like(f), torch.empty_like(c)
        check(lib.c3d_integrate_merge_bwd(ptr(f), ptr(zf), ptr(c), ptr(zc), ptr(nz) if has_noise else None, ptr(d_fea), ptr(d_f),
                                          ptr(d_c), zf.numel() // S, S, C1 - 1, clamp, last_back, white_back, stream_ptr()),
              "c3d_integrate_merge_bwd")
        return d_f, None, d_c, None, None, None, None, None


def integrate_merged(fine, z_fine, coarse, z_coarse, noise=None, clamp_mode="relu", last_back=False, white_back=False):
    """(fea, weights, z_sorted) of the sorted union of the fine and coarse samples; differentiable in fine and coarse."""
    return IntegrateMergedFunction.apply(fine, z_fine, coarse, z_coarse, noise, clamp_mode, last_back, white_back)


def fancy_integration(rgb_sigma, z_vals, device=None, dim_rgb=3, noise_std=0.5, last_back=False, white_back=False,
                      clamp_mode=None, fill_mode=None):
    """Drop-in for exp/pigan/pigan_utils.py:212-273 (same argument meaning, return shapes, RNG draw and error behaviour):
    rgb_sigma (b, rays, samples, dim_rgb + 1), z_vals (b, rays, samples, 1) sorted ->
    (rgb_final (b, rays, dim_rgb), depth_final (b, rays, 1), weights (b, rays, samples, 1)), on the native op.
    The noise is drawn exactly as the reference draws it (one torch.randn of sigma's shape, scaled by noise_std).  Gradients
    flow to rgb_sigma through rgb_final; weights and depth_final are outputs without a graph (the reference's training loop
    never differentiates them: the weights feed sample_pdf under no_grad, generator_nerf_inr.py:537, depth is inference-only).
    fill_mode 'debug' / 'weight' (reference debugging aids) are not on the native path."""
    if clamp_mode not in CLAMP_MODES:
        assert 0, "Need to choose clamp mode"                                # pigan_utils.py:252-253
    if fill_mode is not None:
        raise _lib.C3dError("fancy_integration: fill_mode is a debugging aid of the reference and is not supported natively")
    if rgb_sigma.shape[-1] != dim_rgb + 1:
        raise _lib.C3dError(f"fancy_integration: expected dim_rgb + 1 = {dim_rgb + 1} channels, got {rgb_sigma.shape[-1]}")
    z = z_vals.reshape(rgb_sigma.shape[:-1])
    noise = torch.randn(rgb_sigma.shape[:-1] + (1,), device=rgb_sigma.device if device is None else device) * noise_std
    if not integrate_supported(rgb_sigma, z, noise):
        raise _lib.C3dError(f"fancy_integration: unsupported shape {tuple(rgb_sigma.shape)} (samples <= 32, dim_rgb <= 128, fp32)")
    rgb_final, weights = integrate(rgb_sigma, z, noise, clamp_mode, last_back, white_back)
    depth_final = torch.sum(weights * z, -1, keepdim=True)
    return rgb_final, depth_final, weights.unsqueeze(-1)


def sample_pdf_supported(bins, weights):
    return (weights.dim() == 2 and bins.dim() == 2 and weights.dtype == torch.float32 and bins.dtype == torch.float32
            and 1 <= weights.shape[1] <= 32 and bins.shape == (weights.shape[0], weights.shape[1] + 1))


def sample_pdf_from_u(bins, weights, u, eps=1e-5):
    """The arithmetic of pigan_utils.sample_pdf (L181-209) for given uniforms u (N_rays, N_importance): one native launch."""
    lib = load()
    if not sample_pdf_supported(bins, weights) or u.dim() != 2 or u.shape[0] != weights.shape[0] or u.dtype != torch.float32:
        raise _lib.C3dError(f"sample_pdf: unsupported shapes bins {tuple(bins.shape)} weights {tuple(weights.shape)} u {tuple(u.shape)} "
                            "(fp32, 1..32 weights per ray, bins one longer)")
    b, w, uu = bins.detach().contiguous(), weights.detach().contiguous(), u.detach().contiguous()
    out = torch.empty_like(uu)
    check(lib.c3d_sample_pdf(ptr(b), ptr(w), ptr(uu), ptr(out), w.shape[0], w.shape[1], uu.shape[1], float(eps), stream_ptr()),
          "c3d_sample_pdf")
    return out


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5):
    """Drop-in for exp/pigan/pigan_utils.py:164-209: same arguments, the same torch.rand / torch.linspace draw, same output
    (N_rays, N_importance).  No gradient (the reference calls it under no_grad and detaches the result)."""
    N_rays = weights.shape[0]
    if det:
        u = torch.linspace(0, 1, N_importance, device=bins.device).expand(N_rays, N_importance)       # pigan_utils.py:188-190
    else:
        u = torch.rand(N_rays, N_importance, device=bins.device)                                       # pigan_utils.py:192
    return sample_pdf_from_u(bins, weights, u.contiguous(), eps)
