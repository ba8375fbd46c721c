"""Discriminator class surface of CIPS-3D on the B200-native D ops.

Mirrors exp/cips3d/models/discriminator.py (EqualConv2d L20-54, Blur L67-82, ConvLayer L134-222,
ResBlock L224-252, EqualLinear L254-288, Discriminator_MultiScale L405-585,
Discriminator_MultiScale_Aux L588-664) and exp/cips3d/models/diffaug.py L10-85.  The two native
ops of the reference (fused bias+leaky-ReLU, upfirdn2d blur) are this package's CUDA kernels
(ops.fused_leaky_relu / ops.upfirdn2d, differentiable twice for the R1 penalty); the convolutions
stay cuDNN library calls exactly as in the reference.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .generator import MODEL_REGISTRY, _require_cuda


_SQRT2 = math.sqrt(2.0)
_BLUR_TAPS = (1, 3, 3, 1)


class FusedLeakyReLU(nn.Module):
    """Learned per-channel bias + leaky-ReLU * sqrt(2) in one kernel (exp/comm/op/fused_act.py:73-82)."""

    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope, self.scale = negative_slope, scale

    def forward(self, input):
        return ops.fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


class EqualConv2d(nn.Module):
    """N(0,1) weights with the He constant applied at run time (discriminator.py:20-54).  The product is the library's implicit
    GEMM (as in the reference); the autograd structure is ops.conv2d's: fprop / dgrad / wgrad Functions closed under
    differentiation, so the R1 double backward never takes torch's generic slow path."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None
        self.scale = (in_channel * kernel_size * kernel_size) ** -0.5
        self.stride, self.padding = stride, padding

    def forward(self, input):
        return ops.conv2d(input, self.weight * self.scale, self.bias, self.stride, self.padding)

    def extra_repr(self):
        o, i, k, _ = self.weight.shape
        return f'{i}, {o}, {k}, stride={self.stride}, padding={self.padding}'


def make_kernel(taps):
    """Separable FIR taps -> normalised 2-D kernel (discriminator.py:57-65)."""
    k = torch.as_tensor(taps, dtype=torch.float32)
    k = torch.outer(k, k) if k.ndim == 1 else k
    return k / k.sum()


class Blur(nn.Module):
    """upfirdn2d with up = down = 1 (discriminator.py:67-82); `kernel` is a state_dict buffer."""

    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        k = make_kernel(kernel)
        self.register_buffer('kernel', k * (upsample_factor ** 2) if upsample_factor > 1 else k)
        self.pad = pad

    def forward(self, input):
        return ops.upfirdn2d(input, self.kernel, pad=self.pad)


class ScaledLeakyReLU(nn.Module):
    def __init__(self, negative_slope=0.2):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, input):
        return F.leaky_relu(input, self.negative_slope) * _SQRT2


class ConvLayer(nn.Module):
    """[blur ->] conv [-> bias+lrelu] with the child names of discriminator.py:134-222
    (down_blur / equal_conv / flrelu / slrelu) so checkpoints load unchanged.  A stride-2 layer
    blurs first with pad ((p+1)//2, p//2), p = len(taps) - 2 + k - 1."""

    def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=_BLUR_TAPS, bias=True,
                 activate=True, upsample=False, padding="zero"):
        super().__init__()
        if upsample:
            raise NotImplementedError("transposed-conv upsampling is not used by the CIPS-3D discriminators")
        if padding not in ("zero", "valid"):
            raise ValueError('Padding should be "zero" or "valid"')
        if downsample:
            p = len(blur_kernel) - 2 + kernel_size - 1
            self.down_blur = Blur(blur_kernel, pad=((p + 1) // 2, p // 2))
        self.padding = (kernel_size - 1) // 2 if (padding == "zero" and not downsample) else 0
        self.equal_conv = EqualConv2d(in_channel, out_channel, kernel_size, stride=2 if downsample else 1,
                                      padding=self.padding, bias=bias and not activate)
        if activate and bias:
            self.flrelu = FusedLeakyReLU(out_channel)
        elif activate:
            self.slrelu = ScaledLeakyReLU(0.2)

    def forward(self, x):
        for name in ("down_blur", "equal_conv", "flrelu", "slrelu"):
            mod = self._modules.get(name)
            if mod is not None:
                x = mod(x)
        return x


class ResBlock(nn.Module):
    """(conv2(conv1(x)) + skip(x)) / sqrt(2), one of the two 3x3 convs strided (discriminator.py:224-252)."""

    def __init__(self, in_channel, out_channel, blur_kernel=_BLUR_TAPS, kernel_size=3, downsample=True,
                 first_downsample=False):
        super().__init__()
        down1, down2 = (downsample, False) if first_downsample else (False, downsample)
        mid = in_channel
        self.conv1 = ConvLayer(in_channel, mid, kernel_size, downsample=down1, blur_kernel=blur_kernel)
        self.conv2 = ConvLayer(mid, out_channel, kernel_size, downsample=down2, blur_kernel=blur_kernel)
        self.skip = ConvLayer(in_channel, out_channel, 1, downsample=downsample, blur_kernel=blur_kernel,
                              activate=False, bias=False)

    def forward(self, input):
        return (self.conv2(self.conv1(input)) + self.skip(input)) / _SQRT2


class EqualLinear(nn.Module):
    """Equalised-lr linear, optional fused bias+lrelu (discriminator.py:254-288)."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.full((out_dim,), float(bias_init))) if bias else None
        self.activation, self.lr_mul = activation, lr_mul
        self.scale = lr_mul / math.sqrt(in_dim)

    def forward(self, input):
        w = self.weight * self.scale
        if self.activation:
            return ops.fused_leaky_relu(F.linear(input, w), self.bias * self.lr_mul)
        return F.linear(input, w, self.bias * self.lr_mul)

    def extra_repr(self):
        return f'{self.weight.shape[1]}, {self.weight.shape[0]}'


# --------------------------------------------------------------------------------------
# DiffAugment (semantics of exp/cips3d/models/diffaug.py:10-85; D6 in SURVEY.md: stays torch).
# Re-derived with mask/gather arithmetic instead of index grids; the torch RNG calls (shape,
# order, distribution) are the reference's so augmentations match draw for draw.
# --------------------------------------------------------------------------------------
def _per_sample_uniform(x):
    return torch.rand(x.size(0), 1, 1, 1, dtype=x.dtype, device=x.device)


def _color_jitter(x):
    """brightness (+U-0.5), saturation (x2U about the channel mean), contrast ((U+0.5) about the
    image mean) -- diffaug.py:31-46, in that order."""
    x = x + (_per_sample_uniform(x) - 0.5)
    m = x.mean(dim=1, keepdim=True)
    x = (x - m) * (_per_sample_uniform(x) * 2) + m
    m = x.mean(dim=[1, 2, 3], keepdim=True)
    return (x - m) * (_per_sample_uniform(x) + 0.5) + m


def _translate(x, ratio=0.125):
    """out[b,:,i,j] = x[b,:,i+tx_b,j+ty_b] (zero outside), tx/ty ~ randint(+-ratio*size) -- diffaug.py:49-62."""
    B, _, H, W = x.shape
    sx, sy = int(H * ratio + 0.5), int(W * ratio + 0.5)
    tx = torch.randint(-sx, sx + 1, size=[B, 1, 1], device=x.device)
    ty = torch.randint(-sy, sy + 1, size=[B, 1, 1], device=x.device)
    src_i = torch.arange(H, device=x.device).view(1, H, 1) + tx          # (B,H,1)
    src_j = torch.arange(W, device=x.device).view(1, 1, W) + ty          # (B,1,W)
    ok = ((src_i >= 0) & (src_i < H) & (src_j >= 0) & (src_j < W)).unsqueeze(1)
    flat = (src_i.clamp(0, H - 1) * W + src_j.clamp(0, W - 1)).view(B, 1, H * W).expand(-1, x.size(1), -1)
    return x.reshape(B, x.size(1), H * W).gather(2, flat).view_as(x) * ok.to(x.dtype)


def _cutout(x, ratio=0.2):
    """Zero a (ratio*size)^2 window around a random centre; rows/cols are clamped to the image the
    way the reference's index grid is (so a window hanging off the edge still blanks the border
    row/column) -- diffaug.py:65-79."""
    B, _, H, W = x.shape
    ch, cw = int(H * ratio + 0.5), int(W * ratio + 0.5)
    ox = torch.randint(0, H + (1 - ch % 2), size=[B, 1, 1], device=x.device)
    oy = torch.randint(0, W + (1 - cw % 2), size=[B, 1, 1], device=x.device)
    r = torch.arange(H, device=x.device).view(1, H, 1)
    c = torch.arange(W, device=x.device).view(1, 1, W)
    r_lo, r_hi = (ox - ch // 2).clamp(0, H - 1), (ox - ch // 2 + ch - 1).clamp(0, H - 1)
    c_lo, c_hi = (oy - cw // 2).clamp(0, W - 1), (oy - cw // 2 + cw - 1).clamp(0, W - 1)
    hole = (r >= r_lo) & (r <= r_hi) & (c >= c_lo) & (c <= c_hi)
    if ch == 0 or cw == 0:
        hole = torch.zeros_like(hole)
    return x * (~hole).unsqueeze(1).to(x.dtype)


_POLICIES = {'color': _color_jitter, 'translation': _translate, 'cutout': _cutout}


def DiffAugment(x, policy='color,translation,cutout', channels_first=True):
    if not policy:
        return x
    if not channels_first:
        x = x.permute(0, 3, 1, 2)
    for name in policy.split(','):
        x = _POLICIES[name](x)
    if not channels_first:
        x = x.permute(0, 2, 3, 1)
    return x.contiguous()


def _default_channels(channel_multiplier):
    """discriminator.py:441-451: 512 channels up to 32x32, then halving per octave."""
    ch = {r: 512 for r in (4, 8, 16, 32)}
    for r, base in ((64, 256), (128, 128), (256, 64), (512, 32), (1024, 16)):
        ch[r] = base * channel_multiplier
    return ch


@MODEL_REGISTRY.register(name_prefix=__name__)
class Discriminator_MultiScale(nn.Module):
    """Progressive multi-resolution StyleGAN2 D (discriminator.py:405-585): one 1x1 stem per
    resolution (conv_in[res]), ResBlocks from max_size down to 8 (convs[res]), 3x3 conv at 4x4,
    two equalised linears.  The input resolution picks the entry point; alpha < 1 fades in the
    newest resolution against a bilinearly halved input."""

    def __init__(self, diffaug, max_size, channel_multiplier=2, blur_kernel=_BLUR_TAPS, input_size=3,
                 first_downsample=False, channels=None, stddev_group=4, **kwargs):
        super().__init__()
        self.epoch = self.step = 0
        self.diffaug, self.max_size, self.input_size = diffaug, max_size, input_size
        self.stddev_group, self.stddev_feat = stddev_group, 1
        channels = _default_channels(channel_multiplier) if channels is None else channels
        self.conv_in = nn.ModuleDict({str(res): ConvLayer(input_size, c, 1) for res, c in channels.items()})
        self.convs = nn.ModuleDict()
        res, c_in = max_size, channels[max_size]
        while res > 4:
            c_out = channels[res // 2]
            self.convs[str(res)] = ResBlock(c_in, c_out, blur_kernel, first_downsample=first_downsample)
            res, c_in = res // 2, c_out
        self.final_conv = ConvLayer(c_in + (1 if stddev_group > 1 else 0), channels[4], 3)
        self.space_linear = EqualLinear(channels[4] * 16, channels[4], activation='fused_lrelu')
        self.out_linear = EqualLinear(channels[4], 1)
        self.module_name_list = ['conv_in', 'convs', 'final_conv', 'space_linear', 'out_linear']

    def diff_aug_img(self, img):
        return DiffAugment(img, policy='color,translation,cutout')

    def _minibatch_stddev(self, out):
        """discriminator.py:545-556"""
        n, c, h, w = out.shape
        g = min(n, self.stddev_group)
        y = out.contiguous().view(g, -1, self.stddev_feat, c // self.stddev_feat, h, w)
        y = torch.sqrt(y.var(0, unbiased=False) + 1e-8).mean([2, 3, 4], keepdims=True).squeeze(2)
        return torch.cat([out, y.repeat(g, 1, h, w)], 1)

    # Internal activation layout.  True: channels-last (N, H, W, C in memory) from the stem to the 4x4 head -- the layout the
    # tensor-core convolutions use natively; bias_act, the blur and conv2d keep whatever layout they are given (ops.bias_act,
    # c3d_blur_nhwc).  Measured NEUTRAL on B200 (profiles/r02g_summary.md: cuDNN's own NCHW <-> NHWC transposes, 15 % of a config-5
    # step, disappear and its kernels get 24 ms faster per 2 steps, but the same time reappears as layout-conversion copies around
    # the autograd graph and a slower interleaved blur), so the default stays the NCHW layout of the reference.  The module's contract
    # is the same either way: NCHW-shaped input, (N, 1) output.
    channels_last = False

    def forward(self, input, alpha, summary_ddict=None):
        _require_cuda(input, "Discriminator_MultiScale.forward")
        x = self.diff_aug_img(input) if self.diffaug else input
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        res = 2 ** int(math.log(x.shape[-1], 2))
        out = self.convs[str(res)](self.conv_in[str(res)](x))
        if alpha < 1:
            low = self.conv_in[str(res // 2)](F.interpolate(x, scale_factor=0.5, mode='bilinear'))
            out = alpha * out + (1 - alpha) * low
        res //= 2
        while res > 4:
            out = self.convs[str(res)](out)
            res //= 2
        if self.stddev_group > 0:
            out = self._minibatch_stddev(out)
        out = self.space_linear(self.final_conv(out).flatten(1))
        if summary_ddict is not None:
            with torch.no_grad():
                summary_ddict['logits_norm']['logits_norm'] = out.norm(dim=1).mean().item()
                summary_ddict['w_norm']['w_norm'] = self.out_linear.weight.norm(dim=1).mean().item()
        return self.out_linear(out), None, None


@MODEL_REGISTRY.register(name_prefix=__name__)
class Discriminator_MultiScale_Aux(nn.Module):
    """Main D + a narrower auxiliary D for the NeRF-branch images (discriminator.py:588-664): with
    use_aux_disc the first half of the batch goes to main_disc, the second half to aux_disc."""

    def __init__(self, diffaug, max_size, channel_multiplier=2, first_downsample=False, stddev_group=0, **kwargs):
        super().__init__()
        self.epoch = self.step = 0
        self.main_disc = Discriminator_MultiScale(diffaug=diffaug, max_size=max_size,
                                                  channel_multiplier=channel_multiplier,
                                                  first_downsample=first_downsample, stddev_group=stddev_group)
        aux_channels = {r: 256 for r in (4, 8, 16, 32, 64, 128)}
        aux_channels.update({256: 128, 512: 64, 1024: 32})
        self.aux_disc = Discriminator_MultiScale(diffaug=diffaug, max_size=max_size, channel_multiplier=2,
                                                 first_downsample=True, channels=aux_channels,
                                                 stddev_group=stddev_group)

    def forward(self, input, use_aux_disc=False, summary_ddict=None, alpha=1., **kwargs):
        if not use_aux_disc:
            return self.main_disc(input, alpha, summary_ddict=summary_ddict)
        half = input.shape[0] // 2
        main_out, latent, position = self.main_disc(input[:half], alpha, summary_ddict=summary_ddict)
        aux_out = self.aux_disc(input[half:], alpha)[0]
        return torch.cat([main_out, aux_out], dim=0), latent, position
