"""Generator class surface of CIPS-3D on the B200-native hot path.

Mirrors (constructor kwargs, forward signatures, attribute names, state_dict keys/shapes and
RNG call order) exp/cips3d/models/generator.py of the reference:
  NeRFNetwork L151-376, SinBlock L893-980, ToRGB L983-1006, CIPSNet L1009-1154,
  GeneratorNerfINR L1158-1951, GeneratorNerfINR_freeze_NeRF L1954-2078,
plus exp/comm/models/film_layer.py FiLMLayer L41-107, exp/comm/models/mod_conv_fc.py
SinStyleMod L392-563 and exp/cips3d/models/multi_head_mapping.py L13-153.

Inference (torch.no_grad / no parameter requires grad) runs the fused CUDA kernels:
c3d_ray_siren_fwd for rays->features and c3d_cips_fwd for the per-pixel MLP.  When autograd
needs a graph (training step) the same math runs as differentiable torch CUDA ops -- the
backward kernels are the next hot-path row, not a CPU path:
every entry point refuses non-CUDA tensors.
"""
import contextlib
import logging
import math
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import comm_utils, ops

_log = logging.getLogger("cips3d_b200")
_warned = set()


def _note_torch_path(what, why):
    """The fused sm_100a kernels cover the reference's shipped configurations; any other module shape runs as torch CUDA
    ops.  That must not happen silently: say so once per (module, reason)."""
    if (what, why) not in _warned:
        _warned.add((what, why))
        _log.warning("cips3d_b200: %s runs as eager torch CUDA ops, not the fused kernel: %s", what, why)

try:  # the reference's registry, when the caller runs inside the reference's train.py
    from tl2.proj.fvcore import MODEL_REGISTRY  # type: ignore
except Exception:  # pragma: no cover - tl2 is not installed in this image
    class _Registry(dict):
        def register(self, name_prefix=None, **kw):
            def deco(cls):
                self[f"{name_prefix}.{cls.__name__}" if name_prefix else cls.__name__] = cls
                return cls
            return deco
    MODEL_REGISTRY = _Registry()


def kaiming_leaky_init(m):
    # tl2.proj.pytorch.init_func.kaiming_leaky_init; in-tree twin: multi_head_mapping.py:22-25
    if m.__class__.__name__.find('Linear') != -1:
        torch.nn.init.kaiming_normal_(m.weight, a=0.2, mode='fan_in', nonlinearity='leaky_relu')


def frequency_init(freq):                       # nerf_network.py:29-36, inr_network.py:20-27
    def init(m):
        with torch.no_grad():
            if isinstance(m, nn.Linear):
                num_input = m.weight.size(-1)
                m.weight.uniform_(-np.sqrt(6 / num_input) / freq, np.sqrt(6 / num_input) / freq)
    return init


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: cips3d_b200 runs on CUDA (sm_100a) only; got a {t.device} tensor. "
                           "There is no CPU fallback.")


# ======================================================================================
# layers
# ======================================================================================
class LinearScale(nn.Module):                   # film_layer.py:24-40
    def __init__(self, scale, bias):
        super().__init__()
        self.scale_v = scale
        self.bias_v = bias

    def forward(self, x):
        return x * self.scale_v + self.bias_v


class FiLMLayer(nn.Module):                     # film_layer.py:41-107
    def __init__(self, in_dim, out_dim, style_dim, use_style_fc=True, which_linear=nn.Linear, **kwargs):
        super().__init__()
        self.in_dim, self.out_dim, self.style_dim, self.use_style_fc = in_dim, out_dim, style_dim, use_style_fc
        self.linear = which_linear(in_dim, out_dim)
        self.linear.apply(frequency_init(25))
        self.gain_scale = LinearScale(scale=15, bias=30)
        # FiLM + sine of the autograd graph as the native op (ops.FilmSinFunction), and its linear as the tcgen05 split-fp16
        # GEMM (ops.PointsLinearFunction: forward + data gradient native, weight gradient a library reduction)
        self.fused_film = False
        self.fused_linear = False
        if use_style_fc:
            self.gain_fc = which_linear(style_dim, out_dim)
            self.bias_fc = which_linear(style_dim, out_dim)
            self.gain_fc.weight.data.mul_(0.25)
            self.bias_fc.weight.data.mul_(0.25)
        else:
            self.style_dim = out_dim * 2

    def film_params(self, style):
        """gamma (B,out), beta (B,out)"""
        if self.use_style_fc:
            return self.gain_scale(self.gain_fc(style)), self.bias_fc(style)
        gain, bias = style.view(style.shape[0], 2, -1).unbind(dim=1)
        return self.gain_scale(gain), bias

    def forward(self, x, style):
        gain, bias = self.film_params(style)
        if x.dim() == 3:
            gain, bias = gain.unsqueeze(1), bias.unsqueeze(1)
        elif x.dim() != 2:
            assert 0
        if self.fused_film and x.dim() == 3:
            if self.fused_linear and ops.points_linear_supported(x, self.linear.weight):
                z = ops.points_linear(x, self.linear.weight, self.linear.bias)
            else:
                z = self.linear(x)
            if ops.film_sin_supported(z, gain, bias):
                return ops.film_sin(z, gain, bias)      # one native pass forward, one backward (csrc/film_ops.cu)
            return torch.sin(gain * z + bias)
        return torch.sin(gain * self.linear(x) + bias)


class UniformBoxWarp(nn.Module):                # nerf_network.py:39-45
    def __init__(self, sidelength):
        super().__init__()
        self.scale_factor = 2 / sidelength

    def forward(self, coordinates):
        return coordinates * self.scale_factor


class NeRFNetwork(nn.Module):                   # generator.py:151-376
    def __init__(self, in_dim=3, hidden_dim=256, hidden_layers=2, style_dim=512, rgb_dim=3, device=None,
                 name_prefix='nerf', **kwargs):
        super().__init__()
        self.device, self.in_dim, self.hidden_dim, self.rgb_dim = device, in_dim, hidden_dim, rgb_dim
        self.style_dim, self.hidden_layers, self.name_prefix = style_dim, hidden_layers, name_prefix
        self.module_name_list = []
        self.style_dim_dict = {}
        self.network = nn.ModuleList()
        self.module_name_list.append('network')
        _out_dim = in_dim
        for idx in range(hidden_layers):
            _in_dim, _out_dim = _out_dim, hidden_dim
            _layer = FiLMLayer(in_dim=_in_dim, out_dim=_out_dim, style_dim=style_dim, use_style_fc=True)
            self.network.append(_layer)
            self.style_dim_dict[f'{name_prefix}_w{idx}'] = _layer.style_dim
        self.final_layer = nn.Linear(hidden_dim, 1)
        self.module_name_list.append('final_layer')
        self.color_layer_sine = FiLMLayer(in_dim=hidden_dim, out_dim=hidden_dim // 2, style_dim=style_dim,
                                          use_style_fc=True)
        self.style_dim_dict[f'{name_prefix}_rgb'] = self.color_layer_sine.style_dim
        self.module_name_list.append('color_layer_sine')
        self.color_layer_linear = nn.Sequential(nn.Linear(hidden_dim // 2, rgb_dim))
        self.color_layer_linear.apply(kaiming_leaky_init)
        self.module_name_list.append('color_layer_linear')
        self.dim_styles = sum(self.style_dim_dict.values())
        self.gridwarper = UniformBoxWarp(0.24)
        self.fused_linear = False      # training graph: color_layer_linear on ops.points_linear (see FiLMLayer.fused_linear)

    # ---- what the fused kernel consumes
    def fused_supported(self, num_steps=None):
        """mirrors the argument checks of c3d_ray_siren_fwd (include/cips3d_b200.h): ffhq_exp.yaml's NeRF, 3..32 steps"""
        return (self.in_dim == 3 and self.hidden_dim == 128 and self.hidden_layers == 2 and self.rgb_dim == 32
                and (num_steps is None or 3 <= num_steps <= 32))

    def kernel_weights(self):
        n0, n1 = self.network[0].linear, self.network[1].linear
        return dict(w0=n0.weight, b0=n0.bias, w1=n1.weight, b1=n1.bias,
                    w_sigma=self.final_layer.weight, b_sigma=self.final_layer.bias,
                    wc=self.color_layer_sine.linear.weight, bc=self.color_layer_sine.linear.bias,
                    wl=self.color_layer_linear[0].weight, bl=self.color_layer_linear[0].bias)

    def kernel_film(self, style_dict):
        g0, b0 = self.network[0].film_params(style_dict[f'{self.name_prefix}_w0'])
        g1, b1 = self.network[1].film_params(style_dict[f'{self.name_prefix}_w1'])
        gc, bc = self.color_layer_sine.film_params(style_dict[f'{self.name_prefix}_rgb'])
        return dict(gamma0=g0, beta0=b0, gamma1=g1, beta1=b1, gammac=gc, betac=bc)

    def forward_with_frequencies_phase_shifts(self, input, style_dict, **kwargs):
        x = self.gridwarper(input)
        for index, layer in enumerate(self.network):
            x = layer(x, style_dict[f'{self.name_prefix}_w{index}'])
        sigma = self.final_layer(x)
        x = self.color_layer_sine(x, style_dict[f'{self.name_prefix}_rgb'])
        lin = self.color_layer_linear[0]
        if self.fused_linear and x.dim() == 3 and ops.points_linear_supported(x, lin.weight):
            rbg = ops.points_linear(x, lin.weight, lin.bias)      # tcgen05 split-fp16 GEMM, forward and data gradient
        else:
            rbg = self.color_layer_linear(x)
        return torch.cat([rbg, sigma], dim=-1)

    def forward(self, input, style_dict, ray_directions=None, **kwargs):
        _require_cuda(input, "NeRFNetwork.forward")
        return self.forward_with_frequencies_phase_shifts(input=input, style_dict=style_dict,
                                                          ray_directions=ray_directions, **kwargs)

    def staged_forward(self, transformed_points, transformed_ray_directions_expanded, style_dict, max_points,
                       num_steps):
        batch_size, num_points, _ = transformed_points.shape
        out = torch.zeros((batch_size, num_points, self.rgb_dim + 1), device=transformed_points.device)
        for b in range(batch_size):
            head = 0
            while head < num_points:
                tail = head + max_points
                out[b:b + 1, head:tail] = self(input=transformed_points[b:b + 1, head:tail],
                                               style_dict={n: s[b:b + 1] for n, s in style_dict.items()},
                                               ray_directions=transformed_ray_directions_expanded[b:b + 1, head:tail])
                head += max_points
        return out.view(batch_size, -1, num_steps, self.rgb_dim + 1)


class SinAct(nn.Module):
    def forward(self, x):
        return torch.sin(x)


class SinStyleMod(nn.Module):                   # mod_conv_fc.py:392-563 (kernel_size 1 "bmm" form)
    def __init__(self, in_channel, out_channel, kernel_size=1, style_dim=None, use_style_fc=False,
                 demodulate=True, use_group_conv=False, eps=1e-8, **kwargs):
        super().__init__()
        assert not use_group_conv and kernel_size == 1, "only the shipping 1x1 bmm form is built"
        self.eps, self.in_channel, self.out_channel, self.kernel_size = eps, in_channel, out_channel, kernel_size
        self.style_dim, self.use_style_fc, self.demodulate, self.use_group_conv = style_dim, use_style_fc, demodulate, False
        self.padding = 0
        self.weight = nn.Parameter(torch.randn(1, in_channel, out_channel))
        torch.nn.init.kaiming_normal_(self.weight[0], a=0.2, mode='fan_in', nonlinearity='leaky_relu')
        if use_style_fc:
            self.modulation = nn.Linear(style_dim, in_channel)
            self.modulation.apply(kaiming_leaky_init)
        else:
            self.style_dim = in_channel
        self.sin = SinAct()
        self.norm = nn.LayerNorm(in_channel)     # never used in forward; kept for state_dict parity

    def style_scale(self, style):
        """(B,in): s + 1"""
        s = self.modulation(style) if self.use_style_fc else style
        return s + 1

    def demod_scale(self, s1p):
        """(B,out): rsqrt(sum_i (W[i,o] * s1p[b,i])^2 + eps)"""
        if not self.demodulate:
            return torch.ones(s1p.shape[0], self.out_channel, device=s1p.device, dtype=s1p.dtype)
        return torch.rsqrt((s1p * s1p) @ (self.weight[0] * self.weight[0]) + self.eps)

    def forward(self, x, style, force_bmm=False):
        _require_cuda(x, "SinStyleMod.forward")
        s1p = self.style_scale(style)
        d = self.demod_scale(s1p)
        if x.dim() == 2:
            return ((x * s1p) @ self.weight[0]) * d
        return torch.matmul(x * s1p.unsqueeze(1), self.weight[0]) * d.unsqueeze(1)

    def _load_from_state_dict(self, state_dict, prefix, *a, **k):   # mod_conv_fc.py:184-278 analogue
        key = prefix + "weight"
        if key in state_dict and state_dict[key].dim() == 5:       # (1,out,in,1,1) -> (1,in,out)
            state_dict[key] = state_dict[key][:, :, :, 0, 0].permute(0, 2, 1).contiguous()
        return super()._load_from_state_dict(state_dict, prefix, *a, **k)


class SkipLayer(nn.Module):
    def forward(self, x0, x1):
        return x0 + x1


class SinBlock(nn.Module):                      # generator.py:893-980
    def __init__(self, in_dim, out_dim, style_dim, name_prefix):
        super().__init__()
        self.in_dim, self.out_dim, self.style_dim, self.name_prefix = in_dim, out_dim, style_dim, name_prefix
        self.style_dim_dict = {}
        self.mod1 = SinStyleMod(in_channel=in_dim, out_channel=out_dim, style_dim=style_dim, use_style_fc=True)
        self.style_dim_dict[f'{name_prefix}_0'] = self.mod1.style_dim
        self.act1 = nn.LeakyReLU(0.2, inplace=True)
        self.mod2 = SinStyleMod(in_channel=out_dim, out_channel=out_dim, style_dim=style_dim, use_style_fc=True)
        self.style_dim_dict[f'{name_prefix}_1'] = self.mod2.style_dim
        self.act2 = nn.LeakyReLU(0.2, inplace=True)
        self.skip = SkipLayer()

    def forward(self, x, style_dict, skip=False):
        x_orig = x
        x = self.act1(self.mod1(x, style_dict[f'{self.name_prefix}_0']))
        out = self.act2(self.mod2(x, style_dict[f'{self.name_prefix}_1']))
        if skip and out.shape[-1] == x_orig.shape[-1]:
            out = self.skip(out, x_orig)
        return out


class ToRGB(nn.Module):                         # generator.py:983-1006
    def __init__(self, in_dim, dim_rgb=3, use_equal_fc=False):
        super().__init__()
        assert not use_equal_fc
        self.in_dim, self.dim_rgb = in_dim, dim_rgb
        self.linear = nn.Linear(in_dim, dim_rgb)

    def forward(self, input, skip=None):
        out = self.linear(input)
        return out if skip is None else out + skip


class CIPSNet(nn.Module):                       # generator.py:1009-1154
    def __init__(self, input_dim, style_dim, hidden_dim=256, pre_rgb_dim=32, device=None, name_prefix='inr',
                 **kwargs):
        super().__init__()
        self.device, self.pre_rgb_dim, self.name_prefix = device, pre_rgb_dim, name_prefix
        self.input_dim, self.hidden_dim = input_dim, hidden_dim
        self.channels = {str(2 ** i): hidden_dim for i in range(2, 11)}
        self.module_name_list = []
        self.style_dim_dict = {}
        _out_dim = input_dim
        network, to_rbgs = OrderedDict(), OrderedDict()
        for i, (name, channel) in enumerate(self.channels.items()):
            _in_dim, _out_dim = _out_dim, channel
            blk = SinBlock(in_dim=_in_dim, out_dim=_out_dim, style_dim=style_dim, name_prefix=f'{name_prefix}_w{name}')
            self.style_dim_dict.update(blk.style_dim_dict)
            network[name] = blk
            to_rbgs[name] = ToRGB(in_dim=_out_dim, dim_rgb=pre_rgb_dim, use_equal_fc=False)
        self.network = nn.ModuleDict(network)
        self.to_rgbs = nn.ModuleDict(to_rbgs)
        self.to_rgbs.apply(frequency_init(100))
        self.module_name_list += ['network', 'to_rgbs']
        out_layers = []
        if pre_rgb_dim > 3:
            out_layers.append(nn.Linear(pre_rgb_dim, 3))
        out_layers.append(nn.Tanh())
        self.tanh = nn.Sequential(*out_layers)
        self.tanh.apply(frequency_init(100))
        self.module_name_list.append('tanh')

    def _n_blocks(self, img_size):
        stop = str(2 ** int(np.log2(img_size)))
        names = list(self.network.keys())
        return names.index(stop) + 1 if stop in names else len(names)

    def fused_supported(self):
        """mirrors the argument checks of c3d_cips_fwd (csrc/cips_tc.cu): hidden 512, <= 64 input features, RGB heads"""
        return self.pre_rgb_dim == 3 and self.hidden_dim == 512 and self.input_dim <= 64

    def forward_torch(self, input, style_dict, img_size=1024):
        x, rgb = input, 0
        n_run = self._n_blocks(img_size)
        for idx, (name, block) in enumerate(self.network.items()):
            x = block(x, style_dict, skip=idx >= 4)
            if idx >= 3:
                rgb = self.to_rgbs[name](x, skip=rgb)
            if idx + 1 == n_run:
                break
        return self.tanh(rgb) if not isinstance(rgb, int) else self.tanh(torch.zeros_like(input[..., :3]))

    def kernel_inputs(self, style_dict, n_blocks):
        if (os.environ.get("C3D_STYLE_PREP", "fused") == "fused" and not torch.is_grad_enabled()
                and all(b.mod1.use_style_fc and b.mod1.demodulate for b in self.network.values())):
            return self.kernel_inputs_fused(style_dict, n_blocks)
        ws, s1ps, ds, rw, rb = [], [], [], [], []
        for idx, (name, block) in enumerate(self.network.items()):
            if idx >= n_blocks:
                break
            for j, mod in enumerate((block.mod1, block.mod2)):
                s1p = mod.style_scale(style_dict[f'{block.name_prefix}_{j}'])
                ws.append(mod.weight[0])
                s1ps.append(s1p)
                ds.append(mod.demod_scale(s1p))
            rw.append(self.to_rgbs[name].linear.weight)
            rb.append(self.to_rgbs[name].linear.bias)
        return ws, s1ps, ds, rw, rb

    # 'torch': the autograd graph is built from torch CUDA ops (round-1 behaviour).  'fused': forward and backward chain on the
    # native kernels, weight gradients as fp16 library GEMMs (ops.CipsMLPFunction) -- opt-in, emulation-verified, gradients
    # agree with fp32 autograd to fp16-operand accuracy (~1e-3 relative), not bit for bit.
    train_backend = 'torch'

    def forward_fused_train(self, input, style_dict, img_size=1024):
        n_blocks = self._n_blocks(img_size)
        ws, s1ps, ds, rw, rb = self.kernel_inputs(style_dict, n_blocks)     # s1p / demod stay in the torch graph (-> modulation, W)
        tensors = list(ws) + list(s1ps) + list(ds) + list(rw[3:n_blocks]) + list(rb[3:n_blocks])
        return ops.CipsMLPFunction.apply(input, n_blocks, 4, 3, *tensors)

    def kernel_inputs_fused(self, style_dict, n_blocks):
        """kernel_inputs with s1p / demod of all layers from one native launch (inference only; C3D_STYLE_PREP=fused)"""
        mods, styles, rw, rb = [], [], [], []
        for idx, (name, block) in enumerate(self.network.items()):
            if idx >= n_blocks:
                break
            for j, mod in enumerate((block.mod1, block.mod2)):
                mods.append(mod)
                styles.append(style_dict[f'{block.name_prefix}_{j}'])
            rw.append(self.to_rgbs[name].linear.weight)
            rb.append(self.to_rgbs[name].linear.bias)
        ws = [m.weight[0] for m in mods]
        s1ps, ds = ops.cips_style_prep(styles, [m.modulation.weight for m in mods], [m.modulation.bias for m in mods], ws,
                                       eps=mods[0].eps)
        return ws, s1ps, ds, rw, rb

    def forward(self, input, style_dict, img_size=1024, **kwargs):
        """input (b, n, in) -> (b, n, 3)"""
        _require_cuda(input, "CIPSNet.forward")
        needs_graph = torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in self.parameters())
                                                   or any(s.requires_grad for s in style_dict.values()))
        if not needs_graph and (not self.fused_supported() or input.dim() != 3):
            _note_torch_path("CIPSNet", f"hidden_dim {self.hidden_dim}, input_dim {self.input_dim}, pre_rgb_dim {self.pre_rgb_dim}, "
                                        f"input rank {input.dim()} (fused: 512 / <= 64 / 3 / rank 3)")
        if needs_graph and self.train_backend == 'fused' and self.fused_supported() and input.dim() == 3 and self._n_blocks(img_size) > 3:
            return self.forward_fused_train(input, style_dict, img_size)
        if needs_graph or not self.fused_supported() or input.dim() != 3:
            return self.forward_torch(input, style_dict, img_size)
        n_blocks = self._n_blocks(img_size)
        ws, s1ps, ds, rw, rb = self.kernel_inputs(style_dict, n_blocks)
        return ops.cips_forward(input, ws, s1ps, ds, rw, rb, n_blocks=n_blocks, skip_from=4, rgb_from=3)


class PixelNorm(nn.Module):                     # multi_head_mapping.py:13-19
    def forward(self, input):
        assert input.dim() == 2
        return input * torch.rsqrt(torch.mean(input ** 2, dim=1, keepdim=True) + 1e-8)


class MultiHeadMappingNetwork(nn.Module):       # multi_head_mapping.py:28-153
    def __init__(self, z_dim, hidden_dim, base_layers, head_layers, head_dim_dict, add_norm=False,
                 norm_out=False, **kwargs):
        super().__init__()
        self.z_dim, self.head_dim_dict = z_dim, head_dim_dict
        out_dim = z_dim
        self.module_name_list = []
        self.norm = PixelNorm()
        base_net = []
        for i in range(base_layers):
            in_dim, out_dim = out_dim, hidden_dim
            base_layer_ = nn.Linear(in_features=in_dim, out_features=out_dim)
            base_layer_.apply(kaiming_leaky_init)
            base_net.append(base_layer_)
            if head_layers > 0 or i != base_layers - 1:
                if add_norm:
                    base_net.append(nn.LayerNorm(out_dim))
                base_net.append(nn.LeakyReLU(0.2, inplace=True))
        if len(base_net) > 0:
            if norm_out and head_layers <= 0:
                base_net.append(nn.LayerNorm(out_dim))
            self.base_net = nn.Sequential(*base_net)
            self.num_z = 1
            self.module_name_list.append('base_net')
        else:
            self.base_net = None
            self.num_z = len(head_dim_dict)
        head_in_dim = out_dim
        for name, head_dim in head_dim_dict.items():
            if head_layers > 0:
                head_net = []
                out_dim = head_in_dim
                for i in range(head_layers):
                    in_dim = out_dim
                    out_dim = head_dim if i == head_layers - 1 else hidden_dim
                    head_layer_ = nn.Linear(in_features=in_dim, out_features=out_dim)
                    head_layer_.apply(kaiming_leaky_init)
                    head_net.append(head_layer_)
                    if i != head_layers - 1:
                        head_net.append(nn.LeakyReLU(0.2, inplace=True))
                    elif norm_out:
                        head_net.append(nn.LayerNorm(out_dim))
                head_net = nn.Sequential(*head_net)
                self.module_name_list.append(name)
            else:
                head_net = nn.Identity()
            self.add_module(name, head_net)

    def forward(self, z):
        if self.base_net is not None:
            base_fea = self.base_net(self.norm(z))
            head_inputs = {name: base_fea for name in self.head_dim_dict.keys()}
        else:
            head_inputs = {name: self.norm(z[idx]) for idx, name in enumerate(self.head_dim_dict.keys())}
        return {name: getattr(self, name)(head_inputs[name]) for name in self.head_dim_dict.keys()}


# ======================================================================================
# differentiable torch restatement of the renderer (training graph only)
# ======================================================================================
def _torch_initial_rays(img_size, z_cam, ray_start, ray_end, num_steps, device):
    x = torch.linspace(-1, 1, img_size, device=device)[None, :].expand(img_size, img_size).reshape(-1)
    y = torch.linspace(1, -1, img_size, device=device)[:, None].expand(img_size, img_size).reshape(-1)
    d = torch.stack([x, y, torch.full_like(x, z_cam)], -1)
    return comm_utils.normalize_vecs(d), torch.linspace(ray_start, ray_end, num_steps, device=device)


def get_world_points_and_direction(batch_size, num_steps, img_size, fov, ray_start, ray_end, h_stddev, v_stddev, h_mean, v_mean,
                                   sample_dist, lock_view_dependence, device='cuda', camera_pos=None, camera_lookup=None,
                                   up_vector=None):
    """exp/comm/comm_utils.py:682-763 with the reference's signature, return tuple and RNG order (jitter rand(b, hw, s, 1), then
    the camera draws): (transformed_points (b, hw*s, 3), transformed_ray_directions_expanded (b, hw*s, 3), transformed_ray_origins
    (b, hw, 3), transformed_ray_directions (b, hw, 3), z_vals (b, hw, s, 1), pitch (b, 1), yaw (b, 1)).  O(points) elementwise
    work as torch ops on `device`: the explicit-points form of rows R2-R7, which the fused renderer computes in registers;
    it feeds `points_forward`."""
    dirs_cam, z_lin = _torch_initial_rays(img_size, ops.z_cam_from_fov(fov), ray_start, ray_end, num_steps, device)
    n = dirs_cam.shape[0]
    jitter_u = torch.rand((batch_size, n, num_steps, 1), device=device)[..., 0]                 # perturb_points, L416-437
    off = (jitter_u - 0.5) * (z_lin[1] - z_lin[0])
    z_vals = z_lin[None, None, :] + off
    p_cam = dirs_cam[None, :, None, :] * z_lin[None, None, :, None] + off[..., None] * dirs_cam[None, :, None, :]
    c2w, pitch, yaw = comm_utils.sample_cam2world(batch_size, device, h_stddev, v_stddev, h_mean, v_mean, sample_dist,
                                                  camera_pos=camera_pos, camera_lookup=camera_lookup, up_vector=up_vector)
    Rm, t = c2w[:, :3, :3], c2w[:, :3, 3]
    pts = torch.einsum("bij,bnsj->bnsi", Rm, p_cam) + t[:, None, None, :]                        # transform_sampled_points, L584-679
    dirs_w = torch.einsum("bij,nj->bni", Rm, dirs_cam)
    origins = t[:, None, :].expand(-1, n, -1).contiguous()
    dirs_exp = dirs_w[:, :, None, :].expand(-1, -1, num_steps, -1).reshape(batch_size, n * num_steps, 3)
    if lock_view_dependence:
        dirs_exp = torch.zeros_like(dirs_exp)
        dirs_exp[..., -1] = -1
    return pts.reshape(batch_size, n * num_steps, 3), dirs_exp, origins, dirs_w, z_vals[..., None], pitch, yaw


def _torch_integrate(rgb_sigma, z, noise, clamp_mode, last_back, white_back, dim_rgb):
    rgbs, sig = rgb_sigma[..., :dim_rgb], rgb_sigma[..., dim_rgb]
    deltas = z[..., 1:] - z[..., :-1]
    deltas = torch.cat([deltas, 1e10 * torch.ones_like(deltas[..., :1])], -1)
    if noise is not None:
        sig = sig + noise
    if clamp_mode == 'softplus':
        alphas = 1 - torch.exp(-deltas * F.softplus(sig))
    elif clamp_mode == 'relu':
        alphas = 1 - torch.exp(-deltas * F.relu(sig))
    else:
        assert 0, "Need to choose clamp mode"
    shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-10], -1)
    weights = alphas * torch.cumprod(shifted, -1)[..., :-1]
    wsum = weights.sum(-1)
    if last_back:
        weights = torch.cat([weights[..., :-1], weights[..., -1:] + (1 - wsum)[..., None]], -1)
    rgb = torch.sum(weights[..., None] * rgbs, -2)
    if white_back:
        rgb = rgb + 1 - wsum[..., None]
    return rgb, weights


def _integrate(backend, rgb_sigma, z, noise, clamp_mode, last_back, white_back, dim_rgb):
    """fancy_integration of the training graph: torch ops, or the native op when `backend == 'fused'` and the shape fits."""
    if backend == 'fused' and rgb_sigma.shape[-1] == dim_rgb + 1 and ops.integrate_supported(rgb_sigma, z, noise):
        return ops.integrate(rgb_sigma, z, noise, clamp_mode, last_back, white_back)
    return _torch_integrate(rgb_sigma, z, noise, clamp_mode, last_back, white_back, dim_rgb)


def _sample_pdf(backend, bins, weights, u, eps=1e-5):
    """sample_pdf of the (no_grad) resampling step: torch ops, or one native launch when `backend == 'fused'`."""
    if backend == 'fused' and ops.sample_pdf_supported(bins, weights):
        return ops.sample_pdf_from_u(bins, weights, u, eps)
    return _torch_sample_pdf(bins, weights, u, eps)


def _torch_sample_pdf(bins, weights, u, eps=1e-5):
    n_s = weights.shape[1]
    weights = weights + eps
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    inds = torch.searchsorted(cdf, u.contiguous())
    below, above = torch.clamp_min(inds - 1, 0), torch.clamp_max(inds, n_s)
    cb, ca = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bb, ba = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = ca - cb
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    return bb + (u - cb) / denom * (ba - bb)


# ======================================================================================
# generators
# ======================================================================================
@MODEL_REGISTRY.register(name_prefix=__name__)
class GeneratorNerfINR(nn.Module):              # generator.py:1158-1951
    def __init__(self, z_dim, nerf_cfg, mapping_nerf_cfg, inr_cfg, mapping_inr_cfg, device='cuda', **kwargs):
        super().__init__()
        self.epoch = 0
        self.step = 0
        self.z_dim = z_dim
        self.device = device
        self.module_name_list = []
        self.siren = NeRFNetwork(**nerf_cfg)
        self.module_name_list.append('siren')
        self.mapping_network_nerf = MultiHeadMappingNetwork(
            **{**mapping_nerf_cfg, 'head_dim_dict': self.siren.style_dim_dict})
        self.module_name_list.append('mapping_network_nerf')
        self.inr_net = CIPSNet(**{**inr_cfg, "input_dim": self.siren.rgb_dim})
        self.module_name_list.append('inr_net')
        self.mapping_network_inr = MultiHeadMappingNetwork(
            **{**mapping_inr_cfg, 'head_dim_dict': self.inr_net.style_dim_dict})
        self.module_name_list.append('mapping_network_inr')
        self.aux_to_rbg = nn.Sequential(nn.Linear(self.siren.rgb_dim, 3), nn.Tanh())
        self.aux_to_rbg.apply(frequency_init(25))
        self.module_name_list.append('aux_to_rbg')
        self.filters = nn.Identity()
        # RNG parity: the reference draws integration noise even when nerf_noise == 0
        # (pigan_utils.py:246).  Keep the draws (default) so torch's RNG stream stays aligned.
        self.skip_unused_noise_draws = False
        self.impl = None            # None -> env C3D_IMPL / tcgen05 default
        # measurement hook: run the differentiable torch-CUDA-op restatement even without autograd (the eager
        # PyTorch path the reference would execute on the same GPU); bench.py reports it beside the fused path
        self.force_torch_path = False

    # ---------------------------------------------------------------- small helpers
    def _nerf_grad_needed(self, style_dict=None):
        """does the renderer need an autograd graph?  Through the field's parameters OR through the styles (a frozen siren under a
        trainable mapping_network_nerf still needs the FiLM gradients)"""
        if not torch.is_grad_enabled():
            return False
        if any(p.requires_grad for p in self.siren.parameters()):
            return True
        return style_dict is not None and any(v.requires_grad for k, v in style_dict.items() if k in self.siren.style_dim_dict)

    def z_sampler(self, shape, device, dist='gaussian'):
        if dist == 'gaussian':
            return torch.randn(shape, device=device)
        elif dist == 'uniform':
            return torch.rand(shape, device=device) * 2 - 1

    def get_zs(self, b, batch_split=1):                             # generator.py:1774-1794
        z_nerf = self.z_sampler(shape=(b, self.mapping_network_nerf.z_dim), device=self.device)
        z_inr = self.z_sampler(shape=(b, self.mapping_network_inr.z_dim), device=self.device)
        if batch_split > 1:
            return [{'z_nerf': a, 'z_inr': c} for a, c in
                    zip(z_nerf.split(b // batch_split), z_inr.split(b // batch_split))]
        return {'z_nerf': z_nerf, 'z_inr': z_inr}

    def mapping_network(self, z_nerf, z_inr):                       # generator.py:1796-1802
        style_dict = {}
        style_dict.update(self.mapping_network_nerf(z_nerf))
        style_dict.update(self.mapping_network_inr(z_inr))
        return style_dict

    def generate_avg_frequencies(self, num_samples=10000, device='cuda'):   # generator.py:1804-1817
        zs = self.get_zs(num_samples)
        with torch.no_grad():
            style_dict = self.mapping_network(**zs)
        self.avg_styles = {name: style.mean(0, keepdim=True) for name, style in style_dict.items()}
        return self.avg_styles

    def get_truncated_freq_phase(self, raw_style_dict, avg_style_dict, raw_lambda):  # generator_nerf_inr.py:770-782
        return {name: avg + raw_lambda * (raw_style_dict[name] - avg) for name, avg in avg_style_dict.items()}

    def get_batch_style_dict(self, b, style_dict):
        return {name: style[[b]] for name, style in style_dict.items()}

    def staged_forward(self, *args, **kwargs):
        raise NotImplementedError

    def set_device(self, device):                                   # generator.py:1822-1826 (no-op)
        pass

    # ---------------------------------------------------------------- random draws (reference order)
    def _draw(self, kind, shape, needed=True):
        if not needed and self.skip_unused_noise_draws:
            return None
        fn = torch.rand if kind == 'rand' else torch.randn
        return fn(shape, device=self.device)

    # ---------------------------------------------------------------- renderer front-end
    def render_pixels_fea(self, style_dict, cam2world, jitter_u, pdf_u, noise_c, noise_f, *, img_size, fov,
                          ray_start, ray_end, num_steps, hierarchical_sample, clamp_mode, nerf_noise,
                          white_back, last_back, ray_idx=None, grad=False):
        """rays -> (B,N,32) integrated features.  jitter_u (B,HW,S), noise_c (B,N,S), pdf_u (B*N,S),
        noise_f (B,N,nS); ray_idx: LongTensor ray subset or None."""
        _require_cuda(cam2world, "GeneratorNerfINR")
        fusable = self.siren.fused_supported(num_steps)
        if not fusable and not grad:
            _note_torch_path("NeRFNetwork renderer", f"in_dim {self.siren.in_dim}, hidden_dim {self.siren.hidden_dim}, hidden_layers "
                             f"{self.siren.hidden_layers}, rgb_dim {self.siren.rgb_dim}, num_steps {num_steps} (fused: 3 / 128 / 2 / 32 / 3..32)")
        if not grad and not self.force_torch_path and fusable:
            out = ops.render_features(
                self.siren.kernel_weights(), self.siren.kernel_film(style_dict), cam2world, jitter_u, pdf_u,
                noise_c, noise_f, img_size=img_size, fov=fov, ray_start=ray_start, ray_end=ray_end,
                num_steps=num_steps, hierarchical_sample=hierarchical_sample, clamp_mode=clamp_mode,
                noise_std=nerf_noise, white_back=white_back, last_back=last_back, ray_idx=ray_idx, impl=self.impl)
            return out["pixels_fea"]
        return self._render_torch(style_dict, cam2world, jitter_u, pdf_u, noise_c, noise_f, img_size=img_size,
                                  fov=fov, ray_start=ray_start, ray_end=ray_end, num_steps=num_steps,
                                  hierarchical_sample=hierarchical_sample, clamp_mode=clamp_mode,
                                  nerf_noise=nerf_noise, white_back=white_back, last_back=last_back,
                                  ray_idx=ray_idx)

    # volume integration of the autograd graph: 'torch' (round-1 behaviour) or 'fused' (ops.IntegrateFunction, csrc/integrate_ops.cu:
    # one native pass forward, one backward) -- opt-in like FiLMLayer.fused_film (validated and timed on a B200 in round 2, DESIGN.md 4.13;
    # the default stays the torch-op graph of the reference)
    train_integrate = 'torch'

    def _render_torch(self, style_dict, c2w, jitter_u, pdf_u, noise_c, noise_f, *, img_size, fov, ray_start,
                      ray_end, num_steps, hierarchical_sample, clamp_mode, nerf_noise, white_back, last_back,
                      ray_idx=None):
        """Autograd graph of rows R2-R12 (training only; CUDA tensors)."""
        dev, S = c2w.device, num_steps
        dirs_cam, z_vals = _torch_initial_rays(img_size, ops.z_cam_from_fov(fov), ray_start, ray_end, S, dev)
        off = (jitter_u - 0.5) * (z_vals[1] - z_vals[0])
        z = z_vals[None, None, :] + off
        p_cam = dirs_cam[None, :, None, :] * z_vals[None, None, :, None] + off[..., None] * dirs_cam[None, :, None, :]
        Rm, t = c2w[:, :3, :3], c2w[:, :3, 3]
        pts = torch.einsum("bij,bnsj->bnsi", Rm, p_cam) + t[:, None, None, :]
        dirs_w = torch.einsum("bij,nj->bni", Rm, dirs_cam)
        if ray_idx is not None:
            pts, z, dirs_w = pts[:, ray_idx], z[:, ray_idx], dirs_w[:, ray_idx]
        B, N = z.shape[:2]
        dim_rgb = self.siren.rgb_dim
        coarse = self.siren(pts.reshape(B, N * S, 3), style_dict, None).reshape(B, N, S, -1)
        nc = noise_c * nerf_noise if noise_c is not None else None
        nf = noise_f * nerf_noise if noise_f is not None else None
        if hierarchical_sample:
            with torch.no_grad():                                    # generator_nerf_inr.py:537 (@no_grad)
                _, w = _integrate(self.train_integrate, coarse, z, nc, clamp_mode, False, False, dim_rgb)
                w = w.reshape(B * N, S) + 1e-5
                zz = z.reshape(B * N, S)
                fz = _sample_pdf(self.train_integrate, 0.5 * (zz[:, :-1] + zz[:, 1:]), w[:, 1:-1], pdf_u).reshape(B, N, S)
                fpts = t[:, None, None, :] + dirs_w[:, :, None, :] * fz[..., None]
            fine = self.siren(fpts.reshape(B, N * S, 3), style_dict, None).reshape(B, N, S, -1)
            if (self.train_integrate == 'fused' and fine.shape[-1] == dim_rgb + 1
                    and ops.integrate_merged_supported(fine, fz, coarse, z, nf)):
                # cat + sort + gather + integration in one native pass each way (csrc/integrate_ops.cu, merged form)
                return ops.integrate_merged(fine, fz, coarse, z, nf, clamp_mode, last_back, white_back)[0]
            all_out = torch.cat([fine, coarse], dim=-2)
            all_z, ind = torch.sort(torch.cat([fz, z], dim=-1), dim=-1)
            all_out = torch.gather(all_out, -2, ind[..., None].expand(-1, -1, -1, all_out.shape[-1]))
        else:
            all_out, all_z = coarse, z
        fea, _ = _integrate(self.train_integrate, all_out, all_z, nf, clamp_mode, last_back, white_back, dim_rgb)
        return fea

    # ---------------------------------------------------------------- forward paths
    def forward(self, zs, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev, hierarchical_sample,
                h_mean=math.pi * 0.5, v_mean=math.pi * 0.5, psi=1, sample_dist=None, lock_view_dependence=False,
                clamp_mode='relu', nerf_noise=0., white_back=False, last_back=False, return_aux_img=False,
                grad_points=None, forward_points=None, **kwargs):
        """generator.py:1256-1370.  -> imgs (b or 2b, 3, h, w), pitch_yaw (b or 2b, 2)"""
        return self._forward_impl(zs, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev,
                                  hierarchical_sample, h_mean, v_mean, psi, sample_dist, lock_view_dependence,
                                  clamp_mode, nerf_noise, white_back, last_back, return_aux_img, grad_points,
                                  forward_points, None, None, None)

    def forward_camera_pos_and_lookup(self, zs, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev,
                                      h_mean, v_mean, hierarchical_sample, camera_pos, camera_lookup, psi=1,
                                      sample_dist=None, lock_view_dependence=False, clamp_mode='relu',
                                      nerf_noise=0., white_back=False, last_back=False, return_aux_img=False,
                                      grad_points=None, forward_points=None, up_vector=None, **kwargs):
        """generator.py:1828-1951"""
        return self._forward_impl(zs, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev,
                                  hierarchical_sample, h_mean, v_mean, psi, sample_dist, lock_view_dependence,
                                  clamp_mode, nerf_noise, white_back, last_back, return_aux_img, grad_points,
                                  forward_points, camera_pos, camera_lookup, up_vector)

    def _forward_impl(self, zs, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev,
                      hierarchical_sample, h_mean, v_mean, psi, sample_dist, lock_view_dependence, clamp_mode,
                      nerf_noise, white_back, last_back, return_aux_img, grad_points, forward_points,
                      camera_pos, camera_lookup, up_vector):
        style_dict = self.mapping_network(**zs)
        if psi < 1:
            avg_styles = self.generate_avg_frequencies(device=self.device)
            style_dict = self.get_truncated_freq_phase(raw_style_dict=style_dict, avg_style_dict=avg_styles,
                                                       raw_lambda=psi)
        common = dict(style_dict=style_dict, img_size=img_size, fov=fov, ray_start=ray_start, ray_end=ray_end,
                      num_steps=num_steps, h_stddev=h_stddev, v_stddev=v_stddev, h_mean=h_mean, v_mean=v_mean,
                      hierarchical_sample=hierarchical_sample, sample_dist=sample_dist,
                      lock_view_dependence=lock_view_dependence, clamp_mode=clamp_mode, nerf_noise=nerf_noise,
                      white_back=white_back, last_back=last_back, return_aux_img=return_aux_img,
                      camera_pos=camera_pos, camera_lookup=camera_lookup)
        if grad_points is not None and grad_points < img_size ** 2:
            return self.part_grad_forward(grad_points=grad_points, **common)
        return self.whole_grad_forward(forward_points=forward_points, up_vector=up_vector, **common)

    def _pixels_to_imgs(self, pixels_fea, style_dict, return_aux_img, img_size, pitch, yaw):
        if self.force_torch_path:
            inr_img = self.inr_net.forward_torch(pixels_fea, style_dict)
        else:
            inr_img = self.inr_net(pixels_fea, style_dict)                      # generator.py:1754
        B = inr_img.shape[0]
        inr_img = inr_img.view(B, img_size, img_size, 3).permute(0, 3, 1, 2)
        inr_img = self.filters(inr_img)
        pitch_yaw = torch.cat([pitch, yaw], -1)
        if return_aux_img:
            aux_img = self._aux(pixels_fea).view(B, img_size, img_size, 3).permute(0, 3, 1, 2)
            return torch.cat([inr_img, aux_img]), torch.cat([pitch_yaw, pitch_yaw])
        return inr_img, pitch_yaw

    def _aux(self, pixels_fea):
        return self.aux_to_rbg(pixels_fea)

    # ---------------------------------------------------------------- explicit-points entry (generator.py:1659-1762)
    def _field(self, points, style_dict, ray_directions):
        """the NeRF field on explicit points; subclasses decide whether it carries a graph"""
        return self.siren(input=points, style_dict=style_dict, ray_directions=ray_directions)

    @torch.no_grad()
    def get_fine_points_and_direction(self, coarse_output, z_vals, dim_rgb, clamp_mode, nerf_noise, num_steps,
                                      transformed_ray_origins, transformed_ray_directions):
        """generator_nerf_inr.py:537-598: coarse weights -> sample_pdf over the bin mid-points -> fine points, all without a
        graph.  Same draws in the same order as the reference (randn of sigma's shape, then rand(rays, num_steps))."""
        b = coarse_output.shape[0]
        z = z_vals.reshape(coarse_output.shape[:-1])
        noise = torch.randn(coarse_output.shape[:-1] + (1,), device=coarse_output.device)[..., 0] * nerf_noise   # pigan_utils.py:246
        _, weights = _integrate(self.train_integrate, coarse_output, z, noise, clamp_mode, False, False, dim_rgb)
        weights = weights.reshape(-1, num_steps) + 1e-5
        zz = z.reshape(-1, num_steps)
        u = torch.rand(zz.shape[0], num_steps, device=zz.device)                                              # pigan_utils.py:192
        fine_z = _sample_pdf(self.train_integrate, 0.5 * (zz[:, :-1] + zz[:, 1:]), weights[:, 1:-1], u).detach()
        fine_z = fine_z.reshape(b, -1, num_steps, 1)
        fine_points = transformed_ray_origins.unsqueeze(2) + transformed_ray_directions.unsqueeze(2) * fine_z.expand(-1, -1, -1, 3)
        return fine_points.reshape(b, -1, 3), fine_z

    def points_forward(self, style_dict, transformed_points, transformed_ray_directions_expanded, num_steps, hierarchical_sample,
                       z_vals, clamp_mode, nerf_noise, transformed_ray_origins, transformed_ray_directions, white_back, last_back,
                       return_aux_img, idx_grad=None):
        """generator.py:1659-1762 (reference signature): field on the given points (b, n, s, 3) -> resampling -> merge ->
        fancy_integration -> CIPS MLP [-> aux head]; returns (inr_img (b, n, 3), aux_img (b, n, 3) or None).  The points are
        the caller's, so this is the differentiable form of the path: the field runs as torch ops, integration / resampling /
        merge on the native ops when `train_integrate == 'fused'`, the CIPS MLP on the fused kernels when no graph is needed.
        `forward` itself never comes through here: it renders from the cameras with the fused kernels."""
        _require_cuda(transformed_points, "GeneratorNerfINR.points_forward")
        if idx_grad is not None:
            pick = lambda t: t.index_select(1, idx_grad)                   # comm_utils.gather_points, L262-282  # noqa: E731
            transformed_points, transformed_ray_directions_expanded, z_vals = (
                pick(transformed_points), pick(transformed_ray_directions_expanded), pick(z_vals))
            transformed_ray_origins, transformed_ray_directions = pick(transformed_ray_origins), pick(transformed_ray_directions)
        b, n = transformed_points.shape[:2]
        dim_rgb = self.siren.rgb_dim
        dirs = transformed_ray_directions_expanded.reshape(b, n * num_steps, 3)
        coarse = self._field(transformed_points.reshape(b, n * num_steps, 3), style_dict, dirs).reshape(b, n, num_steps, -1)
        z = z_vals.reshape(b, n, num_steps)
        if hierarchical_sample:
            fine_points, fine_z = self.get_fine_points_and_direction(
                coarse_output=coarse, z_vals=z_vals, dim_rgb=dim_rgb, clamp_mode=clamp_mode, nerf_noise=nerf_noise,
                num_steps=num_steps, transformed_ray_origins=transformed_ray_origins,
                transformed_ray_directions=transformed_ray_directions)
            fine = self._field(fine_points, style_dict, dirs).reshape(b, n, num_steps, -1)
            fz = fine_z[..., 0]
            noise = torch.randn(b, n, 2 * num_steps, 1, device=coarse.device)[..., 0] * nerf_noise            # pigan_utils.py:246
            if self.train_integrate == 'fused' and ops.integrate_merged_supported(fine, fz, coarse, z, noise):
                pixels_fea = ops.integrate_merged(fine, fz, coarse, z, noise, clamp_mode, last_back, white_back)[0]
            else:
                all_z, ind = torch.sort(torch.cat([fz, z], -1), dim=-1)                                        # L1733-1738
                all_out = torch.gather(torch.cat([fine, coarse], -2), -2, ind[..., None].expand(-1, -1, -1, coarse.shape[-1]))
                pixels_fea, _ = _torch_integrate(all_out, all_z, noise, clamp_mode, last_back, white_back, dim_rgb)
        else:
            noise = torch.randn(b, n, num_steps, 1, device=coarse.device)[..., 0] * nerf_noise
            pixels_fea, _ = _integrate(self.train_integrate, coarse, z, noise, clamp_mode, last_back, white_back, dim_rgb)
        inr_img = self.inr_net(pixels_fea, style_dict)
        return inr_img, (self._aux(pixels_fea) if return_aux_img else None)

    def whole_grad_forward(self, style_dict, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev,
                           h_mean, v_mean, hierarchical_sample, sample_dist=None, lock_view_dependence=False,
                           clamp_mode='relu', nerf_noise=0., white_back=False, last_back=False,
                           return_aux_img=True, forward_points=None, camera_pos=None, camera_lookup=None,
                           up_vector=None):
        """generator.py:1378-1534.  The reference's forward_points chunking (a memory workaround) keeps
        its RNG call order here, but all chunks are rendered by ONE fused launch."""
        device = self.device
        B = list(style_dict.values())[0].shape[0]
        HW, S = img_size ** 2, num_steps
        nS = 2 * S if hierarchical_sample else S
        need_noise = nerf_noise != 0
        cam = dict(h_stddev=h_stddev, v_stddev=v_stddev, h_mean=h_mean, v_mean=v_mean, mode=sample_dist)
        if forward_points is not None:
            with torch.no_grad():
                jit, c2ws, pitches, yaws, ncs, us, nfs = [], [], [], [], [], [], []
                for b in range(B):
                    jit.append(self._draw('rand', (1, HW, S, 1)))
                    cp = camera_pos[b:b + 1] if camera_pos is not None and camera_pos.shape[0] == B else camera_pos
                    cl = camera_lookup[b:b + 1] if camera_lookup is not None and camera_lookup.shape[0] == B else camera_lookup
                    c2w, pitch, yaw = comm_utils.sample_cam2world(1, device, camera_pos=cp, camera_lookup=cl,
                                                                  up_vector=up_vector, **cam)
                    c2ws.append(c2w), pitches.append(pitch), yaws.append(yaw)
                    head = 0
                    while head < HW:
                        n = min(forward_points, HW - head)
                        if hierarchical_sample:
                            ncs.append(self._draw('randn', (1, n, S, 1), need_noise))
                            us.append(self._draw('rand', (n, S)))
                        nfs.append(self._draw('randn', (1, n, nS, 1), need_noise))
                        head += forward_points
                cat = lambda lst, d: None if (not lst or lst[0] is None) else torch.cat(lst, d)
                # chunks of one image are consecutive ray ranges -> concatenating restores (B,HW,.)
                noise_c = cat(ncs, 1)
                noise_c = noise_c.view(B, HW, S) if noise_c is not None else None
                noise_f = cat(nfs, 1)
                noise_f = noise_f.view(B, HW, nS) if noise_f is not None else None
                pdf_u = cat(us, 0)
                pixels_fea = self.render_pixels_fea(
                    style_dict, torch.cat(c2ws, 0), torch.cat(jit, 0).view(B, HW, S), pdf_u, noise_c, noise_f,
                    img_size=img_size, fov=fov, ray_start=ray_start, ray_end=ray_end, num_steps=S,
                    hierarchical_sample=hierarchical_sample, clamp_mode=clamp_mode, nerf_noise=nerf_noise,
                    white_back=white_back, last_back=last_back, grad=False)
                return self._pixels_to_imgs(pixels_fea, style_dict, return_aux_img, img_size,
                                            torch.cat(pitches, 0), torch.cat(yaws, 0))
        jitter_u = self._draw('rand', (B, HW, S, 1)).view(B, HW, S)
        c2w, pitch, yaw = comm_utils.sample_cam2world(B, device, camera_pos=camera_pos,
                                                      camera_lookup=camera_lookup, up_vector=up_vector, **cam)
        noise_c = pdf_u = None
        if hierarchical_sample:
            noise_c = self._draw('randn', (B, HW, S, 1), need_noise)
            noise_c = noise_c.view(B, HW, S) if noise_c is not None else None
            pdf_u = self._draw('rand', (B * HW, S))
        noise_f = self._draw('randn', (B, HW, nS, 1), need_noise)
        noise_f = noise_f.view(B, HW, nS) if noise_f is not None else None
        pixels_fea = self.render_pixels_fea(
            style_dict, c2w, jitter_u, pdf_u, noise_c, noise_f, img_size=img_size, fov=fov, ray_start=ray_start,
            ray_end=ray_end, num_steps=S, hierarchical_sample=hierarchical_sample, clamp_mode=clamp_mode,
            nerf_noise=nerf_noise, white_back=white_back, last_back=last_back, grad=self._nerf_grad_needed(style_dict))
        return self._pixels_to_imgs(pixels_fea, style_dict, return_aux_img, img_size, pitch, yaw)

    def part_grad_forward(self, style_dict, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev,
                          h_mean, v_mean, hierarchical_sample, sample_dist=None, lock_view_dependence=False,
                          clamp_mode='relu', nerf_noise=0., white_back=False, last_back=False, return_aux_img=True,
                          grad_points=None, camera_pos=None, camera_lookup=None, **kwargs):
        """generator.py:1536-1657: a random pixel subset carries gradient, the rest is rendered no_grad."""
        device = self.device
        B = list(style_dict.values())[0].shape[0]
        HW, S = img_size ** 2, num_steps
        nS = 2 * S if hierarchical_sample else S
        need_noise = nerf_noise != 0
        jitter_u = self._draw('rand', (B, HW, S, 1)).view(B, HW, S)
        c2w, pitch, yaw = comm_utils.sample_cam2world(B, device, h_stddev=h_stddev, v_stddev=v_stddev,
                                                      h_mean=h_mean, v_mean=v_mean, mode=sample_dist,
                                                      camera_pos=camera_pos, camera_lookup=camera_lookup)
        assert HW > grad_points
        rand_idx = torch.randperm(HW, device=device)
        parts = []
        for idx, with_grad in ((rand_idx[:grad_points], True), (rand_idx[grad_points:], False)):
            n = idx.numel()
            noise_c = pdf_u = None
            if hierarchical_sample:
                noise_c = self._draw('randn', (B, n, S, 1), need_noise)
                noise_c = noise_c.view(B, n, S) if noise_c is not None else None
                pdf_u = self._draw('rand', (B * n, S))
            noise_f = self._draw('randn', (B, n, nS, 1), need_noise)
            noise_f = noise_f.view(B, n, nS) if noise_f is not None else None
            # the grad half inherits the CALLER's grad mode (the reference does not re-enable it: generator.py:1536-1657)
            ctx = contextlib.nullcontext() if with_grad else torch.no_grad()
            with ctx:
                fea = self.render_pixels_fea(
                    style_dict, c2w, jitter_u, pdf_u, noise_c, noise_f, img_size=img_size, fov=fov,
                    ray_start=ray_start, ray_end=ray_end, num_steps=S, hierarchical_sample=hierarchical_sample,
                    clamp_mode=clamp_mode, nerf_noise=nerf_noise, white_back=white_back, last_back=last_back,
                    ray_idx=idx, grad=with_grad and self._nerf_grad_needed(style_dict))
                inr = self.inr_net(fea, style_dict)
                aux = self._aux(fea) if return_aux_img else None
            parts.append((idx, inr, aux))
        (ig, inr_g, aux_g), (ing, inr_n, aux_n) = parts
        inr_img = comm_utils.scatter_points(ig, inr_g, ing, inr_n, HW).view(B, img_size, img_size, 3).permute(0, 3, 1, 2)
        inr_img = self.filters(inr_img)
        pitch_yaw = torch.cat([pitch, yaw], -1)
        if return_aux_img:
            aux_img = comm_utils.scatter_points(ig, aux_g, ing, aux_n, HW).view(B, img_size, img_size, 3).permute(0, 3, 1, 2)
            return torch.cat([inr_img, aux_img]), torch.cat([pitch_yaw, pitch_yaw])
        return inr_img, pitch_yaw


@MODEL_REGISTRY.register(name_prefix=__name__)
class GeneratorNerfINR_freeze_NeRF(GeneratorNerfINR):      # generator.py:1954-2078
    """NeRF + mapping_nerf + aux head run without gradient -> the fused renderer is always used."""

    def load_nerf_ema(self, G_ema):
        self.siren.load_state_dict(G_ema.siren.state_dict())
        self.mapping_network_nerf.load_state_dict(G_ema.mapping_network_nerf.state_dict())
        self.aux_to_rbg.load_state_dict(G_ema.aux_to_rbg.state_dict())

    def mapping_network(self, z_nerf, z_inr):
        style_dict = {}
        with torch.no_grad():
            style_dict.update(self.mapping_network_nerf(z_nerf))
        style_dict.update(self.mapping_network_inr(z_inr))
        return style_dict

    def _nerf_grad_needed(self, style_dict=None):
        return False

    def render_pixels_fea(self, *a, **k):
        k["grad"] = False
        with torch.no_grad():
            return super().render_pixels_fea(*a, **k)

    def _aux(self, pixels_fea):
        with torch.no_grad():
            return self.aux_to_rbg(pixels_fea)

    def _field(self, points, style_dict, ray_directions):                   # generator.py:2013-2044: the field under no_grad
        with torch.no_grad():
            return self.siren(input=points, style_dict=style_dict, ray_directions=ray_directions)
