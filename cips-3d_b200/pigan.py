"""pi-GAN class surface on the native renderer.

Mirrors piGAN_lib/siren/siren.py (CustomMappingNetwork, FiLMLayer, TALLSIREN, SPATIALSIRENBASELINE, the init
functions) and piGAN_lib/generators/generators.py (ImplicitGenerator3d: forward, staged_forward,
generate_avg_frequencies, forward_with_frequencies, staged_forward_with_frequencies): constructor arguments, method
signatures, attribute names, state_dict keys / shapes / order, constructor RNG order (same init bit for bit) and the
torch.rand / randn call order of the forward (perturbation, yaw, pitch, coarse noise, pdf u, final noise).

Inference (no autograd graph needed) runs the whole renderer natively through c3d_pigan_render_fwd -- rays, the
256-wide 8-layer FiLM-SIREN at 2S samples per ray, resampling, compositing; when a graph is required the same math
runs as differentiable torch CUDA ops.  CUDA only."""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from .comm_utils import sample_cam2world

_PIGAN_SAMPLE_DISTS = ('uniform', 'normal', 'gaussian', 'hybrid', 'truncated_gaussian', 'spherical_uniform')


# ------------------------------------------------------------------------------ siren.py:14-45 init functions
def sine_init(m):
    with torch.no_grad():
        if isinstance(m, nn.Linear):
            num_input = m.weight.size(-1)
            m.weight.uniform_(-np.sqrt(6 / num_input) / 30, np.sqrt(6 / num_input) / 30)


def first_layer_sine_init(m):
    with torch.no_grad():
        if isinstance(m, nn.Linear):
            num_input = m.weight.size(-1)
            m.weight.uniform_(-1 / num_input, 1 / num_input)


film_sine_init = sine_init
first_layer_film_sine_init = first_layer_sine_init


def kaiming_leaky_init(m):
    if m.__class__.__name__.find('Linear') != -1:
        torch.nn.init.kaiming_normal_(m.weight, a=0.2, mode='fan_in', nonlinearity='leaky_relu')


def frequency_init(freq):                                    # siren.py:75-81
    def init(m):
        with torch.no_grad():
            if isinstance(m, nn.Linear):
                num_input = m.weight.size(-1)
                m.weight.uniform_(-np.sqrt(6 / num_input) / freq, np.sqrt(6 / num_input) / freq)
    return init


class CustomMappingNetwork(nn.Module):                       # siren.py:47-73
    def __init__(self, z_dim, map_hidden_dim, map_output_dim):
        super().__init__()
        self.network = nn.Sequential(nn.Linear(z_dim, map_hidden_dim), nn.LeakyReLU(0.2, inplace=True),
                                     nn.Linear(map_hidden_dim, map_hidden_dim), nn.LeakyReLU(0.2, inplace=True),
                                     nn.Linear(map_hidden_dim, map_hidden_dim), nn.LeakyReLU(0.2, inplace=True),
                                     nn.Linear(map_hidden_dim, map_output_dim))
        self.network.apply(kaiming_leaky_init)
        with torch.no_grad():
            self.network[-1].weight *= 0.25

    def forward(self, z):
        fo = self.network(z)
        h = fo.shape[-1] // 2
        return fo[..., :h], fo[..., h:]


class FiLMLayer(nn.Module):                                  # siren.py:83-94
    def __init__(self, input_dim, hidden_dim):
        super().__init__()
        self.layer = nn.Linear(input_dim, hidden_dim)
        self.fused_film = False      # opt-in: the native FiLM + sine autograd op (ops.FilmSinFunction, csrc/film_ops.cu)

    def forward(self, x, freq, phase_shift):
        x = self.layer(x)
        if self.fused_film and x.dim() == 3 and ops.film_sin_supported(x, freq, phase_shift):
            return ops.film_sin(x, freq, phase_shift)
        return torch.sin(freq.unsqueeze(1).expand_as(x) * x + phase_shift.unsqueeze(1).expand_as(x))


class UniformBoxWarp(nn.Module):                             # siren.py:154-159
    def __init__(self, sidelength):
        super().__init__()
        self.scale_factor = 2 / sidelength

    def forward(self, coordinates):
        return coordinates * self.scale_factor


class _Siren(nn.Module):
    """Shared body of TALLSIREN (siren.py:97-152) and SPATIALSIRENBASELINE (siren.py:160-215)."""
    _gridwarp = False

    def __init__(self, input_dim=2, z_dim=100, hidden_dim=256, output_dim=1, device=None):
        super().__init__()
        self.device = device
        self.input_dim = input_dim
        self.z_dim = z_dim
        self.hidden_dim = hidden_dim
        self.output_dim = output_dim
        first_in = 3 if self._gridwarp else input_dim             # siren.py:171 hard-codes 3 for the spatial variant
        self.network = nn.ModuleList([FiLMLayer(first_in, hidden_dim)] + [FiLMLayer(hidden_dim, hidden_dim) for _ in range(7)])
        self.final_layer = nn.Linear(hidden_dim, 1)
        self.color_layer_sine = FiLMLayer(hidden_dim + 3, hidden_dim)
        self.color_layer_linear = self._make_color_linear(hidden_dim)
        self.mapping_network = CustomMappingNetwork(z_dim, 256, (len(self.network) + 1) * hidden_dim * 2)
        self.network.apply(frequency_init(25))
        self.final_layer.apply(frequency_init(25))
        self.color_layer_sine.apply(frequency_init(25))
        self.color_layer_linear.apply(frequency_init(25))
        self.network[0].apply(first_layer_film_sine_init)
        if self._gridwarp:
            self.gridwarper = UniformBoxWarp(0.24)

    def _make_color_linear(self, hidden_dim):
        raise NotImplementedError

    def forward(self, input, z, ray_directions, **kwargs):
        frequencies, phase_shifts = self.mapping_network(z)
        return self.forward_with_frequencies_phase_shifts(input, frequencies, phase_shifts, ray_directions, **kwargs)

    def forward_with_frequencies_phase_shifts(self, input, frequencies, phase_shifts, ray_directions, **kwargs):
        """(B,P,3), raw frequencies / phase shifts (B, 9*hidden), (B,P,3) -> (B,P,4) = [rgb in (0,1), sigma]"""
        frequencies = frequencies * 15 + 30
        x = self.gridwarper(input) if self._gridwarp else input
        H = self.hidden_dim
        for index, layer in enumerate(self.network):
            x = layer(x, frequencies[..., index * H:(index + 1) * H], phase_shifts[..., index * H:(index + 1) * H])
        sigma = self.final_layer(x)
        rbg = self.color_layer_sine(torch.cat([ray_directions, x], dim=-1), frequencies[..., -H:], phase_shifts[..., -H:])
        rbg = torch.sigmoid(self._color_linear_raw(rbg))
        return torch.cat([rbg, sigma], dim=-1)

    def _color_linear_raw(self, x):
        return self.color_layer_linear[0](x)

    # ---- what the native renderer needs
    def fused_supported(self):
        return self.hidden_dim <= 256 and len(self.network) <= 8 and self.network[0].layer.in_features == 3

    def kernel_weights(self, frequencies, phase_shifts):
        H, L = self.hidden_dim, len(self.network)
        fr = (frequencies * 15 + 30).float()
        ph = phase_shifts.float()
        return dict(w=[l.layer.weight for l in self.network], b=[l.layer.bias for l in self.network],
                    freq=[fr[:, i * H:(i + 1) * H].contiguous() for i in range(L)] + [fr[:, -H:].contiguous()],
                    phase=[ph[:, i * H:(i + 1) * H].contiguous() for i in range(L)] + [ph[:, -H:].contiguous()],
                    w_sigma=self.final_layer.weight, b_sigma=self.final_layer.bias,
                    wc=self.color_layer_sine.layer.weight, bc=self.color_layer_sine.layer.bias,
                    wl=self.color_layer_linear[0].weight, bl=self.color_layer_linear[0].bias,
                    hidden=H, gridwarp=self._gridwarp)


class TALLSIREN(_Siren):
    """Primary SIREN architecture used in pi-GAN generators (siren.py:97-152)."""
    _gridwarp = False

    def _make_color_linear(self, hidden_dim):
        return nn.Sequential(nn.Linear(hidden_dim, 3), nn.Sigmoid())


class SPATIALSIRENBASELINE(_Siren):
    """TALLSIREN + UniformBoxWarp(0.24) on the input points (siren.py:160-215)."""
    _gridwarp = True

    def _make_color_linear(self, hidden_dim):
        return nn.Sequential(nn.Linear(hidden_dim, 3))


def _require_cuda(t, what):
    if not t.is_cuda:
        raise _lib.C3dError(f"{what}: cips3d_b200 runs on CUDA tensors only (there is no CPU path)")


class ImplicitGenerator3d(nn.Module):                        # generators.py:12-350
    def __init__(self, siren, z_dim, **kwargs):
        super().__init__()
        self.z_dim = z_dim
        self.siren = siren(output_dim=4, z_dim=self.z_dim, input_dim=3, device=None)
        self.epoch = 0
        self.step = 0
        self.force_torch_path = False       # measurement / debugging hook
        self.train_integrate = 'torch'      # 'fused': fancy_integration of the autograd graph as the native op (csrc/integrate_ops.cu)

    def set_device(self, device):
        self.device = device
        self.siren.device = device
        self.generate_avg_frequencies()

    def generate_avg_frequencies(self):
        """generators.py:99-107"""
        z = torch.randn((10000, self.z_dim), device=self.siren.device)
        with torch.no_grad():
            frequencies, phase_shifts = self.siren.mapping_network(z)
        self.avg_frequencies = frequencies.mean(0, keepdim=True)
        self.avg_phase_shifts = phase_shifts.mean(0, keepdim=True)
        return self.avg_frequencies, self.avg_phase_shifts

    # ---- shared renderer: draws in the reference's order, then native kernels or the torch graph
    def _render(self, frequencies, phase_shifts, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev, h_mean,
                v_mean, hierarchical_sample, sample_dist, lock_view_dependence, clamp_mode, nerf_noise, white_back,
                last_back, need_grad):
        dev = frequencies.device
        _require_cuda(frequencies, "ImplicitGenerator3d")
        B, S, HW = frequencies.shape[0], num_steps, img_size * img_size
        nS = 2 * S if hierarchical_sample else S
        jitter_u = torch.rand((B, HW, S, 1), device=dev)[..., 0]                      # perturb_points
        # piGAN_lib/generators/volumetric_rendering.py:162-165: every mode it does not know -- None included, the default of
        # forward / staged_forward and what inverse_render.py passes -- means "use the mean camera" (the CIPS-3D surface
        # asserts instead, comm_utils.py:451-535)
        mode = sample_dist if sample_dist in _PIGAN_SAMPLE_DISTS else 'mean'
        c2w, pitch, yaw = sample_cam2world(B, dev, h_stddev, v_stddev, h_mean, v_mean, mode)
        noise_c = pdf_u = None
        if hierarchical_sample:
            noise_c = torch.randn((B, HW, S, 1), device=dev)[..., 0]                  # coarse fancy_integration
            pdf_u = torch.rand((B * HW, S), device=dev)                               # sample_pdf
        noise_f = torch.randn((B, HW, nS, 1), device=dev)[..., 0]                     # final fancy_integration
        if need_grad or self.force_torch_path or not self.siren.fused_supported():
            rgb, depth = self._render_torch(frequencies, phase_shifts, c2w, jitter_u, pdf_u, noise_c, noise_f, img_size, fov,
                                            ray_start, ray_end, S, hierarchical_sample, lock_view_dependence, clamp_mode,
                                            nerf_noise, white_back, last_back)
        else:
            out = ops.pigan_render(self.siren.kernel_weights(frequencies, phase_shifts), c2w, jitter_u, pdf_u, noise_c, noise_f,
                                   img_size=img_size, fov=fov, ray_start=ray_start, ray_end=ray_end, num_steps=S,
                                   hierarchical_sample=hierarchical_sample, clamp_mode=clamp_mode, noise_std=nerf_noise,
                                   white_back=white_back, last_back=last_back, lock_view=lock_view_dependence, want_depth=True)
            rgb, depth = out["rgb"], out["depth"]
        return rgb, depth, pitch, yaw

    def _render_torch(self, frequencies, phase_shifts, c2w, jitter_u, pdf_u, noise_c, noise_f, img_size, fov, ray_start, ray_end,
                      S, hierarchical_sample, lock_view, clamp_mode, nerf_noise, white_back, last_back):
        from .generator import _integrate, _sample_pdf, _torch_initial_rays
        dev = c2w.device
        dirs_cam, z_vals = _torch_initial_rays(img_size, ops.z_cam_from_fov(fov), ray_start, ray_end, S, dev)
        off = (jitter_u - 0.5) * (z_vals[1] - z_vals[0])
        z = z_vals[None, None, :] + off
        p_cam = dirs_cam[None, :, None, :] * z_vals[None, None, :, None] + off[..., None] * dirs_cam[None, :, None, :]
        Rm, t = c2w[:, :3, :3], c2w[:, :3, 3]
        pts = torch.einsum("bij,bnsj->bnsi", Rm, p_cam) + t[:, None, None, :]
        dirs_w = torch.einsum("bij,nj->bni", Rm, dirs_cam)
        B, N = z.shape[:2]
        dirs_exp = dirs_w[:, :, None, :].expand(-1, -1, S, -1).reshape(B, N * S, 3)
        if lock_view:
            dirs_exp = torch.zeros_like(dirs_exp)
            dirs_exp[..., -1] = -1
        field = self.siren.forward_with_frequencies_phase_shifts
        coarse = field(pts.reshape(B, N * S, 3), frequencies, phase_shifts, ray_directions=dirs_exp).reshape(B, N, S, 4)
        nc = noise_c * nerf_noise if noise_c is not None else None
        nf = noise_f * nerf_noise
        if hierarchical_sample:
            with torch.no_grad():
                _, w = _integrate(self.train_integrate, coarse, z, nc, clamp_mode, False, False, 3)
                w = w.reshape(B * N, S) + 1e-5
                zz = z.reshape(B * N, S)
                fz = _sample_pdf(self.train_integrate, 0.5 * (zz[:, :-1] + zz[:, 1:]), w[:, 1:-1], pdf_u).reshape(B, N, S)
                fpts = t[:, None, None, :] + dirs_w[:, :, None, :] * fz[..., None]
            fine = field(fpts.reshape(B, N * S, 3), frequencies, phase_shifts, ray_directions=dirs_exp).reshape(B, N, S, 4)
            if self.train_integrate == 'fused' and ops.integrate_merged_supported(fine, fz, coarse, z, nf):
                rgb, w, all_z = ops.integrate_merged(fine, fz, coarse, z, nf, clamp_mode, last_back, white_back)
                return rgb, torch.sum(w * all_z, -1)
            all_out = torch.cat([fine, coarse], dim=-2)
            all_z, ind = torch.sort(torch.cat([fz, z], dim=-1), dim=-1)
            all_out = torch.gather(all_out, -2, ind[..., None].expand(-1, -1, -1, 4))
        else:
            all_out, all_z = coarse, z
        rgb, w = _integrate(self.train_integrate, all_out, all_z, nf, clamp_mode, last_back, white_back, 3)
        depth = torch.sum(w * all_z, -1)
        return rgb, depth

    @staticmethod
    def _to_img(rgb, img_size):
        B = rgb.shape[0]
        return rgb.reshape(B, img_size, img_size, 3).permute(0, 3, 1, 2).contiguous() * 2 - 1

    def forward(self, z, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev, h_mean, v_mean, hierarchical_sample,
                sample_dist=None, lock_view_dependence=False, **kwargs):
        """generators.py:26-96 -> pixels (B,3,R,R) in [-1,1], cat([pitch, yaw], -1)"""
        frequencies, phase_shifts = self.siren.mapping_network(z)
        return self.forward_with_frequencies(frequencies, phase_shifts, img_size, fov, ray_start, ray_end, num_steps, h_stddev,
                                             v_stddev, h_mean, v_mean, hierarchical_sample, sample_dist=sample_dist,
                                             lock_view_dependence=lock_view_dependence, **kwargs)

    def forward_with_frequencies(self, frequencies, phase_shifts, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev,
                                 h_mean, v_mean, hierarchical_sample, sample_dist=None, lock_view_dependence=False, **kwargs):
        """generators.py:290-350"""
        need_grad = torch.is_grad_enabled() and (frequencies.requires_grad or any(p.requires_grad for p in self.siren.parameters()))
        rgb, _, pitch, yaw = self._render(frequencies, phase_shifts, img_size, fov, ray_start, ray_end, num_steps, h_stddev,
                                          v_stddev, h_mean, v_mean, hierarchical_sample, sample_dist, lock_view_dependence,
                                          kwargs['clamp_mode'], kwargs['nerf_noise'], kwargs.get('white_back', False),
                                          kwargs.get('last_back', False), need_grad)
        return self._to_img(rgb, img_size), torch.cat([pitch, yaw], -1)

    def staged_forward(self, z, img_size, fov, ray_start, ray_end, num_steps, h_stddev, v_stddev, h_mean, v_mean, psi=1,
                       lock_view_dependence=False, max_batch_size=50000, depth_map=False, near_clip=0, far_clip=2,
                       sample_dist=None, hierarchical_sample=False, **kwargs):
        """generators.py:110-204: inference with truncation.  max_batch_size (the reference's memory workaround) is
        accepted and ignored: the native renderer never materialises per-sample tensors beyond its own workspace.
        -> pixels (B,3,R,R) on the CPU, depth map (B,R,R) on the CPU"""
        self.generate_avg_frequencies()
        with torch.no_grad():
            raw_frequencies, raw_phase_shifts = self.siren.mapping_network(z)
            truncated_frequencies = self.avg_frequencies + psi * (raw_frequencies - self.avg_frequencies)
            truncated_phase_shifts = self.avg_phase_shifts + psi * (raw_phase_shifts - self.avg_phase_shifts)
        return self.staged_forward_with_frequencies(truncated_frequencies, truncated_phase_shifts, img_size, fov, ray_start,
                                                    ray_end, num_steps, h_stddev, v_stddev, h_mean, v_mean, psi=psi,
                                                    lock_view_dependence=lock_view_dependence, max_batch_size=max_batch_size,
                                                    sample_dist=sample_dist, hierarchical_sample=hierarchical_sample, **kwargs)

    def staged_forward_with_frequencies(self, truncated_frequencies, truncated_phase_shifts, img_size, fov, ray_start, ray_end,
                                        num_steps, h_stddev, v_stddev, h_mean, v_mean, psi=0.7, lock_view_dependence=False,
                                        max_batch_size=50000, depth_map=False, near_clip=0, far_clip=2, sample_dist=None,
                                        hierarchical_sample=False, **kwargs):
        """generators.py:207-288"""
        with torch.no_grad():
            rgb, depth, _, _ = self._render(truncated_frequencies, truncated_phase_shifts, img_size, fov, ray_start, ray_end,
                                            num_steps, h_stddev, v_stddev, h_mean, v_mean, hierarchical_sample, sample_dist,
                                            lock_view_dependence, kwargs['clamp_mode'], kwargs['nerf_noise'],
                                            kwargs.get('white_back', False), kwargs.get('last_back', False), False)
            B = rgb.shape[0]
            return self._to_img(rgb, img_size).cpu(), depth.reshape(B, img_size, img_size).contiguous().cpu()


# ---------------------------------------------------------------- function surface of piGAN_lib/generators/volumetric_rendering.py
def fancy_integration(rgb_sigma, z_vals, device, noise_std=0.5, last_back=False, white_back=False, clamp_mode=None, fill_mode=None):
    """volumetric_rendering.py:18-55 (rgb in channels 0-2, sigma in channel 3) on the native op (csrc/integrate_ops.cu):
    same arguments, the same torch.randn draw, (rgb_final (b, rays, 3), depth_final (b, rays, 1), weights (b, rays, s, 1)).
    An unknown clamp_mode raises (the reference raises too: `raise "Need to choose clamp mode"`)."""
    if clamp_mode not in ops.CLAMP_MODES:
        raise TypeError("Need to choose clamp mode")          # what `raise <str>` amounts to in the reference
    return ops.fancy_integration(rgb_sigma, z_vals, device=device, dim_rgb=3, noise_std=noise_std, last_back=last_back,
                                 white_back=white_back, clamp_mode=clamp_mode, fill_mode=fill_mode)


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5):
    """volumetric_rendering.py:205-240 (the same function as exp/pigan/pigan_utils.py:164-209) on c3d_sample_pdf"""
    return ops.sample_pdf(bins, weights, N_importance, det=det, eps=eps)
