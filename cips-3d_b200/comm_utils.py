"""Camera sampling and look-at matrices: host-side mirror of exp/comm/comm_utils.py:441-581,
617-641 (O(B) work on tiny tensors).  torch RNG calls keep the reference's shapes and order
(theta/yaw first, then phi/pitch) so poses match draw for draw."""
import math
import random

import torch


def normalize_vecs(v):
    return v / v.norm(dim=-1, keepdim=True)


def _truncated_unit_normal(bs, device):
    """First of 4 N(0,1) candidates inside (-2,2) per element (comm_utils.py:441-448)."""
    cand = torch.empty((bs, 1, 4), device=device).normal_()
    first_ok = ((cand < 2) & (cand > -2)).max(-1, keepdim=True)[1]
    return cand.gather(-1, first_ok).squeeze(-1)


def sample_camera_positions(device, bs=1, r=1, horizontal_stddev=1, vertical_stddev=1,
                            horizontal_mean=math.pi * 0.5, vertical_mean=math.pi * 0.5, mode='normal'):
    """-> points on the radius-r sphere (bs,3), pitch phi (bs,1) in [0,pi], yaw theta (bs,1).
    Modes: uniform | normal | gaussian | hybrid | truncated_gaussian | spherical_uniform | mean."""
    hs, vs, hm, vm = horizontal_stddev, vertical_stddev, horizontal_mean, vertical_mean
    unif = lambda: torch.rand((bs, 1), device=device) - 0.5
    gauss = lambda: torch.randn((bs, 1), device=device)
    if mode == 'hybrid':                                   # coin flip between wide-uniform and gaussian
        if random.random() < 0.5:
            theta = unif() * 2 * hs * 2 + hm
            phi = unif() * 2 * vs * 2 + vm
        else:
            theta = gauss() * hs + hm
            phi = gauss() * vs + vm
    elif mode == 'uniform':
        theta = unif() * 2 * hs + hm
        phi = unif() * 2 * vs + vm
    elif mode in ('normal', 'gaussian'):
        theta = gauss() * hs + hm
        phi = gauss() * vs + vm
    elif mode == 'truncated_gaussian':
        theta = _truncated_unit_normal(bs, device) * hs + hm
        phi = _truncated_unit_normal(bs, device) * vs + vm
    elif mode == 'spherical_uniform':
        theta = unif() * 2 * hs + hm
        v = (unif() * 2 * (vs / math.pi) + vm / math.pi).clamp(1e-5, 1 - 1e-5)
        phi = torch.arccos(1 - 2 * v)
    elif mode == 'mean':
        theta = torch.full((bs, 1), float(hm), device=device)
        phi = torch.full((bs, 1), float(vm), device=device)
    else:
        assert 0
    phi = phi.clamp(1e-5, math.pi - 1e-5)
    sin_phi = torch.sin(phi)
    pts = torch.zeros((bs, 3), device=device)
    pts[:, 0:1] = r * sin_phi * torch.cos(theta)
    pts[:, 2:3] = r * sin_phi * torch.sin(theta)
    pts[:, 1:2] = r * torch.cos(phi)
    return pts, phi, theta


def create_cam2world_matrix(forward_vector, origin, device=None, up_vector=None):
    """Look-at: columns (-left, up, -forward), then translate to origin (comm_utils.py:538-581)."""
    fwd = normalize_vecs(forward_vector)
    n = fwd.shape[0]
    if up_vector is None:      # (0,1,0) built on the device: torch.tensor(list, device='cuda') would synchronise the host
        up_vector = torch.zeros_like(fwd)
        up_vector[:, 1] = 1
    left = normalize_vecs(torch.cross(up_vector, fwd, dim=-1))
    up = normalize_vecs(torch.cross(fwd, left, dim=-1))
    rot = torch.eye(4, device=device).unsqueeze(0).repeat(n, 1, 1)
    rot[:, :3, :3] = torch.stack((-left, up, -fwd), dim=-1)
    trans = torch.eye(4, device=device).unsqueeze(0).repeat(n, 1, 1)
    trans[:, :3, 3] = origin
    return trans @ rot


def sample_cam2world(bs, device, h_stddev, v_stddev, h_mean, v_mean, mode, camera_pos=None,
                     camera_lookup=None, up_vector=None):
    """Camera part of transform_sampled_points (comm_utils.py:617-641): -> (bs,4,4), pitch, yaw."""
    if camera_pos is None or camera_lookup is None:
        origin, pitch, yaw = sample_camera_positions(
            device, bs=bs, r=1, horizontal_stddev=h_stddev, vertical_stddev=v_stddev,
            horizontal_mean=h_mean, vertical_mean=v_mean, mode=mode)
        look = -origin
    else:
        origin, look = camera_pos, camera_lookup
        pitch = yaw = torch.zeros(bs, 1, device=device)
    return create_cam2world_matrix(look, origin, device=device, up_vector=up_vector), pitch, yaw


def scatter_points(idx_grad, points_grad, idx_no_grad, points_no_grad, num_points):
    """Inverse of the grad / no-grad pixel split (comm_utils.py:240-258)."""
    out = points_grad.new_zeros(points_grad.shape[0], num_points, points_grad.shape[-1])
    out = out.index_copy(1, idx_grad, points_grad)
    return out.index_copy(1, idx_no_grad, points_no_grad.to(out.dtype))


def get_world_points_and_direction(*args, **kwargs):
    """comm_utils.py:682-763 (reference signature); implemented next to the other ray helpers in generator.py"""
    from .generator import get_world_points_and_direction as impl
    return impl(*args, **kwargs)


def gather_points(points, idx_grad):
    """comm_utils.py:262-282: rays `idx_grad` of (b, n, c) or (b, n, s, c)"""
    assert points.dim() in (3, 4)
    return points.index_select(1, idx_grad)


# ---------------------------------------------------------------- host-side helpers of the inference / finetune scripts
def inr_layer_swapping(swapped_net, target_net, gamma_target, swapped_layers, verbose=True):
    """comm_utils.py:28-51 (finetune / interpolation demos): blend the CIPS blocks `network.<name>` and `to_rgbs.<name>` of
    `swapped_net` towards `target_net` in place: p <- (1 - gamma) p + gamma p_target."""
    import logging
    prefixes = tuple(f"{kind}.{name}" for name in swapped_layers for kind in ("network", "to_rgbs"))
    if verbose:
        logging.getLogger('tl').info(f"Layer swapping: {prefixes}")
    target = target_net.state_dict()
    with torch.no_grad():
        for name, p in swapped_net.named_parameters():
            if prefixes and name.startswith(prefixes):
                p.copy_(p * (1 - gamma_target) + target[name].to(p.device) * gamma_target)


def get_yaw_pitch_by_xyz(x, y, z):
    """comm_utils.py:82-85"""
    return math.atan2(z, x), math.atan2(math.sqrt(x ** 2 + z ** 2), y)


def _trajectory(xyz):
    """(xyz float32 (n, 3), lookup = -xyz, yaws, pitchs float64) -- the return convention of comm_utils.py:87-110, 219-238"""
    import numpy as np
    xyz = np.asarray(xyz, dtype=np.float32)
    # the rows are iterated as float32 scalars, as the reference does: x ** 2 + z ** 2 is then rounded in float32
    yp = np.array([get_yaw_pitch_by_xyz(x, y, z) for x, y, z in xyz], dtype=np.float64).reshape(-1, 2)
    return xyz, -xyz, yp[:, 0].copy(), yp[:, 1].copy()


def get_circle_camera_pos_and_lookup(r=1, alpha=3.141592 / 6, num_samples=36, periods=2):
    """comm_utils.py:87-110: `periods` turns around the z axis on the cone of half-angle alpha (camera path of the videos)"""
    import numpy as np
    n = num_samples * periods
    xyz = np.zeros((n, 3), dtype=np.float32)
    rho = r * math.sin(alpha)
    for i, t in enumerate(np.linspace(1, 0, n)):
        beta = t * 2 * math.pi * periods
        xyz[i] = (rho * math.cos(beta), rho * math.sin(beta), r * math.cos(alpha))
    return _trajectory(xyz)


def get_yaw_camera_pos_and_lookup(r=1, num_samples=36):
    """comm_utils.py:219-238: a yaw sweep in the y = 0 plane, theta from 1 to pi - 1"""
    import numpy as np
    xyz = np.zeros((num_samples, 3), dtype=np.float32)
    for i, theta in enumerate(np.linspace(1, math.pi - 1, num_samples)):
        xyz[i] = (r * math.cos(theta), 0, r * math.sin(theta))
    return _trajectory(xyz)


# ---------------------------------------------------------------- camera-space rays as functions (rows R2-R3 of SURVEY 8(a))
def get_initial_rays_trig(bs, num_steps, fov, resolution, ray_start, ray_end, device):
    """comm_utils.py:365-413 (reference signature): camera-space sample points (bs, H*W, num_steps, 3), depths
    (bs, H*W, num_steps, 1) and unit ray directions (bs, H*W, 3); rays row-major, x in [-1, 1], y from +1 down to -1."""
    W, H = resolution
    x = torch.linspace(-1, 1, W, device=device)[None, :].expand(H, W).reshape(-1)
    y = torch.linspace(1, -1, H, device=device)[:, None].expand(H, W).reshape(-1)
    z = -torch.ones_like(x) / math.tan((2 * math.pi * fov / 360) / 2)
    rays_d_cam = normalize_vecs(torch.stack([x, y, z], -1))
    z_vals = torch.linspace(ray_start, ray_end, num_steps, device=device).reshape(1, num_steps, 1).repeat(W * H, 1, 1)
    points = rays_d_cam[:, None, :] * z_vals
    rep = lambda t: t.unsqueeze(0).repeat(bs, *([1] * t.dim()))            # noqa: E731
    return rep(points), rep(z_vals), rep(rays_d_cam)


def perturb_points(points, z_vals, ray_directions, device):
    """comm_utils.py:416-437: one uniform jitter per sample, within the spacing of the first two depths; same draw
    (torch.rand of z_vals' shape) as the reference."""
    spacing = z_vals[:, :, 1:2, :] - z_vals[:, :, 0:1, :]
    offset = (torch.rand(z_vals.shape, device=device) - 0.5) * spacing
    return points + offset * ray_directions.unsqueeze(2), z_vals + offset


def transform_sampled_points(points, z_vals, ray_directions, device, h_stddev=1, v_stddev=1, h_mean=math.pi * 0.5,
                             v_mean=math.pi * 0.5, mode='normal', camera_pos=None, camera_lookup=None, up_vector=None):
    """comm_utils.py:584-679 (reference signature): jitter the samples, draw (or take) the camera, map camera space to world
    space.  -> (points (bs, n, s, 3), z_vals (bs, n, s, 1), ray directions (bs, n, 3), ray origins (bs, n, 3), pitch, yaw)."""
    bs, n, s, _ = points.shape
    points, z_vals = perturb_points(points, z_vals, ray_directions, device)
    c2w, pitch, yaw = sample_cam2world(bs, device, h_stddev, v_stddev, h_mean, v_mean, mode, camera_pos=camera_pos,
                                       camera_lookup=camera_lookup, up_vector=up_vector)
    hom = torch.cat([points, torch.ones_like(points[..., :1])], -1).reshape(bs, n * s, 4)
    world = torch.bmm(c2w, hom.transpose(1, 2)).transpose(1, 2).reshape(bs, n, s, 4)[..., :3]
    dirs = torch.bmm(c2w[:, :3, :3], ray_directions.reshape(bs, n, 3).transpose(1, 2)).transpose(1, 2).reshape(bs, n, 3)
    origins = c2w[:, :3, 3].unsqueeze(1).expand(bs, n, 3)               # cam2world applied to the camera-space origin
    return world, z_vals, dirs, origins, pitch, yaw
