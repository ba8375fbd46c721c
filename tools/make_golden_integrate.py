"""Generate tests/golden/fancy_integration.npz from the UNMODIFIED reference function exp/pigan/pigan_utils.py:212-273
and, for the pdf_* cases, from sample_pdf (pigan_utils.py:164-209)
(needs /root/reference; imported through tools/ref_shim.py):

    python tools/make_golden_integrate.py

Per case: inputs from a numpy PCG64 stream (platform-independent bytes), the noise the reference drew (captured by seeding
torch identically and re-drawing), the three outputs of the real function and the gradient of sum(rgb_final * d_rgb) w.r.t.
rgb_sigma from the real function's autograd graph.  The merged cases additionally run the reference's own
cat + sort + gather (generator.py:1733-1738) on two unsorted halves before the call (no equal depths across the halves:
torch.sort is not stable there, so the reference itself leaves the order of a tie unspecified)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fancy_integration.npz")
# name: (batch, rays, samples, dim_rgb, clamp, last_back, white_back, noise_std, merged)
CASES = {
    "relu_lastback": (2, 13, 24, 32, "relu", True, False, 0.0, False),
    "softplus_white_noise": (2, 11, 24, 32, "softplus", False, True, 0.5, False),
    "coarse_relu_noise": (1, 17, 12, 32, "relu", False, False, 0.7, False),
    "pigan_rgb3_backs": (2, 9, 24, 3, "relu", True, True, 0.3, False),
    "merged_relu_lastback": (2, 13, 24, 32, "relu", True, False, 0.0, True),
    "merged_softplus_noise": (1, 11, 24, 32, "softplus", True, True, 0.4, True),
}


def main():
    ref_shim.install()
    from exp.pigan import pigan_utils as ref_utils
    out = {}
    for idx, (name, (B, N, T, Cn, clamp, lb, wb, nstd, merged)) in enumerate(CASES.items()):
        rng = np.random.Generator(np.random.PCG64(4200 + idx))
        f32 = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))       # noqa: E731
        uni = lambda *s: torch.from_numpy(rng.random(s).astype(np.float32))                # noqa: E731
        rs = f32(B, N, T, Cn + 1)
        opaque = uni(B, N, 1) < 0.3
        rs[..., Cn] = (rs[..., Cn] + 0.3) * torch.where(opaque, 400.0, 8.0)
        d_rgb = f32(B, N, Cn)
        leaf = rs.clone().requires_grad_()
        if merged:                      # halves as the field produces them; the reference sorts the concatenation
            S = T // 2
            z_fine, z_coarse = 0.88 + 0.24 * uni(B, N, S, 1), torch.sort(0.88 + 0.24 * uni(B, N, S, 1), -2).values
            all_z = torch.cat([z_fine, z_coarse], dim=-2)                                   # generator.py:1733-1734
            _, indices = torch.sort(all_z, dim=-2)                                          # generator.py:1735
            z = torch.gather(all_z, -2, indices)                                            # generator.py:1736
            src = torch.gather(leaf, -2, indices.expand(-1, -1, -1, Cn + 1))                # generator.py:1738
            out[f"{name}/z_fine"], out[f"{name}/z_coarse"] = z_fine.numpy(), z_coarse.numpy()
        else:
            z = torch.sort(0.88 + 0.24 * uni(B, N, T, 1), -2).values
            src = leaf
        torch.manual_seed(900 + idx)
        rgb, depth, w = ref_utils.fancy_integration(src, z, device="cpu", dim_rgb=Cn, noise_std=nstd, last_back=lb,
                                                    white_back=wb, clamp_mode=clamp)
        (grad,) = torch.autograd.grad(rgb, leaf, d_rgb)
        torch.manual_seed(900 + idx)
        noise = torch.randn((B, N, T, 1)) * nstd                                            # pigan_utils.py:246
        for k, v in dict(rgb_sigma=rs, z=z, d_rgb=d_rgb, noise=noise, rgb=rgb.detach(), depth=depth.detach(), weights=w.detach(),
                         grad=grad).items():
            out[f"{name}/{k}"] = v.numpy()
        out[f"{name}/cfg"] = np.array([B, N, T, Cn, int(clamp == "softplus"), int(lb), int(wb), int(merged)], dtype=np.int64)
        out[f"{name}/noise_std"] = np.float32(nstd)
    # sample_pdf (pigan_utils.py:164-209): weights as at its call site (coarse weights + 1e-5, inner slice), bins = midpoints
    for idx, (name, (rays, n, k, det)) in enumerate({"pdf_s12": (37, 10, 12, False), "pdf_n32_k40": (9, 32, 40, False),
                                                      "pdf_det": (11, 10, 12, True), "pdf_peaky": (23, 22, 24, False)}.items()):
        rng = np.random.Generator(np.random.PCG64(7700 + idx))
        w = torch.from_numpy(rng.random((rays, n)).astype(np.float32)) ** (8 if "peaky" in name else 1) + 1e-5
        if "peaky" in name:
            w[:, ::3] = 1e-5                                                                  # empty bins (alpha = 0 samples)
        edges = torch.sort(torch.from_numpy((0.88 + 0.24 * rng.random((rays, n + 2))).astype(np.float32)), -1).values
        bins = 0.5 * (edges[:, :-1] + edges[:, 1:])
        torch.manual_seed(500 + idx)
        samples = ref_utils.sample_pdf(bins, w, k, det=det)
        torch.manual_seed(500 + idx)
        u = torch.linspace(0, 1, k).expand(rays, k).contiguous() if det else torch.rand(rays, k)
        for key, v in dict(bins=bins, weights=w, u=u, samples=samples).items():
            out[f"{name}/{key}"] = v.numpy()
        out[f"{name}/cfg"] = np.array([rays, n, k, int(det)], dtype=np.int64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
