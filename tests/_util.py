"""Shared helpers for the parity tests (golden loading, oracle replay, error metrics)."""
import ast
import os

import numpy as np
import torch

from oracle import cips3d_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GEN_CASES = ("r16_synth", "r16_trained_noise", "r8_softplus_backs", "r8_nohier_s24")


def load_gen_case(name, dtype=torch.float32):
    g = np.load(os.path.join(GOLDEN, f"gen_{name}.npz"))
    kw = ast.literal_eval(str(g["kwargs_json"]))
    sd = O.synthetic_state_dict(O.generator_template(), seed=int(g["seed"]),
                                sigma_bias=float(g["sigma_bias"]), dtype=dtype)
    zs = dict(z_nerf=torch.from_numpy(g["z_nerf"]).to(dtype), z_inr=torch.from_numpy(g["z_inr"]).to(dtype))
    draws = {k[5:]: torch.from_numpy(g[k]).to(dtype) for k in g.files if k.startswith("draw_")}
    if "noise_c" not in draws:      # non-hierarchical: only 3 draws exist
        draws["noise_c"] = None
        draws["pdf_u"] = None
    meta = dict(B=int(g["B"]), img_size=int(g["img_size"]), nerf_noise=float(g["nerf_noise"]))
    ref = {k: torch.from_numpy(g[k]) for k in ("img", "pitch_yaw", "pixels_fea", "depth", "all_z", "coarse")}
    return sd, zs, draws, kw, meta, ref


def oracle_replay(name, dtype=torch.float32):
    sd, zs, draws, kw, meta, ref = load_gen_case(name, dtype)
    with torch.no_grad():
        img, py, r = O.generator_forward(sd, zs, draws, img_size=meta["img_size"],
                                         nerf_noise=meta["nerf_noise"], return_aux_img=True,
                                         return_all=True, **kw)
    return img, py, r, ref


def rel_err(a, b):
    """max |a-b| / max|b|  and  L2-relative."""
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item(), \
           ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
