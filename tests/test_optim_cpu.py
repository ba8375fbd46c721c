"""Optimiser tail (clip + Adam + EMA + zero_grad, csrc/optim_ops.cu) on the CPU emulation vs the oracle."""
import copy

import pytest
import torch

from _emu import emulated
from _optim_cases import TOL, run_case


@pytest.mark.parametrize("betas,max_norm", [((0.0, 0.999), 10.0), ((0.9, 0.99), 0.05), ((0.5, 0.999), None), ((0, 0.999), 10)])
def test_emu_fused_adam_ema_matches_oracle(betas, max_norm):
    """(0, 0.999), grad_clip 10 is the reference's configuration (ffhq_exp.yaml:159,171; integer 0 as in the yaml);
    0.05 makes the clip coefficient bite on every step; None skips the norm pass."""
    with emulated(async_mode=0) as pkg:
        worst, _, _ = run_case(pkg, "cpu", betas, max_norm)
    assert worst < TOL, worst


def test_emu_fused_adam_more_tensors_than_one_launch_holds():
    with emulated(async_mode=0) as pkg:
        worst, opt, _ = run_case(pkg, "cpu", (0.0, 0.999), 1.0, steps=2, many=2 * pkg.optim.OPT_MAX_TENSORS + 7, start_itr=0)
    assert worst < TOL, worst


def test_emu_fused_adam_state_dict_round_trips_with_torch_adam():
    """FusedAdam keeps torch.optim.Adam's state layout: a checkpoint of either loads into the other and the next
    step agrees."""
    with emulated(async_mode=0) as pkg:
        _, opt, net = run_case(pkg, "cpu", (0.9, 0.999), None, steps=3, zero_grad=False)
        t_opt = torch.optim.Adam(net.parameters(), lr=2e-3, betas=(0.9, 0.999))
        ckpt = copy.deepcopy(opt.state_dict())     # load_state_dict does not copy tensors that already fit
        t_opt.load_state_dict(copy.deepcopy(ckpt))
        before = [p.detach().clone() for p in net.parameters()]
        grads = [None if p.grad is None else p.grad.clone() for p in net.parameters()]
        t_opt.step()
        after_torch = [p.detach().clone() for p in net.parameters()]
        with torch.no_grad():
            for p, b in zip(net.parameters(), before):
                p.copy_(b)
        opt2 = pkg.FusedAdam(net.parameters(), lr=2e-3, betas=(0.9, 0.999))
        opt2.load_state_dict(copy.deepcopy(ckpt))
        for p, g in zip(net.parameters(), grads):
            p.grad = g
        opt2.step()
        for a, b in zip(net.parameters(), after_torch):
            assert (a.detach() - b).abs().max().item() <= 1e-6 * b.abs().max().item()


def test_emu_clip_grad_norm_function():
    torch.manual_seed(3)
    ps = [torch.nn.Parameter(torch.randn(n)) for n in (5, 4097, 100)]
    for p in ps:
        p.grad = torch.randn_like(p) * 3
    ref = [p.grad.clone() for p in ps]
    n_ref = torch.nn.utils.clip_grad_norm_([torch.nn.Parameter(torch.zeros_like(g)) for g in ref], 1.0)   # no grads -> 0
    assert float(n_ref) == 0.0
    refp = [torch.nn.Parameter(torch.zeros_like(g)) for g in ref]
    for q, g in zip(refp, ref):
        q.grad = g.clone()
    n_ref = torch.nn.utils.clip_grad_norm_(refp, 2.5)
    with emulated(async_mode=0) as pkg:
        n = pkg.optim.clip_grad_norm_(ps, 2.5)
    assert abs(float(n) - float(n_ref)) < 1e-5 * float(n_ref)
    for p, q in zip(ps, refp):
        assert (p.grad - q.grad).abs().max().item() < 1e-5 * q.grad.abs().max().item()


def test_fused_adam_rejects_what_the_kernels_do_not_implement():
    import cips3d_b200
    with pytest.raises(ValueError):
        cips3d_b200.FusedAdam([torch.nn.Parameter(torch.zeros(3))], weight_decay=0.1)
    with pytest.raises(ValueError):
        cips3d_b200.FusedAdam([torch.nn.Parameter(torch.zeros(3))], amsgrad=True)
    opt = cips3d_b200.FusedAdam([torch.nn.Parameter(torch.zeros(3))])
    opt.param_groups[0]["params"][0].grad = torch.ones(3)
    with pytest.raises(cips3d_b200._lib.C3dError):       # CPU tensors: there is no CPU path
        opt.step()
