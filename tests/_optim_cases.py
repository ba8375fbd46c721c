"""Shared body of the optimiser-tail parity tests (CPU emulation and GPU): FusedAdam.step(max_norm, ema, zero_grad)
against the oracle restatement of clip_grad_norm_ + torch.optim.Adam + the reference EMA (oracle/cips3d_oracle.py,
pinned bit for bit against torch and the reference's EMA class in test_oracle_vs_reference.py).

Tolerance: the update itself is elementwise fp32 in torch's operation order (agrees to the last bits); the global
gradient norm is a 10^5..10^7-term fp32 sum whose order differs from torch's (relative 1e-6), and it scales every
gradient when clipping is active -> 2e-5 relative on parameters / moments, written here once."""
import copy

import torch

from oracle import cips3d_oracle as O

TOL = 2e-5


class Net(torch.nn.Module):
    """odd sizes: vector tails, tensors smaller than a chunk, > 1 chunk, an unused parameter, a buffer"""

    def __init__(self, many=0):
        super().__init__()
        self.a = torch.nn.Linear(37, 53)
        self.b = torch.nn.Linear(53, 4099)
        self.c = torch.nn.Parameter(torch.randn(5))
        self.unused = torch.nn.Parameter(torch.randn(7, 3))
        self.register_buffer("buf", torch.randn(11))
        self.extra = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(3 + (i % 5))) for i in range(many)])

    def forward(self, x):
        y = (self.b(torch.tanh(self.a(x))).sum(-1) * self.c.sum()).mean()
        for e in self.extra:
            y = y + e.square().sum()
        return y


def max_rel(a, b):
    return max(((x.detach().cpu() - y.detach().cpu()).abs().max() / (y.detach().abs().max().cpu() + 1e-30)).item()
               for x, y in zip(a, b))


def run_case(pkg, device, betas, max_norm, steps=4, many=0, start_itr=2, zero_grad=True, seed=0):
    torch.manual_seed(seed)
    net = Net(many).to(device)
    net_ema = copy.deepcopy(net)
    opt = pkg.FusedAdam(net.parameters(), lr=2e-3, betas=betas)
    ema = pkg.EMA(net, net_ema, decay=0.999, start_itr=start_itr)
    names = [n for n, _ in net.named_parameters()]
    P = {n: p.detach().cpu().clone() for n, p in net.named_parameters()}
    M = {n: torch.zeros_like(p) for n, p in P.items()}
    V = {n: torch.zeros_like(p) for n, p in P.items()}
    E = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    worst = 0.0
    for it in range(steps):
        x = torch.randn(8, 37, device=device)
        net.zero_grad(set_to_none=True)
        net(x).backward()
        used = [n for n, p in net.named_parameters() if p.grad is not None]
        grads = [dict(net.named_parameters())[n].grad.detach().cpu().clone() for n in used]
        ema_on = it >= start_itr
        n_ref = O.clip_adam_ema_step([P[n] for n in used], grads, [M[n] for n in used], [V[n] for n in used],
                                     [E[n] for n in used], step=it + 1, lr=2e-3, betas=betas, max_norm=max_norm,
                                     ema_decay=0.999 if ema_on else None)
        if ema_on:      # entries without gradient / buffers: comm_model_utils.py:118-120
            for k in E:
                if k not in used:
                    src = P[k] if k in P else net.state_dict()[k].cpu()
                    E[k].copy_(E[k] * 0.999 + src * (1 - 0.999))
        n = opt.step(max_norm=max_norm, ema=ema, itr=it, zero_grad=zero_grad)
        if max_norm is not None:
            assert abs(float(n) - float(n_ref)) <= 1e-5 * float(n_ref)
        got = dict(net.named_parameters())
        worst = max(worst, max_rel([got[k] for k in names], [P[k] for k in names]))
        worst = max(worst, max_rel([opt.state[got[k]]["exp_avg"] for k in used], [M[k] for k in used]))
        worst = max(worst, max_rel([opt.state[got[k]]["exp_avg_sq"] for k in used], [V[k] for k in used]))
        worst = max(worst, max_rel([net_ema.state_dict()[k] for k in E], [E[k] for k in E]))
        if zero_grad:
            assert all(float(got[k].grad.abs().max()) == 0.0 for k in used)
    assert "unused" not in [k for k in names if got[k] in opt.state]          # Adam never creates state without a gradient
    return worst, opt, net
