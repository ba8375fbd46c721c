"""Optimiser tail of the reference's training step on the fused multi-tensor kernels (csrc/optim_ops.cu).

The reference does, per optimiser and per step (exp/cips3d/scripts/train.py:417-438 for D, :468-491 for G):

    total_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip)
    optimizer.step()                      # torch.optim.Adam(lr, betas, weight_decay=0), train.py:173-187
    optimizer.zero_grad()
    ema_model.update(itr=step, source_dict=generator.state_dict())     # G only, comm_model_utils.py:99-121

which is ~8 elementwise launches per parameter tensor (119 tensors in G, 150 in D).  Here:

    total_norm = optimizer.step(max_norm=grad_clip, ema=ema_model, itr=step, zero_grad=True)

is two launches in all: one reads every gradient for the global norm, one applies clip + Adam + EMA (+ zero_grad).
`FusedAdam` keeps torch.optim.Adam's constructor, param_groups and state layout ('step', 'exp_avg',
'exp_avg_sq'), so optimiser checkpoints are interchangeable; `EMA` keeps the reference class's methods.
No CPU path: the kernels run from libcips3d_b200.so only."""
import ctypes as C

import torch

from . import _lib
from ._lib import check, load, ptr, stream_ptr

OPT_MAX_TENSORS = 160


class OptTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("ema", C.c_void_p), ("n", C.c_int64)]


class AdamParams(C.Structure):
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("ema_decay", C.c_double), ("step", C.c_int64), ("zero_grad", C.c_int32)]


def _bind(lib):
    if getattr(lib, "_c3d_optim_bound", False):
        return lib
    lib.c3d_optim_workspace_bytes.restype = C.c_size_t
    lib.c3d_grad_norm.argtypes = [C.POINTER(OptTensor), C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t,
                                  C.c_void_p]
    lib.c3d_adam_ema_step.argtypes = [C.POINTER(OptTensor), C.c_int32, C.POINTER(AdamParams), C.c_void_p, C.c_void_p]
    lib.c3d_ema_update.argtypes = [C.POINTER(OptTensor), C.c_int32, C.c_double, C.c_void_p]
    lib._c3d_optim_bound = True
    return lib


def _f32_flat_ok(t, what):
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise _lib.C3dError(f"{what}: fused optimiser kernels need contiguous float32 tensors, got {t.dtype} "
                            f"contiguous={t.is_contiguous()}")
    return t


def _table(rows):
    arr = (OptTensor * max(len(rows), 1))()
    for i, (p, g, m, v, e) in enumerate(rows):
        arr[i] = OptTensor(ptr(p), ptr(g), ptr(m), ptr(v), ptr(e), p.numel() if p is not None else g.numel())
    return arr


def grad_norm(grads, max_norm=0.0):
    """(total_norm, clip_coef) as a 2-element device tensor: ||grads||_2 and min(1, max_norm / (norm + 1e-6))
    (torch.nn.utils.clip_grad_norm_'s coefficient; 1 when max_norm <= 0).  One pass over the gradients."""
    lib = _bind(load())
    grads = [_f32_flat_ok(g, "grad") for g in grads if g.numel() > 0]
    dev = grads[0].device if grads else torch.device("cuda")
    out = torch.empty(2, device=dev, dtype=torch.float32)
    wsb = lib.c3d_optim_workspace_bytes()
    ws = torch.empty(wsb // 4, device=dev, dtype=torch.float32)
    rows = [(None, g, None, None, None) for g in grads]
    check(lib.c3d_grad_norm(_table(rows), len(rows), float(max_norm), ptr(out), ptr(ws), wsb, stream_ptr()), "c3d_grad_norm")
    return out


def clip_grad_norm_(parameters, max_norm):
    """torch.nn.utils.clip_grad_norm_(parameters, max_norm) with the norm from one fused launch (train.py:420, 474).
    Scales the gradients in place and returns the total norm.  Prefer FusedAdam.step(max_norm=...), which folds the
    scaling into the update and never rewrites the gradients."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    nc = grad_norm(grads, max_norm)
    torch._foreach_mul_(grads, nc[1])
    return nc[0]


class EMA(object):
    """exp/comm/comm_model_utils.py:54-121 (same constructor, `update`, `update_target_dict`)."""

    def __init__(self, source, target, decay=0.9999, start_itr=0):
        self.source = source
        self.target = target
        self.decay = decay
        self.start_itr = start_itr
        self.source_dict = self.source.state_dict()
        self.target_dict = self.target.state_dict()
        self.update_target_dict(source_state_dict=self.source_dict)

    def update_target_dict(self, source_state_dict):
        with torch.no_grad():
            for key in source_state_dict:
                self.target_dict[key].data.copy_(source_state_dict[key].data)

    def active(self, itr=None):
        return not (itr is not None and itr < self.start_itr)          # comm_model_utils.py:108-110

    def target_of(self):
        """{data_ptr of a source tensor: its EMA tensor} (state_dict tensors alias the module's tensors)."""
        return {v.data_ptr(): self.target_dict[k] for k, v in self.source_dict.items()}

    def update(self, itr=None, source_dict=None, skip_ptrs=()):
        """target = target * decay + source * (1 - decay) for every state_dict entry (one launch for the fp32 ones).
        skip_ptrs: source tensors whose EMA was already advanced inside FusedAdam.step."""
        if not self.active(itr):
            return
        if source_dict is None:
            source_dict = self.source_dict
        lib = _bind(load())
        rows = []
        with torch.no_grad():
            for key, src in source_dict.items():
                if src.data_ptr() in skip_ptrs or src.numel() == 0:
                    continue
                tgt = self.target_dict[key]
                if src.dtype == torch.float32 and tgt.dtype == torch.float32 and src.is_contiguous() and tgt.is_contiguous():
                    rows.append((src.data, None, None, None, tgt.data))
                else:       # integer buffers etc.: the reference's expression verbatim
                    tgt.data.copy_(tgt.data * self.decay + src.data * (1 - self.decay))
        if rows:
            check(lib.c3d_ema_update(_table(rows), len(rows), float(self.decay), stream_ptr()), "c3d_ema_update")


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay=0) -- the optimiser train.py:173-187 builds -- with
    clip_grad_norm_, the EMA of the weights and zero_grad folded into `step`.  State per parameter: 'step' (0-dim
    fp32 tensor as in torch), 'exp_avg', 'exp_avg_sq' -> state_dict()s load into torch.optim.Adam and back."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise ValueError("FusedAdam implements the reference's configuration: weight_decay=0, amsgrad=False")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid betas: {betas}")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None, *, max_norm=None, ema=None, itr=None, zero_grad=False):
        """One optimisation step.  max_norm: clip the global gradient norm (over ALL groups, like
        clip_grad_norm_(model.parameters(), max_norm)); ema: an `EMA` whose targets are advanced with the updated
        weights (entries that are not optimised here -- buffers, parameters without gradient -- included);
        itr: iteration passed to the EMA's start_itr gate; zero_grad: write zeros into the gradients afterwards
        (= optimizer.zero_grad(set_to_none=False)).  Returns the total gradient norm (0-dim tensor) when max_norm
        is given, else None."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _bind(load())
        ema_on = ema is not None and ema.active(itr)
        ema_map = ema.target_of() if ema_on else {}
        with_grad = [(g, p) for g in self.param_groups for p in g["params"] if p.grad is not None and p.numel() > 0]
        for _, p in with_grad:
            _f32_flat_ok(p, "param")
            _f32_flat_ok(p.grad, "grad")
            if p.grad.is_sparse:
                raise RuntimeError("Adam does not support sparse gradients")
        total_norm = clip_ptr = None
        if max_norm is not None and with_grad:
            nc = grad_norm([p.grad for _, p in with_grad], max_norm)
            total_norm, clip_ptr = nc[0], nc[1:].data_ptr()
            self._keep = nc
        elif max_norm is not None:
            total_norm = torch.tensor(0.0)
        done_ptrs = set()
        # one launch per (group, step value): hyper-parameters are per group, the bias correction per step count
        buckets = {}
        for gi, (group, p) in enumerate(with_grad):
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            step = int(float(st["step"])) + 1
            st["step"] = st["step"] + 1 if torch.is_tensor(st["step"]) else step
            buckets.setdefault((id(group), step), (group, step, []))[2].append(p)
        for group, step, plist in buckets.values():
            rows = []
            for p in plist:
                st = self.state[p]
                e = ema_map.get(p.data_ptr())
                if e is not None:
                    _f32_flat_ok(e, "ema target")
                    done_ptrs.add(p.data_ptr())
                rows.append((p.data, p.grad, _f32_flat_ok(st["exp_avg"], "exp_avg"), _f32_flat_ok(st["exp_avg_sq"], "exp_avg_sq"),
                             e.data if e is not None else None))
            hp = AdamParams(lr=float(group["lr"]), beta1=float(group["betas"][0]), beta2=float(group["betas"][1]),
                            eps=float(group["eps"]), ema_decay=float(ema.decay) if ema_on else -1.0, step=step,
                            zero_grad=int(bool(zero_grad)))
            check(lib.c3d_adam_ema_step(_table(rows), len(rows), C.byref(hp), clip_ptr, stream_ptr()), "c3d_adam_ema_step")
        if ema_on:       # state_dict entries this optimiser did not touch (buffers, parameters without gradient)
            ema.update(itr=itr, source_dict=ema.source.state_dict(), skip_ptrs=done_ptrs)
        return total_norm if max_norm is not None else loss
