"""Run the package's host layer against the CPU emulation build of the kernels (tools/emu).

Test infrastructure only.  `emulated(...)` temporarily binds cips3d_b200's ctypes layer to
libcips3d_b200_emu.so (the product sources compiled with g++ -DC3D_EMU: every CUDA thread a fiber, mbarrier /
TMEM / tcgen05.mma / bulk copies emulated, asynchronous ops reordered under a seed) and lets CPU tensors through,
so the real host code + the real kernel code run on the CPU.  The product loader itself refuses this library
(`_lib.load`, test_emu_cpu.py::test_product_loader_refuses_emulation_build)."""
import contextlib
import ctypes as C
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("c3d_build_emu", os.path.join(ROOT, "tools", "emu", "build_emu.py"))
build_emu = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(build_emu)

_cdll = None


def emu_lib():
    global _cdll
    if _cdll is None:
        path = build_emu.build()
        _cdll = C.CDLL(path)
        _cdll.c3d_emu_configure.argtypes = [C.c_int, C.c_ulonglong, C.c_int, C.c_int]
        _cdll.c3d_emulated.restype = C.c_int
    return _cdll


def _cpu_ptr(t):
    if t is None:
        return None
    assert not t.is_cuda and (t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)))
    return t.data_ptr()


_POISON = os.environ.get("C3D_EMU_POISON", "1") != "0"


def _poisoned(alloc):
    """torch.empty / empty_like inside the emulation context return POISONED memory (NaN for floating types, 0x7B bytes
    otherwise): fresh CPU pages are zero, a GPU caching allocator's are not -- a kernel that reads workspace or output
    memory it never wrote (relying on zeros) must fail its parity check here, as compute-sanitizer's initcheck would flag it."""
    def wrapper(*a, **k):
        t = alloc(*a, **k)
        if t.device.type == "cpu" and t.numel() and not t.is_quantized and t.layout == torch.strided:
            with torch.no_grad():
                if t.is_floating_point():
                    t.fill_(float("nan"))
                elif t.dtype in (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64):
                    t.fill_(0x7B)
        return t
    return wrapper


@contextlib.contextmanager
def emulated(async_mode=2, seed=1, preempt_permille=20, sms=2):
    """async_mode: 0 = asynchronous ops complete at issue, 1 = as late as possible, 2 = random (seeded)."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import cips3d_b200 as pkg
    lib = emu_lib()
    lib.c3d_emu_configure(async_mode, seed, preempt_permille, sms)
    L, ops, optim = pkg._lib, pkg.ops, pkg.optim
    saved = (L._lib, L.ptr, L.stream_ptr, ops.ptr, ops.stream_ptr, optim.ptr, optim.stream_ptr)
    L._lib = L.bind(lib)
    L.ptr = ops.ptr = optim.ptr = _cpu_ptr
    L.stream_ptr = ops.stream_ptr = optim.stream_ptr = lambda: None
    allocs = (torch.empty, torch.empty_like)
    if _POISON:
        torch.empty, torch.empty_like = _poisoned(torch.empty), _poisoned(torch.empty_like)
    try:
        yield pkg
    finally:
        torch.empty, torch.empty_like = allocs
        L._lib, L.ptr, L.stream_ptr, ops.ptr, ops.stream_ptr, optim.ptr, optim.stream_ptr = saved
