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
    assert not t.is_cuda and t.is_contiguous()
    return t.data_ptr()


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
    try:
        yield pkg
    finally:
        L._lib, L.ptr, L.stream_ptr, ops.ptr, ops.stream_ptr, optim.ptr, optim.stream_ptr = saved
