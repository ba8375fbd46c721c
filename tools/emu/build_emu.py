"""Build libcips3d_b200_emu.so: the product's csrc/*.cu compiled as plain C++ (g++ -DC3D_EMU) against the
functional CPU emulation of CUDA/PTX in c3d_emu.h.  Test infrastructure only -- never loaded by the package."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cips-3d_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libcips3d_b200_emu.so")
CUDA_INC = os.environ.get("CUDA_INC", "/usr/local/cuda/include")
FLAGS = ["-std=c++17", "-O2", "-g", "-fPIC", "-DC3D_EMU", "-ffp-contract=off", "-fno-strict-aliasing", "-Wno-attributes",
         "-Wno-unknown-pragmas", "-Wno-unused-value", "-I", HERE, "-I", CUDA_INC]


def sanitize_flags():
    """C3D_EMU_SANITIZE=address,alignment (any -fsanitize= list): an instrumented build of the emulated library in its own
    directory -- a CPU stand-in for compute-sanitizer memcheck (out-of-bounds global / shared accesses land in ASan's
    redzones: run python with LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0) and for the
    misaligned-address faults x86 never raises (UBSan alignment: float4 / uint4 / uint64 accesses)."""
    san = os.environ.get("C3D_EMU_SANITIZE", "")
    if not san:
        return []
    flags = ["-fsanitize=" + san, "-fno-omit-frame-pointer"]
    if "alignment" in san:      # no libubsan in this image: a misaligned access executes ud2 (SIGILL) instead of reporting
        flags.append("-fsanitize-undefined-trap-on-error")
    return flags


def build(force=False, extra_defs=(), tag=""):
    """tag: build a variant (e.g. with a fault injected through extra_defs) into its own directory."""
    if sanitize_flags():
        tag = (tag + "_" if tag else "") + "san_" + os.environ["C3D_EMU_SANITIZE"].replace(",", "_")
    global_out = OUT_DIR + ("_" + tag if tag else "")
    return _build(global_out, force, extra_defs)


def _build(OUT_DIR, force, extra_defs):
    LIB = os.path.join(OUT_DIR, "libcips3d_b200_emu.so")
    os.makedirs(OUT_DIR, exist_ok=True)
    base_dir = globals()["OUT_DIR"]
    san = sanitize_flags()
    if extra_defs and OUT_DIR != base_dir and not san:
        _build(base_dir, False, ())        # a variant only recompiles the sources that mention one of its defines
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu"))) + [os.path.join(HERE, "emu_impl.cpp"), os.path.join(HERE, "emu_faults.cpp")]
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "*.h")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h")) + [os.path.abspath(__file__)]
    newest = max(os.path.getmtime(d) for d in deps)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) > newest:
        return LIB
    gxx = os.environ.get("CXX", "g++")
    if san:     # the sanitizer runtimes belong to the distribution's compiler ($CXX may point at one without them)
        gxx = os.environ.get("C3D_EMU_SAN_CXX", "/usr/bin/g++")

    def compile_one(src):
        obj = os.path.join(OUT_DIR, os.path.basename(src) + ".o")
        if not san and extra_defs and OUT_DIR != base_dir and not any(d.split("=")[0] in open(src).read() for d in extra_defs):
            return os.path.join(base_dir, os.path.basename(src) + ".o")
        cmd = [gxx] + FLAGS + san + [f"-D{d}" for d in extra_defs] + ["-x", "c++", "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"emu build failed for {src}:\n{r.stderr[-6000:]}")
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, srcs))
    r = subprocess.run([gxx, "-shared"] + san + ["-o", LIB] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emu link failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
