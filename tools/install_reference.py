#!/usr/bin/env python
"""Install the UNMODIFIED reference's hot-path Python modules into baseline/_ref/ (git-ignored, travels to the GPU box
with the gpurun snapshot) so that `bench.py --impl reference` and the reference-eager-on-B200 leg can run the reference's
own code where /root/reference does not exist.

The reference is pure Python on this path and has no setup.py / pyproject, so the contract's
`pip install --target baseline/_ref /root/reference` cannot work ("neither 'setup.py' nor 'pyproject.toml' found"); this
script is the committed recipe instead: a byte-for-byte copy of the package directories the generator / discriminator
import (SURVEY.md section 8(a) file list), nothing edited, nothing added to the git history.  The un-vendored `tl2`
dependency is satisfied at import time by tools/ref_shim.py (no arithmetic).

    python tools/install_reference.py            # no-op (exit 0) when /root/reference is absent
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("CIPS3D_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
# package directories (python files only) the hot path imports, relative to the reference root
DIRS = ["exp/cips3d/models", "exp/comm", "exp/comm/models", "exp/comm/op", "exp/pigan", "exp/pigan/models", "exp/dev/nerf_inr/models",
        "piGAN_lib/generators", "piGAN_lib/siren", "piGAN_lib/discriminators"]
INITS = ["exp", "exp/cips3d", "exp/dev", "exp/dev/nerf_inr", "piGAN_lib"]


def install(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "exp", "cips3d", "models")):
        if verbose:
            print(f"reference not found at {SRC}: nothing installed (baseline/_ref kept as is)")
        return None
    manifest = {}
    for d in DIRS:
        s = os.path.join(SRC, d)
        if not os.path.isdir(s):
            continue
        os.makedirs(os.path.join(DST, d), exist_ok=True)
        for f in sorted(os.listdir(s)):
            if f.endswith((".py", ".cu", ".cpp", ".h")) and os.path.isfile(os.path.join(s, f)):
                shutil.copyfile(os.path.join(s, f), os.path.join(DST, d, f))
                manifest[f"{d}/{f}"] = hashlib.sha256(open(os.path.join(s, f), "rb").read()).hexdigest()[:16]
    for d in INITS:          # namespace glue: the reference's own __init__.py where it has one, else an empty one
        os.makedirs(os.path.join(DST, d), exist_ok=True)
        s = os.path.join(SRC, d, "__init__.py")
        t = os.path.join(DST, d, "__init__.py")
        if os.path.isfile(s):
            shutil.copyfile(s, t)
        elif not os.path.exists(t):
            open(t, "w").close()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"installed {len(manifest)} reference files into {DST}")
    return DST


if __name__ == "__main__":
    install()
    sys.exit(0)
