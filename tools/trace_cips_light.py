"""Merged timeline of the CIPS kernel from the "light" trace build (tools/build_trace_lib.sh 1 -DC3D_CIPS_ABLATE): block 0, tile
iteration 1, a few layers -- who waits for whom at the layer boundary, with the traced CTA close to its untraced timing.
    C3D_LIB_PATH=.../libcips3d_b200_trace_light.so [C3D_CIPS_ABLATE=7] [C3D_CIPS_PAIR=1] python tools/trace_cips_light.py [B] [first layer] [layers]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cips3d_b200
from cips3d_b200 import _lib
from oracle import cips3d_oracle as O
dev = "cuda:0"
G = cips3d_b200.GeneratorNerfINR(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in O.G_CFG.items()}, device=dev).to(dev).eval()
G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L0 = int(sys.argv[2]) if len(sys.argv) > 2 else 4
NL = int(sys.argv[3]) if len(sys.argv) > 3 else 3
zs = G.get_zs(B)
lib = _lib.load()
lib.c3d_debug_cips_trace.argtypes = [C.c_void_p, C.c_int]
CAP = 2 * 40 * 3072
buf = (C.c_ulonglong * CAP)()
with torch.no_grad():
    style = G.mapping_network(**zs)
    x = torch.randn(B, 65536, 32, device=dev)
    for _ in range(3):
        G.inr_net(x, style)
    torch.cuda.synchronize()
    lib.c3d_debug_cips_trace(buf, CAP)        # reset
    G.inr_net(x, style)
    n = lib.c3d_debug_cips_trace(buf, CAP)
NAMES = {1: "issuer0 reached tile", 2: "issuer1 reached tile", 3: "issuer0 epilogue dep ok", 4: "issuer1 epilogue dep ok",
         5: "issuer0 weights landed -> issue", 6: "issuer1 weights landed -> issue", 8: "epilogue: acc_ready seen, chunk",
         9: "epilogue: chunk done, arrived", 10: "producer: stage free, load tile", 11: "peer relay tile"}
ev = []
for i in range(n):
    w, v = buf[2 * i], buf[2 * i + 1]
    blk, warp, tag, a0, t = int(w >> 8), int(w & 0xFF), int((v >> 56) & 0xFF), int((v >> 40) & 0xFFFF), int(v & 0xFFFFFFFFFF)
    ev.append((blk, t, warp, tag, a0 >> 8, a0 & 0xFF))
print("events", n)
for blk in (0, 1):      # setmaxnreg.inc wait of the epilogue warps (tags 20 / 21: before / after)
    t20 = {w: t for b, t, w, tag, l, i in ev if b == blk and tag == 20}
    t21 = {w: t for b, t, w, tag, l, i in ev if b == blk and tag == 21}
    if t20 and t21:
        print(f"block {blk}: setmaxnreg.inc wait per epilogue warp [clk]: " + " ".join(str(t21[w] - t20[w]) for w in sorted(t20) if w in t21))
for blk in (0, 1):
    rows = sorted(e for e in ev if e[0] == blk and L0 <= e[4] < L0 + NL)
    if not rows:
        continue
    t0 = rows[0][1]
    print(f"block {blk}: layers {L0}..{L0 + NL - 1}   [clk since the first stamp | +delta | warp | event | layer | tile or chunk]")
    prev = t0
    for _, t, warp, tag, l, idx in rows:
        if tag == 9:
            idx &= 3
        print(f"  {t - t0:8d} +{t - prev:6d}  w{warp:<2d} {NAMES.get(tag, str(tag)):34s} L{l:<2d} {idx}")
        prev = t
    # per-layer period: acc_ready[0] seen to acc_ready[0] seen
    seen = sorted((l, t) for _, t, warp, tag, l, idx in sorted(e for e in ev if e[0] == blk) if tag == 8 and idx == 0)
    per = [b[1] - a[1] for a, b in zip(seen, seen[1:]) if b[0] == a[0] + 1]
    if per:
        print(f"  layer period (acc_ready[0] to acc_ready[0]) [clk]: " + " ".join(str(p) for p in per))
