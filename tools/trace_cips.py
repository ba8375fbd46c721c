"""Pipeline timeline of the CIPS kernel from in-kernel clock stamps (needs the -DC3D_TRACE build)."""
import os, sys, ctypes as C, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cips3d_b200
from cips3d_b200 import _lib, ops
from oracle import cips3d_oracle as O
dev = "cuda:0"
G = cips3d_b200.GeneratorNerfINR(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in O.G_CFG.items()}, device=dev).to(dev).eval()
G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234))
B = 4
zs = G.get_zs(B)
lib = _lib.load()
with torch.no_grad():
    style = G.mapping_network(**zs)
    x = torch.randn(B, 65536, 32, device=dev)
    for _ in range(3):
        G.inr_net(x, style)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 8192)()
    lib.c3d_debug_cips_trace(buf, 8192)        # reset
    G.inr_net(x, style)
    n = lib.c3d_debug_cips_trace(buf, 8192)
ev = [((v >> 56) & 0xFF, (v >> 40) & 0xFFFF, v & 0xFFFFFFFFFF) for v in buf[:n]]
t0 = min(e[2] for e in ev)
# tags: 1/2 issuer reached tile, 3/4 epilogue dependency satisfied, 5/6 weight tile landed (= MMA issue),
#       8 epilogue saw acc_ready[j] (a0 = layer << 8 | j), 9 epilogue warp finished chunk j (a0 = layer << 8 | warp << 2 | j)
rdy = collections.defaultdict(dict)
for tag, a0, t in ev:
    if tag == 8:
        rdy[a0 >> 8][a0 & 3] = t - t0
done = collections.defaultdict(lambda: collections.defaultdict(list))
for tag, a0, t in ev:
    if tag == 9:
        done[a0 >> 8][a0 & 3].append(t - t0)
issue = collections.defaultdict(dict)
for tag, a0, t in ev:
    if tag in (5, 6):
        issue[a0 >> 8][a0 & 0xFF] = (t - t0, tag - 5)
print("events", n)
layers = sorted(rdy)
prev = None
for l in layers:
    base = rdy[l].get(0, 0)
    line = f"layer {l:2d}: ready[0] at {base:8d}"
    if prev is not None: line += f" (+{base - prev:6d} since previous layer)"
    prev = base
    line += "; ready j: " + " ".join(f"{rdy[l].get(j, 0) - base:6d}" for j in range(4))
    line += "; chunk done (last warp): " + " ".join(f"{max(done[l][j]) - base:6d}" if done[l][j] else "     -" for j in range(4))
    print(line)
    if l + 1 in issue and 2 <= l <= 7:
        print("      layer %d MMA issue times rel. to ready[0] of layer %d (idx:clk/issuer): " % (l + 1, l) +
              " ".join(f"{i}:{issue[l + 1][i][0] - base}/{issue[l + 1][i][1]}" for i in sorted(issue[l + 1])))
