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
names = {1: "mma0 reach", 2: "mma1 reach", 3: "mma0 epi-ok", 4: "mma1 epi-ok", 5: "mma0 w-ok", 6: "mma1 w-ok", 8: "acc_full", 9: "chunk-done"}
# per layer summary
acc = {}
for tag, a0, t in ev:
    if tag == 8:
        acc[a0 >> 8] = t - t0
layers = sorted(acc)
print("events", n)
for l in layers[2:8]:
    nxt = acc.get(l + 1)
    chunk = collections.defaultdict(list)
    for tag, a0, t in ev:
        if tag == 9 and (a0 >> 8) == l:
            chunk[a0 & 3].append(t - t0 - acc[l])
    line = f"layer {l:2d}: acc_full at {acc[l]:8d}; epilogue chunk done (max over warps): " + " ".join(f"c{j}={max(v)}" for j, v in sorted(chunk.items()))
    line += "  (min: " + " ".join(f"{min(v)}" for j, v in sorted(chunk.items())) + ")"
    if nxt: line += f"   next acc_full +{nxt - acc[l]}"
    print(line)
    # MMA tiles of layer l+1 relative to acc[l]
    for me in (0, 1):
        tl = []
        for t_idx in range(32):
            r = [t for tag, a0, t in ev if tag == 1 + me and a0 == ((l + 1) << 8 | t_idx)]
            e = [t for tag, a0, t in ev if tag == 3 + me and a0 == ((l + 1) << 8 | t_idx)]
            w = [t for tag, a0, t in ev if tag == 5 + me and a0 == ((l + 1) << 8 | t_idx)]
            if r: tl.append((t_idx, r[0] - t0 - acc[l], e[0] - r[0], w[0] - e[0]))
        print(f"    issuer {me} tiles of layer {l+1} (idx: reach, epi-wait, w-wait): " + " ".join(f"{i}:{a},{b},{c}" for i, a, b, c in tl))
