"""Pipeline timeline of the CIPS kernel from in-kernel clock stamps (needs the -DC3D_TRACE build; blocks 0 and 1, tile iteration 1).
Per layer: when the issuers reached / were released for / issued each weight tile, split into the waits that gate an MMA
(previous epilogue, own weight half, peer's half), the producer's stage waits, the epilogue's chunk times.
    C3D_LIB_PATH=.../libcips3d_b200_trace.so [C3D_CIPS_PAIR=1] python tools/trace_cips.py [B]"""
import os, sys, ctypes as C, collections, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cips3d_b200
from cips3d_b200 import _lib
from oracle import cips3d_oracle as O
dev = "cuda:0"
G = cips3d_b200.GeneratorNerfINR(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in O.G_CFG.items()}, device=dev).to(dev).eval()
G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
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
ev = []
for i in range(n):
    w, v = buf[2 * i], buf[2 * i + 1]
    ev.append((int(w >> 8), int(w & 0xFF), int((v >> 56) & 0xFF), int((v >> 40) & 0xFFFF), int(v & 0xFFFFFFFFFF)))   # block, warp, tag, a0, t
print("events", n)
if not ev:
    sys.exit(0)
t0 = min(e[4] for e in ev if e[0] == 0)
by = collections.defaultdict(dict)        # (block, tag, layer) -> {idx: t}
for blk, warp, tag, a0, t in ev:
    l, idx = a0 >> 8, a0 & 0xFF
    if tag in (8, 9):
        by[(blk, tag, l)].setdefault(idx if tag == 9 else (idx, warp), t - t0)
    else:
        by[(blk, tag, l)][idx] = t - t0
layers = sorted({k[2] for k in by if k[0] == 0 and k[1] in (1, 2)})
print("leader (block 0): per layer, medians over the layer's tiles of the issuer waits [clk]")
print("layer | first tile reached | epilogue-dependency wait | own-weights wait | peer-half wait | layer span (first reach -> last issue) | producer: stage-free stamps span")
for l in layers:
    waits_dep, waits_w, waits_peer, reach, issue = [], [], [], [], []
    for me in (0, 1):
        r, d, w12, w = by.get((0, 1 + me, l), {}), by.get((0, 3 + me, l), {}), by.get((0, 12, l), {}), by.get((0, 5 + me, l), {})
        for t_idx in r:
            if t_idx in d and t_idx in w:
                waits_dep.append(d[t_idx] - r[t_idx])
                own = w12.get(t_idx, w[t_idx])
                waits_w.append(own - d[t_idx])
                waits_peer.append(w[t_idx] - own)
                reach.append(r[t_idx]); issue.append(w[t_idx])
    prod = by.get((0, 10, l), {})
    if not reach:
        continue
    med = lambda v: int(statistics.median(v)) if v else 0      # noqa: E731
    print(f"{l:5d} | {min(reach):9d} | med {med(waits_dep):6d} max {max(waits_dep):6d} | med {med(waits_w):6d} max {max(waits_w):6d} | "
          f"med {med(waits_peer):6d} max {max(waits_peer):6d} | {max(issue) - min(reach):7d} | {(max(prod.values()) - min(prod.values())) if prod else 0:7d}")
# epilogue: chunk j ready (first warp) and done (last warp), per layer, both blocks
for blk in (0, 1):
    rows = []
    for l in layers:
        rdy = by.get((blk, 8, l), {})
        done = by.get((blk, 9, l), {})
        if not rdy:
            continue
        r = [min([t for (j, w), t in rdy.items() if j == jj] or [0]) for jj in range(4)]
        d = [max([t for idx, t in done.items() if (idx & 3) == jj] or [0]) for jj in range(4)]
        rows.append((l, r, d))
    if rows:
        print(f"block {blk}: epilogue per layer: acc_ready seen (first warp) j=0..3 | chunk done (last warp) j=0..3")
        for l, r, d in rows:
            print(f"  layer {l:2d}: " + " ".join(f"{v:8d}" for v in r) + " | " + " ".join(f"{v:8d}" for v in d))
# peer relay timing
rel = [(k[2], v) for k, v in by.items() if k[0] == 1 and k[1] == 11]
if rel:
    l, v = sorted(rel)[min(3, len(rel) - 1)]
    ts = [v[i] for i in sorted(v)]
    print(f"peer relay (block 1, layer {l}): gaps between consecutive relays [clk]: " + " ".join(str(b - a) for a, b in zip(ts, ts[1:])))

# per-tile detail of one steady-state layer on the leader: producer load issue -> own half landed -> both landed (MMA issue)
for L_ in (5,):
    prod, own, reach = by.get((0, 10, L_), {}), by.get((0, 12, L_), {}), {}
    issue, dep = {}, {}
    for me in (0, 1):
        issue.update(by.get((0, 5 + me, L_), {})); dep.update(by.get((0, 3 + me, L_), {})); reach.update(by.get((0, 1 + me, L_), {}))
    if prod and issue:
        print(f"layer {L_}, leader, per tile t: producer issued load | own half landed (+latency) | MMA issued | issuer reached | dep ok   [clk, relative to the layer's first stamp]")
        base = min(list(prod.values()) + list(issue.values()))
        for t_idx in sorted(issue):
            p_, o_, i_ = prod.get(t_idx), own.get(t_idx, issue[t_idx]), issue[t_idx]
            print(f"  t{t_idx:2d}: load {p_ - base if p_ is not None else -1:7d} | own landed {o_ - base if o_ is not None else -1:7d} "
                  f"(+{(o_ - p_) if (o_ is not None and p_ is not None) else -1:6d}) | mma {i_ - base:7d} | reached {reach.get(t_idx, base) - base:7d} | dep {dep.get(t_idx, base) - base:7d}")
    prod1, rel1 = by.get((1, 10, L_), {}), by.get((1, 11, L_), {})
    if prod1 and rel1:
        b1 = min(prod1.values())
        print(f"layer {L_}, peer (block 1; its own clock): producer issued load | its half landed (+latency)")
        for t_idx in sorted(rel1):
            if t_idx in prod1:
                print(f"  t{t_idx:2d}: load {prod1[t_idx] - b1:7d} | landed {rel1[t_idx] - b1:7d} (+{rel1[t_idx] - prod1[t_idx]:6d})")
