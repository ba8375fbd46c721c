"""Numerical experiment: how many terms of the fp16 hi/lo operand split does the FiLM-SIREN need for the 1e-3 bar?
Evaluates the renderer on the CPU oracle with the linear layers restated as products of fp16-rounded operands (exact in fp64):
3-term (hi*hi + lo*hi + hi*lo, the kernel), two 2-term forms and the single-pass form.  Results: DESIGN.md section 2."""
import sys; sys.path[:0]=["/root/repo"]
import torch, math
from oracle import cips3d_oracle as O
torch.manual_seed(0)
# emulate the ray MLP with operand splits; compare final composited features vs fp64 truth and fp32 reference
def split(x):
    hi = x.float().half(); lo = (x.float()-hi.float()).half()
    return hi.double(), lo.double()
def lin(x, W, mode):
    # x (..,K) double (already fp32-representable), W (N,K)
    if mode=="exact": return x@W.T
    xh,xl = split(x); Wh,Wl = split(W*256.0)
    if mode=="3": r = xh@Wh.T + xl@Wh.T + xh@Wl.T
    elif mode=="a_hilo": r = xh@Wh.T + xl@Wh.T
    elif mode=="w_hilo": r = xh@Wh.T + xh@Wl.T
    elif mode=="1": r = xh@Wh.T
    return r.float().double()/256.0
def film(sd,pre,x,w,mode, first=False):
    g,b = O.film_params({k:v.double() for k,v in sd.items()}, pre, w.double())
    W=sd[pre+".linear.weight"].double(); bias=sd[pre+".linear.bias"].double()
    z = lin(x,W,"exact" if first else mode)+bias   # layer 0 in kernel also split3 K=16; treat exact-ish
    return torch.sin(g[:,None]*z+b[:,None]).float().double()
def nerf(sd,pts,w,mode):
    x=pts*(2/0.24)
    x=film(sd,"siren.network.0",x,w,mode,first=True)
    x=film(sd,"siren.network.1",x,w,mode)
    sig=lin(x,sd["siren.final_layer.weight"].double(),mode)+sd["siren.final_layer.bias"].double()
    c=film(sd,"siren.color_layer_sine",x,w,mode)
    rgb=lin(c,sd["siren.color_layer_linear.0.weight"].double(),mode)+sd["siren.color_layer_linear.0.bias"].double()
    return torch.cat([rgb,sig],-1)
for sb in (0.0,0.3):
    sd=O.synthetic_state_dict(O.generator_template(),seed=77,sigma_bias=sb)
    B,R=1,32; kw=dict(O.G_KWARGS); S=kw["num_steps"]
    g=torch.Generator().manual_seed(1)
    zs={"z_nerf":torch.randn(B,256,generator=g),"z_inr":torch.randn(B,512,generator=g)}
    dr=O.draw_randoms(B,R,S,generator=g)
    sd64={k:v.double() for k,v in sd.items()}
    w=O.mapping_network(sd64,"mapping_network_nerf",zs["z_nerf"].double(),**O.G_CFG["mapping_nerf_cfg"])
    origin,_,_=O.camera_origin(dr["yaw_n"].double(),dr["pitch_n"].double(),kw["h_stddev"],kw["v_stddev"])
    c2w=O.cam2world(-origin,origin)
    def render(mode):
        orig=O.nerf_network
        O.nerf_network=(lambda sd_,pts,w_,**k: nerf(sd,pts,w_,mode)) if mode!="ref" else orig
        try:
            r=O.render_features(sd64,w,c2w,dr["jitter_u"].double(),dr["pdf_u"].double(),None,None,img_size=R,fov=kw["fov"],ray_start=kw["ray_start"],ray_end=kw["ray_end"],num_steps=S)
        finally: O.nerf_network=orig
        return r
    ref=render("ref")
    for mode in ("3","a_hilo","w_hilo","1"):
        r=render(mode)
        for key in ("coarse","pixels_fea"):
            e=(r[key]-ref[key]).abs().reshape(-1,r[key].shape[-1]).amax(-1)/ref[key].abs().max()
            print(f"sb={sb} mode={mode:7s} {key:10s} max-rel {e.max():.2e}  frac>1e-3 {(e>1e-3).double().mean():.4f}  p99 {e.quantile(0.99):.2e}")
