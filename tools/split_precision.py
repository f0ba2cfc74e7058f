"""CPU experiment: does a 2-term bf16 split (3 products at the kind::f16 rate) of the conv operands hold the 1e-4 bar?
Emulates the split on the conv's three GEMMs (forward, data gradient, weight gradient) inside the fp64 oracle."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import nma_oracle as O
from viforssms_b200 import feed
from viforssms_b200.config import ar_config, param_layout
import torch.nn.functional as Fnn

def split(x, mode):
    x32 = x.float()
    if mode == 'bf16':
        hi = x32.bfloat16().float(); lo = (x32 - hi).bfloat16().float()
    elif mode == 'bf16t':   # truncation split
        hi = (x32.view(torch.int32) & ~0xffff).view(torch.float32); r = x32 - hi
        lo = (r.view(torch.int32) & ~0xffff).view(torch.float32)
    elif mode == 'tf32':
        hi = (x32.view(torch.int32) & ~0x1fff).view(torch.float32); lo = x32 - hi
    return hi.double(), lo.double()

MODE = None
class SplitConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, w):
        ctx.save_for_backward(inp, w)
        ih, il = split(inp, MODE); wh, wl = split(w, MODE)
        return orig(ih, wh) + orig(ih, wl) + orig(il, wh)
    @staticmethod
    def backward(ctx, g):
        inp, w = ctx.saved_tensors
        ih, il = split(inp, MODE); wh, wl = split(w, MODE); gh, gl = split(g, MODE)
        cg = lambda gg, ww: torch.nn.grad.conv1d_input(inp.shape, ww, gg)
        cw = lambda ii, gg: torch.nn.grad.conv1d_weight(ii, w.shape, gg)
        return cg(gh, wh) + cg(gh, wl) + cg(gl, wh), cw(ih, gh) + cw(ih, gl) + cw(il, gh)

orig = Fnn.conv1d
def patched(inp, w, b=None):
    if MODE is None or w.shape[2] == 1: return orig(inp, w, b)
    out = SplitConv.apply(inp, w)
    return out + b[None, :, None] if b is not None else out

def run(p=16, seed=1):
    global MODE
    T = 5000
    cfg = ar_config(p=p)
    d = '/root/repo/dat'
    obs = np.loadtxt(d + "/AR_obs_partial.txt", np.float32); obs_bin = np.loadtxt(d + "/AR_obs_binary.txt", np.float32)
    tt = np.loadtxt(d + "/AR_time_till.txt", np.float32)
    rs = np.random.RandomState(seed)
    idx = feed.sample_indices(T, cfg.B, cfg.p, rs)
    g = torch.Generator().manual_seed(seed)
    layout, n = param_layout(cfg)
    params = O.glorot_init(layout, n, g)
    for i in range(cfg.F):
        off, shape = layout[f"f{i}.feat0.w"]
        params[off:off + shape[0] * shape[1]].reshape(shape)[11, :] *= 10.0 / T
    eps = torch.randn(cfg.p, cfg.L0, generator=g)
    theta = torch.tensor([4.0, 0.5, 1.0]).repeat(cfg.p, 1) + 0.1 * torch.randn(cfg.p, 3, generator=g)
    pads = O.pad_series_ar(obs, obs_bin, tt, 10.0, T, cfg.F, cfg.K, 10)
    tf64, _, _ = O.gather_feed_ar(pads, idx, cfg.L0, cfg.B)
    tf = torch.from_numpy(tf64.astype(np.float32)).double()
    MODE = None
    ref = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf)
    for mode in ('tf32', 'bf16', 'bf16t'):
        MODE = mode
        Fnn.conv1d = patched
        out = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf)
        Fnn.conv1d = orig
        terr = ((out["terms"] - ref["terms"]).abs().max() / max(1.0, ref["terms"].abs().max())).item()
        gerr = ((out["grad_params"] - ref["grad_params"]).norm() / ref["grad_params"].norm()).item()
        worst = 0
        for k, (o, s) in layout.items():
            m = int(np.prod(s)); a = out["grad_params"][o:o+m]; b = ref["grad_params"][o:o+m]
            if b.norm() > 0: worst = max(worst, ((a-b).norm()/b.norm()).item())
        therr = ((out["grad_theta"] - ref["grad_theta"]).norm() / ref["grad_theta"].norm()).item()
        print(f"{mode}: terms {terr:.2e} grad {gerr:.2e} worst-var {worst:.2e} gtheta {therr:.2e}")
for seed in (1, 2, 3):
    run(16, seed)
