"""Static description of one NMA (Neural Moving Average) flow stack.

`NMAConfig` is the host-side mirror of `struct nma_config` in
`include/nma_b200.h`; `param_layout` is the host-side mirror of the flat
parameter blob the C-ABI library consumes.  The blob order is the TF variable
creation order of the reference (`AR.py:53-78`, `fitz_nag_NVP.py:71-96`,
`SV_dense.py:53-76`): per flow — 4 feature dense layers (kernel `[in,out]`,
bias), the K-tap conv (kernel `[K,Cin,Cout]`, bias), 3 theta dense layers,
H hidden 1x1 convs (kernel `[C,C]`, bias [, BN gamma, beta]) and the 2-unit
head (kernel `[C,2]`, bias).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

MODEL_AR = 0
MODEL_FHN = 1
MODEL_SV = 2
MODEL_LV = 3      # Lotka-Volterra, fixed theta (lotka_volterra_partial_batch_fix_theta.py)
MODEL_LVR = 4     # Lotka-Volterra, learned theta (lotka_volterra_partial.py, the script whose data the reference ships)
MODEL_LVB = 5     # Lotka-Volterra, learned softplus-theta, p_val windows per iteration (lotka_volterra_partial_batch.py)
LV_MODELS = (MODEL_LV, MODEL_LVR, MODEL_LVB)     # all use the transposed wide feature layer and the 1 + (L0 - 1)-channel conv

MAX_CHAN = 32
MAX_ARRAYS = 8
MAX_FLOWS = 8
C_FIXED = 50  # network_dims[0] in every script of the reference (Appendix H of SURVEY.md)

OBJ_ELBO = 0        # -sum_rows scale*(sde - logq + obs)        (AR.py:184-185,228-229)
OBJ_NEG_OBS = 1     # -sum_rows obs_log_prob                     (AR.py:201-202)
OBJ_PATH_SQ = 2     # sum (lf_sample - c)^2                      (fitz_nag_NVP.py:288-289, SV_dense.py:251-252)


class CConfig(ctypes.Structure):
    """Byte-for-byte `struct nma_config` (include/nma_b200.h)."""
    _fields_ = [
        ("model", ctypes.c_int32),
        ("p", ctypes.c_int32),
        ("K", ctypes.c_int32),
        ("B", ctypes.c_int32),
        ("D", ctypes.c_int32),
        ("F", ctypes.c_int32),
        ("C", ctypes.c_int32),
        ("H", ctypes.c_int32),
        ("bn", ctypes.c_int32),
        ("Cf", ctypes.c_int32),
        ("feat_aug", ctypes.c_int32),
        ("dtheta", ctypes.c_int32),
        ("n_arrays", ctypes.c_int32),
        ("obs_array", ctypes.c_int32),
        ("bin_array", ctypes.c_int32),
        ("head_offset", ctypes.c_int32),
        ("chan_array", ctypes.c_int32 * MAX_CHAN),
        ("chan_offset", ctypes.c_int32 * MAX_CHAN),
        ("scale", ctypes.c_double),
        ("dt", ctypes.c_float),
        ("obs_std", ctypes.c_float),
        ("x0", ctypes.c_float * 2),
        ("n_pinned", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


@dataclass
class NMAConfig:
    model: int = MODEL_AR
    p: int = 50               # rows per step = MC samples = subsequences (AR.py:117,263-265)
    K: int = 50               # kernel_len
    B: int = 50               # batch_dims
    D: int = 1                # flow_dims: latent components interleaved on the flow's time axis
    F: int = 3                # no_flows
    C: int = C_FIXED          # network_dims[0]
    H: int = 1                # len(network_dims) - 2 hidden 1x1 layers
    bn: int = 0               # inference-mode batch-norm affine after each hidden layer
    Cf: int = 14              # channels of time_feats
    feat_aug: int = 0         # SV: [f[1:], f[1:, :-2] - f[:-1, :-2]] (SV_dense.py:53)
    dtheta: int = 3
    scale: float = 100.0      # T / batch_dims (AR.py:184)
    dt: float = 1.0
    obs_std: float = 1.0
    x0: Tuple[float, float] = (0.0, 0.0)
    n_arrays: int = 4
    # feature channel c of window slot j of row r reads
    #   base[chan_array[c]][D*idx_r + j + chan_offset[c]]
    chan_array: List[int] = field(default_factory=list)
    chan_offset: List[int] = field(default_factory=list)
    obs_array: int = 0        # base array holding the evaluated observations
    bin_array: int = 3        # base array holding the observation indicator (AR: padded obs_bin)
    head_offset: int = 0
    n_pinned: int = 1         # leading states of the series that mask / shift pin to x0 (p_val in the LV batch scripts)

    # ---- derived ----
    @property
    def L0(self) -> int:
        """kernel_ext: AR.py:132, fitz_nag_NVP.py:182-183."""
        return self.F * self.K + self.D * self.B + self.D

    def L(self, i: int) -> int:
        return self.L0 - i * self.K

    def Lin(self, i: int) -> int:
        return self.L(i) - 1

    def N(self, i: int) -> int:
        return self.L(i) - self.K

    @property
    def S(self) -> int:
        """slots whose log sigma enters logq (AR.py:84; fitz_nag_NVP.py:278-279)."""
        return self.D * self.B

    @property
    def Cf_in(self) -> int:
        return self.Cf + (self.Cf - 2 if self.feat_aug else 0)

    @property
    def feat_off(self) -> int:
        return 1 if self.feat_aug else 0

    def validate(self) -> None:
        if self.C != C_FIXED:
            raise ValueError("only network_dims[0]=50 is compiled (every reference script uses 50)")
        if self.D not in (1, 2):
            raise ValueError("flow_dims must be 1 or 2")
        if not (1 <= self.F <= MAX_FLOWS):
            raise ValueError("no_flows out of range")
        if len(self.chan_array) != self.Cf or len(self.chan_offset) != self.Cf:
            raise ValueError("channel table must have Cf entries")
        if self.Cf_in > MAX_CHAN:
            raise ValueError("too many feature channels")
        if self.n_arrays > MAX_ARRAYS:
            raise ValueError("too many base arrays")

    def to_c(self) -> CConfig:
        self.validate()
        c = CConfig()
        for name in ("model", "p", "K", "B", "D", "F", "C", "H", "bn", "Cf", "feat_aug",
                     "dtheta", "n_arrays", "obs_array", "bin_array", "head_offset"):
            setattr(c, name, int(getattr(self, name)))
        for i in range(self.Cf):
            c.chan_array[i] = int(self.chan_array[i])
            c.chan_offset[i] = int(self.chan_offset[i])
        c.scale = float(self.scale)
        c.dt = float(self.dt)
        c.obs_std = float(self.obs_std)
        c.x0[0] = float(self.x0[0])
        c.x0[1] = float(self.x0[1])
        c.n_pinned = int(self.n_pinned)
        return c


def param_layout(cfg: NMAConfig) -> Tuple[Dict[str, Tuple[int, Tuple[int, ...]]], int]:
    """name -> (offset, shape) in the flat fp32 blob, plus the total count.

    Names: f{i}.feat{l}.w/.b, f{i}.conv.w/.b, f{i}.th{l}.w/.b, f{i}.hid{l}.w/.b
    [, f{i}.hid{l}.gamma/.beta], f{i}.head.w/.b
    """
    C = cfg.C
    out: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
    off = 0

    def add(name: str, shape: Tuple[int, ...]) -> None:
        nonlocal off
        n = 1
        for s in shape:
            n *= s
        out[name] = (off, shape)
        off += n

    for i in range(cfg.F):
        # LV (lotka_volterra_partial_batch_fix_theta.py:71-82): the 4th feature layer is as wide as the flow's conv
        # input (feat_dims = L_i - 1) and its TRANSPOSE feeds the conv, whose input channels are then 1 + (L0 - 1)
        lv = cfg.model in LV_MODELS
        for l in range(4):
            add(f"f{i}.feat{l}.w", (cfg.Cf_in if l == 0 else C, cfg.Lin(i) if (lv and l == 3) else C))
            add(f"f{i}.feat{l}.b", (cfg.Lin(i) if (lv and l == 3) else C,))
        add(f"f{i}.conv.w", (cfg.K, cfg.L0 if lv else C + 1, C))
        add(f"f{i}.conv.b", (C,))
        for l in range(3):
            add(f"f{i}.th{l}.w", (cfg.dtheta if l == 0 else C, C))
            add(f"f{i}.th{l}.b", (C,))
        for l in range(cfg.H):
            add(f"f{i}.hid{l}.w", (C, C))
            add(f"f{i}.hid{l}.b", (C,))
            if cfg.bn:
                add(f"f{i}.hid{l}.gamma", (C,))
                add(f"f{i}.hid{l}.beta", (C,))
        add(f"f{i}.head.w", (C, 2))
        add(f"f{i}.head.b", (2,))
    return out, off


def ar_config(p=50, K=50, B=50, F=3, H=1, feat_window=10, T=5000, obs_std=1.0, x0=10.0) -> NMAConfig:
    """The AR(1) model of AR.py / hyperparameters.txt.

    Base arrays (AR.py:135-150): 0 = zero-padded observations (one array serves
    all `feat_window` look-ahead channels through `chan_offset`), 1 = bin_feats,
    2 = time_pad, 3 = time_till, 4 = obs_bin.  Channel order AR.py:282-283.
    """
    fw = feat_window
    return NMAConfig(
        model=MODEL_AR, p=p, K=K, B=B, D=1, F=F, H=H, bn=0, Cf=fw + 4, feat_aug=0, dtheta=3,
        scale=float(T) / float(B), dt=1.0, obs_std=obs_std, x0=(x0, 0.0), n_arrays=5,
        chan_array=[0] * fw + [1, 2, 3, 4], chan_offset=list(range(fw)) + [0, 0, 0, 0],
        obs_array=0, bin_array=4)


def fhn_config(p=50, K=20, B=50, F=3, H=3, feat_window=10, target_dims=1000000, dt=0.1) -> NMAConfig:
    """The FitzHugh-Nagumo model of fitz_nag_NVP.py (RealNVP-style coupling on a time axis that interleaves the
    two latent components).

    Base arrays (fitz_nag_NVP.py:187-202): 0 = 'F'-flattened observations padded with no_flows*kernel_len + 2
    zeros (look-ahead channels read it at offsets 0, 5, ..., 5*(feat_window-1)), 1 = bin_feats, 2 = time_pad,
    3 = interleaved time_till (its pad is 2 slots longer than the others), 4 = obs_bin
    [2, target_dims] row-major (bin_feed, fitz_nag_NVP.py:369-370).  Channel order fitz_nag_NVP.py:362-363."""
    fw = feat_window
    return NMAConfig(
        model=MODEL_FHN, p=p, K=K, B=B, D=2, F=F, H=H, bn=1, Cf=fw + 3, feat_aug=0, dtheta=5,
        scale=float(target_dims) / float(B), dt=dt, obs_std=0.1, x0=(0.0, 0.0), n_arrays=5,
        chan_array=[0] * fw + [1, 2, 3], chan_offset=[5 * i for i in range(fw)] + [0, 0, 0],
        obs_array=0, bin_array=4)


def sv_config(p=200, K=50, B=52, F=5, H=3, feat_window=5, target_dims=1508, dt=1.0, x0=-8.5) -> NMAConfig:
    """The stochastic-volatility model of SV_dense.py: a 1-D latent log-volatility flow with the observed price as a
    second, fixed component (`dim_one`), delta-augmented features (SV_dense.py:53) and no observation term.

    Base arrays (SV_dense.py:159-184): 0 = observations padded with no_flows*kernel_len zeros (NOT +1; look-ahead
    channels at offsets 0, 5, ...; also the source of dim_one = obs[idx : idx+B+1], i.e. head_offset = F*K),
    1 = time_pad, 2 = rolling variance, 3 = log rolling variance of the first differences.
    Channel order SV_dense.py:312-320."""
    fw = feat_window
    return NMAConfig(
        model=MODEL_SV, p=p, K=K, B=B, D=1, F=F, H=H, bn=1, Cf=fw + 3, feat_aug=1, dtheta=4,
        scale=float(target_dims) / float(B), dt=dt, obs_std=1.0, x0=(x0, 0.0), n_arrays=4,
        chan_array=[0] * fw + [1, 2, 3], chan_offset=[5 * i for i in range(fw)] + [0, 0, 0],
        obs_array=0, bin_array=0, head_offset=F * K)


def lvr_config(p=50, K=20, B=50, F=3, H=3, feat_window=10, target_dims=500, dt=0.1, x0=(100.0, 100.0)) -> NMAConfig:
    """The Lotka-Volterra model of lotka_volterra_partial.py (learned theta; defaults = the script's :466-480, which runs
    on the dat/LV_*.txt files the reference ships).  Same flow as `lv_config` (transposed wide 4th feature layer, conv
    over 1 + (L0 - 1) channels, coupling / Permute / BN); p rows with their own window each, as in the FHN script.
    ELBO (ibid. :234-275): path = softplus(flow output) * mask + shift, bivariate Euler-Maruyama density on the state
    DIFFERENCES with theta = exp(theta sample) (3 rates), unit-variance Gaussian observations.

    Base arrays (ibid. :186-205): as the FHN model except that bin_feats is 0 on the pad and 1 on the series and the
    lead of time_till stops before 0; the time channel starts at dt as in FHN."""
    fw = feat_window
    return NMAConfig(
        model=MODEL_LVR, p=p, K=K, B=B, D=2, F=F, H=H, bn=1, Cf=fw + 3, feat_aug=0, dtheta=3,
        scale=float(target_dims) / float(B), dt=dt, obs_std=1.0, x0=(float(x0[0]), float(x0[1])), n_arrays=5,
        chan_array=[0] * fw + [1, 2, 3], chan_offset=[5 * i for i in range(fw)] + [0, 0, 0],
        obs_array=0, bin_array=4)


def lvb_config(p=3, K=20, B=151, F=3, H=3, feat_window=10, target_dims=151, dt=0.2, x0=(91.0, 99.0)) -> NMAConfig:
    """lotka_volterra_partial_batch.py (:677-764): the flow, feed and observation model of `lv_config` with a LEARNED theta
    (4 softplus-scale parameters sampled per row, :198-200), the plain bivariate transition density (:339-343) and p = p_val
    windows per iteration tiling p_val concatenated series, the first p_val states of which are pinned (:237-240)."""
    cfg = lv_config(p=p, K=K, B=B, F=F, H=H, feat_window=feat_window, target_dims=target_dims, dt=dt, x0=x0)
    cfg.model = MODEL_LVB
    cfg.n_pinned = p
    return cfg


def lv_config(p=1, K=20, B=151, F=3, H=3, feat_window=10, target_dims=151, dt=0.2, x0=(91.0, 99.0)) -> NMAConfig:
    """The Lotka-Volterra model of lotka_volterra_partial_batch_fix_theta.py (p_val = 1: one subsequence that is the
    whole series): feed, flow (transposed wide 4th feature layer, conv over 1 + (L0 - 1) channels, coupling / Permute /
    BN) and the bivariate ELBO with its softplus bijector chains.

    Base arrays (ibid. :203-222): as the FHN model, except that bin_feats is 0 on the pad and 1 on the series, the
    time channel starts at 0, and the lead of time_till stops before 0.  Channel order :497-498."""
    fw = feat_window
    return NMAConfig(
        model=MODEL_LV, p=p, K=K, B=B, D=2, F=F, H=H, bn=1, Cf=fw + 3, feat_aug=0, dtheta=4,
        scale=float(target_dims) / float(B), dt=dt, obs_std=1.0, x0=(float(x0[0]), float(x0[1])), n_arrays=5,
        chan_array=[0] * fw + [1, 2, 3], chan_offset=[5 * i for i in range(fw)] + [0, 0, 0],
        obs_array=0, bin_array=4)
