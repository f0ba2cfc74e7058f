"""Emulate k_conv_wgrad_bf's data movement + MMA descriptor addressing under the assumed canonical-layout semantics."""
import numpy as np
rs = np.random.RandomState(0)
K = 6; Q = 200; Qalloc = (Q + 255)//256*256 + 640 + K
KT, APOS, STAGES = 64, 72, 4
A_UNITS, B_UNITS = 16*APOS, 16*KT
STAGE_UNITS = 2*A_UNITS + B_UNITS
inp = np.zeros((Qalloc, 64)); inp[:Q, :51] = rs.randn(Q, 51)
dA = np.zeros((Qalloc, 64)); dA[:Q-K+1, :50] = rs.randn(Q-K+1, 50)
# "hi" part only (values exact); lo = small perturbation to test the three products separately
inp_hi, inp_lo = inp, 1e-3*np.where(inp != 0, rs.randn(*inp.shape), 0)
dA_hi, dA_lo = dA, 1e-3*np.where(dA != 0, rs.randn(*dA.shape), 0)
def to_units(x, shift=0):           # [8][Qalloc] units of 8 channels; element arrays as float64 per (unit, 8)
    u = np.zeros((8, Qalloc, 8))
    for c in range(8): u[c, shift:, :] = x[:Qalloc-shift, 8*c:8*c+8]
    return u
in_hi_u, in_lo_u = to_units(inp_hi), to_units(inp_lo)
da_hi_u, da_lo_u = to_units(dA_hi, K-1), to_units(dA_lo, K-1)
npairs = (K+1)//2; ngroups = (npairs+3)//4
nstages_total = (Q + KT - 1)//KT
gW = np.zeros((K, 51, 50))
def mn_major_read(smem, start_unit, lbo_units, sbo_units, MN, Kdim):
    """matrix [MN][K] from units: element (mn,k) = unit[start + (mn//8)*SBO + (k%8) + (k//8)*LBO][mn%8]"""
    out = np.zeros((MN, Kdim))
    for mn in range(MN):
        for k in range(Kdim):
            out[mn, k] = smem[start_unit + (mn//8)*sbo_units + (k % 8) + (k//8)*lbo_units][mn % 8]
    return out
for g in range(ngroups):
    base, rem = npairs//ngroups, npairs % ngroups
    np_ = base + (1 if g < rem else 0); k0 = 2*(g*base + min(g, rem))
    acc = np.zeros((4, 128, 128))
    for si in range(nstages_total):
        smem = np.zeros((STAGE_UNITS, 8))
        q0 = si*KT
        for idx in range(42):
            if idx < 28:
                hl = idx//14; r2 = idx - hl*14; copy = r2//7; c = r2 - copy*7
                src = (in_lo_u if hl else in_hi_u)[c]
                s0 = q0 + k0 + copy
                d0 = hl*A_UNITS + (copy*8 + c)*APOS
                smem[d0:d0+APOS] = src[s0:s0+APOS]
            else:
                j = idx - 28; hl = j//7; c = j - hl*7
                src = (da_lo_u if hl else da_hi_u)[c]
                s0 = q0 + (K-1)
                d0 = 2*A_UNITS + (hl*8 + c)*KT
                smem[d0:d0+KT] = src[s0:s0+KT]
        for pr in range(np_):
            for ks in range(KT//16):
                aoff = 2*pr + 16*ks
                Ah = mn_major_read(smem, 0 + aoff, 8, APOS, 128, 16)        # LBO 128 B = 8 units
                Al = mn_major_read(smem, A_UNITS + aoff, 8, APOS, 128, 16)
                Bw = mn_major_read(smem, 2*A_UNITS + 16*ks, 8, KT, 128, 16)
                acc[pr][:, :128] += Ah @ Bw.T
                acc[pr][:, 64:128] += Al @ Bw[:64].T
    for pr in range(np_):
        for M in range(128):
            j, c = M >> 6, M & 63
            tap = k0 + 2*pr + j
            if tap < K and c < 51:
                gW[tap, c, :] += acc[pr][M, :50] + acc[pr][M, 64:114]
# reference: 3 products
ref = np.zeros((K, 51, 50))
for k in range(K):
    n = Q - K + 1
    A_h, A_l = inp_hi[k:k+n, :51], inp_lo[k:k+n, :51]
    ref[k] = A_h.T @ dA_hi[:n, :50] + A_h.T @ dA_lo[:n, :50] + A_l.T @ dA_hi[:n, :50]
print("max err", np.abs(gW - ref).max(), "scale", np.abs(ref).max())
