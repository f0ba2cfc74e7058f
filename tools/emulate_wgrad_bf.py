"""Replays k_conv_wgrad_bf's data movement (slab copies, per-row stage walk) and MMA descriptor addressing in numpy,
under the canonical-layout semantics the kernel assumes (no-swizzle MN-major: element (mn, k) at
start + (mn / 8) * SBO + (k % 8) * 16 B + (k / 8) * LBO), and compares with the direct sum.  CPU only; it was the
pre-flight check of the kernel's index arithmetic (the hardware semantics themselves are covered by tests/test_gpu_bf16.py)."""
import numpy as np

rs = np.random.RandomState(0)
K, p, Lin = 20, 3, 70                # rows of Lin positions, N = Lin - (K - 1) of them carry dA
N = Lin - (K - 1)
qtot = p * Lin
Qalloc = (qtot + 255) // 256 * 256 + 640 + K
KT, APOS = 64, 72
A_UNITS, B_UNITS = 16 * APOS, 16 * KT
STAGE_UNITS = 2 * A_UNITS + B_UNITS
inp = np.zeros((Qalloc, 64)); inp[:qtot, :51] = rs.randn(qtot, 51)
dAf = np.zeros((Qalloc, 64))          # dA(r, m) at flattened q = r*Lin + m, zero for m >= N
for r in range(p):
    dAf[r * Lin:r * Lin + N, :50] = rs.randn(N, 50)
inp_hi, inp_lo = inp, 1e-3 * np.where(inp != 0, rs.randn(*inp.shape), 0)
dA_hi, dA_lo = dAf, 1e-3 * np.where(dAf != 0, rs.randn(*dAf.shape), 0)


def to_units(x, shift=0):
    u = np.zeros((8, Qalloc, 8))
    for c in range(8):
        u[c, shift:, :] = x[:Qalloc - shift, 8 * c:8 * c + 8]
    return u


in_hi_u, in_lo_u = to_units(inp_hi), to_units(inp_lo)
da_hi_u, da_lo_u = to_units(dA_hi, K - 1), to_units(dA_lo, K - 1)
npairs = (K + 1) // 2
ngroups = (npairs + 3) // 4
rows, Lr, Nv = p, Lin, N
if rows > 1 and ((16 - Nv % 16) % 16) > K - 1:
    Nv = Lr = rows * Lr; rows = 1
spr = (Nv + KT - 1) // KT
nstages_total = rows * spr
gW = np.zeros((K, 51, 50))


def mn_major_read(smem, start_unit, lbo_units, sbo_units, MN, Kdim):
    out = np.zeros((MN, Kdim))
    for mn in range(MN):
        for k in range(Kdim):
            out[mn, k] = smem[start_unit + (mn // 8) * sbo_units + (k % 8) + (k // 8) * lbo_units][mn % 8]
    return out


for g in range(ngroups):
    base, rem = npairs // ngroups, npairs % ngroups
    np_ = base + (1 if g < rem else 0)
    k0 = 2 * (g * base + min(g, rem))
    acc = np.zeros((4, 128, 128))
    for gs in range(nstages_total):
        smem = np.zeros((STAGE_UNITS, 8))
        row = gs // spr
        q0 = row * Lr + (gs - row * spr) * KT
        left = Nv - (gs % spr) * KT
        nks = KT // 16 if left >= KT else (left + 15) // 16
        for idx in range(42):
            if idx < 28:
                hl = idx // 14; r2 = idx - hl * 14; copy = r2 // 7; c = r2 - copy * 7
                src = (in_lo_u if hl else in_hi_u)[c]
                s0 = q0 + k0 + copy
                d0 = hl * A_UNITS + (copy * 8 + c) * APOS
                smem[d0:d0 + APOS] = src[s0:s0 + APOS]
            else:
                j = idx - 28; hl = j // 7; c = j - hl * 7
                src = (da_lo_u if hl else da_hi_u)[c]
                s0 = q0 + (K - 1)
                d0 = 2 * A_UNITS + (hl * 8 + c) * KT
                smem[d0:d0 + KT] = src[s0:s0 + KT]
        for pr in range(np_):
            for ks in range(nks):
                aoff = 2 * pr + 16 * ks
                Ah = mn_major_read(smem, aoff, 8, APOS, 128, 16)            # LBO 128 B = 8 units
                Al = mn_major_read(smem, A_UNITS + aoff, 8, APOS, 128, 16)
                Bw = mn_major_read(smem, 2 * A_UNITS + 16 * ks, 8, KT, 128, 16)
                acc[pr][:, :128] += Ah @ Bw.T
                acc[pr][:, 64:128] += Al @ Bw[:64].T
    for pr in range(np_):
        for M in range(128):
            j, c = M >> 6, M & 63
            tap = k0 + 2 * pr + j
            if tap < K and c < 51:
                gW[tap, c, :] += acc[pr][M, :50] + acc[pr][M, 64:114]

ref = np.zeros((K, 51, 50))
for k in range(K):
    for r in range(p):
        A_h = inp_hi[r * Lin + k:r * Lin + k + N, :51]; A_l = inp_lo[r * Lin + k:r * Lin + k + N, :51]
        D_h = dA_hi[r * Lin:r * Lin + N, :50]; D_l = dA_lo[r * Lin:r * Lin + N, :50]
        ref[k] += A_h.T @ D_h + A_h.T @ D_l + A_l.T @ D_h
print("rows walked per row:", rows > 1, "max err", np.abs(gW - ref).max(), "scale", np.abs(ref).max())
assert np.abs(gW - ref).max() < 1e-9 * np.abs(ref).max()
