"""CPU replay of the tap-pair form of the bf16-split conv (viforssms_b200/csrc/nma_tc_conv2.cu: pp_mma / qq_mma, and
k_tc_pack_w_bf_pair in nma_tc_conv.cu): the pair tile layout [tap a: chunks 0-5][tap b: chunks 0-5][a: chunk 6]
[b: chunk 6], the seven K = 16 instructions per pair with their descriptor arithmetic (start unit, leading offset of
one position for the shared 7th chunk), and the zero-kernel pairing of an odd last tap - against the direct conv.
Assumes the no-swizzle K-major canonical layout (row r, K-chunk c of an instruction at start + r + c * LBO, in 16-byte
units), which the hardware runs of tests/test_gpu_bf16.py confirm for even kernel_len; the odd case has only this
replay so far."""
import numpy as np
import pytest

NPOS_TILE, CH = 256, 8          # positions per tile, channels per 16-byte unit


def pack_pairs(W, K, mode):
    """k_tc_pack_w_bf_pair without the hi/lo split: [npairs][14 chunks][64 rows][8]; mode 0 forward (rows = output
    channel n, chunk elements = input channels c), mode 1 data gradient (rows = input channel n, elements = f, flipped)."""
    npairs = (K + 1) // 2
    out = np.zeros((npairs, 14, 64, CH))
    for pr in range(npairs):
        for u in range(14):
            k = 2 * pr + (0 if u < 6 else 1 if u < 12 else u - 12)
            cch = u if u < 6 else u - 6 if u < 12 else 6
            if k >= K:
                continue
            for n in range(64):
                for e in range(CH):
                    c = 8 * cch + e
                    if mode == 0 and c < 51 and n < 50:
                        out[pr, u, n, e] = W[k, c, n]
                    if mode == 1 and c < 50 and n < 51:
                        out[pr, u, n, e] = W[K - 1 - k, n, c]
    return out


def mma_pairs(inp_units, tiles, K, npos):
    """pp_mma / qq_mma for one 256-position tile: inp_units [7 chunks][npos][8]; returns out[256][64]."""
    acc = np.zeros((NPOS_TILE, 64))
    flat = inp_units.reshape(7 * npos, CH)                     # unit index = chunk * npos + position
    for pr in range((K + 1) // 2):
        row = 2 * pr
        for ks in range(7):
            if ks < 6:
                start, lbo = row + (1 if ks >= 3 else 0) + (ks % 3) * 2 * npos, npos
            else:
                start, lbo = 6 * npos + row, 1                 # chunk 6 at tap a, then one position further on (tap b)
            # A operand [256 positions][16] of this instruction: K-chunk c of row r at unit start + r + c * lbo
            A = np.concatenate([flat[start + np.arange(NPOS_TILE) + c * lbo] for c in range(2)], axis=1)
            Bt = np.concatenate([tiles[pr, 2 * ks], tiles[pr, 2 * ks + 1]], axis=1)      # [64 rows][16]
            acc += A @ Bt.T
    return acc


@pytest.mark.parametrize("K", [6, 7, 1])
@pytest.mark.parametrize("mode", [0, 1])
def test_tap_pair_tiles_and_descriptors_reproduce_the_conv(K, mode):
    rs = np.random.RandomState(K * 3 + mode)
    npos = (NPOS_TILE + K + 7) & ~7                            # pair_npos()
    cin = 51 if mode == 0 else 50
    x = np.zeros((npos, 56))
    x[:, :cin] = rs.standard_normal((npos, cin))
    W = rs.standard_normal((K, 51, 50))
    units = np.stack([x[:, 8 * c:8 * c + 8] for c in range(7)])          # the 7 staged chunk slabs
    got = mma_pairs(units, pack_pairs(W, K, mode), K, npos)
    want = np.zeros((NPOS_TILE, 64))
    for k in range(K):
        if mode == 0:
            want[:, :50] += x[k:k + NPOS_TILE, :51] @ W[k]
        else:
            want[:, :51] += x[k:k + NPOS_TILE, :50] @ W[K - 1 - k].T
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
