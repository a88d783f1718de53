"""CPU emulation of the row-streaming convolution's control logic (csrc/conv_igemm.cu: conv_rowstream_kernel).

The kernel itself is GPU-only and still opt-in (FP_ROWSTREAM=1); this test pins the part that can be checked without
hardware: the item -> (column, row chunk) decode, which filter rows an input row feeds (dy_lo .. dy_hi), the TMEM row
slot each one accumulates into (slot(y) = H-1-(y-yc), so that dy = 0,1,2 are adjacent increasing column blocks), the
resident filter tile order ((dx*cblocks + cb)*3 + dy), the accumulator-opening instruction, and the point at which an
output row is complete.  The loops below mirror the MMA / epilogue loops of the kernel one to one with numpy matrices in
place of shared memory, TMEM and tcgen05.mma."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

BM = 128


def emulate_rowstream(x, w, pad, reflect, bn, cb_size):
    """x: [Hin, Win, C] virtual input (upsample / concat already applied), w: [cout, C, 3, 3] -> [Ho, Wo, cout]."""
    hin, win, c = x.shape
    cout = w.shape[0]
    ho, wo = hin + 2 * pad - 2, win + 2 * pad - 2
    slots = 512 // bn
    h = slots // 2                                   # output rows per item = row slots per TMEM half
    cblocks = c // cb_size
    tiles_per_row = (wo + BM - 1) // BM
    chunks_per_col = (ho + h - 1) // h
    # resident filter: tile[(dx*cblocks + cb)*3 + dy] = [bn rows (cout padded)][cb_size]
    res = np.zeros((9 * cblocks, bn, cb_size), np.float64)
    for dx in range(3):
        for cb in range(cblocks):
            for dy in range(3):
                res[(dx * cblocks + cb) * 3 + dy, :cout] = w[:, cb * cb_size:(cb + 1) * cb_size, dy, dx]
    out = np.full((ho, wo, cout), np.nan)
    tmem = np.full((BM, 512), np.nan)                # NaN = never written: reading an unopened accumulator would show

    def strip(v, x0, cb):                            # 130 input pixels of virtual row v, channel block cb
        s = np.zeros((BM + 2, cb_size), np.float64)
        vy = v
        if reflect:
            vy = -vy if vy < 0 else (2 * hin - 2 - vy if vy >= hin else vy)
        if not 0 <= vy < hin:
            return s
        for r in range(BM + 2):
            vx = x0 - pad + r
            if reflect:
                vx = -vx if vx < 0 else (2 * win - 2 - vx if vx >= win else vx)
            if 0 <= vx < win:
                s[r] = x[vy, vx, cb * cb_size:(cb + 1) * cb_size]
        return s

    for it, item in enumerate(range(tiles_per_row * chunks_per_col)):
        chunk, xseg = item % chunks_per_col, item // chunks_per_col
        x0, yc = xseg * BM, chunk * h
        half = it & 1
        for j in range(h + 2):
            dy_lo = j - (h - 1) if j > h - 1 else 0
            dy_hi = j if j < 2 else 2
            ndy = dy_hi - dy_lo + 1
            q_hi = (h - 1) - (j - dy_lo)
            d_col = (half * h + q_hi) * bn
            for cb in range(cblocks):
                a_strip = strip(yc - pad + j, x0, cb)
                for dx in range(3):
                    t0 = (dx * cblocks + cb) * 3 + dy_lo
                    for k in range(cb_size // 16):
                        a = a_strip[dx:dx + BM, 16 * k:16 * k + 16]                       # M128 x K16
                        b = res[t0:t0 + ndy, :, 16 * k:16 * k + 16].reshape(ndy * bn, 16)  # N = ndy*bn x K16
                        if dy_lo == 0 and cb == 0 and dx == 0 and k == 0:
                            tmem[:, d_col:d_col + bn] = a @ b[:bn].T                       # opens the row's accumulator
                            if ndy > 1:
                                tmem[:, d_col + bn:d_col + ndy * bn] += a @ b[bn:].T
                        else:
                            tmem[:, d_col:d_col + ndy * bn] += a @ b.T
            if j >= 2:                                # output row yc + j - 2 is complete -> epilogue
                slot = half * h + (h - 1) - (j - 2)
                y = yc + j - 2
                if y < ho:
                    n = min(BM, wo - x0)
                    out[y, x0:x0 + n] = tmem[:n, slot * bn:slot * bn + cout]
                tmem[:, slot * bn:(slot + 1) * bn] = np.nan                                # slot handed back
    return out


@pytest.mark.parametrize('cfg', [
    dict(c=32, cout=32, hin=11, win=140, pad=1, reflect=True, bn=32, cb=32),     # H = 8, ragged column and row ends
    dict(c=64, cout=24, hin=9, win=130, pad=2, reflect=False, bn=32, cb=32),     # two channel blocks, full correlation
    dict(c=64, cout=64, hin=10, win=128, pad=1, reflect=True, bn=64, cb=64),     # H = 4
    dict(c=64, cout=40, hin=7, win=200, pad=2, reflect=False, bn=64, cb=32),
])
def test_rowstream_control_logic_matches_conv2d(cfg):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(cfg['hin'], cfg['win'], cfg['c'], generator=g, dtype=torch.float64)
    w = torch.randn(cfg['cout'], cfg['c'], 3, 3, generator=g, dtype=torch.float64)
    xn = x.permute(2, 0, 1)[None]
    if cfg['reflect']:
        ref = F.conv2d(F.pad(xn, (1, 1, 1, 1), mode='reflect'), w)
    else:
        ref = F.conv2d(xn, w, padding=cfg['pad'])
    got = emulate_rowstream(x.numpy(), w.numpy(), cfg['pad'], cfg['reflect'], cfg['bn'], cfg['cb'])
    assert np.isfinite(got).all()
    assert np.abs(got - ref[0].permute(1, 2, 0).numpy()).max() < 1e-9
