"""CPU emulation of the rolling, row-paired weight-gradient scheme (csrc/conv_wgrad.cu: conv_wgrad_roll_kernel, opt-in
FP_WGRAD_ROLL=1).  A CTA walks down a column of row segments; every input row segment X[v] and every dZ row segment is
staged once.  Per input row two instructions are issued per K step (16 pixels):

  pair  : A = [dZ[y-1] (64 ch) ; dZ[y] (64 ch)]  (two ring slots, LBO = slot stride), B = X[v] with the three taps fused
          (N = 3*CB)  ->  TMEM lanes 0-63 accumulate filter row dy = 1, lanes 64-127 filter row dy = 0
  single: A = [dZ[y-2] ; (whatever follows)], same B  ->  lanes 0-63 of a second accumulator: filter row dy = 2

with y = v + pad the newest dZ row.  The dZ ring has 4 slots plus a 5th that duplicates every row with y % 4 == 0, so
that "slot(y-1) + 1" always holds row y.  Rows outside the image are zero rows (TMA out-of-bounds fill), so no
instruction is conditional.  This test mirrors those loops with numpy and checks the result against autograd."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def emulate_wgrad_roll(x, dz, pad, reflect, cb_size, seg):
    """x: [Hin, Win, C] virtual input of one image, dz: [Ho, Wo, cout <= 64] -> dW [cout, C, 3, 3]."""
    hin, win, c = x.shape
    ho, wo, cout = dz.shape
    ncb = c // cb_size
    tiles_per_row = (wo + seg - 1) // seg
    ksteps = seg // 16
    ncols = 3 * cb_size * ncb
    acc_pair = np.full((128, ncols), np.nan)       # lanes 0-63: dy = 1, lanes 64-127: dy = 0
    acc_single = np.full((128, ncols), np.nan)     # lanes 0-63: dy = 2
    ring = np.full((6, seg, 64), np.nan)           # slots 0..3 + duplicate slot 4 (+ one never-written guard slot)
    first = True

    def dz_row(y, x0):
        r = np.zeros((seg, 64))
        if 0 <= y < ho:
            n = min(seg, wo - x0)
            r[:n, :cout] = dz[y, x0:x0 + n]
        return r

    def strip(v, x0, cb):
        s = np.zeros((seg + 2, cb_size))
        vy = v
        if reflect:
            vy = -vy if vy < 0 else (2 * hin - 2 - vy if vy >= hin else vy)
        if not 0 <= vy < hin:
            return s
        for r in range(seg + 2):
            vx = x0 - pad + r
            if reflect:
                vx = -vx if vx < 0 else (2 * win - 2 - vx if vx >= win else vx)
            if 0 <= vx < win:
                s[r] = x[vy, vx, cb * cb_size:(cb + 1) * cb_size]
        return s

    for xseg in range(tiles_per_row):
        x0 = xseg * seg
        for y_new in range(-2, ho + 2):
            row = dz_row(y_new, x0)
            ring[y_new % 4] = row
            if y_new % 4 == 0:
                ring[4] = row
            if y_new < 0:
                continue                              # rows -2, -1 only prime the ring with zeros
            s_old = (y_new - 1) % 4                   # pair: [slot(y-1) ; slot(y-1) + 1] = [dZ[y-1] ; dZ[y]]
            s_old2 = (y_new - 2) % 4                  # single: [slot(y-2) ; slot(y-2) + 1] = [dZ[y-2] ; (ignored)]
            for cb in range(ncb):
                st = strip(y_new - pad, x0, cb)
                for ks in range(ksteps):
                    rows = slice(16 * ks, 16 * ks + 16)
                    a_pair = np.concatenate([ring[s_old][rows], ring[s_old + 1][rows]], 1)      # K16 x M128
                    a_single = np.concatenate([ring[s_old2][rows], ring[s_old2 + 1][rows]], 1)
                    b = np.concatenate([st[16 * ks + dx:16 * ks + dx + 16] for dx in range(3)], 1)  # K16 x N(3*CB)
                    cols = slice(cb * 3 * cb_size, (cb + 1) * 3 * cb_size)
                    if first and ks == 0:             # accumulate = 0 on the very first K step of each accumulator block
                        acc_pair[:, cols] = a_pair.T @ b
                        acc_single[:, cols] = a_single.T @ b
                    else:
                        acc_pair[:, cols] += a_pair.T @ b
                        acc_single[:, cols] += a_single.T @ b
            first = False
    dw = np.zeros((cout, c, 3, 3))
    for cb in range(ncb):
        for dx in range(3):
            cols = slice(cb * 3 * cb_size + dx * cb_size, cb * 3 * cb_size + (dx + 1) * cb_size)
            ch = slice(cb * cb_size, (cb + 1) * cb_size)
            dw[:, ch, 1, dx] = acc_pair[:cout, cols]
            dw[:, ch, 0, dx] = acc_pair[64:64 + cout, cols]
            dw[:, ch, 2, dx] = acc_single[:cout, cols]
    return dw


@pytest.mark.parametrize('cfg', [
    dict(c=64, cout=64, hin=9, win=150, pad=1, reflect=True, cb=64, seg=160),
    dict(c=64, cout=24, hin=7, win=200, pad=1, reflect=False, cb=32, seg=112),    # two channel blocks, two ragged segments
    dict(c=32, cout=2, hin=5, win=96, pad=1, reflect=True, cb=32, seg=96),        # a head: cout = 2
])
def test_rolling_paired_wgrad_matches_autograd(cfg):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(cfg['hin'], cfg['win'], cfg['c'], generator=g, dtype=torch.float64)
    w = torch.randn(cfg['cout'], cfg['c'], 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
    xn = x.permute(2, 0, 1)[None]
    if cfg['reflect']:
        z = F.conv2d(F.pad(xn, (1, 1, 1, 1), mode='reflect'), w)
    else:
        z = F.conv2d(xn, w, padding=cfg['pad'])
    dz = torch.randn(z.shape, generator=g, dtype=torch.float64)
    z.backward(dz)
    # the first K step of a (ring slot, accumulator) pair must not read an unwritten slot: NaNs would show up here
    got = emulate_wgrad_roll(x.numpy(), dz[0].permute(1, 2, 0).numpy(), cfg['pad'], cfg['reflect'], cfg['cb'], cfg['seg'])
    assert np.isfinite(got).all()
    assert np.abs(got - w.grad.numpy()).max() < 1e-9
