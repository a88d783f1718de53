"""Which role paces the row-streaming kernel?  Needs a `make TRIAGE=1` build; FP_DEBUG_FLAGS: 1 no input loads, 2 no MMAs,
4 no epilogue math / stores.  Times one layer shape (default outconv4.0.conv2 forward at bs 64)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from footprints_b200 import _lib
from footprints_b200._lib import FpConvDesc
lib = _lib.load()
def P(t): return ctypes.c_void_p(t.data_ptr()) if t is not None else None
cfgs = {'o4.c2 fwd (reflect, cp.async)': dict(c=32, cout=32, pad=1, mode=1, h=192, w=640),
        'o4.c2 dgrad (zero pad 2, TMA)': dict(c=32, cout=32, pad=2, mode=0, h=192, w=640),
        'block4.post2 fwd': dict(c=64, cout=64, pad=1, mode=1, h=96, w=320)}
b = 64
for name, c in cfgs.items():
    x = torch.randn(b, c['h'], c['w'], c['c'], device='cuda').half()
    kpad = (9 * c['c'] + 63) // 64 * 64
    bn = 32 if c['cout'] <= 32 else 64
    wp = torch.randn(bn, kpad, device='cuda').half() * 0.05
    ho, wo = c['h'] + 2 * c['pad'] - 2, c['w'] + 2 * c['pad'] - 2
    y = torch.empty(b, ho, wo, c['cout'], device='cuda', dtype=torch.float16)
    d = FpConvDesc(batch=b, c0=c['c'], h0=c['h'], w0=c['w'], up0=0, c1=0, kh=3, kw=3, stride=1, pad_t=c['pad'], pad_l=c['pad'],
                   pad_mode=c['mode'], ho=ho, wo=wo, cout=c['cout'], kpad=kpad, npad=bn, bn=bn, act=2, dtype=0, out_f32=0,
                   out_ld=c['cout'], res_ld=0, b_via_tma=1, w_dtype_plus1=0)
    for flags in (0, 1, 2, 4, 3, 5, 6, 7):
        os.environ['FP_DEBUG_FLAGS'] = str(flags)
        lib.fp_reload_options()
        for _ in range(2):
            _lib.check(lib.fp_conv_fwd(ctypes.byref(d), P(x), None, P(wp), None, None, None, P(y), None))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            _lib.check(lib.fp_conv_fwd(ctypes.byref(d), P(x), None, P(wp), None, None, None, P(y), None))
        e1.record(); torch.cuda.synchronize()
        print(f'{name:34s} flags {flags} (skip: {"loads " if flags & 1 else ""}{"mma " if flags & 2 else ""}{"epilogue" if flags & 4 else ""}): {e0.elapsed_time(e1) / 5:.3f} ms')
