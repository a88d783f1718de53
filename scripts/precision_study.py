"""CPU study (test infrastructure): which layers must keep more than 16-bit operand precision for every output
scale to land within 1e-3 of the fp32 reference.  Emulates the CUDA path's roundings on the oracle: per layer,
operands (input activation as stored, weights) are rounded to `prec(layer)`; accumulation is fp32.
  prec: 'f16' | 'bf16' | 'split' (hi+lo pair of the 16-bit type ~ 22/16 significant bits) | 'f32'
Usage: python scripts/precision_study.py [trained|random] [eval|train] [B] [H] [W]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from footprints_b200.synthetic import synthetic_batch, synthetic_state_dict

torch.set_num_threads(8)


def rnd(x, p, base):
    if p == 'f32':
        return x
    dt = torch.float16 if base == 'f16' else torch.bfloat16
    hi = x.to(dt).float()
    if p == 'split':
        lo = (x - hi).to(dt).float()
        return hi + lo
    return hi


def forward(sd, image, prec, base, training=False):
    """prec(name) -> storage/operand precision of layer `name`'s OUTPUT and weights."""
    def conv(name, x, w, b=None, stride=1, padding=0):
        p = prec(name)
        return F.conv2d(x, rnd(w, p, base), b, stride=stride, padding=padding)

    def bn(prefix, x):
        return F.batch_norm(x, sd[f'{prefix}.running_mean'].clone(), sd[f'{prefix}.running_var'].clone(),
                            sd[f'{prefix}.weight'], sd[f'{prefix}.bias'], training, 0.1, 1e-5)

    def st(name, x):       # store the layer output
        return rnd(x, prec(name), base)

    x = st('stem.in', (image - 0.45) / 0.225)
    x = st('stem', F.relu(bn('encoder.layer0.1', conv('stem', x, sd['encoder.layer0.0.weight'], None, 2, 3))))
    feats = [x]
    x = F.max_pool2d(x, 3, 2, 1)
    for li, nb in enumerate((3, 4, 6, 3)):
        base_ = 'encoder.layer1.1' if li == 0 else f'encoder.layer{li + 1}'
        for b in range(nb):
            p = f'{base_}.{b}'
            n = f'enc{li + 1}.{b}'
            stride = 2 if (b == 0 and li > 0) else 1
            idt = x
            z = conv(n + '.conv1', x, sd[f'{p}.conv1.weight'], None, stride, 1)
            if training:
                z = st(n + '.conv1', z)
            out = st(n + '.conv1', F.relu(bn(f'{p}.bn1', z)))
            z = conv(n + '.conv2', out, sd[f'{p}.conv2.weight'], None, 1, 1)
            if training:
                z = st(n + '.conv2', z)
            out = bn(f'{p}.bn2', z)
            if f'{p}.downsample.0.weight' in sd:
                z = conv(n + '.down', x, sd[f'{p}.downsample.0.weight'], None, 2, 0)
                if training:
                    z = st(n + '.down', z)
                idt = st(n + '.down', bn(f'{p}.downsample.1', z))
            x = st(n + '.conv2', F.relu(out + idt))
        feats.append(x)

    def cblock(p, n, x, names):
        for i, nm in zip((1, 2), names):
            x = F.pad(x, (1, 1, 1, 1), mode='reflect')
            x = st(n + nm, F.elu(conv(n + nm, x, sd[f'{p}.conv{i}.weight'], sd[f'{p}.conv{i}.bias'])))
        return x

    outs = {}
    for dname, dec, sig in (('mask', 'mask_decoder', False), ('depth', 'depth_decoder', True)):
        x = feats[4]
        o = {}
        for bi in range(1, 5):
            x = cblock(f'{dec}.block{bi}.pre_concat_conv', f'{dname}.block{bi}', x, ('.pre1', '.pre2'))
            x = F.interpolate(x, scale_factor=2, mode='nearest')
            x = torch.cat([x, feats[4 - bi]], 1)
            x = cblock(f'{dec}.block{bi}.post_concat_conv', f'{dname}.block{bi}', x, ('.post1', '.post2'))
            if bi >= 2:
                h = F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), sd[f'{dec}.outconv{bi - 1}.conv1.weight'],
                             sd[f'{dec}.outconv{bi - 1}.conv1.bias'])
                if sig:
                    h = torch.sigmoid(h)
                o[bi - 1] = F.interpolate(h, scale_factor=(16 >> (bi - 1)), mode='bilinear', align_corners=False)
        x = F.interpolate(x, scale_factor=2, mode='nearest')
        x = cblock(f'{dec}.outconv4.0', f'{dname}.o4', x, ('.c1', '.c2'))
        h = F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), sd[f'{dec}.outconv4.1.conv1.weight'], sd[f'{dec}.outconv4.1.conv1.bias'])
        if sig:
            h = torch.sigmoid(h)
        o[4] = h
        outs[dname] = o
    return {k: torch.cat([outs['mask'][i], outs['depth'][i]], 1) for i, k in enumerate(('1/8', '1/4', '1/2', '1/1'), start=1)}


def deep(name, groups):
    return any(name.startswith(g) for g in groups)


if __name__ == '__main__':
    kind = sys.argv[1] if len(sys.argv) > 1 else 'trained'
    mode = sys.argv[2] if len(sys.argv) > 2 else 'eval'
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    H = int(sys.argv[4]) if len(sys.argv) > 4 else 192
    W = int(sys.argv[5]) if len(sys.argv) > 5 else 640
    sd = synthetic_state_dict('resnet34', 0, trained_like=(kind == 'trained'))
    img = synthetic_batch(B, H, W, 1)['image']
    tr = mode == 'train'
    with torch.no_grad():
        ref = forward(sd, img, lambda n: 'f32', 'f16', tr)
        print('output ranges', {k: (float(v[:, :2].abs().max()), float(v[:, 2:].abs().max())) for k, v in ref.items()})
        G = {
            'none': (),
            'L34+B12': ('enc3', 'enc4', 'mask.block1', 'mask.block2', 'depth.block1', 'depth.block2'),
            'L234+B123': ('enc2', 'enc3', 'enc4', 'mask.block1', 'mask.block2', 'mask.block3', 'depth.block1', 'depth.block2', 'depth.block3'),
            'enc+B12': ('stem', 'enc', 'mask.block1', 'mask.block2', 'depth.block1', 'depth.block2'),
            'enc+B123': ('stem', 'enc', 'mask.block1', 'mask.block2', 'mask.block3', 'depth.block1', 'depth.block2', 'depth.block3'),
            'mask-only L34+B12': ('enc3', 'enc4', 'mask.block1', 'mask.block2'),
            'all': ('',),
        }
        for base in ('f16', 'bf16'):
            for gname, groups in G.items():
                out = forward(sd, img, lambda n: 'split' if (groups and deep(n, groups)) else base, base, tr)
                errs = {k: float((out[k] - ref[k]).abs().max()) for k in ref}
                print(f'{kind:8s} {mode:5s} {base:5s} split={gname:18s}', ' '.join(f'{k}={v:.2e}' for k, v in errs.items()), flush=True)
