#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: images/sec of the Footprints training step (forward with train-mode
BatchNorm + fused loss + backward + Adam) on synthetic KITTI-shaped input (192x640), batch 64 per GPU,
ResNet-34 encoder + two skip decoders, random-init weights.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched under torch.distributed.run)
    python bench.py --impl reference ...                   (the reference algorithm's CPU path, host cores)

One JSON line on rank 0.  `value` = whole-job images/s with inputs resident in HBM; `e2e` = the same through
host (pinned) buffers with the H2D copy of every step's inputs and a D2H read of the loss inside the timed
region; `roofline` = the tcgen05 implicit-GEMM convolution kernel's achieved algorithmic TFLOP/s (CUDA events
around every launch of one extra, untimed-for-`value` step) against the measured bf16 peak; `cpu_baseline` = the
oracle (a port of the reference's PyTorch path) timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'images/sec fwd+bwd KITTI 192x640 bs=64/GPU'
H, W = 192, 640


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='images per GPU')
    ap.add_argument('--hw', default='192x640', help='input height x width (KITTI 192x640; Matterport 512x640 with --batch 16)')
    ap.add_argument('--cpu-batch', type=int, default=4, help='fixed batch of the CPU reference arm / cpu_baseline sample')
    ap.add_argument('--arch', default='resnet34', choices=['resnet34', 'resnet18'])
    ap.add_argument('--mode', default='train', choices=['train', 'infer'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--precise', action='store_true', help='split-precision layers upstream of the low-resolution mask heads (all four output scales within 1e-3 of the fp32 reference)')
    ap.add_argument('--no-precise', action='store_true', help='skip the extra precise-mode measurement')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from the host instead of replaying the CUDA graph of the step')
    args = ap.parse_args()
    global H, W, METRIC
    H, W = (int(v) for v in args.hw.lower().split('x'))
    if (H, W, args.batch) != (192, 640, 64):
        METRIC = f'images/sec fwd+bwd {H}x{W} bs={args.batch}/GPU'
    return args


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i',
                                          str(self.index), '-lms', '200'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        # "under load" = samples in the upper half of what was seen (the sampler also sees the idle edges)
        load = [v for v in sm if v >= 0.5 * sm[-1]] if sm else []
        return {'sm_mhz': load[len(load) // 2] if load else None,
                'sm_max_mhz': int(self.rows[0][1]) if self.rows and self.rows[0][1].isdigit() else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def usable_cores():
    """Host cores this process may actually use: affinity mask, capped by the cgroup CPU quota when there is one
    (a container can see 128 CPUs and own 8; spawning 128 threads then makes the baseline meaninglessly slow)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()
        if quota != 'max':
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except (OSError, ValueError):
        pass
    return n


# ----------------------------------------------------------------------------------------------- reference arm
def _reference_modules():
    """The UNMODIFIED reference (baseline/_ref placed by scripts/install_reference.sh; SURVEY §8c shims in-process), or None."""
    try:
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        import reference_shims
        if reference_shims.import_reference() is None:
            return None
        import footprints.network as fn
        from footprints.training.losses import LossManager
        return fn, LossManager
    except Exception:
        return None


def cpu_train_steps(arch, batch, steps, warmup, threads):
    """The reference's training step on host cores (footprints/training/train.py:153-156,224-225): forward in train mode,
    LossManager, backward, Adam.  Runs the reference's OWN modules when the unmodified package is importable
    (kind 'reference'), else the oracle port of the same arithmetic (kind 'port').  Returns (seconds per step, kind)."""
    import torch
    from footprints_b200.synthetic import synthetic_batch, synthetic_state_dict
    torch.set_num_threads(threads)
    data = synthetic_batch(batch, H, W, 1)
    ref = _reference_modules() if arch == 'resnet34' else None
    if ref is not None:
        fn, LossManager = ref
        torch.manual_seed(0)
        net = fn.FootprintNetwork(pretrained=False).train()
        lm = LossManager((0.1, 100.0), 0.25)
        opt = torch.optim.Adam(net.parameters(), lr=1e-4)
        targets = {k: v for k, v in data.items() if k != 'image'}

        def one():
            losses = lm(net(data['image']), targets)
            net.zero_grad()
            losses['loss'].backward()
            opt.step()
        kind = 'reference'
    else:
        from oracle import footprints_oracle as O
        sd = synthetic_state_dict(arch, 0, trained_like=False)
        params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(('running_mean', 'running_var'))
                      else v.clone()) for k, v in sd.items()}
        opt = torch.optim.Adam([v for v in params.values() if v.is_floating_point() and v.requires_grad], lr=1e-4)

        def one():
            losses = O.loss_forward(O.network_forward(params, data['image'], arch, True, {}), data)
            opt.zero_grad(set_to_none=True)
            losses['loss'].backward()
            opt.step()
        kind = 'port'
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), kind


def run_reference(args):
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = usable_cores()
    # bounded sample with a FIXED batch (the same at every N, so the driver's ratio has a stable denominator)
    per_step = args.cpu_batch
    steps = max(1, min(args.steps, 4))
    t, kind = cpu_train_steps(args.arch, per_step, steps, 1, threads)
    value = per_step / t
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': 1, 'ms_per_step': t * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'training step (fwd train-BN + loss + bwd + Adam), {args.arch} encoder + 2 skip decoders, '
                                   f'{H}x{W}; CPU sample of {per_step} image(s)/step', 'arch': args.arch},
            'cpu_baseline': {'value': value, 'unit': 'images/s', 'cores': threads, 'kind': kind,
                             'sample': f'{steps} steps of batch {per_step} at {H}x{W}, torch {torch.__version__} CPU, '
                                       + ('the unmodified reference modules (baseline/_ref)' if kind == 'reference' else 'oracle port')},
            'e2e': {'value': value, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- ours
def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from footprints_b200 import FootprintNetwork
    from footprints_b200.synthetic import synthetic_batch
    from footprints_b200.trainer import Trainer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        pg = dist.group.WORLD
    blocks = {'resnet34': (3, 4, 6, 3), 'resnet18': (2, 2, 2, 2)}[args.arch]
    torch.manual_seed(0)
    net = FootprintNetwork(pretrained=False, blocks=blocks, precise=args.precise).cuda()
    B = args.batch
    data = synthetic_batch(B, H, W, seed=1 + rank)
    host = {k: v.pin_memory() for k, v in data.items()}
    dev = {k: v.cuda() for k, v in data.items()}
    tkeys = [k for k in data if k != 'image']

    if args.mode == 'train':
        net.train()
        trainer = Trainer(net, lr=1e-4, process_group=pg)
        trainer.broadcast_parameters()
        if not args.no_graph:
            trainer.enable_cuda_graph()
        eng = trainer.eng

        def step_resident():
            return trainer.step(dev['image'], {k: dev[k] for k in tkeys})
    else:
        net.eval()
        eng = net.engine()

        def step_resident():
            return net(dev['image'])['1/1']

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
            dist.barrier()
        return ms

    for _ in range(max(3, args.warmup)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    ms = timed(step_resident, args.steps)
    launches = (eng.launches - l0)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- e2e: host (pinned) inputs -> device every step, loss read back every step
    copy_stream = torch.cuda.Stream()
    bufs = [{k: torch.empty_like(dev[k]) for k in dev} for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    h2d = sum(v.numel() * v.element_size() for k, v in host.items() if args.mode == 'train' or k == 'image')

    def upload(i):
        with torch.cuda.stream(copy_stream):
            for k in (host if args.mode == 'train' else ['image']):
                bufs[i][k].copy_(host[k], non_blocking=True)
            ready[i].record(copy_stream)

    state = {'i': 0}
    d2h = 21 * 4 if args.mode == 'train' else B * 4 * H * W * 4
    host_out = torch.empty((B, 4, H, W), dtype=torch.float32).pin_memory() if args.mode == 'infer' else None

    def step_e2e():
        i = state['i']
        torch.cuda.current_stream().wait_event(ready[i])
        upload(1 - i)                                   # prefetch the next step's batch while this one computes
        if args.mode == 'train':
            out = trainer.step(bufs[i]['image'], {k: bufs[i][k] for k in tkeys})
            out.cpu()                                   # D2H read of the step's losses (train.py's evaluator does this)
        else:
            host_out.copy_(net(bufs[i]['image'])['1/1'], non_blocking=True)   # predict_simple.py:68 (pinned dst)
            torch.cuda.current_stream().synchronize()
        copy_stream.wait_stream(torch.cuda.current_stream())
        state['i'] = 1 - i

    upload(0)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e_value = world * B / (ms_e2e * 1e-3)

    # ---- inference only: the caller's whole loop (evaluation/inference.py:95-110) with BOTH edges on the GPU: pinned uint8
    # camera frames -> H2D -> ResizeToTensor (Pillow-exact Lanczos + ToTensor) -> network -> sigmoid + float16 -> D2H
    e2e_frames = None
    if args.mode == 'infer' and world == 1:
        from footprints_b200.preprocess import ResizeToTensor
        fh, fw = (375, 1242) if (H, W) == (192, 640) else (2 * H, 2 * W)
        frames_host = torch.randint(0, 256, (B, fh, fw, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).pin_memory()
        fbufs = [torch.empty((B, fh, fw, 3), dtype=torch.uint8, device=bufs[0]['image'].device) for _ in range(2)]
        resize = ResizeToTensor((H, W))
        out16 = torch.empty((B, 4, H, W), dtype=torch.float16).pin_memory()

        def upload_frames(i):
            with torch.cuda.stream(copy_stream):
                fbufs[i].copy_(frames_host, non_blocking=True)
                ready[i].record(copy_stream)

        def step_frames():
            i = state['i']
            torch.cuda.current_stream().wait_event(ready[i])
            upload_frames(1 - i)
            out16.copy_(net.predict_for_saving(resize(fbufs[i])), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            copy_stream.wait_stream(torch.cuda.current_stream())
            state['i'] = 1 - i

        torch.cuda.synchronize()
        state['i'] = 0
        upload_frames(0)
        for _ in range(2):
            step_frames()
        ms_frames = timed(step_frames, args.steps) / args.steps
        e2e_frames = {'value': B / (ms_frames * 1e-3), 'unit': 'images/s', 'ms_per_step': ms_frames,
                      'h2d_bytes_per_step': frames_host.numel(), 'd2h_bytes_per_step': out16.numel() * 2,
                      'what': f'pinned uint8 {fh}x{fw} frames -> ResizeToTensor -> forward -> sigmoid + float16 [B,4,{H},{W}] read back'}

    # ---- roofline of the dominant kernel class (tcgen05 implicit-GEMM conv: forward + data-gradient launches)
    roofline = None
    # every rank runs the instrumented step (it contains the gradient all-reduce); only rank 0 records events
    eng.profile = []          # on EVERY rank: an instrumented step runs eagerly (bucketed collectives), a graph replay does not
    overlap = getattr(eng, 'overlap_wgrad', False)
    if hasattr(eng, 'overlap_wgrad'):
        eng.overlap_wgrad = False          # per-launch CUDA-event times need the launches serialised on one stream
    step_resident()
    torch.cuda.synchronize()
    if hasattr(eng, 'overlap_wgrad'):
        eng.overlap_wgrad = overlap
    prof, eng.profile = eng.profile, None
    if rank == 0:
        agg = {}
        for kind, name, flops, a, b, _ in prof:
            t, f, n = agg.get(kind, (0.0, 0.0, 0))
            agg[kind] = (t + a.elapsed_time(b), f + flops, n + 1)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except OSError:
            pass
        peak = peaks.get('bf16_tflops_sustained', 1400.0)
        t_igemm = sum(agg[k][0] for k in ('fwd', 'dgrad') if k in agg)
        f_igemm = sum(agg[k][1] for k in ('fwd', 'dgrad') if k in agg)
        n_igemm = sum(agg[k][2] for k in ('fwd', 'dgrad') if k in agg)
        ach = f_igemm / (t_igemm * 1e-3) / 1e12 if t_igemm > 0 else 0.0
        # longest single launch of that class, timed live; its DRAM traffic comes from the committed `ncu --set full`
        # capture of the same launch (profiles/r01_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum)
        top = max((x for x in prof if x[0] in ('fwd', 'dgrad')), key=lambda x: x[3].elapsed_time(x[4]), default=None)
        traffic, top_launch = None, None
        if top is not None:
            t_top = top[3].elapsed_time(top[4])
            top_launch = {'layer': top[1], 'kind': top[0], 'ms': t_top, 'tflops': top[2] / (t_top * 1e-3) / 1e12}
            try:        # dram__bytes_read.sum + dram__bytes_write.sum of this launch from this round's `ncu --set full`
                tj = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))     # capture; null when absent
                key = top[1].split('.', 1)[-1] + '/' + top[0]          # 'mask.o4.c1' -> 'o4.c1/fwd'
                if B == 64 and (H, W) == (192, 640) and key in tj['bytes_per_launch']:
                    traffic = tj['bytes_per_launch'][key]
            except (OSError, KeyError, ValueError):
                pass
        roofline = {'bound': 'tensor', 'kernel': 'conv_igemm_kernel + conv_igemm_strip_kernel (tcgen05 implicit GEMM; all fwd + dgrad launches of a step)',
                    'achieved': ach, 'peak': peak,
                    'unit': 'TFLOP/s', 'frac': ach / peak,
                    'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback (B200_PROFILING.md, sustained)',
                    'traffic': traffic, 'traffic_of': top_launch, 'launches': n_igemm, 'ms_per_step_in_kernel': t_igemm,
                    'algorithmic_gflop_per_step': f_igemm / 1e9,
                    'by_kind': {k: {'ms': v[0], 'tflops': (v[1] / (v[0] * 1e-3) / 1e12 if v[0] > 0 else 0.0), 'launches': v[2]}
                                for k, v in agg.items()}}

        # second view, the bound SURVEY §8d assigns to the fused upsample/concat/pad decoder convolutions: algorithmic
        # HBM bytes of the decoders' forward launches (inputs once pre-upsample/pre-pad + skip + outputs + weights,
        # 16-bit) / their summed CUDA-event time, against the measured HBM copy bandwidth
        try:
            dec = [x for x in prof if x[0] == 'fwd' and x[1].split('.', 1)[0] in ('mask', 'depth') and x[5] > 0]
            t_dec = sum(x[3].elapsed_time(x[4]) for x in dec)
            b_dec = float(sum(x[5] for x in dec))
            hbm_peak = peaks.get('hbm_gbs', 6500.0)
            if t_dec > 0:
                roofline['decoder_fwd_hbm'] = {'bound': 'hbm', 'achieved': b_dec / (t_dec * 1e-3) / 1e9, 'peak': hbm_peak,
                                               'unit': 'GB/s', 'frac': b_dec / (t_dec * 1e-3) / 1e9 / hbm_peak,
                                               'algorithmic_gb_per_step': b_dec / 1e9, 'ms_per_step': t_dec,
                                               'launches': len(dec)}
        except Exception as exc:               # never lose the bench line over the secondary view
            roofline['decoder_fwd_hbm'] = {'error': repr(exc)}

    # top-level copies of the per-class roofline views (the driver's parser keeps top-level keys)
    extra = {}
    if rank == 0 and roofline is not None:
        def view(rows, label):
            t = sum(x[3].elapsed_time(x[4]) for x in rows)
            fl = sum(x[2] for x in rows)
            if t <= 0:
                return None
            return {'bound': 'tensor', 'kernel': label, 'achieved': fl / (t * 1e-3) / 1e12, 'peak': peak, 'unit': 'TFLOP/s',
                    'frac': fl / (t * 1e-3) / 1e12 / peak, 'ms_per_step': t, 'launches': len(rows),
                    'algorithmic_gflop_per_step': fl / 1e9}
        extra['roofline_wgrad'] = view([x for x in prof if x[0] == 'wgrad'], 'weight-gradient kernels (conv_wgrad_roll / _strip / general + reduce)')
        extra['roofline_encoder_fwd'] = view([x for x in prof if x[0] == 'fwd' and x[1].startswith(('stem', 'enc'))],
                                             'encoder forward convolutions (stem + 36 BasicBlock convolutions)')
        extra['roofline_decoder_fwd'] = view([x for x in prof if x[0] == 'fwd' and x[1].startswith(('mask', 'depth'))],
                                             'decoder forward convolutions (tensor-pipe view)')
        extra['roofline_decoder_hbm'] = roofline.get('decoder_fwd_hbm')
        conv_ms = sum(v[0] for v in agg.values())
        extra['step_breakdown'] = {'conv_kernels_ms': conv_ms, 'other_kernels_ms': max(0.0, ms_per_step - conv_ms),
                                   'note': 'conv = CUDA-event time of every fwd/dgrad/wgrad launch of the instrumented step'}
        flops_step = sum(v[1] for v in agg.values())
        extra['roofline_whole_step'] = {'bound': 'tensor', 'achieved': flops_step / (ms_per_step * 1e-3) / 1e12, 'peak': peak,
                                        'unit': 'TFLOP/s', 'frac': flops_step / (ms_per_step * 1e-3) / 1e12 / peak}

    # precise mode (split precision upstream of the low-resolution mask heads: all four output scales within 1e-3 of the
    # fp32 reference) measured beside the default single-pass mode, same workload
    if rank == 0 and world == 1 and not args.precise and not args.no_precise:
        try:
            del step_resident
            torch.cuda.empty_cache()
            net2 = FootprintNetwork(pretrained=False, blocks=blocks, precise=True).cuda()
            if args.mode == 'train':
                net2.train()
                tr2 = Trainer(net2, lr=1e-4).enable_cuda_graph()
                fn2 = lambda: tr2.step(dev['image'], {k: dev[k] for k in tkeys})
            else:
                net2.eval()
                fn2 = lambda: net2(dev['image'])['1/1']
            for _ in range(3):
                fn2()
            ms2 = timed(fn2, max(3, args.steps // 2)) / max(3, args.steps // 2)
            extra['precise'] = {'value': B / (ms2 * 1e-3), 'unit': 'images/s', 'ms_per_step': ms2,
                                'note': 'FootprintNetwork(precise=True): every output scale within 1e-3 max-abs of the fp32 reference (tests)'}
        except Exception as exc:
            extra['precise'] = {'error': repr(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.mode == 'train':
        threads = usable_cores()
        bs = args.cpu_batch
        t, kind = cpu_train_steps(args.arch, bs, 2, 1, threads)
        cpu = {'value': bs / t, 'unit': 'images/s', 'cores': threads, 'kind': kind,
               'sample': f'2 training steps of batch {bs} at {H}x{W}, fp32, '
                         + ('the unmodified reference modules (baseline/_ref)' if kind == 'reference' else 'oracle port of the reference path')}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': max(3, args.warmup), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f16', 'data': 'synthetic',
                'config': {'workload': ('training step: fwd (train-mode BN) + fused BCE/log-L1 loss + bwd + Adam' if args.mode == 'train'
                                        else 'inference forward') + f', {args.arch} encoder + 2 skip decoders, {H}x{W}, '
                                                                    f'batch {B}/GPU, random-init weights',
                           'arch': args.arch, 'global_batch': B * world, 'parallelism': f'dp{world}',
                           'l2': 'per-step working set (>10 GB of activations) far exceeds the 126 MB L2; no flush needed'},
                'roofline': roofline, 'cpu_baseline': cpu,
                'e2e': {'value': e2e_value, 'unit': 'images/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': h2d,
                        'd2h_bytes_per_step': d2h},
                'gpu_launches': launches, 'clocks': clocks}
        line['config']['precise'] = bool(args.precise)
        line['config']['cuda_graph'] = bool(args.mode == 'train' and not args.no_graph)
        line.update(extra)
        if e2e_frames is not None:
            line['e2e_from_uint8_frames'] = e2e_frames
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
