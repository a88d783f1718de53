"""Discrete-event simulation of the mbarrier protocols of the two opt-in kernels (conv_rowstream_kernel,
conv_wgrad_roll_kernel).  The kernels cannot run here (no GPU), but their producer / MMA-issuer / epilogue roles
synchronise only through mbarriers with one phase bit each, so liveness (no deadlock) and safety (a consumer never
reads a buffer before the matching fill landed, a producer never overwrites a buffer that is still being read) can be
checked by replaying the same wait / arrive / commit sequences under random interleavings and random completion delays
of the asynchronous operations (TMA, cp.async, tcgen05.commit).

mbarrier model: `pending` arrivals per phase; the phase bit flips when they reach zero; `wait(parity)` passes when the
phase with that parity has completed (a fresh barrier lets `wait(1)` through, as on hardware)."""
import random

import pytest


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, 'more arrivals than the barrier was initialised for'
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def passed(self, parity):
        return self.phase != parity


class Sim:
    """Roles are generators yielding ('wait', bar, parity) | ('arrive', bar) | ('async', bar, action) | ('do', fn)."""

    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.roles, self.pending_async = [], []          # async completions: per-issuer FIFO (completion order = issue order)

    def add(self, name, gen):
        self.roles.append([name, gen, None])

    def run(self, max_steps=2_000_000):
        for _ in range(max_steps):
            live = [r for r in self.roles if r[1] is not None]
            if not live and not self.pending_async:
                return
            choices = []
            for r in live:
                if r[2] is None or r[2][1].passed(r[2][2]):
                    choices.append(('role', r))
            for q in self.pending_async:
                if q:
                    choices.append(('async', q))
            assert choices, 'deadlock: ' + ', '.join(f'{r[0]} waits' for r in live)
            kind, obj = self.rng.choice(choices)
            if kind == 'async':
                bar, action = obj.pop(0)
                if action:
                    action()
                bar.arrive()
                self.pending_async = [q for q in self.pending_async if q]
                continue
            obj[2] = None
            try:
                op = next(obj[1])
            except StopIteration:
                obj[1] = None
                continue
            if op[0] == 'wait':
                obj[2] = op
            elif op[0] == 'arrive':
                op[1].arrive()
            elif op[0] == 'async':                            # completes later, in issue order per queue
                op[3].append((op[1], op[2]))
                if not any(q is op[3] for q in self.pending_async):    # identity, not list equality
                    self.pending_async.append(op[3])
            elif op[0] == 'do':
                op[1]()
        raise AssertionError('simulation did not finish')


# ----------------------------------------------------------------------------------------------- row-streaming conv
def rowstream_sim(seed, items_per_cta, stages, H, cblocks, tma_mode):
    sim = Sim(seed)
    n_prod = 1 if tma_mode else 4                             # stands for 256 producer threads
    full = [Bar(1 if tma_mode else n_prod) for _ in range(stages)]
    empty = [Bar(1) for _ in range(stages)]
    row_full = [Bar(1) for _ in range(2 * H)]
    row_empty = [Bar(2) for _ in range(2 * H)]                # stands for 8 epilogue warps
    content = [None] * stages                                 # what the strip buffer holds / is being read
    reading = [False] * stages
    slot_state = ['free'] * (2 * H)                           # free -> open(item) -> done(item) -> free

    def producer(pid):
        s, ph = 0, 0
        q = []
        for it in range(items_per_cta):
            for j in range(H + 2):
                for cb in range(cblocks):
                    yield ('wait', empty[s], ph ^ 1)
                    tag = (it, j, cb)

                    def land(s=s, tag=tag):
                        assert not reading[s], 'strip overwritten while the MMAs still read it'
                        content[s] = tag
                    yield ('async', full[s], land if pid == 0 else None, q)
                    s += 1
                    if s == stages:
                        s, ph = 0, ph ^ 1

    def mma():
        s, ph = 0, 0
        q = []                                                # tcgen05.commit completions (in order)
        for it in range(items_per_cta):
            half, use_par = it & 1, (it >> 1) & 1
            for j in range(H + 2):
                dy_lo = j - (H - 1) if j > H - 1 else 0
                q_hi = (H - 1) - (j - dy_lo)
                if dy_lo == 0:
                    slot = half * H + q_hi
                    yield ('wait', row_empty[slot], use_par ^ 1)

                    def open_(slot=slot, it=it):
                        assert slot_state[slot] == 'free', f'row slot {slot} reopened before the epilogue drained it'
                        slot_state[slot] = ('open', it)
                    yield ('do', open_)
                for cb in range(cblocks):
                    yield ('wait', full[s], ph)

                    def start(s=s, tag=(it, j, cb)):
                        assert content[s] == tag, f'MMA expected strip {tag}, buffer holds {content[s]}'
                        reading[s] = True
                    yield ('do', start)

                    def done(s=s):
                        reading[s] = False
                    yield ('async', empty[s], done, q)
                    s += 1
                    if s == stages:
                        s, ph = 0, ph ^ 1
                if j >= 2:
                    slot = half * H + (H - 1) - (j - 2)

                    def complete(slot=slot, it=it):
                        assert slot_state[slot] == ('open', it)
                        slot_state[slot] = ('done', it)
                    yield ('async', row_full[slot], complete, q)

    def epilogue(eid):
        for it in range(items_per_cta):
            half, use_par = it & 1, (it >> 1) & 1
            for r in range(H):
                slot = half * H + (H - 1 - r)
                yield ('wait', row_full[slot], use_par)

                def read(slot=slot, it=it):
                    assert slot_state[slot][1] == it and slot_state[slot][0] in ('done', 'draining')
                    slot_state[slot] = ('draining', it)
                yield ('do', read)

                def arrive_and_free(slot=slot):
                    row_empty[slot].arrive()
                    if row_empty[slot].pending == row_empty[slot].count:     # this arrival completed the phase
                        slot_state[slot] = 'free'
                yield ('do', arrive_and_free)

    for pid in range(n_prod):
        sim.add(f'producer{pid}', producer(pid))
    sim.add('mma', mma())
    sim.add('epi0', epilogue(0))
    sim.add('epi1', epilogue(1))
    sim.run()


@pytest.mark.parametrize('tma_mode', [True, False])
@pytest.mark.parametrize('stages,H,cblocks', [(3, 4, 1), (8, 8, 1), (4, 4, 2), (5, 8, 2)])
def test_rowstream_barrier_protocol(stages, H, cblocks, tma_mode):
    for seed in range(6):
        rowstream_sim(seed, items_per_cta=5, stages=stages, H=H, cblocks=cblocks, tma_mode=tma_mode)


# ----------------------------------------------------------------------------------------------- rolling weight gradient
def wgrad_roll_sim(seed, cols, Ho, x_stages, ncb):
    sim = Sim(seed)
    x_full = [Bar(2) for _ in range(x_stages)]                # thread 0 + "the other 255" producers
    x_empty = [Bar(1) for _ in range(x_stages)]
    dz_full = [Bar(1) for _ in range(5)]
    dz_empty = [Bar(1) for _ in range(5)]
    ring = [None] * 5                                         # (col, y) held by each dZ slot
    ring_busy = [0] * 5                                       # outstanding MMA reads per slot
    strips = [None] * x_stages
    strips_busy = [False] * x_stages

    def producer(is_tid0):
        xs, empty_ph = 0, 0
        q = []
        for col in range(cols):
            for y in range(-2, Ho + 2):
                if is_tid0:
                    slot = (y + 4) & 3
                    loads = [slot] + ([4] if slot == 0 and y >= 0 else [])
                    for sl in loads:
                        yield ('wait', dz_empty[sl], ((empty_ph >> sl) & 1) ^ 1)
                        empty_ph ^= 1 << sl

                        def land(sl=sl, tag=(col, y)):
                            assert ring_busy[sl] == 0, f'dZ slot {sl} overwritten while MMAs still read it'
                            ring[sl] = tag
                        yield ('async', dz_full[sl], land, q)
                if y < 0:
                    continue
                for j in range(ncb):
                    s = xs % x_stages
                    yield ('wait', x_empty[s], ((xs // x_stages) & 1) ^ 1)

                    def land_x(s=s, tag=(col, y, j)):
                        assert not strips_busy[s]
                        strips[s] = tag
                    yield ('async', x_full[s], land_x if is_tid0 else None, q)
                    xs += 1

    def mma():
        xs, full_ph = 0, 0
        q = []
        for col in range(cols):
            for y in range(Ho + 2):
                waits = ([2, 3] if y == 0 else []) + [y & 3] + ([4] if (y & 3) == 0 else [])
                for sl in waits:
                    yield ('wait', dz_full[sl], (full_ph >> sl) & 1)
                    full_ph ^= 1 << sl
                s_pair, s_single = (y + 3) & 3, (y + 2) & 3

                def check(col=col, y=y, s_pair=s_pair, s_single=s_single):
                    assert ring[s_pair] == (col, y - 1), (ring[s_pair], (col, y - 1))
                    assert ring[s_pair + 1] == (col, y), (ring[s_pair + 1], (col, y))
                    assert ring[s_single] == (col, y - 2), (ring[s_single], (col, y - 2))
                    for sl in (s_pair, s_pair + 1, s_single):
                        ring_busy[sl] += 1
                yield ('do', check)
                for j in range(ncb):
                    s = xs % x_stages
                    yield ('wait', x_full[s], (xs // x_stages) & 1)

                    def start(s=s, tag=(col, y, j)):
                        assert strips[s] == tag, (strips[s], tag)
                        strips_busy[s] = True
                    yield ('do', start)

                    def done(s=s):
                        strips_busy[s] = False
                    yield ('async', x_empty[s], done, q)
                    xs += 1

                def release_reads(s_pair=s_pair, s_single=s_single):
                    for sl in (s_pair, s_pair + 1, s_single):
                        ring_busy[sl] -= 1
                # tcgen05.commit completes after this row's MMAs: first the reads end, then the slots are handed back
                yield ('async', dz_empty[(y + 2) & 3], release_reads, q)
                if (y & 3) == 0:
                    yield ('async', dz_empty[4], None, q)
            yield ('async', dz_empty[Ho & 3], None, q)
            yield ('async', dz_empty[(Ho + 1) & 3], None, q)

    sim.add('tid0', producer(True))
    sim.add('others', producer(False))
    sim.add('mma', mma())
    sim.run()


@pytest.mark.parametrize('Ho', [5, 8, 11])
@pytest.mark.parametrize('x_stages,ncb', [(5, 1), (6, 2), (2, 1)])
def test_wgrad_roll_barrier_protocol(Ho, x_stages, ncb):
    for seed in range(6):
        wgrad_roll_sim(seed, cols=3, Ho=Ho, x_stages=x_stages, ncb=ncb)
