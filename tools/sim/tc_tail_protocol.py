#!/usr/bin/env python
"""Randomised discrete-event model of the mbarrier protocol of deepfm_fwd_fused_kernel<LA, ., TCTAIL=true>
(rec_pangu_b200/csrc/deepfm_fused.cu): weight producer, MMA issuer (main loop + interleaved tail steps), asynchronous
tensor pipe, gather/split warps, epilogue warps.  It checks, under random interleavings,

  * liveness: every actor terminates (no deadlock) for 1..5 tiles per CTA, 1..3 tail layers, any k-block count;
  * the phase/parity arithmetic of every barrier (a wait that passes on a stale phase shows up as a hazard below);
  * data hazards at the moment the tensor pipe EXECUTES an MMA (not when it is issued): the layer-1 operand slot holds
    the k-block it should, the tail operand holds (tile, layer) completely written, an accumulator buffer is not
    overwritten while an epilogue warp still has to read its previous contents, and the epilogue reads what it expects.

It models the protocol, not the arithmetic.  Run: python tools/sim/tc_tail_protocol.py [--seeds N]
"""
import argparse
import random

LB, OPN, FM_BUF, N_EPI, N_GATHER = 3, 2, 4, 8, 4


class Bar:
    def __init__(self, count):
        self.count, self.left, self.phase = count, count, 0

    def arrive(self):
        self.left -= 1
        assert self.left >= 0
        if self.left == 0:
            self.left, self.phase = self.count, self.phase + 1

    def done(self, parity):           # mbarrier.try_wait.parity semantics
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, tiles, nkb, L, rng):
        self.tiles, self.nkb, self.L, self.rng = tiles, nkb, L, rng
        self.G = tiles * nkb
        self.full_b = [Bar(1) for _ in range(LB)]
        self.empty_b = [Bar(1) for _ in range(LB)]
        self.ready_op = [Bar(N_GATHER) for _ in range(OPN)]
        self.empty_op = [Bar(1) for _ in range(OPN)]
        self.tmem_full = [Bar(1) for _ in range(2)]
        self.tmem_empty = [Bar(N_EPI) for _ in range(2)]
        self.fm_ready = [Bar(N_GATHER) for _ in range(FM_BUF)]
        self.tail_w, self.tail_in, self.tail_out = Bar(1), Bar(N_EPI), Bar(1)
        self.pipe = []                                  # issued tcgen05 work, executed in order, asynchronously
        # modelled storage
        self.b_stage = [None] * LB                      # k-block index g held by weight stage s
        self.op_slot = [dict() for _ in range(OPN)]     # gather warp -> g written
        self.fm_tile = [dict() for _ in range(FM_BUF)]
        self.tail_a = dict()                            # epilogue warp -> (tile, round) written
        self.acc = [dict(content=None, kb=0, readers=N_EPI) for _ in range(2)]   # readers = warps that have read `content`
        self.tail_w_loaded = False
        self.log = []

    # ---- tensor pipe: executes one queued item
    def pipe_step(self):
        kind, *a = self.pipe.pop(0)
        if kind == 'commit':
            a[0].arrive()
        elif kind == 'main':
            t, kb, g, s, o = a
            assert self.b_stage[s] == g, ('weight stage', s, self.b_stage[s], g)
            assert all(self.op_slot[o].get(w) == g for w in range(N_GATHER)), ('operand slot', o, self.op_slot[o], g)
            acc = self.acc[t & 1]
            if kb == 0:
                assert acc['readers'] == N_EPI, ('main MMA overwrites an accumulator still being read', t, acc)
                acc.update(content=('main', t), kb=1, readers=0)
            else:
                assert acc['content'] == ('main', t) and acc['kb'] == kb, ('accumulate into wrong contents', t, kb, acc)
                acc['kb'] = kb + 1
            if kb == self.nkb - 1:
                acc['content'] = ('layer', t, 0)        # complete layer-1 pre-activation of tile t
        elif kind == 'tail':
            t, l = a
            assert self.tail_w_loaded, 'tail MMA before its weights landed'
            assert all(self.tail_a.get(w) == (t, l) for w in range(N_EPI)), ('tail operand', self.tail_a, (t, l))
            acc = self.acc[t & 1]
            assert acc['content'] == ('layer', t, l) and acc['readers'] == N_EPI, ('tail MMA overwrites unread accumulator', t, l, acc)
            acc.update(content=('layer', t, l + 1), readers=0)

    # ---- actors (generators): yield a predicate to block on, or None to just give the scheduler a turn
    def producer(self):
        self.tail_w_loaded = True                        # TMA completes asynchronously; modelled as landed + arrive
        self.tail_w.arrive()
        for g in range(self.G):
            s = g % LB
            yield lambda s=s, g=g: self.empty_b[s].done(((g // LB) & 1) ^ 1)
            self.b_stage[s] = g
            yield None
            self.full_b[s].arrive()

    def issuer(self):
        st = dict(t=0, l=0, n=0)

        def tail_ready():
            return self.tail_in.done(st['n'] & 1)

        def tail_issue():
            self.pipe.append(('tail', st['t'], st['l']))
            self.pipe.append(('commit', self.tail_out))
            st['n'] += 1
            st['l'] += 1
            if st['l'] == self.L:
                st['l'] = 0
                st['t'] += 1

        g = 0
        for t in range(self.tiles):
            while st['t'] + 2 <= t:
                yield tail_ready
                if st['n'] == 0:
                    yield lambda: self.tail_w.done(0)
                tail_issue()
            yield lambda t=t: self.tmem_empty[t & 1].done(((t >> 1) & 1) ^ 1)
            for kb in range(self.nkb):
                s, o = g % LB, g % OPN
                if st['t'] < t and tail_ready():
                    if st['n'] == 0:
                        yield lambda: self.tail_w.done(0)
                    tail_issue()
                yield lambda s=s, g=g: self.full_b[s].done((g // LB) & 1)
                yield lambda o=o, g=g: self.ready_op[o].done((g // OPN) & 1)
                self.pipe.append(('main', t, kb, g, s, o))
                self.pipe.append(('commit', self.empty_op[o]))
                self.pipe.append(('commit', self.empty_b[s]))
                g += 1
                yield None
            self.pipe.append(('commit', self.tmem_full[t & 1]))
        while st['t'] < self.tiles:
            yield tail_ready
            if st['n'] == 0:
                yield lambda: self.tail_w.done(0)
            tail_issue()

    def gather(self, w):
        g = 0
        for t in range(self.tiles):
            for kb in range(self.nkb):
                o = g % OPN
                yield lambda o=o, g=g: self.empty_op[o].done(((g // OPN) & 1) ^ 1)
                self.op_slot[o][w] = g
                yield None
                self.ready_op[o].arrive()
                g += 1
            self.fm_tile[t % FM_BUF][w] = t
            self.fm_ready[t % FM_BUF].arrive()

    def epilogue(self, w):
        n_out = 0
        for t in range(self.tiles):
            yield lambda t=t: self.tmem_full[t & 1].done((t >> 1) & 1)
            acc = self.acc[t & 1]
            for r in range(self.L + 1):
                if r > 0:
                    yield lambda n=n_out: self.tail_out.done(n & 1)
                    n_out += 1
                assert acc['content'] == ('layer', t, r), ('epilogue reads', acc['content'], 'expected', ('layer', t, r), 'warp', w)
                acc['readers'] += 1
                yield None
                if r < self.L:
                    self.tail_a[w] = (t, r)
                    yield None
                    self.tail_in.arrive()
            self.tmem_empty[t & 1].arrive()
            yield lambda t=t: self.fm_ready[t % FM_BUF].done((t // FM_BUF) & 1)
            assert all(self.fm_tile[t % FM_BUF].get(x) == t for x in range(N_GATHER)), ('fm tile', t, self.fm_tile[t % FM_BUF])
            yield None

    def run(self):
        actors = {'producer': self.producer(), 'issuer': self.issuer()}
        actors.update({f'gather{w}': self.gather(w) for w in range(N_GATHER)})
        actors.update({f'epi{w}': self.epilogue(w) for w in range(N_EPI)})
        blocked = {k: None for k in actors}              # predicate the actor is waiting on
        steps = 0
        while actors:
            steps += 1
            runnable = [k for k in actors if blocked[k] is None or blocked[k]()]
            choices = runnable + (['pipe'] if self.pipe else [])
            if not choices:
                raise RuntimeError(f'DEADLOCK with actors {sorted(actors)} (tiles={self.tiles}, nkb={self.nkb}, L={self.L})')
            k = self.rng.choice(choices)
            if k == 'pipe':
                self.pipe_step()
                continue
            try:
                blocked[k] = next(actors[k])
            except StopIteration:
                del actors[k]
        while self.pipe:
            self.pipe_step()
        return steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', type=int, default=40)
    args = ap.parse_args()
    n = 0
    for tiles in (1, 2, 3, 4, 5):
        for nkb in (1, 2, 14):
            for L in (1, 2, 3):
                for seed in range(args.seeds):
                    Sim(tiles, nkb, L, random.Random(seed * 7919 + tiles * 131 + nkb * 17 + L)).run()
                    n += 1
    print(f'tc_tail protocol model: {n} randomised runs, no deadlock, no hazard')


if __name__ == '__main__':
    main()
