#!/usr/bin/env python
"""Randomised discrete-event model of the mbarrier / cp.async protocol of cin_wgrad_tc_kernel
(rec_pangu_b200/csrc/cin_tc.cu): two loader warps (Xk tile + lo pass; G / X0 rows) with CW_AHEAD cp.async groups in flight
over a ring of CW_BST stages, eight software-pipelined operand warps (tcgen05.st of operand g in flight while operand g+1 is
formed from the stage), the MMA issuer, the asynchronous tensor pipe and asynchronous copy completion.  It checks, under random
interleavings,

  * liveness: every actor terminates for any number of k-blocks (more, fewer or equal to the ring / look-ahead depths);
  * the phase/parity arithmetic of every barrier;
  * data hazards: an operand warp forms operand (i, mt) from a stage that holds the COMPLETELY landed G / X0 rows of k-block i
    (not being refilled), the tensor pipe executes the MMAs of (i, mt) with the Xk tile of k-block i landed and lo-extended in
    its stage and with all eight warps' columns of operand (i, mt) in its tensor-memory slot, and no stage or slot is rewritten
    before its last reader is done.

It models the protocol, not the arithmetic (racecheck cannot see synchronisation through mbarriers: profiles/r02_sanitize.md).
Run: python tools/sim/cin_wgrad_protocol.py [--seeds N]
"""
import argparse
import random

BST, AHEAD, OPN, MT, N_OP = 6, 4, 4, 4, 8


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
    def __init__(self, nkb, rng, bst=BST, ahead=AHEAD, opn=OPN):
        self.nkb, self.rng, self.BST, self.AHEAD, self.OPN = nkb, rng, bst, ahead, opn
        self.b_full = [Bar(1) for _ in range(bst)]
        self.b_empty = [Bar(1) for _ in range(bst)]
        self.g_full = [Bar(1) for _ in range(bst)]
        self.g_empty = [Bar(N_OP) for _ in range(bst)]
        self.a_ready = [Bar(N_OP) for _ in range(opn)]
        self.a_empty = [Bar(1) for _ in range(opn)]
        self.acc_done = Bar(1)
        self.pipe = []                                   # issued tcgen05 work, executed in order, asynchronously
        self.copies = {'x': [], 'g': []}                 # cp.async groups in flight per loader warp (complete in order)
        self.landed = {'x': 0, 'g': 0}                   # groups completed so far per loader
        self.x_stage = [None] * bst                      # ('landing', j) while its copies are in flight, ('raw', j), ('ready', j) after the lo pass
        self.g_stage = [None] * bst                      # ('landing', j) / ('ready', j)
        self.op_slot = [dict() for _ in range(opn)]      # operand warp -> global operand index written
        self.mmas = 0
        self.wait_slack = 0                              # mutation hook for the tests: 1 = cp.async.wait_group AHEAD (one group too few)
        self.early_release = False                       # mutation hook: operand warps release the G / X0 stage before their last read

    # ---- asynchronous engines
    def pipe_step(self):
        kind, *a = self.pipe.pop(0)
        if kind == 'commit':
            a[0].arrive()
        else:
            i, mt, s, o, ga = a
            assert self.x_stage[s] == ('ready', i), ('MMA reads Xk stage', s, self.x_stage[s], 'expected k-block', i)
            assert all(self.op_slot[o].get(w) == ga for w in range(N_OP)), ('operand slot', o, self.op_slot[o], ga)
            self.mmas += 1

    def copy_step(self, which):
        j, s = self.copies[which].pop(0)
        st = self.x_stage if which == 'x' else self.g_stage
        assert st[s] == ('landing', j), (which, 'stage', s, st[s], j)
        st[s] = ('raw', j) if which == 'x' else ('ready', j)
        self.landed[which] += 1

    # ---- actors (generators): yield a predicate to block on, or None to just give the scheduler a turn
    def loader(self, which):
        full, empty = (self.b_full, self.b_empty) if which == 'x' else (self.g_full, self.g_empty)
        st = self.x_stage if which == 'x' else self.g_stage
        groups = 0                                       # committed groups, empty ones included

        def issue(j):
            s = j % self.BST
            yield lambda s=s, j=j: empty[s].done(((j // self.BST) & 1) ^ 1)
            if st[s] is not None:
                assert st[s] == ('ready', j - self.BST), (which, 'stage refilled before it was consumed', s, st[s], j)
            st[s] = ('landing', j)
            self.copies[which].append((j, s))

        real = []                                        # index in commit order -> number of REAL groups up to and including it
        n_real = 0
        for j in range(self.AHEAD):
            if j < self.nkb:
                yield from issue(j)
                n_real += 1
            real.append(n_real)
            groups += 1
            yield None
        for i in range(self.nkb):
            s = i % self.BST
            # cp.async.wait_group AHEAD-1: all but the newest AHEAD-1 committed groups are complete -> group i has landed
            need = real[groups - self.AHEAD - self.wait_slack] if groups - self.AHEAD - self.wait_slack >= 0 else 0
            yield lambda need=need: self.landed[which] >= need
            if which == 'x':
                assert st[s] == ('raw', i), ('lo pass reads', s, st[s], i)
                yield None
                st[s] = ('ready', i)                     # lo half written, fence.proxy.async
            else:
                assert st[s] == ('ready', i), (which, s, st[s], i)
            full[s].arrive()
            if i + self.AHEAD < self.nkb:
                yield from issue(i + self.AHEAD)
                n_real += 1
            real.append(n_real)
            groups += 1
            yield None

    def issuer(self):
        ga = 0
        for i in range(self.nkb):
            s = i % self.BST
            yield lambda s=s, i=i: self.b_full[s].done((i // self.BST) & 1)
            for mt in range(MT):
                o = ga % self.OPN
                yield lambda o=o, ga=ga: self.a_ready[o].done((ga // self.OPN) & 1)
                self.pipe.append(('mma', i, mt, s, o, ga))
                self.pipe.append(('commit', self.a_empty[o]))
                ga += 1
                yield None
            self.pipe.append(('commit', self.b_empty[s]))
        self.pipe.append(('commit', self.acc_done))

    def operand(self, w):
        def compute(i, mt):
            s = i % self.BST
            assert self.g_stage[s] == ('ready', i), ('operand warp', w, 'reads G/X0 stage', s, self.g_stage[s], 'expected', i, mt)

        def put(ga):
            o = ga % self.OPN
            yield lambda o=o, ga=ga: self.a_empty[o].done(((ga // self.OPN) & 1) ^ 1)
            self.op_slot[o][w] = ga                      # tcgen05.st issued (completion awaited in publish)

        def publish(ga):
            self.a_ready[ga % self.OPN].arrive()

        if self.nkb == 0:
            return
        yield lambda: self.g_full[0].done(0)
        compute(0, 0)
        ga = 0
        for i in range(self.nkb):
            s = i % self.BST
            for mt_next in (1, 2, 3):
                yield from put(ga)
                yield None
                if self.early_release and mt_next == 3:
                    self.g_empty[s].arrive()
                    yield None
                compute(i, mt_next)
                yield None
                publish(ga)
                ga += 1
            if not self.early_release:
                self.g_empty[s].arrive()                 # stage i completely read (compute(i, 3) above)
            yield from put(ga)
            if i + 1 < self.nkb:
                s2 = (i + 1) % self.BST
                yield lambda s2=s2, i=i: self.g_full[s2].done(((i + 1) // self.BST) & 1)
                compute(i + 1, 0)
            yield None
            publish(ga)
            ga += 1
        if w < 4:                                        # the epilogue warps
            yield lambda: self.acc_done.done(0)
            assert self.mmas == self.nkb * MT, ('epilogue before all MMAs', self.mmas)

    def run(self):
        actors = {'loader_x': self.loader('x'), 'loader_g': self.loader('g'), 'issuer': self.issuer()}
        actors.update({f'op{w}': self.operand(w) for w in range(N_OP)})
        blocked = {k: None for k in actors}
        steps = 0
        while actors:
            steps += 1
            runnable = [k for k in actors if blocked[k] is None or blocked[k]()]
            choices = runnable + (['pipe'] if self.pipe else []) + [c for c in ('x', 'g') if self.copies[c]]
            if not choices:
                raise RuntimeError(f'DEADLOCK with actors {sorted(actors)} (nkb={self.nkb})')
            k = self.rng.choice(choices)
            if k == 'pipe':
                self.pipe_step()
            elif k in ('x', 'g'):
                self.copy_step(k)
            else:
                try:
                    blocked[k] = next(actors[k])
                except StopIteration:
                    del actors[k]
        while self.pipe:
            self.pipe_step()
        assert self.mmas == self.nkb * MT
        return steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', type=int, default=40)
    args = ap.parse_args()
    n = 0
    for nkb in (0, 1, 2, 3, 4, 5, 6, 7, 11, 12, 13, 40):
        for seed in range(args.seeds):
            Sim(nkb, random.Random(seed * 7919 + nkb)).run()
            n += 1
    print(f'cin_wgrad protocol model: {n} randomised runs, no deadlock, no hazard')


if __name__ == '__main__':
    main()
