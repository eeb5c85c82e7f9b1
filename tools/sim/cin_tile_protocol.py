#!/usr/bin/env python
"""Randomised discrete-event model of the barrier protocol shared by cin_fwd2_tc_kernel and cin_bwd_tc_kernel
(rec_pangu_b200/csrc/cin_tc.cu): per 128-row tile four operand warps write the TS-mode A operand into one of two tensor-memory
buffers (a_ready / a_empty), the issuer runs NTILES N-tiles of MMAs over it into two alternating accumulator buffers
(acc_full / acc_empty, a counter that keeps running across tiles: with NTILES = 3 the buffer an N-tile lands in alternates
from tile to tile), four epilogue warps drain them.  Checks liveness, the parity arithmetic, that an accumulator buffer is never
overwritten before all four epilogue warps have read its previous contents, that the epilogue reads (tile, N-tile) in order,
and that an operand buffer is not rewritten while MMAs that read it are still queued in the tensor pipe.

Run: python tools/sim/cin_tile_protocol.py [--seeds N]"""
import argparse
import random

N_OP = N_EPI = 4


class Bar:
    def __init__(self, count):
        self.count, self.left, self.phase = count, count, 0

    def arrive(self):
        self.left -= 1
        assert self.left >= 0
        if self.left == 0:
            self.left, self.phase = self.count, self.phase + 1

    def done(self, parity):
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, tiles, ntiles, rng):
        self.tiles, self.NT, self.rng = tiles, ntiles, rng
        self.a_ready = [Bar(N_OP), Bar(N_OP)]
        self.a_empty = [Bar(1), Bar(1)]
        self.acc_full = [Bar(1), Bar(1)]
        self.acc_empty = [Bar(N_EPI), Bar(N_EPI)]
        self.pipe = []
        self.a_buf = [dict(), dict()]                    # operand warp -> tile written
        self.acc = [dict(content=None, readers=N_EPI), dict(content=None, readers=N_EPI)]
        self.early_a_release = False                     # mutation hook: a_empty committed before the tile's last N-tile is issued

    def pipe_step(self):
        kind, *a = self.pipe.pop(0)
        if kind == 'commit':
            a[0].arrive()
        else:
            t, nt, ab, buf = a
            assert all(self.a_buf[ab].get(w) == t for w in range(N_OP)), ('operand buffer', ab, self.a_buf[ab], 'tile', t)
            acc = self.acc[buf]
            assert acc['readers'] == N_EPI, ('MMA overwrites an accumulator still being read', t, nt, acc)
            acc.update(content=(t, nt), readers=0)

    def operand(self, w):
        for t in range(self.tiles):
            ab = t & 1
            yield lambda ab=ab, t=t: self.a_empty[ab].done(((t >> 1) & 1) ^ 1)
            self.a_buf[ab][w] = t
            yield None
            self.a_ready[ab].arrive()

    def issuer(self):
        n_acc = 0
        for t in range(self.tiles):
            ab = t & 1
            yield lambda ab=ab, t=t: self.a_ready[ab].done((t >> 1) & 1)
            for nt in range(self.NT):
                buf = n_acc & 1
                yield lambda buf=buf, n=n_acc: self.acc_empty[buf].done(((n >> 1) & 1) ^ 1)
                if self.early_a_release and nt == self.NT - 1:
                    self.pipe.append(('commit', self.a_empty[ab]))
                    yield None
                self.pipe.append(('mma', t, nt, ab, buf))
                self.pipe.append(('commit', self.acc_full[buf]))
                n_acc += 1
                yield None
            if not self.early_a_release:
                self.pipe.append(('commit', self.a_empty[ab]))

    def epilogue(self, w):
        n_acc = 0
        for t in range(self.tiles):
            for nt in range(self.NT):
                buf = n_acc & 1
                yield lambda buf=buf, n=n_acc: self.acc_full[buf].done((n >> 1) & 1)
                acc = self.acc[buf]
                assert acc['content'] == (t, nt), ('epilogue warp', w, 'reads', acc['content'], 'expected', (t, nt))
                acc['readers'] += 1
                yield None
                self.acc_empty[buf].arrive()
                n_acc += 1

    def run(self):
        actors = {'issuer': self.issuer()}
        actors.update({f'op{w}': self.operand(w) for w in range(N_OP)})
        actors.update({f'epi{w}': self.epilogue(w) for w in range(N_EPI)})
        blocked = {k: None for k in actors}
        while actors:
            runnable = [k for k in actors if blocked[k] is None or blocked[k]()]
            choices = runnable + (['pipe'] if self.pipe else [])
            if not choices:
                raise RuntimeError(f'DEADLOCK with actors {sorted(actors)} (tiles={self.tiles}, NTILES={self.NT})')
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', type=int, default=40)
    args = ap.parse_args()
    n = 0
    for tiles in (0, 1, 2, 3, 4, 5, 9):
        for ntiles in (1, 2, 3, 4):
            for seed in range(args.seeds):
                Sim(tiles, ntiles, random.Random(seed * 7919 + tiles * 31 + ntiles)).run()
                n += 1
    print(f'cin tile protocol model: {n} randomised runs, no deadlock, no hazard')


if __name__ == '__main__':
    main()
