"""Diagnostic: does the sparse re-zero run NEXT TO the one-CTA-per-SM forward kernel?  Times (CUDA events, 20 repeats each)
forward alone, the re-zero alone (full grid and as `blocks` x 128 threads), and both overlapped on two streams.

    python tools/exp/overlap_zero.py [blocks ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

import bench
from rec_pangu_b200 import _lib


def main():
    dev = torch.device('cuda', 0)
    w = bench.WORKLOADS['deepfm']
    model, enc, _ = bench.build_model(w, dev, 1)
    gen = torch.Generator(device=dev).manual_seed(1)
    data = bench.synth_batch(enc, w['B'], gen, device=dev, labels=bench.label_names(w))
    lib = _lib.load()
    store = model.embedding_layer._grad_store

    def setopt(v):
        _lib.check(lib.rpb_set_option(b'rows_zero_blocks', v), 'set_option')

    def fwd_bwd():
        out = model(data)
        out['loss'].backward()
        return out

    for _ in range(3):
        fwd_bwd()
        model.zero_grad()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, n=20):
        ts = []
        for _ in range(n):
            fwd_bwd()                                # leaves pending rows to clean
            pend = list(store.pending)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
            store.pending = pend
            setopt(0)
            model.zero_grad()
            torch.cuda.synchronize()
        ts.sort()
        return ts[len(ts) // 2]

    def only_fwd():
        with torch.no_grad():
            model(data)

    def zero_with(blocks):
        def f():
            setopt(blocks)
            store.clean()
            setopt(0)
        return f

    def overlapped(blocks):
        def f():
            main = torch.cuda.current_stream()
            fork, join = torch.cuda.Event(), torch.cuda.Event()
            fork.record(main)
            with torch.no_grad():
                model(data)
            side.wait_event(fork)
            with torch.cuda.stream(side):
                setopt(blocks)
                store.clean()
                setopt(0)
                join.record(side)
            main.wait_event(join)
        return f

    print(f'forward alone (eager, incl. 2 weight-split launches): {timed(only_fwd):.1f} us')
    print(f're-zero alone, full grid: {timed(zero_with(0)):.1f} us')
    for blocks in [int(a) for a in sys.argv[1:]] or [148, 296, 592]:
        print(f're-zero alone, {blocks} x 128 threads: {timed(zero_with(blocks)):.1f} us;  forward || re-zero: {timed(overlapped(blocks)):.1f} us')


if __name__ == '__main__':
    main()
