"""Gather-kernel experiments on the GPU box: row-load cache policy x L2 fetch granularity."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rec_pangu_b200 import ops, _lib


def timeit(fn, iters=30, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(8e7))
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def main():
    B, F, Nd, D, V = 65536, 26, 13, 16, 1_000_000
    dev = 'cuda'
    lib = _lib.load()
    torch.manual_seed(0)
    tables = [torch.randn(V + 1, D, device=dev) for _ in range(F)]
    NB = 4
    idxs = [[torch.randint(0, V + 1, (B,), device=dev) for _ in range(F)] for _ in range(NB)]
    dense = [[torch.rand(B, device=dev) for _ in range(Nd)] for _ in range(NB)]
    cnt = [0]

    def g():
        i = cnt[0] % NB
        cnt[0] += 1
        with torch.no_grad():
            ops.gather(tables, idxs[i], dense[i], want_fm=True)
    res = {}
    for gran in (None,):
        if gran is not None:
            rc = lib.rpb_set_option(b'l2_fetch_granularity', gran)
            res[f'set_gran_{gran}_rc'] = rc
        for pol in (1, 3):
            lib.rpb_set_option(b'gather_load_policy', pol)
            t = timeit(g)
            res[f'gather_fm gran={gran} policy={pol}'] = round(t, 2)
            print(f'gran={gran} policy={pol}: {t:.1f} us  alg {B * 1928 / t / 1e3:.0f} GB/s', flush=True)
    json.dump(res, open('gpurun_out/exp_gather.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
