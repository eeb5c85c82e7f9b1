"""Per-role cycle breakdown of the one-kernel DeepFM forward (rpb_debug_fused_trace), config 2 shape."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from rec_pangu_b200 import _lib, ops
from rec_pangu_b200.models.ranking import DeepFM

NAMES = ['kernel', 'split: cp.async wait', 'split: wait TMEM slot', 'split: work', 'mma: wait weights', 'mma: wait operands',
         'mma: wait accumulator', 'mma: issue', 'epi: wait accumulator', 'epi: layer-1 part + FM wait', 'epi: tail',
         'weight producer: wait stage', 'gather: fence+sync+TMA store', 'gather: LDS+FM+split', 'gather: tcgen05.st+wait',
         'gather: issue() next requests']


def main():
    lib = _lib.load()
    for kv in filter(None, os.environ.get('RPB_OPTIONS', '').split(',')):
        k, v = kv.split('=')
        _lib.check(lib.rpb_set_option(k.encode(), int(v)), f'rpb_set_option({k})')
    enc = bench.make_enc(bench.WORKLOADS['deepfm'])
    torch.manual_seed(0)
    with torch.device('cuda'):
        model = DeepFM(embedding_dim=16, hidden_units=[64, 64, 64], enc_dict=enc)
    model.set_grad_mode('persistent')
    model.train()
    gen = torch.Generator(device='cuda').manual_seed(1)
    data = bench.synth_batch(enc, 65536, gen, device='cuda')
    for need_grad in (True, False):
        ctx = torch.enable_grad() if need_grad else torch.no_grad()
        with ctx:
            for _ in range(3):
                model(data)
            torch.cuda.synchronize()
            lib.rpb_debug_fused_trace(None, 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model(data)
            e1.record()
            torch.cuda.synchronize()
        out = (C.c_uint64 * 16)()
        lib.rpb_debug_fused_trace(out, 0)
        print(f'materialise x = {need_grad}  (event {e0.elapsed_time(e1) * 1e3:.1f} us incl. split_pack + launch overhead)')
        for n, v in zip(NAMES, out):
            print(f'   {n:32s} {v:10d}')
        ct = (C.c_uint64 * 1024)()
        lib.rpb_debug_fused_cta_times(ct)
        print(f'   (fetch/split kernel: [split: cp.async wait] = fetch warp waits for a free stage, [gather: issue()] = fetch warp requests, '
              f'split warp waits for rows {ct[805]})')
        print(f'   CTA 0 timeline (cycles from start): gather prologue done {ct[800]}, gather loop end {ct[802]} (tile-end sections {ct[801]}), '
              f'last layer-1 MMA issued {ct[804]}, epilogue done {ct[803]}')
        recs = [(ct[4 * i], ct[4 * i + 1], ct[4 * i + 2], ct[4 * i + 3]) for i in range(148)]
        t0 = min(r[1] for r in recs)
        print('   per-CTA (globaltimer): first start 0, last start %.1f us, first end %.1f us, last end %.1f us' % (
            (max(r[1] for r in recs) - t0) / 1e3, (min(r[2] for r in recs) - t0) / 1e3, (max(r[2] for r in recs) - t0) / 1e3))
        for nt in sorted(set(r[3] for r in recs)):
            d = sorted((r[2] - r[1]) / 1e3 for r in recs if r[3] == nt)
            print(f'   CTAs with {nt} tiles: n={len(d)} duration us min {d[0]:.1f} median {d[len(d) // 2]:.1f} max {d[-1]:.1f}')
        by_sm = sorted(recs, key=lambda r: r[2] - r[1])
        print('   slowest CTAs (sm, tiles, us):', [(int(r[0]), int(r[3]), round((r[2] - r[1]) / 1e3, 1)) for r in by_sm[-8:]])
        print('   fastest CTAs (sm, tiles, us):', [(int(r[0]), int(r[3]), round((r[2] - r[1]) / 1e3, 1)) for r in by_sm[:8]])


if __name__ == '__main__':
    main()
