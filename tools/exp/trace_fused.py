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
         'weight producer: wait stage']


def main():
    lib = _lib.load()
    enc = bench.make_enc()
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


if __name__ == '__main__':
    main()
