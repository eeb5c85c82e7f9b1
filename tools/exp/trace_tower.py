"""Per-role cycle breakdown of the layer-1 GEMM with the tower tail in its epilogue (rpb_debug_tc_trace)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from rec_pangu_b200 import _lib, ops
from rec_pangu_b200.models.layers import MLP

NAMES = ['kernel', 'prod wait raw', 'mma wait operands', 'mma wait acc', 'mma issue', 'split wait tma', 'split wait op stage',
         'split work', 'epi wait acc', 'epi layer-1 part', 'k-blocks', 'epi tail']


def run(M, K, hidden, fused):
    lib = _lib.load()
    ops.FUSED_TOWER_EPILOGUE = fused
    torch.manual_seed(0)
    m = MLP(input_dim=K, output_dim=1, hidden_units=hidden, hidden_activations='relu', dropout_rates=0).cuda()
    Ws, bs, relu, drops = m.layer_params()
    params = []
    for W, b in zip(Ws, bs):
        params += [W.detach(), b.detach()]
    ld = (K + 3) // 4 * 4
    x = torch.zeros(M, ld, device='cuda')
    x[:, :K] = torch.randn(M, K, device='cuda')
    label = (torch.rand(M, device='cuda') < 0.3).float()
    cfg = dict(n_hidden=len(hidden), has_out=True, K=K, relu=relu, dropout=drops, training=False, impl=0)
    for _ in range(3):
        ops._tower_fwd(cfg, x, params, addend=label, head=(label, 0.0, 1.0))
    torch.cuda.synchronize()
    lib.rpb_debug_tc_trace(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops._tower_fwd(cfg, x, params, addend=label, head=(label, 0.0, 1.0))
    e1.record()
    torch.cuda.synchronize()
    out = (C.c_uint64 * 16)()
    lib.rpb_debug_tc_trace(out, 0)
    print(f'M={M} K={K} hidden={hidden} fused={fused} (event {e0.elapsed_time(e1) * 1e3:.1f} us incl. split_pack + launch overhead)')
    for n, v in zip(NAMES, out):
        print(f'   {n:22s} {v:10d}')


if __name__ == '__main__':
    run(65536, 429, [64, 64, 64], 1)
    run(65536, 429, [64, 64, 64], 0)
    run(65536, 429, [64, 64], 1)
