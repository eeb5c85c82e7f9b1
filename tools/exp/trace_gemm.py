"""Per-role stall breakdown of the persistent tcgen05 GEMM on the DeepFM layer shapes (rpb_debug_tc_trace)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from rec_pangu_b200 import _lib, ops

NAMES = ['kernel', 'prod wait raw', 'mma wait operands', 'mma wait acc', 'mma issue', 'split wait tma', 'split wait op stage',
         'split work', 'epi wait acc', 'epi work', 'k-blocks']


def run(M, N, K, ld, a_tmem):
    lib = _lib.load()
    lib.rpb_set_option(b'gemm_a_tmem', a_tmem)
    x = torch.zeros(M, ld, device='cuda')
    x[:, :K] = torch.randn(M, K, device='cuda')
    W = torch.randn(N, K, device='cuda') / K ** 0.5
    b = torch.randn(N, device='cuda')
    for _ in range(3):
        ops.linear(x, W, b, K=K, impl=2)
    torch.cuda.synchronize()
    lib.rpb_debug_tc_trace(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.linear(x, W, b, K=K, impl=2)
    e1.record()
    torch.cuda.synchronize()
    out = (C.c_uint64 * 16)()
    lib.rpb_debug_tc_trace(out, 0)
    print(f'M={M} N={N} K={K} a_tmem={a_tmem}  (event {e0.elapsed_time(e1) * 1e3:.1f} us incl. split_pack + launch)')
    for n, v in zip(NAMES, out):
        print(f'   {n:22s} {v:10d}')
    lib.rpb_set_option(b'gemm_a_tmem', 1)


if __name__ == '__main__':
    for a in (1, 0):
        run(65536, 64, 429, 432, a)
        run(65536, 64, 64, 64, a)


def run_wgrad(M, N, K, ldx):
    lib = _lib.load()
    x = torch.zeros(M, ldx, device='cuda')
    x[:, :K] = torch.randn(M, K, device='cuda')
    dy = torch.randn(M, N, device='cuda')
    W = torch.randn(N, K, device='cuda')
    dW = torch.zeros(N, K, device='cuda')
    st = torch.cuda.current_stream().cuda_stream

    def call():
        return lib.rpb_linear_bwd(dy.data_ptr(), N, x.data_ptr(), ldx, W.data_ptr(), None, 0, None, 0, dW.data_ptr(), None,
                                  M, N, K, 2, st)
    for _ in range(3):
        assert call() == 0
    torch.cuda.synchronize()
    lib.rpb_debug_tc_trace(None, 1)
    assert call() == 0
    torch.cuda.synchronize()
    out = (C.c_uint64 * 16)()
    lib.rpb_debug_tc_trace(out, 0)
    print(f'wgrad M={M} N={N} K={K}')
    for n, v in zip(NAMES, out):
        print(f'   {n:22s} {v:10d}')


if __name__ == '__main__':
    run_wgrad(65536, 64, 429, 432)
    run_wgrad(65536, 64, 64, 64)
