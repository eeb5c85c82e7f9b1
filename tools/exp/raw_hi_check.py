"""Does tcgen05 kind::tf32 truncate fp32 operands (ignore the low 13 mantissa bits)?  Compare the weight gradient with the
masked 'hi' operand against the same kernel fed the raw fp32 words."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from rec_pangu_b200 import _lib

lib = _lib.load()
M, N, K, ldx = 65536, 64, 429, 432
torch.manual_seed(0)
x = torch.zeros(M, ldx, device='cuda'); x[:, :K] = torch.randn(M, K, device='cuda')
dy = torch.randn(M, N, device='cuda')
W = torch.randn(N, K, device='cuda')
st = torch.cuda.current_stream().cuda_stream
outs = []
for raw in (0, 1, 0):
    lib.rpb_set_option(b'tf32_raw_hi', raw)
    dW = torch.zeros(N, K, device='cuda')
    rc = lib.rpb_linear_bwd(dy.data_ptr(), N, x.data_ptr(), ldx, W.data_ptr(), None, 0, None, 0, dW.data_ptr(), None, M, N, K, 2, st)
    assert rc == 0
    torch.cuda.synchronize()
    outs.append(dW)
lib.rpb_set_option(b'tf32_raw_hi', 1)
ref = (dy.double().t() @ x[:, :K].double())
print('masked vs masked (atomics order noise) max abs diff', (outs[0] - outs[2]).abs().max().item())
print('masked vs raw                          max abs diff', (outs[0] - outs[1]).abs().max().item())
print('masked vs fp64 max rel err', ((outs[0].double() - ref).abs().max() / ref.abs().max()).item())
print('raw    vs fp64 max rel err', ((outs[1].double() - ref).abs().max() / ref.abs().max()).item())
