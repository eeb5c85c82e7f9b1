"""tcgen05 / TMEM / TMA dense layer (3xTF32) against float64 and against the fp32 SIMT kernel.
Named zz so that it runs after every other GPU test file."""
import pytest
import torch

from test_kernels_gpu import _check_linear

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('M,N,K,relu', [(128, 64, 32, False), (128, 64, 64, False), (1000, 64, 429, True),
                                        (4096, 64, 429, True), (257, 64, 64, True), (300, 36, 70, False),
                                        (512, 392, 100, False), (2048, 256, 512, False), (640, 428, 64, False)])
def test_linear_tcgen05_forward_backward(M, N, K, relu):
    _check_linear(M, N, K, relu, impl=2, tol=1e-5)


def test_tcgen05_matches_simt_on_deepfm_layer1_shape():
    from rec_pangu_b200 import ops
    torch.manual_seed(0)
    M, N, K = 65536, 64, 429
    x = torch.zeros(M, 432, device='cuda')
    x[:, :K] = torch.randn(M, K, device='cuda') * 0.35
    W = torch.randn(N, K, device='cuda') * (2.0 / K) ** 0.5
    b = torch.randn(N, device='cuda') * 0.1
    y1 = ops.linear(x, W, b, K=K, impl=1)
    y2 = ops.linear(x, W, b, K=K, impl=2)
    assert (y1 - y2).abs().max().item() < 2e-5


@pytest.mark.parametrize('M,N,K,relu', [(1000, 64, 429, True), (640, 428, 64, False)])
def test_linear_tcgen05_v1_kernel_still_correct(M, N, K, relu):
    """The one-tile-per-CTA kernel (gemm_v2 = 0) stays available as a cross-check of the persistent kernel."""
    from rec_pangu_b200 import _lib
    lib = _lib.load()
    assert lib.rpb_set_option(b'gemm_v2', 0) == 0
    try:
        _check_linear(M, N, K, relu, impl=2, tol=1e-5)
    finally:
        lib.rpb_set_option(b'gemm_v2', 1)


def test_persistent_gemm_many_tiles_and_resident_weights():
    """More tiles than SMs (persistent loop, TMEM double buffering) with resident (K=64) and streamed (K=429) weights."""
    from rec_pangu_b200 import ops
    torch.manual_seed(1)
    for M, N, K in ((148 * 128 * 3 + 77, 64, 64), (148 * 128 * 2 + 5, 64, 429), (40000, 224, 64)):
        ld = (K + 3) // 4 * 4
        x = torch.zeros(M, ld, device='cuda')
        x[:, :K] = torch.randn(M, K, device='cuda')
        W = torch.randn(N, K, device='cuda') / K ** 0.5
        b = torch.randn(N, device='cuda')
        y1 = ops.linear(x, W, b, K=K, impl=1)
        y2 = ops.linear(x, W, b, K=K, impl=2)
        assert (y1 - y2).abs().max().item() < 3e-5, (M, N, K)


@pytest.mark.parametrize('option', [b'gemm_a_tmem', b'gemm_stack_n'])
@pytest.mark.parametrize('M,N,K,relu', [(1000, 64, 429, True), (4096, 128, 64, False), (148 * 128 * 2 + 5, 64, 429, True),
                                        (300, 36, 70, False)])
def test_linear_tcgen05_operand_modes_still_correct(M, N, K, relu, option):
    """gemm_a_tmem = 0 keeps the split A operand in shared memory (SS-mode MMA) instead of tensor memory (TS mode);
    gemm_stack_n = 0 issues three N = block_n MMAs per k-step instead of two against the stacked [B hi ; B lo] operand."""
    from rec_pangu_b200 import _lib
    lib = _lib.load()
    assert lib.rpb_set_option(option, 0) == 0
    try:
        _check_linear(M, N, K, relu, impl=2, tol=1e-5)
    finally:
        lib.rpb_set_option(option, 1)
