"""CPU model of the three GEMM formulations and the permuted operand layouts of rec_pangu_b200/csrc/cin_tc.cu, checked against
autograd of the oracle's CIN layer (oracle/restatement.py::cin, reference interaction.py:157-169).  The kernels' index maps are
compile-time formulas; this pins the formulas themselves (row permutations of cin_pack_operand_kernel, the (h,u) / (s,h) column
decoding of the epilogues, the P = G x X0 operand of the weight gradient, the 3xTF32 split) without a GPU."""
import numpy as np
import pytest
import torch

U = 16


def pack_operand(W, F, M, mode, Rp):
    """cin_pack_operand_kernel: W [U, F*M] -> [Rp, 32] with permuted rows (zero padded)."""
    out = np.zeros((Rp, 32), dtype=W.dtype)
    rows_valid = U * F if mode == 0 else F * M
    for r in range(rows_valid):
        if mode == 0:
            h, u = divmod(r, U)
            out[r, :M] = W[u, h * M:h * M + M]
        else:
            s, h = divmod(r, F)
            m = (h + s) % M
            out[r, :U] = W[:, h * M + m]
    return out


def layer_reference(x0, xk, W, b):
    """One CIN layer as the reference computes it; returns X_{k+1} [B,U,D]."""
    B, F, D = x0.shape
    had = torch.einsum('bhd,bmd->bhmd', x0, xk).reshape(B, -1, D)
    return torch.einsum('uj,bjd->bud', W, had) + b.view(1, -1, 1)


@pytest.mark.parametrize('F,M,NT,NTILES', [(26, 26, 144, 3), (26, 16, 208, 2), (5, 3, 16, 5)])
def test_forward_formulation_and_column_map(F, M, NT, NTILES):
    rng = np.random.default_rng(F * M)
    B, D = 3, 4
    x0 = rng.standard_normal((B, F, D))
    xk = rng.standard_normal((B, M, D))
    W = rng.standard_normal((U, F * M))
    b = rng.standard_normal(U)
    ref = layer_reference(torch.tensor(x0), torch.tensor(xk), torch.tensor(W), torch.tensor(b)).numpy()
    Wp = pack_operand(W, F, M, 0, NT * NTILES)                       # rows (h,u), K = M columns
    assert NT * NTILES >= U * F
    out = np.zeros((B, U, D))
    for bi in range(B):
        for d in range(D):
            a = np.zeros(32)
            a[:M] = xk[bi, :, d]                                      # TS-mode A operand of the row (b,d), zero padded
            T = Wp @ a                                                # accumulator columns of the row: N = NT*NTILES
            acc = b.copy()
            for j in range(NT * NTILES):                              # CinTCols: j = h*16 + u
                if j < U * F:
                    acc[j % U] += T[j] * x0[bi, j // U, d]
            out[bi, :, d] = acc
    np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize('F,M,NT,NTILES', [(26, 26, 176, 4), (26, 16, 208, 2), (5, 3, 16, 1)])
def test_backward_formulation_and_diagonal_column_map(F, M, NT, NTILES):
    rng = np.random.default_rng(F + M)
    B, D = 2, 3
    x0 = torch.tensor(rng.standard_normal((B, F, D)), requires_grad=True)
    xk = torch.tensor(rng.standard_normal((B, M, D)), requires_grad=True)
    W = torch.tensor(rng.standard_normal((U, F * M)), requires_grad=True)
    b = torch.tensor(rng.standard_normal(U), requires_grad=True)
    G = rng.standard_normal((B, U, D))                                # dL/dX_{k+1}
    (layer_reference(x0, xk, W, b) * torch.tensor(G)).sum().backward()
    Wp = pack_operand(W.detach().numpy(), F, M, 1, NT * NTILES)       # rows (s,h) -> (h, m=(h+s)%M), K = 16 columns
    assert NT * NTILES >= F * M
    dx0 = np.zeros((B, F, D)); dxk = np.zeros((B, M, D))
    seen = set()
    for bi in range(B):
        for d in range(D):
            a = np.zeros(32); a[:U] = G[bi, :, d]
            dz = Wp @ a
            for j in range(NT * NTILES):                              # CinDzCols
                if j < F * M:
                    h = j % F; m = (h + j // F) % M
                    seen.add((h, m))
                    dxk[bi, m, d] += dz[j] * x0[bi, h, d].item()
                    dx0[bi, h, d] += dz[j] * xk[bi, m, d].item()
    assert len(seen) == F * M                                         # the diagonal enumeration is a bijection onto (h, m)
    np.testing.assert_allclose(dx0, x0.grad.numpy(), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(dxk, xk.grad.numpy(), rtol=1e-10, atol=1e-10)
    # neighbouring columns never share an accumulator (what the permutation is for) — a property of the instantiated shapes
    if F == 26:
        for j in range(1, F * M):
            h0, h1 = (j - 1) % F, j % F
            m0, m1 = (h0 + (j - 1) // F) % M, (h1 + j // F) % M
            assert h0 != h1 and m0 != m1
    # weight / bias gradients: dW[(u,h), m] = sum_rows (G[u] * X0[h]) * Xk[m]  (cin_wgrad_tc_kernel), db = sum_rows G
    Gt, X0, Xk = G.transpose(0, 2, 1).reshape(-1, U), x0.detach().numpy().transpose(0, 2, 1).reshape(-1, F), xk.detach().numpy().transpose(0, 2, 1).reshape(-1, M)
    P = (Gt[:, :, None] * X0[:, None, :]).reshape(-1, U * F)          # rows r = (b,d), lane (u,h) = u*F + h
    dW = (P.T @ Xk).reshape(U, F * M)                                 # dst = dW + u*(F*M) + h*M + m
    np.testing.assert_allclose(dW, W.grad.numpy(), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(Gt.sum(0), b.grad.numpy(), rtol=1e-10, atol=1e-10)


def test_three_term_tf32_split_keeps_fp32_products():
    """hi = x & 0xFFFFE000 (what the tensor core reads of a raw fp32), lo = x - hi (exact): hi.hi + lo.hi + hi.lo differs from the
    fp32 product by the dropped lo.lo term only (<= 2^-22 relative), the reason the CIN kernels hold the 1e-4 bound."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal(4096).astype(np.float32)
    y = rng.standard_normal(4096).astype(np.float32)

    def split(v):
        hi = (v.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
        return hi, (v - hi).astype(np.float32)
    xh, xl = split(x); yh, yl = split(y)
    assert np.all((xh.astype(np.float64) + xl.astype(np.float64)) == x.astype(np.float64))           # the split is exact
    approx = xh.astype(np.float64) * yh + xl.astype(np.float64) * yh + xh.astype(np.float64) * yl
    exact = x.astype(np.float64) * y.astype(np.float64)
    assert np.max(np.abs(approx - exact) / np.maximum(np.abs(exact), 1e-30)) < 2.0 ** -20
