"""GPU parity tests of the individual kernels (through the C ABI via rec_pangu_b200.ops) against the oracle."""
import itertools

import pytest
import torch

import oracle
from helpers import make_enc, make_batch, assert_close_rel

pytestmark = pytest.mark.gpu


def _tables(enc, D, seed=0, device='cuda'):
    g = torch.Generator().manual_seed(seed)
    return {c: torch.randn(enc[c]['vocab_size'] + 1, D, generator=g).to(device) for c in oracle.sparse_cols(enc)}


def _sd(tables, prefix='embedding_layer'):
    return {f'{prefix}.embedding_layer.{c}.weight': t for c, t in tables.items()}


@pytest.mark.parametrize('F,Nd,D,B', [(26, 13, 16, 1000), (5, 3, 8, 77), (3, 0, 32, 64), (7, 2, 40, 129),
                                      (4, 1, 4, 33), (6, 2, 10, 50), (2, 5, 1, 40), (64, 64, 16, 31)])
def test_gather_forward_bit_exact(F, Nd, D, B):
    from rec_pangu_b200 import ops
    enc = make_enc(F, Nd, [50 + 7 * i for i in range(F)])
    data = make_batch(enc, B, device='cuda')
    tabs = _tables(enc, D)
    cols = oracle.sparse_cols(enc)
    x, fm, _ = ops.gather([tabs[c] for c in cols], [data[c] for c in cols],
                          [data[c] for c in oracle.dense_cols(enc)], want_fm=True)
    ops.check_index_errors()
    cpu = {k: v.cpu() for k, v in data.items()}
    e_ref = oracle.embedding_layer({k: v.cpu() for k, v in _sd(tabs).items()}, 'embedding_layer', enc, cpu)
    assert x.shape[1] % 4 == 0
    assert torch.equal(x[:, :F * D].view(B, F, D).cpu(), e_ref)             # pure copy => bit exact
    if Nd:
        assert torch.equal(x[:, F * D:F * D + Nd].cpu(), oracle.get_linear_input(enc, cpu))
    assert torch.count_nonzero(x[:, F * D + Nd:]) == 0
    torch.testing.assert_close(fm.cpu(), oracle.fm_layer(e_ref).squeeze(1), rtol=1e-5, atol=1e-4)


def test_gather_float_and_int32_indices_and_lr():
    from rec_pangu_b200 import ops
    F, Nd, D, B = 5, 3, 8, 64
    enc = make_enc(F, Nd, 30)
    data = make_batch(enc, B, device='cuda')
    tabs, lrt = _tables(enc, D), _tables(enc, 1, seed=5)
    cols, dcols = oracle.sparse_cols(enc), oracle.dense_cols(enc)
    x0, _, lr0 = ops.gather([tabs[c] for c in cols], [data[c] for c in cols], [data[c] for c in dcols],
                            lr_tables=[lrt[c] for c in cols])
    x1, _, _ = ops.gather([tabs[c] for c in cols], [data[c].float() for c in cols], [data[c] for c in dcols])
    x2, _, _ = ops.gather([tabs[c] for c in cols], [data[c].int() for c in cols], [data[c] for c in dcols])
    assert torch.equal(x0, x1) and torch.equal(x0, x2)
    ref = torch.stack([lrt[c][data[c], 0] for c in cols], dim=1)
    assert torch.equal(lr0[:, :F], ref)
    assert torch.equal(lr0[:, F:F + Nd], torch.stack([data[c] for c in dcols], dim=1))


def test_gather_out_of_range_index_raises_indexerror():
    from rec_pangu_b200 import ops
    enc = make_enc(3, 0, 10)
    data = make_batch(enc, 16, device='cuda')
    tabs = _tables(enc, 8)
    cols = oracle.sparse_cols(enc)
    data['C2'][5] = 11                       # vocab_size + 1: one past the OOV row
    ops.gather([tabs[c] for c in cols], [data[c] for c in cols])
    with pytest.raises(IndexError):
        ops.check_index_errors()
    data['C2'][5] = -1
    ops.gather([tabs[c] for c in cols], [data[c] for c in cols])
    with pytest.raises(IndexError):
        ops.check_index_errors()
    data['C2'][5] = 10                       # the OOV row itself is legal
    ops.gather([tabs[c] for c in cols], [data[c] for c in cols])
    ops.check_index_errors()


def test_gather_rejects_cpu_tensors_loudly():
    from rec_pangu_b200 import ops
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ops.gather([torch.zeros(4, 8)], [torch.zeros(3, dtype=torch.int64)])


@pytest.mark.parametrize('F,Nd,D,B,unique', [(26, 13, 16, 512, False), (5, 3, 8, 40, True), (7, 2, 40, 65, False),
                                             (6, 2, 10, 50, False)])
def test_gather_backward_matches_autograd_of_reference_ops(F, Nd, D, B, unique):
    from rec_pangu_b200 import ops
    V = 4096 if unique else 37
    enc = make_enc(F, Nd, V)
    data = make_batch(enc, B, device='cuda')
    if unique:
        for c in oracle.sparse_cols(enc):
            data[c] = torch.randperm(V + 1, device='cuda')[:B]
    tabs = {c: t.requires_grad_(True) for c, t in _tables(enc, D).items()}
    lrt = {c: t.requires_grad_(True) for c, t in _tables(enc, 1, seed=3).items()}
    cols, dcols = oracle.sparse_cols(enc), oracle.dense_cols(enc)
    x, fm, lr_in = ops.gather([tabs[c] for c in cols], [data[c] for c in cols], [data[c] for c in dcols],
                              lr_tables=[lrt[c] for c in cols], want_fm=True)
    g = torch.Generator().manual_seed(1)
    wx = torch.randn(x.shape, generator=g).cuda()
    wf = torch.randn(fm.shape, generator=g).cuda()
    wl = torch.randn(lr_in.shape, generator=g).cuda()
    ((x * wx).sum() + (fm * wf).sum() + (lr_in * wl).sum()).backward()
    # oracle on CPU in float64 semantics of the same ops
    cpu = {k: v.cpu() for k, v in data.items()}
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in _sd(tabs).items()}
    sdl = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in _sd(lrt, 'lr.emb_layer').items()}
    e = oracle.embedding_layer(sd, 'embedding_layer', enc, cpu)
    l = oracle.embedding_layer(sdl, 'lr.emb_layer', enc, cpu).squeeze(-1)
    loss = (e.flatten(1) * wx.cpu()[:, :F * D]).sum() + (oracle.fm_layer(e).squeeze(1) * wf.cpu()).sum() \
        + (l * wl.cpu()[:, :F]).sum()
    loss.backward()
    for c in cols:
        ref = sd[f'embedding_layer.embedding_layer.{c}.weight'].grad
        if unique:
            torch.testing.assert_close(tabs[c].grad.cpu(), ref, rtol=1e-5, atol=1e-5)
        else:
            torch.testing.assert_close(tabs[c].grad.cpu(), ref, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(lrt[c].grad.cpu(), sdl[f'lr.emb_layer.embedding_layer.{c}.weight'].grad,
                                   rtol=1e-5, atol=1e-5)


def test_fm_standalone_forward_backward():
    from rec_pangu_b200 import ops
    e = torch.randn(300, 26, 16, device='cuda', requires_grad=True)
    for mode, ref_fn in (('sum', oracle.fm_layer), ('bi', oracle.bi_interaction)):
        e.grad = None
        out = ops.fm_interaction(e, mode)
        w = torch.randn_like(out)
        (out * w).sum().backward()
        ec = e.detach().cpu().double().requires_grad_(True)
        ref = ref_fn(ec)
        (ref * w.cpu().double()).sum().backward()
        torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-5, atol=1e-4)
        torch.testing.assert_close(e.grad.cpu().double(), ec.grad, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize('M,N,K,relu', [(1000, 64, 429, True), (257, 64, 64, True), (129, 1, 64, False),
                                        (300, 33, 70, False), (512, 390, 100, False)])
def test_linear_simt_forward_backward(M, N, K, relu):
    _check_linear(M, N, K, relu, impl=1, tol=2e-5)


def _check_linear(M, N, K, relu, impl, tol):
    from rec_pangu_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    ld = (K + 3) // 4 * 4
    xbuf = torch.zeros(M, ld)
    xbuf[:, :K] = torch.randn(M, K, generator=g)
    W = (torch.randn(N, K, generator=g) / K ** 0.5)
    b = torch.randn(N, generator=g) * 0.1
    x = xbuf.cuda().requires_grad_(True)
    Wc, bc = W.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    y = ops.linear(x, Wc, bc, K=K, impl=impl)
    gy = torch.randn(M, N, generator=g)
    y.backward(gy.cuda())
    xd = xbuf[:, :K].double().requires_grad_(True)
    Wd, bd = W.double().requires_grad_(True), b.double().requires_grad_(True)
    yd = torch.nn.functional.linear(xd, Wd, bd)
    yd.backward(gy.double())
    scale = yd.abs().max().item()
    assert (y.cpu().double() - yd).abs().max().item() <= tol * max(1.0, scale)
    assert (x.grad.cpu()[:, :K].double() - xd.grad).abs().max().item() <= tol * max(1.0, xd.grad.abs().max().item())
    assert torch.count_nonzero(x.grad[:, K:]) == 0
    assert (Wc.grad.cpu().double() - Wd.grad).abs().max().item() <= 5 * tol * max(1.0, Wd.grad.abs().max().item())
    assert (bc.grad.cpu().double() - bd.grad).abs().max().item() <= 5 * tol * max(1.0, bd.grad.abs().max().item())


@pytest.mark.parametrize('drop', [0.0, 0.3])
def test_mlp_matches_oracle(drop):
    from rec_pangu_b200.models.layers import MLP
    torch.manual_seed(0)
    K = 429
    m = MLP(input_dim=K, output_dim=1, hidden_units=[64, 64, 64], hidden_activations='relu', dropout_rates=drop).cuda()
    m.eval()
    x = torch.zeros(700, 432, device='cuda')
    x[:, :K] = torch.randn(700, K, device='cuda')
    x.requires_grad_(True)
    out = m(x, K=K)
    out.sum().backward()
    stride = 3 if drop > 0 else 2
    sd = {'p.' + k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xc = x.detach().cpu()[:, :K].clone().requires_grad_(True)
    ref = oracle.mlp(sd, 'p', xc, 3, stride)
    ref.sum().backward()
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(x.grad.cpu()[:, :K], xc.grad, rtol=1e-4, atol=1e-5)
    for k, p in m.named_parameters():
        torch.testing.assert_close(p.grad.cpu(), sd['p.' + k].grad, rtol=1e-4, atol=1e-4, msg=lambda s: f'{k}: {s}')


def test_mlp_train_mode_dropout_is_consistent():
    """Train-mode dropout: statistical keep rate, and backward uses the same mask as forward."""
    from rec_pangu_b200.models.layers import MLP
    torch.manual_seed(0)
    m = MLP(input_dim=32, output_dim=1, hidden_units=[256], dropout_rates=0.25).cuda()
    m.train()
    with torch.no_grad():
        m.net[0].weight.copy_(torch.eye(256, 32) + 0.0)
        m.net[0].bias.fill_(1.0)                     # all pre-activations positive
        m.net[3].weight.fill_(1.0)
        m.net[3].bias.zero_()
    x = torch.rand(2000, 32, device='cuda').requires_grad_(True)
    out = m(x)
    out.sum().backward()
    # out = sum_j keep_j * (h_j) / (1-p); with h = x@W^T + 1
    h = (x.detach() @ m.net[0].weight.t() + 1.0)
    ratio = (out.detach().squeeze(1) / h.sum(1)).mean().item()
    assert abs(ratio - 1.0) < 0.02
    gW = m.net[3].weight.grad.squeeze(0)             # = sum_b dropped_h[b, j]
    frac_kept = (gW / (h.sum(0) / 0.75)).mean().item()
    assert abs(frac_kept - 0.75) < 0.02


def test_sigmoid_bce_matches_aten():
    from rec_pangu_b200 import ops
    torch.manual_seed(0)
    z = (torch.randn(5000, 1, device='cuda') * 6).requires_grad_(True)
    z.data[0] = 200.0
    z.data[1] = -200.0                               # saturation: log clamp at -100, backward clamp 1e-12
    y = (torch.rand(5000, device='cuda') < 0.3).float()
    pred, loss = ops.sigmoid_bce(z, y)
    (loss * 3.0).backward()
    zc = z.detach().cpu().clone().requires_grad_(True)
    p = torch.sigmoid(zc)
    l = torch.nn.BCELoss()(p.squeeze(-1), y.cpu())
    (l * 3.0).backward()
    torch.testing.assert_close(pred.cpu(), p, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(loss.cpu(), l, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(z.grad.cpu(), zc.grad, rtol=1e-4, atol=1e-8)
    # eps variant used by MMOE (multi_task/mmoe.py:127-128)
    z2 = torch.randn(1000, 1, device='cuda', requires_grad=True)
    y2 = (torch.rand(1000, device='cuda') < 0.5).float()
    p2, l2 = ops.sigmoid_bce(z2, y2, eps=1e-6, scale=0.5)
    l2.backward()
    zc2 = z2.detach().cpu().clone().requires_grad_(True)
    lr = 0.5 * torch.nn.functional.binary_cross_entropy(torch.sigmoid(zc2).squeeze(-1) + 1e-6, y2.cpu())
    lr.backward()
    torch.testing.assert_close(l2.cpu(), lr, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(z2.grad.cpu(), zc2.grad, rtol=1e-4, atol=1e-8)


def _param_like(t):
    return t.detach().clone().cuda().requires_grad_(True)


@pytest.mark.parametrize('B,K,L', [(300, 429, 3), (65, 70, 1), (1000, 845, 4), (33, 128, 8)])
def test_crossnet_forward_backward(B, K, L):
    from rec_pangu_b200.models.layers import CrossNet
    torch.manual_seed(K)
    m = CrossNet(K, L)
    with torch.no_grad():
        for l in m.cross_net:
            l.weight.weight.copy_(torch.randn(1, K) * (1.0 / K) ** 0.5)
            l.bias.copy_(torch.randn(K) * 0.1)
    sd = {'p.' + k: v.detach().clone().double().requires_grad_(True) for k, v in m.state_dict().items()}
    m = m.cuda()
    ld = (K + 3) // 4 * 4
    xb = torch.zeros(B, ld)
    xb[:, :K] = torch.randn(B, K)
    x = xb.cuda().requires_grad_(True)
    out = m(x, K=K)
    w = torch.randn(B, K)
    (out[:, :K] * w.cuda()).sum().backward()
    xd = xb[:, :K].double().requires_grad_(True)
    ref = oracle.crossnet(sd, 'p', xd, L)
    (ref * w.double()).sum().backward()
    torch.testing.assert_close(out[:, :K].cpu().double(), ref, rtol=1e-5, atol=1e-5)
    assert torch.count_nonzero(out[:, K:]) == 0
    torch.testing.assert_close(x.grad[:, :K].cpu().double(), xd.grad, rtol=1e-4, atol=1e-5)
    for k, p in m.named_parameters():
        assert_close_rel(p.grad, sd['p.' + k].grad, 1e-4, k)


@pytest.mark.parametrize('B,F,D,units', [(200, 26, 16, [16, 16, 16]), (37, 6, 8, [4, 5, 3]), (64, 10, 32, [8, 20]),
                                         (50, 32, 16, [16])])
def test_cin_forward_backward(B, F, D, units):
    from rec_pangu_b200.models.layers import CompressedInteractionNet
    torch.manual_seed(F * D)
    m = CompressedInteractionNet(F, units)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape) * (0.3 if p.dim() > 1 else 0.1))
    sd = {'p.' + k: v.detach().clone().double().requires_grad_(True) for k, v in m.state_dict().items()}
    m = m.cuda()
    e0 = torch.randn(B, F, D) * 0.5
    e = e0.cuda().requires_grad_(True)
    out = m(e)
    w = torch.randn(B, 1)
    (out * w.cuda()).sum().backward()
    ed = e0.double().requires_grad_(True)
    ref = oracle.cin(sd, 'p', ed, units)
    (ref * w.double()).sum().backward()
    scale = max(1.0, ref.abs().max().item())
    assert (out.cpu().double() - ref).abs().max().item() <= 2e-5 * scale
    gs = max(1.0, ed.grad.abs().max().item())
    assert (e.grad.cpu().double() - ed.grad).abs().max().item() <= 5e-5 * gs
    for k, p in m.named_parameters():
        r = sd['p.' + k].grad
        assert (p.grad.cpu().double() - r).abs().max().item() <= 1e-4 * max(1.0, r.abs().max().item()), k


@pytest.mark.parametrize('B,save_x', [(1003, 1), (1003, 0), (8, 1), (1, 0)])
def test_cin_tensor_core_ragged_batches_and_both_backward_modes(B, save_x, monkeypatch):
    """cin_tc.cu at the Criteo shape (F = 26, D = 16, units 16-16-16) on batches that are not multiples of the 8-sample tile or
    the 2-sample k-block, with the forward keeping X_1..X_{L-1} (rpb_cin_fwd_save / rpb_cin_bwd_saved) and with the backward
    recomputing them (rpb_cin_fwd / rpb_cin_bwd)."""
    from rec_pangu_b200 import ops
    from rec_pangu_b200.models.layers import CompressedInteractionNet
    monkeypatch.setattr(ops, 'CIN_SAVE_X', save_x)
    F, D, units = 26, 16, [16, 16, 16]
    torch.manual_seed(B + save_x)
    m = CompressedInteractionNet(F, units)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape) * (0.3 if p.dim() > 1 else 0.1))
    sd = {'p.' + k: v.detach().clone().double().requires_grad_(True) for k, v in m.state_dict().items()}
    m = m.cuda()
    e0 = torch.randn(B, F, D) * 0.5
    e = e0.cuda().requires_grad_(True)
    out = m(e)
    w = torch.randn(B, 1)
    (out * w.cuda()).sum().backward()
    ed = e0.double().requires_grad_(True)
    ref = oracle.cin(sd, 'p', ed, units)
    (ref * w.double()).sum().backward()
    scale = max(1.0, ref.abs().max().item())
    assert (out.cpu().double() - ref).abs().max().item() <= 2e-5 * scale
    gs = max(1.0, ed.grad.abs().max().item())
    assert (e.grad.cpu().double() - ed.grad).abs().max().item() <= 5e-5 * gs
    for k, p in m.named_parameters():
        r = sd['p.' + k].grad
        assert (p.grad.cpu().double() - r).abs().max().item() <= 1e-4 * max(1.0, r.abs().max().item()), k


@pytest.mark.parametrize('B,F,D,H,d', [(100, 26, 32, 3, 8), (33, 6, 8, 3, 4), (64, 6, 12, 3, 4), (50, 32, 16, 1, 16)])
def test_autoint_attention_forward_backward(B, F, D, H, d):
    from rec_pangu_b200.models.layers import MultiHeadSelfAttention
    torch.manual_seed(B)
    m = MultiHeadSelfAttention(D, attention_dim=d, num_heads=H, align_to='output')
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape) * (1.0 / D) ** 0.5)
    sd = {'p.' + k: v.detach().clone().double().requires_grad_(True) for k, v in m.state_dict().items()}
    m = m.cuda()
    x0 = torch.randn(B, F, D)
    x = x0.cuda().requires_grad_(True)
    out = m(x)
    w = torch.randn(B, F, H * d)
    (out * w.cuda()).sum().backward()
    xd = x0.double().requires_grad_(True)
    ref = oracle.mhsa(sd, 'p', xd, H, d)
    (ref * w.double()).sum().backward()
    assert out.shape == ref.shape
    torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(x.grad.cpu().double(), xd.grad, rtol=1e-4, atol=5e-5)
    for k, p in m.named_parameters():
        torch.testing.assert_close(p.grad.cpu().double(), sd['p.' + k].grad, rtol=2e-4, atol=2e-4, msg=lambda s: f'{k}: {s}')


@pytest.mark.parametrize('B,F,D', [(100, 26, 16), (33, 6, 8), (40, 7, 32)])
def test_bilinear_layer_forward_backward(B, F, D):
    from rec_pangu_b200.models.layers import BilinearInteractionLayer
    torch.manual_seed(F)
    m = BilinearInteractionLayer(F, D, 'field_interaction')
    sd = {'p.' + k: v.detach().clone().double().requires_grad_(True) for k, v in m.state_dict().items()}
    m = m.cuda()
    e0 = torch.randn(B, F, D)
    e = e0.cuda().requires_grad_(True)
    out = m(e)
    w = torch.randn(out.shape)
    (out * w.cuda()).sum().backward()
    ed = e0.double().requires_grad_(True)
    ref = oracle.bilinear_field_interaction(sd, 'p', ed)
    (ref * w.double()).sum().backward()
    torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-4, atol=1e-5)
    assert_close_rel(e.grad, ed.grad, 1e-4, 'dE')
    for k, p in m.named_parameters():
        assert_close_rel(p.grad, sd['p.' + k].grad, 1e-4, k)


def test_layers_match_reference_golden_on_gpu():
    """The committed reference-layer outputs (tests/golden/layers.npz) through the CUDA layers."""
    from helpers import load_layers, sub_sd
    from rec_pangu_b200.models import layers as L
    G = load_layers()
    e, x, e2 = G['in/e'].cuda(), G['in/x'].cuda(), G['in/e2'].cuda()
    Fn, D = e.shape[1], e.shape[2]
    tol = dict(rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(L.FM_Layer().cuda()(e).cpu(), G['fm/out'], **tol)
    torch.testing.assert_close(L.InnerProductLayer(output='Bi_interaction_pooling')(e).cpu(), G['bi/out'], **tol)
    m = L.MLP(input_dim=x.shape[1], output_dim=1, hidden_units=[16, 8], hidden_activations='relu', dropout_rates=0)
    m.load_state_dict(sub_sd(G, 'mlp'))
    torch.testing.assert_close(m.cuda()(x).cpu(), G['mlp/out'], **tol)
    c = L.CrossNet(x.shape[1], 3)
    c.load_state_dict(sub_sd(G, 'crossnet'))
    torch.testing.assert_close(c.cuda()(x)[:, :x.shape[1]].cpu(), G['crossnet/out'], **tol)
    ci = L.CompressedInteractionNet(Fn, [4, 5, 3])
    ci.load_state_dict(sub_sd(G, 'cin'))
    torch.testing.assert_close(ci.cuda()(e).cpu(), G['cin/out'], **tol)
    bl = L.BilinearInteractionLayer(Fn, D, 'field_interaction')
    bl.load_state_dict(sub_sd(G, 'bilinear'))
    torch.testing.assert_close(bl.cuda()(e).cpu(), G['bilinear/out'], **tol)
    a = L.MultiHeadSelfAttention(D, attention_dim=4, num_heads=3, align_to='output')
    a.load_state_dict(sub_sd(G, 'mhsa'))
    torch.testing.assert_close(a.cuda()(e).cpu(), G['mhsa/out'], **tol)
    a2 = L.MultiHeadSelfAttention(12, attention_dim=4, num_heads=3, align_to='output')
    a2.load_state_dict(sub_sd(G, 'mhsa_nores'))
    torch.testing.assert_close(a2.cuda()(e2).cpu(), G['mhsa_nores/out'], **tol)


def test_gather_without_materialisation_matches():
    """FM-only consumers (FM inference) skip the [B,F,D] write: fm must equal the materialising launch's fm."""
    from rec_pangu_b200 import ops
    enc = make_enc(26, 13, 500)
    data = make_batch(enc, 3000, device='cuda')
    tabs = _tables(enc, 16)
    cols, dcols = oracle.sparse_cols(enc), oracle.dense_cols(enc)
    with torch.no_grad():
        x, fm, _ = ops.gather([tabs[c] for c in cols], [data[c] for c in cols], [data[c] for c in dcols], want_fm=True)
        x2, fm2, _ = ops.gather([tabs[c] for c in cols], [data[c] for c in cols], [data[c] for c in dcols], want_fm=True,
                                want_x=False)
    assert x2.numel() == 0
    torch.testing.assert_close(fm2, fm, rtol=1e-6, atol=1e-6)


def test_dropout_vector_and_scalar_kernels_draw_the_same_masks():
    """rpb_dropout_fwd / _bwd pick a float4 kernel for 16-byte aligned buffers with n % 4 == 0 and a scalar one otherwise; the
    keep decision of element i depends on (seed, i) only, so both must keep exactly the same elements — checked through the C
    ABI with an explicit seed on an aligned buffer and on the same values shifted by one float (misaligned -> scalar path)."""
    from rec_pangu_b200 import _lib
    lib = _lib.load()
    n, p, seed = 4096 * 4, 0.3, 123456789
    torch.manual_seed(0)
    base = torch.rand(n + 4, device='cuda') + 0.5                   # strictly positive: y != 0 <=> kept
    xa = base[:n].clone()                                           # aligned
    shifted = torch.empty(n + 4, device='cuda')
    shifted[1:n + 1].copy_(xa)
    xs = shifted[1:n + 1]                                           # data pointer + 4 bytes
    assert xa.data_ptr() % 16 == 0 and xs.data_ptr() % 16 == 4
    ya, ys = torch.empty_like(xa), torch.empty(n + 4, device='cuda')[1:n + 1]
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.rpb_dropout_fwd(xa.data_ptr(), ya.data_ptr(), n, p, seed, None, st), 'rpb_dropout_fwd')
    _lib.check(lib.rpb_dropout_fwd(xs.data_ptr(), ys.data_ptr(), n, p, seed, None, st), 'rpb_dropout_fwd')
    torch.cuda.synchronize()
    assert torch.equal(ya, ys)
    keep = (ya != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 0.02
    torch.testing.assert_close(ya[ya != 0], xa[ya != 0] / (1 - p), rtol=1e-6, atol=0)
    # backward with a ReLU output mask: same two paths
    relu = (torch.rand(n, device='cuda') - 0.4).clamp_min(0)
    relu_s = torch.empty(n + 4, device='cuda')[1:n + 1]
    relu_s.copy_(relu)
    da, ds = torch.empty_like(xa), torch.empty(n + 4, device='cuda')[1:n + 1]
    _lib.check(lib.rpb_dropout_bwd(xa.data_ptr(), relu.data_ptr(), da.data_ptr(), n, p, seed, None, st), 'rpb_dropout_bwd')
    _lib.check(lib.rpb_dropout_bwd(xs.data_ptr(), relu_s.data_ptr(), ds.data_ptr(), n, p, seed, None, st), 'rpb_dropout_bwd')
    torch.cuda.synchronize()
    assert torch.equal(da, ds)
    assert torch.equal(da != 0, (ya != 0) & (relu > 0))
