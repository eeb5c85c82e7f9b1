"""GPU parity tests of the fused MLP tower tail (rpb_tower_tail_fwd / rpb_tower_tail_bwd through the C ABI) against
the oracle's MLP (oracle/restatement.py::mlp, models/layers/deep.py:62-84) + torch.nn.BCELoss on the CPU."""
import os

import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _mlp(K, hidden, seed=0):
    from rec_pangu_b200.models.layers import MLP
    torch.manual_seed(seed)
    m = MLP(input_dim=K, output_dim=1, hidden_units=hidden, hidden_activations='relu', dropout_rates=0).cuda()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.copy_(torch.randn_like(p) * 0.1)
    return m


@pytest.mark.parametrize('M,K,hidden', [(700, 429, [64, 64, 64]), (64, 32, [64]), (1, 16, [64, 64]),
                                        (4097, 100, [64, 64, 64, 64, 64]), (333, 429, [64, 64])])
def test_tower_mlp_matches_oracle(M, K, hidden):
    """MLP.forward/backward through the tower kernels (dlogit handed in by autograd) vs the CPU oracle."""
    from rec_pangu_b200 import ops
    assert ops.TOWER_TAIL == 1
    m = _mlp(K, hidden)
    ld = (K + 3) // 4 * 4
    x = torch.zeros(M, ld, device='cuda')
    x[:, :K] = torch.randn(M, K, device='cuda')
    x.requires_grad_(True)
    n0 = ops.launch_count()
    out = m(x, K=K)
    n_fwd = ops.launch_count() - n0
    assert n_fwd <= 3, f'tower forward should be layer-1 GEMM (+weight split) + ONE tail kernel, saw {n_fwd} launches'
    w = torch.randn(M, 1, device='cuda')
    (out * w).sum().backward()
    sd = {'p.' + k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xc = x.detach().cpu()[:, :K].clone().requires_grad_(True)
    ref = oracle.mlp(sd, 'p', xc, len(hidden), 2)
    (ref * w.cpu()).sum().backward()
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(x.grad.cpu()[:, :K], xc.grad, rtol=1e-4, atol=1e-5)
    assert torch.count_nonzero(x.grad[:, K:]) == 0
    for k, p in m.named_parameters():
        r = sd['p.' + k].grad
        tol = 1e-4 * max(1.0, r.abs().max().item())
        assert (p.grad.cpu() - r).abs().max().item() <= tol, k


def test_tower_equals_layerwise_path():
    """RPB_TOWER_TAIL=0 (one GEMM / row-dot kernel per layer, 3xTF32) and the fused fp32 tail agree."""
    from rec_pangu_b200 import ops
    m = _mlp(429, [64, 64, 64], seed=3)
    x = torch.zeros(5000, 432, device='cuda')
    x[:, :429] = torch.randn(5000, 429, device='cuda')
    res = []
    for flag in (1, 0):
        ops.TOWER_TAIL = flag
        try:
            xi = x.clone().requires_grad_(True)
            out = m(xi, K=429)
            out.sum().backward()
            res.append((out.detach().clone(), xi.grad.clone(), [p.grad.clone() for p in m.parameters()]))
            m.zero_grad()
        finally:
            ops.TOWER_TAIL = 1
    torch.testing.assert_close(res[0][0], res[1][0], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(res[0][1], res[1][1], rtol=1e-4, atol=2e-5)
    for a, b in zip(res[0][2], res[1][2]):
        assert (a - b).abs().max().item() <= 1e-4 * max(1.0, b.abs().max().item())


@pytest.mark.parametrize('B', [2048, 777])
def test_deepfm_fused_head_matches_oracle(B):
    """DeepFM with sigmoid + BCE produced by the tower kernel (label passed into ops.deepfm_core): pred, loss, logit and
    every gradient vs the oracle; the loss gradient scale comes in through autograd (loss * 0.5)."""
    from helpers import make_enc, make_batch, assert_close_rel, capture_relu_inputs, kink_adjacent_samples, drop_samples
    from rec_pangu_b200 import ops
    from rec_pangu_b200.models.ranking import DeepFM
    enc = make_enc(26, 13, 500)
    torch.manual_seed(1029)
    model = DeepFM(embedding_dim=16, hidden_units=[64, 64, 64], enc_dict=enc)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if 'embedding_layer' in n:
                p.mul_(0.25)
            elif p.dim() == 1:
                p.copy_(torch.randn(p.shape) * 0.05)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().train()
    data_cpu = make_batch(enc, B, seed=7)
    # samples on a ReLU kink (oracle pre-activation within 5e-6 of zero) are dropped from the batch on both sides, every
    # gradient of the rest is held to the strict bound without an outlier budget (helpers.kink_adjacent_samples)
    with torch.no_grad(), capture_relu_inputs() as cap:
        oracle.deepfm({k: v.clone() for k, v in sd.items()}, enc, data_cpu, hidden_units=(64, 64, 64))
    kink = kink_adjacent_samples(cap.inputs)
    assert int(kink.sum()) <= max(4, B // 100)
    data_cpu = drop_samples(data_cpu, ~kink)
    data = {k: v.cuda() for k, v in data_cpu.items()}
    n0 = ops.launch_count()
    out = model(data)
    assert ops.launch_count() - n0 <= 4          # gather, weight split + layer-1 GEMM, tower tail (incl. sigmoid + BCE)
    (out['loss'] * 0.5).backward()
    ops.check_index_errors()
    sdr = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    ref = oracle.deepfm(sdr, enc, data_cpu, hidden_units=(64, 64, 64))
    (ref['loss'] * 0.5).backward()
    assert out['pred'].shape == ref['pred'].shape
    assert (model._last_logit.cpu().double() - ref['logit'].double()).abs().max().item() <= 1e-4
    torch.testing.assert_close(out['pred'].cpu(), ref['pred'], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out['loss'].cpu(), ref['loss'], rtol=1e-5, atol=1e-6)
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        assert_close_rel(p.grad, sdr[k].grad, 1e-4, k, atol=1e-8)
    # the un-fused path (inference head) gives the same probabilities bit for bit
    with torch.no_grad():
        out2 = model(data, is_training=False)
    assert torch.equal(out2['pred'], out['pred'])


def test_deepfm_pred_consumer_gets_gradient():
    """A caller that differentiates `pred` (not only `loss`) still gets the right gradient from the fused node."""
    from helpers import make_enc, make_batch
    from rec_pangu_b200.models.ranking import DeepFM
    enc = make_enc(4, 2, 50)
    torch.manual_seed(5)
    model = DeepFM(embedding_dim=8, hidden_units=[64, 64], enc_dict=enc)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().train()
    data_cpu = make_batch(enc, 300, seed=3)
    out = model({k: v.cuda() for k, v in data_cpu.items()})
    (out['loss'] + out['pred'].sum() * 0.01).backward()
    sdr = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    ref = oracle.deepfm(sdr, enc, data_cpu, hidden_units=(64, 64))
    (ref['loss'] + ref['pred'].sum() * 0.01).backward()
    for k, p in model.named_parameters():
        r = sdr[k].grad
        assert (p.grad.cpu() - r).abs().max().item() <= 1e-4 * max(1.0, r.abs().max().item()), k


@pytest.mark.parametrize('M,K,hidden', [(5000, 429, [64, 64, 64]), (37893, 429, [64, 64, 64]), (512, 40, [64, 64]),
                                        (20000, 429, [64, 64, 64, 64, 64])])
def test_tower_in_gemm_epilogue_equals_standalone_tail(M, K, hidden):
    """rpb_linear_tower_fwd (tail run by the epilogue warps of the layer-1 tcgen05 GEMM) vs rpb_linear_fwd +
    rpb_tower_tail_fwd: same arithmetic per element, so activations, logit, pred and loss agree bit for bit."""
    from rec_pangu_b200 import ops
    m = _mlp(K, hidden, seed=11)
    Ws, bs, relu, drops = m.layer_params()
    params = []
    for W, b in zip(Ws, bs):
        params += [W.detach(), b.detach()]
    ld = (K + 3) // 4 * 4
    x = torch.zeros(M, ld, device='cuda')
    x[:, :K] = torch.randn(M, K, device='cuda')
    addend = torch.randn(M, device='cuda')
    label = (torch.rand(M, device='cuda') < 0.3).float()
    cfg = dict(n_hidden=len(hidden), has_out=True, K=K, relu=relu, dropout=drops, training=False, impl=0)
    res = []
    for flag in (1, 0):
        ops.FUSED_TOWER_EPILOGUE = flag
        try:
            n0 = ops.launch_count()
            logit, acts, pred, loss = ops._tower_fwd(cfg, x, params, addend=addend, head=(label, 0.0, 1.0))
            res.append((logit, acts, pred, loss, ops.launch_count() - n0))
        finally:
            ops.FUSED_TOWER_EPILOGUE = 1
    torch.cuda.synchronize()
    assert res[0][4] == 2 and res[1][4] == 3          # (weight split + fused GEMM) vs (split + GEMM + tail kernel)
    assert torch.equal(res[0][0], res[1][0])
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.equal(a, b)
    assert torch.equal(res[0][2], res[1][2])
    torch.testing.assert_close(res[0][3], res[1][3], rtol=1e-6, atol=1e-7)     # loss: different partial-sum grouping
    ref = torch.nn.functional.binary_cross_entropy(torch.sigmoid(res[0][0].double().squeeze(1)), label.double())
    torch.testing.assert_close(res[0][3].double(), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('B,F,Nd,hidden', [(4096, 26, 13, [64, 64, 64]), (5000, 26, 13, [64, 64]), (1300, 4, 0, [64, 64, 64]),
                                           (2048, 6, 40, [64, 64, 64, 64])])
def test_deepfm_one_kernel_forward_equals_separate_kernels(B, F, Nd, hidden):
    """rpb_deepfm_fwd_fused (gather + FM + layer 1 + tail + loss in one launch, operand rows fetched with cp.async by the
    GEMM's split warps) vs rpb_gather_fwd + rpb_linear_tower_fwd: the feature row is a pure copy (bit exact), layer 1 sees
    the same operands (bit exact), the FM term is summed in another order (1e-5), gradients agree."""
    from helpers import make_enc, make_batch
    from rec_pangu_b200 import ops
    from rec_pangu_b200.models.ranking import DeepFM
    enc = make_enc(F, Nd, 3000)
    torch.manual_seed(2)
    model = DeepFM(embedding_dim=16, hidden_units=hidden, enc_dict=enc)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if 'embedding_layer' in n:
                p.mul_(0.25)
            elif p.dim() == 1:
                p.copy_(torch.randn(p.shape) * 0.05)
    model = model.cuda().train()
    data = make_batch(enc, B, seed=21, device='cuda')
    res = []
    for flag in (1, 0):
        ops.FUSED_GATHER_GEMM = flag
        try:
            model.zero_grad()
            n0 = ops.launch_count()
            out = model(data)
            nl = ops.launch_count() - n0
            out['loss'].backward()
            ops.check_index_errors()
            grads = {n: p.grad.clone() for n, p in model.named_parameters()}
            res.append((out['pred'].detach().clone(), out['loss'].detach().clone(), model._last_logit.clone(), grads, nl))
        finally:
            ops.FUSED_GATHER_GEMM = 1
    assert res[0][4] == 3 and res[1][4] == 3            # (2 weight splits + fused kernel) vs (gather, split + GEMM-with-tail)
    torch.testing.assert_close(res[0][2], res[1][2], rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(res[0][0], res[1][0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(res[0][1], res[1][1], rtol=1e-5, atol=1e-6)
    for n in res[0][3]:
        a, b = res[0][3][n], res[1][3][n]
        assert (a - b).abs().max().item() <= 2e-4 * max(1e-6, b.abs().max().item()) + 1e-7, n
    # inference (no autograd): no feature row is materialised, same probabilities
    with torch.no_grad():
        p_fused = model(data, is_training=False)['pred']
    torch.testing.assert_close(p_fused, res[0][0], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('B,F,Nd,hidden', [(4096, 26, 13, [64, 64, 64]), (5000, 26, 13, [64, 64]), (1300, 4, 0, [64, 64, 64, 64])])
def test_deepfm_one_kernel_forward_tc_tail(B, F, Nd, hidden):
    """rpb_set_option('fused_tc_tail', 1) (the default): the 64x64 tail layers of rpb_deepfm_fwd_fused on tcgen05 (3xTF32,
    operands handed from layer to layer through tensor memory) vs the exact-fp32 CUDA-core tail of the same kernel: logits
    within 1e-4 (north_star tolerance), gradients to the tensor-relative tolerance of the other tcgen05 tests."""
    from helpers import make_enc, make_batch
    from rec_pangu_b200 import ops, _lib
    from rec_pangu_b200.models.ranking import DeepFM
    enc = make_enc(F, Nd, 3000)
    torch.manual_seed(2)
    model = DeepFM(embedding_dim=16, hidden_units=hidden, enc_dict=enc)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if 'embedding_layer' in n:
                p.mul_(0.25)
            elif p.dim() == 1:
                p.copy_(torch.randn(p.shape) * 0.05)
    model = model.cuda().train()
    data = make_batch(enc, B, seed=21, device='cuda')
    res = []
    for flag in (1, 0):
        _lib.check(_lib.load().rpb_set_option(b'fused_tc_tail', flag), 'rpb_set_option(fused_tc_tail)')
        try:
            model.zero_grad()
            out = model(data)
            out['loss'].backward()
            torch.cuda.synchronize()
            ops.check_index_errors()
            grads = {n: p.grad.clone() for n, p in model.named_parameters()}
            res.append((out['pred'].detach().clone(), out['loss'].detach().clone(), model._last_logit.clone(), grads))
        finally:
            _lib.check(_lib.load().rpb_set_option(b'fused_tc_tail', 1), 'rpb_set_option(fused_tc_tail)')
    assert (res[0][2] - res[1][2]).abs().max().item() <= 1e-4
    torch.testing.assert_close(res[0][1], res[1][1], rtol=1e-5, atol=1e-6)
    for n in res[0][3]:
        a, b = res[0][3][n], res[1][3][n]
        assert (a - b).abs().max().item() <= 5e-4 * max(1e-6, b.abs().max().item()) + 1e-7, n


@pytest.mark.parametrize('B,F,Nd,hidden', [(4096, 26, 13, [64, 64, 64]), (5000, 26, 13, [64, 64]), (1300, 4, 0, [64, 64, 64, 64])])
def test_tower_tail_backward_tc(B, F, Nd, hidden):
    """rpb_set_option('tower_bwd_tc', 1): the dz chain of rpb_tower_tail_bwd on tcgen05 (3xTF32 through tensor memory) vs the
    exact-fp32 CUDA-core kernel: identical forward, every gradient within the tensor-relative tolerance."""
    from helpers import make_enc, make_batch
    from rec_pangu_b200 import ops, _lib
    from rec_pangu_b200.models.ranking import DeepFM
    enc = make_enc(F, Nd, 3000)
    torch.manual_seed(2)
    model = DeepFM(embedding_dim=16, hidden_units=hidden, enc_dict=enc)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if 'embedding_layer' in n:
                p.mul_(0.25)
            elif p.dim() == 1:
                p.copy_(torch.randn(p.shape) * 0.05)
    model = model.cuda().train()
    data = make_batch(enc, B, seed=21, device='cuda')
    res = []
    for flag in (1, 0):
        _lib.check(_lib.load().rpb_set_option(b'tower_bwd_tc', flag), 'rpb_set_option(tower_bwd_tc)')
        try:
            model.zero_grad()
            out = model(data)
            out['loss'].backward()
            torch.cuda.synchronize()
            ops.check_index_errors()
            res.append((out['loss'].detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters()}))
        finally:
            _lib.check(_lib.load().rpb_set_option(b'tower_bwd_tc', 0), 'rpb_set_option(tower_bwd_tc)')
    assert torch.equal(res[0][0], res[1][0])
    for n in res[0][1]:
        a, b = res[0][1][n], res[1][1][n]
        assert (a - b).abs().max().item() <= 5e-4 * max(1e-6, b.abs().max().item()) + 1e-7, n


@pytest.mark.parametrize('G', [2, 3, 8])
def test_sharded_fused_core_on_local_shards(G):
    """ops.SHARDED_FUSED with dist.LocalShards (all G shards of every table on this GPU): the sharded variants of the
    one-kernel forward and of the dx-GEMM scatter epilogue resolve owner = id mod G / local row = id div G through the same
    pointer tables they get over NVLink.  Logits equal the unsharded model's, every table gradient (shards re-interleaved)
    and every dense gradient agree."""
    from helpers import make_enc, make_batch
    from rec_pangu_b200 import ops, dist as rdist
    from rec_pangu_b200.models.ranking import DeepFM
    enc = make_enc(6, 3, [1001, 577, 333, 2000, 170, 64])
    torch.manual_seed(4)
    ref = DeepFM(embedding_dim=16, hidden_units=[64, 64], enc_dict=enc)
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if 'embedding_layer' in n:
                p.mul_(0.25)
    sh = DeepFM(embedding_dim=16, hidden_units=[64, 64], enc_dict=enc)
    sh.load_state_dict({k: v.clone() for k, v in ref.state_dict().items()})
    ref, sh = ref.cuda().train(), sh.cuda().train()
    ls = rdist.LocalShards(sh.embedding_layer, G)
    sh.embedding_layer.attach_shards(ls)
    data = make_batch(enc, 1500, seed=5, device='cuda')
    out_r = ref(data)
    out_r['loss'].backward()
    old = ops.SHARDED_FUSED
    ops.SHARDED_FUSED = 1
    try:
        n0 = ops.launch_count()
        out_s = sh(data)
        assert ops.launch_count() - n0 == 3                      # two weight splits (layer 1, tower tail) + the one-kernel forward
        out_s['loss'].backward()
    finally:
        ops.SHARDED_FUSED = old
    torch.cuda.synchronize()
    ops.check_index_errors()
    torch.testing.assert_close(sh._last_logit, ref._last_logit, rtol=0, atol=1e-6)
    for f, t in enumerate(ref.embedding_layer.tables()):
        full = ls.full_grad(f)
        assert (full - t.grad).abs().max().item() <= 1e-4 * max(1e-6, t.grad.abs().max().item()) + 1e-8, f
    dense_r = {n: p.grad for n, p in ref.named_parameters() if not n.startswith('embedding_layer.')}
    for n, p in sh.named_parameters():
        if not n.startswith('embedding_layer.'):
            torch.testing.assert_close(p.grad, dense_r[n], rtol=1e-4, atol=1e-6, msg=lambda s: f'{n}: {s}')


@pytest.mark.parametrize('vec', [0, 1])
@pytest.mark.parametrize('name', ['autoint', 'autoint_l2'])
def test_autoint_vec_matches_reference_golden(name, vec):
    """Both lane I/O variants of the AutoInt attention kernels (rpb_set_option('autoint_vec', 0 | 1); 1 = float4, the default)
    against the fixtures produced by the real reference AutoInt (tests/golden)."""
    from helpers import load_golden
    from rec_pangu_b200 import ops, _lib
    from rec_pangu_b200.models import ranking
    g = load_golden(name)
    m = g['meta']
    model = getattr(ranking, m['model'])(embedding_dim=m['D'], enc_dict=m['enc_dict'], **m['kwargs'])
    model.load_state_dict(g['sd'])
    model = model.cuda().eval()
    data = {k: v.cuda() for k, v in g['data'].items()}
    _lib.check(_lib.load().rpb_set_option(b'autoint_vec', vec), 'rpb_set_option(autoint_vec)')
    try:
        out = model(data)
        out['loss'].backward()
        torch.cuda.synchronize()
        ops.check_index_errors()
    finally:
        _lib.check(_lib.load().rpb_set_option(b'autoint_vec', 1), 'rpb_set_option(autoint_vec)')
    torch.testing.assert_close(out['pred'].cpu(), g['out']['pred'], rtol=1e-5, atol=1e-6)
    grads = dict(model.named_parameters())
    for k, ref in g['grad'].items():
        torch.testing.assert_close(grads[k].grad.cpu(), ref, rtol=1e-4, atol=1e-5, msg=lambda s: f'{k}: {s}')


def test_deepfm_one_kernel_reports_bad_index():
    from helpers import make_enc, make_batch
    from rec_pangu_b200 import ops
    from rec_pangu_b200.models.ranking import DeepFM
    enc = make_enc(4, 2, 100)
    model = DeepFM(embedding_dim=16, hidden_units=[64, 64], enc_dict=enc).cuda()
    data = make_batch(enc, 1024, seed=1, device='cuda')
    data['C3'][517] = 101 + 5                          # one past the OOV row and beyond
    with torch.no_grad():
        model(data, is_training=False)
    with pytest.raises(IndexError):
        ops.check_index_errors()
