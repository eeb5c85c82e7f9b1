"""Model-level parity on the GPU: (1) the committed golden vectors produced by the real reference classes,
(2) the oracle on seeded Criteo-shaped inputs at sizes it finishes in seconds.  Tolerance: |dlogit| <= 1e-4
(BASELINE.json north_star), grads rtol 1e-4 / atol 1e-5."""
import pytest
import torch

import oracle
from helpers import load_golden, make_enc, make_batch, assert_close_rel

pytestmark = pytest.mark.gpu


def _build(name, meta):
    from rec_pangu_b200.models import ranking
    cls = getattr(ranking, meta['model'])
    return cls(embedding_dim=meta['D'], enc_dict=meta['enc_dict'], **meta['kwargs'])


def _logit(p):
    p = p.double().clamp(1e-12, 1 - 1e-12)
    return torch.log(p) - torch.log1p(-p)


RANKING_GOLDEN = ['deepfm', 'deepfm_d16', 'fm', 'wdl', 'nfm', 'dcn', 'xdeepfm', 'autoint', 'autoint_l2', 'fibinet']


@pytest.mark.parametrize('name', RANKING_GOLDEN)
def test_ranking_model_matches_reference_golden(name):
    g = load_golden(name)
    model = _build(name, g['meta'])
    model.load_state_dict(g['sd'])                 # same state_dict keys/shapes as the reference
    model = model.cuda().eval()
    data = {k: v.cuda() for k, v in g['data'].items()}
    out = model(data)
    out['loss'].backward()
    from rec_pangu_b200 import ops
    ops.check_index_errors()
    assert out['pred'].shape == g['out']['pred'].shape
    assert (_logit(out['pred'].cpu()) - _logit(g['out']['pred'])).abs().max().item() <= 1e-4
    torch.testing.assert_close(out['pred'].cpu(), g['out']['pred'], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out['loss'].cpu(), g['out']['loss'], rtol=1e-5, atol=1e-6)
    grads = dict(model.named_parameters())
    for k, ref in g['grad'].items():
        assert grads[k].grad is not None, k
        torch.testing.assert_close(grads[k].grad.cpu(), ref, rtol=1e-4, atol=1e-5, msg=lambda s: f'{k}: {s}')
    # inference path returns only pred
    out2 = model(data, is_training=False)
    assert set(out2.keys()) == {'pred'}
    assert torch.equal(out2['pred'], out['pred'])


CRITEO_MID = dict(F=26, Nd=13, V=1000, B=4096)


@pytest.mark.parametrize('model_name,kw,okw', [
    ('DeepFM', dict(embedding_dim=16, hidden_units=[64, 64, 64]), dict(hidden_units=(64, 64, 64))),
    ('FM', dict(embedding_dim=16), {}),
    ('WDL', dict(embedding_dim=16, hidden_units=[64, 64, 64]), dict(hidden_units=(64, 64, 64))),
    ('NFM', dict(embedding_dim=16, hidden_units=[64, 64, 64]), dict(hidden_units=(64, 64, 64))),
    ('DCN', dict(embedding_dim=16, crossing_layers=3), dict(crossing_layers=3)),
    ('xDeepFM', dict(embedding_dim=16), {}),
    ('AutoInt', dict(embedding_dim=32, num_heads=3), dict(num_heads=3)),
    ('FiBiNet', dict(embedding_dim=16), {}),
])
def test_ranking_model_matches_oracle_criteo_shape(model_name, kw, okw):
    from rec_pangu_b200.models import ranking
    from rec_pangu_b200 import ops
    c = CRITEO_MID
    B = c['B'] if model_name != 'FiBiNet' else 512
    enc = make_enc(c['F'], c['Nd'], c['V'])
    torch.manual_seed(1029)
    model = getattr(ranking, model_name)(enc_dict=enc, **kw)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.copy_(torch.randn(p.shape) * 0.05)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().eval()
    data_cpu = make_batch(enc, B, seed=1029)
    data = {k: v.cuda() for k, v in data_cpu.items()}
    out = model(data)
    out['loss'].backward()
    ops.check_index_errors()
    sdr = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    ref = oracle.MODEL_FORWARDS[model_name](sdr, enc, data_cpu, **okw)
    ref['loss'].backward()
    dl = (model._last_logit.cpu().double() - ref['logit'].double()).abs().max().item()
    torch.testing.assert_close(out['pred'].cpu(), ref['pred'], rtol=1e-5, atol=1e-6)
    assert dl <= 1e-4, f'max |dlogit| = {dl}'
    torch.testing.assert_close(out['loss'].cpu(), ref['loss'], rtol=1e-5, atol=1e-6)
    for k, p in model.named_parameters():
        r = sdr[k].grad
        assert p.grad is not None, k
        assert_close_rel(p.grad, r, 1e-4, k)
