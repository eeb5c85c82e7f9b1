"""Generate tests/golden/*.npz by running the UNMODIFIED reference classes from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference needs `faiss` and `dgl` only for workloads outside the hot path; they are stubbed in
sys.modules exactly as SURVEY.md App. B-2 describes.  Everything is seeded; the .npz files are committed.
Each file holds:  sd/<state_dict key>, data/<column>, out/<key>, grad/<state_dict key>, plus meta/json.
"""
import importlib.machinery
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
# make sure `rec_pangu` resolves to the reference, not to this repo's alias package
sys.path = [p for p in sys.path if os.path.abspath(p or '.') != REPO]
for name in ['faiss', 'dgl', 'dgl.function']:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    sys.modules[name] = m
sys.modules['dgl'].function = sys.modules['dgl.function']
sys.modules['dgl'].DGLGraph = object
os.environ.setdefault('WANDB_MODE', 'disabled')
sys.path.insert(0, '/root/reference')

import numpy as np  # noqa: E402
import torch  # noqa: E402
from rec_pangu.models.ranking import DeepFM, xDeepFM, AutoInt, DCN, FiBiNet, FM, WDL, NFM, AFM, MaskNet, AFN, AOANet, CCPM  # noqa: E402
from rec_pangu.models.multi_task import MMOE, ShareBottom, OMOE, MLMMOE, ESSM, AITM  # noqa: E402
from rec_pangu.models.layers import LR_Layer  # noqa: E402
from rec_pangu.models.layers import (FM_Layer, MLP, CrossNet, CompressedInteractionNet, SENET_Layer,  # noqa: E402
                                     BilinearInteractionLayer, MultiHeadSelfAttention, InnerProductLayer)

assert 'reference' in sys.modules['rec_pangu'].__file__, sys.modules['rec_pangu'].__file__


def make_enc(n_sparse, n_dense, vocabs):
    enc = {f'I{i + 1}': {'min': 0.0, 'max': 1.0} for i in range(n_dense)}
    enc.update({f'C{i + 1}': {'vocab_size': int(vocabs[i])} for i in range(n_sparse)})
    return enc


def make_batch(enc, B, gen, labels=('label',)):
    data = {}
    for c, d in enc.items():
        if 'vocab_size' in d:
            data[c] = torch.randint(0, d['vocab_size'] + 1, (B,), dtype=torch.int64, generator=gen)
            data[c][0] = d['vocab_size']          # always exercise the OOV row
            data[c][1] = data[c][2]               # and a duplicate index (scatter-add collision)
        else:
            data[c] = torch.rand(B, generator=gen)
    for l in labels:
        data[l] = (torch.rand(B, generator=gen) < 0.3).float()
    return data


def save(name, model, enc, data, out, extra_sd=None, meta=None):
    arrs = {}
    sd = dict(model.state_dict())
    if extra_sd:
        sd.update(extra_sd)
    for k, v in sd.items():
        arrs['sd/' + k] = v.detach().numpy()
    for k, v in data.items():
        arrs['data/' + k] = v.numpy()
    for k, v in out.items():
        arrs['out/' + k] = v.detach().numpy()
    for k, p in model.named_parameters():
        if p.grad is not None:
            arrs['grad/' + k] = p.grad.detach().numpy()
    if extra_sd:
        for k, p in extra_sd.items():
            if getattr(p, 'grad', None) is not None:
                arrs['grad/' + k] = p.grad.detach().numpy()
    m = {'enc_dict': enc}
    m.update(meta or {})
    arrs['meta/json'] = np.frombuffer(json.dumps(m).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **arrs)
    print(name, 'saved', sum(a.nbytes for a in arrs.values()) // 1024, 'KiB raw')


def run_model(name, ctor, kwargs, n_sparse=5, n_dense=3, B=48, D=8, seed=1029):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    vocabs = [17, 50, 97, 9, 33, 21, 64][:n_sparse]
    enc = make_enc(n_sparse, n_dense, vocabs)
    model = ctor(embedding_dim=D, enc_dict=enc, **kwargs)
    # biases start at torch defaults / zeros in places; randomise them so the golden exercises them
    with torch.no_grad():
        for k, p in model.named_parameters():
            if p.dim() == 1:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
    model.eval()                                   # Dropout off (reference MLP default p=0.1 for xDeepFM/AutoInt)
    data = make_batch(enc, B, gen)
    out = model(data)
    out['loss'].backward()
    save(name, model, enc, data, out, meta={'model': ctor.__name__, 'kwargs': kwargs, 'D': D})


def run_mmoe(name, bn_training, seed=1029):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    enc = make_enc(4, 2, [17, 50, 97, 9])
    model = MMOE(embedding_dim=8, enc_dict=enc, mmoe_hidden_dim=16, hidden_dim=[16, 8], dropouts=[0.0, 0.0],
                 device='cpu')
    with torch.no_grad():
        for k, p in model.named_parameters():
            if p.dim() == 1:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1 + (1.0 if 'batchnorm' in k and 'weight' in k else 0.0))
        for k, b in model.named_buffers():
            if 'running_mean' in k:
                b.copy_(torch.randn(b.shape, generator=gen) * 0.1)
            if 'running_var' in k:
                b.copy_(torch.rand(b.shape, generator=gen) + 0.5)
    model.train(bn_training)
    sd_before = {k: v.clone() for k, v in model.state_dict().items()}
    data = make_batch(enc, 48, gen, labels=('task1_label', 'task2_label'))
    out = model(data)
    out['loss'].backward()
    extra = {f'gates.{i}': g for i, g in enumerate(model.gates)}
    extra.update({f'gates_bias.{i}': g for i, g in enumerate(model.gates_bias)})
    # state_dict must be the one BEFORE the forward (train-mode BN updates running stats in place)
    model.load_state_dict(sd_before)
    save(name, model, enc, data, out, extra_sd=extra,
         meta={'model': 'MMOE', 'bn_training': bn_training,
               'kwargs': {'mmoe_hidden_dim': 16, 'hidden_dim': [16, 8], 'dropouts': [0.0, 0.0]}, 'D': 8})


def run_multitask(name, ctor, kwargs, bn_training, list_attrs=(), seed=1029):
    """ShareBottom / OMOE / MLMMOE (multi_task/{sharebottom,omoe,mlmmoe}.py).  `list_attrs`: unregistered python lists of
    Parameters (MLMMOE's level_gates / gates / gates_bias) saved next to the state_dict with their gradients."""
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    enc = make_enc(4, 2, [17, 50, 97, 9])
    model = ctor(embedding_dim=8, enc_dict=enc, **kwargs)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if p.dim() == 1:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1 + (1.0 if 'batchnorm' in k and 'weight' in k else 0.0))
            elif k in ('experts', 'experts_bias'):
                p.mul_(0.3)                       # normal_(0,1) experts over a 34-wide input saturate the sigmoid
        for k, b in model.named_buffers():
            if 'running_mean' in k:
                b.copy_(torch.randn(b.shape, generator=gen) * 0.1)
            if 'running_var' in k:
                b.copy_(torch.rand(b.shape, generator=gen) + 0.5)
    model.train(bn_training)
    sd_before = {k: v.clone() for k, v in model.state_dict().items()}
    data = make_batch(enc, 48, gen, labels=('task1_label', 'task2_label'))
    out = model(data)
    out['loss'].backward()
    extra = {}
    for a in list_attrs:
        extra.update({f'{a}.{i}': g for i, g in enumerate(getattr(model, a))})
    model.load_state_dict(sd_before)
    save(name, model, enc, data, out, extra_sd=extra,
         meta={'model': ctor.__name__, 'bn_training': bn_training, 'kwargs': kwargs, 'D': 8, 'list_attrs': list(list_attrs)})


def run_layers(seed=7):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    B, Fn, D = 32, 6, 8
    e = torch.randn(B, Fn, D, generator=gen)
    arrs = {'in/e': e.numpy()}

    def put(tag, mod, y):
        for k, v in mod.state_dict().items():
            arrs[f'{tag}/sd/{k}'] = v.numpy()
        arrs[f'{tag}/out'] = y.detach().numpy()

    put('fm', FM_Layer(), FM_Layer()(e))
    put('bi', InnerProductLayer(output='Bi_interaction_pooling'), InnerProductLayer(output='Bi_interaction_pooling')(e))
    x = torch.randn(B, Fn * D + 3, generator=gen)
    arrs['in/x'] = x.numpy()
    m = MLP(input_dim=Fn * D + 3, output_dim=1, hidden_units=[16, 8], hidden_activations='relu', dropout_rates=0)
    put('mlp', m, m(x))
    c = CrossNet(Fn * D + 3, 3)
    with torch.no_grad():
        for l in c.cross_net:
            l.bias.copy_(torch.randn(l.bias.shape, generator=gen) * 0.1)
    put('crossnet', c, c(x))
    ci = CompressedInteractionNet(Fn, [4, 5, 3])
    put('cin', ci, ci(e))
    s = SENET_Layer(Fn, 3)
    put('senet', s, s(e))
    bl = BilinearInteractionLayer(Fn, D, 'field_interaction')
    put('bilinear', bl, bl(e))
    a = MultiHeadSelfAttention(D, attention_dim=4, num_heads=3, align_to='output')
    put('mhsa', a, a(e))
    e2 = torch.randn(B, Fn, 12, generator=gen)
    arrs['in/e2'] = e2.numpy()
    a2 = MultiHeadSelfAttention(12, attention_dim=4, num_heads=3, align_to='output')   # D == H*d -> no W_res
    put('mhsa_nores', a2, a2(e2))
    np.savez_compressed(os.path.join(HERE, 'layers.npz'), **arrs)
    print('layers saved')


def run_essm(seed=1029):
    """ESSM (multi_task/essm.py), eval mode (its MLPs carry Dropout(0.2) modules)."""
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    enc = make_enc(4, 2, [17, 50, 97, 9])
    kwargs = {'hidden_dim': [16, 8], 'dropouts': [0.2, 0.2]}
    model = ESSM(embedding_dim=8, enc_dict=enc, device='cpu', **kwargs)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if p.dim() == 1:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
    model.eval()
    data = make_batch(enc, 48, gen, labels=('task1_label', 'task2_label'))
    out = model(data)
    out['loss'].backward()
    save('essm', model, enc, data, out, meta={'model': 'ESSM', 'kwargs': kwargs, 'D': 8})


def run_multitask_all():
    run_essm()
    tw = {'dropouts': [0.0, 0.0]}
    for bn in (False, True):
        sfx = '_train' if bn else '_eval'
        run_multitask('sharebottom' + sfx, ShareBottom, dict(hidden_units=[16, 8], **tw), bn)
        run_multitask('omoe' + sfx, OMOE, dict(omoe_hidden_dim=16, hidden_dim=[16, 8], **tw), bn)
        run_multitask('mlmmoe' + sfx, MLMMOE, dict(mmoe_hidden_dim=16, hidden_dim=[16, 8], device='cpu', **tw), bn,
                      list_attrs=('level_gates', 'gates', 'gates_bias'))


def run_lr(seed=3029):
    """ranking/lr.py cannot be constructed in the reference (lr.py:28 calls a reset_parameters() it does not have,
    SURVEY.md App. A-15): the fixture drives what its forward would run — the reference's LR_Layer, sigmoid, BCELoss on
    squeeze(-1) (lr.py:42-55) — with the state_dict keys the class would have (`lr_layer.*`)."""
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    enc = make_enc(5, 3, [17, 50, 97, 9, 33])

    class _LR(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lr_layer = LR_Layer(enc_dict=enc)

        def forward(self, data):
            y = self.lr_layer(data).sigmoid()
            return {'pred': y, 'loss': torch.nn.BCELoss()(y.squeeze(-1), data['label'])}
    model = _LR()
    data = make_batch(enc, 48, gen)
    out = model(data)
    out['loss'].backward()
    save('lr', model, enc, data, out, meta={'model': 'LR', 'kwargs': {}, 'D': 1})


def run_aitm(seed=4029):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed)
    enc = make_enc(4, 2, [17, 50, 97, 9])
    kwargs = {'tower_dims': [16, 16, 8], 'drop_prob': [0.1, 0.1, 0.1]}
    model = AITM(embedding_dim=8, enc_dict=enc, **kwargs)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if p.dim() == 1:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
    model.eval()
    data = make_batch(enc, 48, gen, labels=('task1_label', 'task2_label'))
    out = model(data)
    out['loss'].backward()
    save('aitm', model, enc, data, out, meta={'model': 'AITM', 'kwargs': kwargs, 'D': 8})


def run_round2():
    """Fixtures added in round 2: MaskNet (parallel and serial blocks), LR, AITM."""
    run_model('masknet', MaskNet, {'hidden_units': [16, 8], 'block_num': 3, 'use_parallel': True}, n_sparse=6, seed=5029)
    run_model('masknet_serial', MaskNet, {'hidden_units': [16, 8], 'block_num': 2, 'use_parallel': False}, n_sparse=6, seed=6029)
    run_lr()
    run_aitm()
    run_model('afn', AFN, {'dnn_hidden_units': [16, 8], 'afn_hidden_units': [16, 8], 'logarithmic_neurons': 3}, n_sparse=6, seed=7029)
    run_model('aoanet', AOANet, {'dnn_hidden_units': [16, 8], 'num_interaction_layers': 2, 'num_subspaces': 3}, n_sparse=6, seed=8029)
    run_model('ccpm', CCPM, {'channels': [3, 2], 'kernel_heights': [4, 3]}, n_sparse=6, seed=9029)


if __name__ == '__main__':
    if '--only-round2' in sys.argv:
        run_round2()
        sys.exit(0)
    if '--only-multitask' in sys.argv:            # ShareBottom / OMOE / MLMMOE fixtures only (added after the first set)
        run_multitask_all()
        sys.exit(0)
    if '--only-afm' in sys.argv:                  # AFM fixture only (added after the first set)
        run_model('afm', AFM, {'hidden_units': [16, 8]}, n_sparse=6, seed=2029)
        sys.exit(0)
    run_layers()
    run_model('deepfm', DeepFM, {'hidden_units': [16, 8]})
    run_model('deepfm_d16', DeepFM, {'hidden_units': [32, 16, 8]}, n_sparse=7, n_dense=4, D=16, B=40)
    run_model('xdeepfm', xDeepFM, {'dnn_hidden_units': [16, 8], 'cin_layer_units': [4, 5, 3]})
    run_model('autoint', AutoInt, {'dnn_hidden_units': [16, 8], 'num_heads': 3, 'attention_dim': 4})
    run_model('autoint_l2', AutoInt, {'dnn_hidden_units': [16], 'num_heads': 2, 'attention_dim': 4, 'attention_layers': 2})
    run_model('dcn', DCN, {'crossing_layers': 3})
    run_model('fibinet', FiBiNet, {'hidden_units': [16, 8]}, n_sparse=6)
    run_model('afm', AFM, {'hidden_units': [16, 8]}, n_sparse=6, seed=2029)
    run_model('fm', FM, {})
    run_model('wdl', WDL, {'hidden_units': [16, 8]})
    run_model('nfm', NFM, {'hidden_units': [16, 8]})
    run_mmoe('mmoe_eval', bn_training=False)
    run_mmoe('mmoe_train', bn_training=True)
    run_multitask_all()
    run_round2()
