"""CPU tests of the host-side mirror of the reference API: import paths, constructor signatures, state_dict key
contract, dataset encoding, trainer/pipeline plumbing (with a torch stand-in model — the real
models are CUDA-only), C-ABI symbol export, loud failure without CUDA."""
import ctypes
import inspect
import os
import re

import numpy as np
import pandas as pd
import pytest
import torch

from helpers import load_golden, make_enc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_paths_resolve_to_this_build():
    import rec_pangu
    from rec_pangu.models.ranking import WDL, DeepFM, NFM, FiBiNet, AFM, AFN, AOANet, AutoInt, CCPM, LR, FM, xDeepFM, DCN, MaskNet  # noqa
    from rec_pangu.models.multi_task import AITM, ESSM, MLMMOE, MMOE, OMOE, ShareBottom  # noqa
    from rec_pangu.trainer import RankTrainer  # noqa
    from rec_pangu.model_pipeline import train_model, test_model  # noqa
    from rec_pangu.dataset import get_dataloader, get_single_dataloader, BaseDataset, MultiTaskDataset  # noqa
    assert DeepFM.__module__.startswith('rec_pangu_b200')


def test_constructor_signatures_match_reference():
    from rec_pangu_b200.models import ranking
    expect = {
        'DeepFM': ['embedding_dim', 'hidden_units', 'loss_fun', 'enc_dict'],
        'xDeepFM': ['embedding_dim', 'dnn_hidden_units', 'cin_layer_units', 'loss_fun', 'enc_dict'],
        'AutoInt': ['embedding_dim', 'dnn_hidden_units', 'attention_layers', 'num_heads', 'attention_dim', 'loss_fun', 'enc_dict'],
        'DCN': ['embedding_dim', 'hidden_units', 'crossing_layers', 'loss_fun', 'enc_dict'],
        'FiBiNet': ['embedding_dim', 'hidden_units', 'loss_fun', 'enc_dict'],
        'AFM': ['embedding_dim', 'hidden_units', 'loss_fun', 'enc_dict'],
        'FM': ['embedding_dim', 'loss_fun', 'enc_dict'],
        'WDL': ['embedding_dim', 'hidden_units', 'loss_fun', 'enc_dict'],
        'NFM': ['embedding_dim', 'hidden_units', 'loss_fun', 'enc_dict'],
    }
    for name, args in expect.items():
        sig = inspect.signature(getattr(ranking, name).__init__)
        assert list(sig.parameters)[1:] == args, name
        assert sig.parameters['embedding_dim'].default == 32
        assert sig.parameters['loss_fun'].default == 'torch.nn.BCELoss()'


@pytest.mark.parametrize('name', ['deepfm', 'deepfm_d16', 'fm', 'wdl', 'nfm', 'dcn', 'xdeepfm', 'autoint', 'autoint_l2', 'fibinet', 'afm'])
def test_state_dict_contract_matches_reference(name):
    """Same keys and shapes as the reference's state_dict (so reference checkpoints load and vice versa)."""
    from rec_pangu_b200.models import ranking
    g = load_golden(name)
    m = g['meta']
    model = getattr(ranking, m['model'])(embedding_dim=m['D'], enc_dict=m['enc_dict'], **m['kwargs'])
    sd = model.state_dict()
    assert set(sd.keys()) == set(g['sd'].keys())
    for k, v in g['sd'].items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    model.load_state_dict(g['sd'])


def test_models_fail_loudly_on_cpu_tensors():
    from rec_pangu_b200.models.ranking import DeepFM
    g = load_golden('deepfm')
    model = DeepFM(embedding_dim=g['meta']['D'], enc_dict=g['meta']['enc_dict'], **g['meta']['kwargs'])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model(g['data'])


def test_reset_parameters_follows_reference_init():
    """kaiming_normal_ on every >=2-D parameter incl. embedding tables: std = sqrt(2/fan_in) (base_model.py:42-59)."""
    from rec_pangu_b200.models.ranking import DeepFM
    torch.manual_seed(0)
    enc = make_enc(3, 2, 20000)
    m = DeepFM(embedding_dim=16, enc_dict=enc)
    w = m.embedding_layer.embedding_layer['C1'].weight
    assert w.shape == (20001, 16)
    assert abs(w.std().item() - (2.0 / 16) ** 0.5) < 0.01


def test_dataset_encoding_matches_reference_rules():
    from rec_pangu_b200.dataset import get_dataloader
    df = pd.DataFrame({'u': ['b', 'a', 'c', 'a', 'b', 'z'], 'i': [3, 1, 2, 1, 3, 3], 'x': [0.0, 1.0, 2.0, 3.0, 4.0, 10.0],
                       'click': [0, 1, 0, 1, 1, 0]})
    schema = {'sparse_cols': ['u', 'i'], 'dense_cols': ['x'], 'label_col': 'click', 'task_type': 'ranking'}
    tr, va, te, enc = get_dataloader(df[:5], df[:6], df[:6], schema, batch_size=4)
    assert enc['u'] == {'a': 0, 'b': 1, 'c': 2, 'vocab_size': 3}            # sorted-unique of str values
    assert enc['i'] == {'1': 0, '2': 1, '3': 2, 'vocab_size': 3}
    assert enc['x']['min'] == 0.0 and enc['x']['max'] == 4.0
    batch = next(iter(va))
    assert set(batch.keys()) == {'u', 'i', 'x', 'label'}
    assert batch['u'].dtype == torch.int64 and batch['x'].dtype == torch.float32 and batch['label'].dtype == torch.float32
    full = torch.cat([b['u'] for b in va])
    assert full.tolist() == [1, 0, 2, 0, 1, 3]                              # unseen 'z' -> OOV id = vocab_size
    xs = torch.cat([b['x'] for b in va])
    np.testing.assert_allclose(xs.numpy(), np.array([0, 1, 2, 3, 4, 10]) / (4.0 + 1e-5), rtol=1e-6)


def test_multitask_dataset_wire_format():
    from rec_pangu_b200.dataset import get_dataloader
    df = pd.DataFrame({'u': list('abcabc'), 'x': np.arange(6.0), 'click': [0, 1, 0, 1, 1, 0], 'scroll': [1, 1, 0, 0, 1, 0]})
    schema = {'sparse_cols': ['u'], 'dense_cols': ['x'], 'label_col': ['click', 'scroll'], 'task_type': 'multitask'}
    tr, va, te, enc = get_dataloader(df, df, df, schema, batch_size=6)
    b = next(iter(va))
    assert set(b.keys()) == {'u', 'x', 'task1_label', 'task2_label'}
    assert b['task2_label'].tolist() == [1, 1, 0, 0, 1, 0]


class _TorchStandIn(torch.nn.Module):
    """Tiny torch model with the reference's forward contract, used only to exercise trainer/pipeline plumbing on CPU."""

    def __init__(self, enc):
        super().__init__()
        self.cols = [c for c in enc if 'min' in enc[c]]
        self.fc = torch.nn.Linear(len(self.cols), 1)

    def forward(self, data, is_training=True):
        x = torch.stack([data[c] for c in self.cols], dim=1)
        pred = torch.sigmoid(self.fc(x))
        out = {'pred': pred}
        if is_training:
            out['loss'] = torch.nn.functional.binary_cross_entropy(pred.squeeze(-1), data['label'])
        return out


def test_train_and_test_model_plumbing_and_metric_keys():
    from rec_pangu_b200.dataset import get_dataloader
    from rec_pangu_b200.model_pipeline import train_model, test_model
    rng = np.random.default_rng(0)
    n = 200
    df = pd.DataFrame({'u': rng.integers(0, 5, n).astype(str), 'x': rng.random(n), 'y': rng.random(n)})
    df['click'] = (df['x'] + 0.1 * rng.random(n) > 0.5).astype(int)
    schema = {'sparse_cols': ['u'], 'dense_cols': ['x', 'y'], 'label_col': 'click', 'task_type': 'ranking'}
    tr, va, te, enc = get_dataloader(df, df, df, schema, batch_size=64)
    model = _TorchStandIn(enc)
    opt = torch.optim.Adam(model.parameters(), lr=0.05)
    for _ in range(5):
        m = train_model(model, tr, opt, torch.device('cpu'))
    assert set(m.keys()) == {'train_roc_auc_score', 'train_log_loss'}
    t = test_model(model, te, torch.device('cpu'))
    assert set(t.keys()) == {'roc_auc_score', 'log_loss'}
    assert t['roc_auc_score'] > 0.8


def test_trainer_refuses_to_compute_without_cuda():
    from rec_pangu_b200.trainer import RankTrainer
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        RankTrainer(num_task=1).fit(torch.nn.Linear(1, 1), [], None, epoch=1)


def test_c_abi_library_exports_every_declared_symbol():
    from rec_pangu_b200 import _lib
    lib = _lib.load()                       # dlopen works without a GPU (static cudart, no libcuda link)
    hdr = open(os.path.join(ROOT, 'include', 'rec_pangu_b200.h')).read()
    declared = set(re.findall(r'\b(rpb_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    assert declared == set(_lib.SIGNATURES.keys()), declared ^ set(_lib.SIGNATURES.keys())
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rpb_version() == _lib.ABI_VERSION
    assert b'bad argument' in lib.rpb_error_string(-2)


def test_index_routing_restatement_is_self_consistent():
    import oracle
    rng = np.random.default_rng(1)
    raw = rng.integers(-2 ** 62, 2 ** 62, 10000, dtype=np.int64)
    V, G = 100_000_000, 8
    rows = oracle.hash_to_row(raw, V)
    assert rows.dtype == np.int64 and rows.min() >= 0 and rows.max() < V
    assert np.array_equal(rows, oracle.hash_to_row(raw.copy(), V))          # deterministic
    owner, local = oracle.shard_route(rows, G)
    assert np.array_equal(local * G + owner, rows)                           # bijection idx <-> (owner, local)
    perm, counts, loc = oracle.bucket_by_owner(rows, G)
    assert counts.sum() == len(rows) and np.all(np.diff(owner[perm]) >= 0)
    assert np.array_equal(loc, local[perm])
    assert oracle.bounds_check(rows, V + 1) and not oracle.bounds_check(np.array([V + 1]), V + 1)
    # known-answer vector for the splitmix64 finaliser (first outputs of the reference splitmix64 sequence seeded 0)
    assert int(oracle.splitmix64(np.array([0], dtype=np.uint64))[0]) == 0xE220A8397B1DCDAF


def test_bench_line_stays_strict_json_and_both_arms_print_the_same_config():
    """bench.py: non-finite numbers of a secondary leg are stringified (the headline line stays strict JSON), and the
    `config` object is built by ONE function for both arms (the driver compares them)."""
    import json
    import bench
    line = {'value': 1.0, 'leg': bench._finite({'a': float('nan'), 'b': [float('inf'), 2.0]})}
    assert json.loads(json.dumps(line, allow_nan=False))['leg']['a'] == 'nan'
    for name, w in bench.WORKLOADS.items():
        c1, c8 = bench.make_config(w, 1), bench.make_config(w, 8)
        assert c1['workload'] == c8['workload'] and 'model' not in c1
        assert c8['global_batch'] == 8 * c1['batch_per_gpu']
        assert bench.alg_bytes_per_sample(w) >= w['F'] * (8 + 4 * w['D'])
    assert bench.alg_bytes_per_sample(bench.WORKLOADS['deepfm']) == 1928
    assert bench.alg_bytes_per_sample(bench.WORKLOADS['autoint']) == 3592 + 4 * 26


def test_reference_arm_drives_the_real_reference_classes_when_oracle_ref_is_built():
    """oracle/_ref (oracle/build_ref.py) holds the unmodified reference package; the loader refuses to mix it with this
    repo's `rec_pangu` alias package.  Run in a child interpreter (both packages are called rec_pangu)."""
    import subprocess
    import sys
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('oracle/_ref not built here (python oracle/build_ref.py needs /root/reference)')
    code = ("import sys; sys.path.insert(0, %r); from oracle import ref_loader; r, m = ref_loader.load(); "
            "import rec_pangu, os; assert '_ref' in rec_pangu.__file__; print(r.DeepFM.__module__, m.MMOE.__module__)" % ROOT)
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300, cwd='/tmp')
    assert out.returncode == 0, out.stderr[-500:]
    assert 'rec_pangu.models.ranking.deepfm' in out.stdout
