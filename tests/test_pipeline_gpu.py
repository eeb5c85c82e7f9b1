"""BASELINE.json config 1 (plumbing): the reference's examples/ranking flow — get_dataloader -> Model(enc_dict) ->
RankTrainer.fit -> save_all -> evaluate_model -> predict_dataframe -> reload — on a demo-sized frame with the example's
schema shape (16 sparse + 9 dense columns), through the CUDA models.  The reference's 100-row CSV cannot travel to the
GPU box, so an equally shaped synthetic frame is generated here."""
import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu


def _frame(n=300, seed=0):
    rng = np.random.default_rng(seed)
    sparse = [f's{i}' for i in range(16)]
    dense = [f'd{i}' for i in range(9)]
    df = pd.DataFrame({c: rng.integers(0, 7 + i, n).astype(str) for i, c in enumerate(sparse)})
    for c in dense:
        df[c] = rng.random(n) * 10
    df['click'] = ((df['d0'] + df['d1'] + rng.random(n) * 4) > 11).astype(int)
    schema = {'sparse_cols': sparse, 'dense_cols': dense, 'label_col': 'click', 'task_type': 'ranking'}
    return df, schema


@pytest.mark.parametrize('model_name', ['DeepFM', 'xDeepFM', 'DCN'])
def test_examples_ranking_flow(model_name, tmp_path):
    from rec_pangu.dataset import get_dataloader
    from rec_pangu.models import ranking
    from rec_pangu.trainer import RankTrainer
    df, schema = _frame()
    train_loader, valid_loader, test_loader, enc_dict = get_dataloader(df[:240], df[:270], df[:285], schema, batch_size=512)
    torch.manual_seed(0)
    model = getattr(ranking, model_name)(embedding_dim=8, enc_dict=enc_dict)
    trainer = RankTrainer(num_task=1, model_ckpt_dir=str(tmp_path))
    # the reference's example passes device=cpu; this build redirects it to the GPU (no CPU compute path)
    valid_metric = trainer.fit(model, train_loader, valid_loader, epoch=12, lr=1e-2, device=torch.device('cpu'),
                               use_earlystopping=True, max_patience=50, monitor_metric='roc_auc_score')
    assert set(valid_metric.keys()) == {'roc_auc_score', 'log_loss'}
    assert valid_metric['roc_auc_score'] > 0.7            # the label is learnable from d0 + d1
    trainer.save_all(model, enc_dict, str(tmp_path))
    test_metric = trainer.evaluate_model(model, test_loader, device=torch.device('cuda'))
    assert 0.0 <= test_metric['log_loss'] < 2.0
    preds = trainer.predict_dataframe(model, df[:285], enc_dict, schema)
    assert len(preds) == 285 and all(0.0 <= float(p) <= 1.0 for p in preds)
    ckpt = torch.load(str(tmp_path / 'model.pth'), weights_only=False)
    assert set(ckpt.keys()) == {'model', 'enc_dict'}
    model2 = getattr(ranking, model_name)(embedding_dim=8, enc_dict=ckpt['enc_dict'])
    model2.load_state_dict(ckpt['model'])
    preds2 = trainer.predict_dataframe(model2.cuda(), df[:285], enc_dict, schema)
    np.testing.assert_allclose(np.array(preds2, dtype=np.float64), np.array(preds, dtype=np.float64), rtol=1e-5, atol=1e-6)


def test_multitask_flow_mmoe(tmp_path):
    from rec_pangu.dataset import get_dataloader
    from rec_pangu.models.multi_task import MMOE
    from rec_pangu.trainer import RankTrainer
    df, schema = _frame(400, seed=1)
    df['scroll'] = ((df['d2'] + np.random.default_rng(2).random(400) * 3) > 6).astype(int)
    schema = dict(schema, label_col=['click', 'scroll'], task_type='multitask')
    tr, va, te, enc = get_dataloader(df[:320], df[:360], df[:380], schema, batch_size=128)
    torch.manual_seed(0)
    model = MMOE(embedding_dim=8, mmoe_hidden_dim=32, hidden_dim=[32, 16], enc_dict=enc, device='cuda')
    trainer = RankTrainer(num_task=2, model_ckpt_dir=str(tmp_path))
    m = trainer.fit(model, tr, va, epoch=3, lr=1e-3, device=torch.device('cuda'))
    assert set(m.keys()) == {'test_task1_roc_auc_score', 'test_task1_log_loss', 'test_task2_roc_auc_score', 'test_task2_log_loss'}
    preds = trainer.predict_dataframe(model, df[:380], enc, schema, device=torch.device('cuda'))
    assert len(preds) == 2 and len(preds[0]) == 380


def test_batch_stager_packs_columns_and_survives_reuse():
    """model_pipeline._to_device: one pinned staging buffer + one H2D copy per dtype; device views of consecutive batches
    stay intact while they are in use (double buffering), values are bit-identical to per-key .to(device)."""
    from rec_pangu_b200.model_pipeline import _BatchStager
    st = _BatchStager()
    dev = torch.device('cuda', torch.cuda.current_device())
    g = torch.Generator().manual_seed(0)
    batches = []
    for b in range(5):
        n = 512 if b < 4 else 77                                     # last batch is ragged
        d = {f's{i}': torch.randint(0, 1000, (n,), generator=g) for i in range(6)}
        d.update({f'd{i}': torch.rand(n, generator=g) for i in range(3)})
        d['label'] = (torch.rand(n, generator=g) < 0.3).float()
        batches.append(d)
    prev = None
    for d in batches:
        ref = {k: v.clone() for k, v in d.items()}
        out = st(dict(d), dev)
        assert list(out.keys()) == list(ref.keys())
        for k in ref:
            assert out[k].is_cuda and out[k].dtype == ref[k].dtype and torch.equal(out[k].cpu(), ref[k])
        if prev is not None:                                          # the previous batch's views were not overwritten
            for k in prev[1]:
                assert torch.equal(prev[0][k].cpu(), prev[1][k])
        prev = (out, ref)
    # two int64 columns of one batch are rows of ONE device buffer
    o = st({k: v for k, v in batches[0].items()}, dev)
    assert o['s1'].data_ptr() - o['s0'].data_ptr() == 512 * 8


def test_trainer_with_fused_adam(tmp_path):
    from rec_pangu.dataset import get_dataloader
    from rec_pangu.models.ranking import DeepFM
    from rec_pangu.trainer import RankTrainer
    df, schema = _frame()
    train_loader, valid_loader, _, enc_dict = get_dataloader(df[:240], df[:270], df[:285], schema, batch_size=512)
    torch.manual_seed(0)
    model = DeepFM(embedding_dim=8, enc_dict=enc_dict)
    trainer = RankTrainer(num_task=1, model_ckpt_dir=str(tmp_path))
    m = trainer.fit(model, train_loader, valid_loader, epoch=12, lr=1e-2, device=torch.device('cuda'), optimizer_type='fused_adam')
    assert m['roc_auc_score'] > 0.7
    with pytest.raises(ValueError):
        trainer.fit(model, train_loader, valid_loader, epoch=1, optimizer_type='fused_adam', lr_scheduler_type='StepLR',
                    scheduler_params={'step_size': 1})
