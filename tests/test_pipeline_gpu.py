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
