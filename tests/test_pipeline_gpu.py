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
    """model_pipeline._BatchStager: one pinned staging buffer + one H2D copy per dtype on a copy stream; device views of
    consecutive batches stay intact while they are in use (double buffering), values are bit-identical to per-key
    .to(device); columns that already are the rows of one pinned tensor are copied without re-packing."""
    from rec_pangu_b200.model_pipeline import _BatchStager, _common_pinned_base
    dev = torch.device('cuda', torch.cuda.current_device())
    st = _BatchStager(dev)
    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    g = torch.Generator().manual_seed(0)
    batches = []
    for b in range(5):
        n = 512 if b < 4 else 77                                     # last batch is ragged
        d = {f's{i}': torch.randint(0, 1000, (n,), generator=g) for i in range(6)}
        d.update({f'd{i}': torch.rand(n, generator=g) for i in range(3)})
        d['label'] = (torch.rand(n, generator=g) < 0.3).float()
        batches.append(d)
    # batch 2 comes from a "columnar loader": every column a row view of one pinned buffer per dtype
    buf_i = torch.stack([batches[2][f's{i}'] for i in range(6)]).pin_memory()
    buf_f = torch.stack([batches[2][k] for k in ('d0', 'd1', 'd2', 'label')]).pin_memory()
    col = {f's{i}': buf_i[i] for i in range(6)}
    col.update({k: buf_f[i] for i, k in enumerate(('d0', 'd1', 'd2', 'label'))})
    assert _common_pinned_base([col[f's{i}'] for i in range(6)]) is buf_i
    assert _common_pinned_base([batches[0][f's{i}'] for i in range(6)]) is None
    batches[2] = col
    staged = []
    for d in batches:
        t, out = st.stage(d, copy_stream)
        main.wait_event(st.ready[t])
        kept = {k: v.clone() for k, v in out.items()}                # what a step would read from the staging views
        ev = torch.cuda.Event()
        ev.record(main)
        st.done[t] = ev
        staged.append((d, kept))
    torch.cuda.synchronize()
    for d, kept in staged:
        assert list(kept.keys()) == list(d.keys())
        for k in d:
            assert torch.equal(kept[k].cpu(), d[k]), k


class _Loader:
    """In-memory loader of host dict batches with the two attributes train_model reads (dataset, batch_size)."""

    def __init__(self, batches, batch_size):
        self.batches, self.batch_size = batches, batch_size
        self.dataset = range(len(batches) * batch_size)

    def __iter__(self):
        return (dict(b) for b in self.batches)


@pytest.mark.parametrize('model_name', ['DeepFM', 'xDeepFM'])
def test_train_model_graph_replay_equals_eager_launches(model_name, monkeypatch):
    """model_pipeline.train_model with a graph-safe optimizer (FusedAdam) captures the step once per staging buffer and
    replays it (reference loop: model_pipeline.py:47-58).  Same batches through RPB_TRAIN_GRAPH=0 (eager launches): same
    parameters afterwards.  xDeepFM runs eval-free: its MLP dropout draws per-replay masks (device epoch), so only DeepFM
    (dropout 0) is compared bit-tight; xDeepFM must stay finite and move."""
    from helpers import make_enc, make_batch
    from rec_pangu_b200.models import ranking
    from rec_pangu_b200.model_pipeline import train_model
    from rec_pangu_b200.optim import FusedAdam
    enc = make_enc(26, 13, 500)
    B = 1024
    batches = [make_batch(enc, B, seed=100 + i) for i in range(7)]
    finals = []
    for graph in ('1', '0'):
        monkeypatch.setenv('RPB_TRAIN_GRAPH', graph)
        torch.manual_seed(3)
        model = getattr(ranking, model_name)(embedding_dim=16, enc_dict=enc).cuda()
        opt = FusedAdam(model, lr=1e-2)
        res = train_model(model, _Loader(batches, B), opt, torch.device('cuda'), metric_list=['log_loss'], log_rounds=3)
        assert 'train_log_loss' in res and np.isfinite(res['train_log_loss'])
        finals.append({k: v.detach().clone() for k, v in model.state_dict().items()})
    moved = 0
    for k in finals[0]:
        a, b = finals[0][k].float(), finals[1][k].float()
        assert torch.isfinite(a).all(), k
        if model_name == 'DeepFM':
            assert (a - b).abs().max().item() <= 1e-5 * max(1.0, b.abs().max().item()), k
        moved += int((a - b).abs().max().item() >= 0)
    assert moved == len(finals[0])


def test_train_model_reports_out_of_range_ids_like_the_reference():
    """An id beyond the table raises IndexError from train_model / test_model (ops.check_index_errors at the points that
    synchronise anyway) instead of silently training on row 0 (reference: aten::embedding raises on CPU / asserts on CUDA)."""
    from helpers import make_enc, make_batch
    from rec_pangu_b200.models.ranking import DeepFM
    from rec_pangu_b200.model_pipeline import train_model, test_model
    enc = make_enc(4, 2, 100)
    batches = [make_batch(enc, 256, seed=i) for i in range(3)]
    batches[1]['C2'][17] = 100 + 7
    model = DeepFM(embedding_dim=16, hidden_units=[64, 64], enc_dict=enc).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    with pytest.raises(IndexError):
        train_model(model, _Loader(batches, 256), opt, torch.device('cuda'), log_rounds=100)
    with pytest.raises(IndexError):
        test_model(model, _Loader(batches, 256), torch.device('cuda'))


def test_graph_replays_draw_fresh_dropout_masks():
    """runtime.GraphedStep on a model with active dropout (xDeepFM, MLP dropout 0.1): the host-drawn seeds are frozen by the
    capture, the device-side epoch the kernels mix in is advanced inside the captured step, so two replays on the SAME batch
    and weights give different training-mode predictions (and a third differs from both)."""
    from helpers import make_enc, make_batch
    from rec_pangu_b200.models.ranking import xDeepFM
    from rec_pangu_b200.runtime import ColumnarBatch, GraphedStep
    enc = make_enc(8, 3, 200)
    torch.manual_seed(1)
    model = xDeepFM(embedding_dim=16, enc_dict=enc).cuda()
    model.set_grad_mode('persistent')
    model.train()
    cb = ColumnarBatch(enc, 512, device='cuda', pinned_host=False)
    cb.load_device(make_batch(enc, 512, seed=5, device='cuda'))
    step = GraphedStep(model, cb)                 # no optimizer: weights stay fixed, only the masks may change
    assert step.graph is not None
    preds = []
    for _ in range(3):
        step.replay()
        torch.cuda.synchronize()
        preds.append(step.pred.detach().clone())
    assert not torch.equal(preds[0], preds[1]) and not torch.equal(preds[1], preds[2]) and not torch.equal(preds[0], preds[2])


def test_graphed_ring_zero_first_leaves_exactly_the_last_steps_gradients():
    """GraphedStep.ring(zero_first=True): step j opens with the sparse re-zero of the rows step j-1 touched, on a side stream
    next to the forward kernel.  After any number of ring-ordered replays the table and dense gradients must be exactly those
    of the LAST step alone (nothing stale from earlier batches, nothing zeroed that the last step wrote)."""
    from helpers import make_enc, make_batch
    from rec_pangu_b200.models.ranking import DeepFM
    from rec_pangu_b200.runtime import ColumnarBatch, GraphedStep
    enc = make_enc(8, 3, 5000)
    torch.manual_seed(3)
    model = DeepFM(embedding_dim=16, hidden_units=[64, 64, 64], enc_dict=enc).cuda()
    model.set_grad_mode('persistent')
    model.train()
    B, NB = 2048, 3
    cbs = []
    for i in range(NB):
        cb = ColumnarBatch(enc, B, device='cuda', pinned_host=False)
        cb.load_device(make_batch(enc, B, seed=20 + i, device='cuda'))
        cbs.append(cb)
    steps = GraphedStep.ring(model, cbs, zero_first=True)
    assert all(s.graph is not None for s in steps)
    for k in range(2 * NB + 2):                       # ends on ring position 1
        steps[k % NB].replay()
    torch.cuda.synchronize()
    last = (2 * NB + 1) % NB
    got = {n: g.detach().clone() for n, g in steps[last].grads.items()}      # the step's own tensors (tables: the persistent buffers)
    # reference: the same step alone, eagerly, on zeroed gradients
    for p in model.parameters():
        if p.grad is not None:
            p.grad.zero_()
    model.embedding_layer._grad_store.pending = []      # host-side list of touched rows: stale after graph replays, all rows are clean now
    out = model(cbs[last].as_dict())
    out['loss'].backward()
    torch.cuda.synchronize()
    n_tables = 0
    for n, p in model.named_parameters():
        assert n in got, n
        n_tables += 'embedding_layer' in n
        torch.testing.assert_close(got[n], p.grad, rtol=1e-5, atol=1e-7, msg=lambda m, n=n: f'{n}: {m}')
    assert n_tables >= 8
