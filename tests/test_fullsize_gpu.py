"""Parity at the FULL BASELINE.json sizes through size-independent properties (the oracle cannot run 65536 x 26 x 1M-row
steps in seconds, so these tests tie the full-size CUDA run back to it):

* analytic tables: table[f][row, d] is a closed form of (f, row, d), so the gathered feature row is known bit-exactly
  from the indices alone (config 2 shape; config 5's 100 M-row hashed table);
* sample independence: a random subset of the full batch, re-encoded onto compact tables holding only the rows it
  touches, goes through the CPU oracle — logits of those samples inside the full-size GPU run must match (1e-4);
* additivity of the batch-summed gradients: grad(full batch) == mean of grad(halves);
* checksum of checksums: the sum of every table-gradient row equals the column sums of the per-sample row gradients;
* idempotence: zero_grad leaves the persistent gradient buffers all-zero; a second zero_grad changes nothing.
"""
import numpy as np
import pytest
import torch

import oracle
from helpers import make_enc

pytestmark = pytest.mark.gpu

F, ND, V, D, B = 26, 13, 1_000_000, 16, 65536


def _full_batch(enc, B, seed, vocab=V):
    g = torch.Generator(device='cuda').manual_seed(seed)
    d = {}
    for c, m in enc.items():
        if 'vocab_size' in m:
            d[c] = torch.randint(0, m['vocab_size'] + 1, (B,), dtype=torch.int64, device='cuda', generator=g)
        else:
            d[c] = torch.rand(B, device='cuda', generator=g)
    d['label'] = (torch.rand(B, device='cuda', generator=g) < 0.25).float()
    return d


def _compact(model_sd_rows, enc, data, sel):
    """Re-encode the samples `sel` onto compact tables that hold only the rows they touch (index -> position in the
    sorted unique list); returns (small enc_dict, small batch on CPU, {column: unique rows})."""
    small_enc, small, uniq = {}, {}, {}
    for c, m in enc.items():
        if 'vocab_size' in m:
            ids = data[c][sel].cpu()
            u, inv = torch.unique(ids, return_inverse=True)
            uniq[c] = u
            small[c] = inv
            small_enc[c] = {'vocab_size': int(u.numel()) - 1}       # table rows = vocab_size + 1 = len(u)
        else:
            small[c] = data[c][sel].cpu()
            small_enc[c] = dict(m)
    small['label'] = data['label'][sel].cpu()
    return small_enc, small, uniq


def test_config2_gather_is_bit_exact_on_analytic_tables_and_scatter_checksums():
    from rec_pangu_b200 import ops
    enc = make_enc(F, ND, V)
    data = _full_batch(enc, B, seed=11)
    cols, dcols = oracle.sparse_cols(enc), oracle.dense_cols(enc)
    rows = torch.arange(V + 1, device='cuda', dtype=torch.int32)
    tables = []
    for f in range(F):
        # integers below 2^24 are exact in fp32: (row * 16 + d) mod 2^24, shifted per field
        t = ((rows.long() * D + f * 7919) % (1 << 24)).float().unsqueeze(1) + torch.arange(D, device='cuda').float()
        tables.append(t.contiguous().requires_grad_(True))
    x, fm, _ = ops.gather(tables, [data[c] for c in cols], [data[c] for c in dcols], want_fm=True)
    ops.check_index_errors()
    idx = torch.stack([data[c] for c in cols], dim=1).cpu().numpy()                       # [B, F]
    exp = ((idx * D + (np.arange(F) * 7919)[None, :]) % (1 << 24)).astype(np.float32)[:, :, None] + \
        np.arange(D, dtype=np.float32)[None, None, :]
    assert np.array_equal(x.detach()[:, :F * D].view(B, F, D).cpu().numpy(), exp)         # pure copy => bit exact
    assert torch.equal(x.detach()[:, F * D:F * D + ND], torch.stack([data[c] for c in dcols], dim=1))
    # scatter: every sample adds w[b] * ones to its 26 rows => sum over all rows of grad[f] == sum_b w[b] per column
    w = torch.rand(B, device='cuda')
    (x[:, :F * D] * w.unsqueeze(1)).sum().backward()
    tot = float(w.double().sum())
    for f in (0, 7, 25):
        g = tables[f].grad
        cs = g.double().sum(dim=0)
        assert torch.allclose(cs, torch.full_like(cs, tot), rtol=1e-5), (f, cs, tot)
        # rows never touched stay exactly zero; touched rows hold the sum of their samples' weights
        touched = torch.zeros(V + 1, dtype=torch.bool, device='cuda')
        touched[data[cols[f]]] = True
        assert torch.count_nonzero(g[~touched]) == 0
        ref = torch.zeros(V + 1, device='cuda', dtype=torch.float64).index_add_(0, data[cols[f]], w.double())
        assert torch.allclose(g[:, 3].double(), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('model_name,kw,okw,Bm', [
    ('DeepFM', dict(embedding_dim=16, hidden_units=[64, 64, 64]), dict(hidden_units=(64, 64, 64)), 65536),
    ('xDeepFM', dict(embedding_dim=16), {}, 65536),
    ('AutoInt', dict(embedding_dim=32, num_heads=3), dict(num_heads=3), 32768),
])
def test_fullsize_step_subset_matches_oracle_and_gradients_are_additive(model_name, kw, okw, Bm):
    """configs 2, 3, 4 at their BASELINE.json shapes (26 x 1M-row tables, batch 65536 / 32768)."""
    from rec_pangu_b200.models import ranking
    from rec_pangu_b200 import ops
    enc = make_enc(F, ND, V)
    torch.manual_seed(1029)
    with torch.device('cuda'):
        model = getattr(ranking, model_name)(enc_dict=enc, **kw)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.copy_(torch.randn(p.shape, device='cuda') * 0.05)
            elif 'embedding_layer' in n:
                p.mul_(0.25 if p.shape[1] > 1 else 0.1)          # keep logits in the conditioned range (see test_models_gpu)
    model.eval()                                                 # dropout off (xDeepFM / AutoInt default p = 0.1)
    model.set_grad_mode('persistent')
    data = _full_batch(enc, Bm, seed=5)
    out = model(data)
    out['loss'].backward()
    ops.check_index_errors()
    logit_full = model._last_logit.clone()
    assert torch.isfinite(out['loss']) and float(out['pred'].min()) >= 0.0 and float(out['pred'].max()) <= 1.0
    dense_names = [n for n, _ in model.named_parameters() if 'embedding_layer' not in n]
    g_full = {n: p.grad.clone() for n, p in model.named_parameters() if n in dense_names}
    tname = [n for n, _ in model.named_parameters() if n.endswith('embedding_layer.C3.weight')][0]
    t_full = dict(model.named_parameters())[tname].grad.clone()

    # ---- sample independence + oracle on a subset re-encoded onto compact tables
    sel = torch.randperm(Bm, generator=torch.Generator().manual_seed(1))[:384].cuda()
    small_enc, small, uniq = _compact(None, enc, data, sel)
    sd = {}
    for k, v in model.state_dict().items():
        col = k.split('.')[-2]
        if 'embedding_layer' in k and col in uniq:
            sd[k] = v[uniq[col].cuda()].cpu()
        else:
            sd[k] = v.cpu()
    ref = oracle.MODEL_FORWARDS[model_name](sd, small_enc, small, **okw)
    dl = (logit_full[sel].cpu().double() - ref['logit'].double()).abs().max().item()
    assert dl <= 1e-4, f'{model_name}: max |dlogit| of the subset inside the full-size run vs oracle = {dl}'

    # ---- additivity: grad of the full-batch mean loss == mean of the two half-batch gradients
    model.zero_grad()
    acc = {n: torch.zeros_like(g) for n, g in g_full.items()}
    t_acc = torch.zeros_like(t_full)
    for h in range(2):
        half = {k: v[h * (Bm // 2):(h + 1) * (Bm // 2)].contiguous() for k, v in data.items()}
        model(half)['loss'].backward()
        for n, p in model.named_parameters():
            if n in acc:
                acc[n] += 0.5 * p.grad
        t_acc += 0.5 * dict(model.named_parameters())[tname].grad
        model.zero_grad()
    for n in dense_names:
        scale = max(1e-6, g_full[n].abs().max().item())
        assert (acc[n] - g_full[n]).abs().max().item() <= 2e-4 * scale, n
    assert (t_acc - t_full).abs().max().item() <= 2e-4 * max(1e-9, t_full.abs().max().item())

    # ---- idempotence of the sparse re-zero
    for buf in model.embedding_layer._grad_store.buffers.values():
        assert torch.count_nonzero(buf) == 0
    model.zero_grad()
    for buf in model.embedding_layer._grad_store.buffers.values():
        assert torch.count_nonzero(buf) == 0


def test_config5_hashed_ids_and_100m_row_table_gather():
    """config 5: raw ids -> splitmix64 mod V (bit-exact vs oracle/index_routing.py), gather from a 100 M-row x 40 table
    (16 GB, analytic content), shard routing of the hashed rows for 8 GPUs (owner = row mod 8, local = row div 8)."""
    from rec_pangu_b200 import ops
    Vh, Dh, Bh = 100_000_000, 40, 32768
    g = torch.Generator(device='cuda').manual_seed(3)
    raw = torch.randint(-2 ** 62, 2 ** 62, (Bh,), dtype=torch.int64, device='cuda', generator=g)
    raw[:4] = torch.tensor([0, -1, 2 ** 62 - 1, -2 ** 62], device='cuda')
    rows = ops.hash_to_row(raw, Vh)
    ref_rows = oracle.hash_to_row(raw.cpu().numpy(), Vh)
    assert np.array_equal(rows.cpu().numpy(), ref_rows)                       # integer work: bit exact
    assert int(rows.min()) >= 0 and int(rows.max()) < Vh
    odd = ops.hash_to_row(raw[1:], Vh)                                        # unaligned (8-byte) path
    assert torch.equal(odd, rows[1:])
    owner, local = oracle.shard_route(ref_rows, 8)
    assert np.array_equal(owner.astype(np.int64) + 8 * local, ref_rows)       # routing is a bijection
    # 100 M-row table with closed-form content
    r = torch.arange(Vh + 1, device='cuda', dtype=torch.int32)
    table = ((r % 65536).float() * Dh).unsqueeze(1) + torch.arange(Dh, device='cuda').float()      # 16 GB, exact integers
    del r
    dense = torch.rand(Bh, device='cuda')
    x, _, _ = ops.gather([table], [rows], [dense])
    ops.check_index_errors()
    exp = ((ref_rows % 65536).astype(np.float32) * Dh)[:, None] + np.arange(Dh, dtype=np.float32)[None, :]
    assert np.array_equal(x.detach()[:, :Dh].cpu().numpy(), exp)
    assert torch.equal(x.detach()[:, Dh], dense)
    del table, x
    torch.cuda.empty_cache()
