"""Launched by torchrun (>= 2 GPUs): row-sharded peer-memory DeepFM step == single-GPU step on the global batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.distributed as dist

from helpers import make_enc, make_batch
from rec_pangu_b200 import dist as rdist, ops
from rec_pangu_b200.models.ranking import DeepFM


def check(rank, world, dev, hidden, D=8, B=96, fused=False):
    """hidden = [16, 8]: layer-by-layer MLP kernels; [64, 64]: the fused tower-tail kernels (single-GPU reference with the
    loss fused into the DeepFM core node, sharded model through MLP + the separate sigmoid/BCE head).
    fused (D = 16, B >= 512, 64-wide tower): the sharded model runs the fused core too (ops.SHARDED_FUSED): one-kernel
    forward with remote row requests, dx GEMM with the scatter epilogue into the owners' gradient shards."""
    enc = make_enc(6, 3, [101, 57, 33, 200, 17, 64])
    torch.manual_seed(7)
    ref = DeepFM(embedding_dim=D, hidden_units=hidden, enc_dict=enc)
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if 'embedding_layer' in n:
                p.mul_(0.3)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    ref = ref.to(dev)
    model = DeepFM(embedding_dim=D, hidden_units=hidden, enc_dict=enc)
    model.load_state_dict(sd)
    model = model.to(dev)
    st = rdist.shard_model_tables(model)
    dense_params = [p for n, p in model.named_parameters() if not n.startswith('embedding_layer.')]
    bucket = rdist.DenseGradBucket(dense_params)

    full = make_batch(enc, B * world, seed=3, device=dev)
    mine = {k: v[rank * B:(rank + 1) * B].contiguous() for k, v in full.items()}
    for step in range(2):
        # reference: single GPU, global batch
        ref.zero_grad()
        ro = ref(full)
        ro['loss'].backward()
        # sharded
        model.zero_grad()
        prev = ops.SHARDED_FUSED
        ops.SHARDED_FUSED = 1 if fused else 0
        n0 = ops.launch_count()
        out = model(mine)
        n_fwd = ops.launch_count() - n0
        (out['loss'] / world).backward()
        ops.SHARDED_FUSED = prev
        bucket.all_reduce()
        ops.check_index_errors(dev)
        if fused:
            assert n_fwd == 3, f'fused sharded forward should be 2 weight splits + one kernel, saw {n_fwd} launches'
        torch.testing.assert_close(out['pred'], ro['pred'][rank * B:(rank + 1) * B], rtol=1e-5, atol=1e-6)
        for f, c in enumerate(model.embedding_layer.emb_feature):
            g_full = st.full_grad(f)
            r_full = ref.embedding_layer.embedding_layer[c].weight.grad
            err = (g_full - r_full).abs().max().item()
            assert err <= 1e-5 * max(1.0, r_full.abs().max().item()) + 1e-8, (step, c, err)
            # the published .grad of the local shard is the owner's slice of the full gradient
            p = model.embedding_layer.embedding_layer[c].weight
            torch.testing.assert_close(p.grad, rdist.local_slice(r_full, rank, world), rtol=1e-4, atol=1e-7)
        ref_dense = {n: p.grad for n, p in ref.named_parameters() if not n.startswith('embedding_layer.')}
        for n, p in model.named_parameters():
            if not n.startswith('embedding_layer.'):
                torch.testing.assert_close(p.grad, ref_dense[n], rtol=1e-4, atol=1e-6, msg=lambda s: f'{n}: {s}')
        # weights round trip
        for f, c in enumerate(model.embedding_layer.emb_feature):
            assert torch.equal(st.full_table(f), ref.embedding_layer.embedding_layer[c].weight.data)
    # zero_grad leaves every gradient shard all-zero again
    model.zero_grad()
    for f in range(len(st.cols)):
        assert torch.count_nonzero(st.full_grad(f)) == 0
    dist.barrier()


def check_mmoe_sync_bn(rank, world, dev):
    """MMOE in train mode (BatchNorm towers use BATCH statistics): data-parallel ranks with cross-GPU statistics
    (dist.enable_sync_batchnorm) == the single-process model on the concatenated batch."""
    from rec_pangu_b200.models.multi_task import MMOE
    enc = make_enc(5, 2, [101, 57, 33, 200, 17])
    B = 64
    kw = dict(embedding_dim=8, enc_dict=enc, mmoe_hidden_dim=16, hidden_dim=[16, 8], dropouts=[0.0, 0.0], device='cpu')
    torch.manual_seed(11)
    ref = MMOE(**kw)
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if 'embedding_layer' in n:
                p.mul_(0.3)
            elif n in ('experts', 'experts_bias'):
                p.mul_(0.3)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    gates = [g.detach().clone() for g in ref.gates]
    gates_bias = [g.detach().clone() for g in ref.gates_bias]
    model = MMOE(**kw)
    model.load_state_dict(sd)
    with torch.no_grad():
        for i in range(2):
            model.gates[i].copy_(gates[i])
            model.gates_bias[i].copy_(gates_bias[i])
    ref, model = ref.to(dev).train(), model.to(dev).train()
    st = rdist.shard_model_tables(model)
    bucket = rdist.DenseGradBucket([p for n, p in model.named_parameters() if not n.startswith('embedding_layer.')])
    full = make_batch(enc, B * world, seed=9, labels=('task1_label', 'task2_label'), device=dev)
    mine = {k: v[rank * B:(rank + 1) * B].contiguous() for k, v in full.items()}
    rdist.enable_sync_batchnorm(enabled=False)
    ro = ref(full)
    ro['loss'].backward()
    rdist.enable_sync_batchnorm()
    out = model(mine)
    (out['loss'] / world).backward()
    bucket.all_reduce()
    rdist.enable_sync_batchnorm(enabled=False)
    ops.check_index_errors(dev)
    for k in ('task1_pred', 'task2_pred'):
        torch.testing.assert_close(out[k], ro[k][rank * B:(rank + 1) * B], rtol=1e-4, atol=1e-5, msg=lambda s: f'{k}: {s}')
    ref_grads = {n: p.grad for n, p in ref.named_parameters()}
    for n, p in model.named_parameters():
        if not n.startswith('embedding_layer.'):
            r = ref_grads[n]
            err = (p.grad - r).abs().max().item()
            assert err <= 2e-4 * max(1e-6, r.abs().max().item()) + 1e-7, (n, err)
    for k, v in model.state_dict().items():
        if 'running_mean' in k or 'running_var' in k:
            torch.testing.assert_close(v, ref.state_dict()[k], rtol=1e-4, atol=1e-6, msg=lambda s: f'{k}: {s}')
    g_full = st.full_grad(2)
    r_full = ref.embedding_layer.embedding_layer['C3'].weight.grad
    assert (g_full - r_full).abs().max().item() <= 2e-4 * max(1e-6, r_full.abs().max().item()) + 1e-8
    model.zero_grad()
    dist.barrier()


def check_xdeepfm_lr_shards(rank, world, dev):
    """xDeepFM (LR_Layer with D = 1 tables + CIN + MLP) on row-sharded tables == the single-GPU model on the global batch:
    predictions, the gradient of every main table and of every LR table (un-sharded), every dense gradient.
    Reference: rec_pangu/models/ranking/xdeepfm.py:48-79, layers/shallow.py:14-27."""
    from rec_pangu_b200.models.ranking import xDeepFM
    enc = make_enc(6, 3, [101, 57, 33, 200, 17, 64])
    B, D = 96, 8
    torch.manual_seed(13)
    ref = xDeepFM(embedding_dim=D, enc_dict=enc)
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if 'embedding_layer' in n:
                p.mul_(0.3)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    ref = ref.to(dev).eval()                        # eval: the MLP's dropout(0.1) is off on both sides
    model = xDeepFM(embedding_dim=D, enc_dict=enc)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    st = rdist.shard_model_tables(model)
    st_lr = model.lr_layer.emb_layer._shards
    assert st_lr is not None and st_lr.D == 1
    is_table = lambda n: 'embedding_layer.' in n                      # noqa: E731
    bucket = rdist.DenseGradBucket([p for n, p in model.named_parameters() if not is_table(n)])
    full = make_batch(enc, B * world, seed=5, device=dev)
    mine = {k: v[rank * B:(rank + 1) * B].contiguous() for k, v in full.items()}
    ro = ref(full)
    ro['loss'].backward()
    out = model(mine)
    (out['loss'] / world).backward()
    bucket.all_reduce()
    ops.check_index_errors(dev)
    torch.testing.assert_close(out['pred'], ro['pred'][rank * B:(rank + 1) * B], rtol=1e-5, atol=1e-6)
    for f, c in enumerate(model.embedding_layer.emb_feature):
        for shards, tab in ((st, ref.embedding_layer.embedding_layer[c].weight), (st_lr, ref.lr_layer.emb_layer.embedding_layer[c].weight)):
            g_full, r_full = shards.full_grad(f), tab.grad
            err = (g_full - r_full).abs().max().item()
            assert err <= 1e-4 * max(1e-6, r_full.abs().max().item()) + 1e-8, (c, shards.D, err)
    ref_dense = {n: p.grad for n, p in ref.named_parameters() if not is_table(n)}
    for n, p in model.named_parameters():
        if not is_table(n):
            r = ref_dense[n]
            assert (p.grad - r).abs().max().item() <= 2e-4 * max(1e-6, r.abs().max().item()) + 1e-7, n
    # the reference-layout state_dict comes back bit for bit, LR tables included
    gsd = rdist.gather_state_dict(model)
    for k, v in sd.items():
        assert torch.equal(gsd[k].cpu(), v), k
    model.zero_grad()
    for f in range(len(st.cols)):
        assert torch.count_nonzero(st_lr.full_grad(f)) == 0
    dist.barrier()


def check_optimizer_then_zero_grad(rank, world, dev):
    """Two full steps — backward, torch.optim.Adam on the local shards + dense parameters, model.zero_grad() — on ranks that
    run at different speeds (rank 0 sleeps before its optimizer step) against the single-GPU model on the global batch:
    the re-zero of the gradient shards must not race with a slower rank's optimizer (ops.sharded_clean)."""
    import time
    enc = make_enc(6, 3, [101, 57, 33, 200, 17, 64])
    B, D, hidden = 640, 16, [64, 64]
    torch.manual_seed(17)
    ref = DeepFM(embedding_dim=D, hidden_units=hidden, enc_dict=enc)
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if 'embedding_layer' in n:
                p.mul_(0.3)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    ref = ref.to(dev)
    model = DeepFM(embedding_dim=D, hidden_units=hidden, enc_dict=enc)
    model.load_state_dict(sd)
    model = model.to(dev)
    st = rdist.shard_model_tables(model)
    bucket = rdist.DenseGradBucket([p for n, p in model.named_parameters() if not n.startswith('embedding_layer.')])
    opt_r = torch.optim.Adam(ref.parameters(), lr=1e-2)
    opt_m = torch.optim.Adam(model.parameters(), lr=1e-2)
    for step in range(2):
        full = make_batch(enc, B * world, seed=30 + step, device=dev)
        mine = {k: v[rank * B:(rank + 1) * B].contiguous() for k, v in full.items()}
        ref(full)['loss'].backward()
        opt_r.step()
        ref.zero_grad()
        (model(mine)['loss'] / world).backward()
        bucket.all_reduce()
        if rank == 0:
            torch.cuda.synchronize()
            time.sleep(0.2)                          # a slow rank: its optimizer reads its gradient shard late
        opt_m.step()
        model.zero_grad()
    for f, c in enumerate(model.embedding_layer.emb_feature):
        w_full = st.full_table(f)
        r = ref.embedding_layer.embedding_layer[c].weight.data
        err = (w_full - r).abs().max().item()
        # Adam's update is lr * m / sqrt(v): a row whose gradient is summed in another order moves by a slightly different
        # fraction of lr = 1e-2 (2.3e-5 at N = 8); a LOST gradient row (the race this guards against) moves by ~lr
        assert err <= 2e-4, (c, err)
    for (n, p), (_, q) in zip(model.dnn.named_parameters(), ref.dnn.named_parameters()):
        assert (p - q).abs().max().item() <= 2e-4, n
    dist.barrier()


def check_local_init(rank, world, dev):
    """dist.deferred_tables() + shard_model_tables(init='kaiming'): no full table exists anywhere, every shard is filled
    locally with the reference's init statistics (std = sqrt(2 / D)), padding rows are zero, the model steps."""
    enc = make_enc(4, 2, [50001, 30000, 777, 12345])
    with rdist.deferred_tables():
        model = DeepFM(embedding_dim=16, hidden_units=[64, 64], enc_dict=enc)
    assert all(p.device.type == 'meta' for n, p in model.named_parameters() if 'embedding_layer' in n)
    st = rdist.shard_model_tables(model, init='kaiming')
    model = model.to(dev)
    for f in range(4):
        w = st.weights[f]
        assert w.device.type == 'cuda' and w.shape[0] == rdist.shard_rows(st.rows[f], world)
        std = w[:len(range(rank, st.rows[f], world))].std().item()
        assert abs(std - (2.0 / 16) ** 0.5) < 0.02, (f, std)
    batch = make_batch(enc, 512, seed=40 + rank, device=dev)
    out = model(batch)
    out['loss'].backward()
    ops.check_index_errors(dev)
    assert torch.isfinite(out['loss']).item()
    model.zero_grad()
    dist.barrier()


def main():
    rank, world, local = rdist.init_from_env('nccl')
    dev = torch.device('cuda', local)
    for hidden in ([16, 8], [64, 64]):
        check(rank, world, dev, hidden)
    check(rank, world, dev, [64, 64], D=16, B=640, fused=True)          # fused core on sharded tables (the default for DeepFM)
    if rank == 0:
        print('SHARDED_FUSED_OK world', world, flush=True)
    check_xdeepfm_lr_shards(rank, world, dev)
    if rank == 0:
        print('SHARDED_LR_OK world', world, flush=True)
    check_optimizer_then_zero_grad(rank, world, dev)
    if rank == 0:
        print('SHARDED_OPT_OK world', world, flush=True)
    check_local_init(rank, world, dev)
    check_mmoe_sync_bn(rank, world, dev)
    if rank == 0:
        print('SHARDED_OK world', world, flush=True)
    torch.cuda.synchronize()
    os._exit(0)          # symmetric-memory / NCCL teardown can block at interpreter exit


if __name__ == '__main__':
    main()
