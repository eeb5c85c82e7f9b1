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


def check(rank, world, dev, hidden):
    """hidden = [16, 8]: layer-by-layer MLP kernels; [64, 64]: the fused tower-tail kernels (single-GPU reference with the
    loss fused into the DeepFM core node, sharded model through MLP + the separate sigmoid/BCE head)."""
    enc = make_enc(6, 3, [101, 57, 33, 200, 17, 64])
    B = 96
    torch.manual_seed(7)
    ref = DeepFM(embedding_dim=8, hidden_units=hidden, enc_dict=enc)
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if 'embedding_layer' in n:
                p.mul_(0.3)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    ref = ref.to(dev)
    model = DeepFM(embedding_dim=8, hidden_units=hidden, enc_dict=enc)
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
        out = model(mine)
        (out['loss'] / world).backward()
        bucket.all_reduce()
        ops.check_index_errors(dev)
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


def main():
    rank, world, local = rdist.init_from_env('nccl')
    dev = torch.device('cuda', local)
    for hidden in ([16, 8], [64, 64]):
        check(rank, world, dev, hidden)
    if rank == 0:
        print('SHARDED_OK world', world, flush=True)
    torch.cuda.synchronize()
    os._exit(0)          # symmetric-memory / NCCL teardown can block at interpreter exit


if __name__ == '__main__':
    main()
