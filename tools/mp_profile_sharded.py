"""torchrun diagnostic: per-phase device time of the row-sharded (NVLink peer memory) gather / scatter / re-zero at
config-2 shape, max over ranks."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench
from rec_pangu_b200 import dist as rdist, ops
from rec_pangu_b200.models.ranking import DeepFM


def main():
    rank, world, local = rdist.init_from_env('nccl')
    dev = torch.device('cuda', local)
    enc = bench.make_enc()
    B, D = bench.CFG['B'], bench.CFG['D']
    torch.manual_seed(1029)
    with torch.device(dev):
        model = DeepFM(embedding_dim=D, hidden_units=bench.CFG['hidden'], enc_dict=enc)
    st = rdist.shard_model_tables(model)
    torch.cuda.empty_cache()
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    data = bench.synth_batch(enc, B, gen, device=dev)
    emb = model.embedding_layer
    idx = [data[c] for c in emb.emb_feature]
    dn = [data[c] for c in emb.dense_feature]

    def timed(fn, iters=10):
        for _ in range(2):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(4e7))
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) * 1e3 / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {}
    with torch.no_grad():
        res['gather_fwd_sharded_us'] = timed(lambda: ops.gather_sharded(st, emb.tables(), idx, dn, want_fm=True))
    res['barrier_us'] = timed(lambda: st.barrier())

    def fwd_bwd():
        x, fm, _ = ops.gather_sharded(st, emb.tables(), idx, dn, want_fm=True)
        (x.sum() * 1e-6 + fm.sum() * 1e-6).backward()
    res['gather_fwd+bwd(scatter+barrier)_us'] = timed(fwd_bwd, iters=5)
    res['clean(rows_zero+barrier)_us'] = timed(lambda: (st.pending.append(idx), ops.sharded_clean(st)), iters=5)
    if rank == 0:
        print('SHARDED_PROFILE', world, res, flush=True)
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == '__main__':
    main()
