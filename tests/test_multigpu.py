"""Multi-process tests: (1) CPU, gloo, world_size 2 — host-side sharding math and the dense-gradient bucket;
(2) GPU, >= 2 devices — row-sharded peer-memory DeepFM step against the single-GPU step (torchrun subprocess)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from rec_pangu_b200 import dist as rdist
    import oracle
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # dense bucket: sum of per-rank grads of loss/world == grad of the global-mean loss
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        x = torch.randn(8 * world, 5)
        y = torch.randn(8 * world, 3)
        mine = slice(rank * 8, (rank + 1) * 8)
        loss = ((lin(x[mine]) - y[mine]) ** 2).mean()
        (loss / world).backward()
        bucket = rdist.DenseGradBucket(list(lin.parameters()))
        bucket.all_reduce()
        lin2 = torch.nn.Linear(5, 3)
        lin2.load_state_dict(lin.state_dict())
        ((lin2(x) - y) ** 2).mean().backward()
        for p, p2 in zip(lin.parameters(), lin2.parameters()):
            torch.testing.assert_close(p.grad, p2.grad, rtol=1e-5, atol=1e-6)
        # sharding math == numpy restatement (oracle/index_routing.py), round trip through all_gather
        rows, D = 103, 4
        full = torch.arange(rows * D, dtype=torch.float32).view(rows, D)
        sl = rdist.local_slice(full, rank, world)
        assert sl.shape[0] == rdist.shard_rows(rows, world)
        ids = np.arange(rows)
        owner, local = oracle.shard_route(ids, world)
        mine_ids = ids[owner == rank]
        assert np.array_equal(local[owner == rank], np.arange(len(mine_ids)))
        assert torch.equal(sl[:len(mine_ids)], full[torch.from_numpy(mine_ids)])
        shards = [torch.empty_like(sl) for _ in range(world)]
        dist.all_gather(shards, sl)
        assert torch.equal(rdist.unshard(shards, rows), full)
        q.put((rank, 'ok'))
    except Exception as e:  # noqa
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_dense_bucket_and_shard_math():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == 'ok' for r in res), res


@pytest.mark.gpu
def test_row_sharded_peer_memory_step_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs (run with gpurun --gpus 2)')
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1',
           '--master-port', '29611', os.path.join(ROOT, 'tests', 'mp_sharded_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'SHARDED_OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
