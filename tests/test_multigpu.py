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


class _HostShards:
    """dist.ShardedTables without the device part (CPU tensors, gloo collectives): what the checkpoint helpers touch."""

    def __init__(self, emb, rank, world, rdist):
        self.rank, self.world, self.group = rank, world, None
        self.cols, self.D = list(emb.emb_feature), emb.embedding_dim
        self.rows = [int(emb.enc_dict[c]['vocab_size']) + 1 for c in self.cols]
        self.weights = [rdist.local_slice(emb.embedding_layer[c].weight.data, rank, world) for c in self.cols]
        self._rdist = rdist

    def full_table(self, f):
        shards = [torch.empty_like(self.weights[f]) for _ in range(self.world)]
        dist.all_gather(shards, self.weights[f])
        return self._rdist.unshard(shards, self.rows[f])

    def barrier(self):
        dist.barrier()


def _check_sharded_checkpoint(rank, world, rdist):
    """gather_state_dict / load_state_dict_sharded round-trip a reference-layout state_dict through row shards."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from helpers import make_enc
    from rec_pangu_b200.models.ranking import DeepFM
    enc = make_enc(4, 2, [101, 57, 33, 8])
    torch.manual_seed(5)
    model = DeepFM(embedding_dim=8, hidden_units=[16, 8], enc_dict=enc)
    full_sd = {k: v.clone() for k, v in model.state_dict().items()}
    model.embedding_layer.attach_shards(_HostShards(model.embedding_layer, rank, world, rdist))
    assert model.state_dict()['embedding_layer.embedding_layer.C1.weight'].shape[0] == rdist.shard_rows(102, world)
    got = rdist.gather_state_dict(model)
    assert set(got) == set(full_sd)
    for k, v in full_sd.items():
        assert torch.equal(got[k], v), k
    torch.manual_seed(6)
    other = {k: torch.randn_like(v) for k, v in full_sd.items()}
    rdist.load_state_dict_sharded(model, other)
    st = model.embedding_layer._shards
    for f, c in enumerate(st.cols):
        assert torch.equal(st.weights[f], rdist.local_slice(other[f'embedding_layer.embedding_layer.{c}.weight'], rank, world))
        assert model.embedding_layer.embedding_layer[c].weight.data_ptr() == st.weights[f].data_ptr()     # loaded in place
    assert torch.equal(model.dnn.net[0].weight, other['dnn.net.0.weight'])
    back = rdist.gather_state_dict(model)
    for k, v in other.items():
        assert torch.equal(back[k], v), k
    bad = dict(other)
    bad.pop('dnn.net.0.bias')
    try:
        rdist.load_state_dict_sharded(model, bad)
        raise AssertionError('strict load accepted a missing key')
    except RuntimeError:
        pass


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
        _check_sharded_checkpoint(rank, world, rdist)
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
