"""Multi-GPU execution of the hot path: one process per GPU (torchrun), NCCL for the plumbing collectives,
row-sharded embedding tables in NVLink peer memory for the lookup.

Design.  The reference is single-device; the path shards on the batch axis and on
table rows:

* batch: every rank runs the model on its own B samples (weak scaling); dense (MLP / interaction) gradients are
  averaged with one NCCL all-reduce over a flat bucket;
* tables: each table is ROW-SHARDED over the G GPUs (owner = id mod G, local row = id div G — the integer contract of
  oracle/index_routing.py).  Shards live in symmetric memory (torch.distributed._symmetric_memory: CUDA VMM + NVLink
  peer mappings), so the gather kernel reads a remote row with a plain load over NVLink and the scatter kernel adds a
  remote gradient row with a vector reduction over NVLink: the index/row all-to-all of a classical sharded lookup is
  FUSED into the gather/scatter kernels — no bucketing, no host-visible counts, no separate collective.  NVSwitch gives
  every peer full bandwidth, so the uniform `mod G` placement is also the load-balanced one.
* the only synchronisation is a device-side barrier (NCCL all-reduce of one word on the compute stream) after backward
  (all remote gradient adds have landed) and after zero_grad / optimizer (tables and grad shards are consistent again).
"""
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> (int, int, int):
    """(rank, world, local_rank) from torchrun's environment; initialises the default process group if needed."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_rows(rows: int, world: int) -> int:
    """Rows per shard (identical on every rank; the tail of the last shards is padding)."""
    return (rows + world - 1) // world


def local_slice(full: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rows owned by `rank` under owner = id mod G: full[rank::G], zero-padded to shard_rows."""
    sl = full[rank::world]
    n = shard_rows(full.shape[0], world)
    if sl.shape[0] < n:
        sl = torch.cat([sl, sl.new_zeros((n - sl.shape[0],) + tuple(sl.shape[1:]))], dim=0)
    return sl.contiguous()


def unshard(shards: List[torch.Tensor], rows: int) -> torch.Tensor:
    """Inverse of local_slice over all ranks' shards -> the full [rows, D] table (for state_dict round trips)."""
    world = len(shards)
    out = shards[0].new_empty((rows,) + tuple(shards[0].shape[1:]))
    for r, s in enumerate(shards):
        n = len(range(r, rows, world))
        out[r::world] = s[:n]
    return out


class DenseGradBucket:
    """Flat bucket for the non-embedding parameters: copy grads in, one all-reduce(sum), copy back.

    Convention for data-parallel steps: every rank back-propagates ``loss / world`` (its loss is the mean over its own
    batch), so that summed gradients — both these dense ones and the table gradients that the scatter kernel adds
    straight into the owners' shards — equal the gradient of the mean loss over the global batch."""

    def __init__(self, params: List[torch.nn.Parameter], group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else 'cpu'
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)

    def all_reduce(self, average: bool = False):
        """Three launches around the collective whatever the number of parameters: pack (one multi-tensor copy), all-reduce,
        unpack (one multi-tensor copy) — a per-parameter loop costs two tiny kernels per tensor (18 for DeepFM's tower)."""
        if self.world == 1 or not self.params:
            return
        views = list(self.flat.split([p.numel() for p in self.params]))
        live = [(v, p) for v, p in zip(views, self.params) if p.grad is not None]
        dead = [v for v, p in zip(views, self.params) if p.grad is None]
        if dead:
            torch._foreach_zero_(dead)
        if live:
            torch._foreach_copy_([v for v, _ in live], [p.grad.reshape(-1) for _, p in live])
        dist.all_reduce(self.flat, group=self.group)
        if average:
            self.flat.mul_(1.0 / self.world)
        if live:
            torch._foreach_copy_([p.grad.reshape(-1) if p.grad.is_contiguous() else p.grad for _, p in live],
                                 [v if p.grad.is_contiguous() else v.view_as(p.grad) for v, p in live])


class ShardedTables:
    """Row-sharded embedding tables + gradient shards of one EmbeddingLayer in symmetric (peer-mapped) memory."""

    def __init__(self, emb_layer, group=None, full_tables: Optional[Dict[str, torch.Tensor]] = None, init=None, seed: int = 1029):
        """`init`: None = slice the full tables the layer (or `full_tables`) holds — every rank must hold identical ones;
        'kaiming' | 'xavier' | callable(shard, f, rank) = fill each rank's shard LOCALLY, no full table anywhere (tables that
        exceed one GPU: BASELINE.json config 5).  'kaiming' is the reference's reset_parameters (base_model.py:42-59:
        kaiming_normal_ on [rows, D] => std = sqrt(2 / D)), 'xavier' its _init_weights (std = sqrt(2 / (rows + D)))."""
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.cols = list(emb_layer.emb_feature)
        self.D = emb_layer.embedding_dim
        dev = torch.device('cuda', torch.cuda.current_device())
        self.rows = [int(emb_layer.enc_dict[c]['vocab_size']) + 1 for c in self.cols]
        F, G = len(self.cols), self.world
        self.weights, self.grads, self._handles = [], [], []
        w_ptrs = torch.zeros(F * G, dtype=torch.int64)
        g_ptrs = torch.zeros(F * G, dtype=torch.int64)
        # all gradient shards of the layer live in ONE symmetric allocation: the per-step re-zero is one memset instead of F
        # fill launches (26 x ~2.3 us at config 2), one rendezvous instead of F
        ns = [shard_rows(r, G) for r in self.rows]
        seg = [(n * self.D + 3) // 4 * 4 for n in ns]             # floats per table, 16-byte aligned segments (D = 1 tables)
        offs = [sum(seg[:f]) for f in range(F)]
        self.grad_flat = symm.empty((sum(seg),), dtype=torch.float32, device=dev)
        hgf = symm.rendezvous(self.grad_flat, self.group)
        self.grad_flat.zero_()
        self._handles.append(hgf)
        for f, c in enumerate(self.cols):
            n = ns[f]
            w = symm.empty((n, self.D), dtype=torch.float32, device=dev)
            g = self.grad_flat[offs[f]:offs[f] + n * self.D].view(n, self.D)
            hw = symm.rendezvous(w, self.group)
            if init is None:
                src = full_tables[c] if full_tables is not None else emb_layer.embedding_layer[c].weight.data
                if src.device.type == 'meta':
                    raise RuntimeError('tables were created under dist.deferred_tables(): pass init= to shard_model_tables')
                w.copy_(local_slice(src.to(dev), self.rank, G))
            elif callable(init):
                init(w, f, self.rank)
            else:
                gen = torch.Generator(device=dev).manual_seed(seed + 7919 * f + 104729 * self.rank)
                std = (2.0 / self.D) ** 0.5 if init == 'kaiming' else (2.0 / (self.rows[f] + self.D)) ** 0.5
                w.normal_(0.0, std, generator=gen)
                n_mine = len(range(self.rank, self.rows[f], G))
                if n_mine < n:
                    w[n_mine:].zero_()                       # padding rows of the last shards
            for r in range(G):
                w_ptrs[f * G + r] = int(hw.buffer_ptrs[r])
                g_ptrs[f * G + r] = int(hgf.buffer_ptrs[r]) + offs[f] * 4
            self.weights.append(w)
            self.grads.append(g)
            self._handles.append(hw)
        self.w_tab = w_ptrs.to(dev)
        self.g_tab = g_ptrs.to(dev)
        self._one = torch.zeros(1, device=dev)
        self.pending = []            # idx lists of backward passes whose touched grad rows still need re-zeroing
        self.barrier()

    def barrier(self):
        """Device-side barrier on the current stream (graph-capturable)."""
        dist.all_reduce(self._one, group=self.group)

    def full_table(self, f: int) -> torch.Tensor:
        """All-gather the shards of table f and re-interleave them (checkpointing / tests)."""
        shards = [torch.empty_like(self.weights[f]) for _ in range(self.world)]
        dist.all_gather(shards, self.weights[f], group=self.group)
        return unshard(shards, self.rows[f])

    def full_grad(self, f: int) -> torch.Tensor:
        shards = [torch.empty_like(self.grads[f]) for _ in range(self.world)]
        dist.all_gather(shards, self.grads[f], group=self.group)
        return unshard(shards, self.rows[f])


class LocalShards:
    """All G row shards of every table on ONE device, with the pointer tables the sharded kernels take — a single-GPU
    stand-in for ShardedTables (no symmetric memory, no collectives; `barrier()` is a no-op) used to check the sharded
    address arithmetic of the kernels (owner = id mod G, local row = id div G) without a second GPU: a shard pointer is a
    shard pointer, local or NVLink.  Acts as rank 0: `weights` / `grads` are rank 0's shards (the autograd anchors and the
    published `.grad`), `all_weights[f][g]` / `all_grads[f][g]` hold every shard."""

    def __init__(self, emb_layer, world: int):
        self.world, self.rank, self.group = int(world), 0, None
        self.cols = list(emb_layer.emb_feature)
        self.D = emb_layer.embedding_dim
        self.rows = [int(emb_layer.enc_dict[c]['vocab_size']) + 1 for c in self.cols]
        F, G = len(self.cols), self.world
        self.all_weights, self.all_grads = [], []
        w_ptrs = torch.zeros(F * G, dtype=torch.int64)
        g_ptrs = torch.zeros(F * G, dtype=torch.int64)
        dev = None
        for f, c in enumerate(self.cols):
            full = emb_layer.embedding_layer[c].weight.data
            dev = full.device
            ws = [local_slice(full, g, G) for g in range(G)]
            gs = [torch.zeros_like(w) for w in ws]
            for g in range(G):
                w_ptrs[f * G + g] = ws[g].data_ptr()
                g_ptrs[f * G + g] = gs[g].data_ptr()
            self.all_weights.append(ws)
            self.all_grads.append(gs)
        self.weights = [ws[0] for ws in self.all_weights]
        self.grads = [gs[0] for gs in self.all_grads]
        self.w_tab, self.g_tab = w_ptrs.to(dev), g_ptrs.to(dev)
        self.pending = []

    def barrier(self):
        pass

    def full_table(self, f: int) -> torch.Tensor:
        return unshard(self.all_weights[f], self.rows[f])

    def full_grad(self, f: int) -> torch.Tensor:
        return unshard(self.all_grads[f], self.rows[f])

    def zero_grads(self):
        for gs in self.all_grads:
            for g in gs:
                g.zero_()
        self.pending = []


def _sharded_layers(model):
    """(state_dict prefix, EmbeddingLayer) for every row-sharded EmbeddingLayer of the model — `embedding_layer` and the D = 1
    tables of an LR_Layer (`lr_layer.emb_layer` / `lr.emb_layer`); keys are `<prefix>.embedding_layer.<col>.weight` (App. C)."""
    from .models.layers.embedding import EmbeddingLayer
    return [(name, m) for name, m in model.named_modules() if isinstance(m, EmbeddingLayer) and m._shards is not None]


def _table_key(col: str, prefix: str = 'embedding_layer') -> str:
    return f'{prefix}.embedding_layer.{col}.weight'


def gather_state_dict(model) -> Dict[str, torch.Tensor]:
    """`model.state_dict()` in the REFERENCE's layout from a model whose tables are row-sharded: every table entry
    (`embedding_layer.embedding_layer.<col>.weight`, and the LR tables) is all-gathered and re-interleaved to
    [vocab_size + 1, D]; all other entries (replicated on every rank) are the local ones.  Collective: every rank of the
    shard group must call it and every rank gets the full dict, so `RankTrainer.save_model` on rank 0 writes a checkpoint
    that the reference — or a single-GPU run of this build, or a run sharded over another number of GPUs — loads unchanged
    (trainer.py:133-150)."""
    sd = {k: v for k, v in model.state_dict().items()}
    for prefix, m in _sharded_layers(model):
        st = m._shards
        for f, c in enumerate(st.cols):
            sd[_table_key(c, prefix)] = st.full_table(f)
    return sd


def load_state_dict_sharded(model, state_dict: Dict[str, torch.Tensor], strict: bool = True):
    """Inverse of gather_state_dict: load a reference-layout state_dict into a model with row-sharded tables.  Each rank
    copies ITS rows (owner = id mod G) of every full table into its shard in place (the peer mappings stay valid);
    everything else goes through nn.Module.load_state_dict.  Full-size table entries never touch the device whole."""
    layers = _sharded_layers(model)
    if not layers:
        return model.load_state_dict(state_dict, strict=strict)
    rest = dict(state_dict)
    table_keys = set()
    for prefix, m in layers:
        st = m._shards
        for f, c in enumerate(st.cols):
            key = _table_key(c, prefix)
            table_keys.add(key)
            if key not in rest:
                if strict:
                    raise KeyError(f'missing key in state_dict: {key}')
                continue
            full = rest.pop(key)
            if tuple(full.shape) != (st.rows[f], st.D):
                raise RuntimeError(f'size mismatch for {key}: checkpoint {tuple(full.shape)}, model {(st.rows[f], st.D)}')
            with torch.no_grad():
                st.weights[f].copy_(local_slice(full, st.rank, st.world).to(st.weights[f].device))
    own = {k: v for k, v in model.state_dict().items() if k not in table_keys}
    missing = [k for k in own if k not in rest]
    unexpected = [k for k in rest if k not in own]
    if strict and (missing or unexpected):
        raise RuntimeError(f'load_state_dict_sharded: missing keys {missing}, unexpected keys {unexpected}')
    res = model.load_state_dict(rest, strict=False)
    for _, m in layers:
        m._shards.barrier()
    return res


def enable_sync_batchnorm(group=None, enabled: bool = True):
    """BatchNorm1d layers of the hot path (multi-task towers) normalise with the statistics of the GLOBAL batch: the
    [sum | sum of squares | count] vector in forward and the [dgamma | dbeta] column sums in backward are all-reduced over
    `group` (needed for parity with the single-process reference under data parallelism)."""
    from . import ops
    ops.SYNC_BN_GROUP = (group if group is not None else dist.group.WORLD) if enabled else None


class deferred_tables:
    """Context manager: EmbeddingLayers constructed inside create their nn.Embedding containers on the `meta` device (shape
    only, no storage), so a model whose tables exceed one GPU can be built and then given row shards that are filled
    locally (`shard_model_tables(model, init=...)`).  reset_parameters / _init_weights skip meta tensors."""
    active = False

    def __enter__(self):
        self._prev, deferred_tables.active = deferred_tables.active, True
        return self

    def __exit__(self, *exc):
        deferred_tables.active = self._prev
        return False


def shard_model_tables(model, group=None, init=None, seed: int = 1029) -> ShardedTables:
    """Convert every EmbeddingLayer of the model — `model.embedding_layer` and the D = 1 tables of an LR_Layer
    (models/layers/shallow.py) — to row-sharded peer-memory tables.  init=None: every rank must hold identical full
    tables when this is called (same seed); they are released afterwards.  init='kaiming' | 'xavier' | callable: see
    ShardedTables.  Returns the ShardedTables of `model.embedding_layer`."""
    from .models.layers.embedding import EmbeddingLayer
    main = None
    for i, m in enumerate(mod for mod in model.modules() if isinstance(mod, EmbeddingLayer)):
        st = ShardedTables(m, group, init=init, seed=seed + 1000003 * i)
        m.attach_shards(st)
        if m is model.embedding_layer:
            main = st
    return main
