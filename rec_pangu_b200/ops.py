"""Autograd bindings of the sm_100a kernels (C ABI in include/rec_pangu_b200.h, loaded through ctypes).

PyTorch is plumbing here: it owns device memory, the current stream and the autograd tape; every
numerical step of the hot path is a kernel of librec_pangu_b200.so.  CUDA tensors only — there is no CPU
fallback, and a missing library raises at first use.
"""
import ctypes as C
import os
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import GatherDesc, ScatterDesc, TowerFwdDesc, TowerBwdDesc, check

__all__ = ['GradStore', 'hash_to_row', 'essm_head', 'deepfm_core', 'gather', 'gather_sharded', 'sharded_clean', 'fm_interaction', 'mlp_forward', 'sigmoid_bce', 'linear', 'feature_row_stride',
           'check_index_errors', 'set_gemm_impl', 'get_gemm_impl', 'launch_count', 'reset_launch_count']

_GEMM_IMPL = int(os.environ.get('RPB_GEMM_IMPL', '0'))   # 0 auto, 1 SIMT fp32, 2 tcgen05 3xTF32
# DeepFM backward: 1 = layer-1 dx GEMM scatters table gradients from its epilogue (rpb_linear_dx_scatter),
# 0 = dx GEMM to HBM followed by rpb_gather_bwd.  Both are bit-for-bit the same sums in a different add order.
FUSED_DX_SCATTER = int(os.environ.get('RPB_DX_SCATTER', '1'))
# MLP tower tail (rpb_tower_tail_fwd/bwd): 1 = every 64-wide hidden layer after the first, the Linear(64->1) output, the
# logit sum and (DeepFM) sigmoid + BCE run as ONE fp32 kernel per direction; 0 = one GEMM / row-dot / head kernel each.
TOWER_TAIL = int(os.environ.get('RPB_TOWER_TAIL', '1'))
# 1 = the tail runs inside the epilogue of the layer-1 tcgen05 GEMM (rpb_linear_tower_fwd), 0 = as its own kernel
FUSED_TOWER_EPILOGUE = int(os.environ.get('RPB_TOWER_EPILOGUE', '1'))
# 1 = the tower's small weight-gradient kernels run on a side stream concurrently with the layer-1 one
PARALLEL_WGRAD = int(os.environ.get('RPB_PARALLEL_WGRAD', '1'))
_SIDE = {}


def _side_stream(dev) -> 'torch.cuda.Stream':
    key = torch.device(dev).index or 0
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


# 1 = DeepFM forward as ONE kernel: the layer-1 GEMM gathers its own operand rows (rpb_deepfm_fwd_fused)
FUSED_GATHER_GEMM = int(os.environ.get('RPB_FUSED_GATHER_GEMM', '1'))
# Row-sharded tables (dist.ShardedTables): 1 = DeepFM runs its fused core on them too — the one-kernel forward requests
# remote rows with the same cp.async over NVLink and the dx GEMM's scatter epilogue reduces into the owners' gradient
# shards — instead of the separate gather / MLP / dx GEMM / scatter kernels.  Opt-in until measured on >= 2 GPUs.
SHARDED_FUSED = int(os.environ.get('RPB_SHARDED_FUSED', '1'))
_LAUNCHES = 0           # number of librec_pangu_b200 kernels launched (bench.py reports it as gpu_launches)


def set_gemm_impl(impl: int):
    global _GEMM_IMPL
    assert impl in (0, 1, 2)
    _GEMM_IMPL = impl


def get_gemm_impl() -> int:
    return _GEMM_IMPL


def launch_count() -> int:
    return _LAUNCHES


def reset_launch_count():
    global _LAUNCHES
    _LAUNCHES = 0


def _count(n=1):
    global _LAUNCHES
    _LAUNCHES += n


def _cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f'rec_pangu_b200: {what} must be a CUDA tensor — the hot path is hand-written sm_100a '
                           f'CUDA and has no CPU fallback (got device {t.device})')


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _rowmajor(t: torch.Tensor) -> torch.Tensor:
    """2-D tensor whose last dim is contiguous (row stride arbitrary)."""
    if t.dim() != 2:
        t = t.reshape(t.shape[0], -1)
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def feature_row_stride(F: int, D: int, Nd: int) -> int:
    """Row stride (floats) of the feature-row buffer x: F*D+Nd rounded up to a multiple of 4 (16-byte rows)."""
    return (F * D + Nd + 3) // 4 * 4


# ------------------------------------------------------------------ index error record (pinned, device visible)
_ERR = {}


def _err_record(device) -> torch.Tensor:
    key = torch.device(device).index or 0
    if key not in _ERR:
        _ERR[key] = torch.zeros(4, dtype=torch.int64).pin_memory()
    return _ERR[key]


_DROPOUT_EPOCH = {}
_dropout_calls = 0


def _dropout_epoch(device) -> torch.Tensor:
    """Per-device int64 step counter the dropout kernels mix into their seed when they RUN.  The host-drawn seed of a call is
    baked into a captured CUDA graph; `advance_dropout_epoch()` (inside the captured step: runtime.GraphedStep) makes every
    replay draw a fresh mask, forward and backward of one step seeing the same value."""
    global _dropout_calls
    _dropout_calls += 1
    key = torch.device(device).index or 0
    t = _DROPOUT_EPOCH.get(key)
    if t is None:
        t = _DROPOUT_EPOCH[key] = torch.zeros(1, dtype=torch.int64, device=device)
    return t


def dropout_calls() -> int:
    """Number of dropout kernel launches so far (GraphedStep: does the step need the epoch increment?)."""
    return _dropout_calls


def advance_dropout_epoch(device=None):
    dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
    key = dev.index or 0
    t = _DROPOUT_EPOCH.get(key)
    if t is None:
        t = _DROPOUT_EPOCH[key] = torch.zeros(1, dtype=torch.int64, device=dev)
    t.add_(1)


def check_index_errors(device=None, sync: bool = True):
    """Raise IndexError if any gather since the last check saw an index outside [0, vocab_size]
    (the reference raises IndexError from aten::embedding on CPU, models/layers/embedding.py:61-62)."""
    dev = torch.device(device if device is not None else torch.cuda.current_device())
    if sync:
        torch.cuda.synchronize(dev)
    rec = _err_record(dev)
    if int(rec[0]) != 0:
        f, b, v = int(rec[1]), int(rec[2]), int(rec[3])
        rec.zero_()
        raise IndexError(f'index out of range in embedding gather: field #{f}, sample {b}, index {v}')


# ------------------------------------------------------------------ persistent dense-grad buffers
class GradStore:
    """Persistent [rows, D] gradient buffers for embedding tables ('persistent' grad mode).

    The reference's backward materialises a zero-filled dense grad per table every step (embedding_dense_backward,
    SURVEY.md K14: 1.66 GB of zero-fill at config 2).  Here each table owns one buffer that is all-zero between
    steps: backward scatter-adds into it and publishes it as ``param.grad`` (a dense tensor with exactly the
    reference's content, so torch.optim.Adam(model.parameters()) still works), and ``clean()`` — called from
    ``BaseModel.zero_grad()`` or lazily by the next backward — re-zeroes only the rows the batch touched."""

    def __init__(self):
        self.buffers = {}          # id(param) -> tensor
        self.pending = []          # list of (grads, lr_grads, rows, idx tensors, D)

    def buffer(self, param: torch.Tensor) -> torch.Tensor:
        b = self.buffers.get(id(param))
        if b is None or b.shape != param.shape or b.device != param.device:
            b = torch.zeros_like(param)
            self.buffers[id(param)] = b
        return b

    def clean(self):
        for grads, lr_grads, rows, idx, D in self.pending:
            F = len(idx)
            d = ScatterDesc()
            d.B, d.F, d.D = idx[0].shape[0], F, D
            g_arr = _ptr_list(grads)
            d.grads = g_arr
            if lr_grads is not None:
                l_arr = _ptr_list(lr_grads)
                d.lr_grads = l_arr
            r_arr = (C.c_int64 * F)(*rows)
            i_arr = _ptr_list(idx)
            d.rows, d.idx = r_arr, i_arr
            check(_lib.load().rpb_rows_zero(C.byref(d), _stream()), 'rpb_rows_zero')
            _count()
        self.pending = []


def _ptr_list(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() if t is not None else 0 for t in ts])


# ------------------------------------------------------------------ gather
class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, *tensors):
        F, Nd, D = cfg['F'], cfg['Nd'], cfg['D']
        has_lr, want_fm = cfg['has_lr'], cfg['want_fm']
        tables = tensors[:F]
        lr_tables = tensors[F:2 * F] if has_lr else ()
        o = 2 * F if has_lr else F
        idx = tensors[o:o + F]
        dense = tensors[o + F:o + F + Nd]
        dev = tables[0].device
        B = idx[0].shape[0]
        ldx = feature_row_stride(F, D, Nd)
        x = torch.empty((B, ldx), dtype=torch.float32, device=dev) if cfg.get('want_x', True) else None
        need_grad = cfg['needs_grad']
        fm = torch.empty((B,), dtype=torch.float32, device=dev) if want_fm else None
        fm_s = torch.empty((B, D), dtype=torch.float32, device=dev) if (want_fm and need_grad) else None
        ld_lr = (F + Nd + 3) // 4 * 4 if has_lr else 0
        lr_in = torch.empty((B, ld_lr), dtype=torch.float32, device=dev) if has_lr else None
        rows = [int(t.shape[0]) for t in tables]

        d = GatherDesc()
        d.B, d.F, d.D, d.Nd, d.ldx, d.ld_lr = B, F, D, Nd, ldx, ld_lr
        t_arr = (C.c_void_p * F)(*[t.data_ptr() for t in tables])
        r_arr = (C.c_int64 * F)(*rows)
        i_arr = (C.c_void_p * F)(*[t.data_ptr() for t in idx])
        d.tables, d.rows, d.idx = t_arr, r_arr, i_arr
        if Nd:
            d_arr = (C.c_void_p * Nd)(*[t.data_ptr() for t in dense])
            d.dense = d_arr
        if has_lr:
            l_arr = (C.c_void_p * F)(*[t.data_ptr() for t in lr_tables])
            d.lr_tables = l_arr
        d.x, d.fm, d.fm_s, d.lr_in = _ptr(x), _ptr(fm), _ptr(fm_s), _ptr(lr_in)
        d.err = _err_record(dev).data_ptr()
        check(_lib.load().rpb_gather_fwd(C.byref(d), _stream()), 'rpb_gather_fwd')
        _count()

        ctx.set_materialize_grads(False)
        ctx.cfg = cfg
        ctx.rows = rows
        ctx.n_in = len(tensors)
        ctx.params = (tables, lr_tables)       # Parameter objects (persistent grad mode publishes .grad itself)
        ctx.save_for_backward(x, fm_s, *idx)
        outs = [x if x is not None else torch.empty(0, device=dev)]
        if want_fm:
            outs.append(fm)
        if has_lr:
            outs.append(lr_in)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        cfg = ctx.cfg
        F, Nd, D = cfg['F'], cfg['Nd'], cfg['D']
        has_lr, want_fm = cfg['has_lr'], cfg['want_fm']
        x, fm_s, *idx = ctx.saved_tensors
        gx = gouts[0]
        k = 1
        gfm = None
        glr = None
        if want_fm:
            gfm = gouts[k]
            k += 1
        if has_lr:
            glr = gouts[k]
        B = x.shape[0]
        dev = x.device
        tbl_req = ctx.needs_input_grad[1:1 + F]
        lr_req = ctx.needs_input_grad[1 + F:1 + 2 * F] if has_lr else ()

        grads: List[Optional[torch.Tensor]] = [None] * ctx.n_in
        store = cfg.get('grad_store')
        tables, lr_tables = ctx.params
        if store is not None:
            trainable = [tables[f] for f in range(F) if tbl_req[f]] + \
                        ([lr_tables[f] for f in range(F) if lr_req[f]] if has_lr else [])
            if store.pending and all(t.grad is None for t in trainable):
                store.clean()                  # grads were dropped (optimizer.zero_grad): start from all-zero buffers
            g_tables = [store.buffer(tables[f]) if tbl_req[f] else None for f in range(F)]
            g_lr = [store.buffer(lr_tables[f]) if (has_lr and lr_req[f]) else None for f in range(F)]
        else:
            g_tables = [torch.zeros((ctx.rows[f], D), dtype=torch.float32, device=dev) if tbl_req[f] else None
                        for f in range(F)]
            g_lr = [torch.zeros((ctx.rows[f], 1), dtype=torch.float32, device=dev) if (has_lr and lr_req[f]) else None
                    for f in range(F)]
        have_any = (gx is not None or gfm is not None) and any(g is not None for g in g_tables)
        have_lr = glr is not None and any(g is not None for g in g_lr)
        if have_any or have_lr:
            d = ScatterDesc()
            d.B, d.F, d.D = B, F, D
            if gx is not None:
                gx = _rowmajor(gx)
                d.dx, d.lddx = gx.data_ptr(), gx.stride(0)
            if gfm is not None:
                gfm = gfm.contiguous()
                d.dfm, d.x, d.ldx, d.fm_s = gfm.data_ptr(), x.data_ptr(), x.stride(0), fm_s.data_ptr()
            g_arr = (C.c_void_p * F)(*[g.data_ptr() if g is not None else 0 for g in g_tables])
            d.grads = g_arr
            if have_lr:
                glr = _rowmajor(glr)
                gl_arr = (C.c_void_p * F)(*[g.data_ptr() if g is not None else 0 for g in g_lr])
                d.lr_grads, d.dlr_in, d.ld_dlr = gl_arr, glr.data_ptr(), glr.stride(0)
            r_arr = (C.c_int64 * F)(*ctx.rows)
            i_arr = (C.c_void_p * F)(*[t.data_ptr() for t in idx])
            d.rows, d.idx = r_arr, i_arr
            check(_lib.load().rpb_gather_bwd(C.byref(d), _stream()), 'rpb_gather_bwd')
            _count()
        if store is not None:
            store.pending.append((g_tables, g_lr if has_lr else None, ctx.rows, list(idx), D))
            for f in range(F):
                for prm, buf in ((tables[f], g_tables[f]), (lr_tables[f] if has_lr else None, g_lr[f])):
                    if buf is None:
                        continue
                    if prm.grad is None:
                        prm.grad = buf             # publish the dense grad (same content as the reference's)
                    elif prm.grad is not buf:
                        raise RuntimeError("embedding table .grad was replaced externally; persistent grad mode "
                                           "needs model.zero_grad() / set grad_mode='dense'")
            return (None, *grads)
        for f in range(F):
            grads[f] = g_tables[f]
            if has_lr:
                grads[F + f] = g_lr[f]
        return (None, *grads)


def gather(tables: Sequence[torch.Tensor], idx: Sequence[torch.Tensor], dense: Sequence[torch.Tensor] = (),
           lr_tables: Optional[Sequence[torch.Tensor]] = None, want_fm: bool = False,
           grad_store: Optional['GradStore'] = None, want_x: bool = True):
    """One-launch multi-table gather.  Returns (x [B, ldx], fm [B] | None, lr_in [B, ld_lr] | None).

    ``x[:, :F*D].view(B, F, D)`` is the reference's ``EmbeddingLayer.forward`` output
    (models/layers/embedding.py:49-63); ``x[:, :F*D+Nd]`` is ``cat(emb.flatten(1), get_linear_input(...))``.
    """
    F = len(tables)
    if F == 0:
        raise ValueError('gather needs at least one sparse field')
    if F > _lib.MAX_FIELDS or len(dense) > _lib.MAX_DENSE:
        raise NotImplementedError(f'more than {_lib.MAX_FIELDS} sparse or {_lib.MAX_DENSE} dense fields per gather')
    D = int(tables[0].shape[1])
    for t in tables:
        _cuda(t, 'embedding table')
        if t.dtype != torch.float32 or not t.is_contiguous() or t.shape[1] != D:
            raise ValueError('embedding tables must be contiguous fp32 [rows, D] with one common D')
    idx_l = []
    for t in idx:
        _cuda(t, 'sparse feature column')
        t = t.reshape(-1)
        if t.dtype != torch.int64:          # embedding.py:61: X[col].long()
            t = t.long()
        idx_l.append(t.contiguous())
    B = idx_l[0].shape[0]
    dense_l = []
    for t in dense:
        _cuda(t, 'dense feature column')
        t = t.reshape(-1)
        if t.dtype != torch.float32:
            t = t.float()
        if t.shape[0] != B:
            raise ValueError('dense column batch size mismatch')
        dense_l.append(t.contiguous())
    has_lr = lr_tables is not None
    if has_lr:
        lr_tables = [t if t.is_contiguous() else t.contiguous() for t in lr_tables]
    needs_grad = torch.is_grad_enabled() and (any(t.requires_grad for t in tables) or
                                              (has_lr and any(t.requires_grad for t in lr_tables)))
    if not want_x and (needs_grad or not (want_fm or has_lr)):
        want_x = True                  # backward (and any consumer of the rows) needs the materialised feature row
    cfg = dict(F=F, Nd=len(dense_l), D=D, has_lr=has_lr, want_fm=want_fm, needs_grad=needs_grad,
               grad_store=grad_store, want_x=want_x)
    args = list(tables) + (list(lr_tables) if has_lr else []) + idx_l + dense_l
    outs = _Gather.apply(cfg, *args)
    x = outs[0]
    k = 1
    fm = None
    lr_in = None
    if want_fm:
        fm = outs[k]
        k += 1
    if has_lr:
        lr_in = outs[k]
    return x, fm, lr_in


class _GatherSharded(torch.autograd.Function):
    """Gather over row-sharded tables in NVLink peer memory (dist.ShardedTables).  Gradients are reduced straight into
    the owners' gradient shards by the scatter kernel; after the device barrier each local shard's .grad is published."""

    @staticmethod
    def forward(ctx, st, want_fm, n_idx, need_grad, *tensors):
        F, D, G = len(st.cols), st.D, st.world
        params = tensors[:F]                       # local shard Parameters (autograd anchors)
        idx = tensors[F:F + n_idx]
        dense = tensors[F + n_idx:]
        Nd = len(dense)
        dev = idx[0].device
        B = idx[0].shape[0]
        ldx = feature_row_stride(F, D, Nd)
        x = torch.empty((B, ldx), dtype=torch.float32, device=dev)
        fm = torch.empty((B,), dtype=torch.float32, device=dev) if want_fm else None
        fm_s = torch.empty((B, D), dtype=torch.float32, device=dev) if (want_fm and need_grad) else None
        d = GatherDesc()
        d.B, d.F, d.D, d.Nd, d.ldx, d.ld_lr = B, F, D, Nd, ldx, 0
        r_arr = (C.c_int64 * F)(*st.rows)
        i_arr = _ptr_list(idx)
        d.rows, d.idx = r_arr, i_arr
        if Nd:
            d_arr = _ptr_list(dense)
            d.dense = d_arr
        d.x, d.fm, d.fm_s = x.data_ptr(), _ptr(fm), _ptr(fm_s)
        d.err = _err_record(dev).data_ptr()
        d.G, d.shard_tab = G, st.w_tab.data_ptr()
        check(_lib.load().rpb_gather_fwd(C.byref(d), _stream()), 'rpb_gather_fwd(sharded)')
        _count()
        ctx.set_materialize_grads(False)
        ctx.st, ctx.want_fm, ctx.params = st, want_fm, params
        ctx.n_inputs = 4 + len(tensors)
        ctx.save_for_backward(x, fm_s, *idx)
        return (x, fm) if want_fm else (x,)

    @staticmethod
    def backward(ctx, *gouts):
        st = ctx.st
        x, fm_s, *idx = ctx.saved_tensors
        if st.pending and all(p.grad is None for p in ctx.params if p.requires_grad):
            sharded_clean(st)                         # grads were dropped without model.zero_grad() (optimizer.zero_grad()): lazy re-zero
        gx = gouts[0]
        gfm = gouts[1] if ctx.want_fm else None
        F, D, G = len(st.cols), st.D, st.world
        if gx is not None or gfm is not None:
            d = ScatterDesc()
            d.B, d.F, d.D = x.shape[0], F, D
            if gx is not None:
                gx = _rowmajor(gx)
                d.dx, d.lddx = gx.data_ptr(), gx.stride(0)
            if gfm is not None:
                gfm = gfm.contiguous()
                d.dfm, d.x, d.ldx, d.fm_s = gfm.data_ptr(), x.data_ptr(), x.stride(0), fm_s.data_ptr()
            r_arr = (C.c_int64 * F)(*st.rows)
            i_arr = _ptr_list(idx)
            d.rows, d.idx = r_arr, i_arr
            d.G, d.grad_shard_tab = G, st.g_tab.data_ptr()
            check(_lib.load().rpb_gather_bwd(C.byref(d), _stream()), 'rpb_gather_bwd(sharded)')
            _count()
        st.pending.append(list(idx))
        st.barrier()                                  # every rank's remote gradient adds have landed
        for p, g in zip(ctx.params, st.grads):
            if p.requires_grad:
                p.grad = g
        return (None,) * ctx.n_inputs


def gather_sharded(st, params, idx: Sequence[torch.Tensor], dense: Sequence[torch.Tensor] = (), want_fm: bool = False):
    """Row-sharded counterpart of `gather` (no LR tables).  Returns (x, fm | None, None)."""
    idx_l = []
    for t in idx:
        _cuda(t, 'sparse feature column')
        t = t.reshape(-1)
        if t.dtype != torch.int64:
            t = t.long()
        idx_l.append(t.contiguous())
    dense_l = [t.reshape(-1).float().contiguous() for t in dense]
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    outs = _GatherSharded.apply(st, want_fm, len(idx_l), need_grad, *params, *idx_l, *dense_l)
    return outs[0], (outs[1] if want_fm else None), None


def sharded_clean(st):
    """Re-zero the gradient shards after a step, then a barrier.  Two ways: (b) every rank memsets its OWN shards (local
    HBM at ~6 TB/s; 1/G of the tables: 128 us at G = 2, 32 us at G = 8 for config 2), or (a) every rank clears the rows ITS
    batches touched wherever they live (rpb_rows_zero with the shard table: 64-byte stores, (G-1)/G of them over NVLink) —
    the only affordable way when the shards are large (config 5: a memset of 2 GB per table and step).
    Ordering: in (a) a rank writes into OTHER ranks' gradient shards, which their optimizer may still be reading, so (a)
    starts with a barrier (every rank is past optimizer.step) — in (b) nobody touches foreign memory.  The choice is made
    once per ShardedTables from an all-reduced estimate, so every rank takes the same branch whatever its batch size."""
    F = len(st.cols)
    if not st.pending:
        return
    if st.world > 1:
        mode = getattr(st, '_clean_mode', None)
        if mode is None:
            n_rows = sum(ix[0].shape[0] * F for ix in st.pending)
            dense_us = sum(g.numel() for g in st.grads) * 4 / 6.0e6
            want_dense = 1.0 if dense_us < 8.0e-5 * n_rows else 0.0
            t = torch.tensor([want_dense], device=st.grads[0].device)
            if st.group is not None and torch.distributed.is_initialized():
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN, group=st.group)
            mode = st._clean_mode = 'dense' if float(t.item()) > 0.5 else 'sparse'
        if mode == 'dense':
            flat = getattr(st, 'grad_flat', None)
            if flat is not None:
                flat.zero_()                      # every gradient shard of the layer in one memset
            else:
                for g in st.grads:
                    g.zero_()
            st.barrier()
            st.pending = []
            return
        st.barrier()                      # nobody's optimizer is still reading the shards we are about to write into
    for idx in st.pending:
        d = ScatterDesc()
        d.B, d.F, d.D = idx[0].shape[0], F, st.D
        r_arr = (C.c_int64 * F)(*st.rows)
        i_arr = _ptr_list(idx)
        d.rows, d.idx = r_arr, i_arr
        d.G, d.grad_shard_tab = st.world, st.g_tab.data_ptr()
        check(_lib.load().rpb_rows_zero(C.byref(d), _stream()), 'rpb_rows_zero(sharded)')
        _count()
    st.barrier()
    st.pending = []


def hash_to_row(raw: torch.Tensor, vocab_size: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Hashed-id encoder (config 5): row = splitmix64(raw) mod vocab_size for an int64 CUDA tensor of raw ids.  The
    result addresses rows [0, vocab_size) of a [vocab_size + 1, D] table; row vocab_size stays the OOV slot."""
    _cuda(raw, 'raw id column')
    if raw.dtype != torch.int64:
        raise TypeError(f'hash_to_row expects int64 ids, got {raw.dtype}')
    raw = raw.contiguous()
    if out is None:
        out = torch.empty_like(raw)
    check(_lib.load().rpb_hash_to_row(_ptr(raw), _ptr(out), raw.numel(), int(vocab_size), _stream()), 'rpb_hash_to_row')
    _count()
    return out


# ------------------------------------------------------------------ standalone FM on [B,F,D]
class _FM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e, mode):
        B, F, D = e.shape
        if e.stride(2) != 1 or e.stride(1) != D:
            e = e.contiguous()
        out_sum = torch.empty((B, 1), dtype=torch.float32, device=e.device) if mode == 'sum' else None
        out_bi = torch.empty((B, D), dtype=torch.float32, device=e.device) if mode == 'bi' else None
        check(_lib.load().rpb_fm_fwd(_ptr(e), e.stride(0), B, F, D, _ptr(out_sum), _ptr(out_bi), _stream()), 'rpb_fm_fwd')
        _count()
        ctx.save_for_backward(e)
        ctx.mode = mode
        return out_sum if mode == 'sum' else out_bi

    @staticmethod
    def backward(ctx, g):
        (e,) = ctx.saved_tensors
        B, F, D = e.shape
        g = g.contiguous()
        de = torch.empty((B, F, D), dtype=torch.float32, device=e.device)
        dsum = g if ctx.mode == 'sum' else None
        dbi = g if ctx.mode == 'bi' else None
        check(_lib.load().rpb_fm_bwd(_ptr(e), e.stride(0), B, F, D, _ptr(dsum), _ptr(dbi), _ptr(de), F * D, 0,
                                     _stream()), 'rpb_fm_bwd')
        _count()
        return de, None


def fm_interaction(e: torch.Tensor, mode: str = 'sum') -> torch.Tensor:
    """InnerProductLayer product_sum_pooling ('sum' -> [B,1]) / Bi_interaction_pooling ('bi' -> [B,D])."""
    _cuda(e, 'feature_emb')
    return _FM.apply(e.float(), mode)


# ------------------------------------------------------------------ MLP tower
def _mlp_fwd(cfg, x, params, addend=None):
    """Forward of Linear(+ReLU)(+Dropout) x n_hidden + Linear(out).  `addend` ([M]) is added to a 1-wide output inside
    the final row-dot kernel (DeepFM: logit = fm + dnn).  Returns (out, acts, pre_drop, seeds)."""
    n_hidden, K = cfg['n_hidden'], cfg['K']
    relu, drops, training, impl = cfg['relu'], cfg['dropout'], cfg['training'], cfg['impl']
    lib = _lib.load()
    st = _stream()
    M = x.shape[0]
    acts = []          # acts[i] = input of hidden layer i (post-ReLU/post-dropout of i-1); acts[n_hidden] = last hidden out
    pre_drop = []      # pre-dropout ReLU outputs for layers with active dropout (else None)
    seeds = []
    h, ldh, kdim = x, x.stride(0), K
    for i in range(n_hidden):
        W, b = params[2 * i], params[2 * i + 1]
        N = W.shape[0]
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        check(lib.rpb_linear_fwd(_ptr(h), ldh, _ptr(W), _ptr(b), _ptr(y), N, M, N, kdim, 1 if relu[i] else 0,
                                 impl, st), 'rpb_linear_fwd')
        _count(2 if impl != 1 else 1)
        acts.append(h)
        p = drops[i] if training else 0.0
        if p > 0.0:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            yd = torch.empty_like(y)
            check(lib.rpb_dropout_fwd(_ptr(y), _ptr(yd), y.numel(), p, seed, _dropout_epoch(y.device).data_ptr(), st), 'rpb_dropout_fwd')
            _count()
            pre_drop.append(y)
            seeds.append(seed)
            y = yd
        else:
            pre_drop.append(None)
            seeds.append(0)
        h, ldh, kdim = y, N, N
    W, b = params[2 * n_hidden], params[2 * n_hidden + 1]
    N = W.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    if N == 1:
        check(lib.rpb_rowdot_fwd(_ptr(h), ldh, _ptr(W), _ptr(b), _ptr(addend), None, None, _ptr(out), M, kdim, st),
              'rpb_rowdot_fwd')
        _count()
    else:
        if addend is not None:
            raise NotImplementedError('addend with a multi-column MLP output')
        check(lib.rpb_linear_fwd(_ptr(h), ldh, _ptr(W), _ptr(b), _ptr(out), N, M, N, kdim, 0, impl, st),
              'rpb_linear_fwd')
        _count(2 if impl != 1 else 1)
    acts.append(h)
    return out, acts, pre_drop, seeds


def _mlp_bwd(cfg, acts, pre_drop, seeds, params, g, need_dx_input, layer0_hook=None):
    """Backward of _mlp_fwd.  Returns (gx | None, gparams).  `layer0_hook(dh, lddh, W0) -> bool`, when given, consumes the
    pre-activation gradient of the first layer instead of the dx GEMM (fused dx + scatter); False = not handled."""
    n_hidden, K = cfg['n_hidden'], cfg['K']
    relu, drops, training, impl = cfg['relu'], cfg['dropout'], cfg['training'], cfg['impl']
    lib = _lib.load()
    st = _stream()
    M = acts[0].shape[0]
    dev = acts[0].device
    g = _rowmajor(g)
    gparams: List[Optional[torch.Tensor]] = [None] * len(params)
    # one zero-fill for every dW/db of the tower (the kernels accumulate into them)
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
    zviews, off = [], 0
    for p in params:
        zviews.append(flat[off:off + p.numel()].view(p.shape))
        off += p.numel()

    def mask_for(layer_in_idx):
        """Activation mask to fuse when producing the grad of acts[layer_in_idx] (= output of hidden layer-1)."""
        j = layer_in_idx - 1          # hidden layer that produced this activation
        if j < 0:
            return None, False
        p = drops[j] if training else 0.0
        if p > 0.0:
            return None, True         # dropout active: unfused elementwise backward handles ReLU+dropout
        return (acts[layer_in_idx] if relu[j] else None), False

    def finish_drop(dh, j):
        """dh = grad wrt dropped output of hidden layer j -> grad wrt its pre-activation."""
        p = drops[j]
        out = torch.empty_like(dh)
        check(lib.rpb_dropout_bwd(_ptr(dh), _ptr(pre_drop[j]) if relu[j] else None, _ptr(out), dh.numel(), p,
                                  seeds[j], _dropout_epoch(dh.device).data_ptr(), st), 'rpb_dropout_bwd')
        _count()
        return out

    # output layer (always present: the reference only builds MLPs with output_dim=1 on this path)
    W, b = params[2 * n_hidden], params[2 * n_hidden + 1]
    N = W.shape[0]
    hin = acts[n_hidden]
    kdim = W.shape[1]
    mask, dropped = mask_for(n_hidden)
    dx = torch.empty((M, kdim), dtype=torch.float32, device=dev)
    dW = zviews[2 * n_hidden]
    db = zviews[2 * n_hidden + 1]
    if N == 1:
        gcol = g.reshape(-1) if g.stride(0) == 1 else g[:, 0].contiguous()
        check(lib.rpb_rowdot_bwd(_ptr(gcol), _ptr(hin), hin.stride(0), _ptr(W), _ptr(mask),
                                 mask.stride(0) if mask is not None else 0, _ptr(dx), dx.stride(0),
                                 _ptr(dW), _ptr(db), M, kdim, st), 'rpb_rowdot_bwd')
        _count(2)
    else:
        check(lib.rpb_linear_bwd(_ptr(g), g.stride(0), _ptr(hin), hin.stride(0), _ptr(W), _ptr(mask),
                                 mask.stride(0) if mask is not None else 0, _ptr(dx), dx.stride(0),
                                 _ptr(dW), _ptr(db), M, N, kdim, impl, st), 'rpb_linear_bwd')
        _count(3)
    gparams[2 * n_hidden], gparams[2 * n_hidden + 1] = dW, db
    dh = dx
    lddh = dx.stride(0)
    if dropped:
        dh = finish_drop(dh, n_hidden - 1)

    for i in range(n_hidden - 1, -1, -1):
        W, b = params[2 * i], params[2 * i + 1]
        N, kdim = W.shape[0], (K if i == 0 else W.shape[1])
        hin = acts[i]
        need_dx = i > 0 or need_dx_input
        mask, dropped = mask_for(i)
        dW = zviews[2 * i]
        db = zviews[2 * i + 1]
        if i == 0 and layer0_hook is not None:
            # parameter gradients only; dx is consumed by the hook (fused dx GEMM + embedding-gradient scatter)
            check(lib.rpb_linear_bwd(_ptr(dh), lddh, _ptr(hin), hin.stride(0), _ptr(W), None, 0, None, 0, _ptr(dW), _ptr(db),
                                     M, N, kdim, impl, st), 'rpb_linear_bwd')
            _count(2)
            gparams[0], gparams[1] = dW, db
            if layer0_hook(dh, lddh, W):
                return None, gparams
            dW = db = None                       # not handled: fall through and compute dx the plain way
            need_dx = True
        if need_dx:
            if i == 0:
                dx = torch.empty((M, hin.stride(0)), dtype=torch.float32, device=dev)
                if hin.stride(0) > kdim:
                    dx[:, kdim:].zero_()
            else:
                dx = torch.empty((M, kdim), dtype=torch.float32, device=dev)
        else:
            dx = None
        check(lib.rpb_linear_bwd(_ptr(dh), lddh, _ptr(hin), hin.stride(0), _ptr(W), _ptr(mask),
                                 mask.stride(0) if mask is not None else 0, _ptr(dx),
                                 dx.stride(0) if dx is not None else 0, _ptr(dW), _ptr(db), M, N, kdim, impl, st),
              'rpb_linear_bwd')
        _count(3)
        if dW is not None:
            gparams[2 * i], gparams[2 * i + 1] = dW, db
        dh = dx
        lddh = dx.stride(0) if dx is not None else 0
        if dropped and dh is not None:
            dh = finish_drop(dh, i - 1)
    gx = None
    if dh is not None:
        gx = dh if dh.shape[1] == acts[0].shape[1] else dh[:, :acts[0].shape[1]]
    return gx, gparams


def _tower_ok(cfg, params) -> bool:
    """True when the MLP is Linear(K->64)+ReLU, n_tail x [Linear(64->64)+ReLU], Linear(64->1) with dropout inactive:
    the shape rpb_tower_tail_* is built for (DeepFM/xDeepFM/AutoInt/FiBiNet/WDL defaults, deep.py:36-41)."""
    if not TOWER_TAIL or not cfg.get('has_out', True):
        return False
    n_hidden = cfg['n_hidden']
    if n_hidden < 1 or n_hidden - 1 > 4 or len(params) != 2 * n_hidden + 2:
        return False
    for i in range(n_hidden):
        if not cfg['relu'][i] or (cfg['training'] and cfg['dropout'][i] > 0.0):
            return False
        W, b = params[2 * i], params[2 * i + 1]
        if b is None or W.shape[0] != 64 or (i > 0 and W.shape[1] != 64):
            return False
    Wo, bo = params[2 * n_hidden], params[2 * n_hidden + 1]
    return bo is not None and Wo.shape[0] == 1 and Wo.shape[1] == 64


def _tower_fwd(cfg, x, params, addend=None, head=None):
    """Layer 1 on the tcgen05 GEMM, everything after it in rpb_tower_tail_fwd.  head = (label [M], eps, scale) also
    yields pred [M,1] and the mean-BCE loss from the same launch.  Returns (logit [M,1], acts, pred | None, loss | None)."""
    n_hidden, K, impl = cfg['n_hidden'], cfg['K'], cfg['impl']
    lib = _lib.load()
    st = _stream()
    M, dev = x.shape[0], x.device
    y1 = torch.empty((M, 64), dtype=torch.float32, device=dev)
    n_tail = n_hidden - 1
    hs = [torch.empty((M, 64), dtype=torch.float32, device=dev) for _ in range(n_tail)]
    logit = torch.empty((M, 1), dtype=torch.float32, device=dev)
    d = TowerFwdDesc()
    d.M, d.H, d.n_tail = M, 64, n_tail
    d.h1, d.ldh1 = y1.data_ptr(), 64
    keep = None
    if n_tail:
        keep = (_ptr_list([params[2 * (l + 1)] for l in range(n_tail)]),
                _ptr_list([params[2 * (l + 1) + 1] for l in range(n_tail)]), _ptr_list(hs))
        d.W, d.b, d.h = keep
    d.w_out, d.b_out = params[2 * n_hidden].data_ptr(), params[2 * n_hidden + 1].data_ptr()
    d.addend = addend.data_ptr() if addend is not None else None
    d.logit = logit.data_ptr()
    pred = loss = None
    if head is not None:
        label, eps, scale = head
        pred = torch.empty((M, 1), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        d.label, d.pred, d.loss = label.data_ptr(), pred.data_ptr(), loss.data_ptr()
        d.eps, d.scale = eps, scale
        d.work = _head_work(dev).data_ptr()
    # one launch from the feature row to the loss when the layer-1 GEMM can host the tail in its epilogue
    rc = _lib.ERR_UNSUPPORTED
    if FUSED_TOWER_EPILOGUE and impl != 1:
        rc = lib.rpb_linear_tower_fwd(_ptr(x), x.stride(0), _ptr(params[0]), _ptr(params[1]), K, C.byref(d), st)
        if rc != _lib.ERR_UNSUPPORTED:
            check(rc, 'rpb_linear_tower_fwd')
            _count(2)
    if rc == _lib.ERR_UNSUPPORTED:
        check(lib.rpb_linear_fwd(_ptr(x), x.stride(0), _ptr(params[0]), _ptr(params[1]), _ptr(y1), 64, M, 64, K, 1, impl, st),
              'rpb_linear_fwd')
        _count(2 if impl != 1 else 1)
        check(lib.rpb_tower_tail_fwd(C.byref(d), st), 'rpb_tower_tail_fwd')
        _count()
    del keep
    return logit, [x, y1] + hs, pred, loss


def _deepfm_fused_fwd(cfg, tables, idx, dense, params, head, need_grad, shards=None):
    """rpb_deepfm_fwd_fused: gather + FM + layer 1 + tower tail (+ sigmoid/BCE) in one launch.  Returns None when the
    shape is outside what that kernel takes, else (logit, acts, pred, loss, fm_s, rows) like _gather_fwd_raw + _tower_fwd."""
    F, Nd, D = len(tables), len(dense), int(tables[0].shape[1])
    M = idx[0].shape[0]
    n_hidden = cfg['n_hidden']
    n_tail = n_hidden - 1
    if D != 16 or (F & 1) or n_tail < 1 or M < 512:
        return None
    lib, st, dev = _lib.load(), _stream(), tables[0].device
    ldx = feature_row_stride(F, D, Nd)
    x = torch.empty((M, ldx), dtype=torch.float32, device=dev) if need_grad else None
    fm_s = torch.empty((M, D), dtype=torch.float32, device=dev) if need_grad else None
    rows = [int(t.shape[0]) for t in tables] if shards is None else list(shards.rows)      # sharded: GLOBAL row counts
    g = GatherDesc()
    g.B, g.F, g.D, g.Nd, g.ldx, g.ld_lr = M, F, D, Nd, ldx, 0
    keep = [_ptr_list(tables), (C.c_int64 * F)(*rows), _ptr_list(idx)]
    g.tables, g.rows, g.idx = keep
    if Nd:
        keep.append(_ptr_list(dense))
        g.dense = keep[-1]
    if shards is not None:
        g.G, g.shard_tab = shards.world, shards.w_tab.data_ptr()
    g.x, g.fm_s = _ptr(x), _ptr(fm_s)
    g.err = _err_record(dev).data_ptr()
    y1 = torch.empty((M, 64), dtype=torch.float32, device=dev)
    hs = [torch.empty((M, 64), dtype=torch.float32, device=dev) for _ in range(n_tail)]
    logit = torch.empty((M, 1), dtype=torch.float32, device=dev)
    d = TowerFwdDesc()
    d.M, d.H, d.n_tail = M, 64, n_tail
    d.h1, d.ldh1 = y1.data_ptr(), 64
    keep += [_ptr_list([params[2 * (l + 1)] for l in range(n_tail)]),
             _ptr_list([params[2 * (l + 1) + 1] for l in range(n_tail)]), _ptr_list(hs)]
    d.W, d.b, d.h = keep[-3:]
    d.w_out, d.b_out = params[2 * n_hidden].data_ptr(), params[2 * n_hidden + 1].data_ptr()
    d.logit = logit.data_ptr()
    pred = loss = None
    if head is not None:
        label, eps, scale = head
        pred = torch.empty((M, 1), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        d.label, d.pred, d.loss = label.data_ptr(), pred.data_ptr(), loss.data_ptr()
        d.eps, d.scale = eps, scale
        d.work = _head_work(dev).data_ptr()
    rc = lib.rpb_deepfm_fwd_fused(C.byref(g), _ptr(params[0]), _ptr(params[1]), C.byref(d), st)
    if rc == _lib.ERR_UNSUPPORTED:
        return None
    check(rc, 'rpb_deepfm_fwd_fused')
    _count(3)                     # layer-1 weight split, tail weight split (tcgen05 tail), the forward kernel
    del keep
    return logit, [x, y1] + hs, pred, loss, fm_s, rows


def _tower_bwd(cfg, acts, params, dlogit_in=None, head=None, gloss=None, need_dx_input=False, layer0_hook=None):
    """Backward of _tower_fwd.  dlogit comes from `dlogit_in` ([M]) or is formed in the kernel from head = (pred, label,
    eps, scale) and gloss (0-dim tensor or None = 1).  Returns (gx | None, gparams, dlogit [M])."""
    n_hidden, K, impl = cfg['n_hidden'], cfg['K'], cfg['impl']
    lib = _lib.load()
    st = _stream()
    x = acts[0]
    M, dev = x.shape[0], x.device
    n_tail = n_hidden - 1
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
    zviews, off = [], 0
    for p in params:
        zviews.append(flat[off:off + p.numel()].view(p.shape))
        off += p.numel()
    dzs = [torch.empty((M, 64), dtype=torch.float32, device=dev) for _ in range(n_hidden)]
    dlogit = torch.empty((M,), dtype=torch.float32, device=dev)
    d = TowerBwdDesc()
    d.M, d.H, d.n_tail = M, 64, n_tail
    hin = list(acts[1:1 + n_hidden])
    keep = [_ptr_list(hin), _ptr_list(dzs), _ptr_list([zviews[2 * j + 1] for j in range(n_hidden)])]
    d.hin, d.dz, d.db = keep
    d.ldh1 = hin[0].stride(0)
    if n_tail:
        keep.append(_ptr_list([params[2 * (l + 1)] for l in range(n_tail)]))
        d.W = keep[-1]
    d.w_out = params[2 * n_hidden].data_ptr()
    d.dw_out, d.db_out = zviews[2 * n_hidden].data_ptr(), zviews[2 * n_hidden + 1].data_ptr()
    if dlogit_in is not None:
        d.dlogit_in = dlogit_in.data_ptr()
    else:
        pred, label, eps, scale = head
        d.pred, d.label, d.eps, d.scale = pred.data_ptr(), label.data_ptr(), eps, scale
        if gloss is not None:
            gl = gloss.reshape(1).contiguous().float()
            d.gloss = gl.data_ptr()
    d.dlogit_out = dlogit.data_ptr()
    check(lib.rpb_tower_tail_bwd(C.byref(d), st), 'rpb_tower_tail_bwd')
    _count()
    # weight gradients (bias gradients came out of the tower kernel): dW_i = dz_i^T . input_i.  They only READ the dz
    # tiles, so the small 64x64 ones run on a side stream next to the layer-1 one (fork / join with events; the pattern
    # is capturable into the step's CUDA graph) instead of three latency-bound launches back to back.
    def wgrad(i):
        kdim = K if i == 0 else 64
        check(lib.rpb_linear_bwd(_ptr(dzs[i]), 64, _ptr(acts[i]), acts[i].stride(0), None, None, 0, None, 0,
                                 _ptr(zviews[2 * i]), None, M, 64, kdim, impl, _stream()), 'rpb_linear_bwd')
        _count()

    if n_tail and PARALLEL_WGRAD:
        main, side = torch.cuda.current_stream(), _side_stream(dev)
        fork, join = torch.cuda.Event(), torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            for i in range(n_hidden - 1, 0, -1):
                wgrad(i)
            join.record(side)
        wgrad(0)
        main.wait_event(join)
    else:
        for i in range(n_hidden - 1, -1, -1):
            wgrad(i)
    gparams = list(zviews)
    if layer0_hook is not None and layer0_hook(dzs[0], 64, params[0], dlogit):
        return None, gparams, dlogit
    gx = None
    if need_dx_input or layer0_hook is not None:
        gx = torch.empty((M, x.stride(0)), dtype=torch.float32, device=dev)
        if x.stride(0) > K:
            gx[:, K:].zero_()
        check(lib.rpb_linear_bwd(_ptr(dzs[0]), 64, None, 0, _ptr(params[0]), None, 0, _ptr(gx), gx.stride(0), None, None,
                                 M, 64, K, impl, st), 'rpb_linear_bwd')
        _count(2 if impl != 1 else 1)
        if gx.shape[1] != x.shape[1]:
            gx = gx[:, :x.shape[1]]
    return gx, gparams, dlogit


class _MLP(torch.autograd.Function):
    """Linear(+ReLU)(+Dropout) x n_hidden + Linear(out) as one autograd node so ReLU backward is fused into the
    producing GEMM's epilogue (mask = saved activation)."""

    @staticmethod
    def forward(ctx, cfg, x, *params):
        ctx.tower = _tower_ok(cfg, params)
        if ctx.tower:
            out, acts, _, _ = _tower_fwd(cfg, x, params)
            pre_drop, seeds = [None] * cfg['n_hidden'], [0] * cfg['n_hidden']
        else:
            out, acts, pre_drop, seeds = _mlp_fwd(cfg, x, params)
        ctx.cfg = cfg
        ctx.seeds = seeds
        ctx.n_saved_acts = len(acts)
        ctx.save_for_backward(*acts, *[t if t is not None else x.new_empty(0) for t in pre_drop], *params)
        return out

    @staticmethod
    def backward(ctx, g):
        cfg = ctx.cfg
        n_hidden = cfg['n_hidden']
        saved = ctx.saved_tensors
        na = ctx.n_saved_acts
        acts = saved[:na]
        pre_drop = saved[na:na + n_hidden]
        params = saved[na + n_hidden:]
        if ctx.tower:
            gx, gparams, _ = _tower_bwd(cfg, acts, params, dlogit_in=g.reshape(-1).contiguous(),
                                        need_dx_input=ctx.needs_input_grad[1])
        else:
            gx, gparams = _mlp_bwd(cfg, acts, pre_drop, ctx.seeds, params, g, ctx.needs_input_grad[1])
        return (None, gx if ctx.needs_input_grad[1] else None, *gparams)


class _DeepFMCore(torch.autograd.Function):
    """Whole DeepFM body as ONE autograd node: gather(+FM) -> MLP -> logit = fm + dnn (ranking/deepfm.py:52-61).
    Knowing both consumers of the gathered rows lets backward run the layer-1 dx GEMM with the scatter epilogue
    (rpb_linear_dx_scatter): dx + dlogit*(s - e) is added straight into the table gradients, dx never reaches HBM."""

    @staticmethod
    def forward(ctx, cfg, gcfg, *tensors):
        F, Nd = gcfg['F'], gcfg['Nd']
        tables, idx, dense = tensors[:F], tensors[F:2 * F], tensors[2 * F:2 * F + Nd]
        label = tensors[-1] if gcfg['has_label'] else None
        params = tensors[2 * F + Nd:len(tensors) - (1 if gcfg['has_label'] else 0)]
        need_grad = gcfg['needs_grad']
        ctx.tower = gcfg['tower']
        pred = loss = None
        fused = None
        head = (label, 0.0, 1.0) if label is not None else None
        st = gcfg.get('shards')
        if ctx.tower and FUSED_GATHER_GEMM and cfg['impl'] != 1:
            fused = _deepfm_fused_fwd(cfg, tables, idx, dense, params, head, need_grad, shards=st)
        if st is not None and fused is None:
            raise RuntimeError('fused DeepFM core on row-sharded tables: shape outside rpb_deepfm_fwd_fused (deepfm_core '
                               'should have declined it)')
        if fused is not None:                          # gather + FM + layer 1 + tail + loss: one kernel
            logit, acts, pred, loss, fm_s, rows = fused
            pre_drop, seeds = [None] * cfg['n_hidden'], [0] * cfg['n_hidden']
            x = acts[0] if acts[0] is not None else logit.new_empty(0)
            acts[0] = x
        else:
            x, fm, fm_s, rows = _gather_fwd_raw(tables, idx, dense, want_fm=True, need_grad=need_grad)
            if ctx.tower:
                logit, acts, pred, loss = _tower_fwd(cfg, x, params, addend=fm, head=head)
                pre_drop, seeds = [None] * cfg['n_hidden'], [0] * cfg['n_hidden']
            else:
                logit, acts, pre_drop, seeds = _mlp_fwd(cfg, x, params, addend=fm)
        ctx.set_materialize_grads(False)
        ctx.cfg, ctx.gcfg, ctx.seeds, ctx.rows = cfg, gcfg, seeds, rows
        ctx.tables = tables
        ctx.n_saved_acts = len(acts)
        ctx.n_inputs = 2 + len(tensors)
        extra = [pred, label] if pred is not None else []
        ctx.n_extra = len(extra)
        ctx.save_for_backward(fm_s, *idx, *acts, *[t if t is not None else x.new_empty(0) for t in pre_drop], *params, *extra)
        if pred is not None:
            return logit, pred, loss
        return logit

    @staticmethod
    def backward(ctx, g, g_pred=None, g_loss=None):
        cfg, gcfg = ctx.cfg, ctx.gcfg
        F, Nd, D = gcfg['F'], gcfg['Nd'], gcfg['D']
        n_hidden = cfg['n_hidden']
        saved = ctx.saved_tensors
        fm_s = saved[0]
        idx = list(saved[1:1 + F])
        na = ctx.n_saved_acts
        acts = saved[1 + F:1 + F + na]
        pre_drop = saved[1 + F + na:1 + F + na + n_hidden]
        params = saved[1 + F + na + n_hidden:len(saved) - ctx.n_extra]
        pred, label = (saved[-2], saved[-1]) if ctx.n_extra else (None, None)
        x = acts[0]
        dev = x.device
        tables = ctx.tables
        tbl_req = ctx.needs_input_grad[2:2 + F]
        store = gcfg.get('grad_store')
        st = gcfg.get('shards')
        if st is not None:
            # row-sharded: the scatter reduces into the owners' gradient shards (local HBM or NVLink); g_tables are this
            # rank's own shards and only flag which tables are trainable
            store = None
            if st.pending and all(tables[f].grad is None for f in range(F) if tbl_req[f]):
                sharded_clean(st)                     # lazy re-zero, as GradStore does for unsharded tables
            g_tables = [st.grads[f] if tbl_req[f] else None for f in range(F)]
        elif store is not None:
            trainable = [tables[f] for f in range(F) if tbl_req[f]]
            if store.pending and all(t.grad is None for t in trainable):
                store.clean()
            g_tables = [store.buffer(tables[f]) if tbl_req[f] else None for f in range(F)]
        else:
            g_tables = [torch.zeros((ctx.rows[f], D), dtype=torch.float32, device=dev) if tbl_req[f] else None for f in range(F)]

        def desc(dlogit, dx=None):
            d = ScatterDesc()
            d.B, d.F, d.D = x.shape[0], F, D
            if dx is not None:
                d.dx, d.lddx = dx.data_ptr(), dx.stride(0)
            d.dfm, d.x, d.ldx, d.fm_s = dlogit.data_ptr(), x.data_ptr(), x.stride(0), fm_s.data_ptr()   # dL/dfm = dlogit
            d._keep = (_ptr_list(g_tables), (C.c_int64 * F)(*ctx.rows), _ptr_list(idx), dlogit)
            d.grads, d.rows, d.idx = d._keep[:3]
            if st is not None:
                d.G, d.grad_shard_tab = st.world, st.g_tab.data_ptr()
            return d

        def hook(dh, lddh, W0, dlogit):
            d = desc(dlogit)
            rc = _lib.load().rpb_linear_dx_scatter(_ptr(dh), lddh, _ptr(W0), x.shape[0], W0.shape[0], cfg['K'], C.byref(d),
                                                   _stream())
            if rc == _lib.ERR_UNSUPPORTED:
                return False
            check(rc, 'rpb_linear_dx_scatter')
            _count(2)
            return True

        any_tbl = any(t is not None for t in g_tables)
        use_hook = any_tbl and FUSED_DX_SCATTER
        if ctx.tower:
            # dlogit: formed inside the tower kernel from (pred, label, g_loss) unless someone also consumed logit / pred
            dl_in = None
            if g is not None or g_pred is not None or pred is None:
                dl_in = torch.zeros((x.shape[0],), dtype=torch.float32, device=dev)
                if g is not None:
                    dl_in += g.reshape(-1)
                if g_pred is not None:
                    dl_in += g_pred.reshape(-1) * pred.reshape(-1) * (1.0 - pred.reshape(-1))
                if g_loss is not None:
                    dl = torch.empty_like(dl_in)
                    gl = g_loss.reshape(1).contiguous().float()
                    check(_lib.load().rpb_sigmoid_bce_bwd(_ptr(pred), _ptr(label), _ptr(gl), 0.0, 1.0, _ptr(dl),
                                                          x.shape[0], _stream()), 'rpb_sigmoid_bce_bwd')
                    _count()
                    dl_in += dl
            gx, gparams, dlogit = _tower_bwd(cfg, acts, params, dlogit_in=dl_in,
                                             head=(pred, label, 0.0, 1.0) if dl_in is None else None, gloss=g_loss,
                                             need_dx_input=any_tbl, layer0_hook=hook if use_hook else None)
        else:
            dlogit = g.reshape(-1).contiguous()
            gx, gparams = _mlp_bwd(cfg, acts, pre_drop, ctx.seeds, params, g, any_tbl,
                                   layer0_hook=(lambda dh, lddh, W0: hook(dh, lddh, W0, dlogit)) if use_hook else None)
        if gx is not None and any_tbl:                            # hook declined: plain scatter of dx + FM term
            gx = _rowmajor(gx)
            d = desc(dlogit, gx)
            check(_lib.load().rpb_gather_bwd(C.byref(d), _stream()), 'rpb_gather_bwd')
            _count()
        out: List[Optional[torch.Tensor]] = [None] * ctx.n_inputs
        if st is not None:
            st.pending.append(list(idx))
            st.barrier()                                  # every rank's remote gradient adds have landed
            for f in range(F):
                if g_tables[f] is not None:
                    tables[f].grad = g_tables[f]
        elif store is not None:
            store.pending.append((g_tables, None, ctx.rows, idx, D))
            for f in range(F):
                if g_tables[f] is None:
                    continue
                if tables[f].grad is None:
                    tables[f].grad = g_tables[f]
                elif tables[f].grad is not g_tables[f]:
                    raise RuntimeError("embedding table .grad was replaced externally; persistent grad mode needs "
                                       "model.zero_grad() / set grad_mode='dense'")
        else:
            for f in range(F):
                out[2 + f] = g_tables[f]
        base = 2 + 2 * F + Nd
        for i, gp in enumerate(gparams):
            out[base + i] = gp
        return tuple(out)


def _gather_fwd_raw(tables, idx, dense, want_fm, need_grad):
    """One gather launch outside autograd bookkeeping: returns (x, fm, fm_s, rows)."""
    F, Nd, D = len(tables), len(dense), int(tables[0].shape[1])
    dev = tables[0].device
    B = idx[0].shape[0]
    ldx = feature_row_stride(F, D, Nd)
    x = torch.empty((B, ldx), dtype=torch.float32, device=dev)
    fm = torch.empty((B,), dtype=torch.float32, device=dev) if want_fm else None
    fm_s = torch.empty((B, D), dtype=torch.float32, device=dev) if (want_fm and need_grad) else None
    rows = [int(t.shape[0]) for t in tables]
    d = GatherDesc()
    d.B, d.F, d.D, d.Nd, d.ldx, d.ld_lr = B, F, D, Nd, ldx, 0
    keep = (_ptr_list(tables), (C.c_int64 * F)(*rows), _ptr_list(idx))
    d.tables, d.rows, d.idx = keep
    if Nd:
        dk = _ptr_list(dense)
        d.dense = dk
    d.x, d.fm, d.fm_s = x.data_ptr(), _ptr(fm), _ptr(fm_s)
    d.err = _err_record(dev).data_ptr()
    check(_lib.load().rpb_gather_fwd(C.byref(d), _stream()), 'rpb_gather_fwd')
    _count()
    return x, fm, fm_s, rows


def deepfm_core(tables, idx, dense, weights, biases, n_hidden, relu, dropout, training, grad_store=None,
                impl: Optional[int] = None, label: Optional[torch.Tensor] = None, shards=None):
    """logit [B,1] of DeepFM (FM second order + MLP over [emb | dense]) as one fused autograd node.  With `label` ([B]
    fp32) and a tower-shaped MLP (see _tower_ok) the node also yields sigmoid(logit) and the mean BCE from the same
    launch that finishes the MLP: returns (logit, pred [B,1], loss) instead of logit.

    `shards` (dist.ShardedTables): `tables` are this rank's shard Parameters; the node then needs the one-kernel forward
    (rpb_deepfm_fwd_fused) and returns None when the shape is outside it (the caller runs the unfused sharded path)."""
    F = len(tables)
    D = int(tables[0].shape[1])
    for t in tables:
        _cuda(t, 'embedding table')
    idx_l = []
    for t in idx:
        _cuda(t, 'sparse feature column')
        t = t.reshape(-1)
        idx_l.append((t if t.dtype == torch.int64 else t.long()).contiguous())
    dense_l = [t.reshape(-1).float().contiguous() for t in dense]
    params = []
    for W, b in zip(weights, biases):
        params += [W if W.is_contiguous() else W.contiguous(), b]
    needs_grad = torch.is_grad_enabled() and (any(t.requires_grad for t in tables) or any(p.requires_grad for p in params))
    cfg = dict(n_hidden=n_hidden, has_out=True, K=F * D + len(dense_l), relu=list(relu), dropout=list(dropout),
               training=training, impl=_GEMM_IMPL if impl is None else impl)
    tower = _tower_ok(cfg, params)
    if shards is not None and not (tower and FUSED_GATHER_GEMM and cfg['impl'] != 1 and D == 16 and F % 2 == 0 and
                                   n_hidden >= 2 and idx_l[0].shape[0] >= 512):
        return None
    extra = []
    if label is not None and tower:
        _cuda(label, 'label')
        extra = [label.reshape(-1).float().contiguous()]
    gcfg = dict(F=F, Nd=len(dense_l), D=D, needs_grad=needs_grad, grad_store=grad_store if shards is None else None,
                tower=tower, has_label=bool(extra), shards=shards)
    return _DeepFMCore.apply(cfg, gcfg, *tables, *idx_l, *dense_l, *params, *extra)


def mlp_forward(x: torch.Tensor, K: int, weights: Sequence[torch.Tensor], biases: Sequence[Optional[torch.Tensor]],
                n_hidden: int, has_out: bool, relu: Sequence[bool], dropout: Sequence[float], training: bool,
                impl: Optional[int] = None) -> torch.Tensor:
    """MLP.forward (models/layers/deep.py:74-84) on x[:, :K] (x may be a wider, padded feature-row buffer)."""
    _cuda(x, 'MLP input')
    x = _rowmajor(x.float() if x.dtype != torch.float32 else x)
    params = []
    for W, b in zip(weights, biases):
        if b is None:
            raise NotImplementedError('use_bias=False MLP layers')
        params += [W if W.is_contiguous() else W.contiguous(), b]
    cfg = dict(n_hidden=n_hidden, has_out=has_out, K=K, relu=list(relu), dropout=list(dropout), training=training,
               impl=_GEMM_IMPL if impl is None else impl)
    if n_hidden < 1:
        raise NotImplementedError('MLP needs >= 1 hidden layer (deep.py:40-41)')
    if not has_out:
        # towers without an output layer (AITM click / conversion towers, deep.py:79-84 with output_dim=None): layer by layer
        # through the single-Linear op (same tcgen05 GEMM), ReLU and dropout in between
        h, k = x, K
        for i in range(n_hidden):
            h = linear(h, params[2 * i], params[2 * i + 1], K=k, impl=impl)
            if relu[i]:
                h = torch.relu(h)
            h = globals()['dropout'](h, dropout[i], training)
            k = None
        return h
    return _MLP.apply(cfg, x, *params)


# ------------------------------------------------------------------ single Linear (autograd)
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, K, W, b, impl):
        M, N = x.shape[0], W.shape[0]
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        lib = _lib.load()
        if N == 1:
            check(lib.rpb_rowdot_fwd(_ptr(x), x.stride(0), _ptr(W), _ptr(b), None, None, None, _ptr(y), M, K, _stream()),
                  'rpb_rowdot_fwd')
            _count()
        else:
            check(lib.rpb_linear_fwd(_ptr(x), x.stride(0), _ptr(W), _ptr(b), _ptr(y), N, M, N, K, 0, impl, _stream()),
                  'rpb_linear_fwd')
            _count(2 if impl != 1 else 1)
        ctx.save_for_backward(x, W)
        ctx.K, ctx.impl, ctx.has_b = K, impl, b is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        K, impl = ctx.K, ctx.impl
        M, N = x.shape[0], W.shape[0]
        g = _rowmajor(g)
        lib = _lib.load()
        need_dx = ctx.needs_input_grad[0]
        dx = None
        if need_dx:
            dx = torch.empty((M, x.shape[1]), dtype=torch.float32, device=x.device)
            if x.shape[1] > K:
                dx[:, K:].zero_()
        dW = torch.zeros_like(W)
        db = torch.zeros((N,), dtype=torch.float32, device=x.device) if ctx.has_b else None
        if N == 1:
            gcol = g.reshape(-1) if g.stride(0) == 1 else g[:, 0].contiguous()
            check(lib.rpb_rowdot_bwd(_ptr(gcol), _ptr(x), x.stride(0), _ptr(W), None, 0, _ptr(dx),
                                     dx.stride(0) if dx is not None else 0, _ptr(dW), _ptr(db), M, K, _stream()),
                  'rpb_rowdot_bwd')
            _count(2)
        else:
            check(lib.rpb_linear_bwd(_ptr(g), g.stride(0), _ptr(x), x.stride(0), _ptr(W), None, 0, _ptr(dx),
                                     dx.stride(0) if dx is not None else 0, _ptr(dW), _ptr(db), M, N, K, impl,
                                     _stream()), 'rpb_linear_bwd')
            _count(3)
        return dx, None, dW, db, None


def linear(x: torch.Tensor, W: torch.Tensor, b: Optional[torch.Tensor], K: Optional[int] = None,
           impl: Optional[int] = None) -> torch.Tensor:
    """nn.Linear on x[:, :K] (K defaults to W.shape[1])."""
    _cuda(x, 'linear input')
    x = _rowmajor(x.float() if x.dtype != torch.float32 else x)
    K = W.shape[1] if K is None else K
    return _Linear.apply(x, K, W if W.is_contiguous() else W.contiguous(), b, _GEMM_IMPL if impl is None else impl)


# ------------------------------------------------------------------ sigmoid + BCE head
_WORK = {}


def _head_work(dev) -> torch.Tensor:
    key = torch.device(dev).index or 0
    if key not in _WORK:
        _WORK[key] = torch.zeros(2 + 1024 * 2, dtype=torch.int32, device=dev)
    return _WORK[key]


class _SigmoidBCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logit, label, eps, scale):
        M = logit.numel()
        z = logit.reshape(-1).contiguous()
        pred = torch.empty((M,), dtype=torch.float32, device=logit.device)
        loss = torch.empty((), dtype=torch.float32, device=logit.device) if label is not None else None
        lab = label.reshape(-1).contiguous().float() if label is not None else None
        check(_lib.load().rpb_sigmoid_bce_fwd(_ptr(z), _ptr(lab), _ptr(pred), _ptr(loss), eps, scale, M,
                                              _ptr(_head_work(logit.device)), _stream()), 'rpb_sigmoid_bce_fwd')
        _count()
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(pred, lab if lab is not None else pred.new_empty(0))
        ctx.eps, ctx.scale, ctx.shape, ctx.has_label = eps, scale, logit.shape, label is not None
        pred_out = pred.view(logit.shape)
        if label is None:
            return pred_out
        return pred_out, loss

    @staticmethod
    def backward(ctx, gpred, gloss=None):
        pred, lab = ctx.saved_tensors
        M = pred.numel()
        dlogit = None
        if ctx.has_label and gloss is not None:
            dlogit = torch.empty((M,), dtype=torch.float32, device=pred.device)
            gl = gloss.reshape(1).contiguous().float()
            check(_lib.load().rpb_sigmoid_bce_bwd(_ptr(pred), _ptr(lab), _ptr(gl), ctx.eps, ctx.scale, _ptr(dlogit), M,
                                                  _stream()), 'rpb_sigmoid_bce_bwd')
            _count()
        if gpred is not None and ctx.needs_input_grad[0]:
            # someone consumed `pred` directly: sigmoid backward (rare; plain elementwise plumbing)
            extra = gpred.reshape(-1) * pred * (1.0 - pred)
            dlogit = extra if dlogit is None else dlogit + extra
        return (dlogit.view(ctx.shape) if dlogit is not None else None), None, None, None


def sigmoid_bce(logit: torch.Tensor, label: Optional[torch.Tensor], eps: float = 0.0, scale: float = 1.0):
    """pred = sigmoid(logit); loss = scale * mean BCE(pred + eps, label) (torch.nn.BCELoss semantics)."""
    _cuda(logit, 'logit')
    if label is None:
        return _SigmoidBCE.apply(logit, None, eps, scale), None
    _cuda(label, 'label')
    return _SigmoidBCE.apply(logit, label, eps, scale)


class _ESSMHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, y1, y2, w):
        M = z1.numel()
        a, b = z1.reshape(-1).contiguous(), z2.reshape(-1).contiguous()
        click = torch.empty((M,), dtype=torch.float32, device=z1.device)
        conv = torch.empty_like(click)
        has_label = y1 is not None
        loss = torch.empty((), dtype=torch.float32, device=z1.device) if has_label else None
        t1 = y1.reshape(-1).contiguous().float() if has_label else None
        t2 = y2.reshape(-1).contiguous().float() if has_label else None
        check(_lib.load().rpb_essm_head_fwd(_ptr(a), _ptr(b), _ptr(t1), _ptr(t2), _ptr(click), _ptr(conv), _ptr(loss), w, M,
                                            _ptr(_head_work(z1.device)), _stream()), 'rpb_essm_head_fwd')
        _count()
        ctx.set_materialize_grads(False)
        ctx.w, ctx.shape, ctx.has_label = w, z1.shape, has_label
        if has_label:
            ctx.save_for_backward(click, conv, t1, t2)
        c_out, v_out = click.view(z1.shape), conv.view(z2.shape)
        ctx.mark_non_differentiable(c_out, v_out)
        return (c_out, v_out, loss) if has_label else (c_out, v_out)

    @staticmethod
    def backward(ctx, gc, gv, gloss=None):
        if not ctx.has_label or gloss is None:
            return None, None, None, None, None
        click, conv, t1, t2 = ctx.saved_tensors
        M = click.numel()
        dz1, dz2 = torch.empty_like(click), torch.empty_like(click)
        gl = gloss.reshape(1).contiguous().float()
        check(_lib.load().rpb_essm_head_bwd(_ptr(click), _ptr(conv), _ptr(t1), _ptr(t2), _ptr(gl), ctx.w, _ptr(dz1), _ptr(dz2),
                                            M, _stream()), 'rpb_essm_head_bwd')
        _count()
        return dz1.view(ctx.shape), dz2.view(ctx.shape), None, None, None


def essm_head(z1: torch.Tensor, z2: torch.Tensor, y1: Optional[torch.Tensor] = None, y2: Optional[torch.Tensor] = None,
              w_ctr: float = 0.5):
    """ESSM head (multi_task/essm.py:50-75): (click, conversion[, loss]) with click = sigmoid(z1), conversion = sigmoid(z2),
    loss = mean BCE(click * conversion, y2) + w_ctr * mean BCE(click, y1).  The predictions are outputs only (gradients
    flow through the loss)."""
    _cuda(z1, 'ctr logit')
    _cuda(z2, 'cvr logit')
    return _ESSMHead.apply(z1, z2, y1, y2, float(w_ctr))


# ------------------------------------------------------------------ DCN CrossNet
def _ptr_array(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() if t is not None else 0 for t in ts])


class _CrossNet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, K, L, *params):            # params = w_0..w_{L-1} ([1,K] or [K]), b_0..b_{L-1}
        ws, bs = params[:L], params[L:]
        B = x0.shape[0]
        ldo = (K + 3) // 4 * 4
        out = torch.empty((B, ldo), dtype=torch.float32, device=x0.device)
        need = any(ctx.needs_input_grad)
        S = torch.empty((B, L), dtype=torch.float32, device=x0.device) if need else None
        check(_lib.load().rpb_crossnet_fwd(_ptr(x0), x0.stride(0), K, L, _ptr_array(ws), _ptr_array(bs), _ptr(out), ldo,
                                           _ptr(S), B, _stream()), 'rpb_crossnet_fwd')
        _count()
        ctx.K, ctx.L = K, L
        ctx.save_for_backward(x0, S, *params)
        return out

    @staticmethod
    def backward(ctx, g):
        x0, S, *params = ctx.saved_tensors
        K, L = ctx.K, ctx.L
        ws, bs = params[:L], params[L:]
        B = x0.shape[0]
        g = _rowmajor(g)
        dx0 = torch.empty((B, x0.shape[1]), dtype=torch.float32, device=x0.device) if ctx.needs_input_grad[0] else None
        dws = [torch.zeros_like(w) for w in ws]
        dbs = [torch.zeros_like(b) for b in bs]
        check(_lib.load().rpb_crossnet_bwd(_ptr(x0), x0.stride(0), K, L, _ptr_array(ws), _ptr_array(bs), _ptr(S), _ptr(g),
                                           g.stride(0), _ptr(dx0), dx0.stride(0) if dx0 is not None else 0,
                                           _ptr_array(dws), _ptr_array(dbs), B, _stream()), 'rpb_crossnet_bwd')
        _count(5)
        return (dx0, None, None, *dws, *dbs)


def crossnet(x0: torch.Tensor, K: int, weights, biases) -> torch.Tensor:
    """CrossNet.forward (interaction.py:137-141) on x0[:, :K]; returns [B, round_up(K,4)] (pad columns zero)."""
    _cuda(x0, 'CrossNet input')
    x0 = _rowmajor(x0.float() if x0.dtype != torch.float32 else x0)
    ws = [w if w.is_contiguous() else w.contiguous() for w in weights]
    return _CrossNet.apply(x0, K, len(ws), *ws, *biases)


# ------------------------------------------------------------------ xDeepFM CIN
# 1: the CIN forward keeps X_1..X_{L-1} for backward when the tensor-core kernels take the shape (RPB_CIN_SAVE_X=0: recompute)
CIN_SAVE_X = int(os.environ.get('RPB_CIN_SAVE_X', '1'))


class _CIN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e, L, *params):                # params = W_0..W_{L-1} ([U,Cin,1]), b_0..b_{L-1}
        Ws, bs = params[:L], params[L:]
        B, F, D = e.shape
        units = [int(w.shape[0]) for w in Ws]
        tot = sum(units)
        pooled = torch.empty((B, tot), dtype=torch.float32, device=e.device)
        u_arr = (C.c_int32 * L)(*units)
        # training: keep X_1 .. X_{L-1} (what autograd keeps in the reference) so that backward does not recompute the forward —
        # tensor-core shapes only; RPB_ERR_UNSUPPORTED means "use the recomputing pair"
        xsave = None
        needs_grad = any(ctx.needs_input_grad)           # (grad mode is off inside Function.forward: ask the context)
        if needs_grad and L > 1 and CIN_SAVE_X:
            ldx = (tot - units[-1]) * D
            xsave = torch.empty((B, ldx), dtype=torch.float32, device=e.device)
            rc = _lib.load().rpb_cin_fwd_save(_ptr(e), e.stride(0), B, F, D, L, u_arr, _ptr_array(Ws), _ptr_array(bs),
                                              _ptr(pooled), tot, _ptr(xsave), ldx, _stream())
            if rc == _lib.ERR_UNSUPPORTED:
                xsave = None
            else:
                check(rc, 'rpb_cin_fwd_save')
        if xsave is None:
            check(_lib.load().rpb_cin_fwd(_ptr(e), e.stride(0), B, F, D, L, u_arr, _ptr_array(Ws), _ptr_array(bs),
                                          _ptr(pooled), tot, _stream()), 'rpb_cin_fwd')
        _count(2)
        ctx.L, ctx.units = L, units
        ctx.has_x = xsave is not None
        if xsave is not None:
            ctx.save_for_backward(e, xsave, *params)
        else:
            ctx.save_for_backward(e, *params)
        return pooled

    @staticmethod
    def backward(ctx, g):
        if ctx.has_x:
            e, xsave, *params = ctx.saved_tensors
        else:
            e, *params = ctx.saved_tensors
            xsave = None
        L, units = ctx.L, ctx.units
        Ws, bs = params[:L], params[L:]
        B, F, D = e.shape
        g = _rowmajor(g)
        de = torch.empty((B, F, D), dtype=torch.float32, device=e.device)
        dWs = [torch.zeros_like(w) for w in Ws]
        dbs = [torch.zeros_like(b) for b in bs]
        u_arr = (C.c_int32 * L)(*units)
        if xsave is not None:
            check(_lib.load().rpb_cin_bwd_saved(_ptr(e), e.stride(0), B, F, D, L, u_arr, _ptr_array(Ws), _ptr_array(bs), _ptr(g),
                                                g.stride(0), _ptr(de), F * D, 0, _ptr_array(dWs), _ptr_array(dbs),
                                                _ptr(xsave), xsave.stride(0), _stream()), 'rpb_cin_bwd_saved')
        else:
            check(_lib.load().rpb_cin_bwd(_ptr(e), e.stride(0), B, F, D, L, u_arr, _ptr_array(Ws), _ptr_array(bs), _ptr(g),
                                          g.stride(0), _ptr(de), F * D, 0, _ptr_array(dWs), _ptr_array(dbs), _stream()),
                  'rpb_cin_bwd')
        _count(3)
        return (de, None, *dWs, *dbs)


def cin(e: torch.Tensor, weights, biases) -> torch.Tensor:
    """CompressedInteractionNet up to the pooled vector [B, sum(units)] (interaction.py:157-169)."""
    _cuda(e, 'feature_emb')
    if e.dtype != torch.float32:
        e = e.float()
    B, F, D = e.shape
    if e.stride(2) != 1 or e.stride(1) != D:
        e = e.contiguous()
    Ws = [w if w.is_contiguous() else w.contiguous() for w in weights]
    return _CIN.apply(e, len(Ws), *Ws, *biases)


# ------------------------------------------------------------------ AutoInt interacting layer
class _AutoIntCore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkvr, res, B, F, H, d):
        HD = H * d
        out = torch.empty((B, F, HD), dtype=torch.float32, device=qkvr.device)
        check(_lib.load().rpb_autoint_attn_fwd(_ptr(qkvr), qkvr.stride(0), _ptr(res), res.stride(0) if res is not None else 0,
                                               _ptr(out), B, F, H, d, _stream()), 'rpb_autoint_attn_fwd')
        _count()
        ctx.dims = (B, F, H, d)
        ctx.has_res_proj = res is None
        ctx.save_for_backward(qkvr, out)
        return out

    @staticmethod
    def backward(ctx, g):
        qkvr, out = ctx.saved_tensors
        B, F, H, d = ctx.dims
        HD = H * d
        g = g.contiguous()
        dq = torch.empty((B * F, 4 * HD), dtype=torch.float32, device=qkvr.device)
        check(_lib.load().rpb_autoint_attn_bwd(_ptr(qkvr), qkvr.stride(0), 1 if ctx.has_res_proj else 0, _ptr(out), _ptr(g),
                                               _ptr(dq), 4 * HD, B, F, H, d, _stream()), 'rpb_autoint_attn_bwd')
        _count()
        if ctx.has_res_proj:
            return dq, None, None, None, None, None
        return dq[:, :3 * HD], dq[:, 3 * HD:], None, None, None, None


def autoint_attention(X: torch.Tensor, Wq, Wk, Wv, Wres, num_heads: int, attention_dim: int) -> torch.Tensor:
    """MultiHeadSelfAttention.forward (attention.py:98-101) for the AutoInt configuration -> [B, F, H*d].
    Projections run on the dense-layer GEMM (tcgen05 when the shape qualifies), the attention core in one kernel."""
    _cuda(X, 'attention input')
    B, F, D = X.shape
    Xc = X.float().contiguous().view(B * F, D)          # [B*F, D] matrix for the projection GEMM
    Wcat = torch.cat([Wq, Wk, Wv] + ([Wres] if Wres is not None else []), dim=0)     # [(3|4)*H*d, D]
    qkvr = linear(Xc, Wcat, None)
    res = None if Wres is not None else Xc
    return _AutoIntCore.apply(qkvr, res, B, F, num_heads, attention_dim)


# ------------------------------------------------------------------ MMOE pieces
class _MatmulKN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, K, Wkn, bias, impl):
        M, N = x.shape[0], Wkn.shape[1]
        ldy = (N + 3) // 4 * 4
        y = torch.empty((M, ldy), dtype=torch.float32, device=x.device)
        if ldy > N:
            y[:, N:].zero_()
        check(_lib.load().rpb_matmul_kn_fwd(_ptr(x), x.stride(0), _ptr(Wkn), Wkn.stride(0), _ptr(bias), _ptr(y), ldy, M, N, K,
                                            impl, _stream()), 'rpb_matmul_kn_fwd')
        _count(2 if impl != 1 else 1)
        ctx.save_for_backward(x, Wkn)
        ctx.K, ctx.impl, ctx.has_b = K, impl, bias is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, Wkn = ctx.saved_tensors
        K, impl = ctx.K, ctx.impl
        M, N = x.shape[0], Wkn.shape[1]
        g = _rowmajor(g)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, x.shape[1]), dtype=torch.float32, device=x.device)
            if x.shape[1] > K:
                dx[:, K:].zero_()
        dW = torch.zeros_like(Wkn)
        db = torch.zeros((N,), dtype=torch.float32, device=x.device) if ctx.has_b else None
        check(_lib.load().rpb_matmul_kn_bwd(_ptr(g), g.stride(0), _ptr(x), x.stride(0), _ptr(Wkn), Wkn.stride(0), _ptr(dx),
                                            dx.stride(0) if dx is not None else 0, _ptr(dW), _ptr(db), M, N, K, impl,
                                            _stream()), 'rpb_matmul_kn_bwd')
        _count(4)
        return dx, None, dW, db, None


def matmul_kn(x: torch.Tensor, Wkn: torch.Tensor, bias: Optional[torch.Tensor], K: Optional[int] = None,
              impl: Optional[int] = None) -> torch.Tensor:
    """x[:, :K] @ Wkn[K, N] + bias -> [M, round_up(N,4)] (pad columns zero)."""
    _cuda(x, 'matmul input')
    x = _rowmajor(x.float() if x.dtype != torch.float32 else x)
    K = Wkn.shape[0] if K is None else K
    return _MatmulKN.apply(x, K, Wkn if Wkn.is_contiguous() else Wkn.contiguous(), bias,
                           _GEMM_IMPL if impl is None else impl)


class _MMOECombine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eo, Hh, E, T):
        B = eo.shape[0]
        out = torch.empty((T, B, Hh), dtype=torch.float32, device=eo.device)
        gate = torch.empty((B, T * E), dtype=torch.float32, device=eo.device)
        check(_lib.load().rpb_mmoe_combine_fwd(_ptr(eo), eo.stride(0), B, Hh, E, T, _ptr(out), _ptr(gate), _stream()),
              'rpb_mmoe_combine_fwd')
        _count()
        ctx.dims = (Hh, E, T)
        ctx.save_for_backward(eo, gate)
        return out

    @staticmethod
    def backward(ctx, g):
        eo, gate = ctx.saved_tensors
        Hh, E, T = ctx.dims
        B = eo.shape[0]
        g = g.contiguous()
        deo = torch.empty_like(eo)
        check(_lib.load().rpb_mmoe_combine_bwd(_ptr(eo), eo.stride(0), _ptr(gate), _ptr(g), B, Hh, E, T, _ptr(deo),
                                               deo.stride(0), _stream()), 'rpb_mmoe_combine_bwd')
        _count()
        return deo, None, None, None


def mmoe_combine(eo: torch.Tensor, Hh: int, E: int, T: int) -> torch.Tensor:
    """Gate softmax + gated sum over experts (mmoe.py:90-104): eo [B, >= Hh*E + T*E] -> [T, B, Hh]."""
    _cuda(eo, 'experts_out')
    return _MMOECombine.apply(_rowmajor(eo), Hh, E, T)


# Process group over which BatchNorm batch statistics are shared (None = per-process statistics, the single-GPU case).
# Set by rec_pangu_b200.dist.enable_sync_batchnorm(); with it the data-parallel towers normalise with the statistics of
# the GLOBAL batch, i.e. exactly what the single-process reference computes on the concatenated batch.
SYNC_BN_GROUP = None


class _BatchNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, momentum, eps):
        M, N = x.shape
        lib, st = _lib.load(), _stream()
        group = SYNC_BN_GROUP if training else None
        count = float(M)
        if training:
            stats = torch.zeros((2 * N + 1,), dtype=torch.float32, device=x.device)
            check(lib.rpb_bn_stats(_ptr(x), M, N, _ptr(stats[:N]), _ptr(stats[N:2 * N]), st), 'rpb_bn_stats')
            _count()
            if group is not None:
                import torch.distributed as dist
                stats[2 * N] = float(M)
                dist.all_reduce(stats, group=group)                  # sums, sums of squares and the sample count
                count = float(M) * dist.get_world_size(group)        # equal per-rank batches (weak scaling)
            mean = stats[:N] / count
            var = (stats[N:2 * N] / count - mean * mean).clamp_min_(0.0)     # biased (normalisation) variance; [N]-sized plumbing
            with torch.no_grad():
                running_mean.mul_(1 - momentum).add_(mean, alpha=momentum)
                running_var.mul_(1 - momentum).add_(var * (count / max(count - 1, 1)), alpha=momentum)
        else:
            mean, var = running_mean, running_var
        invstd = torch.rsqrt(var + eps)
        y = torch.empty_like(x)
        check(lib.rpb_bn_apply(_ptr(x), _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(beta), _ptr(y), M, N, st), 'rpb_bn_apply')
        _count()
        ctx.save_for_backward(x, gamma, mean.contiguous(), invstd)
        ctx.training, ctx.group, ctx.count = training, group, count
        return y

    @staticmethod
    def backward(ctx, g):
        x, gamma, mean, invstd = ctx.saved_tensors
        M, N = x.shape
        g = g.contiguous()
        lib, st = _lib.load(), _stream()
        dx = torch.empty_like(x)
        dgb = torch.zeros((2, N), dtype=torch.float32, device=x.device)        # [dgamma ; dbeta] of THIS rank's samples
        if ctx.group is None:
            check(lib.rpb_bn_bwd(_ptr(g), _ptr(x), _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(dx), _ptr(dgb[0]),
                                 _ptr(dgb[1]), M, N, 1 if ctx.training else 0, st), 'rpb_bn_bwd')
            _count(2)
            return dx, dgb[0], dgb[1], None, None, None, None, None
        import torch.distributed as dist
        check(lib.rpb_bn_bwd_stats(_ptr(g), _ptr(x), _ptr(mean), _ptr(invstd), _ptr(dgb[0]), _ptr(dgb[1]), M, N, st),
              'rpb_bn_bwd_stats')
        tot = dgb.clone()
        dist.all_reduce(tot, group=ctx.group)                                  # column sums over the global batch
        check(lib.rpb_bn_bwd_dx(_ptr(g), _ptr(x), _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(tot[0]), _ptr(tot[1]), _ptr(dx),
                                M, N, 1.0 / ctx.count, 1, st), 'rpb_bn_bwd_dx')
        _count(2)
        # parameter gradients stay the local sums: the dense-gradient all-reduce (dist.DenseGradBucket) adds the ranks
        return dx, dgb[0], dgb[1], None, None, None, None, None


def batch_norm(x: torch.Tensor, bn: torch.nn.BatchNorm1d, training: bool) -> torch.Tensor:
    """nn.BatchNorm1d (mmoe.py:54-56) on a contiguous [M, N] CUDA tensor; updates running stats in training mode."""
    _cuda(x, 'batch-norm input')
    if training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return _BatchNorm.apply(x.contiguous(), bn.weight, bn.bias, bn.running_mean, bn.running_var, training,
                            bn.momentum if bn.momentum is not None else 0.1, bn.eps)


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        y = torch.empty_like(x)
        check(_lib.load().rpb_dropout_fwd(_ptr(x), _ptr(y), x.numel(), p, seed, _dropout_epoch(x.device).data_ptr(), _stream()), 'rpb_dropout_fwd')
        _count()
        ctx.p, ctx.seed = p, seed
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dx = torch.empty_like(g)
        check(_lib.load().rpb_dropout_bwd(_ptr(g), None, _ptr(dx), g.numel(), ctx.p, ctx.seed, _dropout_epoch(g.device).data_ptr(), _stream()), 'rpb_dropout_bwd')
        _count()
        return dx, None, None


def dropout(x: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    """nn.Dropout (counter-based mask, recomputed in backward)."""
    if not training or p <= 0.0:
        return x
    return _Dropout.apply(x.contiguous(), float(p), int(torch.randint(0, 2 ** 62, (1,)).item()))


# ------------------------------------------------------------------ LayerNorm (MaskNet's MaskBlock)
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, K, gamma, beta, eps):
        M = x.shape[0]
        y = torch.empty((M, K), dtype=torch.float32, device=x.device)
        mean = torch.empty((M,), dtype=torch.float32, device=x.device)
        rstd = torch.empty((M,), dtype=torch.float32, device=x.device)
        check(_lib.load().rpb_layernorm_fwd(_ptr(x), x.stride(0), _ptr(gamma), _ptr(beta), eps, _ptr(y), y.stride(0), _ptr(mean),
                                            _ptr(rstd), M, K, _stream()), 'rpb_layernorm_fwd')
        _count()
        ctx.K = K
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, g):
        x, gamma, mean, rstd = ctx.saved_tensors
        M, K = x.shape[0], ctx.K
        g = _rowmajor(g)
        dx = torch.zeros_like(x) if x.shape[1] != K else torch.empty_like(x)      # pad columns of a feature row get zero
        dgamma = torch.zeros_like(gamma)
        dbeta = torch.zeros_like(gamma)
        check(_lib.load().rpb_layernorm_bwd(_ptr(g), g.stride(0), _ptr(x), x.stride(0), _ptr(gamma), _ptr(mean), _ptr(rstd),
                                            _ptr(dx), dx.stride(0), _ptr(dgamma), _ptr(dbeta), M, K, _stream()), 'rpb_layernorm_bwd')
        _count()
        return dx, None, dgamma, dbeta, None


def layer_norm(x: torch.Tensor, ln: torch.nn.LayerNorm, K: Optional[int] = None) -> torch.Tensor:
    """torch.nn.LayerNorm over the first K columns of x [M, >= K] (K defaults to the module's normalized_shape)."""
    _cuda(x, 'layer_norm input')
    K = int(ln.normalized_shape[0]) if K is None else K
    return _LayerNorm.apply(_rowmajor(x), K, ln.weight, ln.bias, float(ln.eps))


# ------------------------------------------------------------------ FiBiNet: SENET + bilinear (both passes) fused
class _FiBiNet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, F, D, Nd, W1, W2, Wb):
        B = x.shape[0]
        P = F * (F - 1) // 2
        K = 2 * P * D + Nd
        ldc = (K + 3) // 4 * 4
        comb = torch.empty((B, ldc), dtype=torch.float32, device=x.device)
        A = torch.empty((B, F), dtype=torch.float32, device=x.device)
        R = W1.shape[0]
        check(_lib.load().rpb_fibinet_fwd(_ptr(x), x.stride(0), B, F, D, Nd, _ptr(W1), R, _ptr(W2), _ptr(Wb), _ptr(comb), ldc,
                                          _ptr(A), _stream()), 'rpb_fibinet_fwd')
        _count()
        ctx.dims = (F, D, Nd)
        ctx.save_for_backward(x, A, W1, W2, Wb)
        return comb

    @staticmethod
    def backward(ctx, g):
        x, A, W1, W2, Wb = ctx.saved_tensors
        F, D, Nd = ctx.dims
        B = x.shape[0]
        g = _rowmajor(g)
        dx = torch.empty((B, x.shape[1]), dtype=torch.float32, device=x.device)
        dW1, dW2, dWb = torch.zeros_like(W1), torch.zeros_like(W2), torch.zeros_like(Wb)
        check(_lib.load().rpb_fibinet_bwd(_ptr(x), x.stride(0), B, F, D, _ptr(W1), W1.shape[0], _ptr(W2), _ptr(Wb), _ptr(A),
                                          _ptr(g), g.stride(0), _ptr(dx), dx.stride(0), _ptr(dW1), _ptr(dW2), _ptr(dWb),
                                          _stream()), 'rpb_fibinet_bwd')
        _count(3)
        return dx, None, None, None, dW1, dW2, dWb


def fibinet_interaction(x: torch.Tensor, F: int, D: int, Nd: int, W1: torch.Tensor, W2: torch.Tensor,
                        Wb: torch.Tensor) -> torch.Tensor:
    """[bilinear(E) | bilinear(SENET(E)) | dense] MLP input of FiBiNet (fibinet.py:59-66) from the feature row x."""
    _cuda(x, 'feature row')
    return _FiBiNet.apply(_rowmajor(x), F, D, Nd, W1.contiguous(), W2.contiguous(), Wb.contiguous())


def senet(e: torch.Tensor, W1: torch.Tensor, W2: torch.Tensor) -> torch.Tensor:
    raise NotImplementedError('standalone SENET_Layer.forward: use FiBiNet (fused rpb_fibinet_* kernels)')


def bilinear(e: torch.Tensor, Wb: torch.Tensor) -> torch.Tensor:
    """Standalone BilinearInteractionLayer('field_interaction').forward -> [B, P, D] through the fused FiBiNet kernel
    (zero excitation weights switch the SENET branch off; its half of the output row is ignored)."""
    _cuda(e, 'feature_emb')
    B, F, D = e.shape
    x = e.float().contiguous().view(B, F * D)
    R = max(1, F // 3)
    W1 = torch.zeros((R, F), dtype=torch.float32, device=e.device)
    W2 = torch.zeros((F, R), dtype=torch.float32, device=e.device)
    P = F * (F - 1) // 2
    comb = _FiBiNet.apply(x, F, D, 0, W1, W2, Wb.contiguous())
    return comb[:, :P * D].reshape(B, P, D)
