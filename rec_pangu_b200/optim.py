"""Fused optimizer for the hot path.

`FusedAdam(model, lr)` has torch.optim.Adam's hyper-parameters (the reference creates
`torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-08, weight_decay=0)`, trainer.py:75).
Dense parameters get the element-wise Adam kernel; embedding tables in 'persistent' grad mode get the row-sparse kernel
(`rpb_sparse_adam`), which touches only the rows of the current batch and re-zeroes the gradient buffer in the same
pass.  Semantics of the sparse part, default: torch.optim.SparseAdam ("lazy" Adam) — a row that receives no gradient in a
step is left untouched, whereas the reference's dense Adam would still move it by its momentum.  `exact=True` removes that
deviation without touching more rows ("exact-lazy"): every row carries the number of the last step it received
(`stamp`); a forward pre-hook replays, for exactly the rows the batch is about to read, the zero-gradient steps they missed
(`rpb_sparse_adam_catchup`), and `flush()` — also run before every `model.state_dict()` — does it for all rows.  The
trained model then equals the one torch.optim.Adam(model.parameters()) produces (tests/test_models_gpu.py, all rows)."""
import ctypes as C

import torch

from . import _lib, ops
from ._lib import AdamMultiDesc, SparseAdamDesc, check
from .models.layers.embedding import EmbeddingLayer


class FusedAdam:
    graph_safe = True           # step counter lives on the device: model_pipeline.train_model may capture the step as a CUDA graph

    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, exact: bool = False):
        self.model, self.lr, self.betas, self.eps = model, lr, betas, eps
        self.exact = exact
        self.step_count = 0
        self.step_dev = None                       # device int32 step counter (graph-replay safe)
        self.emb_layers = [m for m in model.modules() if isinstance(m, EmbeddingLayer) and m._shards is None]
        for m in self.emb_layers:
            m.grad_mode = 'persistent'
        table_ids = {id(p) for m in self.emb_layers for p in m.tables()}
        self.dense = [p for p in model.parameters() if id(p) not in table_ids and p.requires_grad]
        self.state = {}
        self._hooks = []
        if exact:
            self._hooks.append(model.register_forward_pre_hook(self._pre_forward))
            self._hooks.append(model.register_state_dict_pre_hook(lambda module, prefix, keep_vars: self.flush()))

    # ---- exact-lazy Adam: bring rows up to date before they are read
    def _table_state(self, p):
        s = self._st(p)
        if 'stamp' not in s:
            s['stamp'] = torch.zeros(p.shape[0], dtype=torch.int32, device=p.device)
        return s

    @torch.no_grad()
    def _pre_forward(self, module, args):
        """Forward pre-hook: the rows this batch is about to read receive the zero-gradient steps they missed."""
        if self.step_dev is None or not args or not isinstance(args[0], dict):
            return
        data = args[0]
        lib, st = _lib.load(), C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for emb in self.emb_layers:
            params = [p for p in emb.tables()]
            if not all(c in data for c in emb.emb_feature) or not any(p.requires_grad for p in params):
                continue
            idx = []
            for c in emb.emb_feature:
                t = data[c].reshape(-1)
                idx.append((t if t.dtype == torch.int64 else t.long()).contiguous())
            F, D = len(idx), int(params[0].shape[1])
            states = [self._table_state(p) if p.requires_grad else None for p in params]
            arr = lambda ts: (C.c_void_p * F)(*[t.data_ptr() if t is not None else 0 for t in ts])  # noqa: E731
            d = SparseAdamDesc()
            d.B, d.F, d.D, d.step = idx[0].shape[0], F, D, self.step_count
            d.lr, d.beta1, d.beta2, d.eps = self.lr, self.betas[0], self.betas[1], self.eps
            keep = (arr([p if s is not None else None for p, s in zip(params, states)]), arr([None] * F),
                    arr([s['m'] if s else None for s in states]), arr([s['v'] if s else None for s in states]),
                    arr([s['stamp'] if s else None for s in states]), (C.c_int64 * F)(*[int(p.shape[0]) for p in params]), arr(idx))
            d.weights, d.grads, d.exp_avg, d.exp_avg_sq, d.stamps, d.rows, d.idx = keep
            d.step_dev = self.step_dev.data_ptr()
            check(lib.rpb_sparse_adam_catchup(C.byref(d), st), 'rpb_sparse_adam_catchup')
            ops._count()

    @torch.no_grad()
    def flush(self):
        """Every row of every table receives the steps it missed (exact mode; a no-op otherwise): call before reading the
        tables as a whole — `model.state_dict()` does it through a hook."""
        if not self.exact or self.step_dev is None:
            return
        lib, st = _lib.load(), C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for emb in self.emb_layers:
            for p in emb.tables():
                if not p.requires_grad:
                    continue
                s = self._table_state(p)
                check(lib.rpb_sparse_adam_flush(p.data_ptr(), s['m'].data_ptr(), s['v'].data_ptr(), s['stamp'].data_ptr(),
                                                int(p.shape[0]), int(p.shape[1]), self.lr, self.betas[0], self.betas[1], self.eps,
                                                self.step_dev.data_ptr(), self.step_count, st), 'rpb_sparse_adam_flush')
                ops._count(2)

    def _st(self, p):
        s = self.state.get(id(p))
        if s is None:
            s = {'m': torch.zeros_like(p), 'v': torch.zeros_like(p)}
            self.state[id(p)] = s
        return s

    @torch.no_grad()
    def step(self):
        """One optimizer step.  The step number lives on the device (`self.step_dev`, incremented on the stream), so the
        whole step — including its bias correction and the per-step row stamps — can be captured into a CUDA graph and
        replayed."""
        self.step_count += 1
        lib, st = _lib.load(), C.c_void_p(torch.cuda.current_stream().cuda_stream)
        live = [p for p in self.dense if p.grad is not None]
        if self.step_dev is None and (live or self.emb_layers):
            dev = live[0].device if live else next(self.model.parameters()).device
            self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        if self.step_dev is not None:
            self.step_dev.add_(1)
        for c0 in range(0, len(live), _lib.ADAM_MAX_TENSORS):
            chunk = live[c0:c0 + _lib.ADAM_MAX_TENSORS]
            n = len(chunk)
            states = [self._st(p) for p in chunk]
            grads = [p.grad.contiguous() for p in chunk]
            arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
            keep = (arr(chunk), arr(grads), arr([s['m'] for s in states]), arr([s['v'] for s in states]),
                    (C.c_int64 * n)(*[p.numel() for p in chunk]), grads)
            d = AdamMultiDesc()
            d.count, d.step = n, self.step_count
            d.lr, d.beta1, d.beta2, d.eps = self.lr, self.betas[0], self.betas[1], self.eps
            d.params, d.grads, d.exp_avg, d.exp_avg_sq, d.numel = keep[:5]
            d.step_dev = self.step_dev.data_ptr()
            check(lib.rpb_adam_multi(C.byref(d), st), 'rpb_adam_multi')
            ops._count()
        for emb in self.emb_layers:
            store = emb._grad_store
            for grads, lr_grads, rows, idx, D in store.pending:
                self._sparse(lib, st, emb.tables(), grads, rows, idx, D)
                if lr_grads is not None:
                    # D=1 LR tables rode in the same gather: their params are matched by buffer identity
                    lr_params = [self._param_of(store, g) for g in lr_grads]
                    self._sparse_scalar(lr_params, lr_grads, rows, idx)
            store.pending = []                     # gradient rows were re-zeroed by the fused kernel

    def _param_of(self, store, buf):
        for p in self.model.parameters():
            if store.buffers.get(id(p)) is buf:
                return p
        return None

    def _sparse(self, lib, st, params, grads, rows, idx, D):
        F = len(idx)
        d = SparseAdamDesc()
        d.B, d.F, d.D, d.step = idx[0].shape[0], F, D, self.step_count
        d.lr, d.beta1, d.beta2, d.eps = self.lr, self.betas[0], self.betas[1], self.eps
        states = [self._st(p) if (g is not None and p is not None) else None for p, g in zip(params, grads)]
        for p, s in zip(params, states):
            if s is not None and 'stamp' not in s:
                s['stamp'] = torch.zeros(p.shape[0], dtype=torch.int32, device=p.device)
        arr = lambda ts: (C.c_void_p * F)(*[t.data_ptr() if t is not None else 0 for t in ts])  # noqa: E731
        w_arr = arr([p if (g is not None and s is not None) else None for p, g, s in zip(params, grads, states)])
        g_arr = arr(grads)
        m_arr = arr([s['m'] if s else None for s in states])
        v_arr = arr([s['v'] if s else None for s in states])
        s_arr = arr([s['stamp'] if s else None for s in states])
        r_arr = (C.c_int64 * F)(*rows)
        i_arr = arr(idx)
        d.weights, d.grads, d.exp_avg, d.exp_avg_sq, d.stamps, d.rows, d.idx = w_arr, g_arr, m_arr, v_arr, s_arr, r_arr, i_arr
        d.step_dev = self.step_dev.data_ptr()
        check(lib.rpb_sparse_adam(C.byref(d), st), 'rpb_sparse_adam')
        ops._count()

    def _sparse_scalar(self, params, grads, rows, idx):
        """D = 1 tables (LR_Layer): the scalar variant of rpb_sparse_adam — same stamp claim, graph-safe."""
        self._sparse(_lib.load(), C.c_void_p(torch.cuda.current_stream().cuda_stream), params, grads, rows, idx, 1)

    def zero_grad(self, set_to_none: bool = True):
        self.model.zero_grad(set_to_none=set_to_none)
