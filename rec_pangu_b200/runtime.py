"""Step runners around the model API: CUDA-graph capture of a whole forward+backward step and a columnar
pinned-host staging buffer (the B200-side replacement of the reference's per-key H2D loop,
rec_pangu/model_pipeline.py:47-58: `for key in data: data[key] = data[key].to(device)`; `model(data)`;
`loss.backward()`; `model.zero_grad()`)."""
from typing import Callable, Dict, List, Optional

import torch

from .models.utils import sparse_feature_names, dense_feature_names


class ColumnarBatch:
    """One batch as three stacked tensors — idx [F, B] int64, dense [Nd, B] fp32, labels [T, B] fp32 — instead of
    F+Nd+T separate ones, so a host batch crosses PCIe in 3 copies and the device dict is made of row views."""

    def __init__(self, enc_dict: Dict, batch_size: int, label_names=('label',), device='cuda', pinned_host=True):
        self.sparse = sparse_feature_names(enc_dict)
        self.dense = dense_feature_names(enc_dict)
        self.labels = list(label_names)
        F, Nd, T, B = len(self.sparse), len(self.dense), len(self.labels), batch_size
        self.idx = torch.zeros((F, B), dtype=torch.int64, device=device)
        self.dns = torch.zeros((max(Nd, 1), B), dtype=torch.float32, device=device)
        self.lab = torch.zeros((max(T, 1), B), dtype=torch.float32, device=device)
        self.h_idx = self.h_dns = self.h_lab = None
        if pinned_host:
            self.h_idx = torch.zeros((F, B), dtype=torch.int64).pin_memory()
            self.h_dns = torch.zeros((max(Nd, 1), B), dtype=torch.float32).pin_memory()
            self.h_lab = torch.zeros((max(T, 1), B), dtype=torch.float32).pin_memory()

    def as_dict(self) -> Dict[str, torch.Tensor]:
        d = {c: self.idx[i] for i, c in enumerate(self.sparse)}
        d.update({c: self.dns[i] for i, c in enumerate(self.dense)})
        d.update({c: self.lab[i] for i, c in enumerate(self.labels)})
        return d

    def fill_host(self, data: Dict[str, torch.Tensor]):
        for i, c in enumerate(self.sparse):
            self.h_idx[i].copy_(data[c].long())
        for i, c in enumerate(self.dense):
            self.h_dns[i].copy_(data[c])
        for i, c in enumerate(self.labels):
            self.h_lab[i].copy_(data[c])

    def h2d(self):
        """3 async copies from pinned memory on the current stream; returns the bytes moved."""
        self.idx.copy_(self.h_idx, non_blocking=True)
        self.dns.copy_(self.h_dns, non_blocking=True)
        self.lab.copy_(self.h_lab, non_blocking=True)
        return self.h_idx.numel() * 8 + self.h_dns.numel() * 4 + self.h_lab.numel() * 4

    def load_device(self, data: Dict[str, torch.Tensor]):
        for i, c in enumerate(self.sparse):
            self.idx[i].copy_(data[c])
        for i, c in enumerate(self.dense):
            self.dns[i].copy_(data[c])
        for i, c in enumerate(self.labels):
            self.lab[i].copy_(data[c])


class GraphedStep:
    """Captures `out = model(batch); out['loss'].backward(); post(); model.zero_grad()` into one CUDA graph on
    static input buffers.  `post` (e.g. an optimizer step or a gradient collective) runs between backward and
    zero_grad.  ``replay()`` re-launches the whole step with one cudaGraphLaunch; ``loss`` / ``pred`` are the static
    output tensors.

    ``zero_first``: the same K steps with the loop boundary where the reference's loop has it (`optimizer.zero_grad()` opens
    the iteration, model_pipeline.py:44): every step starts with the `model.zero_grad()` of the step before it, issued on a
    side stream AFTER the forward kernel has been launched and as a few small blocks per SM (`rows_zero_blocks`), so that the
    sparse re-zero of the table gradients (random 64-byte stores, DRAM-bound) runs next to the one-CTA-per-SM forward kernel
    (latency-bound, 20 % of the HBM roofline) instead of after backward; joined before backward.  The gradients of the last
    step stay in `.grad` after a replay."""

    _zero_streams = {}

    def __init__(self, model: torch.nn.Module, batch: ColumnarBatch, post: Optional[Callable[[], None]] = None,
                 warmup: int = 3, use_graph: bool = True, loss_scale: float = 1.0, zero_first: bool = False,
                 zero_blocks: int = 148, defer_capture: bool = False):
        self.model, self.batch, self.post = model, batch, post
        self.zero_first = zero_first
        self.zero_blocks = int(zero_blocks)
        self.loss_scale = loss_scale           # data parallel: back-propagate loss / world (dist.DenseGradBucket)
        self.data = batch.as_dict()
        self.graph = None
        self.loss = self.pred = None
        self.grads = None
        self.launches_per_step = 0
        from . import ops
        self._advance_epoch = False
        self._use_graph = use_graph
        d0 = ops.dropout_calls()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager(zero_first=False)
        # a step that draws dropout masks starts by advancing the device-side epoch the kernels mix into their seeds: the
        # host-drawn seeds are frozen by the capture, the epoch is not (same trick as FusedAdam's device step counter)
        self._advance_epoch = ops.dropout_calls() > d0
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if not defer_capture:
            if zero_first:
                raise ValueError('zero_first steps re-zero what the step BEFORE them touched: build them with GraphedStep.ring()')
            self._capture()

    def _capture(self):
        from . import ops
        n0 = ops.launch_count()
        if self._use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._eager()
            self.graph = g
        else:
            self._eager()
        self.launches_per_step = ops.launch_count() - n0

    @classmethod
    def ring(cls, model, batches, post=None, warmup: int = 3, use_graph: bool = True, loss_scale: float = 1.0,
             zero_first: bool = True, zero_blocks: int = 148):
        """One step per batch, to be replayed round-robin in this order.  With ``zero_first`` step i opens with the re-zero of
        the rows step i-1 touched (the last one for i = 0): the captures therefore follow one eager backward of the last
        batch, in ring order, each leaving its own touched-row list for the next capture to pick up."""
        steps = [cls(model, b, post=post, warmup=warmup, use_graph=use_graph, loss_scale=loss_scale, zero_first=zero_first,
                     zero_blocks=zero_blocks, defer_capture=True) for b in batches]
        if zero_first:
            last = steps[-1]
            out = model(last.data)
            (out['loss'] if loss_scale == 1.0 else out['loss'] * loss_scale).backward()
            if post is not None:
                post()
            torch.cuda.synchronize()
        for st in steps:
            st._capture()
        return steps

    def _eager(self, zero_first=None):
        zero_first = self.zero_first if zero_first is None else zero_first
        if self._advance_epoch:
            from . import ops
            ops.advance_dropout_epoch(self.batch.idx.device)
        join = None
        if zero_first:
            main = torch.cuda.current_stream()
            key = main.device.index or 0
            side = self._zero_streams.get(key)
            if side is None:
                side = self._zero_streams[key] = torch.cuda.Stream(device=main.device)
            fork, join = torch.cuda.Event(), torch.cuda.Event()
            fork.record(main)
        out = self.model(self.data)
        if join is not None:
            from . import _lib
            side.wait_event(fork)
            with torch.cuda.stream(side):
                _lib.check(_lib.load().rpb_set_option(b'rows_zero_blocks', self.zero_blocks), 'rpb_set_option(rows_zero_blocks)')
                try:
                    self.model.zero_grad(set_to_none=True)
                finally:
                    _lib.check(_lib.load().rpb_set_option(b'rows_zero_blocks', 0), 'rpb_set_option(rows_zero_blocks)')
                join.record(side)
            torch.cuda.current_stream().wait_event(join)
        (out['loss'] if self.loss_scale == 1.0 else out['loss'] * self.loss_scale).backward()
        if self.post is not None:
            self.post()
        if not zero_first:
            self.model.zero_grad(set_to_none=True)
        self.loss = out['loss'].detach()
        self.pred = out.get('pred', None)
        if zero_first:      # this step's gradient tensors (dense ones are re-created by every captured step: `.grad` follows the last CAPTURE)
            self.grads = {n: p.grad for n, p in self.model.named_parameters() if p.grad is not None}

    def replay(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._eager()
