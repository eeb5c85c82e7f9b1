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

    ``zero_first`` (experimental, not yet measured): the same K steps with the loop boundary moved — every step starts with
    the `model.zero_grad()` of the step before it, issued on a side stream so that the sparse re-zero of the table
    gradients (random 64-byte stores, DRAM-bound) overlaps the forward kernel (shared-memory-pipe-bound) and is joined
    before backward; the gradients of the last step stay in `.grad` after a replay."""

    _zero_streams = {}

    def __init__(self, model: torch.nn.Module, batch: ColumnarBatch, post: Optional[Callable[[], None]] = None,
                 warmup: int = 3, use_graph: bool = True, loss_scale: float = 1.0, zero_first: bool = False):
        self.model, self.batch, self.post = model, batch, post
        self.zero_first = zero_first
        self.loss_scale = loss_scale           # data parallel: back-propagate loss / world (dist.DenseGradBucket)
        self.data = batch.as_dict()
        self.graph = None
        self.loss = self.pred = None
        self.launches_per_step = 0
        from . import ops
        self._advance_epoch = False
        d0 = ops.dropout_calls()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager()
        # a step that draws dropout masks starts by advancing the device-side epoch the kernels mix into their seeds: the
        # host-drawn seeds are frozen by the capture, the epoch is not (same trick as FusedAdam's device step counter)
        self._advance_epoch = ops.dropout_calls() > d0
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = ops.launch_count()
        if use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._eager()
            self.graph = g
        else:
            self._eager()
        self.launches_per_step = ops.launch_count() - n0

    def _eager(self):
        if self._advance_epoch:
            from . import ops
            ops.advance_dropout_epoch(self.batch.idx.device)
        join = None
        if self.zero_first:
            main = torch.cuda.current_stream()
            key = main.device.index or 0
            side = self._zero_streams.get(key)
            if side is None:
                side = self._zero_streams[key] = torch.cuda.Stream(device=main.device)
            fork, join = torch.cuda.Event(), torch.cuda.Event()
            fork.record(main)
            side.wait_event(fork)
            with torch.cuda.stream(side):
                self.model.zero_grad(set_to_none=True)
                join.record(side)
        out = self.model(self.data)
        if join is not None:
            torch.cuda.current_stream().wait_event(join)
        (out['loss'] if self.loss_scale == 1.0 else out['loss'] * self.loss_scale).backward()
        if self.post is not None:
            self.post()
        if not self.zero_first:
            self.model.zero_grad(set_to_none=True)
        self.loss = out['loss'].detach()
        self.pred = out.get('pred', None)

    def replay(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._eager()
