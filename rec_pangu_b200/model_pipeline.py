"""train_model / test_model — the per-batch hot loop the model sits behind (reference:
rec_pangu/model_pipeline.py:17-219).  Signatures and returned metric-dict keys are the reference's; what happens per batch
is the reference's sequence — move the batch to the device, model(data), loss.backward(), optimizer.step(),
model.zero_grad() — run the B200 way:

* a host batch dict crosses PCIe as ONE copy per dtype (`_BatchStager`) instead of one per column, on a copy stream that
  runs one batch ahead of the compute stream;
* when the optimizer is graph-safe (`rec_pangu_b200.optim.FusedAdam`: step counter on the device) and the batch shape
  repeats, the whole step is captured ONCE per staging buffer as a CUDA graph (`_GraphedTrainStep`) and replayed — the
  eager path pays ~120 us of Python + ctypes per kernel launch, the replay one cudaGraphLaunch;
* predictions and labels stay on the device until the epoch ends, the running AUC is computed only when it is logged, and
  the embedding-index error record (ops.check_index_errors) is read at the points that synchronise anyway.
"""
import os
import time
from typing import Dict, List, Optional

import numpy as np
import torch

try:                                           # optional, as in the reference
    from loguru import logger
except Exception:                              # pragma: no cover
    import logging
    logger = logging.getLogger('rec_pangu_b200')

_METRICS = ['roc_auc_score', 'log_loss']


def _metric(name, y, p):
    from sklearn.metrics import roc_auc_score, log_loss
    y, p = np.asarray(y), np.asarray(p)
    if name == 'log_loss':
        return round(float(log_loss(y, np.clip(p, 1e-7, 1 - 1e-7))), 4)      # reference: log_loss(..., eps=1e-7)
    return round(float(roc_auc_score(y, p)), 4)


def _common_pinned_base(cols: List[torch.Tensor]) -> Optional[torch.Tensor]:
    """If the columns are, in order, the rows of ONE contiguous pinned [n, B] host tensor (a columnar loader hands out row
    views of its staging buffer), return that tensor: it can cross PCIe as it is, nothing to pack."""
    base = cols[0]._base
    if base is None or base.dim() != 2 or not base.is_contiguous() or base.shape[0] != len(cols) or not base.is_pinned():
        return None
    step = base.stride(0) * base.element_size()
    p0 = base.data_ptr()
    for i, c in enumerate(cols):
        if c._base is not base or c.data_ptr() != p0 + i * step or not c.is_contiguous():
            return None
    return base


class _BatchStager:
    """Host batch dict -> device in ONE async copy per dtype instead of one per column (the reference moves every key of
    the batch dict separately, model_pipeline.py:47-50: ~40 small H2D copies per step at the Criteo shape).  Columns of equal
    dtype and length are packed as rows of a pinned [n_cols, B] staging buffer (skipped when they already are the rows of one
    pinned tensor), copied on `stream`, and handed out as row views of a device buffer (the kernels take per-column
    pointers, so nothing is re-packed).  `n_buf` device/host buffer sets alternate; `ready[t]` is recorded after the copies
    of set t, `done[t]` must be recorded by the consumer after the last kernel that reads set t."""

    def __init__(self, device: torch.device, n_buf: int = 2):
        self.device, self.n_buf = device, n_buf
        self.host, self.dev = {}, {}
        self.ready = [torch.cuda.Event() for _ in range(n_buf)]
        self.done = [None] * n_buf
        self.host_free = [None] * n_buf
        self.turn = -1

    @staticmethod
    def stageable(data) -> bool:
        return all(isinstance(v, torch.Tensor) and not v.is_cuda and v.dim() == 1 for v in data.values())

    def signature(self, data):
        return tuple((k, v.dtype, v.shape[0]) for k, v in data.items())

    def stage(self, data: Dict[str, torch.Tensor], stream: torch.cuda.Stream):
        """Returns (set index t, dict of device row views).  The copies run on `stream`."""
        self.turn = (self.turn + 1) % self.n_buf
        t = self.turn
        groups = {}
        for k, v in data.items():
            groups.setdefault((v.dtype, v.shape[0]), []).append(k)
        if self.done[t] is not None:
            stream.wait_event(self.done[t])                  # the step that last read device set t has finished
        out = {}
        with torch.cuda.stream(stream):
            for (dtype, n), keys in groups.items():
                sig = (dtype, n, len(keys))
                if sig not in self.dev:
                    self.dev[sig] = [torch.empty((len(keys), n), dtype=dtype, device=self.device) for _ in range(self.n_buf)]
                d = self.dev[sig][t]
                cols = [data[k] for k in keys]
                src = _common_pinned_base(cols)
                if src is None:
                    if sig not in self.host:
                        self.host[sig] = [torch.empty((len(keys), n), dtype=dtype).pin_memory() for _ in range(self.n_buf)]
                    src = self.host[sig][t]
                    if self.host_free[t] is not None:
                        self.host_free[t].synchronize()      # the DMA that last read this pinned buffer has finished
                    for i, c in enumerate(cols):
                        src[i].copy_(c)
                d.copy_(src, non_blocking=True)
                for i, k in enumerate(keys):
                    out[k] = d[i]
            self.ready[t].record(stream)
            self.host_free[t] = self.ready[t]
        return t, {k: out[k] for k in data.keys()}


def _graph_safe(optimizer) -> bool:
    return bool(getattr(optimizer, 'graph_safe', False)) and os.environ.get('RPB_TRAIN_GRAPH', '1') != '0'


class _GraphedTrainStep:
    """`out = model(data); out['loss'].backward(); optimizer.step(); model.zero_grad()` on ONE staging-buffer set, captured
    as a CUDA graph after `warm` eager runs (workspaces grow, lazy state is created) and replayed afterwards."""

    def __init__(self, model, optimizer, data_views: Dict[str, torch.Tensor], warm: int = 2):
        self.model, self.optimizer, self.data = model, optimizer, data_views
        self.calls, self.warm, self.graph, self.out = 0, warm, None, None
        self.failed = False
        from . import ops
        self._ops = ops
        self._advance_epoch = False

    def _eager(self):
        if self._advance_epoch:
            self._ops.advance_dropout_epoch(next(iter(self.data.values())).device)
        out = self.model(self.data)
        out['loss'].backward()
        self.optimizer.step()
        self.model.zero_grad()
        return {k: v.detach() for k, v in out.items()}

    def run(self):
        self.calls += 1
        if self.graph is not None:
            self.graph.replay()
            return self.out
        if self.calls <= self.warm or self.failed:
            d0 = self._ops.dropout_calls()
            out = self._eager()
            self._advance_epoch = self._advance_epoch or self._ops.dropout_calls() > d0
            return out
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.out = self._eager()
            self.graph = g
            return self.out                                   # capture does not execute: the caller replays via run() next time
        except Exception as ex:                               # capture refused (an op that syncs, ...): stay eager, say so once
            self.failed = True
            logger.info(f'train_model: CUDA-graph capture of the step failed ({ex!r}); running eager launches')
            torch.cuda.synchronize()
            return self._eager()


def _batch_outputs(output, data, num_task, preds, labels):
    for i in range(num_task):
        pk, lk = ('pred', 'label') if num_task == 1 else (f'task{i + 1}_pred', f'task{i + 1}_label')
        preds[i].append(output[pk].detach().reshape(-1).clone())          # static graph outputs / staging buffers are reused
        labels[i].append(data[lk].detach().reshape(-1).clone())


def _tail(chunks: List[torch.Tensor], n: int) -> torch.Tensor:
    """Last n elements of the concatenation without concatenating everything (the reference cats all batches every log round)."""
    got, take = 0, []
    for c in reversed(chunks):
        take.append(c)
        got += c.shape[0]
        if got >= n:
            break
    return torch.cat(list(reversed(take)))[-n:]


def train_model(model: torch.nn.Module, train_loader, optimizer, device: torch.device,
                metric_list: List[str] = ['roc_auc_score', 'log_loss'], num_task: int = 1, use_wandb: bool = False,
                log_rounds: int = 100) -> dict:
    from . import ops
    device = torch.device(device)
    model.train()
    max_iter = int(len(train_loader.dataset) / train_loader.batch_size)
    preds = [[] for _ in range(num_task)]
    labels = [[] for _ in range(num_task)]
    start_time = time.time()
    on_gpu = device.type == 'cuda'
    # staging buffers, copy stream and captured steps live as long as the optimizer (= the training run): epoch 2 replays the
    # graphs epoch 1 captured
    state = getattr(optimizer, '_rpb_train_state', None)
    if on_gpu and (state is None or state['model'] is not model or state['device'] != device):
        state = {'model': model, 'device': device, 'stager': _BatchStager(device), 'copy_stream': torch.cuda.Stream(device=device),
                 'graphed': {}}
        try:
            optimizer._rpb_train_state = state
        except Exception:                              # an optimizer that refuses attributes: per-call state
            pass
    stager = state['stager'] if on_gpu else None
    copy_stream = state['copy_stream'] if on_gpu else None
    graphed: Dict[tuple, _GraphedTrainStep] = state['graphed'] if on_gpu else {}
    use_graph = on_gpu and _graph_safe(optimizer)

    def staged(it):
        """Yields (set index | None, device batch): batch i+1 is staged (packed + H2D on the copy stream) while step i runs."""
        for data in it:
            if on_gpu and stager.stageable(data):
                yield stager.stage(data, copy_stream)
            else:
                yield None, {k: v.to(device, non_blocking=True) for k, v in data.items()}

    main = torch.cuda.current_stream(device) if on_gpu else None
    nxt = None
    it = staged(iter(train_loader))
    idx = -1
    while True:
        cur = nxt if nxt is not None else next(it, None)
        if cur is None:
            break
        idx += 1
        nxt = next(it, None)                          # enqueue the copies of the next batch before this step is launched
        t, data = cur
        if t is not None:
            main.wait_event(stager.ready[t])
        if use_graph and t is not None:
            key = (t, tuple((k, v.dtype, v.shape[0]) for k, v in data.items()))
            step = graphed.get(key)
            if step is None:
                step = graphed[key] = _GraphedTrainStep(model, optimizer, data)
            output = step.run()
            if step.graph is not None and step.calls == step.warm + 1:
                output = step.run()                    # the capture itself executed nothing: replay it for this batch
        else:
            output = model(data)
            loss = output['loss']
            loss.backward()
            optimizer.step()
            model.zero_grad()
        _batch_outputs(output, data, num_task, preds, labels)
        if t is not None:
            ev = torch.cuda.Event()
            ev.record(main)
            stager.done[t] = ev
        if use_wandb:
            import wandb
            wandb.log({'train_loss': output['loss'].item()})
        if idx % log_rounds == 0:
            iter_time = time.time() - start_time
            remaining = round(((iter_time / (idx + 1)) * (max_iter - idx + 1)) / 60, 2)
            msg = f'Iter {idx}/{max_iter} Remaining time:{remaining} min Loss:{round(float(output["loss"].detach().cpu()), 4)}'
            if on_gpu:
                ops.check_index_errors(device, sync=False)        # the .cpu() above synchronised: the record is current
            if num_task == 1:
                y = _tail(labels[0], 1000).cpu().numpy()
                p = _tail(preds[0], 1000).cpu().numpy()
                if len(np.unique(y)) > 1:
                    msg += f' AUC:{_metric("roc_auc_score", y, p)}'
            if device.type != 'cpu':
                from .utils import get_gpu_usage
                msg += f' GPU Mem:{get_gpu_usage(device)}'
            logger.info(msg)
    if on_gpu:
        ops.check_index_errors(device, sync=True)                 # out-of-range ids raise IndexError as in the reference
    res = {}
    for i in range(num_task):
        if not metric_list or not preds[i]:
            continue
        y = torch.cat(labels[i]).cpu().numpy()
        p = torch.cat(preds[i]).cpu().numpy()
        for metric in metric_list:
            assert metric in _METRICS, 'metric :{} not supported! metric must be in {}'.format(metric, _METRICS)
            key = f'train_{metric}' if num_task == 1 else f'train_task{i + 1}_{metric}'
            res[key] = _metric(metric, y, p)
    return res


def test_model(model: torch.nn.Module, test_loader, device: torch.device,
               metric_list: List[str] = ['roc_auc_score', 'log_loss'], num_task: int = 1) -> dict:
    from . import ops
    device = torch.device(device)
    model.eval()
    preds = [[] for _ in range(num_task)]
    labels = [[] for _ in range(num_task)]
    on_gpu = device.type == 'cuda'
    stager = _BatchStager(device) if on_gpu else None
    copy_stream = torch.cuda.Stream(device=device) if on_gpu else None
    with torch.no_grad():
        for data in test_loader:
            t = None
            if on_gpu and stager.stageable(data):
                t, data = stager.stage(data, copy_stream)
                torch.cuda.current_stream(device).wait_event(stager.ready[t])
            else:
                data = {k: v.to(device) for k, v in data.items()}
            output = model(data)              # the reference evaluates with is_training=True (App. A-14): loss is computed
            _batch_outputs(output, data, num_task, preds, labels)
            if t is not None:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(device))
                stager.done[t] = ev
    if on_gpu:
        ops.check_index_errors(device, sync=True)
    res = {}
    for i in range(num_task):
        y = torch.cat(labels[i]).cpu().numpy()
        p = torch.cat(preds[i]).cpu().numpy()
        for metric in metric_list:
            assert metric in _METRICS, f"Unsupported metric: {metric}. Supported metrics are {_METRICS}."
            key = metric if num_task == 1 else f'test_task{i + 1}_{metric}'
            res[key] = _metric(metric, y, p)
    return res
