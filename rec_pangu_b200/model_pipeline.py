"""train_model / test_model — the per-batch hot loop the model sits behind (reference:
rec_pangu/model_pipeline.py:17-219).  Signatures and returned metric-dict keys are the reference's; the loop body
keeps its order (H2D per key -> model(data) -> loss.backward() -> optimizer.step() -> model.zero_grad()) but
predictions/labels stay on the device until the epoch ends (one D2H instead of two syncs per iteration) and the
running AUC is computed only when it is logged."""
import time
from typing import List

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


class _BatchStager:
    """Host batch -> device in ONE copy per dtype instead of one per column (the reference moves every key of the batch dict
    separately, model_pipeline.py:47-50: ~40 small H2D copies per step at the Criteo shape).  Columns of equal dtype and
    length are packed as rows of a pinned [n_cols, B] staging buffer, copied with a single async cudaMemcpy, and handed to
    the model as row views of the device buffer (the kernels take per-column pointers, so nothing is re-packed).  Two
    device buffers alternate so the copy of batch i+1 never overwrites what the step of batch i is still reading."""

    def __init__(self):
        self.host, self.dev, self.done, self.turn = {}, {}, {}, 0

    def __call__(self, data, device):
        if device.type != 'cuda' or not all(isinstance(v, torch.Tensor) and not v.is_cuda and v.dim() == 1 for v in data.values()):
            for key in data.keys():
                data[key] = data[key].to(device, non_blocking=True)
            return data
        self.turn ^= 1
        t = self.turn
        groups = {}
        for k, v in data.items():
            groups.setdefault((v.dtype, v.shape[0]), []).append(k)
        out = {}
        for (dtype, n), keys in groups.items():
            sig = (dtype, n, len(keys))
            if sig not in self.host:
                self.host[sig] = [torch.empty((len(keys), n), dtype=dtype).pin_memory() for _ in range(2)]
                self.dev[sig] = [torch.empty((len(keys), n), dtype=dtype, device=device) for _ in range(2)]
                self.done[sig] = [None, None]
            h, d = self.host[sig][t], self.dev[sig][t]
            if self.done[sig][t] is not None:
                self.done[sig][t].synchronize()          # the DMA that last read this pinned buffer has finished
            for i, k in enumerate(keys):
                h[i].copy_(data[k])
            d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            self.done[sig][t] = ev
            for i, k in enumerate(keys):
                out[k] = d[i]
        return {k: out[k] for k in data.keys()}


_stager = _BatchStager()


def _to_device(data, device):
    return _stager(data, device)


def train_model(model: torch.nn.Module, train_loader, optimizer, device: torch.device,
                metric_list: List[str] = ['roc_auc_score', 'log_loss'], num_task: int = 1, use_wandb: bool = False,
                log_rounds: int = 100) -> dict:
    model.train()
    max_iter = int(len(train_loader.dataset) / train_loader.batch_size)
    preds = [[] for _ in range(num_task)]
    labels = [[] for _ in range(num_task)]
    start_time = time.time()
    for idx, data in enumerate(train_loader):
        data = _to_device(data, device)
        output = model(data)
        loss = output['loss']
        loss.backward()
        optimizer.step()
        model.zero_grad()
        for i in range(num_task):
            pk, lk = ('pred', 'label') if num_task == 1 else (f'task{i + 1}_pred', f'task{i + 1}_label')
            preds[i].append(output[pk].detach().reshape(-1))
            labels[i].append(data[lk].detach().reshape(-1).clone())      # the staging buffer is reused two batches later
        if use_wandb:
            import wandb
            wandb.log({'train_loss': loss.item()})
        if idx % log_rounds == 0:
            iter_time = time.time() - start_time
            remaining = round(((iter_time / (idx + 1)) * (max_iter - idx + 1)) / 60, 2)
            msg = f'Iter {idx}/{max_iter} Remaining time:{remaining} min Loss:{round(float(loss.detach().cpu()), 4)}'
            if num_task == 1:
                y = torch.cat(labels[0])[-1000:].cpu().numpy()
                p = torch.cat(preds[0])[-1000:].cpu().numpy()
                if len(np.unique(y)) > 1:
                    msg += f' AUC:{_metric("roc_auc_score", y, p)}'
            if device.type != 'cpu':
                from .utils import get_gpu_usage
                msg += f' GPU Mem:{get_gpu_usage(device)}'
            logger.info(msg)
    res = {}
    for i in range(num_task):
        y = torch.cat(labels[i]).cpu().numpy()
        p = torch.cat(preds[i]).cpu().numpy()
        for metric in metric_list:
            assert metric in _METRICS, 'metric :{} not supported! metric must be in {}'.format(metric, _METRICS)
            key = f'train_{metric}' if num_task == 1 else f'train_task{i + 1}_{metric}'
            res[key] = _metric(metric, y, p)
    return res


def test_model(model: torch.nn.Module, test_loader, device: torch.device,
               metric_list: List[str] = ['roc_auc_score', 'log_loss'], num_task: int = 1) -> dict:
    model.eval()
    preds = [[] for _ in range(num_task)]
    labels = [[] for _ in range(num_task)]
    with torch.no_grad():
        for data in test_loader:
            data = _to_device(data, device)
            output = model(data)              # the reference evaluates with is_training=True (App. A-14): loss is computed
            for i in range(num_task):
                pk, lk = ('pred', 'label') if num_task == 1 else (f'task{i + 1}_pred', f'task{i + 1}_label')
                preds[i].append(output[pk].detach().reshape(-1))
                labels[i].append(data[lk].detach().reshape(-1).clone())      # the staging buffer is reused two batches later
    res = {}
    for i in range(num_task):
        y = torch.cat(labels[i]).cpu().numpy()
        p = torch.cat(preds[i]).cpu().numpy()
        for metric in metric_list:
            assert metric in _METRICS, f"Unsupported metric: {metric}. Supported metrics are {_METRICS}."
            key = metric if num_task == 1 else f'test_task{i + 1}_{metric}'
            res[key] = _metric(metric, y, p)
    return res
