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


def _to_device(data, device):
    for key in data.keys():
        data[key] = data[key].to(device, non_blocking=True)
    return data


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
            labels[i].append(data[lk].detach().reshape(-1))
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
                labels[i].append(data[lk].detach().reshape(-1))
    res = {}
    for i in range(num_task):
        y = torch.cat(labels[i]).cpu().numpy()
        p = torch.cat(preds[i]).cpu().numpy()
        for metric in metric_list:
            assert metric in _METRICS, f"Unsupported metric: {metric}. Supported metrics are {_METRICS}."
            key = metric if num_task == 1 else f'test_task{i + 1}_{metric}'
            res[key] = _metric(metric, y, p)
    return res
