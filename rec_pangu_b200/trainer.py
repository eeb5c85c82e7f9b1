"""RankTrainer — reference: rec_pangu/trainer.py:23-236 (same constructor, fit/evaluate/save/predict methods,
Adam(lr, betas=(0.9,0.999), eps=1e-8) as trainer.py:75, same checkpoint dict layout {'model'[, 'enc_dict']})."""
import os
from typing import Optional

import torch
import torch.utils.data as D
from torch.optim import lr_scheduler

from .dataset import BaseDataset, MultiTaskDataset
from .model_pipeline import train_model, test_model, logger
from .utils import beautify_json


def _compute_device(device) -> torch.device:
    """The reference's examples pass device=torch.device('cpu') (examples/ranking/run_ranking_example.py:32).  This
    build has no CPU compute path: a 'cpu' request is redirected to the current CUDA device (the data loader keeps
    producing host tensors; train_model moves each batch), and without a GPU it fails loudly."""
    device = torch.device(device)
    if device.type == 'cuda':
        return device
    if not torch.cuda.is_available():
        raise RuntimeError('rec_pangu_b200 computes on CUDA (sm_100a) only and no GPU is visible; there is no CPU fallback')
    logger.info(f'device={device} requested; rec_pangu_b200 computes on cuda:{torch.cuda.current_device()}')
    return torch.device('cuda', torch.cuda.current_device())


class RankTrainer:
    def __init__(self, num_task: int = 1, wandb_config: dict = None, model_ckpt_dir: str = './model_ckpt'):
        self.num_task = num_task
        self.wandb_config = wandb_config
        self.model_ckpt_dir = model_ckpt_dir
        self.use_wandb = self.wandb_config is not None
        if self.use_wandb:
            import wandb
            wandb.login(key=self.wandb_config['key'])
            self.wandb_config.pop('key')

    # ---- pieces of fit(): each returns what the epoch loop needs, so the loop itself stays a dozen lines
    @staticmethod
    def _make_optimizer(model, lr, optimizer_type, lr_scheduler_type):
        if optimizer_type == 'fused_adam':
            if lr_scheduler_type:
                raise ValueError('lr schedulers drive torch.optim optimizers; use optimizer_type="adam" with a scheduler')
            from .optim import FusedAdam
            return FusedAdam(model, lr=lr, betas=(0.9, 0.999), eps=1e-08)
        if optimizer_type == 'adam':               # the reference's optimizer, trainer.py:75
            return torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-08, weight_decay=0)
        raise ValueError(f'Unknown optimizer_type: {optimizer_type}')

    @staticmethod
    def _make_scheduler(optimizer, lr_scheduler_type, scheduler_params):
        if not lr_scheduler_type:
            return None
        known = {'StepLR': lr_scheduler.StepLR, 'ExponentialLR': lr_scheduler.ExponentialLR,
                 'CosineAnnealingLR': lr_scheduler.CosineAnnealingLR}
        if lr_scheduler_type not in known:
            raise ValueError('Unknown scheduler type: {}'.format(lr_scheduler_type))
        return known[lr_scheduler_type](optimizer, **(scheduler_params or {}))

    class _EarlyStop:
        """Best-so-far bookkeeping of the monitored validation metric (larger is better, as in the reference)."""

        def __init__(self, enabled, metric, patience):
            self.enabled, self.metric, self.patience = enabled, metric, patience
            self.best_epoch, self.best = -1, -1

        def update(self, epoch, valid_metric):
            """Returns (improved, stop)."""
            if not self.enabled:
                return False, False
            assert self.metric in valid_metric.keys(), f'{self.metric} not in Valid Metric {valid_metric.keys()}'
            improved = valid_metric[self.metric] > self.best
            if improved:
                self.best_epoch, self.best = epoch, valid_metric[self.metric]
            return improved, epoch - self.best_epoch >= self.patience

    def fit(self, model, train_loader, valid_loader: Optional = None, epoch: int = 10, lr: float = 1e-3,
            device: torch.device = torch.device('cpu'), use_earlystopping: bool = False, max_patience: int = 999,
            monitor_metric: Optional[str] = None, lr_scheduler_type: str = "", scheduler_params: Optional[dict] = {},
            optimizer_type: str = 'adam'):
        """Same arguments as the reference (trainer.py:51-61) plus `optimizer_type`: 'adam' = torch.optim.Adam over the dense
        gradients exactly as the reference (trainer.py:75), 'fused_adam' = rec_pangu_b200.optim.FusedAdam (row-sparse Adam
        on the touched table rows, dense parameters in one launch, the whole step replayed as a CUDA graph by train_model).
        Per epoch, as the reference: train, step the scheduler, validate, checkpoint `e_<i>` (and `best`), early-stop.
        Returns the last validation metric dict (None without a valid_loader; the reference raises UnboundLocalError there)."""
        wandb = None
        if self.use_wandb:
            import wandb
            wandb.init(**self.wandb_config)
        device = _compute_device(device)
        model = model.to(device)
        optimizer = self._make_optimizer(model, lr, optimizer_type, lr_scheduler_type)
        scheduler = self._make_scheduler(optimizer, lr_scheduler_type, scheduler_params)
        stopper = self._EarlyStop(use_earlystopping, monitor_metric, max_patience)
        valid_metric = None
        logger.info('Model Starting Training ')
        for i in range(1, epoch + 1):
            train_metric = train_model(model, train_loader, optimizer=optimizer, device=device, num_task=self.num_task,
                                       use_wandb=self.use_wandb)
            if scheduler is not None:
                scheduler.step()
                logger.info(f"Epoch {i} LR:{round(scheduler.get_last_lr()[0], 6)}")
            logger.info(f"Train Metric:{train_metric}")
            if valid_loader is None:
                continue
            valid_metric = test_model(model, valid_loader, device, num_task=self.num_task)
            self.save_train_model(model, self.model_ckpt_dir, f'e_{i}')
            if wandb is not None:
                wandb.log(valid_metric)
            improved, stop = stopper.update(i, valid_metric)
            if improved:
                self.save_train_model(model, self.model_ckpt_dir, 'best')
            if stop:
                logger.info(f"EarlyStopping at the Epoch {i} Valid Metric:{valid_metric}")
                break
            logger.info(f"Valid Metric:{valid_metric}")
        if wandb is not None:
            wandb.finish()
        return valid_metric

    @staticmethod
    def _save(obj, model_ckpt_dir, name):
        os.makedirs(model_ckpt_dir, exist_ok=True, mode=0o777)
        torch.save(obj, os.path.join(model_ckpt_dir, name))
        logger.info(f'Model Saved to {model_ckpt_dir}')

    @staticmethod
    def _state_dict(model):
        """state_dict in the REFERENCE's layout: row-sharded tables are all-gathered back to [vocab_size + 1, D] (collective:
        every rank calls it), so the file loads in the reference, on one GPU, or on another shard count."""
        emb = getattr(model, 'embedding_layer', None)
        if emb is not None and getattr(emb, '_shards', None) is not None:
            from . import dist as rdist
            return rdist.gather_state_dict(model)
        return model.state_dict()

    def _save_state(self, model, extra: dict, model_ckpt_dir: str, name: str):
        sd = self._state_dict(model)                   # collective when sharded; only rank 0 writes
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_rank() != 0:
            return
        self._save(dict({'model': sd}, **extra), model_ckpt_dir, name)

    def save_model(self, model, model_ckpt_dir: str):
        self._save_state(model, {}, model_ckpt_dir, 'model.pth')

    def save_all(self, model, enc_dict: dict, model_ckpt_dir: str):
        self._save_state(model, {'enc_dict': enc_dict}, model_ckpt_dir, 'model.pth')

    def save_train_model(self, model, model_ckpt_dir: str, model_str: str):
        self._save_state(model, {}, model_ckpt_dir, f'model_{model_str}.pth')

    def evaluate_model(self, model, test_loader, device: torch.device = torch.device('cpu')):
        test_metric = test_model(model, test_loader, _compute_device(device), num_task=self.num_task)
        logger.info(f"Test Metric:{beautify_json(test_metric)}")
        return test_metric

    def predict_dataloader(self, model, test_loader, device: torch.device = torch.device('cpu')):
        device = _compute_device(device)
        model.eval()
        outs = [[] for _ in range(self.num_task)]
        with torch.no_grad():
            for data in test_loader:
                for key in data.keys():
                    data[key] = data[key].to(device)
                output = model(data, is_training=False)
                for i in range(self.num_task):
                    k = 'pred' if self.num_task == 1 else f'task{i + 1}_pred'
                    outs[i].append(output[k].reshape(-1))
        if device.type == 'cuda':
            from . import ops
            ops.check_index_errors(device, sync=True)         # out-of-range ids raise IndexError as in the reference
        res = [list(torch.cat(o).cpu().numpy()) for o in outs]
        return res[0] if self.num_task == 1 else res

    def predict_dataframe(self, model, test_df, enc_dict: dict, schema: dict,
                          device: torch.device = torch.device('cpu'), batch_size: int = 1024):
        if schema['task_type'] == 'ranking':
            test_dataset = BaseDataset(schema, test_df, enc_dict=enc_dict)
        elif schema['task_type'] == 'multitask':
            test_dataset = MultiTaskDataset(schema, test_df, enc_dict=enc_dict)
        test_loader = D.DataLoader(test_dataset, batch_size=batch_size, shuffle=False, num_workers=0)
        return self.predict_dataloader(model, test_loader, device=device)
