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

    def fit(self, model, train_loader, valid_loader: Optional = None, epoch: int = 10, lr: float = 1e-3,
            device: torch.device = torch.device('cpu'), use_earlystopping: bool = False, max_patience: int = 999,
            monitor_metric: Optional[str] = None, lr_scheduler_type: str = "", scheduler_params: Optional[dict] = {},
            optimizer_type: str = 'adam'):
        """Same arguments as the reference (trainer.py:51-61) plus `optimizer_type`: 'adam' = torch.optim.Adam over the dense
        gradients exactly as the reference (trainer.py:75), 'fused_adam' = rec_pangu_b200.optim.FusedAdam (row-sparse Adam
        on the touched table rows, dense parameters in one launch; lazy-Adam semantics on untouched rows)."""
        if self.use_wandb:
            import wandb
            wandb.init(**self.wandb_config)
        device = _compute_device(device)
        model = model.to(device)
        if optimizer_type == 'fused_adam':
            if lr_scheduler_type != "":
                raise ValueError('lr schedulers drive torch.optim optimizers; use optimizer_type="adam" with a scheduler')
            from .optim import FusedAdam
            optimizer = FusedAdam(model, lr=lr, betas=(0.9, 0.999), eps=1e-08)
        elif optimizer_type == 'adam':
            optimizer = torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-08, weight_decay=0)
        else:
            raise ValueError(f'Unknown optimizer_type: {optimizer_type}')
        if lr_scheduler_type == 'StepLR':
            scheduler = lr_scheduler.StepLR(optimizer, **scheduler_params)
        elif lr_scheduler_type == 'ExponentialLR':
            scheduler = lr_scheduler.ExponentialLR(optimizer, **scheduler_params)
        elif lr_scheduler_type == 'CosineAnnealingLR':
            scheduler = lr_scheduler.CosineAnnealingLR(optimizer, **scheduler_params)
        elif lr_scheduler_type == "":
            scheduler = None
        else:
            raise ValueError('Unknown scheduler type: {}'.format(lr_scheduler_type))
        logger.info('Model Starting Training ')
        best_epoch, best_metric, valid_metric = -1, -1, None
        for i in range(1, epoch + 1):
            train_metric = train_model(model, train_loader, optimizer=optimizer, device=device, num_task=self.num_task,
                                       use_wandb=self.use_wandb)
            if scheduler is not None:
                scheduler.step()
                logger.info(f"Epoch {i} LR:{round(scheduler.get_last_lr()[0], 6)}")
            logger.info(f"Train Metric:{train_metric}")
            if valid_loader is not None:
                valid_metric = test_model(model, valid_loader, device, num_task=self.num_task)
                self.save_train_model(model, self.model_ckpt_dir, f'e_{i}')
                if self.use_wandb:
                    import wandb
                    wandb.log(valid_metric)
                if use_earlystopping:
                    assert monitor_metric in valid_metric.keys(), f'{monitor_metric} not in Valid Metric {valid_metric.keys()}'
                    if valid_metric[monitor_metric] > best_metric:
                        best_epoch, best_metric = i, valid_metric[monitor_metric]
                        self.save_train_model(model, self.model_ckpt_dir, 'best')
                    if i - best_epoch >= max_patience:
                        logger.info(f"EarlyStopping at the Epoch {i} Valid Metric:{valid_metric}")
                        break
                logger.info(f"Valid Metric:{valid_metric}")
        if self.use_wandb:
            import wandb
            wandb.finish()
        return valid_metric

    @staticmethod
    def _save(obj, model_ckpt_dir, name):
        os.makedirs(model_ckpt_dir, exist_ok=True, mode=0o777)
        torch.save(obj, os.path.join(model_ckpt_dir, name))
        logger.info(f'Model Saved to {model_ckpt_dir}')

    def save_model(self, model, model_ckpt_dir: str):
        self._save({'model': model.state_dict()}, model_ckpt_dir, 'model.pth')

    def save_all(self, model, enc_dict: dict, model_ckpt_dir: str):
        self._save({'model': model.state_dict(), 'enc_dict': enc_dict}, model_ckpt_dir, 'model.pth')

    def save_train_model(self, model, model_ckpt_dir: str, model_str: str):
        self._save({'model': model.state_dict()}, model_ckpt_dir, f'model_{model_str}.pth')

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
