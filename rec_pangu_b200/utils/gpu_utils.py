import torch


def get_gpu_usage(device=None) -> str:
    """'reserved G/total G' string (reference: rec_pangu/utils/gpu_utils.py:7-19)."""
    reserved = torch.cuda.max_memory_reserved(device) / 1024 ** 3
    total = torch.cuda.get_device_properties(device).total_memory / 1024 ** 3
    return '{:.2f} G/{:.2f} G'.format(reserved, total)


def set_device(device_id: int = -1) -> torch.device:
    """reference: rec_pangu/utils/gpu_utils.py:22-48."""
    if not isinstance(device_id, int):
        raise TypeError("Device ID should be an integer.")
    return torch.device('cpu') if device_id < 0 else torch.device(f'cuda:{device_id}')
