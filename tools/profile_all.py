"""Run two eager training steps of one model at its BASELINE.json config shape (for ncu kernel captures)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', required=True)
    ap.add_argument('--steps', type=int, default=2)
    a = ap.parse_args()
    from rec_pangu_b200.models import ranking, multi_task
    F, Nd, V = 26, 13, 1_000_000
    cfg = {
        'DeepFM': (ranking.DeepFM, dict(embedding_dim=16), 65536),
        'xDeepFM': (ranking.xDeepFM, dict(embedding_dim=16), 65536),
        'AutoInt': (ranking.AutoInt, dict(embedding_dim=32, num_heads=3), 32768),
        'DCN': (ranking.DCN, dict(embedding_dim=16), 65536),
        'FiBiNet': (ranking.FiBiNet, dict(embedding_dim=16), 16384),
        'MMOE': (multi_task.MMOE, dict(embedding_dim=40, device='cuda'), 32768),
    }[a.model]
    enc = {f'I{i + 1}': {'min': 0.0, 'max': 1.0} for i in range(Nd)}
    enc.update({f'C{i + 1}': {'vocab_size': V} for i in range(F)})
    torch.manual_seed(0)
    with torch.device('cuda'):
        model = cfg[0](enc_dict=enc, **cfg[1])
    with torch.no_grad():
        for n, p in model.named_parameters():
            if 'embedding_layer' in n:
                p.mul_(0.2)
    model.train()
    if hasattr(model, 'set_grad_mode'):
        model.set_grad_mode('persistent')
    B = cfg[2]
    g = torch.Generator(device='cuda').manual_seed(1)
    data = {c: torch.randint(0, V + 1, (B,), device='cuda', generator=g) for c in enc if 'vocab_size' in enc[c]}
    data.update({c: torch.rand(B, device='cuda', generator=g) for c in enc if 'min' in enc[c]})
    for k in ('label', 'task1_label', 'task2_label'):
        data[k] = (torch.rand(B, device='cuda', generator=g) < 0.25).float()
    for _ in range(a.steps):
        out = model(data)
        out['loss'].backward()
        model.zero_grad()
    torch.cuda.synchronize()
    print(a.model, 'ok loss', float(out['loss']))


if __name__ == '__main__':
    main()
