"""Shared test helpers: golden-fixture loading and synthetic Criteo-shaped inputs."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    g = {'sd': {}, 'data': {}, 'out': {}, 'grad': {}}
    for k in z.files:
        top, rest = k.split('/', 1)
        if top == 'meta':
            g['meta'] = json.loads(bytes(z[k]).decode())
        else:
            g[top][rest] = torch.from_numpy(z[k].copy())
    return g


def load_layers():
    z = np.load(os.path.join(GOLDEN, 'layers.npz'))
    return {k: torch.from_numpy(z[k].copy()) for k in z.files}


def sub_sd(flat, tag):
    p = tag + '/sd/'
    return {k[len(p):]: v for k, v in flat.items() if k.startswith(p)}


def make_enc(n_sparse, n_dense, vocab):
    enc = {f'I{i + 1}': {'min': 0.0, 'max': 1.0} for i in range(n_dense)}
    vs = vocab if isinstance(vocab, (list, tuple)) else [vocab] * n_sparse
    enc.update({f'C{i + 1}': {'vocab_size': int(vs[i])} for i in range(n_sparse)})
    return enc


def make_batch(enc, B, seed=1029, labels=('label',), device='cpu'):
    gen = torch.Generator().manual_seed(seed)
    data = {}
    for c, d in enc.items():
        if 'vocab_size' in d:
            data[c] = torch.randint(0, d['vocab_size'] + 1, (B,), dtype=torch.int64, generator=gen)
        else:
            data[c] = torch.rand(B, generator=gen)
    for l in labels:
        data[l] = (torch.rand(B, generator=gen) < 0.25).float()
    return {k: v.to(device) for k, v in data.items()}


class capture_relu_inputs:
    """Records the input of every ReLU the oracle evaluates inside the `with` block (oracle/restatement.py uses F.relu only).
    Entries are CPU tensors whose first dimension is the batch."""

    def __enter__(self):
        import torch.nn.functional as F
        self._F, self._orig, self.inputs = F, F.relu, []

        def relu(x, *a, **k):
            self.inputs.append(x.detach())
            return self._orig(x, *a, **k)
        F.relu = relu
        return self

    def __exit__(self, *exc):
        self._F.relu = self._orig
        return False


def kink_adjacent_samples(relu_inputs, tau=5e-6):
    """Samples that have at least one ReLU pre-activation within `tau` of the kink.  The CUDA path reproduces every
    pre-activation to ~1e-6 or better (3xTF32 products, fp32 sums in another order), so only for these samples can the GPU take the other
    ReLU branch than the CPU oracle — which flips that ONE sample's whole gradient contribution on and off: an
    ill-conditioning of the reference's own formulation at that input, not an arithmetic error.  Everything else in the
    batch is smooth.  Returns a bool mask [B]."""
    mask = None
    for z in relu_inputs:
        m = (z.abs() < tau).reshape(z.shape[0], -1).any(dim=1)
        mask = m if mask is None else (mask | m)
    return mask


def drop_samples(data, keep):
    return {k: v[keep.to(v.device)] for k, v in data.items()}


def assert_close_rel(a, b, tol, what='', atol=1e-7, outlier_frac=0.0):
    """|a-b| <= tol * max|b| + atol element-wise, for batch-reduced gradients whose fp32 summation order (atomics /
    tiling) differs from the sequential CPU oracle.  `outlier_frac` tolerates a small fraction of elements beyond the
    bound: on large random batches a ReLU pre-activation (or a saturated sigmoid) within fp32 noise of its kink takes
    the other branch on the GPU than on the CPU, which changes that ONE sample's whole gradient contribution — an
    ill-conditioning of the reference's own formulation, not a kernel error."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    ref = b.abs().max().item()
    bad = ((a - b).abs() > tol * ref + atol)
    nbad = int(bad.sum().item())
    allowed = int(outlier_frac * a.numel())
    err = (a - b).abs().max().item()
    assert nbad <= allowed, (f'{what}: {nbad} of {a.numel()} elements beyond tol (allowed {allowed}); '
                             f'max abs err {err:.3e} vs max |ref| {ref:.3e} (tol {tol})')
