"""torchrun diagnostic: per-kernel device time of the N-GPU DeepFM step (config-2 shape, tables row-sharded in NVLink peer
memory): torch.profiler (CUPTI) around eager steps on every rank, kernel durations averaged per step, MAX over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/mp_profile_sharded.py

Prints a markdown table on rank 0 (kept as profiles/r02_scale<N>_kernels.md).  Device times are per kernel; the step replayed
as a CUDA graph (bench.py) is their sum minus overlap plus launch gaps."""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity

import bench
from rec_pangu_b200 import dist as rdist


def main():
    rank, world, local = rdist.init_from_env('nccl')
    dev = torch.device('cuda', local)
    w = bench.WORKLOADS[os.environ.get('WORKLOAD', 'deepfm')]
    model, enc, st = bench.build_model(w, dev, world)
    torch.cuda.empty_cache()
    bucket = rdist.DenseGradBucket([p for n, p in model.named_parameters() if 'embedding_layer.' not in n])
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    labels = bench.label_names(w)
    data = bench.synth_batch(enc, w['B'], gen, device=dev, labels=labels)

    def step():
        out = model(data)
        (out['loss'] / world).backward()
        bucket.all_reduce()
        model.zero_grad()

    for _ in range(3):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    n_steps = 4
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n_steps):
            step()
        torch.cuda.synchronize()
    per = defaultdict(float)
    cnt = defaultdict(int)
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.split('(')[0].replace('void ', '').replace('rpb::', '')[:70]
            per[name] += ev.device_time_total if hasattr(ev, 'device_time_total') else ev.cuda_time_total
            cnt[name] += 1
    names = sorted(per, key=lambda k: -per[k])[:24]
    # same kernel set on every rank (same code path): max over ranks of the per-step time
    vals = torch.tensor([per[n] / n_steps for n in names], device=dev, dtype=torch.float64)
    obj = [names]
    dist.broadcast_object_list(obj, src=0)
    vals = torch.tensor([per.get(n, 0.0) / n_steps for n in obj[0]], device=dev, dtype=torch.float64)
    vmax = vals.clone()
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f'SHARDED_PROFILE world={world} workload={w["model"]} batch_per_gpu={w["B"]}')
        print('| kernel | launches / step | us / step (rank 0) | us / step (max over ranks) |')
        print('|---|---|---|---|')
        for n, a, b in zip(obj[0], vals.tolist(), vmax.tolist()):
            print(f'| `{n}` | {cnt[n] / n_steps:.1f} | {a:.1f} | {b:.1f} |')
        print(f'| sum of the listed kernels | | {sum(vals.tolist()):.1f} | {sum(vmax.tolist()):.1f} |', flush=True)
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == '__main__':
    main()
