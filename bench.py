#!/usr/bin/env python
"""bench.py — hot-path throughput on B200 (BASELINE.json metric: "DeepFM samples/sec at 1/2/4/8 B200; embedding-gather
HBM GB/s vs peak").

Default workload = BASELINE.json configs[1]: DeepFM, synthetic Criteo shape — 26 sparse fields x 1M-row tables (D=16),
13 dense fields, hidden [64,64,64], batch 65536 per GPU (weak scaling), random-init weights, uniform ids.
One "step" = one pass of the hot path over one batch: forward (gather -> FM -> MLP -> sigmoid/BCE) + backward (MLP grads,
FM grad, scatter-add into the per-table gradient buffers) + sparse re-zero of those buffers (`model.zero_grad()`), i.e.
rec_pangu/model_pipeline.py:52-58 without optimizer.step (the optimizer is a "next" row; `train_step` /
`train_model` report it beside the headline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload deepfm|xdeepfm|autoint|mmoe_cfg5]

`--workload` selects the other BASELINE.json configs (xdeepfm = configs[2], autoint = configs[3], mmoe_cfg5 = configs[4]).
N > 1 is launched by torchrun (one rank per GPU, NCCL).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

T_START = time.time()          # before `import torch`: a box that is slow to start must not also pay for the optional legs

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

SEED = 1029

# ----------------------------------------------------------------------------------------------- workloads
# F sparse fields x V-row tables (V+1 rows with the OOV slot) of width D, Nd dense fields, batch B per GPU.
WORKLOADS = {
    'deepfm': dict(config=1, model='DeepFM', kind='ranking', F=26, Nd=13, V=1_000_000, D=16, B=65536,
                   kw=dict(hidden_units=[64, 64, 64]), okw=dict(hidden_units=(64, 64, 64)),
                   text='DeepFM criteo-shape: 26 sparse x 1M-row tables (D=16) + 13 dense, MLP 64-64-64, fwd + bwd (per-table '
                        'grads, sparse re-zero), batch 65536 per GPU, BASELINE.json configs[1]'),
    'xdeepfm': dict(config=2, model='xDeepFM', kind='ranking', F=26, Nd=13, V=1_000_000, D=16, B=65536, kw={}, okw={},
                    text='xDeepFM criteo-shape: 26 sparse x 1M-row tables (D=16) + LR tables (D=1) + 13 dense, CIN 16-16-16, '
                         'MLP 64-64-64 (dropout 0.1), fwd + bwd, batch 65536 per GPU, BASELINE.json configs[2]'),
    'autoint': dict(config=3, model='AutoInt', kind='ranking', F=26, Nd=13, V=1_000_000, D=32, B=32768,
                    kw=dict(num_heads=3), okw=dict(num_heads=3),
                    text='AutoInt criteo-shape: 26 sparse x 1M-row tables (D=32) + LR tables + 13 dense, 1 interacting layer x 3 '
                         'heads (d=8), MLP 64-64-64 (dropout 0.1), fwd + bwd, batch 32768 per GPU, BASELINE.json configs[3]'),
    'mmoe_cfg5': dict(config=4, model='MMOE', kind='multitask', F=8, Nd=13, V=100_000_000, D=40, B=65536, kw={}, okw={},
                      text='MMOE 2-task: 8 hashed 100M-row tables (D=40, 16 GB each, row-sharded over the GPUs) + 13 dense, 3 experts '
                           'x 128, towers 128-64, fwd + bwd, batch 65536 per GPU, BASELINE.json configs[4]'),
}


def alg_bytes_per_sample(w):
    """F*(8 + 4*D) + 4*Nd + 4 — ids + one read of each gathered row + dense features + one fp32 output
    (+ 4*F for the D=1 LR rows of the models that have an LR_Layer)."""
    lr = 4 * w['F'] if w['model'] in ('xDeepFM', 'AutoInt') else 0
    return w['F'] * (8 + 4 * w['D']) + 4 * w['Nd'] + 4 + lr


def make_enc(w):
    enc = {f'I{i + 1}': {'min': 0.0, 'max': 1.0} for i in range(w['Nd'])}
    enc.update({f'C{i + 1}': {'vocab_size': w['V']} for i in range(w['F'])})
    return enc


def label_names(w):
    return ('label',) if w['kind'] == 'ranking' else ('task1_label', 'task2_label')


def synth_batch(enc, B, gen, device='cpu', labels=('label',)):
    d = {}
    for c, m in enc.items():
        if 'vocab_size' in m:
            d[c] = torch.randint(0, m['vocab_size'] + 1, (B,), dtype=torch.int64, generator=gen, device=device)
        else:
            d[c] = torch.rand(B, generator=gen, device=device)
    for l in labels:
        d[l] = (torch.rand(B, generator=gen, device=device) < 0.25).float()
    return d


def make_config(w, world):
    """The `config` object both arms print (the driver compares them): what the workload is, nothing about how it ran."""
    return {'workload': w['text'], 'batch_per_gpu': w['B'], 'global_batch': w['B'] * world,
            'parallelism': 'single' if world == 1 else f'dp{world}: batch-parallel ranks, tables row-sharded over the GPUs',
            'l2': 'inputs larger than L2: 4 rotating batches over tables >> 126 MB'}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index=0, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {'hw_slowdown': nv.nvmlClocksThrottleReasonHwSlowdown,
                 'hw_thermal_slowdown': nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 'sw_thermal_slowdown': nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 'sw_power_cap': nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


def _finite(o):
    """Replace non-finite floats by strings: the headline line must stay strict JSON whatever a secondary leg produced."""
    if isinstance(o, float):
        return o if o == o and o not in (float('inf'), float('-inf')) else str(o)
    if isinstance(o, dict):
        return {str(k): _finite(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_finite(v) for v in o]
    return o


# ----------------------------------------------------------------------------------------------- reference arm (CPU)
def _port_state_dict(w, gen):
    """Random-init weights in the reference's state_dict layout for the oracle port (DeepFM only)."""
    F, Nd, V, D = w['F'], w['Nd'], w['V'], w['D']
    hidden = w['kw']['hidden_units']
    sd = {}
    for i in range(F):
        sd[f'embedding_layer.embedding_layer.C{i + 1}.weight'] = (torch.randn(V + 1, D, generator=gen) * (2.0 / D) ** 0.5)
    dims = [F * D + Nd] + hidden
    for i in range(len(hidden)):
        sd[f'dnn.net.{2 * i}.weight'] = torch.randn(dims[i + 1], dims[i], generator=gen) * (2.0 / dims[i]) ** 0.5
        sd[f'dnn.net.{2 * i}.bias'] = torch.zeros(dims[i + 1])
    k = 2 * len(hidden)
    sd[f'dnn.net.{k}.weight'] = torch.randn(1, dims[-1], generator=gen) * (2.0 / dims[-1]) ** 0.5
    sd[f'dnn.net.{k}.bias'] = torch.zeros(1)
    for v in sd.values():
        v.requires_grad_(True)
    return sd


def cpu_reference_run(w, steps, warmup, budget_s=150.0, max_rows=None):
    """The reference's own CPU torch path of the workload on ALL host cores.  kind = "reference": the unmodified reference
    classes from oracle/_ref (oracle/build_ref.py installs them in the build container; the directory travels to the GPU
    box), driven exactly like rec_pangu/model_pipeline.py:52-58 — model(data); loss.backward(); model.zero_grad().
    kind = "port": the oracle restatement (oracle/restatement.py) when oracle/_ref is absent.  One step = one batch of up
    to B samples; if `warmup + steps` full batches would take longer than `budget_s`, every step runs the same bounded
    sample of B' < B samples (stated in `sample`).  Tables are capped at `max_rows` rows per field when the host cannot hold
    the workload (config 5: 100M x 40 floats x 8 tables); ids are drawn below the cap."""
    torch.set_num_threads(os.cpu_count() or 1)           # torchrun exports OMP_NUM_THREADS=1
    cores = torch.get_num_threads()
    from oracle import ref_loader
    w = dict(w)
    note = ''
    if max_rows is not None and w['V'] > max_rows:
        note = f'; tables capped at {max_rows} rows per field on the host (ids drawn below the cap)'
        w['V'] = max_rows
    enc = make_enc(w)
    labels = label_names(w)
    gen = torch.Generator().manual_seed(SEED)
    kind = 'reference' if ref_loader.available() else 'port'
    if kind == 'reference':
        ranking, multi_task = ref_loader.load()
        torch.manual_seed(SEED)
        cls = getattr(ranking if w['kind'] == 'ranking' else multi_task, w['model'])
        model = cls(embedding_dim=w['D'], enc_dict=enc, **w['kw'])
        if hasattr(model, 'set_device'):
            model.set_device(torch.device('cpu'))
        model.train()

        def step(batch):
            out = model(batch)
            out['loss'].backward()
            model.zero_grad()
    else:
        import oracle
        if w['model'] != 'DeepFM':
            raise RuntimeError('oracle/_ref is not built and the oracle port is wired for DeepFM only')
        sd = _port_state_dict(w, gen)

        def step(batch):
            out = oracle.deepfm(sd, enc, batch, hidden_units=tuple(w['kw']['hidden_units']))
            out['loss'].backward()
            for v in sd.values():
                v.grad = None
    B = w['B']
    full = [synth_batch(enc, B, gen, labels=labels) for _ in range(2)]
    batches, rows = full, B
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        step(batches[it % 2])
        dt = time.perf_counter() - t0
        if it == 0 and dt * (warmup + steps) > budget_s and rows == B:
            # bounded sample: keep the step COUNT, shrink the rows per step so the whole run fits the budget
            rows = max(1024, int(B * budget_s / (dt * (warmup + steps))) // 1024 * 1024)
            batches = [{k: v[:rows].clone() for k, v in b.items()} for b in full]
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return {'value': rows * len(times) / total, 'ms_per_step': 1e3 * total / len(times), 'cores': cores, 'kind': kind,
            'sample': f'{len(times)} steps x {rows} samples per step (of batch {B}; fwd + bwd + zero_grad, dense table grads as '
                      f'the reference) after {warmup} warm-up steps{note}'}


def run_reference(args, w):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if rank != 0:
        return
    r = cpu_reference_run(w, args.steps, args.warmup, max_rows=2_000_000 if w['V'] > 2_000_000 else None)
    line = {
        'metric': f'{w["model"]} samples/sec (forward+backward hot path)', 'value': r['value'], 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'impl': 'reference', 'config': make_config(w, world),
        'cpu_baseline': {'value': r['value'], 'unit': 'samples/s', 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']},
        'e2e': {'value': r['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def _median(xs):
    s = sorted(xs)
    return s[len(s) // 2]


def build_model(w, dev, world):
    """Model of the workload on `dev`; for N > 1 the tables are created as row shards (no full table anywhere)."""
    from rec_pangu_b200.models import ranking, multi_task
    enc = make_enc(w)
    cls = getattr(ranking if w['kind'] == 'ranking' else multi_task, w['model'])
    torch.manual_seed(SEED)
    st = None
    if world > 1:
        from rec_pangu_b200 import dist as rdist
        with rdist.deferred_tables():               # table Parameters are created empty; the shards are filled locally below
            with torch.device(dev):
                model = cls(embedding_dim=w['D'], enc_dict=enc, **w['kw'])
        st = rdist.shard_model_tables(model, init='kaiming')
    else:
        with torch.device(dev):
            model = cls(embedding_dim=w['D'], enc_dict=enc, **w['kw'])
    if hasattr(model, 'set_device'):
        model.set_device(dev)
    model.set_grad_mode('persistent')
    model.train()
    return model, enc, st


def run_ours(args, w):
    from rec_pangu_b200 import ops, _lib
    from rec_pangu_b200.runtime import ColumnarBatch, GraphedStep

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    B = w['B']
    for kv in filter(None, os.environ.get('RPB_OPTIONS', '').split(',')):       # tuning knobs, e.g. RPB_OPTIONS=wgrad_stages=2
        k, v = kv.split('=')
        _lib.check(_lib.load().rpb_set_option(k.encode(), int(v)), f'rpb_set_option({k})')
    model, enc, st = build_model(w, dev, world)
    labels = label_names(w)

    # N > 1: batch-parallel ranks, tables row-sharded over the GPUs in NVLink peer memory (the lookup and
    # the gradient scatter cross NVLink inside the gather/scatter kernels), dense grads summed with one NCCL all-reduce.
    post, loss_scale = None, 1.0
    if world > 1:
        from rec_pangu_b200 import dist as rdist
        torch.cuda.empty_cache()
        bucket = rdist.DenseGradBucket([p for n, p in model.named_parameters() if not n.startswith('embedding_layer.')
                                        and '.emb_layer.' not in n])
        post, loss_scale = bucket.all_reduce, 1.0 / world

    NB = 4
    gen = torch.Generator(device=dev).manual_seed(SEED + rank)
    cbs, steps_g = [], []
    for i in range(NB):
        cb = ColumnarBatch(enc, B, label_names=labels, device=dev, pinned_host=(i == 0))
        cb.load_device(synth_batch(enc, B, gen, device=dev, labels=labels))
        cbs.append(cb)
    use_graph = not args.eager
    launch_mode = 'cuda_graph' if use_graph else 'eager'
    # one GPU: the step opens with zero_grad (where the reference's loop has it), its sparse re-zero running next to the forward
    # kernel; N > 1: the sharded gradient tables are cleaned lazily by the next backward (ops.sharded_clean), step order unchanged
    zero_first = bool(args.zero_first) and world == 1
    try:
        steps_g = GraphedStep.ring(model, cbs, post=post, use_graph=use_graph, loss_scale=loss_scale, zero_first=zero_first)
    except Exception as e:      # capture not possible: time the eager path instead (still the same kernels)
        if rank == 0:
            print(f'[bench] CUDA-graph capture failed ({e!r}); falling back to eager launches', file=sys.stderr)
        launch_mode = 'eager'
        torch.cuda.synchronize()
        steps_g = GraphedStep.ring(model, cbs, post=post, use_graph=False, loss_scale=loss_scale, zero_first=zero_first)
    ops.check_index_errors(dev)
    launches_per_step = steps_g[0].launches_per_step

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_window(fn, n):
        """Exactly n calls of fn(i) between two events, a barrier + synchronize on both sides, max over ranks (ms)."""
        barrier()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---------------- timed region: K steps, inputs resident in HBM, rotating over NB batches (tables >> L2).  N > 1: the
    # window is repeated (NVLink contention makes a single 10-30 ms window noisy) and the MEDIAN window is reported.
    repeats = args.repeats if args.repeats > 0 else (1 if world == 1 else 5)
    # the steps are replayed strictly in ring order (with zero_first step j re-zeroes what step j-1 touched)
    ring = {'pos': 0}

    def step_next():
        j = ring['pos']
        steps_g[j].replay()
        ring['pos'] = (j + 1) % NB
        return j

    for i in range(args.warmup):
        step_next()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    windows = [timed_window(lambda i: step_next(), args.steps) for _ in range(repeats)]
    clocks = sampler.stop()
    ms = _median(windows)
    value = world * B * args.steps / (ms * 1e-3)

    # ---------------- e2e: every step's inputs start in pinned HOST memory: H2D (3 copies) -> step -> D2H read of the loss.
    # Double-buffered: the copy of batch i+1 runs on a copy stream while batch i computes (each batch is still copied
    # exactly once per step inside the timed region).
    for cb in cbs:
        if cb.h_idx is None:
            cb.h_idx = torch.zeros_like(cb.idx, device='cpu').pin_memory()
            cb.h_dns = torch.zeros_like(cb.dns, device='cpu').pin_memory()
            cb.h_lab = torch.zeros_like(cb.lab, device='cpu').pin_memory()
        cb.h_idx.copy_(cb.idx)
        cb.h_dns.copy_(cb.dns)
        cb.h_lab.copy_(cb.lab)
    main_stream = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream()
    ev_copied = [torch.cuda.Event() for _ in range(NB)]
    ev_done = [torch.cuda.Event() for _ in range(NB)]
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()          # D2H landing zone of the per-step loss
    ev_loss = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {'h2d': 0, 'loss': 0.0}

    def e2e_loop(n):
        """Every step: H2D of its batch from pinned memory (3 copies), the graph replay, and a D2H copy of its loss that the
        host reads.  The loss of step i is read while step i+1 runs (async metrics: the copy is queued right behind the
        step, the host waits for its event one iteration later), so the device never idles on a host round trip; the last
        loss is read before the timed region closes.  Batches follow the ring order of the steps."""
        main_stream.synchronize()
        with torch.cuda.stream(copy_stream):
            e2e_state['h2d'] = cbs[ring['pos']].h2d()
            ev_copied[ring['pos']].record(copy_stream)
        for i in range(n):
            cur = ring['pos']
            nxt = (cur + 1) % NB
            copy_stream.wait_event(ev_done[nxt])                # buffer `nxt` was last read NB-1 steps ago (no-op before its first use)
            with torch.cuda.stream(copy_stream):
                cbs[nxt].h2d()
                ev_copied[nxt].record(copy_stream)
            main_stream.wait_event(ev_copied[cur])
            step_next()
            ev_done[cur].record(main_stream)
            lh = i % 2
            loss_host[lh:lh + 1].copy_(steps_g[cur].loss.reshape(1), non_blocking=True)       # D2H of this step's result
            ev_loss[lh].record(main_stream)
            if i >= 1:
                ev_loss[1 - lh].synchronize()                    # loss of step i-1 has landed
                e2e_state['loss'] = float(loss_host[1 - lh])
        ev_loss[(n - 1) % 2].synchronize()
        e2e_state['loss'] = float(loss_host[(n - 1) % 2])

    e2e_loop(max(3, args.warmup))
    e2e_windows = []
    for _ in range(repeats):
        barrier()
        e0.record()
        e2e_loop(args.steps)
        e1.record()
        barrier()
        e2e_windows.append(max_over_ranks(e0.elapsed_time(e1)))
    ms_e2e = _median(e2e_windows)
    # leave the ring at its start: the last step replayed is the last one captured, so the host-side list of touched rows
    # matches the device again and a plain zero_grad() cleans exactly what is dirty before the secondary legs run
    while ring['pos'] != 0:
        step_next()
    torch.cuda.synchronize()
    if zero_first:
        model.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
    e2e = {'value': world * B * args.steps / (ms_e2e * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': e2e_state['h2d'],
           'd2h_bytes_per_step': 4, 'loss': e2e_state['loss'], 'ms_per_step': ms_e2e / args.steps,
           'overlap': 'H2D of batch i+1 on a copy stream during step i; loss of step i read by the host during step i+1'}

    line = {
        'metric': f'{w["model"]} samples/sec (forward+backward hot path)', 'value': value, 'unit': 'samples/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': make_config(w, world),
        'run': {'launch': launch_mode, 'repeats': repeats, 'window_ms': [round(x, 4) for x in windows],
                'e2e_window_ms': [round(x, 4) for x in e2e_windows], 'timing': 'median window of `repeats`, each exactly `steps` steps, '
                'CUDA events, barrier + synchronize on both sides, max over ranks',
                'gemm': {0: 'auto(tcgen05 3xTF32)', 1: 'simt fp32', 2: 'tcgen05 3xTF32'}[ops.get_gemm_impl()],
                'grad_mode': 'persistent' if world == 1 else 'sharded',
                'zero_grad': ('opens the step (side stream, 148 x 128-thread blocks next to the forward kernel), joined before backward'
                              if zero_first else 'closes the step'),
                'sharded_fused_core': bool(ops.SHARDED_FUSED) if world > 1 else None,
                'tables_per_gpu_bytes': sum(p.numel() * 4 for n, p in model.named_parameters() if 'embedding_layer' in n)},
        'e2e': e2e, 'gpu_launches': launches_per_step * args.steps, 'clocks': clocks, 'roofline': None,
        'train_step': None, 'train_model': None, 'zipf_ids': None, 'torch_eager_gpu_baseline': None,
    }

    # Everything below is secondary.  The line exists from here on and is filled in leg by leg; a watchdog prints it as it
    # stands if the secondary legs ever fail to come back (a hang there must not cost the headline numbers above).
    _lock, _state = threading.Lock(), {'done': False}

    def emit_line(note=None):
        with _lock:
            if _state['done']:
                return
            _state['done'] = True
            if note is not None:
                line['watchdog'] = note
            text = None
            for _ in range(5):
                try:
                    try:
                        text = json.dumps(_finite(line), allow_nan=False)
                    except (ValueError, TypeError):
                        text = json.dumps(line)
                    break
                except RuntimeError:             # dict touched by the main thread while serialising
                    time.sleep(0.05)
            print(text if text is not None else json.dumps({k: line[k] for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step')}), flush=True)

    watchdog = None
    if world == 1 and rank == 0:
        def _fire():
            emit_line('secondary legs did not return within 420 s of the headline measurement; line printed by the watchdog')
            sys.stdout.flush()
            os._exit(0)
        watchdog = threading.Timer(420.0, _fire)
        watchdog.daemon = True
        watchdog.start()

    def graph_time(make, n):
        """us per replay of NB graphs built by make(cb) (one per rotating batch), n timed replays."""
        gs, keep = [], []
        for cb in cbs:
            keep.append(make(cb))
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                keep.append(make(cb))
            gs.append(g)
        for i in range(3):
            gs[i % NB].replay()
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            gs[i % NB].replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / n

    if world == 1:
        line['roofline'] = roofline_leg(w, model, cbs, ms / args.steps, args.steps, graph_time)

    # ---------------- same step + optimizer: FusedAdam between backward and zero_grad — dense
    # parameters in one multi-tensor launch, table rows row-sparsely with the gradient re-zero fused in (rpb_sparse_adam),
    # step counter on the device so the captured graph keeps its bias correction.  Reported beside the headline.
    if world == 1 and not args.no_train_step:
        try:
            from rec_pangu_b200.optim import FusedAdam
            opt = FusedAdam(model, lr=1e-3)
            tsteps = [GraphedStep(model, cb, post=opt.step, use_graph=use_graph) for cb in cbs]
            for i in range(max(3, args.warmup)):
                tsteps[i % NB].replay()
            torch.cuda.synchronize()
            e0.record()
            for i in range(args.steps):
                tsteps[i % NB].replay()
            e1.record()
            torch.cuda.synchronize()
            ms_t = e0.elapsed_time(e1)
            line['train_step'] = {'value': B * args.steps / (ms_t * 1e-3), 'unit': 'samples/s', 'ms_per_step': ms_t / args.steps,
                                  'what': 'forward + backward + FusedAdam (row-sparse Adam on the touched table rows, gradient re-zero fused), '
                                          'runtime.GraphedStep on device-resident batches',
                                  'gpu_launches_per_step': tsteps[0].launches_per_step, 'loss_after': float(tsteps[0].loss.item())}
            del tsteps
            # ---- the drop-in entry point itself: model_pipeline.train_model over an in-memory loader of HOST batches
            # (rec_pangu/model_pipeline.py:17-125), optimizer_type='fused_adam' as RankTrainer.fit would build it
            line['train_model'] = train_model_leg(w, model, enc, opt, dev, labels, args.steps)
            del opt
        except Exception as ex:          # a secondary leg must never cost the headline line
            if line['train_step'] is None:
                line['train_step'] = {'error': repr(ex)}
            else:
                line['train_model'] = {'error': repr(ex)}

    # ---------------- secondary legs (N = 1 only): Zipf(1.05)-distributed ids and stock PyTorch eager on the
    # same GPU (the oracle's functional restatement of the reference forward run on CUDA tensors = the "existing Blackwell
    # path" a user of the reference gets from `.to('cuda')`: F separate embedding lookups + stack + cat + addmm chain, dense
    # [V+1, D] table gradients from autograd)
    if world == 1 and not args.no_extras:
        try:
            import numpy as np
            rng = np.random.default_rng(SEED)
            zsteps = []
            for i in range(2):
                cb = ColumnarBatch(enc, B, label_names=labels, device=dev, pinned_host=False)
                d = synth_batch(enc, B, gen, device=dev, labels=labels)
                for c in cb.sparse:
                    d[c] = torch.from_numpy(((rng.zipf(1.05, B) - 1) % (w['V'] + 1)).astype('int64')).to(dev)
                cb.load_device(d)
                zsteps.append(GraphedStep(model, cb, use_graph=use_graph))
            for i in range(3):
                zsteps[i % 2].replay()
            torch.cuda.synchronize()
            e0.record()
            for i in range(args.steps):
                zsteps[i % 2].replay()
            e1.record()
            torch.cuda.synchronize()
            ms_z = e0.elapsed_time(e1)
            line['zipf_ids'] = {'value': B * args.steps / (ms_z * 1e-3), 'unit': 'samples/s', 'ms_per_step': ms_z / args.steps,
                                'ids': 'zipf(1.05) - 1 mod (V+1) per field'}
            del zsteps
        except Exception as ex:
            line['zipf_ids'] = {'error': repr(ex)}
        try:
            line['torch_eager_gpu_baseline'] = eager_gpu_leg(w, model, enc, cbs[0].as_dict())
        except Exception as ex:
            line['torch_eager_gpu_baseline'] = {'error': repr(ex)}
            try:
                torch.cuda.synchronize()
            except Exception:
                pass

    if rank == 0:
        line['wall_s_before_cpu_baseline'] = round(time.time() - T_START, 1)
        if world == 1 and not args.no_cpu_baseline:
            try:
                del steps_g
                torch.cuda.empty_cache()
            except Exception:
                pass
            line['cpu_baseline'] = cpu_baseline_child(args.workload)
        if watchdog is not None:
            watchdog.cancel()
        emit_line()
    if dist is not None:
        # symmetric-memory + NCCL teardown can block for minutes at interpreter exit; the numbers are out, leave hard
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def cpu_baseline_child(workload):
    """`cpu_baseline` = the reference arm of this same file on a bounded sample (3 steps after 1 warm-up, ~10-30 s of CPU
    work), in a CHILD process: the reference package is also called `rec_pangu` and must not meet this repo's alias package
    in one interpreter, and its 1.7 GB of host tables are gone when the child exits.  Never raises."""
    import subprocess
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--workload', workload, '--steps', '3',
                            '--warmup', '3', '--cpu-budget', '30'], capture_output=True, text=True, timeout=400,
                           env={k: v for k, v in os.environ.items() if k not in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS')})
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith('{'):
                return json.loads(ln)['cpu_baseline']
        return {'error': f'no result (rc {r.returncode}): ' + (r.stderr or '')[-300:]}
    except Exception as ex:
        return {'error': repr(ex)}


def roofline_leg(w, model, cbs, ms_step, steps, graph_time):
    """`roofline` of the dominant memory kernel of the step's forward, on the ALGORITHMIC bytes of the embedding + interaction
    stage over the measured HBM copy bandwidth.  DeepFM: the one-kernel training forward the step launches
    (deepfm_fwd_fs_kernel: gather + dense pack + FM + layer-1 tcgen05 GEMM + tcgen05 tower tail + BCE; the 2 weight-split
    launches in front of it are inside the interval).  Other models: their gather launch (multi-table gather + dense pack +
    LR rows), timed alone."""
    from rec_pangu_b200 import ops
    B = w['B']
    peak, peak_src = measured_peaks()
    alg = alg_bytes_per_sample(w)
    out = None
    try:
        emb = model.embedding_layer
        tables = emb.tables()
        lr_tables = None
        for name in ('lr_layer', 'lr'):
            if hasattr(model, name):
                lr_tables = getattr(model, name).emb_layer.tables()

        def gather_only(cb, want_x=True):
            d = cb.as_dict()
            idx = [d[c] for c in emb.emb_feature]
            dn = [d[c] for c in emb.dense_feature]
            with torch.no_grad():
                return ops.gather(tables, idx, dn, lr_tables=lr_tables, want_fm=(w['model'] == 'DeepFM'), want_x=want_x)

        us_g = graph_time(lambda cb: gather_only(cb, True), steps)
        gather = {'kernel': 'gather_fwd_tile_kernel (multi-table gather + dense pack' + (' + FM second order' if w['model'] == 'DeepFM' else ' + LR rows')
                            + ', x materialised), timed alone', 'bound': 'hbm', 'achieved': alg * B / (us_g * 1e-6) / 1e9, 'peak': peak,
                  'unit': 'GB/s', 'frac': alg * B / (us_g * 1e-6) / 1e9 / peak, 'traffic': None, 'us_per_launch': us_g,
                  'alg_bytes_per_launch': alg * B, 'peak_source': peak_src}
        out = gather
        if w['model'] == 'DeepFM':
            us_n = graph_time(lambda cb: gather_only(cb, False), steps)
            gather['traffic'] = 185.7e6
            gather['traffic_source'] = 'profiles/r01_deepfm_step_ncu_full.md (ncu --set full: dram__bytes_read 125.3 MB + write 60.4 MB per launch)'
            gather['no_materialise'] = {'us_per_launch': us_n, 'achieved': alg * B / (us_n * 1e-6) / 1e9,
                                        'frac': alg * B / (us_n * 1e-6) / 1e9 / peak}
            us_f = graph_time(lambda cb: model(cb.as_dict()), steps)
            ach = alg * B / (us_f * 1e-6) / 1e9
            saved = alg + 4 * ((w['F'] * w['D'] + w['Nd'] + 3) // 4 * 4) + 4 * 64 * len(w['kw']['hidden_units']) + 4 * w['D'] + 8
            out = {'kernel': 'deepfm_fwd_fs_kernel (gather + dense pack + FM + layer-1 tcgen05 GEMM + tcgen05 tower tail + BCE in one '
                             'launch, training variant: x and activations stored), the forward of the timed step',
                   'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                   'traffic': 250.5e6, 'traffic_source': 'profiles/r02_deepfm_step_ncu.md (ncu --set full of deepfm_fwd_fs_kernel<6, 0, 8>: dram__bytes_read 127.7 MB + write 122.8 MB per launch)',
                   'us_per_launch': us_f, 'alg_bytes_per_launch': alg * B, 'peak_source': peak_src,
                   'share_of_step': us_f / (1e3 * ms_step), 'forward_only_samples_per_s': B / (us_f * 1e-6),
                   # the same launch against the bytes a TRAINING forward has to move: the algorithmic reads plus what it must
                   # leave behind for backward (feature row x, h1..h3, fm_s, logit, pred)
                   'with_saved_activations': {'bytes_per_sample': saved, 'achieved': saved * B / (us_f * 1e-6) / 1e9,
                                              'frac': saved * B / (us_f * 1e-6) / 1e9 / peak},
                   'floor_note': 'random 64-byte rows are not bound by bytes: 1.7 M row reads alone take 49.5 us with 8 requesting warps '
                                 'per SM (tools/exp/exp_rowfetch.cu: LDGSTS, LDG, TMA gather4, cp.async.bulk all >= 49 us), with the x store '
                                 '74 us; the kernel adds h1..h3.  The limit is the rate at which an SM gets row requests accepted, not DRAM '
                                 'latency: an L2 prefetch warp running ahead made the kernel 34 us slower (profiles/r02_rowfetch.md)',
                   'gather_only': gather}
    except Exception as ex:
        out = dict(out or {}, error=repr(ex))
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
    return out


class _MemLoader:
    """In-memory stand-in for the reference's DataLoader: yields dict batches of HOST tensors (what default-collate of
    BaseDataset.__getitem__ produces, base_dataset.py:105-124) and has the two attributes train_model reads."""

    def __init__(self, batches, batch_size):
        self.batches, self.batch_size = batches, batch_size
        self.dataset = range(len(batches) * batch_size)

    def __iter__(self):
        for b in self.batches:
            yield dict(b)

    def __len__(self):
        return len(self.batches)


def train_model_leg(w, model, enc, opt, dev, labels, steps):
    """model_pipeline.train_model (the reference's entry point, same signature) over `n` host batches: wall clock around
    the call, device synchronised on both sides.  Includes the host-side packing of every batch dict, the H2D copies, the
    step (captured once as a CUDA graph inside train_model when the optimizer is graph-safe), and the end-of-epoch metrics."""
    from rec_pangu_b200.model_pipeline import train_model
    B = w['B']
    n = max(8, min(steps, 40))
    gen = torch.Generator().manual_seed(SEED + 7)

    def columnar(d):
        """The batch as a columnar loader hands it out: every column a row view of one pinned [n_cols, B] buffer per dtype."""
        out, groups = {}, {}
        for k, v in d.items():
            groups.setdefault(v.dtype, []).append(k)
        for dt, keys in groups.items():
            buf = torch.empty((len(keys), B), dtype=dt).pin_memory()
            for i, k in enumerate(keys):
                buf[i].copy_(d[k])
                out[k] = buf[i]
        return {k: out[k] for k in d}
    host = [columnar(synth_batch(enc, B, gen, labels=labels)) for _ in range(4)]
    loader = _MemLoader([host[i % 4] for i in range(n)], B)
    warm = _MemLoader([host[i % 4] for i in range(8)], B)          # 2 eager runs + the graph capture per staging buffer happen here
    num_task = 1 if w['kind'] == 'ranking' else 2
    train_model(model, warm, opt, dev, metric_list=[], num_task=num_task, log_rounds=10 ** 9)
    torch.cuda.synchronize()
    runs = []
    for _ in range(3):                       # wall clock over ~40 ms of work: one host hiccup moves a single run by tens of percent
        t0 = time.perf_counter()
        train_model(model, loader, opt, dev, metric_list=[], num_task=num_task, log_rounds=10 ** 9)
        torch.cuda.synchronize()
        runs.append(time.perf_counter() - t0)
    dt = sorted(runs)[1]
    return {'value': B * n / dt, 'unit': 'samples/s', 'ms_per_step': 1e3 * dt / n, 'steps': n,
            'runs_ms_per_step': [round(1e3 * r / n, 4) for r in runs], 'reported': 'median of 3 calls',
            'what': 'model_pipeline.train_model(model, loader of host dict batches (row views of pinned columnar buffers), FusedAdam, '
                    'device): H2D + fwd + bwd + optimizer + zero_grad per batch (step replayed as a CUDA graph), predictions kept '
                    'for the epoch metrics; wall clock'}


def eager_gpu_leg(w, model, enc, dd):
    """Stock PyTorch eager ops of the reference forward + autograd backward on the same GPU (oracle restatement on CUDA tensors)."""
    import oracle
    if w['V'] > 2_000_000:
        return {'skipped': 'dense [V+1, D] autograd table gradients of 100M-row tables do not fit next to the tables'}
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    if w['model'] == 'MMOE':
        return {'skipped': 'gates are unregistered parameters in the reference; not wired for the eager leg'}
    fwd = oracle.MODEL_FORWARDS[w['model']]
    n_e = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(2 + n_e):
        if it == 2:
            torch.cuda.synchronize()
            e0.record()
        out = fwd(sd, enc, dd, **w['okw'])
        out['loss'].backward()
        for v in sd.values():
            v.grad = None
    e1.record()
    torch.cuda.synchronize()
    ms_e = e0.elapsed_time(e1)
    del sd, out
    torch.cuda.empty_cache()
    return {'value': w['B'] * n_e / (ms_e * 1e-3), 'unit': 'samples/s', 'ms_per_step': ms_e / n_e,
            'what': 'stock PyTorch eager ops of the reference forward + autograd backward on the same B200 (fp32, eval-mode dropout)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='deepfm', choices=sorted(WORKLOADS))
    ap.add_argument('--repeats', type=int, default=0, help='timed windows of `steps` steps (median reported); 0 = 1 at N=1, 5 at N>1')
    ap.add_argument('--zero-first', type=int, default=0, help='1 GPU: 1 = the step opens with zero_grad on a side stream next to the forward kernel (measured slower: the two kernels do not share an SM, profiles/r02_zero_first.md); 0 = zero_grad closes the step')
    ap.add_argument('--eager', action='store_true', help='time eager launches instead of CUDA-graph replays')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train-step', action='store_true', help='skip the secondary forward+backward+optimizer timings')
    ap.add_argument('--no-extras', action='store_true', help='skip the Zipf-id and stock-PyTorch-eager-GPU secondary timings')
    ap.add_argument('--no-experiments', action='store_true', help='accepted for the older trip scripts; nothing to skip any more')
    ap.add_argument('--cpu-budget', type=float, default=150.0, help='reference arm: seconds of CPU work before steps shrink to a bounded sample')
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    w = WORKLOADS[args.workload]
    if args.impl == 'reference':
        global _CPU_BUDGET
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == '__main__':
    main()
