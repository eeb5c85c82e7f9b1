#!/usr/bin/env python
"""bench.py — DeepFM hot-path throughput on B200 (BASELINE.json metric: "DeepFM samples/sec at 1/2/4/8 B200;
embedding-gather HBM GB/s vs peak").

Workload (BASELINE.json configs[1]): DeepFM, synthetic Criteo shape — 26 sparse fields x 1M-row tables (D=16),
13 dense fields, hidden [64,64,64], batch 65536 per GPU (weak scaling), random-init weights, uniform ids.
One "step" = one pass of the hot path over one batch: forward (gather -> FM -> MLP -> sigmoid/BCE) + backward
(MLP grads, FM grad, scatter-add into the dense per-table gradient buffers) + sparse re-zero of those buffers
(`model.zero_grad()`), i.e. rec_pangu/model_pipeline.py:52-58 without optimizer.step (the
optimizer is a "next" row, not part of the path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

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

CFG = dict(F=26, Nd=13, V=1_000_000, D=16, B=65536, hidden=[64, 64, 64])
SEED = 1029
ALG_BYTES_PER_SAMPLE = CFG['F'] * (8 + 4 * CFG['D']) + 4 * CFG['Nd'] + 4        # 1928 B


def make_enc():
    enc = {f'I{i + 1}': {'min': 0.0, 'max': 1.0} for i in range(CFG['Nd'])}
    enc.update({f'C{i + 1}': {'vocab_size': CFG['V']} for i in range(CFG['F'])})
    return enc


def synth_batch(enc, B, gen, device='cpu'):
    d = {}
    for c, m in enc.items():
        if 'vocab_size' in m:
            d[c] = torch.randint(0, m['vocab_size'] + 1, (B,), dtype=torch.int64, generator=gen, device=device)
        else:
            d[c] = torch.rand(B, generator=gen, device=device)
    d['label'] = (torch.rand(B, generator=gen, device=device) < 0.25).float()
    return d


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


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_reference_run(steps, warmup, B, threads=None, min_seconds=None):
    """The reference's own CPU torch path (oracle port: oracle/restatement.py restates DeepFM.forward op for op;
    /root/reference cannot travel to the GPU box).  One step = forward + loss.backward() on one batch of B samples,
    same shapes/weights layout as the GPU arm (dense [V+1,D] table grads zero-filled by autograd, as the reference)."""
    import oracle
    if threads:
        torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    enc = make_enc()
    gen = torch.Generator().manual_seed(SEED)
    F, Nd, V, D = CFG['F'], CFG['Nd'], CFG['V'], CFG['D']
    sd = {}
    for i in range(F):
        sd[f'embedding_layer.embedding_layer.C{i + 1}.weight'] = (torch.randn(V + 1, D, generator=gen) * (2.0 / D) ** 0.5)
    dims = [F * D + Nd] + CFG['hidden']
    for i in range(len(CFG['hidden'])):
        sd[f'dnn.net.{2 * i}.weight'] = torch.randn(dims[i + 1], dims[i], generator=gen) * (2.0 / dims[i]) ** 0.5
        sd[f'dnn.net.{2 * i}.bias'] = torch.zeros(dims[i + 1])
    k = 2 * len(CFG['hidden'])
    sd[f'dnn.net.{k}.weight'] = torch.randn(1, dims[-1], generator=gen) * (2.0 / dims[-1]) ** 0.5
    sd[f'dnn.net.{k}.bias'] = torch.zeros(1)
    for v in sd.values():
        v.requires_grad_(True)
    batches = [synth_batch(enc, B, gen) for _ in range(2)]
    times = []
    for it in range(warmup + steps):
        if min_seconds is not None and len(times) >= 3 and sum(times) >= min_seconds:
            break                                   # bounded sample: about min_seconds of CPU work
        t0 = time.perf_counter()
        out = oracle.deepfm(sd, enc, batches[it % 2], hidden_units=tuple(CFG['hidden']))
        out['loss'].backward()
        for v in sd.values():
            v.grad = None
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return {'value': B * len(times) / total, 'ms_per_step': 1e3 * total / len(times), 'cores': cores,
            'sample': f'{len(times)} steps x {B} samples (fwd+bwd, dense table grads) after {warmup} warm-up'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    B = CFG['B']
    steps = max(1, min(args.steps, 20))
    r = cpu_reference_run(steps, min(args.warmup, 2), B)
    line = {
        'metric': 'DeepFM samples/sec (forward+backward hot path)', 'value': r['value'], 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': min(args.warmup, 2), 'ms_per_step': r['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'impl': 'reference',
        'config': {'workload': 'DeepFM criteo-shape (26 sparse x 1M vocab, 13 dense, D=16, MLP 64-64-64), '
                               'CPU torch path of the reference restated op-for-op (oracle port)',
                   'batch_per_step': B, 'note': 'bounded sample: at most 20 full-size steps'},
        'cpu_baseline': {'value': r['value'], 'unit': 'samples/s', 'cores': r['cores'], 'kind': 'port', 'sample': r['sample']},
        'e2e': {'value': r['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    from rec_pangu_b200 import ops
    from rec_pangu_b200.models.ranking import DeepFM
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

    B, F, Nd, D = CFG['B'], CFG['F'], CFG['Nd'], CFG['D']
    for kv in filter(None, os.environ.get('RPB_OPTIONS', '').split(',')):       # tuning knobs, e.g. RPB_OPTIONS=wgrad_stages=2
        k, v = kv.split('=')
        from rec_pangu_b200 import _lib
        _lib.check(_lib.load().rpb_set_option(k.encode(), int(v)), f'rpb_set_option({k})')
    enc = make_enc()
    torch.manual_seed(SEED)
    with torch.device(dev):
        model = DeepFM(embedding_dim=D, hidden_units=CFG['hidden'], enc_dict=enc)
    model.set_grad_mode('persistent')
    model.train()

    # N > 1: batch-parallel ranks, tables row-sharded over the GPUs in NVLink peer memory (the lookup and
    # the gradient scatter cross NVLink inside the gather/scatter kernels), dense grads summed with one NCCL all-reduce.
    post, loss_scale, st = None, 1.0, None
    if world > 1:
        from rec_pangu_b200 import dist as rdist
        st = rdist.shard_model_tables(model)
        torch.cuda.empty_cache()
        bucket = rdist.DenseGradBucket([p for n, p in model.named_parameters() if not n.startswith('embedding_layer.')])
        post, loss_scale = bucket.all_reduce, 1.0 / world

    NB = 4
    gen = torch.Generator(device=dev).manual_seed(SEED + rank)
    cbs, steps_g = [], []
    for i in range(NB):
        cb = ColumnarBatch(enc, B, device=dev, pinned_host=(i == 0))
        cb.load_device(synth_batch(enc, B, gen, device=dev))
        cbs.append(cb)
    use_graph = not args.eager
    launch_mode = 'cuda_graph' if use_graph else 'eager'
    try:
        for cb in cbs:
            steps_g.append(GraphedStep(model, cb, post=post, use_graph=use_graph, loss_scale=loss_scale))
    except Exception as e:      # capture not possible: time the eager path instead (still the same kernels)
        if rank == 0:
            print(f'[bench] CUDA-graph capture failed ({e!r}); falling back to eager launches', file=sys.stderr)
        launch_mode = 'eager'
        torch.cuda.synchronize()
        steps_g = [GraphedStep(model, cb, post=post, use_graph=False, loss_scale=loss_scale) for cb in cbs]
    ops.check_index_errors(dev)
    launches_per_step = steps_g[0].launches_per_step

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- timed region: K steps, inputs resident in HBM, rotating over NB batches (tables 1.66 GB >> L2)
    for i in range(args.warmup):
        steps_g[i % NB].replay()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        steps_g[i % NB].replay()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---------------- e2e: every step's inputs start in pinned HOST memory: H2D (3 copies) -> step -> D2H read of the loss.
    # Double-buffered: the copy of batch i+1 runs on a copy stream while batch i computes (each batch is still copied
    # exactly once per step inside the timed region).
    for cb in cbs[:2]:
        if cb.h_idx is None:
            cb.h_idx = torch.zeros_like(cb.idx, device='cpu').pin_memory()
            cb.h_dns = torch.zeros_like(cb.dns, device='cpu').pin_memory()
            cb.h_lab = torch.zeros_like(cb.lab, device='cpu').pin_memory()
        cb.h_idx.copy_(cb.idx)
        cb.h_dns.copy_(cb.dns)
        cb.h_lab.copy_(cb.lab)
    main_stream = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream()
    ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
    ev_done = [torch.cuda.Event(), torch.cuda.Event()]

    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()          # D2H landing zone of the per-step loss
    ev_loss = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop(n):
        """Every step: H2D of its batch from pinned memory (3 copies), the graph replay, and a D2H copy of its loss that the
        host reads.  The loss of step i is read while step i+1 runs (async metrics: the copy is queued right behind the
        step, the host waits for its event one iteration later), so the device never idles on a host round trip; the last
        loss is read before the timed region closes."""
        h2d_bytes, lossv = 0, 0.0
        with torch.cuda.stream(copy_stream):
            h2d_bytes = cbs[0].h2d()
            ev_copied[0].record(copy_stream)
        for i in range(n):
            cur, nxt = i % 2, (i + 1) % 2
            if i >= 1:
                copy_stream.wait_event(ev_done[nxt])            # buffer `nxt` was last read by step i-1
            with torch.cuda.stream(copy_stream):
                cbs[nxt].h2d()
                ev_copied[nxt].record(copy_stream)
            main_stream.wait_event(ev_copied[cur])
            steps_g[cur].replay()
            ev_done[cur].record(main_stream)
            loss_host[cur:cur + 1].copy_(steps_g[cur].loss.reshape(1), non_blocking=True)     # D2H of this step's result
            ev_loss[cur].record(main_stream)
            if i >= 1:
                ev_loss[nxt].synchronize()                       # loss of step i-1 has landed
                lossv = float(loss_host[nxt])
        ev_loss[(n - 1) % 2].synchronize()
        lossv = float(loss_host[(n - 1) % 2])
        return h2d_bytes, lossv

    e2e_loop(max(3, args.warmup))
    barrier()
    e0.record()
    h2d, lossv = e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e = {'value': world * B * args.steps / (ms_e2e * 1e-3), 'unit': 'samples/s', 'h2d_bytes_per_step': h2d,
           'd2h_bytes_per_step': 4, 'loss': lossv, 'overlap': 'H2D of batch i+1 on a copy stream during step i; loss of step i read by the host during step i+1'}

    line = {
        'metric': 'DeepFM samples/sec (forward+backward hot path)', 'value': value, 'unit': 'samples/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'DeepFM criteo-shape: 26 sparse x 1M-row tables (D=16) + 13 dense, MLP 64-64-64, '
                               'fwd + bwd (dense per-table grads, sparse re-zero), BASELINE.json configs[1]',
                   'batch_per_gpu': B, 'global_batch': B * world, 'parallelism': (f'dp{world}: batch-parallel ranks, tables row-sharded in NVLink peer memory (fused P2P gather/scatter), '
                                   f'dense grads NCCL all-reduce') if world > 1 else 'single',
                   'launch': launch_mode, 'l2': 'inputs larger than L2: 4 rotating batches over 1.66 GB of tables',
                   'gemm': {0: 'auto(tcgen05 3xTF32)', 1: 'simt fp32', 2: 'tcgen05 3xTF32'}[ops.get_gemm_impl()],
                   'grad_mode': 'persistent' if world == 1 else 'sharded',
                   'sharded_fused_core': bool(ops.SHARDED_FUSED) if world > 1 else None},
        'e2e': e2e, 'gpu_launches': launches_per_step * args.steps, 'clocks': clocks, 'roofline': None,
        'train_step': None, 'zipf_ids': None, 'torch_eager_gpu_baseline': None,
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

    roofline = None
    if world == 1:
        # ---------------- roofline of the dominant memory kernel: the fused gather+FM forward (rpb_gather_fwd)
        tables = model.embedding_layer.tables()
        gg = []
        with torch.no_grad():
            for cb in cbs:
                d = cb.as_dict()
                idx = [d[c] for c in model.embedding_layer.emb_feature]
                dn = [d[c] for c in model.embedding_layer.dense_feature]
                ops.gather(tables, idx, dn, want_fm=True)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    ops.gather(tables, idx, dn, want_fm=True)
                gg.append(g)
        for i in range(3):
            gg[i % NB].replay()
        torch.cuda.synchronize()
        e0.record()
        for i in range(args.steps):
            gg[i % NB].replay()
        e1.record()
        torch.cuda.synchronize()
        us_gather = e0.elapsed_time(e1) * 1e3 / args.steps
        peak, peak_src = measured_peaks()
        achieved = ALG_BYTES_PER_SAMPLE * B / (us_gather * 1e-6) / 1e9
        # same stage without materialising the [B,F,D] rows (what FM inference runs; SURVEY.md §8d defines the
        # algorithmic bytes of the embedding+interaction forward without the optional materialisation)
        gn = []
        with torch.no_grad():
            for cb in cbs:
                d = cb.as_dict()
                idx = [d[c] for c in model.embedding_layer.emb_feature]
                dn = [d[c] for c in model.embedding_layer.dense_feature]
                ops.gather(tables, idx, dn, want_fm=True, want_x=False)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    ops.gather(tables, idx, dn, want_fm=True, want_x=False)
                gn.append(g)
        for i in range(3):
            gn[i % NB].replay()
        torch.cuda.synchronize()
        e0.record()
        for i in range(args.steps):
            gn[i % NB].replay()
        e1.record()
        torch.cuda.synchronize()
        us_nomat = e0.elapsed_time(e1) * 1e3 / args.steps
        gather_only = {'kernel': 'gather_fwd_tile_kernel (multi-table gather + dense pack + FM second order, x materialised), timed alone',
                    'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                    'traffic': 185.7e6, 'traffic_source': 'profiles/r01_deepfm_step_ncu_full.md (ncu --set full: dram__bytes_read 125.3 MB + write 60.4 MB per launch)',
                    'us_per_launch': us_gather, 'alg_bytes_per_launch': ALG_BYTES_PER_SAMPLE * B, 'peak_source': peak_src,
                    'no_materialise': {'us_per_launch': us_nomat,
                                       'achieved': ALG_BYTES_PER_SAMPLE * B / (us_nomat * 1e-6) / 1e9,
                                       'frac': ALG_BYTES_PER_SAMPLE * B / (us_nomat * 1e-6) / 1e9 / peak}}
        roofline = gather_only
        # ---------------- the kernel the timed step actually launches for this stage: the one-kernel DeepFM forward
        # (rpb_deepfm_fwd_fused: gather + FM + layer-1 tcgen05 GEMM + tower tail + loss; 29 % of the step in
        # profiles/r01_bench_launches.csv).  Timed live as a graph-captured TRAINING forward (feature row and activations
        # stored for backward; the 3 us weight-split launch in front of it is inside the interval), same algorithmic
        # bytes as the stage it replaces.  If anything here fails the stand-alone gather kernel stays the
        # reported roofline kernel.
        try:
            fw, keep_out = [], []
            for cb in cbs:
                d = cb.as_dict()
                keep_out.append(model(d))
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    keep_out.append(model(d))
                fw.append(g)
            for i in range(3):
                fw[i % NB].replay()
            torch.cuda.synchronize()
            e0.record()
            for i in range(args.steps):
                fw[i % NB].replay()
            e1.record()
            torch.cuda.synchronize()
            us_fwd = e0.elapsed_time(e1) * 1e3 / args.steps
            ach = ALG_BYTES_PER_SAMPLE * B / (us_fwd * 1e-6) / 1e9
            roofline = {'kernel': 'deepfm_fwd_fused_kernel (gather + dense pack + FM + layer-1 tcgen05 GEMM + tower tail + BCE in one '
                                  'launch, training variant: x and activations stored), the forward of the timed step',
                        'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                        'traffic': 248.6e6, 'traffic_source': 'profiles/r01_deepfm_step_ncu_full.md (ncu --set full: dram__bytes_read 125.9 MB + write 122.7 MB per launch)',
                        'us_per_launch': us_fwd, 'alg_bytes_per_launch': ALG_BYTES_PER_SAMPLE * B, 'peak_source': peak_src,
                        'share_of_step': us_fwd / (1e3 * ms / args.steps),
                        'forward_only_samples_per_s': B / (us_fwd * 1e-6),
                        # the same launch against the bytes a TRAINING forward has to move: the algorithmic reads plus what it
                        # must leave behind for backward (feature row x, h1..h3, fm_s, logit, pred)
                        'with_saved_activations': (lambda bps: {'bytes_per_sample': bps, 'achieved': bps * B / (us_fwd * 1e-6) / 1e9,
                                                                'frac': bps * B / (us_fwd * 1e-6) / 1e9 / peak})(
                            ALG_BYTES_PER_SAMPLE + 4 * ((F * D + Nd + 3) // 4 * 4) + 4 * 64 * len(CFG['hidden']) + 4 * D + 8),
                        'note': 'bound by the shared-memory (MIO) pipe, not by HBM: the CUDA-core tower tail shares it with the '
                                'gather warps (profiles/r01_experiments.md); the stage alone is gather_only',
                        'gather_only': gather_only}
            del fw, keep_out
        except Exception as ex:      # keep the line: the stand-alone gather kernel remains the roofline kernel
            roofline = dict(gather_only, fused_forward_error=repr(ex))
            try:
                torch.cuda.synchronize()
            except Exception:
                pass


    line['roofline'] = roofline

    # ---------------- same step + optimizer: FusedAdam between backward and zero_grad — dense
    # parameters in one multi-tensor launch, table rows row-sparsely with the gradient re-zero fused in (rpb_sparse_adam),
    # step counter on the device so the captured graph keeps its bias correction.  Reported beside the headline.
    train_step = None
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
            train_step = {'value': B * args.steps / (ms_t * 1e-3), 'unit': 'samples/s', 'ms_per_step': ms_t / args.steps,
                          'what': 'forward + backward + FusedAdam (row-sparse Adam on the touched table rows, gradient re-zero fused)',
                          'gpu_launches_per_step': tsteps[0].launches_per_step, 'loss_after': float(tsteps[0].loss.item())}
            del tsteps, opt
        except Exception as ex:          # a secondary leg must never cost the headline line
            train_step = {'error': repr(ex)}

    line['train_step'] = train_step

    # ---------------- secondary legs (N = 1 only): Zipf(1.05)-distributed ids and stock PyTorch eager on the
    # same GPU (the oracle's functional restatement of the reference forward run on CUDA tensors = the "existing Blackwell
    # path" a user of the reference gets from `.to('cuda')`: F separate embedding lookups + stack + cat + addmm chain, dense
    # [V+1, D] table gradients from autograd)
    zipf = eager_gpu = None
    if world == 1 and not args.no_extras:
        try:
            import numpy as np
            rng = np.random.default_rng(SEED)
            zsteps = []
            for i in range(2):
                cb = ColumnarBatch(enc, B, device=dev, pinned_host=False)
                d = synth_batch(enc, B, gen, device=dev)
                for c in cb.sparse:
                    d[c] = torch.from_numpy(((rng.zipf(1.05, B) - 1) % (CFG['V'] + 1)).astype('int64')).to(dev)
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
            zipf = {'value': B * args.steps / (ms_z * 1e-3), 'unit': 'samples/s', 'ms_per_step': ms_z / args.steps,
                    'ids': 'zipf(1.05) - 1 mod (V+1) per field'}
            del zsteps
            import oracle
            sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
            dd = cbs[0].as_dict()
            n_e = 5
            for it in range(2 + n_e):
                if it == 2:
                    torch.cuda.synchronize()
                    e0.record()
                out = oracle.deepfm(sd, enc, dd, hidden_units=tuple(CFG['hidden']))
                out['loss'].backward()
                for v in sd.values():
                    v.grad = None
            e1.record()
            torch.cuda.synchronize()
            ms_e = e0.elapsed_time(e1)
            eager_gpu = {'value': B * n_e / (ms_e * 1e-3), 'unit': 'samples/s', 'ms_per_step': ms_e / n_e,
                         'what': 'stock PyTorch eager ops of the reference forward + autograd backward on the same B200 (fp32)'}
            del sd, out
            torch.cuda.empty_cache()
        except Exception as ex:
            if zipf is None:
                zipf = {'error': repr(ex)}
            else:
                eager_gpu = {'error': repr(ex)}

    line['zipf_ids'], line['torch_eager_gpu_baseline'] = zipf, eager_gpu
    if rank == 0:
        # opt-in variants not yet measured on hardware: in a child process with a hard timeout, only when this run has been
        # quick so far (a slow box must not be pushed past "minutes"), after every measurement of this process is final
        t_parent = time.time() - T_START
        if world == 1 and not args.no_experiments and not args.no_extras:
            if t_parent < 180:
                line['experiments'] = experiments_in_children(args.steps, deadline=T_START + 400)
                line['experiments']['wall_s'] = round(time.time() - T_START - t_parent, 1)
            else:
                line['experiments'] = {'skipped': f'this run had already taken {t_parent:.0f} s'}
        line['wall_s_before_cpu_baseline'] = round(time.time() - T_START, 1)
        if world == 1 and not args.no_cpu_baseline:
            try:
                torch.cuda.empty_cache()
            except Exception:
                pass
            try:
                r = cpu_reference_run(steps=60, warmup=1, B=CFG['B'], min_seconds=10.0)   # ~10 s of CPU work, <= 60 steps
                line['cpu_baseline'] = {'value': r['value'], 'unit': 'samples/s', 'cores': r['cores'], 'kind': 'port',
                                        'sample': r['sample']}
            except Exception as ex:
                line['cpu_baseline'] = {'error': repr(ex), 'kind': 'port'}
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


# ----------------------------------------------------------------------------------------------- experiments (child process)
def run_experiments(args):
    """`bench.py --experiment safe|tc_fwd|tc_bwd|all` — run by the default bench in CHILD processes (own CUDA context, hard
    timeout) after all of its own measurements are done, so that nothing here can touch the headline numbers.  Measures the
    opt-in variants that were written after the last GPU call of round 1 and have not run on hardware yet; prints one JSON
    object per finished leg (the last line is the result).  Groups: `safe` = everything that cannot trap (default-kernel
    checks, hints, address arithmetic), `tc_fwd` / `tc_bwd` = one new tcgen05 kernel each, in separate processes because a
    protocol bug there traps the CUDA context (`tc_bwd` also runs both together and `all_on` when told that `tc_fwd` passed).

    * zero_first: GraphedStep(zero_first=True) — the sparse re-zero of the table gradients overlapped with the next forward
      (each replay clears the rows ITS batch touched two replays earlier: same work per step, one step later).
    * afm_golden: the AFM class against the fixture of the real reference AFM (default kernels).
    * sharded_fused_local: the fused core on row-sharded tables (RPB_SHARDED_FUSED) with all shards on this one GPU
      (dist.LocalShards): parity of logits and of every gradient against the unsharded model.
    * l2_persist / all_on: an L2 persisting access-policy window on the feature row x (hint only), alone and with every
      variant that passed parity.
    * autoint_vec: the AutoInt attention kernels with float4 lane I/O at the config-4 shape: bit-identity, step time.
    * l2_fetch_32B: the default step with cudaLimitMaxL2FetchGranularity = 32 (aimed at the scatter epilogue's line fetches).
    * fused_tc_tail / tower_bwd_tc / both_tc: rpb_set_option(...) — the tower-tail layers of the one-kernel forward, and the dz
      chain of the tower-tail backward, on tcgen05: parity against the default kernels on the same batch (logit / loss /
      gradients), then step (and forward-only) timings."""
    from rec_pangu_b200 import ops, _lib
    from rec_pangu_b200.models.ranking import DeepFM
    from rec_pangu_b200.runtime import ColumnarBatch, GraphedStep
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    B, D = CFG['B'], CFG['D']
    enc = make_enc()
    torch.manual_seed(SEED)
    with torch.device(dev):
        model = DeepFM(embedding_dim=D, hidden_units=CFG['hidden'], enc_dict=enc)
    model.set_grad_mode('persistent')
    model.train()
    gen = torch.Generator(device=dev).manual_seed(SEED)
    NB = 2
    cbs = []
    for i in range(NB):
        cb = ColumnarBatch(enc, B, device=dev, pinned_host=False)
        cb.load_device(synth_batch(enc, B, gen, device=dev))
        cbs.append(cb)
    K = max(20, min(args.steps, 100))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def time_graphs(gs):
        for i in range(4):
            gs[i % NB].replay()
        torch.cuda.synchronize()
        e0.record()
        for i in range(K):
            gs[i % NB].replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K

    def fwd_graphs(infer=False):
        """Graph-captured forward per batch: training (grad enabled: x, activations stored) or inference (no_grad, is_training
        False: nothing materialised)."""
        gs, keep = [], []
        for cb in cbs:
            d = cb.as_dict()
            with (torch.no_grad() if infer else torch.enable_grad()):
                keep.append(model(d, is_training=not infer))
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    keep.append(model(d, is_training=not infer))
            gs.append(g)
        return gs, keep

    which = args.experiment if args.experiment in ('safe', 'tc_fwd', 'tc_bwd') else 'all'
    res = {'steps': K, 'group': which}

    def emit():                      # one line per finished leg: if the parent's timeout strikes, the last line is what it keeps
        print(json.dumps(res), flush=True)

    base = None
    try:
        base = [GraphedStep(model, cb) for cb in cbs]
        res['default_ms_per_step'] = time_graphs(base)
        fg, keep = fwd_graphs()
        res['default_fwd_us'] = 1e3 * time_graphs(fg)
        del fg, keep
        fg, keep = fwd_graphs(infer=True)
        res['default_fwd_infer_us'] = 1e3 * time_graphs(fg)
        del fg, keep
    except Exception as ex:
        res['default_error'] = repr(ex)
    emit()
    if which in ('safe', 'all'):
        try:
            zs = [GraphedStep(model, cb, zero_first=True) for cb in cbs]
            res['zero_first'] = {'ms_per_step': time_graphs(zs), 'launches_per_step': zs[0].launches_per_step}
            del zs
        except Exception as ex:
            res['zero_first'] = {'error': repr(ex)}
        try:
            # the rotated graphs leave the rows of the batch before last populated: start the parity leg from all-zero buffers
            model.zero_grad()
            for buf in model.embedding_layer._grad_store.buffers.values():
                buf.zero_()
            torch.cuda.synchronize()
        except Exception as ex:
            res['reset_error'] = repr(ex)
    emit()
    if which in ('safe', 'all'):
        # ---- AFM (the FiBiNet class under the reference's other name, added after the last GPU call): the fixture produced by
        # the real reference AFM, through the default kernels
        try:
            sys.path.insert(0, os.path.join(ROOT, 'tests'))
            from helpers import load_golden
            from rec_pangu_b200.models import ranking
            g_ = load_golden('afm')
            m_ = g_['meta']
            afm = getattr(ranking, m_['model'])(embedding_dim=m_['D'], enc_dict=m_['enc_dict'], **m_['kwargs'])
            afm.load_state_dict(g_['sd'])
            afm = afm.to(dev).eval()
            o_ = afm({k: v.to(dev) for k, v in g_['data'].items()})
            o_['loss'].backward()
            torch.cuda.synchronize()
            gr_ = dict(afm.named_parameters())
            res['afm_golden'] = {
                'max_abs_dpred': float((o_['pred'].cpu() - g_['out']['pred']).abs().max()),
                'dloss': abs(float(o_['loss'].item()) - float(g_['out']['loss'])),
                'max_rel_dgrad': max(float((gr_[k].grad.cpu() - v).abs().max() / v.abs().max().clamp_min(1e-12)) for k, v in g_['grad'].items())}
            res['afm_golden']['parity_ok'] = bool(res['afm_golden']['max_abs_dpred'] <= 1e-5 and res['afm_golden']['dloss'] <= 1e-5 and
                                                  res['afm_golden']['max_rel_dgrad'] <= 1e-4)
            del afm, o_, gr_
        except Exception as ex:
            res.setdefault('afm_golden', {})['error'] = repr(ex)
    emit()
    if which in ('safe', 'all'):
        # ---- fused core on row-sharded tables, checked on ONE GPU: dist.LocalShards keeps all G shards of every table on this
        # device, so the sharded variants of the one-kernel forward and of the dx-GEMM scatter epilogue see the same pointer
        # tables as over NVLink; a smaller vocabulary keeps the second copy of the tables cheap
        try:
            from rec_pangu_b200 import dist as rdist
            G_loc, V_loc, B_loc = 4, 50_000, 8192
            enc_s = {f'I{i + 1}': {'min': 0.0, 'max': 1.0} for i in range(CFG['Nd'])}
            enc_s.update({f'C{i + 1}': {'vocab_size': V_loc} for i in range(CFG['F'])})
            torch.manual_seed(3)
            with torch.device(dev):
                ref_m = DeepFM(embedding_dim=D, hidden_units=CFG['hidden'], enc_dict=enc_s)
                sh_m = DeepFM(embedding_dim=D, hidden_units=CFG['hidden'], enc_dict=enc_s)
            sh_m.load_state_dict({k: v.clone() for k, v in ref_m.state_dict().items()})
            ref_m.train()
            sh_m.train()
            ls = rdist.LocalShards(sh_m.embedding_layer, G_loc)
            sh_m.embedding_layer.attach_shards(ls)
            bt = synth_batch(enc_s, B_loc, gen, device=dev)
            out_r = ref_m(bt)
            out_r['loss'].backward()
            ops.SHARDED_FUSED = 1
            try:
                n0 = ops.launch_count()
                out_s = sh_m(bt)
                n_fwd = ops.launch_count() - n0
                out_s['loss'].backward()
            finally:
                ops.SHARDED_FUSED = 0
            torch.cuda.synchronize()
            ops.check_index_errors(dev)
            tg = 0.0
            for f, t in enumerate(ref_m.embedding_layer.tables()):
                tg = max(tg, float((ls.full_grad(f) - t.grad).abs().max() / t.grad.abs().max().clamp_min(1e-12)))
            dense_r = {n: p.grad for n, p in ref_m.named_parameters() if not n.startswith('embedding_layer.')}
            dg = max(float((p.grad - dense_r[n]).abs().max() / dense_r[n].abs().max().clamp_min(1e-12))
                     for n, p in sh_m.named_parameters() if not n.startswith('embedding_layer.'))
            sl = {'shards': G_loc, 'forward_launches': n_fwd,
                  'max_abs_dlogit': float((sh_m._last_logit - ref_m._last_logit).abs().max()),
                  'dloss': abs(float(out_s['loss'].item()) - float(out_r['loss'].item())),
                  'max_rel_dgrad_tables': tg, 'max_rel_dgrad_dense': dg}
            sl['parity_ok'] = bool(n_fwd == 2 and sl['max_abs_dlogit'] <= 1e-6 and sl['dloss'] <= 1e-6 and tg <= 1e-4 and dg <= 1e-4)
            res['sharded_fused_local'] = sl
            del ref_m, sh_m, ls, out_r, out_s
            torch.cuda.empty_cache()
        except Exception as ex:
            res.setdefault('sharded_fused_local', {})['error'] = repr(ex)
    emit()
    # ---- tcgen05 variants: parity first (eager, same batch, against the default kernels), then timings.  A protocol bug
    # traps the context (every mbarrier wait is bounded), which ends this process's measurements but nothing else.
    lib = _lib.load()
    d0 = cbs[0].as_dict()

    def eager_step(opts):
        for k in ('fused_tc_tail', 'tower_bwd_tc'):
            _lib.check(lib.rpb_set_option(k.encode(), 1 if k in opts else 0), f'rpb_set_option({k})')
        model.zero_grad()
        out = model(d0)
        out['loss'].backward()
        torch.cuda.synchronize()
        ops.check_index_errors(dev)
        gw = {n: p.grad.detach().clone() for n, p in model.named_parameters() if not n.startswith('embedding_layer.')}
        gt = model.embedding_layer.tables()[0].grad.detach().clone()
        return model._last_logit.clone(), float(out['loss'].item()), gw, gt

    ref = None
    variants = []
    if which in ('tc_fwd', 'all'):
        variants.append(('fused_tc_tail', ('fused_tc_tail',)))
    if which in ('tc_bwd', 'all'):
        variants.append(('tower_bwd_tc', ('tower_bwd_tc',)))
    if which == 'all' or (which == 'tc_bwd' and args.tc_fwd_ok):      # together only where the forward variant has passed parity
        variants.append(('both_tc', ('fused_tc_tail', 'tower_bwd_tc')))
    for name, opts in variants:
        try:
            if ref is None:
                ref = eager_step(())
            l0, loss0, gw0, gt0 = ref
            l1, loss1, gw1, gt1 = eager_step(opts)
            rel = {n: float((gw1[n] - gw0[n]).abs().max() / gw0[n].abs().max().clamp_min(1e-12)) for n in gw0}
            tc = {'max_abs_dlogit': float((l1 - l0).abs().max()), 'loss_default': loss0, 'loss_variant': loss1,
                  'max_rel_dgrad_dense': max(rel.values()), 'worst_dense_grad': max(rel, key=rel.get),
                  'max_rel_dgrad_table0': float((gt1 - gt0).abs().max() / gt0.abs().max().clamp_min(1e-12)),
                  'tolerance': 'north_star: |dlogit| <= 1e-4; gradients 5e-4 of the tensor maximum'}
            tc['parity_ok'] = bool(tc['max_abs_dlogit'] <= 1e-4 and abs(loss1 - loss0) <= 1e-5 and
                                   tc['max_rel_dgrad_dense'] <= 5e-4 and tc['max_rel_dgrad_table0'] <= 5e-4)
            res[name] = tc
            model.zero_grad()
            ts = [GraphedStep(model, cb) for cb in cbs]          # captured with the options on
            tc['ms_per_step'] = time_graphs(ts)
            tc['launches_per_step'] = ts[0].launches_per_step
            if 'fused_tc_tail' in opts:
                fg, keep = fwd_graphs()
                tc['fwd_us'] = 1e3 * time_graphs(fg)
                del fg, keep
                fg, keep = fwd_graphs(infer=True)
                tc['fwd_infer_us'] = 1e3 * time_graphs(fg)
                del fg, keep
                if name == 'fused_tc_tail':                  # per-role cycles of CTA 0 (rpb_debug_fused_trace): what bounds it now
                    import ctypes
                    lib.rpb_debug_fused_trace(None, 1)
                    model(d0)
                    torch.cuda.synchronize()
                    out16 = (ctypes.c_uint64 * 16)()
                    lib.rpb_debug_fused_trace(out16, 0)
                    names = ['kernel', 'gather_cp_wait', 'gather_wait_slot', 'gather_work', 'mma_wait_weights', 'mma_wait_operands',
                             'mma_wait_acc', 'mma_issue', 'epi_wait_acc', 'epi_rounds', 'epi_head', 'producer_wait']
                    tc['trace_cycles_cta0'] = {k: int(v) for k, v in zip(names, out16)}
            del ts
        except Exception as ex:
            res.setdefault(name, {})['error'] = repr(ex)
        emit()
    try:
        for k in (b'fused_tc_tail', b'tower_bwd_tc'):
            lib.rpb_set_option(k, 0)
        model.zero_grad()
    except Exception:
        pass
    emit()
    if which in ('safe', 'all') or res.get('both_tc', {}).get('parity_ok'):
        # ---- L2 persisting window on the feature row x (rpb_set_option('l2_persist', 1)): the forward kernel's stores of x and the
        # re-reads by the layer-1 weight gradient and the scatter epilogue carry an access-policy window (a hint: same results),
        # alone and then together with everything else that passed parity above
        try:
            _lib.check(lib.rpb_set_option(b'l2_persist', 1), 'rpb_set_option(l2_persist)')
            model.zero_grad()
            ps = [GraphedStep(model, cb) for cb in cbs]
            res['l2_persist'] = {'ms_per_step': time_graphs(ps), 'loss': float(ps[0].loss.item()),
                                 'loss_default': float(base[0].loss.item()) if base is not None else None}
            del ps
            if res.get('both_tc', {}).get('parity_ok') and res.get('tower_bwd_tc', {}).get('parity_ok'):
                for k in (b'fused_tc_tail', b'tower_bwd_tc'):
                    _lib.check(lib.rpb_set_option(k, 1), 'rpb_set_option')
                model.zero_grad()
                allon = [GraphedStep(model, cb, zero_first=True) for cb in cbs]
                res['all_on'] = {'ms_per_step': time_graphs(allon), 'loss': float(allon[0].loss.item()),
                                 'what': 'fused_tc_tail + tower_bwd_tc + l2_persist + zero_first'}
                del allon
        except Exception as ex:
            res.setdefault('l2_persist', {})['error'] = repr(ex)
        try:
            for k in (b'fused_tc_tail', b'tower_bwd_tc', b'l2_persist'):
                lib.rpb_set_option(k, 0)
            model.zero_grad()
            for buf in model.embedding_layer._grad_store.buffers.values():
                buf.zero_()
        except Exception:
            pass
    emit()
    if which in ('safe', 'all'):
        # ---- AutoInt attention kernels with float4 lane I/O (rpb_set_option('autoint_vec', 1)): BASELINE.json config 4 shape
        # (B = 32768, 26 fields, D = 32, 3 heads x 8) on a 100k-row vocabulary (the attention kernels do not see the vocabulary).
        # Same arithmetic in the same order, so everything must be bit-identical; eval() keeps dropout out of the comparison.
        try:
            from rec_pangu_b200.models.ranking import AutoInt
            enc_a = {f'I{i + 1}': {'min': 0.0, 'max': 1.0} for i in range(CFG['Nd'])}
            enc_a.update({f'C{i + 1}': {'vocab_size': 100_000} for i in range(CFG['F'])})
            torch.manual_seed(5)
            with torch.device(dev):
                am = AutoInt(embedding_dim=32, num_heads=3, enc_dict=enc_a)
            am.set_grad_mode('persistent')
            am.eval()
            B_a = 32768
            ab = ColumnarBatch(enc_a, B_a, device=dev, pinned_host=False)
            ab.load_device(synth_batch(enc_a, B_a, gen, device=dev))
            da = ab.as_dict()

            def a_step(flag):
                _lib.check(lib.rpb_set_option(b'autoint_vec', flag), 'rpb_set_option(autoint_vec)')
                am.zero_grad()
                out = am(da)
                out['loss'].backward()
                torch.cuda.synchronize()
                ops.check_index_errors(dev)
                return (out['pred'].detach().clone(), float(out['loss'].item()),
                        {n: p.grad.detach().clone() for n, p in am.named_parameters() if p.grad is not None and p.numel() < 10_000_000})

            p0, l0_, g0 = a_step(0)
            p1, l1_, g1 = a_step(1)
            av = {'pred_equal': bool(torch.equal(p0, p1)), 'dloss': abs(l1_ - l0_),
                  'max_rel_dgrad': max(float((g1[n] - g0[n]).abs().max() / g0[n].abs().max().clamp_min(1e-12)) for n in g0)}
            av['parity_ok'] = bool(av['pred_equal'] and av['dloss'] == 0.0 and av['max_rel_dgrad'] <= 1e-5)
            res['autoint_vec'] = av
            am.zero_grad()
            am.train()
            for flag, key in ((0, 'default_ms_per_step'), (1, 'ms_per_step')):
                _lib.check(lib.rpb_set_option(b'autoint_vec', flag), 'rpb_set_option(autoint_vec)')
                gs = GraphedStep(am, ab)
                for _ in range(3):
                    gs.replay()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(20):
                    gs.replay()
                e1.record()
                torch.cuda.synchronize()
                av[key] = e0.elapsed_time(e1) / 20
                del gs
            lib.rpb_set_option(b'autoint_vec', 0)
            del am, ab
            torch.cuda.empty_cache()
        except Exception as ex:
            res.setdefault('autoint_vec', {})['error'] = repr(ex)
    emit()
    if which in ('safe', 'all'):
        # ---- L2 fetch granularity 32 B (cudaLimitMaxL2FetchGranularity, device-wide hint): the scatter epilogue's `red`s fetch
        # whole 128-byte lines (340 MB read for 109 MB of reductions, profiles/r01_deepfm_step_ncu_full.md); replays of the
        # graphs captured above, so only the limit differs (runs after the tail leg; if that one trapped, this reports the error)
        try:
            if base is not None:
                lib = _lib.load()
                _lib.check(lib.rpb_set_option(b'l2_fetch_granularity', 32), 'rpb_set_option(l2_fetch_granularity)')
                res['l2_fetch_32B'] = {'ms_per_step': time_graphs(base)}      # last leg: the limit is not restored
        except Exception as ex:
            res['l2_fetch_32B'] = {'error': repr(ex)}
    print(json.dumps(res), flush=True)
    sys.stdout.flush()
    os._exit(0)            # a trapped context must not turn teardown into a hang


def _finite(o):
    """Replace non-finite floats by strings: the headline line must stay strict JSON whatever an experiment produced."""
    if isinstance(o, float):
        return o if o == o and o not in (float('inf'), float('-inf')) else str(o)
    if isinstance(o, dict):
        return {str(k): _finite(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_finite(v) for v in o]
    return o


def experiments_in_children(steps, deadline):
    """The three experiment groups, one child process each, as long as `deadline` (time.time()) allows.  Never raises."""
    out = {}
    for name, budget in (('safe', 110), ('tc_fwd', 80), ('tc_bwd', 100)):
        remaining = deadline - time.time()
        if remaining < 40:
            out[name] = {'skipped': 'time budget of the optional legs used up'}
            continue
        extra = ['--tc-fwd-ok'] if name == 'tc_bwd' and out.get('tc_fwd', {}).get('fused_tc_tail', {}).get('parity_ok') else []
        t0 = time.time()
        out[name] = experiments_in_child(steps, int(min(budget, remaining)), name, extra)
        out[name]['wall_s'] = round(time.time() - t0, 1)
    return out


def experiments_in_child(steps, budget_s=120, group='all', extra=()):
    """Run `bench.py --experiment <group>` in a child process; returns its JSON object or {'error': ...}.  Never raises."""
    import subprocess
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), '--experiment', group, '--steps', str(steps), *extra],
                           capture_output=True, text=True, timeout=budget_s)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith('{'):
                return _finite(json.loads(ln))
        return {'error': f'no result (rc {r.returncode}): ' + (r.stderr or '')[-300:]}
    except subprocess.TimeoutExpired as ex:            # keep what the child had finished (it prints one line per leg)
        out = ex.stdout.decode(errors='replace') if isinstance(ex.stdout, bytes) else (ex.stdout or '')
        for ln in reversed(out.strip().splitlines()):
            if ln.startswith('{') and ln.endswith('}'):
                try:
                    r = _finite(json.loads(ln))
                    r['timeout'] = f'child killed after {budget_s} s; legs finished until then are kept'
                    return r
                except Exception:
                    continue
        return {'error': f'timeout after {budget_s} s'}
    except Exception as ex:
        return {'error': repr(ex)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--eager', action='store_true', help='time eager launches instead of CUDA-graph replays')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train-step', action='store_true', help='skip the secondary forward+backward+optimizer timing')
    ap.add_argument('--no-extras', action='store_true', help='skip the Zipf-id and stock-PyTorch-eager-GPU secondary timings')
    ap.add_argument('--no-experiments', action='store_true', help='skip the child-process measurements of the opt-in variants')
    ap.add_argument('--experiment', default=None, help='internal: run the opt-in variant measurements (child process of the default bench): safe | tc_fwd | tc_bwd | all')
    ap.add_argument('--tc-fwd-ok', action='store_true', help='internal: the tc_fwd group passed parity, so tc_bwd may also run both tcgen05 variants together')
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.experiment is not None:
        run_experiments(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
