#!/usr/bin/env python
"""Benchmark of the TeacherGNN aggregation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] --steps K --warmup W # CPU restatement of the reference path

A "step" is one full-graph training step of the 2-layer GCN teacher (Initial topology:
Linear F->H, 2 x GCNConv H->H with the Initial residual, Linear H->C): forward, NLL loss on the train
rows, backward, Adam step.  It aggregates 2*L*E edges (L forward + L transposed aggregations).
N=1 workload = BASELINE.json configs[3]: synthetic power-law graph, 10M nodes / 100M directed edges,
256-dim fp32.  N>1 = the same graph node-sliced over N GPUs ("strong" scaling); the per-aggregation exchange
is fused into the producing GEMM's epilogue (NVLink peer stores of the rows each peer gathers, --exchange push,
default) or an NCCL all-gather of the row blocks (--exchange nccl).

Prints ONE JSON line (rank 0).  Extra keys beside the driver contract: roofline (forward aggregation
kernel), roofline_kernels (every C-ABI kernel), cpu_baseline, e2e, clocks, gpu_launches.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'edges-aggregated/sec (2-layer GCN fwd+bwd)'
UNIT = 'edges/s'


def parse():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=10)
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    p.add_argument('--config', default='cfg4', choices=['cfg4', 'cfg5'],
                   help='cfg4 = BASELINE.json configs[3] (10M nodes / 100M edges / 256-dim fp32, 2 layers, the metric\'s '
                        'config; N>1 slices the same graph: strong scaling).  cfg5 = configs[4] (50M nodes / 1B edges / '
                        '128-dim bf16, 3-layer GCN + SE on every layer, 8 GPUs; generated shard by shard, 6.25M nodes / '
                        '125M edges per GPU: weak scaling when run on fewer GPUs)')
    p.add_argument('--nodes', type=int, default=0, help='0 = the config\'s size')
    p.add_argument('--edges', type=int, default=0, help='directed edges incl. one self loop per node; 0 = the config\'s')
    p.add_argument('--dim', type=int, default=0)
    p.add_argument('--classes', type=int, default=64)
    p.add_argument('--layers', type=int, default=0)
    p.add_argument('--se', default='', help='whetherHasSE flags (configs[3] has no SE; configs[4]: 010 = every layer '
                                            'of the Initial topology)')
    p.add_argument('--se-reg', type=float, default=0.5, help='coefficient of the SE regulariser in the loss')
    p.add_argument('--exchange', default='push', choices=['push', 'nccl'],
                   help='N>1: rows pushed to the peers from the producing GEMM epilogue (default) or NCCL all-gather')
    p.add_argument('--panels', type=int, default=0,
                   help='N>1, push: column panels the exchange is pipelined in against the aggregation '
                        '(0 = auto: 4 at 8 GPUs where the push dominates, else 1; measured, see DESIGN.md 7)')
    p.add_argument('--push-ctas', type=int, default=64, help='N>1, panels>1: grid cap of a pushing GEMM')
    p.add_argument('--src-panels', type=int, default=0,
                   help='source panels the neighbour lists are grouped by (1, 2, 4; 0 = auto = 1: plain edge-list order; '
                        'the same value must be used at every N for bit-identical sums)')
    p.add_argument('--passes', default='auto', choices=['auto', 'on', 'off'],
                   help='N>1, push, src-panels>1: pipeline every dense exchange by source panel at full row width '
                        '(producing GEMM per panel, aggregation pass per panel through a carry buffer).  auto = off: '
                        'measured slower than the one-launch exchange (2 x B200: 142 / 121 ms with 2 / 4 panels vs '
                        '86 ms, profiles/r02w_*), kept as a tested feature of the library')
    p.add_argument('--no-e2e', action='store_true')
    p.add_argument('--no-cpu-baseline', action='store_true')
    p.add_argument('--cpu-nodes', type=int, default=1_000_000, help='sample size of the CPU baseline / reference arm')
    p.add_argument('--cpu-steps', type=int, default=2)
    p.add_argument('--dropout', type=float, default=0.0,
                   help='TeacherGNN dropout (reference defaults are 0.2-0.6, base_options.py:20,190-220); > 0 takes the '
                        'layers off the pre-scaled hand-off, reported as a second line under profiles/')
    p.add_argument('--profile', default='', help='write a torch.profiler kernel table of 2 steps to this file')
    a = p.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1')) if a.impl == 'ours' else max(1, a.gpus)
    if a.config == 'cfg5':
        per_gpu_nodes, per_gpu_edges = 6_250_000, 125_000_000
        a.nodes = a.nodes or per_gpu_nodes * world
        a.edges = a.edges or per_gpu_edges * world
        a.dim, a.layers, a.se, a.storage = a.dim or 128, a.layers or 3, a.se or '010', 'bf16'
    else:
        a.nodes, a.edges = a.nodes or 10_000_000, a.edges or 100_000_000
        a.dim, a.layers, a.se, a.storage = a.dim or 256, a.layers or 2, a.se or '000', 'fp32'
    return a


def workload_name(a):
    return (f'synthetic power-law (gamma=2.5) N={a.nodes} E={a.edges} d={a.dim} {a.storage}, {a.layers}-layer GCN '
            f'(Initial topology, C={a.classes}, SE={a.se}' + (f', dropout={a.dropout}' if a.dropout else '') +
            '), fwd+bwd+Adam')


def workload_config(a, edges):
    """The ``config`` object of both arms: the workload only (implementation details live in ``impl_details``)."""
    es = 2 if a.storage == 'bf16' else 4
    return {'workload': workload_name(a), 'edges': int(edges), 'edges_aggregated_per_step': 2 * a.layers * int(edges),
            'l2': 'inputs exceed L2 (feature matrix %.1f GB vs 126 MB)' % (a.nodes * a.dim * es / 1e9)}


def model_args(a, n_nodes, device):
    from types import SimpleNamespace
    m = SimpleNamespace(type_trick='Initial', type_model='GCN', num_layers=a.layers, dim_hidden=a.dim,
                        num_feats=a.dim, num_classes=a.classes, dropout=a.dropout, res_alpha=0.1, layer_agg='concat',
                        transductive=True, N_nodes=n_nodes, device=device, dataset='synthetic',
                        dim_learnable_input=0, lamda=0.5, num_groups=None, skip_weight=None, graph_dropout=0.0,
                        layerwise_dropout=False, dim_commonEmb=a.classes)
    m.TeacherGNN = SimpleNamespace(whetherHasSE=[int(c) for c in a.se], change_to_featureless=False)
    return m


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    BAD = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20}
    NOTE = {'sw_power_cap': 0x4}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        return False

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


# ------------------------------------------------------------------------------------------------
# the reference arm / cpu_baseline leg: the oracle on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(a, steps, warmup):
    """Times the CPU restatement of the reference path (oracle/, DGL-style OpenMP CSR aggregation,
    torch CPU GEMMs) on a bounded sample of the workload: same generator, degree law, d, L and topology,
    fewer nodes (average degree kept).  Returns (edges_per_s, ms_per_step, sample description, threads)."""
    import torch
    from oracle import coldbrew_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n = min(a.cpu_nodes, a.nodes)
    e_dir = int(round(a.edges * (n / a.nodes)))
    und = max(1, (e_dir - n) // 2)
    ei = O.powerlaw_graph(n, und, seed=0)
    E = ei.shape[1]
    margs = O.make_args(type_trick='Initial', num_layers=a.layers, dim_hidden=a.dim, num_feats=a.dim,
                        num_classes=a.classes, N_nodes=n, whetherHasSE=a.se, dataset='Cora')
    torch.manual_seed(3)
    model = O.OracleTeacherGNN(margs, None)
    model.model.model.plan = O.CsrPlan(ei, n)
    x = torch.randn(n, a.dim, generator=torch.Generator().manual_seed(1))
    y = torch.randint(0, a.classes, (n,), generator=torch.Generator().manual_seed(2))
    idx = torch.arange(n // 10)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    model.train()

    def step():
        opt.zero_grad(set_to_none=True)
        loss = O.teacher_loss(model, x, ei, y, idx, 0.5)
        loss.backward()
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, steps)
    sample = (f'N={n} E={E} d={a.dim} L={a.layers} (1/{max(1, a.nodes // n)} of the workload, same degree law), '
              f'{steps} steps after {warmup} warm-up, OpenMP CSR aggregation + torch CPU GEMM')
    return 2 * a.layers * E / dt, dt * 1e3, sample, min(threads, O.c_oracle_threads())


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, a.steps)
    warm = max(0, a.warmup)       # the same warm-up count as the repo's arm
    # bounded sample: ~11 us of host time per node per step on 8 cores (less with more cores); the largest sample
    # that keeps the whole --steps/--warmup run inside a 4-minute box, capped at a fifth of the workload
    a.cpu_nodes = int(max(50_000, min(a.nodes // 5, 240.0 / ((steps + warm) * 11e-6))))
    val, ms, sample, threads = cpu_reference_run(a, steps, warm)
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': steps,
            'warmup': warm, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            # the same workload description as the repo's arm prints; what was actually timed is in cpu_baseline
            'config': workload_config(a, a.edges),
            'timed_on': {'what': 'bounded sample of the workload, see cpu_baseline.sample', 'sample_nodes': a.cpu_nodes,
                         'sample_fraction_of_workload': round(a.cpu_nodes / a.nodes, 5)},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from gnn_tail_generalization_b200 import _cabi, dist as cbdist, ops, synth
    from gnn_tail_generalization_b200.GNN_model.GNN_normalizations import TeacherGNN

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device; there is no CPU fallback for this path '
                         '(use --impl reference for the CPU restatement)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_kind = 'measured' if 'hbm_gbs' in peaks else 'fallback'

    # ---- workload (generated on the device) ------------------------------------------------------
    N, d, L = a.nodes, a.dim, a.layers
    bf16 = a.storage == 'bf16'
    st_dtype = torch.bfloat16 if bf16 else torch.float32
    es = 2 if bf16 else 4
    und = max(1, (a.edges - N) // 2)
    if a.config == 'cfg5':
        # configs[4]: every rank generates and keeps only its own in-edges (the 10^9-edge list never exists) and
        # builds its slice from them (cb_graph_create_local); the out-edges of a symmetric graph are the same list
        # with its rows swapped
        ei = synth.powerlaw_graph_sharded(N, und, rank, world, seed=0, device=dev)
        e_cnt = torch.tensor([ei.shape[1]], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(e_cnt)
        E = int(e_cnt)
        graph = cbdist.SlicedGraph(ei, N, rank, world, local_out_edges=ei.flip(0))
    else:
        ei = synth.powerlaw_graph(N, und, seed=0, device=dev)          # identical on every rank, sliced by the build
        E = ei.shape[1]
        if a.src_panels <= 0:
            a.src_panels = 1
        graph = cbdist.SlicedGraph(ei, N, rank, world, src_panels=a.src_panels)
    del ei
    a.src_panels = graph.src_panels
    use_passes = a.src_panels > 1 and world > 1 and a.passes == 'on'
    if use_passes:
        a.panels = 1
        a.push_ctas = a.push_ctas if a.push_ctas != 64 else 0     # the per-panel GEMMs run on the full grid by default
    if a.panels <= 0:
        # panels pay where the push dominates AND a panel row is still a decent gather: 4 x 256-byte rows at cfg4 on
        # 8 GPUs (measured: 48.2 -> 44.7 ms); a 128-dim bf16 row cut in 4 is 64 bytes per gather (cfg5, measured:
        # 220.6 ms with 4 panels)
        a.panels = 4 if (world >= 8 and d * es // 4 >= 256) else 1
    if world > 1 and a.exchange == 'push':
        try:
            graph.enable_push(d, a.panels, a.push_ctas, elem_bytes=es, src_passes=use_passes)
        except RuntimeError as e:   # raised on every rank together (PeerExchange); reported in config.parallelism
            if rank == 0:
                print(f'bench: {e}; using the NCCL all-gather exchange', file=sys.stderr, flush=True)
            a.exchange = 'nccl (peer buffers unavailable)'
    torch.cuda.empty_cache()
    lo, hi = graph.row_begin, graph.row_end
    rows = hi - lo
    if a.config == 'cfg5':
        x = synth.features(rows, d, seed=1000 + rank, device=dev).to(st_dtype)      # this rank's rows only
        y = synth.labels(rows, a.classes, seed=2000 + rank, device=dev)
    else:
        x = synth.features(N, d, seed=1, device=dev)[lo:hi].clone() if world > 1 else synth.features(N, d, 1, dev)
        y = synth.labels(N, a.classes, seed=2, device=dev)[lo:hi]
        x = x.to(st_dtype)
    n_train = N // 10
    if a.config == 'cfg5':
        idx = torch.arange(rows // 10, device=dev)                      # train rows = first 10% of every rank's rows
    else:
        idx = torch.arange(max(0, min(n_train, hi) - lo), device=dev)   # train rows = first 10% of the nodes
    torch.manual_seed(3)
    model = TeacherGNN(model_args(a, rows, str(dev)), None).to(dev)     # SE tables: this rank's rows (row-sharded)
    cbdist.broadcast_dense_params(model, world)     # the SE tables consume RNG: make the replicated weights equal
    cbdist.attach_graph(model, graph)
    model.train()
    se_coef = a.se_reg
    se_opt = None
    if any(c == '1' for c in a.se):
        # SE tables + Adam moments are the largest thing in HBM: one fused pass per table (cb_se_adam_step) instead
        # of autograd's norm backward + torch.optim.Adam's passes; bf16 forward reads a bf16 shadow of the fp32 master
        from gnn_tail_generalization_b200 import se_optim
        se_opt = se_optim.FusedSEAdam(model, lr=1e-3, se_reg=se_coef, shadow_dtype=torch.bfloat16 if bf16 else None)
        opt = torch.optim.Adam(se_opt.other_parameters(), lr=1e-3)
    else:
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            fn()
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    def allsum(t):
        t = t.detach().double().reshape(-1).clone()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t]

    # parity evidence the driver can compare across N: the forward pass at the INITIAL weights is bit-identical for
    # every world size (per-row arithmetic does not depend on the slicing), so its fp64 checksums must agree to
    # the last printed digit; the loss after the timed steps agrees to fp32 all-reduce reassociation (~1e-6).
    model.eval()
    with torch.no_grad():
        lg0 = model.get_3_embs(x, None, idx).emb4classi_full.double()
        logits_checksum = [float('%.12e' % v) for v in allsum(torch.stack([lg0.sum(), lg0.abs().sum(), (lg0 * lg0).sum()]))]
        del lg0
    model.train()

    last = {}

    def step(x=x):
        opt.zero_grad(set_to_none=True)
        res = model.get_3_embs(x, None, idx)
        nll = F.nll_loss(F.log_softmax(res.emb4classi.float(), 1), y[idx], reduction='sum') / n_train
        loss = nll if model.se_reg_all is None else nll + se_coef * model.se_reg_all
        loss.backward()
        cbdist.allreduce_dense_grads(model, world)
        opt.step()
        if se_opt is not None:
            se_opt.step()
        last['nll'] = nll.detach()
        return loss

    for _ in range(max(3, a.warmup)):
        step()
    launches0, exch0 = _cabi.launch_count(), graph.exchanged_bytes
    with ClockSampler(local) as clk:
        ms_step = timed(step, a.steps)
    launches = _cabi.launch_count() - launches0
    exch_per_step = (graph.exchanged_bytes - exch0) // a.steps
    value = 2 * L * E / (ms_step * 1e-3)
    nll_after = allsum(last['nll'])[0]        # global train NLL of the last timed step (each rank holds its share)

    if a.profile and rank == 0:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
        with open(a.profile, 'w') as f:
            f.write(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=90))

    # ---- per-kernel probe: CUDA events around every C-ABI launch, over a second timed region ----
    sink = []
    ops.set_timing_sink(sink)
    probe_steps = min(a.steps, 5)
    ms_probe = timed(step, probe_steps)
    ops.set_timing_sink(None)
    torch.cuda.synchronize()
    per = {}
    for name, e0, e1, nbytes, flops in sink:
        r = per.setdefault(name, {'launches': 0, 'ms': 0.0, 'bytes': 0, 'flops': 0})
        r['launches'] += 1
        r['ms'] += e0.elapsed_time(e1)
        r['bytes'] += nbytes
        r['flops'] += flops
    kernels = {}
    for name, r in per.items():
        avg_ms = r['ms'] / r['launches']
        gbs = r['bytes'] / r['launches'] / (avg_ms * 1e-3) / 1e9
        kernels[name] = {'launches_per_step': r['launches'] / probe_steps, 'avg_ms': round(avg_ms, 4),
                         'alg_bytes': r['bytes'] // r['launches'], 'achieved_gbs': round(gbs, 1),
                         'frac': round(gbs / hbm_peak, 4), 'share_of_step': round(r['ms'] / probe_steps / ms_probe, 4)}
        if r['flops']:
            # 3 TF32 tensor-core products per fp32 product: flops counts the MMA work actually issued
            kernels[name]['tensor_tflops_tf32'] = round(r['flops'] / r['launches'] / (avg_ms * 1e-3) / 1e12, 1)
    dom = kernels.get('agg_forward_bf16' if bf16 else 'agg_forward', {})
    roofline = {'bound': 'hbm', 'kernel': 'k_agg (cb_agg_forward_bf16)' if bf16 else 'k_agg (cb_agg_forward)', 'achieved': dom.get('achieved_gbs'),
                'peak': hbm_peak, 'peak_source': f'{peak_kind} copy bandwidth (MEASURED_PEAKS.json hbm_gbs)',
                'unit': 'GB/s', 'frac': dom.get('frac'), 'traffic': None,
                'avg_launch_ms': dom.get('avg_ms'), 'alg_bytes_per_launch': dom.get('alg_bytes')}
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
        if a.config == 'cfg4' and world == 1:      # the capture is of this kernel at this size
            roofline['traffic'] = tr.get('k_agg_forward_dram_bytes_per_launch')
            roofline['traffic_source'] = tr.get('source')
    except Exception:
        pass

    # ---- end to end: host features in pinned memory -> device every step, loss read back ---------
    # The step's features travel host -> device inside the timed region, every step.  Like any input
    # pipeline, the copy of step i+1 runs on a side stream into the second of two device buffers while
    # step i computes; the first copy of the timed region is fully exposed.
    e2e = None
    if not a.no_e2e:
        x_host = torch.empty((rows, d), dtype=st_dtype, pin_memory=True)
        x_host.copy_(x)
        result = torch.zeros(1, dtype=torch.float32, pin_memory=True)
        bufs = [x, torch.empty_like(x)]
        copy_stream = torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        state = {'i': 0, 'primed': False, 'left': 0}

        def upload(slot):
            copy_stream.wait_event(consumed[slot])          # the step that read this buffer has finished
            with torch.cuda.stream(copy_stream):
                bufs[slot].copy_(x_host, non_blocking=True)
                copied[slot].record(copy_stream)

        def e2e_step():
            cur = state['i'] & 1
            if not state['primed']:                          # first step of a region: nothing was prefetched
                upload(cur)
                state['primed'] = True
            state['left'] -= 1
            if state['left'] > 0:
                upload(cur ^ 1)                              # next step's features, overlapped with this step
            torch.cuda.current_stream().wait_event(copied[cur])
            loss = step(bufs[cur])
            consumed[cur].record()
            result.copy_(loss.detach().reshape(1), non_blocking=True)
            state['i'] += 1

        def e2e_region(steps):
            # exactly one upload per step; the first one of a region is fully exposed
            state['primed'], state['left'] = False, steps
            torch.cuda.synchronize()
            for ev in consumed:
                ev.record()
            return timed(e2e_step, steps)

        # the upload alone, every rank at once (no compute in flight): what the host side can deliver to N GPUs
        # concurrently.  Where this is below step-time bandwidth the e2e number is host-bound, not kernel-bound.
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(copy_stream):
            h0.record(copy_stream)
            for _ in range(3):
                bufs[1].copy_(x_host, non_blocking=True)
            h1.record(copy_stream)
        barrier()
        h2d_ms = torch.tensor([h0.elapsed_time(h1) / 3], device=dev)
        if world > 1:
            dist.all_reduce(h2d_ms, op=dist.ReduceOp.MAX)
        h2d_gbs = rows * d * es / (float(h2d_ms) * 1e-3) / 1e9

        e2e_region(2)
        e2e_steps = max(2, a.steps)
        ms_e2e = e2e_region(e2e_steps)
        e2e = {'value': 2 * L * E / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': rows * d * es * world,
               'd2h_bytes_per_step': 4 * world, 'ms_per_step': ms_e2e, 'steps': e2e_steps,
               'h2d_copies_in_region': e2e_steps,
               'h2d_alone_ms_per_rank': round(float(h2d_ms), 3), 'h2d_alone_gbs_per_rank': round(h2d_gbs, 1),
               'h2d_note': 'upload of one step\'s features timed alone with all ranks copying at once (max over ranks): '
                           'when it exceeds ms_per_step of the device-timed run the end-to-end step is bound by the '
                           'host -> device path, not by the kernels',
               'what': 'features copied from pinned host memory every step (double-buffered, the copy of step '
                       'i+1 overlaps step i on a side stream), loss read back to the host'}
        del x_host, bufs

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        del model, opt
        torch.cuda.empty_cache()
        v, ms_cpu, sample, threads = cpu_reference_run(a, a.cpu_steps, 1)
        cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample, 'ms_per_step': ms_cpu}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps,
                'warmup': max(3, a.warmup), 'ms_per_step': ms_step, 'higher_is_better': True,
                'scaling': 'weak' if a.config == 'cfg5' else 'strong',
                'vs_baseline': None, 'dtype': 'bf16' if bf16 else 'f32', 'data': 'synthetic',
                'config': workload_config(a, E),
                'impl_details': {
                    'parallelism': (f'node-slice x{world}, exchange={a.exchange}, panels={a.panels}, '
                                    f'source-panel passes={"%d" % a.src_panels if use_passes else "off"}' if world > 1
                                    else 'single GPU'),
                    'neighbour_order': (f'rows grouped by {a.src_panels} source panels (same order at every N)'
                                        if a.src_panels > 1 else 'edge-list order'),
                    'gemm': ('tcgen05 kind::f16 on bf16 operands as stored, fp32 accumulate in TMEM' if bf16 else
                             'tcgen05 3xTF32 split (fp32-class accuracy), fp32 accumulate in TMEM'),
                    'se_optimizer': ('cb_se_adam_step (fused Adam + ||E|| gradient' +
                                     (', fp32 master + bf16 shadow' if bf16 else '') + ')') if se_opt else None},
                'roofline': roofline, 'roofline_kernels': kernels, 'cpu_baseline': cpu, 'e2e': e2e,
                'clocks': clk.summary(), 'gpu_launches': launches,
                'parity': {'logits_checksum_initial_weights': logits_checksum,
                           'what': '[sum, sum|.|, sum of squares] in fp64 over the [N, C] logits of one forward at the '
                                   'seeded initial weights, all-reduced: identical for every --gpus N',
                           'train_nll_after_timed_steps': float('%.8e' % nll_after),
                           'steps_taken': max(3, a.warmup) + a.steps},
                'exchange_bytes_per_step_per_rank': exch_per_step}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
