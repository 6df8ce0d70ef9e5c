#!/usr/bin/env python
"""bench.py — scenes/sec of the reverse-diffusion CCSP sampling loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                    # our arm (CUDA path)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # reference arm (CPU)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...       # N > 1 (one rank per GPU)

A "step" is one full `GaussianDiffusion.sample` of the workload BASELINE.json quotes the metric on
(configs[1]): RandomSplitQualitativeWorld, N=8 objects, T=1000 timesteps, ULA with K=10 steps per
timestep (11 000 denoiser evaluations), batch = 1024 scenes PER GPU (weak scaling: ONE global batch of
N x 1024 scenes goes through parallel.ShardedSampler — contiguous scene shards, no collective inside the
loop, one NCCL all-gather of the final poses per step).  At N > 1 the line also carries `strong`: a fixed
global batch of 1024 scenes over the N GPUs against the same batch on one GPU.

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import sys as _sys

if '--impl' in _sys.argv and 'reference' in _sys.argv:
    # the CPU arm uses every host core; torchrun exports OMP_NUM_THREADS=1, which would pin numpy's BLAS to one thread
    for _k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_k] = str(os.cpu_count() or 1)
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(world='RandomSplitQualitativeWorld', n_obj=8, T=1000, K=10, batch_per_gpu=1024)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--math', default=None, help='fp32 | tf32x3 | bf16x3 | tf32 | bf16 (default: best exact mode available)')
    ap.add_argument('--timesteps', type=int, default=WORKLOAD['T'], help='dev only: anything but 1000 is NOT the headline config')
    ap.add_argument('--batch', type=int, default=WORKLOAD['batch_per_gpu'], help='dev only')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-strong', action='store_true', help='N > 1: skip the fixed-global-batch (strong scaling) measurement')
    ap.add_argument('--no-configs', action='store_true', help='N = 1: skip the extra configs 1/3/4/5 keys')
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# shared: workload, algorithmic FLOPs
# --------------------------------------------------------------------------------------------------
def algorithmic_flops(E, n, P):
    """per denoiser evaluation: first-layer pose part + decoder per edge, pose encoder per node."""
    l1 = E * (2 * 512 * 512)
    dec = E * 2 * (2 * 256 * 128 + 2 * 128 * P)
    enc = n * (2 * P * 128 + 2 * 128 * 256)
    return dict(l1=l1, dec=dec, enc=enc, total=l1 + dec + enc)


def ncu_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture (profiles/ncu_traffic.json, written by scripts/summarize_ncu.py); None when no capture is committed."""
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(kernel_key)
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_sustained=d['bf16_tflops_sustained'], bf16_burst=d['bf16_tflops'], hbm=d['hbm_gbs'], source='measured (MEASURED_PEAKS.json)')
    return dict(bf16_sustained=1400.0, bf16_burst=1590.0, hbm=6650.0, source='fallback (B200_PROFILING.md)')


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, uuid):
        self.uuid, self.proc, self.lines = uuid, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', self.uuid, f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [s.strip() for s in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v == 'Active':
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm),
                    reasons=sorted(reasons))


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's OWN implementation (unmodified networks/{ddpm,denoise_fn}.py, staged under oracle/_ref by
# `python -m oracle.make_ref`, loaded through oracle/ref_shim.py) on all host cores; the numpy port only if it is absent
# --------------------------------------------------------------------------------------------------
def reference_sample_time(batch, sd, dims, mode, T_full, K, timesteps_sampled=1, device='cpu', warm=0):
    """Time `GaussianDiffusion.sample` of the UNMODIFIED reference for `timesteps_sampled` timesteps (each 1+K denoiser
    evaluations: p_sample + K ULA steps) at the full batch and extrapolate to T_full: the per-timestep cost does not
    depend on t (same graph, same K — SURVEY.md §8d).  The sampled timesteps are the LAST ones of the T_full cosine
    schedule (passed as `betas`).  Returns (seconds_per_full_run, seconds_measured, kind)."""
    import torch
    from oracle import ref_shim
    if not ref_shim.reference_available():
        full, dt = port_sample_time(batch, sd, dims, mode, T_full, K, timesteps_sampled)
        return full, dt, 'port'
    dfn, ddpm = ref_shim.load_reference()
    if device == 'cpu':
        torch.set_num_threads(os.cpu_count() or 1)
    m = dfn.ConstraintDiffuser(dims=dims, hidden_dim=256, input_mode=mode, EBM='ULA', device=device, verbose=False)
    betas = torch.tensor(ddpm.cosine_beta_schedule(T_full)[-timesteps_sampled:])
    gd = ddpm.GaussianDiffusion(m, timesteps=timesteps_sampled, EBM='ULA', samples_per_step=K, betas=betas,
                                step_sizes='2*self.betas')
    if device != 'cpu':
        gd = gd.to(device)
    gd.eval()
    missing, unexpected = gd.load_state_dict(sd, strict=False)
    assert not unexpected and not [k for k in missing if k.startswith('denoise_fn.')], (missing, unexpected)
    sync = (lambda: torch.cuda.synchronize()) if device != 'cpu' else (lambda: None)
    try:
        for _ in range(warm):
            gd.sample(batch)
        sync()
        t0 = time.perf_counter()
        gd.sample(batch)                           # the reference's public entry point (ddpm.py:342-351)
        sync()
        dt = time.perf_counter() - t0
    finally:
        torch.set_grad_enabled(True)               # p_sample_loop flips the global flag (ddpm.py:262-265)
    return dt / timesteps_sampled * T_full, dt, 'reference'


def port_sample_time(batch, sd, dims, mode, T_full, K, timesteps_sampled=1):
    """Fallback when oracle/_ref is missing: the numpy restatement (oracle/ccsp_oracle.py), same sampling protocol."""
    from oracle import ccsp_oracle as orc
    from diffusion_ccsp_b200 import synthetic
    den = orc.OracleDenoiser({k: v.numpy() for k, v in sd.items()}, dims, mode)
    gd = orc.OracleDiffusion(den, timesteps=timesteps_sampled, EBM='ULA', samples_per_step=K,
                             schedule={k: v[-timesteps_sampled:] for k, v in orc.make_schedule(T_full).items()})
    noise = synthetic.make_noise(timesteps_sampled, K, batch.num_nodes, dims[-1][0], seed=1).numpy()
    t0 = time.perf_counter()
    gd.p_sample_loop(batch, noise)
    dt = time.perf_counter() - t0
    return dt / timesteps_sampled * T_full, dt


def cpu_model():
    try:
        for ln in open('/proc/cpuinfo'):
            if ln.startswith('model name'):
                return ln.split(':', 1)[1].strip()
    except Exception:
        pass
    return 'unknown'


def host_threads():
    import torch
    return int(torch.get_num_threads())


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from diffusion_ccsp_b200 import scenes, synthetic
    mode, dims = 'qualitative', synthetic.DIMS['qualitative']
    sd = synthetic.load_trained_checkpoint()
    batch = scenes.qualitative_batch(args.batch, WORKLOAD['n_obj'])
    T, K = args.timesteps, WORKLOAD['K']
    times, kind = [], None
    for i in range(args.warmup + args.steps):
        full, _, kind = reference_sample_time(batch, sd, dims, mode, T, K, 1)
        if i >= args.warmup:
            times.append(full)
    sec = sum(times) / len(times)
    val = args.batch / sec
    cores = host_threads()
    sample = (f'each step = GaussianDiffusion.sample over 1 of {T} timesteps (11 denoiser evaluations) at the full batch of '
              f'{args.batch} scenes, extrapolated x{T} (per-timestep cost is t-independent); ms_per_step is the extrapolated '
              f'full-run time, not the wall time of the step')
    line = dict(impl='reference', metric='scenes_per_sec', value=val, unit='scenes/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='fp32', data='synthetic',
                config=config_dict(args, T, K, args.batch),
                cpu_baseline=dict(value=val, unit='scenes/s', cores=cores, kind=kind, sample=sample, cpu=cpu_model(),
                                  code=('unmodified networks/{ddpm,denoise_fn}.py via oracle/_ref, torch %s, %d threads'
                                        % (__import__('torch').__version__, cores)) if kind == 'reference' else
                                  'oracle/ccsp_oracle.py numpy port (oracle/_ref missing: run python -m oracle.make_ref)',
                                  note='one CPU replica of the per-GPU workload on rank 0, whatever --gpus is'),
                e2e=dict(value=val, unit='scenes/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def config_dict(args, T, K, B):
    """identical in both arms (the driver compares the dicts)"""
    return dict(workload=workload_name(args), timesteps=T, ula_steps=K, scenes_per_gpu=B,
                denoiser_evals_per_step=T * (1 + K))


def workload_name(args):
    return (f"RandomSplitQualitativeWorld N={WORKLOAD['n_obj']}, T={args.timesteps}, ULA K={WORKLOAD['K']}, "
            f"batch={args.batch} scenes/GPU (BASELINE.json configs[1])")


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from diffusion_ccsp_b200 import _abi, scenes, synthetic
    from diffusion_ccsp_b200.checker import SolvedChecker
    from diffusion_ccsp_b200.ddpm import GaussianDiffusion
    from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser
    from diffusion_ccsp_b200.parallel import ShardedSampler

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torchrun)'
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # keep stdout to the one JSON line: the image exports NCCL_DEBUG=VERSION, which makes NCCL printf its banner there
        if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
            del os.environ['NCCL_DEBUG']
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)

    math = args.math or default_math()
    mode, dims = 'qualitative', synthetic.DIMS['qualitative']
    P = dims[-1][0]
    sd = synthetic.load_trained_checkpoint()          # trained by this repo's own training step (scripts/train_fixture.py)
    T, K, B = args.timesteps, WORKLOAD['K'], args.batch
    den = ConstraintDiffuser(dims=dims, input_mode=mode, device=dev, verbose=False, math=math)
    gd = GaussianDiffusion(den, timesteps=T, EBM='ULA', samples_per_step=K).eval()
    gd.load_state_dict(sd, strict=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    uuid = uuid if uuid.startswith('GPU-') else 'GPU-' + uuid
    evals = T * (1 + K)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_run(sampler, steps, warmup, seed0, timing=False):
        """W warm-up + K timed steps of the PRODUCT's sharded path (parallel.ShardedSampler: shard -> local loop -> one all-gather),
        inputs resident in HBM (the shard's plan is compiled before the timed region); CUDA events, barrier + synchronize on
        both sides, max over ranks."""
        plan = den.plan_for(sampler.local)

        def step(i):
            flush.zero_()                                                  # L2 flush between steps
            return sampler.sample(seed=seed0 + i)                          # global poses on every rank

        for i in range(warmup):
            step(i)
        barrier()
        if timing:
            plan.set_timing(max(1, evals // 128))
        clk = ClockSampler(uuid)
        clk.start()
        _abi.reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            full = step(warmup + i)
        e1.record()
        barrier()
        clocks = clk.stop()
        launches = _abi.launch_count()
        ms = max_over_ranks(e0.elapsed_time(e1) / steps)
        tm = plan.get_timing() if timing else None
        plan.set_timing(0)
        return ms, full, clocks, launches, tm, plan

    # ---- headline: weak scaling, 1024 scenes per GPU — ONE global batch of world x B scenes through the sharded sampler --------
    global_batch = scenes.qualitative_batch(world * B, WORKLOAD['n_obj'], seed=0)
    sampler = ShardedSampler(gd, global_batch)
    batch = sampler.local
    n, E = batch.num_nodes, batch.num_edges
    ms, full, clocks, launches, tm, plan = timed_run(sampler, args.steps, args.warmup, 1000, timing=True)
    value = world * B / (ms / 1e3)
    out = full[sampler.n0:sampler.n1]
    free = out[~batch.mask.bool().to(dev)]
    fin = torch.isfinite(free)
    result_stats = dict(finite_frac=float(fin.float().mean()), max_abs=float(free[fin].abs().max()),
                        frac_in_unit_box=float((free.abs() <= 1.05).float().mean()))

    # ---- N1: solved fraction of the last step's output (SolvedChecker: one launch, poses stay on the device) -------------
    checker = SolvedChecker(batch, dims, mode, dev)
    solved = checker(out)
    torch.cuda.synchronize(dev)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(20):
        solved, counts = checker(out, return_counts=True)
    c1.record()
    torch.cuda.synchronize(dev)
    check_ms = c0.elapsed_time(c1) / 20
    st = torch.tensor([float(solved.sum()), float(solved.numel()), float((counts[:, 0] > 0).sum()), float((counts[:, 1] > 0).sum())],
                      device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(st)                                                 # end-of-run counters
    solved_frac = float(st[0] / st[1])
    solved_info = dict(solved_frac=solved_frac, solved_scenes=int(st[0]), scenes=int(st[1]), with_collisions=int(st[2]),
                       with_missing_constraints=int(st[3]), check_ms_per_batch=check_ms,
                       checker='k_check_solved (CUDA, 1 launch per batch; clamp + rows + SAT + 13 relations + set inclusion)',
                       weights='checkpoint trained by this repo (15k Adam steps of ccsp_train_step on 24k generated scenes): far from the '
                               "paper's 300k-step models, and N = 8 is beyond the sizes the reference trains on (2-5 tiles)")
    if rank == 0:
        # the same checkpoint in the regime the reference evaluates (N = 4, solve_csp.py test sets of 2-4 tiles): one try
        b4 = scenes.qualitative_batch(64, 4)
        s4 = SolvedChecker(b4, dims, mode, dev)(gd.sample(b4, seed=77))
        solved_info['n4'] = dict(scenes=64, solved_frac=float(s4.float().mean()), timesteps=T)
    barrier()

    # ---- strong scaling: a FIXED global batch of B scenes over the N GPUs (B / N per GPU), and the same batch on one GPU -----
    strong = None
    if world > 1 and not args.no_strong:
        gb1 = scenes.qualitative_batch(B, WORKLOAD['n_obj'], seed=0)
        ssteps = max(1, min(args.steps, 3))
        ms_n, _, _, _, _, _ = timed_run(ShardedSampler(gd, gb1), ssteps, 1, 3000)
        den.drop_plans()
        class _Solo:                                                        # every rank runs the WHOLE batch alone: the N = 1 time
            local = gb1

            @staticmethod
            def sample(seed):
                return gd.p_sample_loop(gb1, seed=seed)
        solo = _Solo()
        ms_1, _, _, _, _, _ = timed_run(solo, ssteps, 1, 4000)
        strong = dict(scaling='strong', global_batch=B, scenes_per_gpu=B // world, ms_per_step=ms_n, value=B / (ms_n / 1e3),
                      unit='scenes/s', ms_per_step_one_gpu=ms_1, speedup=ms_1 / ms_n, efficiency=ms_1 / ms_n / world, steps=ssteps,
                      ms_per_evaluation=ms_n / evals,
                      limit='per-launch fixed cost of the node + edge kernel pair (prologue, first ring fill, tail: ~0.03 ms per evaluation) '
                            'does not shrink with the shard')
        den.drop_plans()
        plan = den.plan_for(batch)

    # ---- e2e: public API with HOST buffers: plan build (H2D) + sample + D2H, wall clock ------------
    e2e = None
    if not args.no_e2e:
        pinned = scenes.SceneBatch(batch.x.pin_memory(), batch.edge_index.pin_memory(), batch.edge_attr.pin_memory(),
                                   batch.mask.pin_memory())
        host_out = torch.empty((n, P), dtype=torch.float32).pin_memory()

        e2e_parts = dict(plan_ms=0.0, sample_ms=0.0, d2h_ms=0.0)

        def e2e_step(i):
            ta = time.perf_counter()
            den.drop_plans()
            pl = den.plan_for(pinned)                                    # host batch -> HBM-resident plan (H2D inside)
            tb_ = time.perf_counter()
            res = gd.sample(pinned, seed=5000 + i, node_offset=sampler.n0)  # synchronises (reference bookkeeping)
            tc_ = time.perf_counter()
            host_out.copy_(res, non_blocking=False)
            td = time.perf_counter()
            e2e_parts['plan_ms'] += (tb_ - ta) * 1e3; e2e_parts['sample_ms'] += (tc_ - tb_) * 1e3; e2e_parts['d2h_ms'] += (td - tc_) * 1e3
            return pl.h2d_bytes

        e2e_step(0)
        for k in e2e_parts:
            e2e_parts[k] = 0.0
        barrier()
        clk2 = ClockSampler(uuid)
        clk2.start()
        t0 = time.perf_counter()
        nsteps = max(1, min(args.steps, 3))
        for i in range(nsteps):
            h2d = e2e_step(1 + i)
        barrier()
        dt = max_over_ranks((time.perf_counter() - t0) / nsteps)
        clocks2 = clk2.stop()
        e2e = dict(value=world * B / dt, unit='scenes/s', h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(n * P * 4),
                   ms_per_step=dt * 1e3, steps=nsteps, clocks=clocks2,
                   breakdown_ms={k: v / nsteps for k, v in e2e_parts.items()},
                   api='GaussianDiffusion.sample(batch) with host batch; plan rebuilt every step')

    # ---- the other BASELINE.json configs at their own per-GPU shard sizes, full T (N = 1 only; extra keys, not the headline) ----
    configs = None
    if world == 1 and not args.no_configs:
        configs = other_configs(dev, math, T, K, flush)

    if rank == 0:
        pk = peaks()
        fl = algorithmic_flops(E, n, P)
        s = max(tm['samples'], 1)
        l1_ms, dec_ms, node_ms = tm['ms_edge_l1'] / s, tm['ms_edge_dec'] / s, tm['ms_node'] / s
        fused = dec_ms == 0.0
        dom_flops = fl['l1'] + (fl['dec'] if fused else 0)
        achieved = dom_flops / (l1_ms * 1e-3) / 1e12 if l1_ms > 0 else None
        roofline = dict(bound='tensor', kernel='k_edge (first layer' + (' + decoder, fused)' if fused else ')'),
                        achieved=achieved, peak=pk['bf16_sustained'], unit='TFLOP/s',
                        frac=(achieved / pk['bf16_sustained']) if achieved else None,
                        traffic=ncu_traffic('k_edge_fused2_tc' if fused else 'k_edge_l1_tc'),
                        traffic_source='profiles/ncu_traffic.json (committed ncu --set full capture, not measured in this run)',
                        peak_source=pk['source'] + ': bf16_tflops_sustained (kernel timed inside a long step)',
                        algorithmic_flops_per_launch=dom_flops, avg_launch_ms=l1_ms, launches_sampled=tm['samples'],
                        note=('FP32 FMA validation path: tensor-core fraction is expected to be tiny' if math == 'fp32' else
                              '3-term split: algorithmic FLOPs counted once; ceiling = 1/3 of the tensor peak of the operand type'))
        line = dict(metric='scenes_per_sec', value=value, unit='scenes/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype=math, data='synthetic',
                    config=config_dict(args, T, K, B),
                    workload_detail=dict(nodes_per_gpu=n, edges_per_gpu=E, global_scenes=world * B,
                                         path='parallel.ShardedSampler: one global batch, contiguous scene shards, Philox keyed on the global node id, one all_gather_into_tensor per step',
                                         weights='diffusion_ccsp_b200/data/denoise_fn_qualitative_fp16.npz: 9.15 M parameters trained by this repo (scripts/train_fixture.py)',
                                         noise='in-kernel Philox4x32-10', l2='explicit 256 MiB flush between steps; static term + activations (2 x %d MB) exceed L2' % (plan.edge_rows * 512 * 4 >> 20)),
                    clocks=clocks, e2e=e2e, gpu_launches=int(launches), result=result_stats,
                    solved=solved_info, solved_scenes_per_sec=value * solved_frac,
                    roofline=roofline,
                    kernels=dict(avg_ms=dict(edge_l1=l1_ms, edge_dec=dec_ms, node=node_ms),
                                 share_of_step=dict(edge_l1=l1_ms * evals / ms, edge_dec=dec_ms * evals / ms, node=node_ms * evals / ms),
                                 shares_additive=False,
                                 shares_note='the sampling events serialise the PDL-overlapped node/edge pair, so the shares sum to more than 1'),
                    algorithmic_tflops=fl['total'] * evals * world / (ms * 1e-3) / 1e12)
        if strong is not None:
            line['strong'] = strong
        if configs is not None:
            line['configs'] = configs
        if world == 1 and not args.no_cpu_baseline:
            full_s, measured, kind = reference_sample_time(batch, sd, dims, mode, T, K, 1)
            line['cpu_baseline'] = dict(value=B / full_s, unit='scenes/s', cores=host_threads(), kind=kind, cpu=cpu_model(),
                                        sample=f'1 of {T} timesteps (11 denoiser evaluations, {measured:.1f} s) at the full batch, extrapolated x{T}')
            if kind == 'reference':
                # the tougher baseline of BASELINE.md §3: the same unmodified PyTorch code with device='cuda' on this B200
                try:
                    full_c, meas_c, _ = reference_sample_time(batch, sd, dims, mode, T, K, 2, device='cuda', warm=1)
                    line['reference_cuda'] = dict(value=B / full_c, unit='scenes/s', kind='reference',
                                                  sample=f'2 of {T} timesteps ({meas_c * 1e3:.0f} ms) after 1 warm-up sample, extrapolated',
                                                  code='unmodified networks/{ddpm,denoise_fn}.py, device=cuda, torch eager FP32 (allow_tf32 off)')
                except Exception as ex:      # never let the extra baseline break the bench line
                    line['reference_cuda'] = dict(unavailable=repr(ex)[:200])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def other_configs(dev, math, T, K, flush):
    """BASELINE.json configs 1, 3, 4, 5 at full T through the public API (one warm-up sample at T = 20, then one timed sample):
    device-timed scenes/s of ONE GPU's shard.  Weights are seeded-init (no checkpoint exists for these worlds): throughput only."""
    import gc
    import torch
    from diffusion_ccsp_b200 import scenes, synthetic
    from diffusion_ccsp_b200.ddpm import GaussianDiffusion
    from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser
    out = {}
    cases = [
        ('1', 'qualitative', False, lambda: scenes.qualitative_batch(8, 4), 100, 'RandomSplitQualitativeWorld N=4, T=100, batch=8'),
        ('3', 'diffuse_pairwise', False, lambda: scenes.make_batch('boxes', 4096, 12, seed=1), T, 'RandomSplitWorld N=12, batch=4096, 1 GPU'),
        ('4', 'diffuse_pairwise', True, lambda: scenes.make_batch('triangles', 1024, 10, seed=2), T, 'TriangularRandomSplitWorld N=10, batch 8192 / 8 GPUs = 1024 per GPU'),
        ('5', 'robot_box', False, lambda: scenes.make_batch('robot_box', 256, 6, seed=3), T, '3D panda-box packing N=6, batch 2048 / 8 GPUs = 256 per GPU'),
    ]
    for key, mode, tri, factory, Tc, desc in cases:
        dims = synthetic.dims_for(mode, tri)
        b = factory()
        den = ConstraintDiffuser(dims=dims, input_mode=mode, device=dev, verbose=False, math=math)
        sd = synthetic.make_state_dict(dims, mode, seed=0)
        warm = GaussianDiffusion(den, timesteps=20, EBM='ULA', samples_per_step=K).eval()
        warm.load_state_dict(sd, strict=False)
        warm.sample(b, seed=1)
        gd = GaussianDiffusion(den, timesteps=Tc, EBM='ULA', samples_per_step=K).eval()
        gd.load_state_dict(sd, strict=False)
        den.plan_for(b)
        # one evaluation at the last timestep: the model's [T, C, 512] time table for THIS T is built before the timed region
        den(torch.zeros((b.num_nodes, dims[-1][0])), b, torch.tensor([Tc - 1]), eval=True)
        gc.collect()                     # models / plans of earlier cases are destroyed (cudaFree of their cached blocks: 100s of ms) HERE,
        flush.zero_()                    # not by a collection that happens to run inside the timed region
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gd.p_sample_loop(b, seed=2)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        ev = Tc * (1 + K)
        fl = algorithmic_flops(b.num_edges, b.num_nodes, dims[-1][0])['total']
        out[key] = dict(workload=desc, timesteps=Tc, scenes_per_gpu=b.num_graphs, nodes=b.num_nodes, edges=b.num_edges, ms_per_step=ms,
                        ms_per_evaluation=ms / ev, value=b.num_graphs / (ms / 1e3), unit='scenes/s (one GPU, device-timed, 1 step)',
                        algorithmic_tflops=fl * ev / (ms * 1e-3) / 1e12)
        den.drop_plans()
        del den, gd, warm
    return out


def default_math():
    return 'bf16x3'


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
