#!/usr/bin/env python
"""Benchmark of the RegTR correspondence-prediction hot path (BASELINE.json metric:
pairs/sec on 3DMatch-size clouds).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--pairs B] [--config C]

One "step" = one pass of the hot path (pyramid -> KPConv encoder -> cross-encoder -> regressor ->
Kabsch) over one batch of B synthetic 3DMatch-shaped pairs per GPU.  Default workload: BASELINE
config 2 (one ~20k-point pair, batch 1, 1 GPU).  N>1 (torchrun): every rank runs its own pairs
(weak scaling, no data-path collective) and the poses are all-gathered once per step.

Prints ONE JSON line on rank 0 (see the keys in `main`).  Timing: CUDA events per step on the
launching stream, W >= 3 warm-up steps, L2 flushed between timed steps, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WEIGHT_SEED = 2024
POOL = 8          # distinct synthetic pairs rotated through the steps
_T0 = time.time()


def log(msg):
    if int(os.environ.get('RANK', 0)) == 0:
        print(f'[bench {time.time() - _T0:7.1f}s] {msg}', file=sys.stderr, flush=True)


def cpu_threads():
    """Threads for the CPU legs: all host cores up to 32 (beyond that the ATen ops of this workload,
    ~1 ms each, slow down from fork/join overhead); reported as `cores`."""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get('REGTR_CPU_THREADS', 32))))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', type=int, default=2, help='BASELINE.json config id (2..5)')
    ap.add_argument('--pairs', type=int, default=None, help='pairs per GPU per step (default: from config)')
    ap.add_argument('--cpu-baseline', type=int, default=1, help='time the CPU port beside the GPU run (N=1 only)')
    ap.add_argument('--checks', type=int, default=1, help='report pose error vs the oracle on one pair')
    ap.add_argument('--attention', default='fp32', choices=['fp32', 'bf16_tc'],
                    help="fp32: parity kernel (default, pose within 1e-4); bf16_tc: tcgen05 tensor-core core")
    ap.add_argument('--inflight', type=int, default=10,
                    help='independent pairs in flight per GPU (CUDA-graph executors on private streams); 1 = serial')
    ap.add_argument('--graph', type=int, default=1, help='1: CUDA-graph executor (GraphedRegTR); 0: eager forward')
    return ap.parse_args()


def pairs_per_gpu(args):
    if args.pairs:
        return args.pairs
    return {2: 1, 3: 8, 4: 8, 5: 4}.get(args.config, 1)


def workload_name(args, B):
    n_pts = '~30k' if args.config == 5 else '~20k'
    return f'BASELINE config {args.config}: synthetic 3DMatch-like pairs, {n_pts} pts/cloud, {B} pair(s)/GPU/step'


def make_pool(config_id, n, first=0):
    from regtr_b200.synthetic import make_batch
    cid = 2 if config_id in (2, 3, 4) else config_id
    b = make_batch(cid, n, first_pair=first)
    return list(zip(b['src_xyz'], b['tgt_xyz']))


# ------------------------------------------------------------------------------ clocks

class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                                          '-i', str(index), '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Start of the window whose samples are reported (the timed region)."""
        self.t0 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        t0 = getattr(self, 't0', 0.0)
        rows = [r for t, r in self.rows if t >= t0] or [r for _, r in self.rows[-3:]]
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------- CPU baseline

def cpu_forward_fn(cfg, sd):
    """The reference's CPU path, restated: its own C++ pre-processing core when oracle/_ref is
    built (kind 'port' either way: the PyTorch modules are the oracle's restatement)."""
    from oracle import pre, regtr_oracle
    ref = pre.RefCpp() if pre.RefCpp.available() else None

    def run(src, tgt):
        with torch.no_grad():
            if ref is not None:
                meta = pre.preprocess_refcpp(cfg, [src, tgt], ref)
            else:
                meta = pre.preprocess(cfg, [src, tgt])
            return regtr_oracle.forward(sd, cfg, [src], [tgt], meta=meta)['pose']
    return run, ('reference C++ core (oracle/_ref: nanoflann + grid_subsampling)' if ref is not None
                 else 'oracle C restatement (brute-force first-K)')


def time_cpu(cfg, sd, pool, n_pairs, warm=1):
    cores = cpu_threads()
    torch.set_num_threads(cores)
    run, pre_kind = cpu_forward_fn(cfg, sd)
    for i in range(warm):
        run(*pool[i % len(pool)])
    ts = []
    for i in range(n_pairs):
        t0 = time.perf_counter()
        run(*pool[i % len(pool)])
        ts.append(time.perf_counter() - t0)
    return dict(value=1.0 / statistics.median(ts), unit='pairs/s', cores=cores, kind='port',
                sample=f'{n_pairs} pair(s) of the same workload, median; 1 warm-up; pre-processing = {pre_kind}; '
                       f'torch CPU fp32 with {cores} threads'), ts


# ------------------------------------------------------------------------------- main

def kpconv_algorithmic_bytes(info):
    """bytes = Nq*K*4 + nnz*(12 + Cin*4) + Nq*12 + Nq*Cout*4 + 15*Cin*Cout*4 (fp32)."""
    nnz = int((info['idx'] < info['Ns']).sum().item())
    Nq, K, Cin, Cout = info['Nq'], info['K'], info['Cin'], info['Cout']
    return Nq * K * 4 + nnz * (12 + Cin * 4) + Nq * 12 + Nq * Cout * 4 + 15 * Cin * Cout * 4


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    from regtr_b200.config import get_config
    from regtr_b200.weights import random_state_dict
    cfg = get_config('3dmatch')
    sd = random_state_dict(cfg, WEIGHT_SEED)
    B = pairs_per_gpu(args)
    pool = make_pool(args.config, min(POOL, max(2, args.steps)))
    run, pre_kind = cpu_forward_fn(cfg, sd)
    cores = cpu_threads()
    torch.set_num_threads(cores)
    log(f'reference arm: {cores} threads of {os.cpu_count()} cores')
    for i in range(max(args.warmup, 1)):
        run(*pool[i % len(pool)])
    # each step = a bounded sample of the workload: ONE pair (the workload has B*N pairs/step); the run is
    # additionally bounded in time (REGTR_REF_BUDGET_S, default 120 s) so that large K still ends in minutes
    budget = float(os.environ.get('REGTR_REF_BUDGET_S', 120))
    t0 = time.perf_counter()
    done = 0
    for i in range(args.steps):
        run(*pool[i % len(pool)])
        done += 1
        if done >= 3 and time.perf_counter() - t0 > budget:
            break
    dt = time.perf_counter() - t0
    v = done / dt
    line = dict(metric='pairs/sec on 3DMatch-size clouds (~20k pts); pose err vs ref', value=v, unit='pairs/s',
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1000 * dt / done,
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                impl='reference', executed_steps=done,
                config=dict(workload=workload_name(args, B), sample='1 pair per step (bounded sample)'),
                cpu_baseline=dict(value=v, unit='pairs/s', cores=cores, kind='port',
                                  sample=f'{done} steps x 1 pair; pre-processing = {pre_kind}'),
                e2e=dict(value=v, unit='pairs/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        return run_reference(args, rank, world)

    import torch.distributed as dist
    from regtr_b200 import ops
    from regtr_b200.config import get_config
    from regtr_b200.dist import gather_poses
    from regtr_b200.regtr import GraphedRegTR, PipelinedRegTR, RegTR
    from regtr_b200.weights import random_state_dict

    assert torch.cuda.is_available(), 'bench.py --impl b200 needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'   # keep stdout to the single JSON line (NCCL prints its banner there)
        dist.init_process_group('nccl', device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    B = pairs_per_gpu(args)

    cfg = get_config('3dmatch')
    cfg.attention_impl = args.attention
    sd = random_state_dict(cfg, WEIGHT_SEED)
    model = RegTR(cfg).to(dev).eval()
    model.load_state_dict(sd, strict=True)
    runner = GraphedRegTR(model) if args.graph else model      # the public call a user makes
    depth = args.inflight if args.graph else 1
    pipe = PipelinedRegTR(model, depth) if depth > 1 else None

    # rank-local pool of distinct pairs (generated rank-locally)
    n_pool = max(POOL, B)
    pool = make_pool(args.config, n_pool, first=rank * n_pool)
    host = [(torch.from_numpy(s).pin_memory(), torch.from_numpy(t).pin_memory()) for s, t in pool]
    resident = [(s.to(dev), t.to(dev)) for s, t in host]

    def batch_at(step, src):
        ids = [(step * B + j) % n_pool for j in range(B)]
        return {'src_xyz': [src[i][0] for i in ids], 'tgt_xyz': [src[i][1] for i in ids]}, ids

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(step):
        batch, _ = batch_at(step, resident)
        out = runner(batch)
        return gather_poses(out['pose'], B * world) if world > 1 else out['pose']

    def e2e_step(step):
        batch, ids = batch_at(step, host)
        if not args.graph:
            batch = {k: [t.to(dev, non_blocking=True) for t in v] for k, v in batch.items()}
        out = runner(batch)                     # graph executor: packs the host clouds, one H2D copy
        if world > 1:
            return gather_poses(out['pose'], B * world).cpu(), ids
        return (out['pose_host'] if 'pose_host' in out else out['pose'].cpu()), ids

    sampler = ClockSampler(local) if rank == 0 else None     # started early: nvidia-smi needs ~0.1 s to report
    log(f'pool ready ({n_pool} pairs, {sum(len(a) + len(b) for a, b in pool) // n_pool} pts/pair); warm-up')
    # ---------------- warm-up
    for i in range(W):
        device_step(i)
    if pipe is not None:                        # capture every (slot, capacity bucket) graph before timing
        pipe.warm([batch_at(i, resident)[0] for i in range(-(-n_pool // B))])
        for i in range(2 * depth):
            pipe.submit(batch_at(i, host)[0])
        pipe.drain()
    barrier()

    def run_pipelined(src):
        """K steps with `depth` forwards in flight; returns (elapsed ms on the device, bytes in/out)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = []
        t_cpu = time.perf_counter()
        for i in range(K):
            done = pipe.submit(batch_at(W + i, src)[0], pre_hook=flush.zero_)
            if done is not None:
                outs.append(done['pose_host'].clone())
        t_cpu = time.perf_counter() - t_cpu
        log(f'host time in submit/result: {1e3 * t_cpu / K:.3f} ms/step')
        outs += [o['pose_host'].clone() for o in pipe.drain()]
        pipe.join()
        if world > 1:                           # the run's poses, gathered over the ranks (288 B / pair)
            gather_poses(torch.cat(outs, dim=1).to(dev, non_blocking=True), B * K * world)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), outs

    # ---------------- timed region: K steps, HBM-resident inputs
    if sampler:
        sampler.mark()
    launches0 = ops.LAUNCHES
    barrier()
    if pipe is None:
        evs = []
        for i in range(K):
            flush.zero_()                                   # L2 flush between timed steps (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            device_step(W + i)
            e1.record()
            evs.append((e0, e1))
        barrier()
        t_dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    else:
        t_dev_ms, _ = run_pipelined(resident)               # includes the per-step L2 flush kernels
        barrier()
    launches = ops.LAUNCHES - launches0
    log(f'device-resident: {t_dev_ms / K:.3f} ms/step')

    # ---------------- end-to-end: host buffers in, pose on the host out
    h2d = sum(host[j][0].numel() * 4 + host[j][1].numel() * 4 for j in batch_at(W, host)[1])
    if pipe is None:
        for i in range(2):
            e2e_step(i)
        barrier()
        evs = []
        d2h = 0
        for i in range(K):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pose_h, ids = e2e_step(W + i)
            e1.record()
            evs.append((e0, e1))
            d2h = pose_h.numel() * 4
        barrier()
        t_e2e_ms = sum(a.elapsed_time(b) for a, b in evs)
    else:
        run_pipelined(host)
        barrier()
        t_e2e_ms, outs = run_pipelined(host)
        d2h = outs[-1].numel() * 4
        barrier()
    clocks = sampler.stop() if sampler else None
    log(f'e2e: {t_e2e_ms / K:.3f} ms/step; clocks {clocks}')

    # ---------------- max over ranks
    t = torch.tensor([t_dev_ms, t_e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_dev_ms, t_e2e_ms = t.tolist()

    # ---------------- roofline of the dominant kernel (KPConv gather + contraction), rank 0
    roof = None
    if rank == 0:
        ops.KPCONV_TRACE = []
        nsteps = min(K, 5)
        for i in range(nsteps):                 # eager forward: the trace hooks live in ops.kpconv
            flush.zero_()
            model(batch_at(W + i, resident)[0])
        torch.cuda.synchronize()
        tr, ops.KPCONV_TRACE = ops.KPCONV_TRACE, None
        # Re-time the dominant kernel (k_kpconv_agg: neighbour gather + influence + aggregation) launch
        # by launch on its own stream: [L2 flush][event][kernel][event], 3 repetitions each, over the
        # 11 KPConv calls of the last traced step.  The flush (~80 us) hides the launch latency.
        last = tr[-11:] if len(tr) >= 11 else tr
        gather_ms = warm_ms = 0.0
        per_launch = []
        for _, _, info in last:
            best = best_w = None
            for _ in range(3):
                flush.zero_()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                ev[0].record()
                ops.kpconv_aggregate(*info['args'], row_flags=info.get('row_flags'))      # cold: L2 just flushed
                ev[1].record()
                ev[2].record()
                ops.kpconv_aggregate(*info['args'], row_flags=info.get('row_flags'))      # inputs L2-resident, as
                ev[3].record()                                                            # behind their producer
                torch.cuda.synchronize()
                t, tw = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
                best = t if best is None else min(best, t)
                best_w = tw if best_w is None else min(best_w, tw)
            gather_ms += best
            warm_ms += best_w
            per_launch.append(dict(Nq=info['Nq'], Cin=info['Cin'], us=round(best * 1e3, 1),
                                   us_l2_warm=round(best_w * 1e3, 1)))
        tot_ms = sum(a.elapsed_time(b) for a, b, _ in tr) / nsteps          # whole KPConv op incl. weight GEMM (eager)
        step_bytes = sum(kpconv_algorithmic_bytes(info) for _, _, info in last)
        peak, peak_src = peaks()
        ach = step_bytes / (gather_ms * 1e-3) / 1e9
        roof = dict(bound='hbm',
                    kernel='k_kpconv_agg_mma / k_kpconv_c1 (+k_row_flags where the flags are not fused upstream): KPConv neighbour '
                           'gather + kernel-point influence + aggregation, 11 launches/pair, each timed alone right after '
                           'an L2 flush; bytes = SURVEY 8d algorithmic bytes of the KPConv op',
                    achieved=ach, peak=peak, unit='GB/s', frac=ach / peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the 11 launches of one
                    # `ncu --set full` capture of this workload (profiles/r01_ncu_kpconv_agg_mma_summary.csv):
                    # 55.5 MB per pair against 590 MB algorithmic -- the 126 MB L2 absorbs the row reuse
                    traffic=55.5e6 / 11, traffic_unit='bytes/launch (ncu capture, batch 1)', peak_source=peak_src,
                    algorithmic_bytes_per_launch=step_bytes / max(len(last), 1),
                    algorithmic_bytes_per_step=step_bytes, gather_ms_per_step=gather_ms,
                    # the same launches with their inputs still in L2 (the state behind the producer kernel
                    # inside a forward); `achieved`/`frac` above are the conservative cold-L2 figures
                    gather_ms_per_step_l2_warm=warm_ms, frac_l2_warm=step_bytes / (warm_ms * 1e-3) / 1e9 / peak,
                    kpconv_op_ms_per_step_eager=tot_ms, per_launch=per_launch)

    # ---------------- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and args.cpu_baseline:
        log('cpu baseline ...')
        cpu, _ = time_cpu(cfg, sd, pool, n_pairs=10)        # ~10 s of CPU work on the box's cores
        log(f'cpu baseline {cpu["value"]:.3f} pairs/s on {cpu["cores"]} threads')

    # ---------------- pose error vs the oracle on one pair (reported, also covered by tests/)
    pose_err = None
    if rank == 0 and args.checks:
        try:
            from oracle import pre, regtr_oracle
            s, tg = pool[0]
            out = model({'src_xyz': [resident[0][0]], 'tgt_xyz': [resident[0][1]]})
            ref = regtr_oracle.forward(sd, cfg, [s], [tg], meta=pre.preprocess(cfg, [s, tg]))
            pose_err = float((out['pose'].cpu() - ref['pose']).abs().max())
        except Exception as exc:  # pragma: no cover
            pose_err = f'unavailable: {exc}'

    if rank == 0:
        pairs = B * world * K
        line = dict(
            metric='pairs/sec on 3DMatch-size clouds (~20k pts); pose err vs ref',
            value=pairs / (t_dev_ms * 1e-3), unit='pairs/s', n_gpus=world, steps=K, warmup=W,
            ms_per_step=t_dev_ms / K, higher_is_better=True, scaling='weak', vs_baseline=None,
            dtype='f32', data='synthetic',
            config=dict(workload=workload_name(args, B), pairs_per_gpu_per_step=B,
                        parallelism=f'pair-level data parallel x{world}', l2_flush_between_steps=True,
                        weights='seeded random init (no pretrained weights offline)', precision_mode='fp32 parity (3xTF32 tcgen05 GEMMs, 3xTF32 mma.sync attention + KPConv aggregation)' if args.attention == 'fp32'
                        else 'fast (3xTF32 GEMMs, bf16 tcgen05 attention core)',
                        executor=(f'cuda-graph, {depth} pairs in flight (PipelinedRegTR)' if depth > 1 else
                                  'cuda-graph (GraphedRegTR)') if args.graph else 'eager',
                        timed_region='K steps incl. one 256 MB L2-flush kernel per step' if depth > 1 else
                                     'per-step CUDA events, L2 flush between steps not timed'),
            e2e=dict(value=pairs / (t_e2e_ms * 1e-3), unit='pairs/s', h2d_bytes_per_step=h2d,
                     d2h_bytes_per_step=d2h, ms_per_step=t_e2e_ms / K),
            gpu_launches=launches, clocks=clocks, roofline=roof, cpu_baseline=cpu,
            pose_err_vs_oracle=pose_err, impl='b200')
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
