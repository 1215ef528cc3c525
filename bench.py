#!/usr/bin/env python
"""Headline benchmark: particle-steps/s and FTLE fields/s on the 0.25 deg ERA5-shape regional
configuration (BASELINE.json configs[1], "C2": 281x321, 6-hourly u,v, 48 h backward FTLE,
SETTLS_order=4, cubic interpolation, f64).

A *step* is one batch of B independent start times (a rolling series of B+8 six-hourly levels;
every window is exactly configs[1]; B defaults to 1184 = 8 windows per SM) pushed through the whole hot path:
prefilter + pack (new levels only once) -> departure-point integration -> fused FTLE epilogue.

  python bench.py --gpus N --steps K --warmup W            # this repository (CUDA, sm_100a)
  python bench.py --impl reference ...                     # the reference's CPU arithmetic (oracle port)

`value` is whole-job particle-steps/s with the winds already resident in HBM; `e2e` is the same
metric through the public rolling API with host (pinned) buffers, H2D and D2H inside the timed
region.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S_ORDER = 4
NT = 9                   # 48 h at 6 h
DT = -21600
METRIC = 'particle-steps/s'


DEFAULT_BATCH = {'C2': 1184, 'C4': 1184, 'C3': 8}
DEFAULT_BATCH_ROWBANDS = {'C2': 8, 'C4': 8, 'C3': 1}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=None,
                    help='start times per step and per GPU.  Default for C2/C4: 1184 = 8 windows per SM = one GPU\'s share '
                         '(1095) of a year of hourly start times sharded 8-way (BASELINE configs[3]) rounded up to whole waves '
                         'of 296; for C3 (1 M particles per window): 8')
    ap.add_argument('--xclamp', default='outer', choices=['outer', 'pointwise'],
                    help="x-boundary: 'outer' = what the reference executes (quirk Q6)")
    ap.add_argument('--order', type=int, default=3, choices=[1, 3])
    ap.add_argument('--precision', default='f64', choices=['f64', 'f32', 'f32fast'],
                    help="'f64' parity path (default, the headline); 'f32' f32 packed winds with f64 tap arithmetic; "
                         "'f32fast' f32 winds and f32 tap arithmetic (tolerance-tested fast path)")
    ap.add_argument('--workload', default='C2', choices=['C2', 'C3', 'C4'])
    ap.add_argument('--sharding', default='starts', choices=['starts', 'rowbands'],
                    help="multi-GPU sharding: 'starts' = independent start times per rank (weak scaling, the headline); "
                         "'rowbands' = every rank integrates a band of particle rows of the SAME windows (strong scaling, BASELINE "
                         "configs[2]: use with --workload C3)")
    ap.add_argument('--host-dtype', default='f64', choices=['f64', 'f32'],
                    help="dtype of the host winds of the end-to-end path; 'f32' (how ERA5 is stored) halves the upload but then the "
                         "f64 path follows the reference's f32 dtype propagation (round32: two gathers per SETTLS stage)")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--chunk', type=int, default=148, help='windows per launch in the pipelined end-to-end path')
    args = ap.parse_args()
    if args.batch is None:
        args.batch = (DEFAULT_BATCH_ROWBANDS if args.sharding == 'rowbands' else DEFAULT_BATCH)[args.workload]
    return args


def workload(name):
    from lagrangiancoherence_b200 import synthetic as S
    if name == 'C2':
        lat, lon = S.grid_c2()
        return lat, lon, NT, DT, 'C2 regional 0.25deg 281x321, 6-hourly, 48 h backward (nt=9)'
    if name == 'C4':
        lat, lon = S.grid_c2()
        return lat, lon, 49, -3600, 'C4 rolling series on the C2 grid 281x321, hourly, 48 h backward (nt=49)'
    lat, lon = S.grid_c3()
    return lat, lon, 73, -3600, 'C3 near-global 0.25deg 721x1440, hourly, 72 h backward (nt=73)'


def config_dict(args, desc, B, world):
    if args.sharding == 'rowbands':
        shard = f'row bands x{world} (2-row recomputed halo, winds replicated' + ('' if world == 1 else ', bands gathered inside the timed region') + ')'
        what = f'step = {B} field(s), particle rows split over the GPUs'
    else:
        shard = f'start-times x{world}' + ('' if world == 1 else ', finished fields gathered on rank 0 inside the timed region ('
                                                    + os.environ.get('LCS_BENCH_GATHER', 'p2p') + ')')
        what = f'step = {B} rolling start times per GPU'
    return {'workload': f'{desc}, SETTLS_order={S_ORDER}, interp_order={args.order}, {args.precision} winds, '
                        f'xclamp={args.xclamp}; {what}',
            'grid': desc.split(',')[0], 'windows_per_step_per_gpu': B, 'xclamp': args.xclamp,
            'interp_order': args.order, 'settls_order': S_ORDER, 'sharding': shard,
            'l2': 'flushed between timed steps (256 MiB write)'}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for t, line in self.rows:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                clk, mx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(clk)
                for n, val in zip(names, parts[2:6]):
                    if val.lower().startswith('active'):
                        reasons.add(n)
        if not sm:      # timed region shorter than the sampling period: take the nearest samples
            sm = [float(l.split(',')[0]) for _, l in self.rows[-3:] if l.split(',')[0].strip().replace('.', '').isdigit()]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ----------------------------------------------------------------------------- reference arm / CPU baseline
def _oracle_window(job):
    """One configs[1] window through the CPU oracle (scipy map_coordinates + numba-typed stencil +
    scipy.linalg.norm, i.e. the reference's arithmetic without xarray bookkeeping)."""
    s, name, order, xclamp = job
    from oracle import lcs_oracle as O
    from lagrangiancoherence_b200 import synthetic as S
    lat, lon, nt, dt, _ = workload(name)
    nt = cpu_levels(name, nt)
    u, v = S.era5_like_winds(lat, lon, nt, t0=s)
    t = time.perf_counter()
    O.lcs_field(u, v, lat, lon, dt, SETTLS_order=S_ORDER, traj_interp_order=order, xclamp=xclamp)
    return time.perf_counter() - t


def cpu_levels(name, nt):
    """Wind levels of one CPU sample window: the whole window, except on C3 where one 73-level window of 1 M particles
    is minutes of scipy time per core -- there the sample is the first 2 intervals (cost is linear in the intervals)."""
    return min(nt, 3) if name == 'C3' else nt


def cpu_pool_run(name, order, xclamp, cores, nwindows, first=0):
    """`nwindows` independent windows over `cores` worker processes; returns wall seconds."""
    import multiprocessing as mp
    jobs = [(first + i, name, order, xclamp) for i in range(nwindows)]
    t = time.perf_counter()
    if cores == 1:
        for j in jobs:
            _oracle_window(j)
    else:
        with mp.get_context('fork').Pool(cores) as pool:
            pool.map(_oracle_window, jobs, chunksize=1)
    return time.perf_counter() - t


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    lat, lon, nt, dt, desc = workload(args.workload)
    cores = os.cpu_count() or 1
    psteps_window = lat.size * lon.size * (cpu_levels(args.workload, nt) - 1)
    for _ in range(args.warmup):
        cpu_pool_run(args.workload, args.order, args.xclamp, cores, cores)
    times = [cpu_pool_run(args.workload, args.order, args.xclamp, cores, cores, first=100 * (k + 1))
             for k in range(args.steps)]
    total = float(np.sum(times))
    value = psteps_window * cores * args.steps / total
    sample = (f'bounded sample of the config above: {cores} of its windows per step, one per worker process '
              f'({cores} processes), {args.steps} steps')
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'particle-steps/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config_dict(args, desc, args.batch, max(args.gpus, 1)),
            'fields_per_s': cores * args.steps / total,
            'cpu_baseline': {'value': value, 'unit': 'particle-steps/s', 'cores': cores, 'kind': 'port',
                             'sample': sample},
            'e2e': {'value': value, 'unit': 'particle-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0,
            'note': 'oracle port of the reference arithmetic (scipy.ndimage.map_coordinates + numba-typed stencil + '
                    'scipy.linalg.norm); the reference itself needs xarray, absent from this image'}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- CUDA arm
def ncu_traffic(args, B):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    `ncu --set full` capture of this very command line (profiles/r02_traffic.json); None for any other config."""
    try:
        table = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))
    except OSError:
        return None
    key = f'{args.workload}:{args.xclamp}:p{args.order}:{args.precision}:B{B}:{args.sharding}'
    entry = table.get(key)
    return entry['dram_bytes_per_launch'] if entry else None


def measured_peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        peaks = {}
    if 'hbm_gbs' in peaks:
        return peaks['hbm_gbs'], 'MEASURED_PEAKS.json (measured copy)'
    return 6650.0, 'fallback 6650 GB/s (B200_PROFILING.md)'


def run_b200(args):
    import torch
    import torch.distributed as dist
    from lagrangiancoherence_b200 import synthetic as S, _lib, rolling
    from lagrangiancoherence_b200.engine import FtleEngine, _ptr, _stream, precision_args

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    from lagrangiancoherence_b200.affinity import bind_host_to_gpu
    numa = bind_host_to_gpu(local) if world > 1 and not os.environ.get('LCS_NO_NUMA_BIND') else None
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    lat, lon, nt, dt, desc = workload(args.workload)
    B = args.batch
    rowbands = args.sharding == 'rowbands'
    nlev = B + nt - 1
    npts = lat.size * lon.size
    # start-time sharding: rank r owns start times [r*B, (r+1)*B) of one long synthetic series (weak scaling);
    # row bands: every rank holds the SAME B windows and integrates its band of particle rows (strong scaling)
    u, v = S.era5_like_winds(lat, lon, nlev, t0=0 if rowbands else rank * B)
    host_dtype = np.float32 if (args.precision != 'f64' or args.host_dtype == 'f32') else np.float64
    u, v = u.astype(host_dtype), v.astype(host_dtype)
    h_u = torch.from_numpy(u).pin_memory()
    h_v = torch.from_numpy(v).pin_memory()
    d_u, d_v = h_u.to(dev), h_v.to(dev)
    eng = FtleEngine(lat, lon, dt, SETTLS_order=S_ORDER, interp_order=args.order, xmode=args.xclamp,
                     device=dev, **precision_args(args.precision))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    adv_ms = []

    # ---- the gather of finished fields onto rank 0 (the only exchange on this path)
    gather_mode = os.environ.get('LCS_BENCH_GATHER', 'p2p')          # 'p2p' (NVLink peer copies) | 'nccl'
    peer = None
    gathered = None
    if world > 1 and not rowbands:
        if gather_mode == 'p2p':
            # symmetric memory needs CUDA VMM handle exchange between the ranks; where the box forbids it every rank
            # falls back to NCCL (agreed through an all_reduce so that no rank is left behind)
            try:
                from lagrangiancoherence_b200.peer import PeerFields
                peer = PeerFields([B] * world, lat.size, lon.size, device=dev, dst=0)
            except Exception as exc:                                    # noqa: BLE001
                sys.stderr.write(f'[bench] rank {rank}: peer gather unavailable ({exc!r}); using NCCL gather\n')
                peer = None
            agree = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)
            if int(agree.item()) == 0:
                peer, gather_mode = None, 'nccl'
        if peer is None and rank == 0:
            gathered = torch.empty((world, B, lat.size, lon.size), dtype=torch.float64, device=dev)
    x = torch.empty((B, lat.size, lon.size), dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    sigma_out = torch.empty_like(x)

    def step_starts(timed):
        st = eng.stage(d_u, d_v, reuse=True)
        if world == 1:
            a, b = ev(), ev()
            a.record()
            eng.advect(st, nsteps=nt - 1, nwindows=B, out=(x, y))
            b.record()
            sigma = eng.epilogue(x, y, out=sigma_out)
            if timed:
                adv_ms.append([(a, b)])
            return sigma
        # N > 1: the batch goes in parts (whole waves of 296 windows where B allows, else halves) so that the push of
        # a finished part's fields to rank 0 (NVLink copy engines, or NCCL gather) runs while the next part is integrated
        nparts = B // 296 if (B % 296 == 0 and B >= 592) else (2 if B >= 2 else 1)
        bounds = [B * i // nparts for i in range(nparts + 1)]
        works, evs, sig = [], [], []
        for lo, n in ((bounds[i], bounds[i + 1] - bounds[i]) for i in range(nparts)):
            a, b = ev(), ev()
            a.record()
            eng.advect(st, nsteps=nt - 1, nwindows=n, level0=lo, out=(x[lo:lo + n], y[lo:lo + n]))
            b.record()
            evs.append((a, b))
            sig.append(eng.epilogue(x[lo:lo + n], y[lo:lo + n], out=sigma_out[lo:lo + n]))
            if peer is not None:
                peer.push(lo, sig[-1])
            else:
                dst_list = [gathered[r, lo:lo + n] for r in range(world)] if rank == 0 else None
                works.append(dist.gather(sig[-1], dst_list, dst=0, async_op=True))
        for w in works:
            w.wait()
        if peer is not None:
            torch.cuda.current_stream(dev).wait_stream(peer.side)       # the group barrier is the timed region's own
        if timed:
            adv_ms.append(evs)
        return sig

    def step_bands(timed):
        st = eng.stage(d_u, d_v, reuse=True)                            # replicated: particles of any band roam the whole domain
        a, b = ev(), ev()
        a.record()
        out0, out1, in0, in1 = rolling.shard_rows(lat.size, world, rank)
        xb, yb = eng.advect(st, nsteps=nt - 1, nwindows=B, rows=(in0, in1), xrank=rolling.band_xrank(eng, world, B))
        b.record()
        band = eng.epilogue(xb, yb, in_row0=in0, out_rows=(out0, out1))
        if timed:
            adv_ms.append([(a, b)])
        return rolling.gather_bands(band, lat.size) if world > 1 else band

    step = step_bands if rowbands else step_starts

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    last = None
    for _ in range(args.warmup):
        last = step(False)                          # held like the timed steps hold it: the caching allocator reaches its
    barrier()                                       # steady state here, not with a cudaMalloc inside the second timed step
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.lcs_kernel_launches()
    mallocs0 = torch.cuda.memory_stats(dev).get('num_device_alloc', 0)
    t_wall0 = time.time()
    step_ms = []
    for _ in range(args.steps):
        flush.fill_(1)                              # evict L2 between timed iterations
        barrier()
        a, b = ev(), ev()
        a.record()
        last = step(True)
        b.record()
        barrier()
        step_ms.append(a.elapsed_time(b))
    t_wall1 = time.time()
    launches_timed = int(lib.lcs_kernel_launches() - launches0)
    mallocs_timed = int(torch.cuda.memory_stats(dev).get('num_device_alloc', 0) - mallocs0)
    eng.check_finite()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    total_ms = float(np.sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    advect_ms = float(np.mean([sum(a.elapsed_time(b) for a, b in evs) for evs in adv_ms]))
    work_ranks = 1 if rowbands else world                           # row bands split ONE batch; start times add batches
    psteps_step = work_ranks * B * npts * (nt - 1)
    value = psteps_step * args.steps / (total_ms * 1e-3)

    # ---- outside the timed region: rank 0 checks what it gathered against its own single-GPU recomputation
    gather_check = None
    if world > 1 and rank == 0:
        if rowbands:
            full = last
            st = eng.stage(d_u, d_v)
            xs, ys = eng.advect(st, nsteps=nt - 1, nwindows=B)
            ref = eng.epilogue(xs, ys)
            gather_check = {'what': 'row bands gathered on rank 0 vs the whole field integrated by rank 0 alone',
                            'fields': int(B), 'bit_identical': bool(torch.equal(full, ref))}
        else:
            got = peer.local if peer is not None else gathered
            ok, checked = True, 0
            for r in range(world):
                for s_ in sorted({0, B // 2, B - 1}):
                    uu, vv = S.era5_like_winds(lat, lon, nt, t0=r * B + s_)
                    uu, vv = uu.astype(host_dtype), vv.astype(host_dtype)
                    st = eng.stage(torch.from_numpy(uu).to(dev), torch.from_numpy(vv).to(dev))
                    xs, ys = eng.advect(st, nsteps=nt - 1, nwindows=1)
                    ok = ok and bool(torch.equal(eng.epilogue(xs, ys)[0], got[r, s_]))
                    checked += 1
            gather_check = {'what': 'fields gathered on rank 0 (%s) vs single-window recomputation on rank 0' % gather_mode,
                            'windows': checked, 'bit_identical': ok}

    # ---- end to end through the public API: pinned host winds in, host fields out, every step
    e2e_ms = []
    if rowbands:
        h_out = torch.empty((B, lat.size, lon.size), dtype=torch.float64).pin_memory() if rank == 0 else None
        bu, bv = torch.empty_like(d_u), torch.empty_like(d_v)
        for i in range(args.warmup + args.steps):
            flush.fill_(1)
            barrier()
            a, b = ev(), ev()
            a.record()
            bu.copy_(h_u, non_blocking=True)
            bv.copy_(h_v, non_blocking=True)
            full = rolling.ftle_row_bands(eng, bu, bv, nwindows=B) if world > 1 else eng.epilogue(*eng.advect(eng.stage(bu, bv), nsteps=nt - 1, nwindows=B))
            if rank == 0:
                h_out.copy_(full, non_blocking=True)
            b.record()
            barrier()
            if i >= args.warmup:
                e2e_ms.append(a.elapsed_time(b))
        api = 'lagrangiancoherence_b200.rolling.ftle_row_bands (pinned host winds in on every rank, gathered field out on rank 0)'
        h2d, d2h = 2 * nlev * npts * u.itemsize, B * npts * 8
    else:
        h_out = torch.empty((B, lat.size, lon.size), dtype=torch.float64).pin_memory()
        for i in range(args.warmup + args.steps):
            flush.fill_(1)
            barrier()
            a, b = ev(), ev()
            a.record()
            rolling.rolling_ftle(h_u, h_v, lat, lon, nt, dt, SETTLS_order=S_ORDER, interp_order=args.order,
                                 xclamp=args.xclamp, precision=args.precision, device=dev, out=h_out, chunk=args.chunk,
                                 engine=eng)
            b.record()
            barrier()
            if i >= args.warmup:
                e2e_ms.append(a.elapsed_time(b))
        api = ('lagrangiancoherence_b200.rolling.rolling_ftle (pinned host winds in, pinned host fields out; every rank keeps '
               'its own start times -- the consumer is the host, nothing is gathered)')
        h2d, d2h = 2 * nlev * npts * u.itemsize, B * npts * 8
    e2e_total = float(np.sum(e2e_ms))
    if world > 1:
        t = torch.tensor([e2e_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_total = float(t.item())
    e2e_value = psteps_step * args.steps / (e2e_total * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: the departure-point integrator.
    # Bytes per particle-step: the reference formulation gathers (2+4S)(p+1)^2 values of w bytes (SURVEY 8(d):
    # 2304 B at S=4, p=3, f64).  The ES layout pre-combines the SETTLS operand 2f_k - f_{k+1} per grid point, so
    # the kernel gathers (1+S)(p+1)^2 two-value elements = 1280 B.  The roofline uses the bytes the kernel moves
    # through the bounding L1 data pipe and the measured ceiling for exactly that element size; the reference-
    # formulation figure is reported beside it, informational (it exceeds its own ceiling: the algorithmic saving).
    elt = 8 if args.precision == 'f64' else 4
    taps = (args.order + 1) ** 2
    survey_bytes_pstep = (2 + 4 * S_ORDER) * taps * elt
    issued_bytes_pstep = (1 + S_ORDER) * taps * 2 * elt
    if eng.pair_dtype == _lib.LCS_F64 and host_dtype == np.float32 and eng.f32_propagation:
        issued_bytes_pstep = (1 + 2 * S_ORDER) * taps * 2 * elt      # round32: levels k and k+1 sampled separately
    rows_rank = lat.size if not rowbands else (lambda r: r[3] - r[2])(rolling.shard_rows(lat.size, world, rank))
    psteps_launch = B * rows_rank * lon.size * (nt - 1)
    achieved = psteps_launch * issued_bytes_pstep / (advect_ms * 1e-3) / 1e9
    achieved_survey = psteps_launch * survey_bytes_pstep / (advect_ms * 1e-3) / 1e9
    # measured ceiling: lcs_gather_peak = same taps x taps pattern and thread tiling, coherent positions, integer
    # position arithmetic only, all rounds independent, on an L2-resident level.
    sink = torch.zeros(1, dtype=torch.float64, device=dev)
    iters = 40
    buf = torch.zeros((lat.size * lon.size + 8) * 4 * elt, dtype=torch.uint8, device=dev)
    Bp = max(B, 296)                                                    # enough windows to fill the machine

    def measure_peak(vec):
        gp = lambda: _lib.check(lib.lcs_gather_peak(_ptr(buf), eng.pair_dtype, vec, lat.size, lon.size, lat.size, lon.size,
                                                    Bp, args.order + 1, 0.0, iters, _ptr(sink), _stream(dev)), 'lcs_gather_peak')
        for _ in range(3):
            gp()
        torch.cuda.synchronize(dev)
        best = 1e30
        for _ in range(5):
            a, b = ev(), ev()
            a.record(); gp(); b.record()
            torch.cuda.synchronize(dev)
            best = min(best, a.elapsed_time(b))
        return Bp * npts * iters * taps * vec * elt / (best * 1e-3) / 1e9
    gather_peak_es = measure_peak(2)         # 2-value elements: what the ES kernel issues
    gather_peak = measure_peak(4)            # 4-value elements: the reference formulation's taps
    # ---- ONE field (north star: "a 48 h backward FTLE field on one B200 at >= 60 % of its gather roofline"): the
    # integrator of a single window of this workload, 30 launches back to back between two events so that the host's
    # launch path (25 us per call from Python) is not in the figure; both x-boundaries.  Outside the timed region.
    single_field = None
    if world == 1 and not rowbands:
        single_field = {'what': 'lcs_advect of ONE window of the workload, mean of 30 back-to-back launches (CUDA events)',
                        'bytes_per_particle_step': issued_bytes_pstep, 'peak': gather_peak_es, 'unit': 'GB/s'}
        st1 = eng.stage(d_u[:nt], d_v[:nt])
        for xm in ('pointwise', 'outer'):
            e1 = eng if xm == args.xclamp else FtleEngine(lat, lon, dt, SETTLS_order=S_ORDER, interp_order=args.order, xmode=xm,
                                                          device=dev, **precision_args(args.precision))
            for _ in range(5):
                e1.advect(st1, nsteps=nt - 1, nwindows=1, out=(x[:1], y[:1]))
            torch.cuda.synchronize(dev)
            a, b = ev(), ev()
            a.record()
            for _ in range(30):
                e1.advect(st1, nsteps=nt - 1, nwindows=1, out=(x[:1], y[:1]))
            b.record()
            torch.cuda.synchronize(dev)
            us = a.elapsed_time(b) / 30 * 1e3
            gbs = npts * (nt - 1) * issued_bytes_pstep / (us * 1e-6) / 1e9
            single_field[xm] = {'advect_us': us, 'achieved': gbs, 'frac': gbs / gather_peak_es}
        del st1
    hbm_peak, hbm_src = measured_peaks()
    # compulsory HBM traffic of the integrator: every staged level read once, final positions written once
    hbm_bytes = (nlev * 2 - 1) * npts * 2 * elt + B * 2 * rows_rank * lon.size * 8
    traffic = ncu_traffic(args, B)
    outer = args.xclamp == 'outer'
    kernel_name = ('advect_outer_group_kernel (one cooperative launch per step: resident CTA groups draw windows, group '
                   'barriers per sub-step)' if outer else 'advect_fused_kernel (one launch per step)')
    hbm = {'peak': hbm_peak, 'unit': 'GB/s', 'peak_source': hbm_src, 'compulsory_bytes': hbm_bytes,
           'compulsory_frac': hbm_bytes / (advect_ms * 1e-3) / 1e9 / hbm_peak}
    if traffic is not None:
        hbm.update({'measured_bytes': traffic, 'achieved': traffic / (advect_ms * 1e-3) / 1e9,
                    'frac': traffic / (advect_ms * 1e-3) / 1e9 / hbm_peak,
                    'source': 'dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full capture of this '
                              'command, profiles/r02_traffic.json) over the live CUDA-event time'})
    note = ('L1/L2 -> SM gather throughput (SURVEY 8(d)); not tensor: nothing on this path is a dense contraction.  HBM: '
            + ('the outer-clamp kernel also streams its position / Euler-sample state (48 B per particle sub-step) through DRAM -- '
               'see roofline.hbm.frac for the measured share of the copy bandwidth; shrinking the windows in flight until that '
               'state is L2-resident was measured and is slower (DESIGN.md 4), the kernel is bound by L1 wavefronts and issue '
               'slots' if outer else 'compulsory traffic only (levels read once, positions written once), see roofline.hbm'))
    line = {
        'metric': METRIC, 'value': value, 'unit': 'particle-steps/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'step_ms_rank0': [round(t, 3) for t in step_ms],
        'advect_ms_rank0': [round(sum(a.elapsed_time(b) for a, b in evs), 3) for evs in adv_ms],
        'cuda_mallocs_in_timed_region': mallocs_timed,
        'higher_is_better': True,
        'scaling': 'strong' if rowbands else 'weak',
        'vs_baseline': None, 'dtype': {'f64': 'f64', 'f32': 'f32 winds / f64 tap arithmetic and positions',
                                       'f32fast': 'f32 winds and tap arithmetic / f64 index map and positions'}[args.precision],
        'data': 'synthetic', 'config': config_dict(args, desc, B, world),
        'fields_per_s': work_ranks * B * args.steps / (total_ms * 1e-3),
        'interpolations_per_s': value * (2 + 4 * S_ORDER),
        'e2e': {'value': e2e_value, 'unit': 'particle-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_total / args.steps, 'fields_per_s': work_ranks * B * args.steps / (e2e_total * 1e-3),
                'api': api},
        'gpu_launches': launches_timed,
        'gpu_launches_per_step': launches_timed / args.steps,
        'clocks': clocks,
        'host_binding': numa,
        'gather_check': gather_check,
        'single_field': single_field,
        'roofline': {
            'bound': 'l1-gather',
            'bound_note': note,
            'kernel': kernel_name,
            'achieved': achieved, 'peak': gather_peak_es, 'unit': 'GB/s', 'frac': achieved / gather_peak_es,
            'bytes_per_particle_step': issued_bytes_pstep,
            'peak_source': 'lcs_gather_peak measured in this run: %dx%d-tap gathers of %d-B elements (what the ES kernel '
                           'issues) on an L2-resident level, same thread tiling, coherent positions, no dependent arithmetic'
                           % (args.order + 1, args.order + 1, 2 * elt),
            # informational: SURVEY 8(d)'s (2+4S)(p+1)^2 w bytes per particle-step over the same kernel time, against the
            # ceiling for 4-value taps; above 1 because the ES layout needs 2 values per tap, not 4
            'reference_formulation': {'bytes_per_particle_step': survey_bytes_pstep, 'achieved': achieved_survey,
                                      'peak': gather_peak, 'unit': 'GB/s', 'frac': achieved_survey / gather_peak,
                                      'note': 'informational only'},
            'advect_ms_per_step': advect_ms,
            'traffic': traffic,
            'hbm': hbm,
        },
    }
    if not args.no_cpu_baseline and world == 1:          # the CPU leg is timed at N = 1 only (rank 0 is the only rank there)
        cores = os.cpu_count() or 1
        lat2, lon2, nt2, _, desc2 = workload(args.workload)
        nwin = cores if args.workload == 'C3' else 4 * cores          # ~10 s of wall on the 16 cores of the GPU box
        secs = cpu_pool_run(args.workload, args.order, args.xclamp, cores, nwin)
        line['cpu_baseline'] = {'value': lat2.size * lon2.size * (cpu_levels(args.workload, nt2) - 1) * nwin / secs, 'unit': 'particle-steps/s',
                                'cores': cores, 'kind': 'port',
                                'sample': f'{nwin} windows of the same workload'
                                          + (' (first 2 intervals of each)' if args.workload == 'C3' else '') + f', over {cores} worker processes '
                                          f'({secs:.1f} s wall): oracle = scipy map_coordinates + numba-typed stencil '
                                          f'+ scipy.linalg.norm'}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
