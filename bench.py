#!/usr/bin/env python
"""
bench.py -- ground-plane hypotheses/s (detections x planes) of the polling hot path on N x B200.

    python bench.py --gpus N --steps K --warmup W              (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[3] = "C4": 4096 DISTINCT KITTI-size images x 100 synthetic
detections x road_planes_database_22k (21634 planes) PER GPU (weak scaling: images are sharded, the database is
replicated, there is no collective on the data path -- SURVEY.md section 8.5).  A step is one pass of the hot
path over that batch.

    value      whole-job hypotheses/s with the inputs already resident in HBM: sum over the K steps of the
               polling kernel's CUDA-event time on its launching stream, max over ranks; L2 is flushed
               between steps (the working set is smaller than L2)
    e2e        the same metric through the public numpy-in/numpy-out call ``fit_road_planes`` (C ABI
               ``gpp_fit_host``) from PINNED HOST buffers, host->device and device->host copies inside the
               timed region, wall clock between barriers;  e2e_pageable: the same from plain numpy arrays into
               fresh result arrays (the drop-in call as the reference's callers would make it)
    parity_*   the first 8 images of the LAST TIMED step against the C oracle (rank 0)
    roofline   FP32 CUDA-core bound (this path is neither HBM- nor tensor-bound: 0.006 B/hypothesis):
               achieved = 148 algorithmic FLOP/hypothesis x hypotheses / kernel time;
               peak = FFMA rate measured by libgpp's microbenchmark in this same run (MEASURED_PEAKS.json has no
               FP32 entry), nominal 74.4 TFLOP/s beside it; FMA-pipe / MUFU / issue-slot utilisation and DRAM traffic
               from the committed ncu capture of the same launch (profiles/)
    strong     N > 1: BASELINE.json's multi-GPU configuration itself -- the 4096 images of C4 SHARDED over the N ranks
               (kernel and end-to-end, time = max over ranks, efficiency against the one-GPU time of the same run)
    multi_gpu_single_process   N > 1: one caller driving all N GPUs (gpp_fit_host_multi), from
               pinned and from pageable host arrays, rank 0 while the other ranks wait on the host
    other_workloads   N = 1: C3, C2 and the reference's own call shape (one image, with and without padding rows):
               kernel time, numpy-call latency, roofline fraction
    cpu_baseline  the oracle port (numpy restatement of the reference graph) timed on this box's host cores
               on a bounded sample of the same workload

--impl reference times the reference's CPU implementation of the path: TensorFlow is not installable here,
so it is the oracle port (numpy restatement, one process per host core over images), on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W_ALG = 148.0                      # algorithmic FLOP per hypothesis, SURVEY.md section 8.4
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12
KERNEL_OF_MODE = {'verified': 'gpp::poll3_kernel<32, verified>',
                  'fast': 'gpp::poll3_kernel<32, fast>', 'exact': 'gpp::poll3_kernel<32, exact>'}
METRIC = 'ground-plane hypotheses/sec (dets x planes)'
UNIT = 'hypotheses/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='gpp', choices=['gpp', 'reference'])
    ap.add_argument('--images', type=int, default=4096, help='images per GPU (C4: 4096)')
    ap.add_argument('--dets', type=int, default=100)
    ap.add_argument('--planes', default='22k', choices=['10', '100', '1k', '10k', '22k'])
    ap.add_argument('--mode', default=os.environ.get('GPP_BENCH_MODE', 'verified'), choices=['exact', 'fast', 'verified'])
    ap.add_argument('--cpu-sample-images', type=int, default=0, help='images in the CPU-baseline sample (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def load_planes(tag):
    return np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % tag))


def make_workload(images, dets, planes, seed):
    """Seeded synthetic KITTI-shaped detections: `images` DISTINCT images, generated on the
    host in chunks (numpy; not part of any timed region)."""
    from gpp_b200.utils import synthetic
    parts = []
    for c0 in range(0, images, 512):
        parts.append(synthetic.synth_detections(min(512, images - c0), dets, planes, seed=seed * 1000 + c0 // 512))
    cat = lambda k: np.ascontiguousarray(np.concatenate([p[k] for p in parts], axis=0))  # noqa: E731
    return cat(0), cat(1), cat(2), cat(3).astype(np.float32)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        rows = [l for (t, l) in self.lines if t0 <= t <= t1 + 0.15] or [l for (_, l) in self.lines]
        sm, smax, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for l in rows:
            f = [x.strip() for x in l.split(',')]
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except Exception:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU legs
def _numpy_port_worker(job):
    from oracle.fit_road_planes_ref import fit_road_planes_ref
    boxes, dims, orient, P_inv, planes = job
    fit_road_planes_ref(boxes, dims, orient, P_inv, planes)
    return boxes.shape[0] * boxes.shape[1] * planes.shape[0]


def cpu_numpy_port(boxes, dims, orient, P_inv, planes, procs):
    """The oracle port (numpy restatement of fit_road_planes.py) over `procs` processes, one image per job."""
    B = boxes.shape[0]
    jobs = [(boxes[b:b + 1], dims[b:b + 1], orient[b:b + 1], P_inv[b:b + 1], planes) for b in range(B)]
    t0 = time.time()
    if procs <= 1:
        hyp = sum(_numpy_port_worker(j) for j in jobs)
    else:
        import multiprocessing as mp
        with mp.get_context('fork').Pool(procs) as pool:
            pool.map(_numpy_port_worker, jobs[:procs])          # untimed: worker start-up + imports
            t0 = time.time()
            hyp = sum(pool.map(_numpy_port_worker, jobs, chunksize=1))
    return hyp, time.time() - t0


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores."""
    if rank != 0:
        return
    planes = load_planes(args.planes)
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    n_img = args.cpu_sample_images or max(procs, 8)
    boxes, dims, orient, P_inv = make_workload(n_img, args.dets, planes, seed=3)
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    times, hyp = [], 0
    for i in range(args.warmup + args.steps):
        h, dt = cpu_numpy_port(boxes, dims, orient, P_inv, planes, procs)
        if i >= args.warmup:
            times.append(dt)
            hyp += h
    total = sum(times)
    value = hyp / total
    sample = '%d images x %d detections x %d planes per step (bounded sample of the 4096-image C4 batch)' % (
        n_img, args.dets, planes.shape[0])
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / max(1, args.steps),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, planes),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': procs, 'kind': 'port', 'sample': sample,
                         'what': 'numpy restatement of keras_retinanet_3D/layers/fit_road_planes.py (TensorFlow is '
                                 'not installable here), one process per host core over images',
                         'host_cpu_count': cores},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(args, planes):
    return {'workload': 'C4: %d distinct images x %d detections x road_planes_database_%s (%d planes) per GPU, KITTI P2 '
                        '1242x375 scaled 1333/1242, every row valid, key-point noise 1.5 px' % (
                            args.images, args.dets, args.planes, planes.shape[0]),
            'images_per_gpu': args.images, 'detections_per_image': args.dets, 'planes': int(planes.shape[0]),
            'mode': args.mode, 'sharding': 'images sharded, plane database replicated, no collective',
            'l2': 'flushed between timed steps (256 MiB write); working set < L2'}


def ncu_pipe_summary():
    """FMA-pipe / issue utilisation of the headline kernel from the newest committed ncu summary (profiles/): the
    148-FLOP roofline fraction counts algorithmic work, these say how busy the pipes really were."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_poll3_verified_c4_ncu.txt')))
    if not files:
        return None
    with open(files[-1]) as f:
        text = f.read()
    def grab(key):
        m = re.search(re.escape(key) + r'\s+([0-9.]+)', text)
        return float(m.group(1)) if m else None
    return {'source': 'profiles/' + os.path.basename(files[-1]) + ' (ncu --set full capture of this launch)',
            'fma_pipe_active_pct': grab('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
            'xu_pipe_pct': grab('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
            'issue_active_pct': grab('smsp__issue_active.avg.pct_of_peak_sustained_active'),
            'warps_active_pct': grab('sm__warps_active.avg.pct_of_peak_sustained_active'),
            'dram_bytes_per_launch': (lambda r, w: None if r is None or w is None else r * 1e6 + w * 1e3)(
                grab('dram__bytes_read.sum'), grab('dram__bytes_write.sum'))}


# ----------------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import gpp_b200

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the polling path has no CPU fallback '
                         '(use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL announces its version on stdout when the communicator is created; stdout is reserved for the ONE JSON
        # line, so file descriptor 1 points at stderr until the first collective is through
        import ctypes
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
            ctypes.CDLL(None).fflush(None)
            host_group = dist.new_group(backend='gloo')      # host-side waits that leave the GPUs idle (see below)
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    planes = load_planes(args.planes)
    N = int(planes.shape[0])
    B, D = args.images, args.dets
    # weak leg: every rank has its own batch (seed 3 + rank); the strong leg shards rank 0's batch (seed 3)
    boxes, dims, orient, P_inv = make_workload(B, D, planes, seed=3 + rank)
    hyp_per_step = float(B) * D * N
    poller = gpp_b200.get_poller(local_rank)
    poller.set_planes(planes)

    # ---------------- FP32 peak (roofline denominator), measured in this run
    ffma = poller.microbench(0)
    peak_tflops = 2.0 * ffma['ops_per_s'] / 1e12

    # ---------------- device-resident leg: `value`
    tb, td = torch.from_numpy(boxes).to(dev), torch.from_numpy(dims).to(dev)
    to, tp = torch.from_numpy(orient).to(dev), torch.from_numpy(P_inv).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def device_step(t=None):
        flush.fill_(1)                                       # evict L2 (untimed: kernel time comes from events)
        return poller.fit_torch(*(t or (tb, td, to, tp)), mode=args.mode, return_index=True)

    # nvidia-smi's start-up takes a driver lock that stalls a running kernel for tens of ms: start the sampler
    # BEFORE the warm-up steps so that only its steady 100 ms polling overlaps the timed region
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.5)
    for _ in range(args.warmup):
        device_step()
    barrier()
    launches0 = poller.launch_count()
    t_clock0 = time.time()
    kernel_ms = []
    barrier()
    w0 = time.time()
    for _ in range(args.steps):
        timed_out = device_step()
        torch.cuda.synchronize()
        kernel_ms.append(poller.last_kernel_ms())            # CUDA events around the kernel, launching stream
    barrier()
    wall_dev = time.time() - w0
    t_clock1 = time.time()
    launches = poller.launch_count() - launches0
    dev_ms = max_over_ranks(sum(kernel_ms))
    value = world * hyp_per_step * args.steps / (dev_ms * 1e-3)

    # ---------------- parity of the run that was just timed: the first 8 images against the C oracle (rank 0)
    parity = None
    if rank == 0:
        from oracle import c_oracle
        n_chk = min(8, B)
        want = c_oracle.fit_road_planes_c(boxes[:n_chk], dims[:n_chk], orient[:n_chk], P_inv[:n_chk], planes, return_index=True)
        got = [t[:n_chk].cpu().numpy() for t in timed_out]
        bad = int(np.sum(got[3] != want[3]))
        bits = all(np.array_equal(g, w, equal_nan=True) for g, w in zip(got, want))
        parity = {'parity_checked': True, 'parity_mismatches': bad, 'outputs_bit_identical': bool(bits),
                  'parity_rows': int(n_chk * D), 'parity_against': 'oracle/gpp_oracle.c on the last timed step'}

    # ---------------- end-to-end leg: public numpy API, host buffers, copies inside the timed region
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    pin_in = [pinned(a) for a in (boxes, dims, orient, P_inv)]
    pin_out = [torch.empty((B, D, 4, 3), dtype=torch.float32, pin_memory=True),
               torch.empty((B, D, 1, 4), dtype=torch.float32, pin_memory=True),
               torch.empty((B, D), dtype=torch.float32, pin_memory=True)]
    outs = [t.numpy() for t in pin_out]
    nb, nd, no, npi = [t.numpy() for t in pin_in]

    def time_e2e(step, steps):
        for _ in range(max(1, args.warmup)):
            step()
        barrier()
        e0 = time.time()
        for _ in range(steps):
            step()
        barrier()
        return max_over_ranks(time.time() - e0)

    def e2e_step():
        gpp_b200.fit_road_planes(nb, nd, no, npi, planes, mode=args.mode, device=local_rank, out=outs)
        return float(outs[2][0, 0])                          # the step's result is read on the host

    def e2e_pageable_step():                                 # the drop-in call as the reference's callers make it
        return float(gpp_b200.fit_road_planes(boxes, dims, orient, P_inv, planes, mode=args.mode, device=local_rank)[2][0, 0])

    e2e_s = time_e2e(e2e_step, args.steps)
    e2e_value = world * hyp_per_step * args.steps / e2e_s
    pg_steps = max(3, args.steps // 2)
    e2e_pg_s = time_e2e(e2e_pageable_step, pg_steps)
    h2d = int(nb.nbytes + nd.nbytes + no.nbytes + npi.nbytes)
    d2h = int(sum(o.nbytes for o in outs))
    clocks = sampler.stop(t_clock0, max(t_clock1, time.time()))

    # ---------------- strong scaling (BASELINE.json configs[3]): rank 0's 4096 images sharded over the N ranks
    strong = None
    if world > 1:
        from gpp_b200.sharding import shard_bounds
        full = make_workload(B, D, planes, seed=3)            # the same batch on every rank; rank r polls its shard
        b0, b1 = shard_bounds(B, world, rank)
        ts = [torch.from_numpy(a[b0:b1]).to(dev) for a in full]
        for _ in range(args.warmup):
            device_step(ts)
        barrier()
        sk = []
        for _ in range(args.steps):
            device_step(ts)
            torch.cuda.synchronize()
            sk.append(poller.last_kernel_ms())
        strong_ms = max_over_ranks(sum(sk)) / args.steps
        sh_in = [pinned(a[b0:b1]) for a in full]
        sh_np = [t.numpy() for t in sh_in]
        sh_out = [o[:b1 - b0] for o in outs]

        def strong_e2e_step():
            gpp_b200.fit_road_planes(sh_np[0], sh_np[1], sh_np[2], sh_np[3], planes, mode=args.mode, device=local_rank, out=sh_out)
            return float(sh_out[2][0, 0])
        strong_e2e_s = time_e2e(strong_e2e_step, args.steps) / args.steps
        t1_kernel = dev_ms / args.steps                      # one GPU polling all 4096 images (the weak leg of this run)
        t1_e2e = e2e_s / args.steps
        strong = {'images_total': B, 'images_per_gpu': [shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world)],
                  'kernel_ms': strong_ms, 'hyp_per_s_kernel': hyp_per_step / (strong_ms * 1e-3),
                  'e2e_ms': 1e3 * strong_e2e_s, 'hyp_per_s_e2e': hyp_per_step / strong_e2e_s,
                  'one_gpu_kernel_ms': t1_kernel, 'one_gpu_e2e_ms': 1e3 * t1_e2e,
                  'efficiency_kernel': t1_kernel / (world * strong_ms), 'efficiency_e2e': t1_e2e / (world * strong_e2e_s),
                  'what': 'C4 = %d images sharded over %d GPUs, one process per GPU, time = max over ranks; efficiency = '
                          'T(1 GPU, same run) / (N x T(N GPUs))' % (B, world)}

    # ---------------- the single-process multi-GPU driver (numpy caller on a multi-GPU box), rank 0 while the others wait
    # (the other ranks wait in a gloo barrier: an NCCL barrier would keep a spinning kernel of ANOTHER process on
    # their GPUs, and kernels of two processes time-slice a GPU instead of sharing it)
    multi = None
    n_vis = torch.cuda.device_count()
    if world > 1:
        barrier()
        dist.barrier(group=host_group)
    if rank == 0 and world > 1 and n_vis >= world:
        devs = list(range(world))                             # exactly the GPUs this job was given
        for d in devs:
            gpp_b200.get_poller(d).set_planes(planes)
        multi = {'devices': len(devs)}
        for label, ins, out_arrays in (('pinned', (nb, nd, no, npi), outs), ('pageable', (boxes, dims, orient, P_inv), None)):
            ms = []
            for i in range(4):
                c0 = time.time()
                gpp_b200.fit_road_planes_multi(ins[0], ins[1], ins[2], ins[3], planes, devices=devs, mode=args.mode, out=out_arrays)
                if i:
                    ms.append(1e3 * (time.time() - c0))
            multi[label + '_ms'] = float(np.mean(ms))
            multi[label + '_hyp_per_s'] = hyp_per_step / (np.mean(ms) * 1e-3)
            multi[label + '_efficiency_vs_one_gpu_e2e'] = (1e3 * (e2e_s if label == 'pinned' else e2e_pg_s) /
                                                           (args.steps if label == 'pinned' else pg_steps)) / (len(devs) * np.mean(ms))
        multi['what'] = 'fit_road_planes_multi: ONE process, one host thread + one handle per GPU, C4 (%d images) sharded' % B
    if world > 1:
        dist.barrier(group=host_group)
        barrier()

    # ---------------- the other arithmetic modes, for context (kernel only, 3 steps each)
    other_modes = {}
    for other in ('verified', 'exact', 'fast'):
        if other == args.mode:
            continue
        oms = []
        for i in range(4):
            poller.fit_torch(tb, td, to, tp, mode=other)
            torch.cuda.synchronize()
            if i:
                oms.append(poller.last_kernel_ms())
        other_modes[other] = world * hyp_per_step / (max_over_ranks(float(np.mean(oms))) * 1e-3)

    # ---------------- the other configurations of BASELINE.json and the reference's own call shape (N = 1 only)
    other_workloads = None
    if rank == 0 and world == 1:
        other_workloads = {}
        for tag, (img, db, nv) in (('C3_64x100x10k', (64, '10k', 100)), ('C2_1x100x1k', (1, '1k', 100)),
                                   ('single_image_1x100x22k', (1, '22k', 100)),
                                   ('single_image_15_valid_rows_85_padding', (1, '22k', 15))):
            from gpp_b200.utils import synthetic
            pl = load_planes(db)
            wb, wd, wo, wp = synthetic.synth_detections(img, args.dets, pl, seed=11, n_valid=nv)
            wp = wp.astype(np.float32)
            poller.set_planes(pl)
            tw = [torch.from_numpy(a).to(dev) for a in (wb, wd, wo, wp)]
            kms, calls = [], []
            for i in range(30):                      # sub-millisecond kernels after a database switch: 10 warm-up calls
                poller.fit_torch(*tw, mode=args.mode)
                torch.cuda.synchronize()
                if i >= 10:
                    kms.append(poller.last_kernel_ms())
            feed = np.expand_dims(np.asfortranarray(pl.astype(np.float64)), axis=0)     # the callers' (1, N, 4) float64 feed
            for i in range(30):
                c0 = time.perf_counter()
                gpp_b200.fit_road_planes(wb, wd, wo, wp, feed, mode=args.mode, device=local_rank)
                if i >= 5:
                    calls.append(time.perf_counter() - c0)
            hyp = float(img) * args.dets * pl.shape[0]
            k = float(np.median(kms))
            other_workloads[tag] = {'kernel_ms': k, 'hyp_per_s_kernel': hyp / (k * 1e-3),
                                    'numpy_call_ms': 1e3 * float(np.median(calls)),
                                    'roofline_frac': W_ALG * hyp / (k * 1e-3) / 1e12 / peak_tflops}
        poller.set_planes(planes)

    # ---------------- CPU baseline beside it (rank 0, N = 1 only, bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle
        n_img = args.cpu_sample_images or 8
        h, dt = cpu_numpy_port(boxes[:n_img], dims[:n_img], orient[:n_img], P_inv[:n_img], planes, 1)
        c_threads = c_oracle.max_threads()
        c_img = min(args.images, max(16, 2 * c_threads))
        t0 = time.time()
        c_oracle.fit_road_planes_c(boxes[:c_img], dims[:c_img], orient[:c_img], P_inv[:c_img], planes)
        c_dt = time.time() - t0
        cpu = {'value': h / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
               'sample': '%d images x %d detections x %d planes (%.1f s)' % (n_img, D, N, dt),
               'what': 'numpy restatement of the reference graph, single process',
               'c_port_value': c_img * D * N / c_dt, 'c_port_cores': c_threads,
               'c_port_sample': '%d images (%.1f s), fused C restatement, all host threads' % (c_img, c_dt),
               'host_cpu_count': os.cpu_count()}

    if rank == 0:
        pipes = ncu_pipe_summary() if (args.mode == 'verified' and args.images == 4096 and args.planes == '22k') else None
        ms_per_step = dev_ms / args.steps
        achieved = W_ALG * value / world / 1e12              # per-GPU TFLOP/s of algorithmic work
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, planes),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': 1e3 * e2e_s / args.steps,
                    'api': 'gpp_b200.fit_road_planes (numpy in/out, pinned host buffers) -> gpp_fit_host'},
            'e2e_pageable': {'value': world * hyp_per_step * pg_steps / e2e_pg_s, 'unit': UNIT,
                             'ms_per_step': 1e3 * e2e_pg_s / pg_steps, 'steps': pg_steps,
                             'api': 'gpp_b200.fit_road_planes on plain (pageable) numpy arrays, results in fresh arrays: '
                                    'the drop-in call'},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {'bound': 'fp32', 'achieved': achieved, 'peak': peak_tflops, 'unit': 'TFLOP/s',
                         'frac': achieved / peak_tflops,
                         'traffic': pipes['dram_bytes_per_launch'] if pipes else None, 'traffic_unit': 'bytes/launch',
                         'traffic_source': pipes['source'] if pipes else None,
                         'peak_source': 'libgpp FFMA microbenchmark, same run (MEASURED_PEAKS.json has no FP32 entry)',
                         'nominal_peak': NOMINAL_FP32_TFLOPS, 'frac_of_nominal': achieved / NOMINAL_FP32_TFLOPS,
                         'flop_per_hypothesis': W_ALG, 'kernel': KERNEL_OF_MODE[args.mode],
                         'kernel_ms_per_launch': ms_per_step,
                         'fma_pipe_active_pct': pipes['fma_pipe_active_pct'] if pipes else None,
                         'xu_pipe_pct': pipes['xu_pipe_pct'] if pipes else None,
                         'issue_active_pct': pipes['issue_active_pct'] if pipes else None,
                         'pipe_source': pipes['source'] if pipes else None,
                         'frac_note': 'achieved counts the ALGORITHMIC 148 FLOP of every hypothesis (SURVEY.md 8.4); the kernel '
                                      'drops most rows of 64 planes after less than half a hypothesis, so frac can exceed 1 -- '
                                      'how busy the pipes are is fma_pipe_active_pct / xu_pipe_pct / issue_active_pct'},
            'cpu_baseline': cpu,
            'other_modes': dict(other_modes, unit=UNIT),
            'other_workloads': other_workloads,
            'strong': strong,
            'multi_gpu_single_process': multi,
            'wall_ms_per_step_device_leg': 1e3 * wall_dev / args.steps,
            'kernel_ms_steps': [round(x, 3) for x in kernel_ms],
        }
        line.update(parity or {})
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
