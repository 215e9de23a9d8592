#!/usr/bin/env python
"""Benchmark of the LRGNet grow engine on synthetic S3DIS-shaped rooms (BASELINE.json metric: segmented points/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rooms R] [--impl reference]

One "step" = one full pass of the hot path over the workload: every room of the synthetic Area-5-shaped set
(68 rooms, ~20k raw points each, seeds 1000+room): feature preparation on the device (the reference's points/s timer starts
before it, test_region_grow.py:120,317), every room grown to completion, filled.  `value` is measured with the raw points
already resident in HBM (CUDA events on the engine stream, max over ranks); `e2e` is the same pass through the public
host-buffer calls (raw points in pinned host memory in, per-raw-point labels out, copies inside the timed region);
`e2e_features` is the pass on 13-D features prepared beforehand (feature-level API, host buffers in, labels out).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

FLOPS_PER_STEP = 271712256                     # BASELINE.md section 4 (algorithmic, factored heads)
BRANCH_FLOPS_PER_STEP = 2 * 2 * 512 * 82752    # both branches, 512 points, 82,752 MAC per point
HEAD_FLOPS_PER_STEP = 2 * 2 * 512 * (64 * 256 + 256 * 128 + 128 * 2)
GPROJ_FLOPS_PER_STEP = 2 * 2 * 1024 * 256
BRANCH_TILE_FLOPS = 2 * 128 * 82752                      # one 128-row tile of one branch
HEAD_TILE_FLOPS = 2 * 128 * (64 * 256 + 256 * 128)       # tensor part of one 128-row head tile (128->2 runs on the FMA pipe)
METRIC = 'segmented_points_per_sec'
UNIT = 'points/s'


def load_peaks():
    try:
        return json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'sm_max_mhz': 1965.0}, 'fallback'


def make_workload(n_rooms, seed_base, cache=True):
    """Synthetic raw rooms (x y z r g b obj_id cls_id rows): (raw_offsets (R+1) int64, raw_points (sum Nr, 8) float32)."""
    from tools import rooms
    path = '/tmp/lrg_bench_rooms_v3_%d_%d.npz' % (n_rooms, seed_base)
    if cache and os.path.exists(path):
        z = np.load(path)
        return z['raw_offsets'], z['raw_points']
    raw_rows = [rooms.generate_room(seed_base + r) for r in range(n_rooms)]
    raw_off = np.zeros(n_rooms + 1, np.int64)
    np.cumsum([len(r) for r in raw_rows], out=raw_off[1:])
    raw_points = np.ascontiguousarray(np.concatenate(raw_rows), np.float32)
    if cache:
        try:
            np.savez(path, raw_offsets=raw_off, raw_points=raw_points)
        except Exception:
            pass
    return raw_off, raw_points


def host_features(raw_off, raw_points, room):
    """13-D features + seed order of one room by the oracle's host restatement of test_region_grow.py:119-173 -- only the CPU
    arms (cpu_baseline, --impl reference) call this; the GPU arms prepare the features on the device."""
    from oracle import feature_prep
    f = feature_prep.prepare_features(raw_points[raw_off[room]:raw_off[room + 1]], 0.1)
    return f['points'], f['order'].astype(np.int32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def pinned_array(lib_mod, shape, dtype):
    import ctypes as C
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    lib_mod.check(lib_mod.lib().lrg_host_alloc(C.byref(p), max(n, 1)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


# ----------------------------------------------------------------------------------------------- CPU arms
def cpu_sample(weights, points, order, n_steps, literal, room_id=0):
    """Time `n_steps` grow steps of the oracle driver (port of the reference) on one room.  TEST INFRA used as baseline."""
    from oracle import lrg_driver, lrg_forward
    fwd = lambda a, b: lrg_forward.forward(weights, a, b)
    g = lrg_driver.RoomGrower(points, order, fwd, lrg_driver.PhiloxRng(0), room_id=room_id, literal_update=literal)
    t0 = time.perf_counter()
    for seed_id in np.arange(len(points))[order]:
        if g.visited[seed_id]:
            continue
        g.begin_region(seed_id)
        while g.total_steps < n_steps:
            st = g.prepare_step()
            if st is None:
                break
            add, rmv = fwd(st['inlier'], st['neighbor'])
            if g.apply_step(add[0], rmv[0]) is not None:
                break
        if g.total_steps >= n_steps:
            break
    return g.total_steps, time.perf_counter() - t0


def load_weights():
    with np.load(os.path.join(REPO, 'tests', 'golden', 'lrgnet_model5.npz')) as z:
        return {k: z[k] for k in z.files}


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path (port: oracle/lrg_driver.py with the literal
    per-point update loop of test_region_grow.py:282-287 + numpy forward on all host threads); the Python reference cannot
    travel to the GPU box (TensorFlow/h5py absent), so this is the oracle port."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    raw_off, raw_points = make_workload(args.rooms, 1000)
    raw_counts = np.diff(raw_off)
    weights = load_weights()
    feats = {}
    steps_per_room = None
    sample_steps = args.ref_sample_steps
    times, nsteps = [], []
    for it in range(args.warmup + args.steps):
        room = it % args.rooms
        if room not in feats:
            feats[room] = host_features(raw_off, raw_points, room)      # outside the timed sample (the GPU arms time it)
        p, o = feats[room]
        n, dt = cpu_sample(weights, p, o, sample_steps if it >= args.warmup else max(5, sample_steps // 10), literal=True, room_id=room)
        if it >= args.warmup:
            times.append(dt)
            nsteps.append(n)
    steps_per_s = sum(nsteps) / sum(times)
    # points/s = steps/s x (raw points per grow step of this workload); the ratio comes from the workload statistics file
    # written by the GPU arm when available, else from the reference logs (~948 steps per ~20k-point room)
    pts_per_step = workload_points_per_step(args.rooms, raw_counts)
    value = steps_per_s * pts_per_step
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * float(np.mean(times)), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'area5_synthetic_%d_rooms_20k_raw' % args.rooms, 'rooms': args.rooms, 'resolution': 0.1,
                   'scope': 'grow driver + LrgNet forward on precomputed 13-D features'},
        'grow_steps_per_sec': steps_per_s,
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
                         'sample': '%d grow steps per timed step on one room (literal per-point update loop, numpy forward); points/s = steps/s x %.2f raw points per grow step' % (sample_steps, pts_per_step)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def workload_points_per_step(n_rooms, raw_counts):
    """Raw points per grow step of the workload: from the GPU arm's last run on this box if there was one, else from the
    committed statistics of the same deterministic workload (profiles/workload_stats.json), else from the reference logs."""
    try:
        s = json.load(open('/tmp/lrg_bench_stats_%d.json' % n_rooms))
        return float(s['raw_points']) / float(s['grow_steps'])
    except Exception:
        pass
    try:
        s = json.load(open(os.path.join(REPO, 'profiles', 'workload_stats.json')))['area5_synthetic_%d_rooms_20k_raw' % n_rooms]
        return float(s['raw_points']) / float(s['grow_steps'])
    except Exception:
        return float(np.mean(raw_counts)) / 948.0


# ----------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--rooms', type=int, default=68, help='rooms per GPU (Area 5 has 68)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--ref-sample-steps', type=int, default=120)
    ap.add_argument('--cpu-baseline-steps', type=int, default=150)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the statistics / random-restart / beam-search measurements')
    ap.add_argument('--slots', type=int, default=0)
    ap.add_argument('--lockstep-timing', action='store_true', help='also time the lock-step loop kernel by kernel')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        return run_reference_arm(args)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local_rank))

    from learn_region_grow_b200 import _lib, parallel
    from learn_region_grow_b200.engine import Engine
    peaks, peaks_src = load_peaks()
    weights = load_weights()
    # weak scaling: every rank grows its own Area-5-sized set of rooms (different seeds)
    raw_off, raw_points = make_workload(args.rooms, 1000 + rank * args.rooms)
    raw_counts = np.diff(raw_off)
    total_raw = int(raw_counts.sum())
    eng = Engine(1, 1, 512, 512, 13, 0, device=local_rank)
    eng.load_weights(weights)
    params = dict(resolution=0.1, seed=0, max_slots=args.slots, room_id_base=rank * args.rooms)

    # raw points resident in HBM; one preparation up front sizes the equalised rooms (and yields the features of `e2e_features`)
    d_raw = torch.from_numpy(np.ascontiguousarray(raw_points, np.float32)).cuda()
    offsets = eng.upload_raw_concatenated(raw_off, d_raw, 0.1)
    feat = eng.prepared_features()
    points, order = feat['points'], feat['order']
    # pinned host buffers for the secondary end-to-end arm
    h_points = pinned_array(_lib, points.shape, np.float32); h_points[...] = points
    h_order = pinned_array(_lib, order.shape, np.int32); h_order[...] = order
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')      # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lengths = [int(offsets[-1])] * world      # every rank has the same room sizes only by seed; gather actual below
    if dist is not None:
        t = torch.tensor([int(offsets[-1])], device='cuda')
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        lengths = [int(x.item()) for x in allt]

    def gather_labels():
        if dist is None:
            return
        local = torch.as_tensor(parallel.DeviceArray(eng.labels_device_ptr(True), int(offsets[-1])), device='cuda')
        parallel.allgather_labels(local, lengths)

    # ---- resident arm: `value` (raw points resident in HBM -> device feature preparation -> grow -> fill)
    launches = 0
    prep_ms_list = []
    stats = None
    dev_ms = []
    grow_ms_list = []
    for it in range(args.warmup + args.steps):
        flush.fill_(it & 0xFF)
        if it == args.warmup:
            barrier()
            sampler = ClockSampler(local_rank)
            sampler.start()
            wall0 = time.perf_counter()
        eng.upload_raw_concatenated(raw_off, d_raw, 0.1)       # device feature preparation from the resident raw points
        prep_ms = eng.prepare_ms()
        stats = eng.segment_resident(**params)
        t_ag0 = time.perf_counter()
        gather_labels()
        torch.cuda.synchronize()
        t_ag = time.perf_counter() - t_ag0
        if it >= args.warmup:
            pr = eng.profile()
            dev_ms.append(prep_ms + pr['grow_ms'] + pr['fill_ms'] + (1e3 * t_ag if dist is not None else 0.0))
            grow_ms_list.append(pr['grow_ms'])
            prep_ms_list.append(prep_ms)
            launches += pr['kernel_launches'] + 6      # + 4 feature-preparation kernels, pack, flag reset
    barrier()
    wall = time.perf_counter() - wall0
    total_ms = float(sum(dev_ms))
    if dist is not None:
        t = torch.tensor([total_ms, wall], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, wall = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    grow_steps = int(stats['grow_steps'].sum())
    value = world * total_raw / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel: the persistent grow kernel (one launch per pass; CUDA events on the engine stream)
    pr = eng.profile()
    peak_tf = peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops'))
    grow_ms = float(np.mean(grow_ms_list))
    achieved_tf = grow_steps * FLOPS_PER_STEP / (grow_ms * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(REPO, 'profiles', 'grow_kernel_traffic.json')))['dram_bytes_per_launch']
    except Exception:
        pass
    roofline = {
        'kernel': 'lrg_grow_kernel' if pr['persistent'] else 'lock-step kernels', 'bound': 'tensor', 'achieved': achieved_tf, 'peak': peak_tf,
        'unit': 'TFLOP/s', 'frac': achieved_tf / peak_tf if peak_tf else None, 'traffic': traffic, 'peak_source': peaks_src + ' bf16 sustained',
        'launches_per_pass': 1 if pr['persistent'] else None, 'avg_launch_ms': grow_ms,
        'algorithmic_flops_per_launch': grow_steps * FLOPS_PER_STEP,
        'note': 'algorithmic = 271.7 MFLOP per grow step (512+512 rows, factored heads); the kernel evaluates only distinct rows '
                '(padding duplicates reuse logits) as 3xTF32 on tcgen05; the run is latency-bound by the longest room; '
                'grid = one CTA per SM: 132 work-item CTAs + 16 pooled-projection server CTAs (weights resident in shared memory), '
                'unless LRG_FLAG_NO_PROJ_SERVERS',
    }
    # the driver phases' side of the roofline: algorithmic bytes per grow step = one pass over the room's state
    # (14 B per point) + the gathered tiles and logits (62,464 B) + the fp32 weights amortised over the rooms stepped together;
    # (the 36 B per inlier of the median are left out: the mean inlier count is not tracked)
    n_eq_mean = float(offsets[-1]) / max(args.rooms, 1)
    bytes_step = 14.0 * n_eq_mean + 62464.0 + 3164176.0 / max(args.rooms, 1)
    hbm_peak = peaks.get('hbm_gbs')
    roofline['hbm_side'] = {'algorithmic_bytes_per_grow_step': bytes_step, 'achieved_gbs': grow_steps * bytes_step / (grow_ms * 1e-3) / 1e9,
                            'peak_gbs': hbm_peak, 'frac': (grow_steps * bytes_step / (grow_ms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None,
                            'note': 'neither roofline binds: the pass is the longest room\'s chain of sequential steps (DESIGN.md section 6)'}
    if pr['persistent']:
        items, busy = pr['items'], pr['busy_ms']
        executed = 3.0 * (items['branch'] * BRANCH_TILE_FLOPS + items['head'] * HEAD_TILE_FLOPS)
        roofline.update({
            'items_per_pass': items, 'busy_ms_by_item': busy,
            'avg_us_per_item': {k: (1e3 * busy[k] / items[k] if items[k] else None) for k in items},
            'sm_busy_frac': sum(busy.values()) / (148.0 * grow_ms),
            'executed_tf32_tflops': executed / (grow_ms * 1e-3) / 1e12,
            'frac_of_tf32_peak_executed': executed / (grow_ms * 1e-3) / 1e12 / (peak_tf / 2.0) if peak_tf else None,
            'tensor_tile_busy_tflops': executed / ((busy['branch'] + busy['head']) * 1e-3 / 148.0) / 1e12 if (busy['branch'] + busy['head']) > 0 else None,
        })
        try:
            # the reference's own timing buckets (comp_time_analysis, test_region_grow.py:40-51,120-317), per pass: 'feature' =
            # feature preparation (device time); 'net' = the forward (branch + projection + head handlers), 'neighbor' + 'inlier'
            # = the driver step handler -- SM time summed over CTAs, since the phases of different rooms overlap on the device
            net_ms = float(busy['branch'] + busy['gproj'] + busy['head'])
            drv_ms = float(busy['step'])
            roofline['reference_buckets'] = {
                'feature_ms_device': float(np.mean(prep_ms_list)), 'net_sm_ms': net_ms, 'neighbor_plus_inlier_sm_ms': drv_ms,
                'net_share_of_sm_time': net_ms / (net_ms + drv_ms) if (net_ms + drv_ms) > 0 else None,
                'note': 'the pooled projection runs on 16 server CTAs and is not in net_sm_ms unless LRG_FLAG_NO_PROJ_SERVERS'}
        except Exception:
            pass
    if args.lockstep_timing:
        # per-kernel CUDA-event times of the lock-step loop (the A/B path; one {step, branch, gproj, head} quartet per iteration)
        eng.segment_resident(flags=_lib.FLAG_KERNEL_TIMING, **params)
        kt = eng.profile()
        roofline['lockstep_kernel_ms'] = {k: kt[k] for k in ('step_kernel_ms', 'branch_kernel_ms', 'gproj_kernel_ms', 'head_kernel_ms')}
        roofline['lockstep_iterations'] = kt['iterations']

    # ---- secondary end-to-end arm: 13-D features prepared on the host beforehand, host buffers in, labels out
    h2d = h_points.nbytes + h_order.nbytes + offsets.nbytes
    d2h = int(offsets[-1]) * 4 + args.rooms * 32
    for it in range(2):
        eng.segment_concatenated(offsets, h_points, h_order, **params)
    barrier()
    e0 = time.perf_counter()
    e2e_ms = []
    for it in range(args.steps):
        t_it = time.perf_counter()
        labels, _ = eng.segment_concatenated(offsets, h_points, h_order, **params)
        gather_labels()
        e2e_ms.append(1e3 * (time.perf_counter() - t_it))
    barrier()
    e2e_s = time.perf_counter() - e0
    # ---- end to end (`e2e`): RAW points (x y z r g b ...) in pinned host memory -> device feature preparation -> growing ->
    # fill -> per-raw-point labels back on the host; copies inside the timed region
    h_raw = pinned_array(_lib, raw_points.shape, np.float32); h_raw[...] = raw_points
    raw_ms = []
    for it in range(1 + args.steps):
        barrier()
        t_it = time.perf_counter()
        eng.upload_raw_concatenated(raw_off, h_raw, 0.1)
        st_raw = eng.segment_resident(**params)
        lab_raw = eng.raw_labels(True)
        gather_labels()
        barrier()
        if it >= 1:
            raw_ms.append(1e3 * (time.perf_counter() - t_it))
    raw_s = float(np.mean(raw_ms)) * 1e-3
    if dist is not None:
        t = torch.tensor([raw_s], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        raw_s = float(t[0])
    e2e_raw = {'value': world * total_raw / raw_s, 'unit': UNIT, 'h2d_bytes_per_step': int(h_raw.nbytes + raw_off.nbytes),
               'd2h_bytes_per_step': int(total_raw * 4 + args.rooms * 32), 'ms_per_step': [round(x, 2) for x in raw_ms],
               'grow_steps_per_pass': int(st_raw['grow_steps'].sum()),
               'scope': 'raw points (x y z r g b) in pinned host memory -> device feature preparation (test_region_grow.py:119-173) -> grow -> fill -> per-raw-point labels on the host'}
    # ---- the rows either side of the path, once each on the raw rooms just uploaded (not part of `value` / `e2e`):
    # segmentation statistics (test_region_grow.py:319-349), the random-restart driver (test_random_restart.py) and the
    # beam-search driver (test_beam_search.py)
    extras = {}
    if not args.no_extras:
        obj_raw = [raw_points[raw_off[i]:raw_off[i + 1], 6].astype(np.int32) for i in range(args.rooms)]
        eng.room_metrics(obj_raw, raw=True)
        t0 = time.perf_counter()
        mt = eng.room_metrics(obj_raw, raw=True)
        extras['statistics'] = {'ms_per_call': 1e3 * (time.perf_counter() - t0), 'rooms': args.rooms,
                                'mean': {k: float(np.nanmean(mt[k])) for k in ('nmi', 'ami', 'ars', 'prc', 'rcl', 'iou')},
                                'scope': 'obj_id of the raw points in, NMI/AMI/ARS/PRC/RCL/IOU per room out (contingency tables and expected mutual information on the device)'}
        eng.segment_resident(num_restarts=10, **params)
        st_rr = eng.segment_resident(num_restarts=10, **params)
        pr_rr = eng.profile()
        mt_rr = eng.room_metrics(obj_raw, raw=True)
        extras['random_restart'] = {'num_restarts': 10, 'grow_ms_per_pass': pr_rr['grow_ms'], 'grow_steps_per_pass': int(st_rr['grow_steps'].sum()),
                                    'grow_steps_per_sec': float(st_rr['grow_steps'].sum()) / (pr_rr['grow_ms'] * 1e-3),
                                    'points_per_sec': total_raw / ((pr_rr['grow_ms'] + pr_rr['fill_ms']) * 1e-3), 'per_gpu': True,
                                    'mean': {k: float(np.nanmean(mt_rr[k])) for k in ('nmi', 'ami', 'ars', 'prc', 'rcl', 'iou')},
                                    'scope': 'test_random_restart.py: 10 restarts per seed as parallel lanes, largest region kept'}
        eng.segment_resident(beam_width=3, search_width=3, **params)
        st_bs = eng.segment_resident(beam_width=3, search_width=3, **params)
        pr_bs = eng.profile()
        mt_bs = eng.room_metrics(obj_raw, raw=True)
        extras['beam_search'] = {'beam_width': 3, 'search_width': 3, 'grow_ms_per_pass': pr_bs['grow_ms'], 'grow_steps_per_pass': int(st_bs['grow_steps'].sum()),
                                 'grow_steps_per_sec': float(st_bs['grow_steps'].sum()) / (pr_bs['grow_ms'] * 1e-3),
                                 'points_per_sec': total_raw / ((pr_bs['grow_ms'] + pr_bs['fill_ms']) * 1e-3), 'per_gpu': True,
                                 'mean': {k: float(np.nanmean(mt_bs[k])) for k in ('nmi', 'ami', 'ars', 'prc', 'rcl', 'iou')},
                                 'scope': 'test_beam_search.py: 3 candidates x 3 expansions per round as parallel lanes, largest masks kept'}
    clocks = sampler.stop()          # nvidia-smi keeps sampling through the timed regions
    if dist is not None:
        t = torch.tensor([e2e_s], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = world * total_raw * args.steps / e2e_s

    if rank == 0:
        try:
            json.dump({'raw_points': total_raw, 'grow_steps': grow_steps}, open('/tmp/lrg_bench_stats_%d.json' % args.rooms, 'w'))
        except Exception:
            pass
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        p0, o0 = host_features(raw_off, raw_points, 0)
        pts_per_step = total_raw / grow_steps
        n_lit, t_lit = cpu_sample(weights, p0, o0, args.cpu_baseline_steps, literal=True)
        n_vec, t_vec = cpu_sample(weights, p0, o0, args.cpu_baseline_steps * 2, literal=False)
        cpu_baseline = {'value': n_lit / t_lit * pts_per_step, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
                        'sample': 'first %d grow steps of room 0 (%d pts), oracle port with the reference\'s literal per-point update loop, numpy forward on all host threads; points/s = steps/s x %.2f raw points per grow step of this workload' % (n_lit, len(p0), pts_per_step),
                        'grow_steps_per_sec': n_lit / t_lit,
                        'vectorised_port': {'value': n_vec / t_vec * pts_per_step, 'grow_steps_per_sec': n_vec / t_vec}}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': 'area5_synthetic_%d_rooms_20k_raw' % args.rooms, 'rooms_per_gpu': args.rooms, 'resolution': 0.1,
                       'equalized_points_per_gpu': int(offsets[-1]), 'raw_points_per_gpu': total_raw, 'l2': 'flushed between steps (256 MB fill)',
                       'scope': 'raw points resident in HBM -> device feature preparation (test_region_grow.py:119-173) -> grow driver + LrgNet forward -> fill',
                       'rng': 'philox4x32-10 seed 0', 'weights': 'lrgnet_model5 (golden)'},
            'feature_prep_ms_per_pass': float(np.mean(prep_ms_list)), 'grow_ms_per_pass': float(np.mean(grow_ms_list)),
            'grow_steps_per_sec': world * grow_steps / (ms_per_step * 1e-3), 'grow_steps_per_pass': grow_steps,
            'longest_room_steps': int(stats['grow_steps'].max()),
            'wall_s_timed_region': wall, 'clocks': clocks, 'gpu_launches': int(launches),
            'e2e': e2e_raw,
            'e2e_features': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                             'ms_per_step': [round(x, 2) for x in e2e_ms],
                             'scope': '13-D features + seed order prepared beforehand, in pinned host memory -> grow -> fill -> labels per equalised point on the host'},
            'roofline': roofline, 'cpu_baseline': cpu_baseline, 'extras': extras,
            'flops_per_grow_step': FLOPS_PER_STEP,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
