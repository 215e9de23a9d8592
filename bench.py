#!/usr/bin/env python
"""Benchmark of the LRGNet grow engine on synthetic S3DIS-shaped rooms (BASELINE.json metric: segmented points/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config {1,2,3,4,5}] [--impl reference]

One "step" = one full pass of the hot path over the workload: raw points -> feature preparation on the device (the
reference's points/s timer starts before it, test_region_grow.py:120,317) -> every room grown to completion -> fill.
`value` is measured with the raw points already resident in HBM (CUDA events on the engine stream, max over ranks); `e2e` is
the same pass through the public host-buffer calls (raw points in pinned host memory in, per-raw-point labels out, copies
inside the timed region).

Workloads (BASELINE.json `configs`), selected with --config:
  2 (default)  "--area 5": 68 synthetic Area-5-shaped rooms (~20k raw points, seeds 1000+room) PER GPU -- weak scaling, the
               line BENCH / SCALE record.  Its `extras` carry the other configurations measured in the same run:
               config1 (one grow step), config3, config4_strong (the 272 rooms sharded over the N ranks), config5, the
               calibrated workload, the local-search drivers, the statistics kernels and the tf_ops.
  4            S3DIS areas 1-6: 272 rooms (seeds 1000+global room id) sharded over the ranks by room (LPT on the raw point
               counts, parallel.shard_rooms), one all-gather of labels -- STRONG scaling.
  3            "--area scannet": 312 rooms with log-uniform raw sizes in [5k, 60k] (seeds 2000+room), sharded the same way.
  5            Semantic-KITTI-shaped: 19 scenes of ~300k raw points at resolution 0.3 (seeds 3000+scene), sharded by scene.
  1            one LrgNet grow step: us per forward call / per resident grow step beside the numpy forward.
`--impl reference` times the reference's CPU implementation of the same pass (see run_reference_arm).
"""
import os
import sys

if '--impl' in sys.argv and 'reference' in sys.argv:
    # the CPU arm uses every host thread; torchrun exports OMP_NUM_THREADS=1 -- undo that before numpy loads its BLAS
    for _k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_k] = str(os.cpu_count())

import argparse
import json
import subprocess
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

FLOPS_PER_STEP = 271712256                     # BASELINE.md section 4 (algorithmic, factored heads)
BRANCH_TILE_FLOPS = 2 * 128 * 82752                      # one 128-row tile of one branch
HEAD_TILE_FLOPS = 2 * 128 * (64 * 256 + 256 * 128)       # tensor part of one 128-row head tile (128->2 runs on the FMA pipe)
METRIC = 'segmented_points_per_sec'
UNIT = 'points/s'
# The "calibrated" variant of the room generator (tools/calib_probe.py, profiles/r2j_calib.txt): 13 boxes per room instead of
# 20-40 bring a room to the ~50 clusters / ~950 grow steps of the reference's S3DIS logs; noise as the reference's.
CALIBRATED = dict(n_boxes=13)


def load_peaks():
    try:
        return json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'sm_max_mhz': 1965.0}, 'fallback'


def all_host_threads():
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass


# ----------------------------------------------------------------------------------------------- workloads
def _room(kind, g, calibrated=False):
    from tools import rooms
    if kind == 'kitti':
        return rooms.generate_outdoor_scene(3000 + g)
    if kind == 'scannet':
        rs = np.random.RandomState(2000 + g)
        n = int(np.exp(rs.uniform(np.log(5000), np.log(60000))))
        return rooms.generate_room(2000 + g, n_raw=n)
    return rooms.generate_room(1000 + g, **(CALIBRATED if calibrated else {}))


def concat_rooms(rows):
    raw_off = np.zeros(len(rows) + 1, np.int64)
    np.cumsum([len(r) for r in rows], out=raw_off[1:])
    pts = np.ascontiguousarray(np.concatenate(rows), np.float32) if rows else np.zeros((0, 8), np.float32)
    return raw_off, pts


def make_workload(n_rooms, seed_base, cache=True):
    """Synthetic raw rooms (x y z r g b obj_id cls_id rows) seeds seed_base..seed_base+n_rooms-1:
    (raw_offsets (R+1) int64, raw_points (sum Nr, 8) float32)."""
    path = '/tmp/lrg_bench_rooms_v3_%d_%d.npz' % (n_rooms, seed_base)
    if cache and os.path.exists(path):
        z = np.load(path)
        return z['raw_offsets'], z['raw_points']
    raw_off, raw_points = concat_rooms([_room('s3dis', seed_base - 1000 + r) for r in range(n_rooms)])
    if cache:
        try:
            np.savez(path, raw_offsets=raw_off, raw_points=raw_points)
        except Exception:
            pass
    return raw_off, raw_points


WORKLOADS = {
    # config: (name, kind, total units, resolution, scaling)
    2: ('area5_synthetic_%d_rooms_20k_raw', 's3dis', None, 0.1, 'weak'),
    3: ('scannet_synthetic_312_rooms_5k_60k_raw', 'scannet', 312, 0.1, 'strong'),
    4: ('s3dis_areas_1_6_synthetic_272_rooms_20k_raw', 's3dis', 272, 0.1, 'strong'),
    5: ('semantic_kitti_synthetic_19_scenes_300k_raw_res0.3', 'kitti', 19, 0.3, 'strong'),
}


def sharded_workload(config, rank, world):
    """The rooms of a strong-scaling configuration that fall to this rank: every rank derives the same LPT shard table from
    the raw point counts (parallel.shard_rooms), then keeps its own rooms.  Returns (global ids, rows, all counts, shards)."""
    from learn_region_grow_b200 import parallel
    name, kind, total, res, _ = WORKLOADS[config]
    rows = [_room(kind, g) for g in range(total)]
    counts = np.array([len(r) for r in rows], np.int64)
    shards = parallel.shard_rooms(counts, world)
    mine = [int(g) for g in shards[rank]]
    return mine, [rows[g] for g in mine], counts, shards


# ----------------------------------------------------------------------------------------------- helpers
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


def load_weights():
    with np.load(os.path.join(REPO, 'tests', 'golden', 'lrgnet_model5.npz')) as z:
        return {k: z[k] for k in z.files}


# ----------------------------------------------------------------------------------------------- CPU arms
def cpu_room(weights, raw_room, resolution, literal=False, max_steps=None, room_id=0):
    """The reference's CPU path on ONE room, whole: raw points -> host feature preparation (test_region_grow.py:119-173) -> grow
    (:175-306) -> fill (:308-316), by the oracle port (oracle/feature_prep.py, oracle/lrg_driver.py, numpy forward on all host
    threads).  TEST INFRASTRUCTURE used as the baseline.  Returns (grow steps, seconds, seconds of the preparation alone)."""
    from oracle import feature_prep, lrg_driver, lrg_forward
    fwd = lambda a, b: lrg_forward.forward(weights, a, b)
    t0 = time.perf_counter()
    f = feature_prep.prepare_features(raw_room, resolution)
    t_prep = time.perf_counter() - t0
    g = lrg_driver.RoomGrower(f['points'], f['order'].astype(np.int32), fwd, lrg_driver.PhiloxRng(0), resolution=resolution, room_id=room_id,
                              literal_update=literal)
    if max_steps is None:
        g.run()
        g.fill()
    else:
        for seed_id in np.arange(len(f['points']))[f['order']]:
            if g.visited[seed_id]:
                continue
            g.begin_region(seed_id)
            while g.total_steps < max_steps:
                st = g.prepare_step()
                if st is None:
                    break
                add, rmv = fwd(st['inlier'], st['neighbor'])
                if g.apply_step(add[0], rmv[0]) is not None:
                    break
            if g.total_steps >= max_steps:
                break
    return g.total_steps, time.perf_counter() - t0, t_prep


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the SAME pass on the same workload, one whole room per timed
    step (raw points -> feature preparation -> grow -> fill; points/s = the rooms' raw points / their time), on all host threads.
    The Python reference itself cannot run on the GPU box (TensorFlow / h5py are not installable and /root/reference does not
    travel), so this is the oracle PORT of test_region_grow.py -- with its voxel-set update vectorised (one numpy.isin per step
    instead of the reference's per-point Python loop, :282-287) and the feature preparation vectorised, i.e. FASTER than the
    reference; the literal loop is sampled beside it (`literal_grow_steps_per_sec`).  Warm-up steps are bounded samples (40 grow
    steps)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    all_host_threads()
    config = args.config if args.config in WORKLOADS else 2
    name, kind, total, res, scaling = WORKLOADS[config]
    n_units = args.rooms if config == 2 else total
    cap = 400 if config == 5 else None              # (a 300k-point scene does not finish in minutes on the host: bounded sample)
    weights = load_weights()
    times, nsteps, npts, preps = [], [], [], []
    for it in range(args.warmup + args.steps):
        g = it % n_units
        raw = _room(kind, g)
        if it < args.warmup:
            cpu_room(weights, raw, res, max_steps=40, room_id=g)
            continue
        n, dt, tp = cpu_room(weights, raw, res, room_id=g, max_steps=cap)
        times.append(dt); nsteps.append(n); npts.append(len(raw)); preps.append(tp)
    if cap is None:
        value = float(sum(npts) / sum(times))
        sample = '%d whole rooms of the workload, one per step (%d raw points, %d grow steps in all)' % (len(times), sum(npts), sum(nsteps))
    else:
        # bounded sample: points/s = steps/s x raw points per grow step of the workload (profiles/workload_stats.json)
        pps = 300000.0 / 25000.0
        try:
            s = json.load(open(os.path.join(REPO, 'profiles', 'workload_stats.json')))[name]
            pps = float(s['raw_points']) / float(s['grow_steps'])
        except Exception:
            pass
        value = float(sum(nsteps) / sum(times) * pps)
        sample = 'first %d grow steps of %d scenes; points/s = grow steps/s x %.2f raw points per grow step of the workload' % (cap, len(times), pps)
    n_lit, t_lit, tp_lit = cpu_room(weights, _room(kind, 0), res, literal=True, max_steps=60)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * float(np.mean(times)), 'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': (name % args.rooms) if config == 2 else name, 'config': config, 'resolution': res,
                   'scope': 'raw points -> feature preparation (test_region_grow.py:119-173) -> grow driver + LrgNet forward -> fill, one whole room per step',
                   'rooms_timed': [it % n_units for it in range(args.warmup, args.warmup + args.steps)]},
        'grow_steps_per_sec': float(sum(nsteps) / sum(times)), 'feature_prep_share': float(sum(preps) / sum(times)),
        'literal_grow_steps_per_sec': float(n_lit / max(t_lit - tp_lit, 1e-9)),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
                         'sample': sample + ': oracle port of test_region_grow.py with the voxel-set update vectorised, numpy forward on all host threads'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- GPU arm
class Rig:
    """One rank's engine, stream-side helpers and timing loop."""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        import torch
        self.torch = torch
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group('nccl', rank=self.rank, world_size=self.world, device_id=torch.device('cuda', self.local_rank))
            self.dist = dist
        from learn_region_grow_b200 import _lib, parallel
        from learn_region_grow_b200.engine import Engine
        self._lib, self.parallel = _lib, parallel
        self.weights = load_weights()
        self.eng = Engine(1, 1, 512, 512, 13, 0, device=self.local_rank)
        self.eng.load_weights(self.weights)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')      # > 126 MB L2

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, vals, op):
        if self.dist is None:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device='cuda', dtype=self.torch.float64)
        self.dist.all_reduce(t, op=op)
        return [float(x) for x in t]

    def allmax(self, *vals):
        return self._reduce(vals, self.dist.ReduceOp.MAX if self.dist else None)

    def allsum(self, *vals):
        return self._reduce(vals, self.dist.ReduceOp.SUM if self.dist else None)

    def allgather_rows(self, vals):
        if self.dist is None:
            return [[float(v) for v in vals]]
        t = self.torch.tensor(list(vals), device='cuda', dtype=self.torch.float64)
        allt = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(allt, t)
        return [[float(x) for x in row] for row in allt]

    def gather_labels(self, n_eq, lengths):
        """The path's one collective: all-gather of the int32 instance labels (padded to the longest shard)."""
        if self.dist is None:
            return None
        local = self.torch.as_tensor(self.parallel.DeviceArray(self.eng.labels_device_ptr(True), int(n_eq)), device='cuda') if n_eq > 0 else \
            self.torch.zeros(0, dtype=self.torch.int32, device='cuda')
        return self.parallel.allgather_labels(local, lengths)

    def timed_passes(self, raw_off, raw_points, resolution, params, steps, warmup, sampler=None):
        """W warm-up + K timed passes with the raw points resident in HBM; per-pass device time = CUDA events on the engine stream
        for preparation / grow / fill + the label all-gather (the caller takes the max over ranks)."""
        torch, eng = self.torch, self.eng
        total_raw = int(raw_off[-1])
        d_raw = torch.from_numpy(np.ascontiguousarray(raw_points, np.float32)).cuda()
        offsets = eng.upload_raw_concatenated(raw_off, d_raw, resolution)
        lengths = [int(r[0]) for r in self.allgather_rows([int(offsets[-1])])]
        dev_ms, grow_ms, prep_ms_l, fill_ms = [], [], [], []
        launches = 0
        stats = None
        wall0 = time.perf_counter()
        for it in range(warmup + steps):
            self.flush.fill_(it & 0xFF)
            if it == warmup:
                self.barrier()
                if sampler is not None:
                    sampler.start()
                wall0 = time.perf_counter()
            eng.upload_raw_concatenated(raw_off, d_raw, resolution)       # device feature preparation from the resident raw points
            prep_ms = eng.prepare_ms()
            stats = eng.segment_resident(resolution=resolution, **params)
            t_ag0 = time.perf_counter()
            self.gather_labels(offsets[-1], lengths)
            torch.cuda.synchronize()
            t_ag = time.perf_counter() - t_ag0
            if it >= warmup:
                pr = eng.profile()
                dev_ms.append(prep_ms + pr['grow_ms'] + pr['fill_ms'] + (1e3 * t_ag if self.dist is not None else 0.0))
                grow_ms.append(pr['grow_ms']); prep_ms_l.append(prep_ms); fill_ms.append(pr['fill_ms'])
                # grow (+ lock-step kernels) + fill (2) + flag reset + raw labels, and what the upload launched (feature preparation
                # with its sort passes, pack, spatial index) -- counted by the library
                launches += pr['kernel_launches'] + 2 + eng.prepare_launches()
        self.barrier()
        wall = time.perf_counter() - wall0
        return dict(total_raw=total_raw, offsets=offsets, lengths=lengths, dev_ms=dev_ms, grow_ms=grow_ms, prep_ms=prep_ms_l, fill_ms=fill_ms,
                    launches=launches, stats=stats, wall=wall, profile=eng.profile())

    def e2e_passes(self, raw_off, raw_points, resolution, params, steps, offsets, lengths):
        """The same pass end to end: RAW points in pinned host memory -> device feature preparation -> growing -> fill ->
        per-raw-point labels back on the host (+ the all-gather); wall clock, copies inside the timed region."""
        eng = self.eng
        h_raw = pinned_array(self._lib, raw_points.shape, np.float32)
        h_raw[...] = raw_points
        ms = []
        for it in range(1 + steps):
            self.barrier()
            t_it = time.perf_counter()
            eng.upload_raw_concatenated(raw_off, h_raw, resolution)
            eng.segment_resident(resolution=resolution, **params)
            eng.raw_labels(True)
            self.gather_labels(offsets[-1], lengths)
            self.barrier()
            if it >= 1:
                ms.append(1e3 * (time.perf_counter() - t_it))
        return ms, h_raw


def roofline_of(res, peaks, peaks_src, n_rooms):
    """The dominant kernel -- the persistent grow kernel, one launch per pass -- against the measured peaks."""
    pr, stats = res['profile'], res['stats']
    grow_steps = int(stats['grow_steps'].sum()) if len(stats) else 0
    wasted = int(stats['spec_wasted_steps'].sum()) if len(stats) else 0
    peak_tf = peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops'))
    grow_ms = float(np.mean(res['grow_ms']))
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
        'note': 'algorithmic = 271.7 MFLOP per committed grow step (512+512 rows, factored heads); the kernel evaluates only distinct rows '
                '(padding duplicates reuse logits) as 3xFP16 on tcgen05 (kind::f16, A operand in TMEM), plus the steps of discarded '
                'speculative attempts; grid = one CTA per SM (cooperative launch): 132 work-item CTAs + 16 pooled-projection server CTAs '
                '(weights resident in registers); neither roofline binds -- the pass is bounded by per-room chains of dependent steps',
        'speculative_steps_discarded_per_pass': wasted,
    }
    # the driver phases' side of the roofline: algorithmic bytes per grow step = one pass over the room's state
    # (14 B per point) + the gathered tiles and logits (62,464 B) + the fp32 weights amortised over the rooms stepped together
    n_eq_mean = float(res['offsets'][-1]) / max(n_rooms, 1)
    bytes_step = 14.0 * n_eq_mean + 62464.0 + 3164176.0 / max(n_rooms, 1)
    hbm_peak = peaks.get('hbm_gbs')
    roofline['hbm_side'] = {'algorithmic_bytes_per_grow_step': bytes_step, 'achieved_gbs': grow_steps * bytes_step / (grow_ms * 1e-3) / 1e9,
                            'peak_gbs': hbm_peak, 'frac': (grow_steps * bytes_step / (grow_ms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None}
    if pr['persistent']:
        items, busy = pr['items'], pr['busy_ms']
        executed = 3.0 * (items['branch'] * BRANCH_TILE_FLOPS + items['head'] * HEAD_TILE_FLOPS)     # (upper bound: the parts of a split branch tile count as tiles)
        roofline.update({
            'items_per_pass': items, 'busy_ms_by_item': busy,
            'avg_us_per_item': {k: (1e3 * busy[k] / items[k] if items[k] else None) for k in items},
            'sm_busy_frac': sum(busy.values()) / (148.0 * grow_ms),
            'executed_f16_tflops_upper_bound': executed / (grow_ms * 1e-3) / 1e12,
        })
        try:
            # the reference's own timing buckets (comp_time_analysis, test_region_grow.py:40-51,120-317), per pass: 'feature' =
            # feature preparation (device time); 'net' = the forward (branch + head handlers), 'neighbor' + 'inlier' = the driver
            # step handler -- SM time summed over CTAs, since the phases of different rooms overlap on the device
            net_ms = float(busy['branch'] + busy['gproj'] + busy['head'])
            drv_ms = float(busy['step'])
            roofline['reference_buckets'] = {
                'feature_ms_device': float(np.mean(res['prep_ms'])), 'net_sm_ms': net_ms, 'neighbor_plus_inlier_sm_ms': drv_ms,
                'net_share_of_sm_time': net_ms / (net_ms + drv_ms) if (net_ms + drv_ms) > 0 else None,
                'note': 'SM time summed over CTAs (phases of different rooms overlap on the device, so there is no per-phase wall split); '
                        'the pooled projection runs on 16 server CTAs and is not in net_sm_ms'}
        except Exception:
            pass
    return roofline


def measure_workload(rig, config, args, peaks, peaks_src, steps, warmup, with_e2e=True, sampler=None):
    """One configuration on this rank's share of the rooms; returns the fields of a bench line (rank 0 prints them)."""
    name, kind, total, res, scaling = WORKLOADS[config]
    if config == 2:
        rows = None
        raw_off, raw_points = make_workload(args.rooms, 1000 + rig.rank * args.rooms)
        n_rooms_local, room_base, n_units = args.rooms, rig.rank * args.rooms, args.rooms * rig.world
        name = name % args.rooms
        shard_info = None
    else:
        mine, rows, counts, shards = sharded_workload(config, rig.rank, rig.world)
        raw_off, raw_points = concat_rooms(rows)
        n_rooms_local, room_base, n_units = len(rows), 0, total
        shard_info = {'rooms_per_rank': [len(s) for s in shards], 'raw_points_per_rank': [int(counts[s].sum()) for s in shards],
                      'partition': 'LPT on raw point counts (parallel.shard_rooms)'}
    params = dict(seed=0, max_slots=args.slots, room_id_base=room_base, spec_lanes=args.spec_lanes)
    r = rig.timed_passes(raw_off, raw_points, res, params, steps, warmup, sampler)
    mine_ms = float(sum(r['dev_ms'])) / steps
    total_ms, wall = rig.allmax(float(sum(r['dev_ms'])), r['wall'])
    ms_per_step = total_ms / steps
    stats = r['stats']
    grow_steps_local = int(stats['grow_steps'].sum()) if len(stats) else 0
    longest_local = float(stats['grow_steps'].max()) if len(stats) else 0.0
    raw_all, steps_all, eq_all = rig.allsum(r['total_raw'], grow_steps_local, int(r['offsets'][-1]))
    longest = rig.allmax(longest_local)[0]
    out = {
        'metric': METRIC, 'value': raw_all / (ms_per_step * 1e-3), 'unit': UNIT, 'n_gpus': rig.world, 'steps': steps, 'warmup': warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': name, 'config': config, 'units': int(n_units), 'resolution': res, 'raw_points': int(raw_all), 'equalized_points': int(eq_all),
                   'l2': 'flushed between steps (256 MB fill)', 'rng': 'philox4x32-10 seed 0', 'weights': 'lrgnet_model5 (golden)',
                   'arithmetic': '3xFP16 split products on tcgen05, fp32 accumulation (logits within 2e-4 of the fp64 graph)',
                   'scope': 'raw points resident in HBM -> device feature preparation (test_region_grow.py:119-173) -> grow driver + LrgNet forward -> fill'
                            + (' -> one all-gather of labels' if rig.world > 1 else '')},
        'feature_prep_ms_per_pass': float(np.mean(r['prep_ms'])), 'grow_ms_per_pass': float(np.mean(r['grow_ms'])), 'fill_ms_per_pass': float(np.mean(r['fill_ms'])),
        'grow_steps_per_sec': steps_all / (ms_per_step * 1e-3), 'grow_steps_per_pass': int(steps_all), 'longest_room_steps': int(longest),
        'wall_s_timed_region': wall, 'gpu_launches': int(r['launches']),
    }
    if shard_info is not None:
        out['config']['sharding'] = shard_info
        rowsr = rig.allgather_rows([mine_ms, longest_local, float(np.mean(r['grow_ms']))])
        out['per_rank'] = {'ms_per_pass': [x[0] for x in rowsr], 'longest_room_steps': [int(x[1]) for x in rowsr], 'grow_ms_per_pass': [x[2] for x in rowsr],
                           'note': 'the value follows the slowest rank (max over ranks): with rooms sharded a rank is bounded by its longest room\'s chain'}
    out['roofline'] = roofline_of(r, peaks, peaks_src, n_rooms_local)
    if with_e2e:
        ms, h_raw = rig.e2e_passes(raw_off, raw_points, res, params, steps, r['offsets'], r['lengths'])
        raw_s = rig.allmax(float(np.mean(ms)) * 1e-3)[0]
        out['e2e'] = {'value': raw_all / raw_s, 'unit': UNIT, 'h2d_bytes_per_step': int(h_raw.nbytes + raw_off.nbytes),
                      'd2h_bytes_per_step': int(r['total_raw'] * 4 + n_rooms_local * 48), 'ms_per_step': [round(x, 2) for x in ms],
                      'scope': 'raw points (x y z r g b) in pinned host memory -> device feature preparation -> grow -> fill -> per-raw-point labels on the host'
                               + (' -> all-gather' if rig.world > 1 else '') + '; bytes are per rank'}
    out['_local'] = dict(raw_off=raw_off, raw_points=raw_points, params=params, res=res, rows=rows)
    return out


def config1(rig):
    """BASELINE.json configs[0]: one room, ONE LrgNet grow step.  us per forward call (host tiles in / logits out; tiles resident),
    us per grow step of a room alone on the device (driver step + forward, no host round trip), beside the numpy forward."""
    from oracle import lrg_forward
    eng, torch = rig.eng, rig.torch
    z = np.load(os.path.join(REPO, 'tests', 'golden', 'forward_graphdef.npz'))
    inl, nb = np.ascontiguousarray(z['inlier'][:1]), np.ascontiguousarray(z['neighbor'][:1])
    for _ in range(5):
        eng.forward(inl, nb)
    t0 = time.perf_counter()
    for _ in range(200):
        add, rmv = eng.forward(inl, nb)
    host_us = (time.perf_counter() - t0) / 200 * 1e6
    err = float(max(np.abs(add - z['add_f64'][:1]).max(), np.abs(rmv - z['remove_f64'][:1]).max()))
    d_i, d_n = torch.from_numpy(inl).cuda(), torch.from_numpy(nb).cuda()
    d_a, d_r = torch.zeros(1, 512, 2, device='cuda'), torch.zeros(1, 512, 2, device='cuda')
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        eng.forward_device(d_i, d_n, d_a, d_r, 1, stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(200):
        eng.forward_device(d_i, d_n, d_a, d_r, 1, stream)
    e1.record()
    torch.cuda.synchronize()
    dev_us = e0.elapsed_time(e1) * 1e3 / 200
    # one room alone, resident: the whole grow step on the device
    raw_off, raw = make_workload(1, 1000)
    eng.upload_raw_concatenated(raw_off, raw, 0.1)
    per = {}
    for lanes in (1, 4):
        eng.segment_resident(resolution=0.1, seed=0, spec_lanes=lanes)
        st = eng.segment_resident(resolution=0.1, seed=0, spec_lanes=lanes)
        per[lanes] = 1e3 * eng.profile()['grow_ms'] / max(int(st['grow_steps'].sum()), 1)
    all_host_threads()
    W = rig.weights
    for _ in range(2):
        lrg_forward.forward(W, inl, nb)
    t0 = time.perf_counter()
    for _ in range(10):
        lrg_forward.forward(W, inl, nb)
    np_us = (time.perf_counter() - t0) / 10 * 1e6
    return {'forward_host_us_per_call': host_us, 'forward_device_us_per_call': dev_us, 'max_abs_logit_error_vs_shipped_graph_f64': err,
            'resident_room_us_per_grow_step': per[1], 'resident_room_us_per_committed_grow_step_4_lanes': per[4],
            'numpy_forward_us_per_call': np_us, 'host_cores': os.cpu_count(),
            'scope': 'one (1,512,13)+(1,512,13) tile pair of the driver (tests/golden/forward_graphdef.npz); lrg_forward_host = H2D + 3 kernels + D2H, '
                     'lrg_forward_device = 3 kernels; resident = one 20k-point room alone on the GPU, grow kernel time / committed grow steps'}


def tfops_extras(rig, peaks):
    """tf_ops at the PointNet++ call shapes of the reference's consumer (train_pointnet.py:181-190), through the C ABI on device
    buffers: us per launch and algorithmic GB/s against the measured copy bandwidth.  (The reference's own kernels are timed beside
    ours by tests/test_tfops_perf_gpu.py -> profiles/*tfops*; the bench does not execute anything under oracle/ here.)"""
    import ctypes as C
    torch = rig.torch
    L = rig._lib.lib()
    chk = rig._lib.check
    p = lambda t: C.c_void_p(t.data_ptr())

    def timeit(fn, reps=20, warm=3):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    hbm = peaks.get('hbm_gbs') or 1.0
    rng = np.random.RandomState(5)
    out = {}
    for B in (1, 100):
        n, m, ns, c, radius = 1024, 256, 32, 64, 0.2
        xyz = torch.from_numpy(rng.rand(B, n, 3).astype(np.float32)).cuda()
        feat = torch.from_numpy(rng.randn(B, n, c).astype(np.float32)).cuda()
        idx = torch.zeros(B, m, dtype=torch.int32, device='cuda')
        q = torch.zeros(B, m, 3, device='cuda')
        gi = torch.zeros(B, m, ns, dtype=torch.int32, device='cuda')
        cnt = torch.zeros(B, m, dtype=torch.int32, device='cuda')
        g = torch.zeros(B, m, ns, c, device='cuda')
        ops = [
            ('farthest_point_sample', lambda: chk(L.lrg_farthest_point_sampling(B, n, m, p(xyz), None, p(idx), None)), B * (n * 12 + m * 4)),
            ('gather_point', lambda: chk(L.lrg_gather_point(B, n, m, p(xyz), p(idx), p(q), None)), B * m * 28),
            ('query_ball_point', lambda: chk(L.lrg_query_ball_point(B, n, m, radius, ns, p(xyz), p(q), p(gi), p(cnt), None)), B * (n * 12 + m * (16 + 4 * ns))),
            ('group_point', lambda: chk(L.lrg_group_point(B, n, c, m, ns, p(feat), p(gi), p(g), None)), B * m * ns * (4 + 8 * c)),
        ]
        nx = torch.zeros(B, m, 3, device='cuda')
        npts = torch.zeros(B, m, ns, 3 + c, device='cuda')
        gx = torch.zeros(B, m, ns, 3, device='cuda')
        ops.append(('sample_and_group', lambda: chk(L.lrg_sample_and_group(B, n, m, radius, ns, c, p(xyz), p(feat), None, p(idx), p(nx), p(npts), p(gi), p(cnt),
                                                                           p(gx), None)), B * (n * (12 + 4 * c) + m * ns * (4 + 4 * (3 + c)))))
        row = {}
        for name, fn, nbytes in ops:
            us = timeit(fn)
            row[name] = {'us': us, 'algorithmic_gbs': nbytes / us / 1e3, 'frac_of_hbm_peak': nbytes / us / 1e3 / hbm}
        x1 = torch.from_numpy(rng.rand(B, n, 3).astype(np.float32)).cuda()
        x2 = torch.from_numpy(rng.rand(B, m, 3).astype(np.float32)).cuda()
        d3 = torch.zeros(B, n, 3, device='cuda')
        k3 = torch.zeros(B, n, 3, dtype=torch.int32, device='cuda')
        w3 = torch.full((B, n, 3), 1.0 / 3.0, device='cuda')
        pts = torch.from_numpy(rng.randn(B, m, 256).astype(np.float32)).cuda()
        o3 = torch.zeros(B, n, 256, device='cuda')
        us = timeit(lambda: chk(L.lrg_three_nn(B, n, m, p(x1), p(x2), p(d3), p(k3), None)))
        row['three_nn'] = {'us': us, 'algorithmic_gbs': B * (n * 36 + m * 12) / us / 1e3, 'frac_of_hbm_peak': B * (n * 36 + m * 12) / us / 1e3 / hbm}
        us = timeit(lambda: chk(L.lrg_three_interpolate(B, m, 256, n, p(pts), p(k3), p(w3), p(o3), None)))
        nb = B * (n * (24 + 4 * 256) + m * 256 * 4)
        row['three_interpolate'] = {'us': us, 'algorithmic_gbs': nb / us / 1e3, 'frac_of_hbm_peak': nb / us / 1e3 / hbm}
        out['B%d_n1024_m256' % B] = row
    # one large cloud: 65,536 points in the registers of a cluster of 8 CTAs (argmax through distributed shared memory)
    n, m = 65536, 256
    xyz = torch.from_numpy(rng.rand(1, n, 3).astype(np.float32)).cuda()
    idx = torch.zeros(1, m, dtype=torch.int32, device='cuda')
    us = timeit(lambda: chk(L.lrg_farthest_point_sampling(1, n, m, p(xyz), None, p(idx), None)), reps=5, warm=1)
    out['B1_n65536_m256'] = {'farthest_point_sample': {'us': us, 'us_per_round': us / m, 'algorithmic_gbs': (n * 12 + m * 4) / us / 1e3,
                                                       'kernel': 'lrg_fps_cluster_kernel<8> (thread-block cluster, DSMEM exchange)'}}
    out['note'] = 'CUDA events, 20 launches after 3 warm-up; arrays of a few MB are L2-resident (stated, not flushed); hbm peak = MEASURED_PEAKS hbm_gbs'
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', type=int, default=2, choices=[1, 2, 3, 4, 5], help='BASELINE.json configuration (see the module docstring)')
    ap.add_argument('--rooms', type=int, default=68, help='config 2: rooms per GPU (Area 5 has 68)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='config 2: skip the other configurations, the local-search drivers, statistics and tf_ops')
    ap.add_argument('--slots', type=int, default=0)
    ap.add_argument('--spec-lanes', type=int, default=0, help='speculative lanes per room (0 = engine default, 1 = off)')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        return run_reference_arm(args)

    rig = Rig(args)
    eng, rank, world = rig.eng, rig.rank, rig.world
    peaks, peaks_src = load_peaks()
    sampler = ClockSampler(rig.local_rank)

    if args.config == 1:
        c1 = config1(rig)
        if rank == 0:
            us = c1['resident_room_us_per_committed_grow_step_4_lanes']
            print(json.dumps({'metric': 'grow_steps_per_sec', 'value': 1e6 / us, 'unit': 'grow steps/s', 'n_gpus': 1, 'steps': 1, 'warmup': 1,
                              'ms_per_step': us * 1e-3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                              'config': {'workload': 'one_room_one_grow_step', 'config': 1}, 'config1': c1}))
        return

    line = measure_workload(rig, args.config, args, peaks, peaks_src, args.steps, args.warmup, with_e2e=True, sampler=sampler)
    local = line.pop('_local')
    extras = {}
    if args.config == 2 and not args.no_extras:
        total_raw = int(local['raw_off'][-1])
        params = local['params']
        # (the raw rooms of the last e2e pass are still uploaded)
        if world == 1:
            obj_raw = [local['raw_points'][local['raw_off'][i]:local['raw_off'][i + 1], 6].astype(np.int32) for i in range(args.rooms)]
            eng.room_metrics(obj_raw, raw=True)
            t0 = time.perf_counter()
            mt = eng.room_metrics(obj_raw, raw=True)
            extras['statistics'] = {'ms_per_call': 1e3 * (time.perf_counter() - t0), 'rooms': args.rooms,
                                    'mean': {k: float(np.nanmean(mt[k])) for k in ('nmi', 'ami', 'ars', 'prc', 'rcl', 'iou')},
                                    'scope': 'obj_id of the raw points in, NMI/AMI/ARS/PRC/RCL/IOU per room out (test_region_grow.py:319-349)'}
            for key, kw, scope in (('random_restart', dict(num_restarts=10), 'test_random_restart.py: 10 restarts per seed as parallel lanes, largest region kept'),
                                   ('beam_search', dict(beam_width=3, search_width=3), 'test_beam_search.py: 3 candidates x 3 expansions per round as parallel lanes'),
                                   ('beam_search_ml', dict(beam_width=3, search_width=3, scoring='ml'),
                                    'test_beam_search.py --scoring ml: candidates ranked by accumulated log-probability (:238-256,263-264)')):
                kw2 = dict(params)
                kw2.pop('spec_lanes', None)
                kw2.update(kw)
                eng.segment_resident(resolution=0.1, **kw2)
                st2 = eng.segment_resident(resolution=0.1, **kw2)
                pr2 = eng.profile()
                m2 = eng.room_metrics(obj_raw, raw=True)
                extras[key] = dict(kw, grow_ms_per_pass=pr2['grow_ms'], grow_steps_per_pass=int(st2['grow_steps'].sum()),
                                   grow_steps_per_sec=float(st2['grow_steps'].sum()) / (pr2['grow_ms'] * 1e-3),
                                   points_per_sec=total_raw / ((pr2['grow_ms'] + pr2['fill_ms']) * 1e-3),
                                   mean={k: float(np.nanmean(m2[k])) for k in ('nmi', 'ami', 'ars', 'prc', 'rcl', 'iou')}, scope=scope)
            # one lane per room (no speculation): what the same kernel does when every room is a single chain
            eng.segment_resident(resolution=0.1, **dict(params, spec_lanes=1))
            st1 = eng.segment_resident(resolution=0.1, **dict(params, spec_lanes=1))
            extras['no_speculation'] = {'grow_ms_per_pass': eng.profile()['grow_ms'], 'grow_steps_per_pass': int(st1['grow_steps'].sum())}
            extras['config1'] = config1(rig)
            extras['tfops'] = tfops_extras(rig, peaks)
            # the calibrated variant of the same generator (~50 clusters / ~950 grow steps per room like the reference's S3DIS logs)
            raw_off_c, raw_c = concat_rooms([_room('s3dis', g, calibrated=True) for g in range(args.rooms)])
            rc = rig.timed_passes(raw_off_c, raw_c, 0.1, params, 2, 1)
            stc = rc['stats']
            ms_c = float(np.mean(rc['dev_ms']))
            extras['calibrated_workload'] = {
                'generator': CALIBRATED, 'value': int(raw_off_c[-1]) / (ms_c * 1e-3), 'unit': UNIT, 'ms_per_pass': ms_c,
                'grow_steps_per_pass': int(stc['grow_steps'].sum()), 'regions_per_room': float(stc['regions'].mean()), 'clusters_per_room': float(stc['clusters'].mean()),
                'grow_steps_per_room': float(stc['grow_steps'].mean()), 'longest_room_steps': int(stc['grow_steps'].max()),
                'grow_steps_per_sec': float(stc['grow_steps'].sum()) / (ms_c * 1e-3),
                'note': 'the default rooms (20-40 boxes) over-segment: ~2,650 grow steps and ~120 clusters per room against 948 / ~50 in the '
                        'reference\'s S3DIS logs; this variant of the same generator is closer to that balance of chain length and throughput'}
        # BASELINE.json configs 4 (strong scaling over the ranks), 3 and 5: two timed passes each after one warm-up
        for cfg, key in ((4, 'config4_strong'), (3, 'config3'), (5, 'config5')):
            if cfg != 4 and world > 1:
                continue
            ln = measure_workload(rig, cfg, args, peaks, peaks_src, 2, 1, with_e2e=False)
            ln.pop('_local')
            keep = ('value', 'unit', 'scaling', 'n_gpus', 'ms_per_step', 'grow_ms_per_pass', 'feature_prep_ms_per_pass', 'fill_ms_per_pass', 'grow_steps_per_sec',
                    'grow_steps_per_pass', 'longest_room_steps', 'config', 'per_rank')
            extras[key] = {k: ln[k] for k in keep if k in ln}
            rf = ln['roofline']
            extras[key]['roofline'] = {k: rf.get(k) for k in ('kernel', 'bound', 'achieved', 'peak', 'unit', 'frac', 'avg_launch_ms', 'sm_busy_frac', 'hbm_side',
                                                             'avg_us_per_item', 'speculative_steps_discarded_per_pass')}
    clocks = sampler.stop()          # nvidia-smi keeps sampling through the timed regions

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        all_host_threads()
        name, kind, total, res, _ = WORKLOADS[args.config]
        raw0 = local['rows'][0] if local['rows'] is not None else local['raw_points'][local['raw_off'][0]:local['raw_off'][1]]
        cap = None if len(raw0) <= 40000 else 400                  # (a 300k-point scene does not finish in seconds on the host)
        n_vec, t_vec, tp = cpu_room(rig.weights, raw0, res, max_steps=cap)
        n_lit, t_lit, tp_lit = cpu_room(rig.weights, raw0, res, literal=True, max_steps=60)
        whole = cap is None
        cpu_baseline = {'value': len(raw0) / t_vec if whole else (n_vec / t_vec) * line['config']['raw_points'] / max(line['grow_steps_per_pass'], 1),
                        'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
                        'sample': ('room 0 of the workload, whole (%d raw points, %d grow steps, %.1f s): ' % (len(raw0), n_vec, t_vec) if whole else
                                   'first %d grow steps of unit 0 (%d raw points); points/s = steps/s x raw points per grow step of the workload: ' % (n_vec, len(raw0)))
                                  + 'raw points -> host feature preparation -> grow -> fill by the oracle port (voxel-set update vectorised, numpy forward on all host threads)',
                        'grow_steps_per_sec': n_vec / max(t_vec - tp, 1e-9), 'feature_prep_s': tp,
                        'literal_update_loop': {'grow_steps_per_sec': n_lit / max(t_lit - tp_lit, 1e-9),
                                                'note': 'the reference\'s per-point Python loop (test_region_grow.py:282-287), first 60 grow steps'}}
    if rank == 0:
        line['clocks'] = clocks
        line['cpu_baseline'] = cpu_baseline
        line['extras'] = extras
        line['flops_per_grow_step'] = FLOPS_PER_STEP
        print(json.dumps(line))
    if rig.dist is not None:
        rig.dist.barrier()
        rig.dist.destroy_process_group()


if __name__ == '__main__':
    main()
