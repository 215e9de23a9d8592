"""CPU restatement of the region-grow driver -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows /root/reference/test_region_grow.py:175-316 statement by statement (voxel-*set* semantics for the
add/remove update, numpy median, numpy float32 arithmetic), parameterised by an RNG object so that it can
either replay the reference's global ``numpy.random`` stream bit for bit (``NumpyLegacyRng``) or consume the
counter-based Philox stream the CUDA engine uses (``PhiloxRng``).

Parity: PINNED against the unmodified reference driver executed under import shims
(oracle/make_golden.py -> tests/golden/driver_trace_*.npz; tests/test_oracle_driver.py).
"""
import numpy as np
import scipy.special

from . import philox


# ----------------------------------------------------------------------------- RNG objects
class NumpyLegacyRng:
    """Replays the reference's draws: one global MT19937 stream (test_region_grow.py:21)."""

    def __init__(self, seed=0, state=None):
        self.rs = state if state is not None else np.random.RandomState(seed)

    def begin_step(self, room, step, lane=0, seed_point=0):
        pass

    def sample(self, count, k, which):
        # test_region_grow.py:237-240 / :249-252
        if count >= k:
            return np.asarray(self.rs.choice(count, k, replace=False))
        return np.asarray(list(range(count)) + list(self.rs.choice(count, k - count, replace=True)))

    def uniform(self, k, which):
        return self.rs.random_sample(k)       # float64, compared with a float32 confidence (:266-267)


class PhiloxRng:
    """The engine's own stream: every draw is a pure function of (seed, room, step, stream, element)."""

    def __init__(self, seed=0):
        self.seed = int(seed)
        self.room = 0
        self.step = 0
        self.lane = 0
        self.seed_point = 0

    def _stream(self, stream):
        return (self.seed_point << 8) | (stream + 8 * self.lane)

    def begin_step(self, room, step, lane=0, seed_point=0):
        # Coordinates of a draw: (room, seed point of the region, step within the region, stream, element).  Keyed by the
        # region rather than by the room's running step count so that regions of a room can be grown out of order / side by
        # side; restart lane l (test_random_restart.py) draws from streams 8*l + {0..5}.
        self.room, self.step, self.lane, self.seed_point = int(room), int(step), int(lane), int(seed_point)

    def sample(self, count, k, which):
        key_stream = philox.STREAM_INLIER_KEY if which == 'inlier' else philox.STREAM_NEIGHBOR_KEY
        pad_stream = philox.STREAM_INLIER_PAD if which == 'inlier' else philox.STREAM_NEIGHBOR_PAD
        if count >= k:
            # k smallest (key, index) pairs, emitted in ascending index order
            keys = philox.draw_u32(self.seed, self.room, self.step, self._stream(key_stream), count)
            chosen = np.lexsort((np.arange(count), keys))[:k]
            return np.sort(chosen)
        r = philox.draw_u32(self.seed, self.room, self.step, self._stream(pad_stream), k - count).astype(np.uint64)
        pad = (r * np.uint64(count)) >> np.uint64(32)          # multiply-high range reduction
        return np.concatenate([np.arange(count), pad.astype(np.int64)])

    def uniform(self, k, which):
        stream = philox.STREAM_ADD_UNIFORM if which == 'add' else philox.STREAM_REMOVE_UNIFORM
        return philox.u32_to_unit_float(philox.draw_u32(self.seed, self.room, self.step, self._stream(stream), k))


# ----------------------------------------------------------------------------- confidence
def confidence(logits):
    """test_region_grow.py:262-263: softmax over the two logits, column 1 (float32 in, float32 out)."""
    return scipy.special.softmax(np.asarray(logits, np.float32), axis=-1)[:, 1]


def voxelize(xyz, resolution):
    """test_region_grow.py:175 -- float32 division by the python float, round-half-even, to int64."""
    return np.round(np.asarray(xyz, np.float32) / resolution).astype(int)


# ----------------------------------------------------------------------------- the state machine
class RoomGrower:
    """One room of test_region_grow.py:175-316.

    ``forward_fn(inlier (1,Ni,F) f32, neighbor (1,Nj,F) f32) -> (add (1,Nj,2), remove (1,Ni,2))``.
    """

    def __init__(self, points, order, forward_fn, rng, resolution=0.1, num_inlier=512, num_neighbor=512,
                 cluster_threshold=10, room_id=0, max_steps_per_region=None, literal_update=False):
        self.points = np.ascontiguousarray(points, dtype=np.float32)
        self.order = np.asarray(order)
        self.forward_fn = forward_fn
        self.rng = rng
        self.resolution = resolution
        self.Ni, self.Nj = num_inlier, num_neighbor
        self.cluster_threshold = cluster_threshold
        self.room_id = room_id
        self.max_steps_per_region = max_steps_per_region
        self.literal_update = literal_update      # True: the reference's per-point python loop (:282-287), for CPU timing
        n = len(self.points)
        self.point_voxels = voxelize(self.points[:, :3], resolution)              # :175
        self.cluster_label = np.zeros(n, dtype=int)                               # :176
        self.cluster_id = 1                                                       # :177
        self.visited = np.zeros(n, dtype=bool)                                    # :178
        self.total_steps = 0
        self.regions = []          # (seed, steps, size, reason, labelled)
        self.trace = None          # optional list of per-step dicts

    def _begin_rng_step(self):
        self.rng.begin_step(self.room_id, self.steps, 0, self.seed_id)

    # -- region lifecycle ---------------------------------------------------------------------------
    def begin_region(self, seed_id):
        self.seed_id = int(seed_id)
        self.currentMask = np.zeros(len(self.points), dtype=bool)                 # :198-199
        self.currentMask[seed_id] = True
        seed_voxel = self.point_voxels[seed_id]
        self.minDims = seed_voxel.copy()                                          # :200-203
        self.maxDims = seed_voxel.copy()
        self.seqMinDims = self.minDims
        self.seqMaxDims = self.maxDims
        self.steps = 0
        self.stuck = 0

    def stop_growing(self, reason):                                               # :210-217
        self.visited[self.currentMask] = True
        size = int(np.sum(self.currentMask))
        labelled = size > self.cluster_threshold
        if labelled:
            self.cluster_label[self.currentMask] = self.cluster_id
            self.cluster_id += 1
        self.regions.append((self.seed_id, self.steps, size, reason, labelled))

    def prepare_step(self):
        """:220-254.  Returns None after stopping with 'noneighbor', else the step's tiles."""
        points, pv = self.points, self.point_voxels
        currentPoints = points[self.currentMask, :].copy()
        newMinDims = self.minDims.copy() - 1
        newMaxDims = self.maxDims.copy() + 1
        mask = np.logical_and(np.all(pv >= newMinDims, axis=1), np.all(pv <= newMaxDims, axis=1))
        mask = np.logical_and(mask, np.logical_not(self.currentMask))
        mask = np.logical_and(mask, np.logical_not(self.visited))
        expandPoints = points[mask, :].copy()
        if len(expandPoints) == 0:                                                # :233-235
            self.stop_growing('noneighbor')
            return None
        self._begin_rng_step()
        subset_i = self.rng.sample(len(currentPoints), self.Ni, 'inlier')         # :237-240
        center = np.median(currentPoints, axis=0)                                 # :241
        expandPoints[:, :2] -= center[:2]                                         # :243-244
        expandPoints[:, 6:] -= center[6:]
        inlier = np.zeros((1, self.Ni, points.shape[1]), dtype=np.float32)
        inlier[0, :, :] = currentPoints[subset_i, :]                              # :245-247
        inlier[0, :, :2] -= center[:2]
        inlier[0, :, 6:] -= center[6:]
        subset_j = self.rng.sample(len(expandPoints), self.Nj, 'neighbor')        # :249-252
        neighbor = np.zeros((1, self.Nj, points.shape[1]), dtype=np.float32)
        neighbor[0, :, :] = expandPoints[subset_j, :]                             # :253
        self._step = dict(center=center, inlier=inlier, neighbor=neighbor,
                          inlier_idx=np.nonzero(self.currentMask)[0][subset_i],
                          neighbor_idx=np.nonzero(mask)[0][subset_j],
                          n_inlier=len(currentPoints), n_neighbor=len(expandPoints))
        return self._step

    def apply_step(self, add_logits, rmv_logits, add_mask=None, rmv_mask=None):
        """:262-306.  Returns the stop reason or None if the region keeps growing.
        ``add_mask`` / ``rmv_mask`` override the sampled masks (teacher forcing in parity tests); the uniforms
        are drawn either way so the stream stays aligned."""
        st = self._step
        add_conf = confidence(add_logits)
        rmv_conf = confidence(rmv_logits)
        u_add = self.rng.uniform(len(add_conf), 'add')                            # :266
        u_rmv = self.rng.uniform(len(rmv_conf), 'remove')                         # :267
        st.update(add_conf=add_conf, rmv_conf=rmv_conf, u_add=u_add, u_rmv=u_rmv)
        if add_mask is None:
            add_mask = u_add < add_conf
        if rmv_mask is None:
            rmv_mask = u_rmv < rmv_conf
        st.update(add_mask=np.asarray(add_mask, bool), rmv_mask=np.asarray(rmv_mask, bool))
        center = st['center']
        addPoints = st['neighbor'][0, :, :][st['add_mask']]                       # :270-273
        addPoints[:, :2] += center[:2]
        addVoxels = voxelize(addPoints[:, :3], self.resolution)
        rmvPoints = st['inlier'][0, :, :][st['rmv_mask']]                         # :274-277
        rmvPoints[:, :2] += center[:2]
        rmvVoxels = voxelize(rmvPoints[:, :3], self.resolution)
        if self.literal_update:
            addSet = set([tuple(p) for p in addVoxels])                           # :273,277
            rmvSet = set([tuple(p) for p in rmvVoxels])
            updated = False
            point_voxels, currentMask = self.point_voxels, self.currentMask
            for i in range(len(point_voxels)):                                    # :282-287, as written
                if not currentMask[i] and tuple(point_voxels[i]) in addSet:
                    currentMask[i] = True
                    updated = True
                if tuple(point_voxels[i]) in rmvSet:
                    currentMask[i] = False
        else:
            in_add = _rows_in(self.point_voxels, addVoxels)                       # same set membership, vectorised
            in_rmv = _rows_in(self.point_voxels, rmvVoxels)
            updated = bool(np.any(np.logical_and(~self.currentMask, in_add)))
            self.currentMask = np.logical_and(np.logical_or(self.currentMask, in_add), ~in_rmv)
        self.steps += 1                                                           # :288
        self.total_steps += 1
        if self.trace is not None:
            self.trace.append(dict(st, seed=self.seed_id, size_after=int(self.currentMask.sum())))
        if not updated:                                                           # :304-306
            self.stop_growing('noexpand')
            return 'noexpand'
        if not self.currentMask.any():
            # the reference would raise on min() of an empty array here; unreachable unless a re-rounded
            # voxel collides -- treated as a stop so the oracle stays total
            self.stop_growing('empty')
            return 'empty'
        self.minDims = self.point_voxels[self.currentMask, :].min(axis=0)         # :292-293
        self.maxDims = self.point_voxels[self.currentMask, :].max(axis=0)
        if not np.any(self.minDims < self.seqMinDims) and not np.any(self.maxDims > self.seqMaxDims):
            if self.stuck >= 1:                                                   # :295-299
                self.stop_growing('stuck')
                return 'stuck'
            self.stuck += 1
        else:
            self.stuck = 0                                                        # :300-301
        self.seqMinDims = np.minimum(self.seqMinDims, self.minDims)               # :302-303
        self.seqMaxDims = np.maximum(self.seqMaxDims, self.maxDims)
        if self.max_steps_per_region is not None and self.steps >= self.max_steps_per_region:
            self.stop_growing('maxsteps')      # engine safety cap; None (no cap) is the reference behaviour
            return 'maxsteps'
        return None

    # -- whole room ---------------------------------------------------------------------------------
    def grow_region(self, seed_id):
        self.begin_region(seed_id)
        while True:
            st = self.prepare_step()
            if st is None:
                return 'noneighbor'
            add, rmv = self.forward_fn(st['inlier'], st['neighbor'])
            reason = self.apply_step(np.asarray(add)[0], np.asarray(rmv)[0])
            if reason is not None:
                return reason

    def run(self):
        for seed_id in np.arange(len(self.points))[self.order]:                   # :183-188
            if self.visited[seed_id]:
                continue
            self.grow_region(seed_id)
        return self.cluster_label

    def fill(self):
        return fill_unlabeled(self.points, self.cluster_label)


class RestartRoomGrower(RoomGrower):
    """One room of /root/reference/test_random_restart.py:141-303: every seed is grown NUM_RESTARTS (:24) times from the
    same ``visited`` state and the restart that ends with the most points is kept (``restart_scoring = 'np'``, the default
    :40; numpy.argmax -> the first of equal scores, :177).  Only the kept mask becomes visited / labelled (:178-181).

    ('ml' scoring is not restated: after the first restart the reference resets ``maskLogProb`` to a list (:196), which
    numpy turns into an empty array on the next ``+=`` (:272), so every later score is empty -- the mode does not work.)

    RNG: with ``NumpyLegacyRng`` the restarts consume the single global stream one after the other like the reference; with
    ``PhiloxRng`` restart ``l`` is lane ``l`` with its own streams and its own step counter (steps taken by that lane in
    this room), so that the device can grow the restarts of a seed side by side.
    """

    def __init__(self, *args, num_restarts=10, **kw):
        super().__init__(*args, **kw)
        self.num_restarts = int(num_restarts)
        self.lane = 0
        self.lane_steps = [0] * self.num_restarts
        self.restart_score, self.restart_mask = [], []
        self.lane_regions = []      # (seed, lane, size, reason) of every restart

    def _begin_rng_step(self):
        self.rng.begin_step(self.room_id, self.steps, self.lane, self.seed_id)

    def stop_growing(self, reason):                                               # test_random_restart.py:170-197
        self.restart_score.append(int(np.sum(self.currentMask)))                  # :174
        self.restart_mask.append(self.currentMask)
        self.lane_regions.append((self.seed_id, self.lane, self.restart_score[-1], reason))
        if len(self.restart_score) == self.num_restarts:
            best = self.restart_mask[int(np.argmax(self.restart_score))]          # :177
            self.visited[best] = True
            size = int(np.sum(best))
            labelled = size > self.cluster_threshold
            if labelled:
                self.cluster_label[best] = self.cluster_id
                self.cluster_id += 1
            self.regions.append((self.seed_id, self.seed_steps, size, reason, labelled))
            self.restart_score, self.restart_mask = [], []

    def apply_step(self, *a, **kw):
        self.seed_steps += 1                                                      # `steps` is not reset by a restart (:163,:188-196)
        self.lane_steps[self.lane] += 1
        return super().apply_step(*a, **kw)

    def run(self):
        for seed_id in np.arange(len(self.points))[self.order]:
            if self.visited[seed_id]:
                continue
            self.seed_steps = 0
            for lane in range(self.num_restarts):
                self.lane = lane
                self.begin_region(seed_id)                                        # :188-196
                self.grow_region_from_current()
        return self.cluster_label

    def grow_region_from_current(self):
        while True:
            st = self.prepare_step()
            if st is None:
                return 'noneighbor'
            add, rmv = self.forward_fn(st['inlier'], st['neighbor'])
            reason = self.apply_step(np.asarray(add)[0], np.asarray(rmv)[0])
            if reason is not None:
                return reason


class BeamRoomGrower(RoomGrower):
    """One room of /root/reference/test_beam_search.py:142-285 (BEAM_WIDTH :24, SEARCH_WIDTH :25, ``scoring = 'np'`` :41).

    Per seed the reference keeps a queue ``Q`` of at most BEAM_WIDTH (score, mask) candidates.  Every round each candidate
    is expanded SEARCH_WIDTH times by one sampled grow step (:207-268); an expansion that added a point joins ``newQ`` with
    the score ``numpy.sum(newMask)`` (:266); the next ``Q`` is the BEAM_WIDTH best of ``newQ`` (stable sort, descending,
    :273).  At the head of every round (``qid == 0``, :179-190) ``bestMask = Q[0]`` and the stuck logic of the plain driver
    runs on its bounding box; the search ends when it sticks twice or no expansion added anything, and ``bestMask`` becomes
    visited / labelled (:276-279).

    ``scoring='ml'`` (:46-47): a candidate's score is its parent's score plus the log-probability of the step under the
    network's own confidences (:238-256,263-264) -- over ALL 512 (padded) tile rows of each set, ``log(conf)`` for a row whose
    re-rounded voxel is in the sampled add / remove set and ``log(1 - conf)`` otherwise, each divided by NUM_NEIGHBOR_POINT
    (both sets, :243,255) and accumulated row by row in float32 (numpy float32 scalars).

    The script as shipped needs Python 2 (``range(n) + list(...)``, :212,224); oracle/make_golden.py runs it unmodified with
    a list-returning ``range`` in its globals (oracle/run_reference.py).

    RNG: with ``NumpyLegacyRng`` the expansions consume the single global stream in the reference's order (candidate-major);
    with ``PhiloxRng`` expansion (q, s) of round r draws at (room, seed point, step r, lane q*SEARCH_WIDTH + s), so that the
    device can run the expansions of a round side by side.
    """

    def __init__(self, *args, beam_width=3, search_width=3, scoring='np', **kw):
        super().__init__(*args, **kw)
        assert scoring in ('np', 'ml')
        self.scoring = scoring
        self.step_log_prob = None   # 'ml': addLogProb + rmvLogProb of the last expansion (float32)
        self.B, self.W = int(beam_width), int(search_width)
        self.lane = 0
        self.round = 0
        self.lane_steps = [0] * (self.B * self.W)
        self.lane_log = []          # (seed, round, lane, updated, size) of every expansion

    def _begin_rng_step(self):
        self.rng.begin_step(self.room_id, self.round, self.lane, self.seed_id)

    def expand(self, mask, lane, forced=None):
        """One sampled grow step from ``mask`` (:191-268).  Returns None if the shell is empty, else (updated, newMask).
        ``forced(step) -> (add_logits, rmv_logits, add_mask, rmv_mask)`` lets a parity test supply the device's masks."""
        pv = self.point_voxels
        self.lane = lane
        self.currentMask = mask.copy()
        self.minDims = pv[mask, :].min(axis=0)                                    # :176-177
        self.maxDims = pv[mask, :].max(axis=0)
        st = self._prepare_tiles()
        if st is None:
            return None
        if forced is None:
            add, rmv = self.forward_fn(st['inlier'], st['neighbor'])
            add_mask = rmv_mask = None
        else:
            add, rmv, add_mask, rmv_mask = forced(st)
        updated = self._apply_masks(np.asarray(add)[0], np.asarray(rmv)[0], add_mask, rmv_mask)
        self.steps += 1                                                           # :261
        self.total_steps += 1
        self.lane_steps[lane] += 1
        self.lane_log.append((self.seed_id, self.round, lane, updated, int(self.currentMask.sum())))
        return updated, self.currentMask

    def _prepare_tiles(self):
        """prepare_step without the plain driver's stop bookkeeping (an empty shell just yields no candidate, :206)."""
        saved = self.stop_growing
        self.stop_growing = lambda reason: None
        try:
            return self.prepare_step()
        finally:
            self.stop_growing = saved

    def _apply_masks(self, add_logits, rmv_logits, add_mask, rmv_mask):
        """:230-260 -- sample the masks, re-voxelise the selected rows, set semantics of the update.  Returns ``updated``."""
        st = self._step
        add_conf = confidence(add_logits)
        rmv_conf = confidence(rmv_logits)
        u_add = self.rng.uniform(len(add_conf), 'add')                            # :232
        u_rmv = self.rng.uniform(len(rmv_conf), 'remove')                         # :233
        if add_mask is None:
            add_mask = u_add < add_conf
        if rmv_mask is None:
            rmv_mask = u_rmv < rmv_conf
        st.update(add_conf=add_conf, rmv_conf=rmv_conf, u_add=u_add, u_rmv=u_rmv,
                  add_mask=np.asarray(add_mask, bool), rmv_mask=np.asarray(rmv_mask, bool))
        center = st['center']
        addPoints = st['neighbor'][0, :, :][st['add_mask']]                       # :234-237
        addPoints[:, :2] += center[:2]
        addVoxels = voxelize(addPoints[:, :3], self.resolution)
        rmvPoints = st['inlier'][0, :, :][st['rmv_mask']]                         # :245-248
        rmvPoints[:, :2] += center[:2]
        rmvVoxels = voxelize(rmvPoints[:, :3], self.resolution)
        if self.scoring == 'ml':                                                  # :238-256
            parts = []
            for tile, conf, sel in ((st['neighbor'][0], add_conf, addVoxels), (st['inlier'][0], rmv_conf, rmvVoxels)):
                rows = np.array(tile[:, :3])
                rows[:, :2] += center[:2]                                         # :240,252 (the tile row is un-centred in place)
                hit = _rows_in(voxelize(rows, self.resolution), sel)
                with np.errstate(divide='ignore'):
                    terms = np.log(np.where(hit, conf, np.float32(1) - conf)) / np.float32(self.Nj)
                acc = np.float32(0)
                for t in terms:                                                   # row-by-row float32 accumulation (:243-245)
                    acc = acc + t
                parts.append(acc)
            self.step_log_parts = parts
        in_add = _rows_in(self.point_voxels, addVoxels)                           # :251-258
        in_rmv = _rows_in(self.point_voxels, rmvVoxels)
        updated = bool(np.any(np.logical_and(~self.currentMask, in_add)))
        self.currentMask = np.logical_and(np.logical_or(self.currentMask, in_add), ~in_rmv)
        if self.trace is not None:
            self.trace.append(dict(st, seed=self.seed_id, size_after=int(self.currentMask.sum())))
        return updated

    def grow_seed(self, seed_id, forced=None):
        """:153-279 for one seed.  Returns the stop reason ('stuck' or 'exhausted')."""
        pv = self.point_voxels
        self.seed_id = int(seed_id)
        seedMask = np.zeros(len(self.points), dtype=bool)                         # :154-155
        seedMask[seed_id] = True
        seqMin = pv[seed_id].copy()                                               # :156-159
        seqMax = pv[seed_id].copy()
        self.steps = 0
        stuck = 0
        bestMask = seedMask
        Q = [(0, seedMask)]                                                       # :164
        self.round = 0
        reason = 'exhausted'
        while len(Q) > 0:                                                         # :169
            bestMask = Q[0][1]                                                    # :179
            mn, mx = pv[bestMask, :].min(axis=0), pv[bestMask, :].max(axis=0)
            if not np.any(mn < seqMin) and not np.any(mx > seqMax):               # :180-186
                if stuck >= 1:
                    reason = 'stuck'
                    break
                stuck += 1
            else:
                stuck = 0
            seqMin = np.minimum(seqMin, mn)                                       # :187-188
            seqMax = np.maximum(seqMax, mx)
            newQ = []
            for q, (score, mask) in enumerate(Q):
                for s in range(self.W):                                           # :207
                    lane = q * self.W + s
                    r = self.expand(mask, lane, None if forced is None else (lambda st, lane=lane: forced(st, lane)))
                    if r is None:
                        break                                                     # empty shell: no expansion of this candidate (:206)
                    updated, newMask = r
                    if updated:                                                   # :262-267 ('np': the score is the region size)
                        if self.scoring == 'ml':                                  # :263-264
                            new_score = score + self.step_log_parts[0] + self.step_log_parts[1]
                            if forced is not None and getattr(forced, 'score', None) is not None:
                                new_score = forced.score(lane, score, self.step_log_parts, new_score)
                        else:
                            new_score = int(np.sum(newMask))
                        newQ.append((new_score, newMask))
            Q = sorted(newQ, key=lambda x: x[0], reverse=True)[:self.B]           # :273
            self.round += 1
        self.visited[bestMask] = True                                             # :276
        size = int(np.sum(bestMask))
        labelled = size > self.cluster_threshold
        if labelled:                                                              # :277-279
            self.cluster_label[bestMask] = self.cluster_id
            self.cluster_id += 1
        self.regions.append((self.seed_id, self.steps, size, reason, labelled))
        return reason

    def run(self, forced=None):
        for seed_id in np.arange(len(self.points))[self.order]:                   # :143-145
            if self.visited[seed_id]:
                continue
            self.grow_seed(seed_id, forced)
        return self.cluster_label


class _LaneGrower(RoomGrower):
    """One speculative lane of SpeculativeRoomGrower: a region grown on the shared (committed) ``visited`` array; stopping
    only records the reason -- the controller commits in seed order."""

    def stop_growing(self, reason):
        self.finished = reason


class SpeculativeRoomGrower:
    """DESIGN STUDY for intra-room parallelism that keeps the plain driver's result (test_region_grow.py:183-306) -- not a
    reference restatement.  ``lanes`` regions of one room grow side by side (one grow step per lane and tick), seeds handed
    out in curvature order, regions COMMITTED strictly in that order.  A region's trajectory depends on the regions before it
    only through ``visited`` inside the boxes it looked at (bounding box +- 1 voxel, :222-229); so when a commit makes points
    visited that lie inside the envelope a younger lane has looked at so far, that lane starts over (or is dropped if its
    seed itself was swallowed: in seed order it would never have been a seed); points outside the envelope are seen as
    visited from then on, exactly as in the sequential run.  Draws are keyed by (room, seed point, step in region) (PhiloxRng),
    so a region does not care when or on which lane it runs.  ``run()`` returns the labels, which equal RoomGrower's
    (tests/test_oracle_driver.py); ``ticks`` is the makespan in grow steps, ``wasted`` the steps of discarded attempts."""

    def __init__(self, points, order, forward_fn, seed=0, lanes=2, room_id=0, validate='early', **kw):
        """``validate='early'``: a commit restarts the younger lanes it invalidates at once (needs the committer to read the
        other lanes' envelopes).  ``validate='commit'``: nobody looks at another lane -- a region is checked only when it
        reaches the head of the window, against the points committed since it started, and grown again there if they touch
        its envelope (simpler on a device: the head lane reads immutable committed lists; conflicts cost more)."""
        self.points, self.order, self.forward_fn = points, np.asarray(order), forward_fn
        self.seed, self.L, self.room_id, self.kw = seed, int(lanes), room_id, kw
        self.validate = validate
        self.ticks = self.wasted = self.useful = self.restarts = self.dropped = 0

    def _lane(self, seed_id, visited, template):
        g = _LaneGrower(self.points, self.order, self.forward_fn, PhiloxRng(self.seed), room_id=self.room_id, **self.kw) \
            if template is None else template
        g.visited = visited
        g.finished = None
        g.begin_region(seed_id)
        g.envelope = (g.point_voxels[seed_id] - 1, g.point_voxels[seed_id] + 1)
        g.seen_commits = len(getattr(self, 'commit_log', []))        # commits that were already visible when the region began
        return g

    def run(self):
        n = len(self.points)
        visited = np.zeros(n, dtype=bool)
        label = np.zeros(n, dtype=int)
        cluster_id, cursor = 1, 0
        window = []                      # lanes in seed order; window[0] commits next
        threshold = self.kw.get('cluster_threshold', 10)
        self.regions = []
        self.commit_log = []             # voxels of every committed region, in commit order
        while True:
            while len(window) < self.L:                                        # hand out the next unvisited seeds (:183-188)
                issued = {g.seed_id for g in window}
                while cursor < n and (visited[self.order[cursor]] or self.order[cursor] in issued):
                    cursor += 1
                if cursor >= n:
                    break
                window.append(self._lane(int(self.order[cursor]), visited, None))
                cursor += 1
            if not window:
                break
            stepped = False
            for g in window:                                                   # one grow step per running lane
                if g.finished is not None:
                    continue
                st = g.prepare_step()
                if st is not None:
                    add, rmv = self.forward_fn(st['inlier'], st['neighbor'])
                    g.apply_step(np.asarray(add)[0], np.asarray(rmv)[0])
                    g.envelope = (np.minimum(g.envelope[0], g.seqMinDims - 1), np.maximum(g.envelope[1], g.seqMaxDims + 1))
                    stepped = True
            self.ticks += 1 if stepped else 0
            while window and window[0].finished is not None:                   # commit in seed order (:210-217)
                g = window[0]
                if self.validate == 'commit':
                    if visited[g.seed_id]:                                     # swallowed by a region committed meanwhile
                        window.pop(0)
                        self.wasted += g.steps
                        self.dropped += 1
                        continue
                    hit = any(np.any(np.all((v >= g.envelope[0]) & (v <= g.envelope[1]), axis=1)) for v in self.commit_log[g.seen_commits:])
                    if hit:                                                    # grown on a stale visited set: again, now at the head
                        self.wasted += g.steps
                        self.restarts += 1
                        self._lane(g.seed_id, visited, g)
                        break
                window.pop(0)
                mask = g.currentMask
                visited[mask] = True
                size = int(mask.sum())
                if size > threshold:
                    label[mask] = cluster_id
                    cluster_id += 1
                self.regions.append((g.seed_id, g.steps, size, g.finished, size > threshold))
                self.useful += g.steps
                vox = g.point_voxels[mask]
                self.commit_log.append(vox)
                if self.validate == 'commit':
                    continue
                for k, y in enumerate(list(window)):                           # younger lanes that looked at these points
                    if np.any(np.all((vox >= y.envelope[0]) & (vox <= y.envelope[1]), axis=1)):
                        self.wasted += y.steps
                        if visited[y.seed_id]:
                            window.remove(y)
                            self.dropped += 1
                        else:
                            self._lane(y.seed_id, visited, y)
                            self.restarts += 1
        self.cluster_label = label
        return label


def _rows_in(voxels, query):
    """Row-wise membership of (N,3) int voxels in a set of (M,3) voxels (python ``tuple in set`` at :283-286)."""
    if len(query) == 0:
        return np.zeros(len(voxels), dtype=bool)
    both = np.concatenate([voxels, query]).astype(np.int64)
    lo = both.min(axis=0)
    span = both.max(axis=0) - lo + 1
    def key(v):
        v = v.astype(np.int64) - lo
        return (v[:, 0] * span[1] + v[:, 1]) * span[2] + v[:, 2]
    return np.isin(key(voxels), key(query))


def fill_unlabeled(points, cluster_label):
    """test_region_grow.py:308-316: unlabeled points take the label of the nearest labelled point (13-D, fp32)."""
    nonzero_idx = np.nonzero(cluster_label)[0]
    filled = cluster_label.copy()
    if len(nonzero_idx) == 0:
        return filled           # the reference would raise in argmin([]); nothing to copy from
    nonzero_points = points[nonzero_idx, :]
    for i in np.nonzero(cluster_label == 0)[0]:
        d = np.sum((nonzero_points - points[i]) ** 2, axis=1)
        filled[i] = cluster_label[nonzero_idx[np.argmin(d)]]
    return filled
