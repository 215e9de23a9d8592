"""Room sharding over the GPUs of one box and the single collective of the path.

Rooms are independent units (the ``for room_id`` loop of test_region_grow.py:110 carries no state across
rooms), so each rank grows its own rooms with no communication; instance labels are exchanged ONCE, after the last
room, with one all-gather of int32 labels padded to the longest shard.
"""
import numpy as np


def shard_rooms(counts, world_size):
    """Longest-processing-time assignment of rooms to ranks by point count.  Returns a list of index arrays."""
    counts = np.asarray(counts, dtype=np.int64)
    order = np.argsort(-counts, kind='stable')
    load = np.zeros(world_size, dtype=np.int64)
    shards = [[] for _ in range(world_size)]
    for r in order:
        k = int(np.argmin(load))
        shards[k].append(int(r))
        load[k] += counts[r]
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def allgather_labels(local_labels, shard_lengths, group=None):
    """local_labels: 1-D int32 torch tensor (this rank's concatenated room labels, device of the backend);
    shard_lengths: total label count of every rank (known to all ranks from the room table).
    Returns the list of every rank's labels (views into one gathered buffer)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    longest = int(max(shard_lengths)) if len(shard_lengths) else 0
    padded = torch.zeros(longest, dtype=torch.int32, device=local_labels.device)
    padded[:local_labels.numel()] = local_labels
    out = torch.empty(world * longest, dtype=torch.int32, device=local_labels.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return [out[r * longest:r * longest + int(shard_lengths[r])] for r in range(world)]


def scatter_back(gathered, shards, counts):
    """Per-room label arrays in global room order from the gathered per-rank label vectors."""
    rooms = [None] * len(counts)
    for r, idx in enumerate(shards):
        pos = 0
        for room in idx:
            rooms[room] = gathered[r][pos:pos + counts[room]]
            pos += counts[room]
    return rooms


class DeviceArray:
    """Wraps a raw device pointer for ``torch.as_tensor`` through ``__cuda_array_interface__`` (zero copy)."""

    def __init__(self, ptr, n, typestr='<i4'):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': typestr, 'data': (int(ptr), False), 'version': 2}
