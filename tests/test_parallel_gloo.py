"""World-size-2 gloo test of the only collective on the path: the final all-gather of instance labels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from learn_region_grow_b200 import parallel


def test_shard_rooms_balances_and_partitions():
    counts = [20000, 5000, 18000, 7000, 12000, 3000, 9000]
    shards = parallel.shard_rooms(counts, 3)
    allr = np.sort(np.concatenate(shards))
    assert allr.tolist() == list(range(7))
    loads = [sum(counts[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(counts)
    assert [len(s) for s in parallel.shard_rooms([5, 4], 4)] == [1, 1, 0, 0]


def _worker(rank, world, port, counts, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    shards = parallel.shard_rooms(counts, world)
    # fake "segmentation": label = 1000*room + position, so misplaced data is detectable
    local = [np.arange(counts[r], dtype=np.int32) + 1000 * r for r in shards[rank]]
    local = torch.from_numpy(np.concatenate(local) if local else np.zeros(0, np.int32))
    lengths = [int(sum(counts[r] for r in s)) for s in shards]
    gathered = parallel.allgather_labels(local, lengths)
    rooms = parallel.scatter_back(gathered, shards, counts)
    ok = all(np.array_equal(rooms[r].numpy(), np.arange(counts[r], dtype=np.int32) + 1000 * r) for r in range(len(counts)))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_labels_world2():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    counts = [7, 3, 11, 0, 5]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
