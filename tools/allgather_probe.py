"""Where the label all-gather's wall time goes (torchrun --nproc-per-node 2 tools/allgather_probe.py)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from learn_region_grow_b200 import parallel
rank = int(os.environ.get('RANK', 0)); torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
dist.init_process_group('nccl')
n = 852443
src = torch.arange(n, dtype=torch.int32, device='cuda')
lengths = [n, n]
def t(fn, reps=5):
    out = []
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); out.append(1e3 * (time.perf_counter() - t0))
    return out
arr = parallel.DeviceArray(src.data_ptr(), n)
print(rank, 'as_tensor', ['%.2f' % x for x in t(lambda: torch.as_tensor(arr, device='cuda'))])
loc = torch.as_tensor(arr, device='cuda')
print(rank, 'allgather_labels', ['%.2f' % x for x in t(lambda: parallel.allgather_labels(loc, lengths))])
pad = torch.zeros(n, dtype=torch.int32, device='cuda'); out = torch.empty(2 * n, dtype=torch.int32, device='cuda')
print(rank, 'all_gather_into_tensor only', ['%.2f' % x for x in t(lambda: dist.all_gather_into_tensor(out, pad))])
dist.destroy_process_group()
