"""Per-GPU device->host bandwidth with all GPUs of the box copying at the same
time (one process per GPU, no NCCL): is the 14.6 MB event slice every rank hands
to its host per snapshot link-bound at 8 GPUs?   torchrun --nproc-per-node N tools/d2h_probe.py"""
import os
import time

import torch

rank = int(os.environ.get('LOCAL_RANK', '0'))
world = int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(rank)
nbytes = 14_600_000
dev = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
big = torch.empty(434_000_000, dtype=torch.uint8, pin_memory=True)
dbig = torch.empty(434_000_000, dtype=torch.uint8, device='cuda')
s = torch.cuda.Stream()
# crude start alignment: all ranks wait for the same wall-clock second
t_go = (int(time.time()) // 5 + 2) * 5
for name, src, dst, reps in (('d2h_14.6MB', dev, host, 300), ('h2d_434MB', big, dbig, 12),
                             ('d2h_14.6MB_again', dev, host, 300)):
    while time.time() < t_go:
        pass
    t_go += 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        dst.copy_(src, non_blocking=True)
        e0.record(s)
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record(s)
    s.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print('rank %d of %d  %-18s %.3f ms  %.1f GB/s' % (rank, world, name, ms, src.numel() / ms / 1e6),
          flush=True)
