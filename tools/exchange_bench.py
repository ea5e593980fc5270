#!/usr/bin/env python
"""The multi-GPU event exchange alone (no tracking kernel): every rank packs a
synthetic event list of the bench's size, `Comm.start_merge` / `finish_merge`
run in the bench's pipeline pattern, and rank 0 prints the time per exchange
(device time by CUDA events, host time per phase).  Answers "is it the
exchange?" for the 8-GPU runs of profiles/r01_scaling.md.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29512 tools/exchange_bench.py [--events 1460000] [--halos 1000] [--steps 30]
  (--backend gloo: CPU run with the numpy stand-ins of the kernels, for testing this script)
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np                      # noqa: E402
import torch                            # noqa: E402
import torch.distributed as dist        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--events', type=int, default=1460000, help='events per rank and step')
    ap.add_argument('--halos', type=int, default=1000)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--backend', default='nccl', choices=['nccl', 'gloo'])
    ap.add_argument('--mode', default='slice', choices=['slice', 'all'])
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    from nbody_orbit_analysis_b200 import sharded
    import exchange_emul as emul
    if a.backend == 'nccl':
        import datetime
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local),
                                timeout=datetime.timedelta(seconds=120))
        from nbody_orbit_analysis_b200.tracker import OrbitTracker
        trk = OrbitTracker()
        device = torch.device('cuda', local)
    else:
        dist.init_process_group('gloo')
        emul.install(sharded)
        trk = emul.EmulTracker()
        device = torch.device('cpu')
    comm = sharded.Comm(world, rank, device=device)

    # a rank's events: a sorted uniform sample of the positions of an unsharded
    # previous snapshot of `n_prev` particles, the share with id mod world == rank
    n_prev_local = 8 * a.events
    rng = np.random.default_rng(1234 + rank)
    gpos = np.sort(rng.choice(world * n_prev_local, n_prev_local, replace=False)).astype(np.int64)
    starts = np.linspace(0, world * n_prev_local, a.halos, endpoint=False).astype(np.int64)
    seg_begin = np.searchsorted(gpos, starts)
    results = []
    for k in range(4):                  # a few distinct lists, reused round-robin
        m = int(a.events * (0.9 + 0.05 * k))
        sel = np.sort(rng.choice(n_prev_local, m, replace=False)).astype(np.int64)
        res = emul.EmulResult(0, gpos, sel, rng.integers(0, 2 ** 40, m),
                              rng.standard_normal(m).astype(np.float16), seg_begin)
        for name in ('d_sel', 'd_ids_buf', 'd_ang_buf', 'd_small'):
            setattr(res, name, getattr(res, name).to(device))
        res.prev_gen.gpos = res.prev_gen.gpos.to(device)
        if a.backend == 'nccl':
            res.compacted = torch.cuda.Event()
            res.compacted.record()
        results.append(res)

    def sync():
        if a.backend == 'nccl':
            torch.cuda.synchronize()
        dist.barrier()

    to_host = 'slice' if a.mode == 'slice' else True
    phases = {'start': 0.0, 'finish': 0.0}
    pending = None
    total_events = 0
    for step in range(-3, a.steps):
        if step == 0:
            sync()
            phases = {'start': 0.0, 'finish': 0.0}
            t_wall = time.perf_counter()
        res = results[step % len(results)]
        res.step = step
        trk._step = step + 2            # as in the bench: snapshot step+1 is submitted
        t0 = time.perf_counter()
        h = comm.start_merge(trk, res, to_host=to_host)
        t1 = time.perf_counter()
        if pending is not None:
            done = comm.finish_merge(pending)
            total_events = done.n_events
        t2 = time.perf_counter()
        phases['start'] += t1 - t0
        phases['finish'] += t2 - t1
        pending = h
    comm.finish_merge(pending)
    sync()
    wall = time.perf_counter() - t_wall
    if rank == 0:
        print(json.dumps({
            'n_ranks': world, 'backend': a.backend, 'mode': a.mode,
            'events_per_rank': a.events, 'global_events': int(total_events),
            'ms_per_exchange_wall': 1e3 * wall / a.steps,
            'host_ms_per_exchange': {k: 1e3 * v / a.steps for k, v in phases.items()},
            'capacity_records_per_rank': comm._cap, 'host_cores': os.cpu_count()}))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
