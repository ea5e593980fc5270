#!/usr/bin/env python
"""Benchmark of the orbit-tracking hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (B200)
    python bench.py --impl reference --steps K --warmup W    # CPU arm

metric  : particle-snapshots/sec = sum over timed snapshots of the particles in
          all region blocks / time of the tracking path (halo frame + ID match
          + apsis detection + angle update + ordered event compaction; N > 1:
          + exchange and merge of the event lists in the reference's order).
step    : one snapshot.  Workload at N=1: BASELINE config[1], 256^3 particles,
          1000 halos, pericentric; per-GPU work is fixed as N grows (weak
          scaling, particles sharded by ID).
value   : inputs resident in HBM (generated there by csrc/oa_synth.cu), results
          (ordered event lists) left in HBM; only the event count and the
          region offsets are read back.  No bulk host<->device copy.
value_results_to_host : the same pass with the event lists also copied to
          pinned host memory every snapshot (what `value` meant in round 1; at
          8 GPUs this is bound by the box's host links, profiles/r02_scaling.md).
e2e     : same steps through OrbitTracker.step() with HOST (pinned) snapshot
          arrays: H2D of ids/pos/vel and D2H of the events inside the timing.
roofline: fused tracking kernel, algorithmic bytes = 72 B per particle-snapshot
           / its CUDA-event duration, vs MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

B_ALG_F32 = 72.0   # 32 in + 18 state read + 4 match + 18 write
FALLBACK_HBM_GBS = 6650.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--particles', type=int, default=256 ** 3,
                    help='particles per GPU (universe size)')
    ap.add_argument('--halos', type=int, default=1000)
    ap.add_argument('--mode', default='pericentric')
    ap.add_argument('--depth', type=int, default=2,
                    help='snapshots submitted ahead of the one being collected')
    ap.add_argument('--profile', action='store_true',
                    help='cProfile of the timed host loop (to stderr)')
    ap.add_argument('--config', type=int, default=1, choices=[0, 1, 2, 3, 4],
                    help='index into BASELINE.json configs (1 = the headline, '
                         'default; 2 = its 100k-halo shape per GPU; 0, 3, 4: '
                         'tools/bench_configs.py)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-particles', type=int, default=400000,
                    help='particles per snapshot in the CPU-baseline sample')
    return ap.parse_args()


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
SAMPLER_SRC = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
period = float(sys.argv[2])
try:
    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons
except AttributeError:
    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons
smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
while True:
    sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
    print('%.6f,%d,%d,%.1f,%d' % (time.time(), sm, smax, pw, reasons(h)), flush=True)
    time.sleep(period)
"""


class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region, taken by a
    separate process through NVML (the same counters `nvidia-smi
    --query-gpu=clocks.sm,clocks_event_reasons.*` prints).  A polling
    `nvidia-smi -lms` re-initialises NVML for every sample, which stalls this
    process's CUDA calls for milliseconds -- measured 2.2 -> 12.6 ms per step --
    so the light-weight loop above is used instead."""
    REASONS = ((0x8, 'hw_slowdown'), (0x40, 'hw_thermal_slowdown'),
               (0x20, 'sw_thermal_slowdown'), (0x4, 'sw_power_cap'))

    def __init__(self, index=0, period=None):
        period = float(os.environ.get('OA_BENCH_CLOCK_PERIOD', '0.01')) \
            if period is None else period
        self.index, self.period = index, period
        self.rows = []
        self.proc = None

    def start(self):
        import tempfile
        try:
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = int(vis.split(',')[self.index]) if vis else self.index
        except (ValueError, IndexError):
            idx = self.index
        try:
            # output goes to a file: no reader thread competes for the GIL of
            # the (host-bound) benchmark loop
            self.out = tempfile.NamedTemporaryFile(
                'w+', prefix='oa_clocks_', suffix='.csv', delete=False)
            self.proc = subprocess.Popen(
                [sys.executable, '-c', SAMPLER_SRC, str(idx), str(self.period)],
                stdout=self.out, stderr=subprocess.DEVNULL)
            time.sleep(0.5)          # NVML initialisation happens before timing
        except OSError:
            self.proc = None

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        time.sleep(2 * self.period)
        self.proc.terminate()
        self.proc.wait()
        self.out.flush()
        self.out.seek(0)
        self.rows = [ln.strip() for ln in self.out.read().splitlines()]
        self.out.close()
        os.unlink(self.out.name)
        sm, smax, reasons = [], [], set()
        for line in list(self.rows):
            f = line.split(',')
            try:
                ts, c, cmax, mask = float(f[0]), float(f[1]), float(f[2]), int(f[4])
            except (ValueError, IndexError):
                continue
            if ts < t0 or ts > t1:
                continue
            sm.append(c)
            smax.append(cmax)
            for bit, name in self.REASONS:
                if mask & bit:
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peak():
    path = os.path.join(HERE, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured'
    except (OSError, KeyError, ValueError):
        return FALLBACK_HBM_GBS, 'fallback'


# ---------------------------------------------------------------------------
# CPU arm: the oracle (numpy restatement of the reference) on host cores
# ---------------------------------------------------------------------------
def cpu_track(snapshots, catalog, mode):
    """Run the oracle over a list of host snapshots; returns
    (seconds in the tracking path, particle-snapshots, per-step outputs)."""
    from oracle import orbit_oracle as oracle
    prev, outs = None, []
    seconds, count = 0.0, 0
    near_zero = []
    for s, (snap, (pos, rad, bulk)) in enumerate(zip(snapshots, catalog)):
        exists = np.arange(len(pos))
        t0 = time.perf_counter()
        state, out = oracle.track_snapshot(snap, exists, pos, bulk, 0.0, mode,
                                           prev)
        dt = time.perf_counter() - t0
        if s > 0:
            seconds += dt
            count += len(snap['ids'])
        outs.append(out)
        prev = state
        # particles whose |v_r| lies within float eps of zero (north_star:
        # "listed separately"): relative to their speed in the halo frame
        lens = np.diff(np.append(snap['region_offsets'], len(snap['ids'])))
        vrel = np.linalg.norm(
            snap['velocities'].astype(np.float64) - np.repeat(
                np.asarray(bulk, dtype=np.float64), lens, axis=0), axis=1)
        hit = np.flatnonzero(np.abs(state.radial_vels) <=
                             8 * np.finfo(np.float32).eps * vrel)
        near_zero += [(s, int(snap['ids'][i])) for i in hit]
    cpu_track.near_zero = near_zero
    return seconds, count, outs


def host_sample(args, n_steps, seed_rank=0):
    """Bounded sample of the workload generated on the host: the first halos
    of the configuration up to ~cpu_particles particles per snapshot."""
    from nbody_orbit_analysis_b200.synth import SynthSim
    sim = SynthSim(args.particles, args.halos, n_steps + 1, dtype=np.float32,
                   catalogue_dtype=np.float32)
    ncols = int(np.searchsorted(np.cumsum(sim.sizes), args.cpu_particles)) + 1
    ncols = min(ncols, sim.n_halos)
    cols = np.arange(ncols)
    snaps, cat = [], []
    for t in range(n_steps + 1):
        pos = sim.halo_centre(t)[cols].astype(np.float32)
        cat.append((pos, sim.radius[cols].astype(np.float32),
                    sim.vh[cols].astype(np.float32)))
        snaps.append(sim.load_snapshot_data(sim.snapshot_numbers[t], pos, None,
                                            cols=cols))
    return snaps, cat, ncols


def _reference_worker(job):
    """One host core: generate the snapshots of two halos of the configuration
    (untimed) and run the oracle over them; the last K steps are timed."""
    particles, halos, cols, warmup, steps, mode = job
    from nbody_orbit_analysis_b200.synth import SynthSim
    from oracle import orbit_oracle as oracle
    n_steps = warmup + steps
    sim = SynthSim(particles, halos, n_steps + 1, dtype=np.float32,
                   catalogue_dtype=np.float32)
    cols = np.array(sorted(cols))
    snaps, cat = [], []
    for t in range(n_steps + 1):
        pos = sim.halo_centre(t)[cols].astype(np.float32)
        cat.append((pos, sim.vh[cols].astype(np.float32)))
        snaps.append(sim.load_snapshot_data(sim.snapshot_numbers[t], pos, None,
                                            cols=cols))
    prev, secs, count = None, 0.0, 0
    with np.errstate(all='ignore'):
        for s, (snap, (pos, bulk)) in enumerate(zip(snaps, cat)):
            t0 = time.perf_counter()
            prev, _ = oracle.track_snapshot(snap, np.arange(len(cols)), pos,
                                            bulk, 0.0, mode, prev)
            dt = time.perf_counter() - t0
            if s > warmup:
                secs += dt
                count += len(snap['ids'])
    return secs, count, len(snaps[-1]['ids'])


def reference_available():
    """The unmodified reference package is importable (mounted tree in the build
    container, or the copy vendored into oracle/_ref by `make -C oracle ref`)."""
    try:
        from oracle import reference_harness
        return reference_harness.available()
    except Exception:
        return False


def reference_track(snaps, cats, main_branches, snapshot_numbers, mode,
                    first_timed):
    """Run the UNMODIFIED reference `track_orbits` (single process, npool=None:
    its pathos pool is a slow-down) over prebuilt host
    snapshots.  Returns (seconds, particle-snapshots) of the iterations
    `first_timed` .. end: from the moment the loader callback of iteration
    `first_timed` returns until the call returns (tracking + result assembly +
    file write through the h5py stand-in; the callbacks only look data up)."""
    import tempfile
    from oracle.reference_harness import load_reference
    ref = load_reference()
    index = {int(sn): t for t, sn in enumerate(snapshot_numbers)}
    marks = {}

    def regions(snapshot_number, halo_ids):
        return cats[index[int(snapshot_number)]]

    def loader(snapshot_number, positions, radii):
        t = index[int(snapshot_number)]
        marks[t] = time.perf_counter()
        return snaps[t]
    tmp = tempfile.mkdtemp(prefix='oa_refarm_')
    with np.errstate(all='ignore'):
        ref.track_orbits.track_orbits(
            snapshot_numbers, main_branches, regions, loader,
            os.path.join(tmp, 'ref.h5'), mode=mode, npool=None, verbose=False)
    end = time.perf_counter()
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    count = sum(len(snaps[t]['ids']) for t in range(first_timed, len(snaps)))
    return end - marks[first_timed], count


def _reference_worker_real(job):
    """One host core: two halos of the configuration through the real
    reference; generation untimed, the last K snapshots timed."""
    particles, halos, cols, warmup, steps, mode = job
    import warnings
    warnings.filterwarnings('ignore')
    from nbody_orbit_analysis_b200.synth import SynthSim
    n_steps = warmup + steps
    sim = SynthSim(particles, halos, n_steps + 1, dtype=np.float32,
                   catalogue_dtype=np.float32)
    cols = np.array(sorted(cols))
    snaps, cats = [], []
    for t in range(n_steps + 1):
        pos = sim.halo_centre(t)[cols].astype(np.float32)
        cats.append((pos, sim.radius[cols].astype(np.float32),
                     sim.vh[cols].astype(np.float32)))
        snaps.append(sim.load_snapshot_data(sim.snapshot_numbers[t], pos, None,
                                            cols=cols))
    secs, count = reference_track(snaps, cats, sim.main_branches[:, cols],
                                  sim.snapshot_numbers, mode, warmup + 1)
    return secs, count, len(snaps[-1]['ids'])


def run_reference(args):
    """--impl reference: the reference's CPU path on this host's cores.

    The unmodified reference package runs (vendored into oracle/_ref by `make
    -C oracle ref`, so that it travels to the GPU box; the oracle port only if
    that copy is missing, or with OA_REF_ARM=port).  The reference's own
    parallel axis is "one halo per worker"
    (track_orbits.py:189-194); its pathos pool re-pickles the whole snapshot
    per task and is a slow-down, so the fair CPU arm is
    one forked worker per core, each tracking one halo of the configuration
    through the same W+K snapshots.  value = particles of the timed steps of
    all workers / the slowest worker's time in them."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    cores = max(1, min(cores, (args.halos - 4) // 2))
    # halos 4 .. 4+2*cores-1 of the configuration, a large and a small one per
    # worker so that the workers carry about the same number of particles
    jobs = [(args.particles, args.halos, (4 + w, 4 + 2 * cores - 1 - w),
             args.warmup, args.steps, args.mode) for w in range(cores)]
    wall0 = time.perf_counter()
    real = reference_available() and os.environ.get('OA_REF_ARM') != 'port'
    with mp.get_context('fork').Pool(cores) as pool:
        out = pool.map(_reference_worker_real if real else _reference_worker,
                       jobs, chunksize=1)
    wall = time.perf_counter() - wall0
    secs = max(o[0] for o in out)
    count = sum(o[1] for o in out)
    value = count / secs
    sample = ('%s; halos 4..%d of %d (two per core, %d particles/snapshot in total) '
              'of the %d-particle configuration, %d timed snapshots; '
              'generation untimed, whole run %.0f s' % (
                  'UNMODIFIED reference orbitanalysis.track_orbits (npool=None) in '
                  'one forked worker per core, h5py replaced by the in-repo '
                  'stand-in' if real else 'oracle port (reference not importable)',
                  3 + 2 * cores, args.halos, sum(o[2] for o in out),
                  args.particles, args.steps, wall))
    line = {
        'impl': 'reference', 'metric': 'particle-snapshots/sec',
        'value': value, 'unit': 'particle-snapshots/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * secs / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32/f64',
        'data': 'synthetic',
        'config': workload_config(args),
        'cpu_baseline': {'value': value, 'unit': 'particle-snapshots/s',
                         'cores': cores, 'kind': 'reference' if real else 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'particle-snapshots/s',
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def workload_config(args):
    return {
        'workload': '%d particles/GPU (256^3 = BASELINE config[1]), %d halos, '
                    '%s tracking, fp32 data + int64 IDs, catalogue bulk '
                    'velocities, periodic box' % (
                        args.particles, args.halos, args.mode),
        'particles_per_gpu': args.particles, 'halos': args.halos,
        'mode': args.mode,
        'l2': 'inputs larger than L2 (>=500 MB per step, distinct every step)',
        'sharding': 'particle ID mod n_gpus',
        'results': 'GPU arm, `value`: the ordered (N > 1: merged) event lists '
                   'stay in HBM, the host reads the event count and '
                   'region_offsets; `value_results_to_host` and `e2e` also copy '
                   'the lists to pinned host memory every snapshot',
    }


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from nbody_orbit_analysis_b200.synth import DeviceSynth
    from nbody_orbit_analysis_b200.tracker import OrbitTracker
    from nbody_orbit_analysis_b200 import sharded

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        # a short collective timeout: a rank that dies must not leave the others
        # (and the box) hanging for the default ten minutes
        dist.init_process_group('nccl', device_id=torch.device('cuda', local),
                                timeout=datetime.timedelta(seconds=120))
    K, W = args.steps, args.warmup
    n_snap = W + K + 1

    gen = DeviceSynth(args.particles, args.halos, rank=rank, world=world)
    exists = np.arange(args.halos)
    snaps = [gen.snapshot(t) for t in range(n_snap)]
    cats = [gen.regions(t) for t in range(n_snap)]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    comm = sharded.Comm(world, rank) if world > 1 else None

    cat_ahead = {}
    # wall-clock of the host phases of the timed loop (per rank; rank 0 reports):
    # shows whether a step is bound by the GPU or by host code / collectives
    phases = {}

    def timed(name, fn, *a, **k):
        t0 = time.perf_counter()
        out = fn(*a, **k)
        phases[name] = phases.get(name, 0.0) + time.perf_counter() - t0
        return out

    CAT_BLOCK = int(os.environ.get('OA_CAT_BLOCK', '16'))
    cat_done = {}

    def cat_block(b):
        """Broadcast the catalogues of snapshots [b CAT_BLOCK, (b+1) CAT_BLOCK)
        as ONE catalogue (the rows of the snapshots back to back)."""
        ts = range(b * CAT_BLOCK, min((b + 1) * CAT_BLOCK, n_snap))
        return comm.start_broadcast(*[np.concatenate([cats[t][k] for t in ts])
                                      for k in range(3)])

    def catalogue(t):
        """This snapshot's catalogue; multi-GPU: broadcast from rank 0 (SURVEY
        8(e)), a block of snapshots per collective, one block ahead, so that it
        neither stalls the submission nor costs a collective per snapshot."""
        if comm is None:
            return cats[t]
        b = t // CAT_BLOCK
        if b not in cat_done:
            h = cat_ahead.pop(b, None) or cat_block(b)
            if (b + 1) * CAT_BLOCK < n_snap:
                cat_ahead[b + 1] = cat_block(b + 1)
            pos, rad, bulk = comm.finish_broadcast(h)
            cat_done.clear()
            cat_done[b] = (pos, rad, bulk)
        pos, rad, bulk = cat_done[b]
        lo = (t - b * CAT_BLOCK) * args.halos
        hi = lo + args.halos
        return pos[lo:hi], rad[lo:hi], bulk[lo:hi]

    def submit_step(trk, t, host=None):
        pos, rad, bulk = timed('catalogue', catalogue, t)
        if host is None:
            dev, n, offsets = snaps[t]
            pending = timed('submit', trk.submit_device,
                            dev, n, np.float32, np.int64, offsets, exists, pos,
                            bulk, 0.0, box_size=gen.host.box, gpos=dev.get('gpos'))
        else:
            pending = timed('submit', trk.submit, host[t], exists, pos, bulk, 0.0,
                            gpos=host[t].get('_gpos'))
        if comm is not None:
            # the exchange's pack kernels right behind the snapshot's selection
            timed('prepack', comm.prepack, trk, pending)
        return pending

    exchange = {'inflight': None}
    run_cfg = {'results': 'host'}

    def collect_step(trk, pending):
        """Results of one snapshot.  Multi-GPU: its event exchange is started
        here and finished one snapshot later (it overlaps the next kernels);
        returns (local result, finished global result or None)."""
        res = timed('collect', trk.collect, pending)
        done = []
        if comm is not None and os.environ.get('OA_BENCH_NO_EXCHANGE') == '1':
            # diagnostic: tracking only, no event exchange (local event counts)
            if res.apsis_offsets is not None:
                done.append(res)
            return res, done
        if comm is not None and res.apsis_offsets is not None:
            if comm.batch_size > 1:
                # OA_EXCHANGE_BATCH=K: K snapshots per exchange, finished one
                # batch behind
                timed('start_merge', comm.stage_merge, trk, res)
                while len(comm._launched) > 1:
                    done += timed('finish_merge', comm.finish_batch,
                                  comm._launched[0])
                return res, done
            # every rank hands its 1/world share of the merged lists to the
            # host (parallel write of the result datasets)
            h = timed('start_merge', comm.start_merge, trk, res,
                      to_host='slice' if run_cfg['results'] == 'host' else False)
            prev, exchange['inflight'] = exchange['inflight'], h
            if prev is not None:
                done.append(timed('finish_merge', comm.finish_merge, prev))
        elif comm is None:
            done.append(res)
        return res, done

    def flush_exchange():
        """Finish whatever exchange is still in flight; list of results."""
        if comm is not None and comm.batch_size > 1:
            comm.launch_batch(flush_exchange.trk)
            done = []
            while comm._launched:
                done += comm.finish_batch(comm._launched[0])
            return done
        h, exchange['inflight'] = exchange['inflight'], None
        return [comm.finish_merge(h)] if h is not None else []

    def timed_run(host=None, results='host'):
        """W+1 untimed snapshots, then K timed ones.  Up to `depth` snapshots
        are submitted ahead of the one whose results are being collected
        (software pipeline: host-side collection, H2D and D2H overlap the
        kernels of the following snapshots).

        results='hbm' : a snapshot's (merged, ordered) event lists stay in HBM;
                        the host reads the event count and `region_offsets`
                        (device-resident pass: no bulk host<->device copy in it);
        results='host': they are also copied to pinned host memory every
                        snapshot -- all of them at N = 1, every rank's 1/N slice
                        of the merged lists at N > 1."""
        run_cfg['results'] = results
        trk = OrbitTracker(mode=args.mode)
        flush_exchange.trk = trk
        trk.events_on_device = comm is not None or results == 'hbm'
        if comm is not None:       # room for the NCCL kernels of the exchange
            trk.sm_reserve = int(os.environ.get('OA_SM_RESERVE', '12'))
        # nvidia-smi is started before the warm-up: its NVML initialisation
        # stalls CUDA calls for tens of ms and must not fall in the timed region
        sampler = ClockSampler(local)
        per_rank_clocks = world > 1 and os.environ.get('OA_BENCH_RANK_CLOCKS') == '1'
        if rank == 0 or per_rank_clocks:
            sampler.start()
        from collections import deque
        depth = max(1, args.depth)     # snapshots in flight (tracker ring = 3)
        if comm is not None and comm.batch_size == 1:
            # the exchange of snapshot k is finished one snapshot later and may
            # have to be repeated from k's device buffers: keep them un-recycled
            # (a batch is exchanged from staging buffers of its own)
            depth = 1
        queue = deque()
        for t in range(0, W + 1):      # same pipelined pattern as the timed loop
            queue.append(submit_step(trk, t, host))
            if len(queue) > depth:
                collect_step(trk, queue.popleft())
        while queue:
            collect_step(trk, queue.popleft())
        flush_exchange()
        trk.timing = []
        launches0 = trk.launches
        phases.clear()
        # profiling builds of the partitioned join (-DOA_PJOIN_STATS=1) count
        # cycles per stage: cleared here, read after the timed steps
        import ctypes as _C
        from nbody_orbit_analysis_b200._lib import lib as _oalib
        pj_stats = (_C.c_uint64 * 16)()
        stats_fn = _oalib.oa_pj2_stats if os.environ.get(
            'OA_TRACK_IMPL') == 'pj2' else _oalib.oa_pjoin_stats
        stats_fn(pj_stats, 1)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), \
            torch.cuda.Event(enable_timing=True)
        wall0 = time.time()
        ev0.record()
        n_part = n_events = 0
        last = None

        def take(p):
            nonlocal n_part, n_events, last
            last, done = collect_step(trk, p)
            n_part += last.n
            n_events += sum(d.n_events for d in done)
        prof = None
        if args.profile and rank == 0:
            import cProfile
            prof = cProfile.Profile()
            prof.enable()
        for t in range(W + 1, W + K + 1):
            queue.append(submit_step(trk, t, host))
            if len(queue) > depth:
                take(queue.popleft())
        while queue:
            take(queue.popleft())
        n_events += sum(d.n_events for d in flush_exchange())
        if prof is not None:
            import pstats
            prof.disable()
            st_ = pstats.Stats(prof, stream=sys.stderr).sort_stats('tottime')
            st_.print_stats(45)
            st_.sort_stats('cumulative').print_stats(30)
        ev1.record()
        barrier()
        wall1 = time.time()
        ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop(wall0, wall1) if (rank == 0 or per_rank_clocks) else None
        stats_on = stats_fn(pj_stats, 0)
        kern_ms = [a.elapsed_time(b) for a, b, _ in trk.timing]
        kern_n = [n for _, _, n in trk.timing]
        stats = torch.tensor([ms, float(n_part), float(n_events)],
                             dtype=torch.float64, device='cuda')
        if world > 1:
            mx = stats.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stats.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, n_part = float(mx[0]), float(sm[1])
            n_events = float(stats[2])      # already global after the merge
        per_rank = None
        if world > 1:
            mine = {'rank': rank, 'kernel_ms': float(np.mean(kern_ms)) if kern_ms else None,
                    'step_ms': float(stats[0]), 'clocks': clocks}
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            per_rank = gathered
        return {'ms': ms, 'wall_ms': 1e3 * (wall1 - wall0), 'per_rank': per_rank,
                'host_phases_ms_per_step': {k: round(1e3 * v / K, 4)
                                            for k, v in phases.items()},
                'particles': n_part, 'events': n_events,
                'launches': trk.launches - launches0, 'clocks': clocks,
                'kern_ms': kern_ms, 'kern_n': kern_n, 'last': last,
                'pj_stats': list(pj_stats) if stats_on else None}

    # `value`: inputs AND results resident in HBM (the host reads counts and
    # offsets only).  OA_BENCH_TO_HOST=1 restores the round-1 definition (event
    # lists copied to the host inside the timed region); that pass is otherwise
    # run right after and reported beside `value` as `value_results_to_host`.
    results_in = 'host' if os.environ.get('OA_BENCH_TO_HOST') == '1' else 'hbm'
    dev_run = timed_run(results=results_in)
    value = dev_run['particles'] / (dev_run['ms'] * 1e-3)
    to_host_run = None
    to_host_error = None
    if results_in == 'hbm' and os.environ.get('OA_BENCH_TO_HOST') != '0' and \
            os.environ.get('OA_BENCH_NO_EXCHANGE') != '1':
        if world > 1:
            to_host_run = timed_run(results='host')
        else:
            try:         # (one rank: a secondary pass must not cost the line)
                to_host_run = timed_run(results='host')
            except Exception as exc:
                import traceback
                traceback.print_exc(file=sys.stderr)
                to_host_error = '%s: %s' % (type(exc).__name__, exc)

    # ---- end to end: host (pinned) snapshots ---------------------------------
    e2e = None
    run_e2e = not args.no_e2e
    if run_e2e:
        # every snapshot of the run sits in pinned host memory (distinct inputs
        # every step): make sure the node has room for all ranks' copies
        need = n_snap * args.particles * 40 * world
        try:
            import psutil
            if psutil.virtual_memory().available < 1.5 * need:
                run_e2e = False
                e2e_skipped = ('host memory: %d GB of pinned snapshots needed, '
                               '%d GB available' % (
                                   need >> 30,
                                   psutil.virtual_memory().available >> 30))
        except ImportError:
            pass
        if world > 1:       # all ranks take the same branch
            flag = torch.tensor([int(run_e2e)], device='cuda')
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if run_e2e and not int(flag.item()):
                run_e2e, e2e_skipped = False, 'host memory on another rank'
    if run_e2e:
        host = []
        for t in range(n_snap):
            dev, n, offsets = snaps[t]
            h = {}
            for k, name in (('ids', 'ids'), ('pos', 'coordinates'),
                            ('vel', 'velocities'), ('gpos', '_gpos')):
                if dev.get(k) is None:
                    continue
                buf = torch.empty(dev[k].shape, dtype=dev[k].dtype,
                                  pin_memory=True)
                buf.copy_(dev[k])
                arr = buf.numpy()
                h[name] = arr.reshape(-1, 3) if k in ('pos', 'vel') else arr
                h['_pin_' + k] = buf
            h['masses'] = 1.0
            h['region_offsets'] = offsets[:-1]
            h['box_size'] = gen.host.box
            h['redshift'] = 0.0
            host.append(h)
        torch.cuda.synchronize()
        e2e_run = timed_run(host, results='host')
        n_per_step = e2e_run['particles'] / K / world
        ev_per_step = e2e_run['events'] / K
        # ids + pos + vel (+ global block positions when sharded) of one rank
        h2d_rank = int(n_per_step * (40 if world > 1 else 32) + args.halos * 72)
        e2e = {'value': e2e_run['particles'] / (e2e_run['ms'] * 1e-3),
               'unit': 'particle-snapshots/s',
               # whole job, like `value`: all ranks' uploads; every rank reads back
               # its 1/N slice of the merged lists + the offsets
               'h2d_bytes_per_step': h2d_rank * world,
               'd2h_bytes_per_step': int(ev_per_step * 10
                                         + world * (args.halos * 8 + 8)),
               'h2d_bytes_per_step_per_gpu': h2d_rank,
               'ms_per_step': e2e_run['ms'] / K,
               'events_per_step': ev_per_step}
        # what the end-to-end number is bound by: the host->device link
        e2e['h2d_gb_per_s_per_gpu'] = h2d_rank / (
            e2e['ms_per_step'] * 1e-3) / 1e9
        # same data, two passes: the global event totals must be identical
        e2e['events_equal_device_run'] = bool(
            e2e_run['events'] == dev_run['events'])
    host_results = None
    if to_host_run is not None:
        ev_per_step = to_host_run['events'] / K
        host_results = {
            'value': to_host_run['particles'] / (to_host_run['ms'] * 1e-3),
            'unit': 'particle-snapshots/s',
            'ms_per_step': to_host_run['ms'] / K,
            'd2h_bytes_per_step_per_gpu': int(ev_per_step * 10 / world
                                              + args.halos * 8 + 8),
            'events_equal_device_run': bool(
                to_host_run['events'] == dev_run['events']),
            'kernel_ms': float(np.mean(to_host_run['kern_ms'])),
            'host_phases_ms_per_step': to_host_run['host_phases_ms_per_step'],
            'what': 'the same device-resident pass with the event lists also '
                    'copied to pinned host memory every snapshot (%s); the '
                    'round-1 definition of `value`' % (
                        'all of them' if world == 1 else
                        "every rank's 1/%d slice of the merged lists" % world)}

    # ---- end to end through the drop-in ENTRY POINT --------------------------------
    # track_orbits() itself: pageable numpy arrays from the loader callback (the
    # staging memcpy into pinned memory included), the pipelined driver, and the
    # result file written per snapshot.  Wall clock from the loader call of the
    # third snapshot (allocation / pinning warm-up before it) to the return.
    e2e_entry = None
    if world == 1 and not args.no_e2e:
        try:
            e2e_entry = entry_point_e2e(args, snaps, cats, gen, min(K, 12) + 2)
        except Exception as exc:      # a secondary figure must not cost the line
            import traceback
            traceback.print_exc(file=sys.stderr)
            e2e_entry = {'error': '%s: %s' % (type(exc).__name__, exc)}

    # ---- roofline of the fused kernel ------------------------------------------
    peak, peak_kind = measured_peak()
    k_ms = float(np.mean(dev_run['kern_ms']))
    k_n = float(np.mean(dev_run['kern_n']))
    achieved = k_n * B_ALG_F32 / (k_ms * 1e-3) / 1e9
    impl = os.environ.get('OA_TRACK_IMPL', 'hash')
    roofline = {
        'kernel': {'hash': 'oa_track_kernel<float,float,double> (fused frame + '
                           'hash match + apsis + angle update + table insert)',
                   'pjoin': 'oa_pjoin_kernel (frame + partition scatter + '
                            'shared-memory hash join + apsis + angle update)',
                   'pj2': 'oa_pj2_kernel (TMA-staged frame + fixed-capacity '
                          'partition scatter + shared-memory hash join + apsis + '
                          'angle update)'}[impl],
        'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
        'frac': achieved / peak, 'peak_kind': peak_kind, 'traffic': None,
        'algorithmic_bytes_per_particle': B_ALG_F32,
        'kernel_ms': k_ms, 'kernel_ms_min': float(np.min(dev_run['kern_ms'])),
        'kernel_ms_max': float(np.max(dev_run['kern_ms'])),
        'particles_per_launch': k_n,
        'kernel_particle_snapshots_per_s': k_n / (k_ms * 1e-3),
        'kernel_share_of_step': k_ms * K / dev_run['ms'],
    }
    traffic_file = os.path.join(HERE, 'profiles', 'traffic.json')
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                roofline['traffic'] = json.load(fh).get(
                    {'hash': 'oa_track_kernel', 'pjoin': 'oa_pjoin_kernel',
                     'pj2': 'oa_pj2_kernel'}[impl])
        except (OSError, ValueError):
            pass

    # ---- CPU baseline on a bounded sample + parity of that sample -------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args, snaps, cats, gen, torch)
    multi_parity = None
    if world > 1 and not args.no_cpu:
        multi_parity = multi_rank_parity(args, snaps, cats, gen, comm, world,
                                         rank, torch, dist)

    if rank == 0:
        line = {
            'metric': 'particle-snapshots/sec', 'value': value,
            'unit': 'particle-snapshots/s', 'n_gpus': world, 'steps': K,
            'warmup': W, 'ms_per_step': dev_run['ms'] / K,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 frame / f64 v_r / int64 ids / f16 angles',
            'data': 'synthetic', 'config': workload_config(args),
            'clocks': dev_run['clocks'], 'gpu_launches': dev_run['launches'],
            'events_per_step': dev_run['events'] / K,
            'track_impl': impl,
            'host_phases_ms_per_step': dev_run['host_phases_ms_per_step'],
            'roofline': roofline,
        }
        if dev_run.get('per_rank'):
            line['per_rank'] = dev_run['per_rank']
        if comm is not None and getattr(comm, 'n_prepacked', 0):
            line['exchange_prepacked'] = comm.n_prepacked
        if comm is not None and getattr(comm, 'phase_ms', None):
            n_x = max(comm.phase_ms.get('exchanges', 1), 1)
            line['exchange_phases_ms'] = {
                k: round(v / n_x, 4) for k, v in comm.phase_ms.items()
                if k != 'exchanges'}
        if dev_run.get('pj_stats') and impl == 'pj2':
            st = dev_run['pj_stats']
            tot = float(st[12]) or 1.0      # cycles of thread 0 of every CTA
            line['pj2_stage_profile'] = {
                'consumer_share': {'scatter': round(st[0] / tot, 4),
                                   'join': round(st[1] / tot, 4),
                                   'wait_full_stage': round(st[6] / tot, 4)},
                'producer_share': {'wait_dependency': round(st[4] / tot, 4),
                                   'wait_free_stage': round(st[5] / tot, 4)},
                'items': {'scatter': int(st[8]), 'join': int(st[9])},
                'cycles_per_item': {'scatter': round(st[0] / max(st[8], 1), 1),
                                    'join': round(st[1] / max(st[9], 1), 1)}}
        elif dev_run.get('pj_stats'):
            st = dev_run['pj_stats']
            names = ('join', 'scatter', 'scan', 'count')
            tot = float(st[12]) or 1.0
            line['pjoin_stage_profile'] = {
                'share_of_cta_cycles': {n: round(st[i] / tot, 4)
                                        for i, n in enumerate(names)},
                'waiting_share': round(st[4] / tot, 4),
                'items': {n: int(st[8 + i]) for i, n in enumerate(names)},
                'cycles_per_item': {n: round(st[i] / max(st[8 + i], 1), 1)
                                    for i, n in enumerate(names)}}
        if host_results is not None:
            line['value_results_to_host'] = host_results
        elif to_host_error is not None:
            line['value_results_to_host'] = {'error': to_host_error}
        if e2e is not None:
            line['e2e'] = e2e
        elif not args.no_e2e:
            line['e2e_skipped'] = e2e_skipped
        if e2e_entry is not None:
            line['e2e_entry_point'] = e2e_entry
        if cpu is not None:
            line['cpu_baseline'] = cpu
        if multi_parity is not None:
            line['parity_multi_gpu'] = multi_parity
        bad = (multi_parity is not None
               and multi_parity['parity_vs_oracle'] != 'ok') or (
            e2e is not None and not e2e['events_equal_device_run']) or (
            host_results is not None
            and not host_results['events_equal_device_run']) or (
            cpu is not None and cpu['parity_vs_gpu_on_sample'] != 'ok')
        line['parity'] = 'MISMATCH' if bad else 'ok'
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def multi_rank_parity(args, snaps, cats, gen, comm, world, rank, torch, dist):
    """N > 1: a sample of WHOLE halos (the last halos of the configuration, all
    ranks' shards of them put back together in the global block order) is
    tracked (i) by the oracle on rank 0's host and (ii) by all ranks through
    OrbitTracker + both exchange paths (all-gather to every rank, all-to-all
    slices); the merged global event lists must equal the oracle's.
    Reference order: track_orbits.py:199-217, 315-316."""
    from nbody_orbit_analysis_b200 import sharded
    from nbody_orbit_analysis_b200.tracker import OrbitTracker
    sizes = gen.host.sizes
    n_h = len(sizes)
    tail = np.cumsum(sizes[::-1]) * world
    ncols = int(min(np.searchsorted(tail, args.cpu_particles) + 1, n_h))
    h0 = n_h - ncols
    n_s = min(len(snaps), 5)
    full, cat = [], []
    for t in range(n_s):
        dev, n, offsets = snaps[t]
        lo = int(offsets[h0])
        mine = {'ids': dev['ids'][lo:n].cpu().numpy(),
                'pos': dev['pos'][3 * lo:3 * n].cpu().numpy().reshape(-1, 3),
                'vel': dev['vel'][3 * lo:3 * n].cpu().numpy().reshape(-1, 3),
                'key': dev['gpos'][lo:n].cpu().numpy(),
                'halo': np.repeat(np.arange(ncols), np.diff(offsets[h0:]))}
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        # global block order: by halo, then by the order key (comparable across
        # the ranks: position / shuffle key in the unsharded block)
        key = np.concatenate([p['key'] for p in parts])
        halo = np.concatenate([p['halo'] for p in parts])
        order = np.lexsort((key, halo))
        halo = halo[order]
        full.append({
            'ids': np.concatenate([p['ids'] for p in parts])[order],
            'coordinates': np.concatenate([p['pos'] for p in parts])[order],
            'velocities': np.concatenate([p['vel'] for p in parts])[order],
            'masses': 1.0,
            'region_offsets': np.searchsorted(halo, np.arange(ncols)).astype(np.int64),
            'box_size': gen.host.box, 'redshift': 0.0, 'H0': 0.0,
            'Omega_m': 0.3, 'Omega_L': 0.7})
        pos, rad, bulk = cats[t]
        cat.append((pos[h0:], rad[h0:], bulk[h0:]))
    outs = None
    if rank == 0:
        _, _, outs = cpu_track(full, cat, args.mode)
    exists = np.arange(ncols)
    report = {'halos': ncols, 'snapshots': n_s - 1,
              'particles_per_snapshot': int(len(full[-1]['ids']))}
    ok_all = True
    def check(t, ids, ang, offsets):
        exp = outs[t]
        a, b = ang.astype(np.float32), exp['apsis_angles'].astype(np.float32)
        return (np.array_equal(ids, exp['apsis_ids'])
                and np.array_equal(offsets, exp['apsis_offsets'])
                and a.shape == b.shape and bool(
                    np.allclose(a, b, rtol=2e-3, atol=2e-3, equal_nan=True)))

    def whole(res):
        """This rank's slice of a merged result -> the global lists (rank 0)."""
        if res.host_ready is not None:
            res.host_ready.synchronize()
        lo, hi = res.host_slice
        pieces = [None] * world
        dist.all_gather_object(pieces, (lo, np.array(res.apsis_ids[:hi - lo]),
                                        np.array(res.apsis_angles[:hi - lo])))
        pieces.sort(key=lambda q: q[0])
        return (np.concatenate([q[1] for q in pieces]),
                np.concatenate([q[2] for q in pieces]))

    for path in (True, 'slice', 'batch'):
        trk = OrbitTracker(mode=args.mode)
        trk.events_on_device = True
        c = sharded.Comm(world, rank)
        if path == 'batch':
            c.batch_size = 3          # 4 event-bearing snapshots: a full batch + a flushed one
        ok, n_ev, staged = True, 0, []
        for t in range(n_s):
            local, gpos = sharded.shard_snapshot(full[t], rank, world)
            res = trk.step(local, exists, cat[t][0], cat[t][2], 0.0, gpos=gpos)
            if t == 0:
                continue
            if path == 'batch':
                c.stage_merge(trk, res)
                staged.append(t)
                if t == n_s - 1:
                    c.launch_batch(trk)
                done = []
                while c._launched:
                    done += c.finish_batch(c._launched[0])
                for r in done:
                    tt = staged.pop(0)
                    ids, ang = whole(r)
                    if rank == 0:
                        n_ev += len(outs[tt]['apsis_ids'])
                        ok &= check(tt, ids, ang, r.apsis_offsets)
                continue
            res = c.merge_events(trk, res, to_host=path)
            if path == 'slice':
                ids, ang = whole(res)
            else:
                lo, hi = res.host_slice
                ids, ang = np.array(res.apsis_ids[:hi - lo]), \
                    np.array(res.apsis_angles[:hi - lo])
            if rank == 0:
                n_ev += len(outs[t]['apsis_ids'])
                ok &= check(t, ids, ang, res.apsis_offsets)
        ok &= not staged
        report[{True: 'all_gather', 'slice': 'all_to_all',
                'batch': 'batched_all_to_all'}[path]] = 'ok' if ok else 'MISMATCH'
        report['sample_events'] = n_ev
        ok_all &= ok
    report['parity_vs_oracle'] = 'ok' if ok_all else 'MISMATCH'
    return report


def entry_point_e2e(args, snaps, cats, gen, n_e):
    """Time ``track_orbits()`` (the call a user of the reference makes,
    ``track_orbits.py:9-11``) over ``n_e`` snapshots of this run's data."""
    import shutil
    import tempfile
    from nbody_orbit_analysis_b200 import track_orbits
    n_e = min(n_e, len(snaps))
    host = []
    for t in range(n_e):
        dev, n, offsets = snaps[t]
        host.append({
            'ids': dev['ids'].cpu().numpy(),                   # pageable
            'coordinates': dev['pos'].cpu().numpy().reshape(-1, 3),
            'velocities': dev['vel'].cpu().numpy().reshape(-1, 3),
            'masses': 1.0, 'region_offsets': offsets[:-1].copy(),
            'box_size': gen.host.box, 'redshift': 0.0, 'H0': 0.0,
            'Omega_m': 0.3, 'Omega_L': 0.7})
    marks = {}

    def regions(sn, halo_ids):
        return cats[int(sn)]

    def loader(sn, pos, rad):
        marks.setdefault(int(sn), time.perf_counter())
        return host[int(sn)]
    tmp = tempfile.mkdtemp(prefix='oa_entry_')
    out = os.path.join(tmp, 'entry.h5')
    mb = np.tile(np.arange(args.halos, dtype=np.int64), (n_e, 1))
    track_orbits.track_orbits(np.arange(n_e), mb, regions, loader, out,
                              mode=args.mode, verbose=False)
    t1 = time.perf_counter()
    first = 2
    secs = t1 - marks[first]
    count = sum(len(host[t]['ids']) for t in range(first, n_e))
    file_bytes = os.path.getsize(out)
    shutil.rmtree(tmp, ignore_errors=True)
    steps = n_e - first
    return {'value': count / secs, 'unit': 'particle-snapshots/s',
            'ms_per_step': 1e3 * secs / steps, 'snapshots': steps,
            'h2d_bytes_per_step': int(count / steps * 32 + args.halos * 72),
            'file_bytes_per_step': int(file_bytes / max(n_e - 1, 1)),
            'h2d_gb_per_s': count / steps * 32 / (secs / steps) / 1e9,
            'what': 'nbody_orbit_analysis_b200.track_orbits.track_orbits(...) '
                    'with loader callbacks returning PAGEABLE numpy arrays, '
                    'result groups written per snapshot (storage backend: '
                    'h5py if importable, else the in-repo container); wall '
                    'clock'}


def cpu_baseline(args, snaps, cats, gen, torch):
    """Oracle on the first halos of the first snapshots of THIS run's data,
    timed on one host core, and compared with the GPU path on the same data."""
    from nbody_orbit_analysis_b200.tracker import OrbitTracker
    sizes = gen.host.sizes
    ncols = int(np.searchsorted(np.cumsum(sizes), args.cpu_particles)) + 1
    ncols = min(ncols, len(sizes))
    n_s = min(len(snaps), 7)
    host, cat = [], []
    for t in range(n_s):
        dev, n, offsets = snaps[t]
        hi = int(offsets[ncols])
        host.append({
            'ids': dev['ids'][:hi].cpu().numpy(),
            'coordinates': dev['pos'][:3 * hi].cpu().numpy().reshape(-1, 3),
            'velocities': dev['vel'][:3 * hi].cpu().numpy().reshape(-1, 3),
            'masses': 1.0, 'region_offsets': offsets[:ncols].copy(),
            'box_size': gen.host.box, 'redshift': 0.0, 'H0': 0.0,
            'Omega_m': 0.3, 'Omega_L': 0.7})
        pos, rad, bulk = cats[t]
        cat.append((pos[:ncols], rad[:ncols], bulk[:ncols]))
    secs, count, outs = cpu_track(host, cat, args.mode)
    # same sample through the CUDA path
    trk = OrbitTracker(mode=args.mode)
    ok, n_ev = True, 0
    ulp_hist = np.zeros(4, dtype=np.int64)     # float16 event angles: 0, 1, 2, >2 ulp
    for t in range(n_s):
        res = trk.step(host[t], np.arange(ncols), cat[t][0], cat[t][2], 0.0)
        if t > 0:
            exp = outs[t]
            n_ev += len(exp['apsis_ids'])
            ga = np.asarray(res.apsis_angles, dtype=np.float16).view(np.int16)
            ea = np.asarray(exp['apsis_angles'], dtype=np.float16).view(np.int16)
            if ga.shape == ea.shape:
                du = np.abs(ga.astype(np.int64) - ea.astype(np.int64))
                ulp_hist += np.bincount(np.minimum(du, 3), minlength=4)[:4]
            ok &= np.array_equal(res.apsis_ids, exp['apsis_ids'])
            ok &= np.array_equal(res.apsis_offsets, exp['apsis_offsets'])
            a, b = res.apsis_angles.astype(np.float32), \
                exp['apsis_angles'].astype(np.float32)
            # (a mismatch must end up in the JSON line, not in a traceback)
            ok &= a.shape == b.shape and bool(
                np.allclose(a, b, rtol=2e-3, atol=2e-3, equal_nan=True))
    out = {
        'value': count / secs, 'unit': 'particle-snapshots/s', 'cores': 1,
        'kind': 'port',
        'sample': 'first %d halos (%d particles/snapshot) x %d snapshots of '
                  'this run, oracle (numpy port of the reference) on 1 core; '
                  'host has %d cores' % (ncols, len(host[-1]['ids']), n_s - 1,
                                         os.cpu_count()),
        'seconds': secs,
        'parity_vs_gpu_on_sample': 'ok' if ok else 'MISMATCH',
        'sample_events': n_ev,
        # float16 accumulated angles of the events, GPU vs oracle (arccos is not
        # correctly rounded in either libm; a 1-ulp difference of a float16
        # accumulator carries forward)
        'angle_ulp_histogram': {'0': int(ulp_hist[0]), '1': int(ulp_hist[1]),
                                '2': int(ulp_hist[2]), '>2': int(ulp_hist[3])},
        # |v_r| <= 8 float32 eps x |v - v_bulk|: the sign of such a v_r could
        # differ between two correct float evaluations; the CUDA path rounds
        # where numpy rounds, so these particles still agree with the reference
        'vr_within_eps_of_zero': {
            'count': len(cpu_track.near_zero),
            'first': [list(x) for x in cpu_track.near_zero[:8]],
            'what': '(snapshot of the sample, particle ID)'},
    }
    if reference_available():
        # the UNMODIFIED reference on the same sample, one core (npool=None is
        # its fastest setting); the port's figure stays as a second number
        mb = np.tile(np.arange(ncols, dtype=np.int64), (n_s, 1))
        r_secs, r_count = reference_track(host, cat, mb, np.arange(n_s), args.mode, 1)
        out.update({
            'value': r_count / r_secs, 'kind': 'reference', 'seconds': r_secs,
            'port_value': count / secs,
            'sample': 'first %d halos (%d particles/snapshot) x %d snapshots of '
                      'this run through the UNMODIFIED reference '
                      'orbitanalysis.track_orbits (npool=None, 1 core, h5py '
                      'replaced by the in-repo stand-in); parity of the sample '
                      'checked between the GPU path and the oracle port; host '
                      'has %d cores' % (ncols, len(host[-1]['ids']), n_s - 1,
                                        os.cpu_count())})
    return out


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    elif a.config in (0, 3, 4):
        sys.path.insert(0, os.path.join(HERE, 'tools'))
        import bench_configs
        bench_configs.run(a)
    else:
        if a.config == 2:
            a.halos = 100000
        run_b200(a)
