"""``bench.py --config N``: the other configurations of BASELINE.json (the default
run is config[1], the headline).  One JSON line per run, same keys as the
headline line where they apply; every GPU number has the oracle (CPU, one core,
bounded sample) timed beside it and its result compared with the oracle's.

  --config 0   example_script-style run: 64^3 particles, 1 NFW halo, periodic
               box, 32 snapshots, pericentres -- the drop-in track_orbits() on
               the GPU against the oracle's track_orbits() on the CPU, whole run
  --config 2   per-GPU shape of the 100k-halo configuration (--halos 100000 on
               256^3 particles per GPU): the default device-resident bench with
               many tiny blocks (handled by bench.run_b200 itself)
  --config 3   track_orbits_onthefly (apocentric): one call = one snapshot pair
  --config 4   consumers: progenitors.find_main_progenitors per snapshot and
               postprocessing.Apsides.collate_apsides over a tracked history
"""
import json
import os
import shutil
import tempfile
import time

import numpy as np


def _line(metric, unit, value, steps, ms, config, extra):
    line = {'metric': metric, 'value': value, 'unit': unit, 'n_gpus': 1,
            'steps': steps, 'warmup': 0, 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'data': 'synthetic', 'config': config}
    line.update(extra)
    if MISMATCHES:
        line['parity_detail'] = MISMATCHES[:8]
    print(json.dumps(line))
    return line


MISMATCHES = []          # first differences found, reported in the JSON line


def _trees_equal(got, exp):
    """Integer datasets identical, float datasets within the float16-angle
    tolerance of the GPU tests; differences are noted in MISMATCHES."""
    if sorted(got) != sorted(exp):
        MISMATCHES.append('keys: %s' % sorted(set(got) ^ set(exp))[:4])
        return False
    ok = True
    for k in sorted(exp):
        a, b = np.asarray(got[k]), np.asarray(exp[k])
        if a.shape != b.shape:
            MISMATCHES.append('%s: shape %s vs %s' % (k, a.shape, b.shape))
            ok = False
        elif a.dtype.kind == 'f' or b.dtype.kind == 'f':
            close = np.isclose(a.astype(np.float64), b.astype(np.float64),
                               rtol=2e-3, atol=2e-3, equal_nan=True)
            if not close.all():
                MISMATCHES.append('%s: %d of %d floats differ' % (
                    k, int((~close).sum()), close.size))
                ok = False
        elif not np.array_equal(a, b):
            MISMATCHES.append('%s: %d of %d integers differ' % (
                k, int((a != b).sum()), a.size))
            ok = False
    return ok


# ---------------------------------------------------------------------------
def config0(args):
    """BASELINE config[0]: the reference's own CPU-runnable case."""
    import torch
    from nbody_orbit_analysis_b200 import storage, track_orbits
    from nbody_orbit_analysis_b200.synth import SynthSim
    from oracle import orbit_oracle as oracle
    n_snap = 32
    sim = SynthSim(64 ** 3, 1, n_snap, dtype=np.float32,
                   catalogue_dtype=np.float64, nfw=True)
    cache = {}

    def loader(sn, pos, rad):             # generation is not part of the path
        if int(sn) not in cache:
            cache[int(sn)] = sim.load_snapshot_data(sn, pos, rad)
        return cache[int(sn)]
    for t, sn in enumerate(sim.snapshot_numbers):
        p, r, _ = sim.regions(sn, sim.main_branches[t])
        loader(sn, p, r)
    tmp = tempfile.mkdtemp(prefix='oa_cfg0_')
    f_gpu, f_cpu = os.path.join(tmp, 'gpu.h5'), os.path.join(tmp, 'cpu.h5')
    a = (sim.snapshot_numbers, sim.main_branches, sim.regions, loader)
    track_orbits.track_orbits(*a, f_gpu, verbose=False)         # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    track_orbits.track_orbits(*a, f_gpu, verbose=False)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    with np.errstate(all='ignore'):
        oracle.track_orbits(*a, f_cpu, storage=storage)
    t_cpu = time.perf_counter() - t0
    count = sum(len(cache[int(sn)]['ids']) for sn in sim.snapshot_numbers[1:])
    ok = _trees_equal(storage.tree(f_gpu), storage.tree(f_cpu))
    shutil.rmtree(tmp, ignore_errors=True)
    return _line(
        'particle-snapshots/sec', 'particle-snapshots/s', count / t_gpu,
        n_snap - 1, 1e3 * t_gpu / (n_snap - 1),
        {'workload': 'BASELINE config[0]: 64^3 particles, 1 NFW halo, periodic '
                     'box, 32 snapshots, pericentric, through track_orbits() '
                     '(host arrays in, result file out; wall clock)'},
        {'dtype': 'f32 frame / f64 v_r / int64 ids / f16 angles',
         'e2e': {'value': count / t_gpu, 'unit': 'particle-snapshots/s',
                 'h2d_bytes_per_step': int(count / (n_snap - 1) * 32),
                 'd2h_bytes_per_step': None},
         'cpu_baseline': {'value': count / t_cpu, 'unit': 'particle-snapshots/s',
                          'cores': 1, 'kind': 'port',
                          'sample': 'the whole run through oracle.track_orbits',
                          'parity_vs_gpu': 'ok' if ok else 'MISMATCH'},
         'parity': 'ok' if ok else 'MISMATCH'})


# ---------------------------------------------------------------------------
def config3(args):
    """track_orbits_onthefly, apocentric (BASELINE config[3]'s entry point, one
    GPU, 256^3-particle universe): one call per snapshot pair."""
    import torch
    from nbody_orbit_analysis_b200 import storage
    from nbody_orbit_analysis_b200 import track_orbits_onthefly as otf
    from nbody_orbit_analysis_b200.synth import SynthSim
    from oracle import orbit_oracle as oracle
    n_calls = min(args.steps, 6)
    halos = min(args.halos, 200)
    sim = SynthSim(args.particles // 8, halos, n_calls + 2, dtype=np.float32,
                   catalogue_dtype=np.float32)
    snaps = {}
    for t, sn in enumerate(sim.snapshot_numbers):
        p, r = sim.regions_onthefly(sn, sim.main_branches[t])[:2]
        snaps[int(sn)] = sim.load_snapshot_data(sn, p, r)

    def loader(sn, pos, rad):
        return snaps[int(sn)]
    links = [np.stack((sim.main_branches[t], sim.main_branches[t - 1]))
             for t in range(1, len(sim.snapshot_numbers))]
    tmp = tempfile.mkdtemp(prefix='oa_cfg3_')

    def run(fn, path, **kw):
        secs = []
        for t in range(1, len(sim.snapshot_numbers)):
            t0 = time.perf_counter()
            fn(sim.snapshot_numbers[t], links[t - 1], sim.regions_onthefly,
               loader, path, mode='apocentric', verbose=False, **kw)
            torch.cuda.synchronize()
            secs.append(time.perf_counter() - t0)
        return secs
    f_gpu, f_cpu = os.path.join(tmp, 'gpu_{}.h5'), os.path.join(tmp, 'cpu_{}.h5')
    s_gpu = run(otf.track_orbits, f_gpu)
    with np.errstate(all='ignore'):
        s_cpu = run(oracle.track_orbits_onthefly, f_cpu, storage=storage)
    count = [len(snaps[int(sn)]['ids']) for sn in sim.snapshot_numbers[1:]]
    ok = all(_trees_equal(storage.tree(f_gpu.format('%0.3d' % sn)),
                          storage.tree(f_cpu.format('%0.3d' % sn)))
             for sn in sim.snapshot_numbers[1:])
    shutil.rmtree(tmp, ignore_errors=True)
    # the first call allocates and pins: timed from the second
    v_gpu = sum(count[1:]) / sum(s_gpu[1:])
    v_cpu = sum(count[1:]) / sum(s_cpu[1:])
    return _line(
        'particle-snapshots/sec', 'particle-snapshots/s', v_gpu, len(count) - 1,
        1e3 * float(np.mean(s_gpu[1:])),
        {'workload': 'track_orbits_onthefly, apocentric: %d particles in %d '
                     'halo blocks per call, every call re-frames the previous '
                     'snapshot like the reference (track_orbits_onthefly.py:'
                     '71-205); host arrays in, result file out; wall clock'
                     % (count[-1], halos)},
        {'dtype': 'f32',
         'e2e': {'value': v_gpu, 'unit': 'particle-snapshots/s',
                 'h2d_bytes_per_step': int(np.mean(count)) * 64,
                 'd2h_bytes_per_step': None},
         'cpu_baseline': {'value': v_cpu, 'unit': 'particle-snapshots/s',
                          'cores': 1, 'kind': 'port',
                          'sample': 'the same calls through '
                                    'oracle.track_orbits_onthefly',
                          'parity_vs_gpu': 'ok' if ok else 'MISMATCH'},
         'parity': 'ok' if ok else 'MISMATCH'})


# ---------------------------------------------------------------------------
def config4(args):
    """The two consumers of the tracking path (BASELINE config[4])."""
    import torch
    from nbody_orbit_analysis_b200 import postprocessing, progenitors, storage
    from nbody_orbit_analysis_b200 import track_orbits
    from nbody_orbit_analysis_b200.synth import SynthSim
    from oracle import orbit_oracle as oracle
    rng = np.random.default_rng(7)
    # (a) find_main_progenitors: H halos of ~170 members, 100 tracked central
    # particles per descendant (progenitors.py:59-117)
    H = 100000
    lens = rng.integers(120, 220, H)
    halo_offsets = np.concatenate(([0], np.cumsum(lens)))[:-1]
    n_ids = int(lens.sum())
    halo_pids = rng.permutation(n_ids * 2)[:n_ids].astype(np.int64)
    tracked_offsets = np.arange(H, dtype=np.int64) * 100
    pick = (halo_offsets[:, None] + rng.integers(0, 120, (H, 100))).reshape(-1)
    tracked_pids = halo_pids[pick]
    tracked_pids[rng.random(len(tracked_pids)) < 0.1] = -5      # unbound since
    progenitors.find_main_progenitors(halo_pids, halo_offsets, tracked_pids,
                                      tracked_offsets)           # warm-up
    torch.cuda.synchronize()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        got = progenitors.find_main_progenitors(halo_pids, halo_offsets,
                                                tracked_pids, tracked_offsets)
    torch.cuda.synchronize()
    t_gpu = (time.perf_counter() - t0) / reps
    hs = 2000                                                   # CPU sample
    t0 = time.perf_counter()
    exp = oracle.find_main_progenitors(
        halo_pids[:halo_offsets[hs]], halo_offsets[:hs],
        tracked_pids[:100 * hs], tracked_offsets[:hs])
    t_cpu = time.perf_counter() - t0
    sample = progenitors.find_main_progenitors(
        halo_pids[:halo_offsets[hs]], halo_offsets[:hs],
        tracked_pids[:100 * hs], tracked_offsets[:hs])
    ok_a = [int(v) for v in sample] == [int(v) for v in exp] and len(got) == H
    prog = {'halo_particle_ids_per_s': (n_ids + 100 * H) / t_gpu,
            'ms_per_snapshot': 1e3 * t_gpu, 'halos': H, 'halo_ids': n_ids,
            'tracked_ids': 100 * H,
            'cpu_ids_per_s': (int(halo_offsets[hs]) + 100 * hs) / t_cpu,
            'cpu_sample_halos': hs, 'parity': 'ok' if ok_a else 'MISMATCH'}
    # (b) collate_apsides over a tracked history (postprocessing.py:30-174)
    n_snap = 24
    sim = SynthSim(2000000, 200, n_snap, dtype=np.float32,
                   catalogue_dtype=np.float32)
    tmp = tempfile.mkdtemp(prefix='oa_cfg4_')
    f_trk = os.path.join(tmp, 'track.h5')
    track_orbits.track_orbits(sim.snapshot_numbers, sim.main_branches,
                              sim.regions, sim.load_snapshot_data, f_trk,
                              verbose=False)
    tree = storage.tree(f_trk)
    n_events = sum(len(v) for k, v in tree.items() if k.endswith('er_IDs'))
    f_gpu, f_cpu = os.path.join(tmp, 'c_gpu.h5'), os.path.join(tmp, 'c_cpu.h5')
    t0 = time.perf_counter()
    postprocessing.Apsides(f_trk).collate_apsides(
        savefile=f_gpu, save_final_counts=True, verbose=False)
    torch.cuda.synchronize()
    t_gpu_b = time.perf_counter() - t0
    t0 = time.perf_counter()
    oracle.Apsides(f_trk, storage=storage).collate_apsides(
        savefile=f_cpu, save_final_counts=True, verbose=False)
    t_cpu_b = time.perf_counter() - t0
    ok_b = _trees_equal(storage.tree(f_gpu), storage.tree(f_cpu))
    shutil.rmtree(tmp, ignore_errors=True)
    post = {'events_per_s': n_events / t_gpu_b, 'seconds': t_gpu_b,
            'events': n_events, 'snapshots': n_snap - 1,
            'cpu_events_per_s': n_events / t_cpu_b, 'cpu_seconds': t_cpu_b,
            'parity': 'ok' if ok_b else 'MISMATCH',
            'what': 'collate_apsides + save_final_apsis_counts, file in / file '
                    'out (incremental merge of a sorted (halo, ID, count) table, '
                    'reference: np.unique over the whole history per snapshot)'}
    return _line(
        'halo-particle IDs matched/sec', 'ids/s', prog['halo_particle_ids_per_s'],
        reps, prog['ms_per_snapshot'],
        {'workload': 'BASELINE config[4]: progenitors.find_main_progenitors '
                     '(%d halos, %d member IDs, 100 tracked IDs per descendant) '
                     'and postprocessing.Apsides.collate_apsides (%d events of a '
                     '%d-snapshot history); host arrays / files in and out, wall '
                     'clock' % (H, n_ids, n_events, n_snap)},
        {'dtype': 'int64',
         'progenitors': prog, 'postprocessing': post,
         'cpu_baseline': {'value': prog['cpu_ids_per_s'], 'unit': 'ids/s',
                          'cores': 1, 'kind': 'port',
                          'sample': 'first %d halos through '
                                    'oracle.find_main_progenitors' % hs},
         'parity': 'ok' if ok_a and ok_b else 'MISMATCH'})


def run(args):
    fn = {0: config0, 3: config3, 4: config4}[args.config]
    return fn(args)
