"""GPU parity of the partitioned hash join (``OrbitTracker(impl='pjoin')``,
csrc/oa_pjoin.cu) -- against the oracle through the drop-in ``track_orbits``
and, at sizes the oracle does not reach, against the hash-table kernel on the
same device (events, offsets AND float16 angles bit-identical: both paths do
the same arithmetic with the same CUDA ``acosf``).

Both generations of the partitioned join run here: ``impl='pjoin'``
(csrc/oa_pjoin.cu: COUNT / SCAN / SCATTER / JOIN stages) and ``impl='pj2'``
(csrc/oa_pj2.cu: fixed-capacity partitions, producer warp + TMA stages).
"""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]
IMPLS = ['pjoin', 'pj2']

CASES = [
    (60000, 37, 6, {}),
    (60000, 37, 6, {'late_halos': 0.3}),
    (40000, 300, 5, {'hubble': True}),
    (30000, 5, 5, {'catalogue_bulk': False}),
    (30000, 1, 5, {'nfw': True, 'periodic': False}),
    (5000, 2000, 4, {}),
]


@pytest.mark.parametrize('impl', IMPLS)
@pytest.mark.parametrize('target', [None, 300])
@pytest.mark.parametrize('mode', ['pericentric', 'apocentric'])
@pytest.mark.parametrize('case', CASES, ids=[
    'plain', 'late', 'hubble_300h', 'nobulk', 'nfw_nonperiodic', 'tiny_blocks'])
def test_track_orbits_pjoin_matches_oracle(case, mode, target, impl, tmp_path,
                                           monkeypatch):
    from test_gpu_track import compare_track_trees
    from nbody_orbit_analysis_b200 import pj2, pjoin, storage, track_orbits
    from nbody_orbit_analysis_b200.synth import SynthSim
    from oracle import orbit_oracle as oracle
    monkeypatch.setenv('OA_TRACK_IMPL', impl)
    if target is not None:       # small partitions: every stage at these sizes
        monkeypatch.setattr(pjoin, 'TARGET', target)
        monkeypatch.setattr(pjoin, 'LAG_PARTICLES', 1 << 12)
        # (pj2: partitions of ~40 records, many groups, regions that grow and
        # shrink across partition counts)
        monkeypatch.setattr(pj2, 'TARGET', 40)
        monkeypatch.setattr(pj2, 'GROUP_PARTICLES', 1 << 12)
    n, nh, ns, kw = case
    sim = SynthSim(n, nh, ns, dtype=np.float32, catalogue_dtype=np.float32, **kw)
    f_gpu, f_cpu = str(tmp_path / 'gpu.h5'), str(tmp_path / 'cpu.h5')
    args = (sim.snapshot_numbers, sim.main_branches, sim.regions,
            sim.load_snapshot_data)
    track_orbits.track_orbits(*args, f_gpu, mode=mode, verbose=False)
    oracle.track_orbits(*args, f_cpu, mode=mode, storage=storage)
    got, exp = storage.tree(f_gpu), storage.tree(f_cpu)
    assert sum(len(v) for k, v in exp.items() if k.endswith('er_IDs')) > 0
    compare_track_trees(got, exp, data_f64=False,
                        derived_bulk=not kw.get('catalogue_bulk', True))


@pytest.mark.parametrize('impl', IMPLS)
@pytest.mark.parametrize('n,halos', [(3000000, 40), (2000000, 2000)])
def test_pjoin_equals_hash_kernel_at_scale(n, halos, impl):
    """Default partition size, regions of up to ~10^5 particles, several
    groups of regions in flight: identical event lists from both kernels."""
    import torch
    from nbody_orbit_analysis_b200.synth import DeviceSynth
    from nbody_orbit_analysis_b200.tracker import OrbitTracker
    gen = DeviceSynth(n, halos)
    exists = np.arange(halos)
    trk = {i: OrbitTracker(impl=i) for i in ('hash', impl)}
    n_events = 0
    for t in range(5):
        dev, m, offsets = gen.snapshot(t)
        pos, rad, bulk = gen.regions(t)
        res = {i: tr.step_device(dev, m, np.float32, np.int64, offsets,
                                 exists, pos, bulk, 0.0,
                                 box_size=gen.host.box)
               for i, tr in trk.items()}
        if t == 0:
            continue
        a, b = res['hash'], res[impl]
        assert a.n_events == b.n_events > 0
        assert np.array_equal(a.apsis_offsets, b.apsis_offsets)
        assert np.array_equal(a.apsis_ids, b.apsis_ids)
        assert np.array_equal(a.apsis_angles.view(np.int16),
                              b.apsis_angles.view(np.int16))
        n_events += a.n_events
    torch.cuda.synchronize()
    assert n_events > 0


def test_full_size_properties_and_agreement():
    """BASELINE config[1] at FULL size (256^3 particles, 1000 halos: the bench
    workload), where the oracle does not reach: size-independent properties of
    the result, and agreement of the two independent implementations.

    * the event positions are strictly ascending over the previous snapshot
      (the reference's event order, ``track_orbits.py:315-316``);
    * every event ID is the ID of the previous-snapshot particle at that
      position (membership, ``:316``), offsets are non-decreasing and end at
      the event count (``:214-215``);
    * the hash kernel and the second partitioned join -- different kernels,
      different data layouts -- give identical IDs, offsets and float16 angles;
    * a second run gives the same bits."""
    import torch
    from nbody_orbit_analysis_b200.synth import DeviceSynth
    from nbody_orbit_analysis_b200.tracker import OrbitTracker
    n, halos = 256 ** 3, 1000
    gen = DeviceSynth(n, halos)
    exists = np.arange(halos)

    def run(impl):
        trk = OrbitTracker(impl=impl)
        out, prev_ids = [], None
        for t in range(4):
            dev, m, offsets = gen.snapshot(t)
            pos, rad, bulk = gen.regions(t)
            res = trk.step_device(dev, m, np.float32, np.int64, offsets, exists,
                                  pos, bulk, 0.0, box_size=gen.host.box)
            if t > 0:
                sel = res.apsis_prev_index
                assert res.n_events == sel.numel() > 100000
                assert bool(torch.all(sel[1:] > sel[:-1]))
                assert torch.equal(res.d_ids, prev_ids[sel])
                off = res.apsis_offsets
                assert off[0] == 0 and off[-1] == res.n_events
                assert (np.diff(off) >= 0).all() and len(off) == halos + 1
                out.append((res.apsis_ids.copy(), off.copy(),
                            res.apsis_angles.view(np.int16).copy()))
            prev_ids = dev['ids']
        torch.cuda.synchronize()
        return out
    a, b, a2 = run('hash'), run('pj2'), run('hash')
    for (ia, oa, ga), (ib, ob, gb), (ic, oc, gc) in zip(a, b, a2):
        assert np.array_equal(ia, ib) and np.array_equal(oa, ob)
        assert np.array_equal(ga, gb)
        assert np.array_equal(ia, ic) and np.array_equal(ga, gc)
