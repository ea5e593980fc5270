"""Device-side state machine of the orbit-tracking path.

``OrbitTracker`` is the GPU replacement of one iteration of the reference's
per-snapshot loop body (reference ``track_orbits.py:147-217``: the per-halo
``track(j)`` calls plus result assembly).  ``submit`` stages one snapshot in HBM
and enqueues, without any host synchronisation, the fused tracking kernel, the
ordered event compaction and the device->host copy of the small result arrays;
``collect`` waits for that snapshot's results.  Calling ``submit`` for snapshot
s+1 before ``collect`` for snapshot s overlaps the host side (user callbacks,
file writing, D2H) with GPU work.  ``step`` = ``submit`` + ``collect``.

PyTorch is used for device/pinned buffers, streams and events only; every
kernel is launched through the C ABI (``_lib``).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, pjoin, pj2
from ._lib import lib, check, ptr

_F = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}
_ITEMSIZE = {torch.uint8: 1, torch.int8: 1, torch.int16: 2, torch.float16: 2,
             torch.int32: 4, torch.float32: 4, torch.int64: 8, torch.float64: 8}


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.OrbitB200Error(
            "orbit-b200 needs a CUDA device (B200, sm_100a); there is no CPU "
            "fallback.")


class Generation:
    """Carried state of one processed snapshot (reference
    ``track_orbits.py:234-240``), resident in HBM."""
    __slots__ = ('n', 'rec', 'tab', 'mark', 'index_bits', 'offsets',
                 'halo_exists', 'frame_f64', 'ids_dtype', 'gpos', 'buckets',
                 'n_buckets',
                 'halo_ids64',
                 # partitioned-join generations (impl='pjoin'): records are
                 # stored partitioned, events take their IDs from `ids`
                 'pjoin', 'ids', 'part_off', 'pj_bits', 'pj_pb',
                 # second-generation partitioned join (impl='pj2'): fixed-
                 # capacity partitions, fill word per partition
                 'pj2', 'fill', 'pj2_plan')


class StepResult:
    """Results of one snapshot.  ``host_ready`` (multi-GPU merge only) is the
    event after which ``apsis_ids`` / ``apsis_angles`` are valid on the host."""
    __slots__ = ('host_ready', 'host_slice', 'step', 'n', 'apsis_ids', 'apsis_angles', 'apsis_offsets', 'hinds',
                 'bulk_velocities', 'angles', 'diag', 'n_events',
                 'apsis_prev_index', 'prev_gen', 'd_ids', 'd_ang', 'compacted',
                 'd_sel', 'd_ids_buf', 'd_ang_buf', 'd_small', 'prepack')


class Pending:
    """A submitted snapshot whose results have not been collected yet."""
    pass


def _as_f(arr, name):
    arr = np.asarray(arr)
    if arr.dtype not in _F:
        if arr.dtype.kind in 'iuf':
            arr = arr.astype(np.float64)
        else:
            raise TypeError("%s must be a float array, got %s" % (
                name, arr.dtype))
    return arr


STAGE_MIN_BYTES = 8 << 20     # smaller staging copies stay on the calling thread
_stage_threads = None


def stage_threads():
    """Threads of a staging copy (``oa_host_copy``): the cores this process may
    run on, shared between the ranks of the node, at most 8 (a copy is bound by
    memory bandwidth before that, and every thread is started per call);
    ``OA_STAGE_THREADS`` overrides."""
    global _stage_threads
    if _stage_threads is None:
        import os
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        try:
            local = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1')))
        except ValueError:
            local = 1
        n = max(1, min(8, cores // local))
        try:
            n = max(1, int(os.environ.get('OA_STAGE_THREADS', n)))
        except ValueError:
            pass
        _stage_threads = n
    return _stage_threads


def _stage_copy(dst, src):
    """Pageable host tensor -> pinned staging tensor (same dtype and length,
    both contiguous).  Large copies are split over a few threads by the C
    library (non-temporal memcpy, the GIL released for its duration): one
    core's copy is several times slower than the host-to-device link."""
    nbytes = src.numel() * src.element_size()
    if nbytes < STAGE_MIN_BYTES or dst.dtype != src.dtype or \
            dst.numel() != src.numel():
        dst.copy_(src)
        return
    check(lib.oa_host_copy(dst.data_ptr(), src.data_ptr(), nbytes,
                           stage_threads()))


class OrbitTracker:
    """Per-snapshot GPU tracker.

    Parameters
    ----------
    mode : 'pericentric' | 'apocentric'
    device : torch device (default: current CUDA device)
    onthefly : use the arithmetic of ``track_orbits_onthefly.py`` (frame and
        v_r in the data dtype, no Hubble flow) and produce per-match outputs.
    """

    def __init__(self, mode='pericentric', device=None, onthefly=False,
                 impl=None, ring=None):
        require_cuda()
        # ring: slots every device buffer cycles through.  3 (default) lets two
        # snapshots be in flight (submit s+1 before collecting s); 2 is the
        # minimum (previous + current generation) and is the memory plan for
        # snapshots that fill HBM: ~79 B per region-particle and slot (32 inputs +
        # 32 record + ~13 table + 2 mark), i.e. 158 B/particle instead of 237
        import os
        if ring is None and os.environ.get('OA_TRACKER_RING'):
            ring = int(os.environ['OA_TRACKER_RING'])
        if ring is not None:
            if int(ring) < 2:
                raise ValueError("ring must be at least 2")
            self.RING = int(ring)
        # 'hash' : oa_track_fused (global hash table, every option)
        # 'pjoin': oa_pjoin_step (partitioned shared-memory join; float32 data
        #          and catalogue, plain tracking without per-particle outputs)
        import os
        impl = impl or os.environ.get('OA_TRACK_IMPL', 'hash')
        if impl not in ('hash', 'pjoin', 'pj2'):
            raise ValueError("impl must be 'hash', 'pjoin' or 'pj2'")
        if impl != 'hash' and onthefly:
            raise ValueError("impl=%r does not cover the on-the-fly path" % impl)
        self.impl = impl
        self._planner = None
        if mode not in _lib.OA_MODE:
            raise ValueError("mode must be 'pericentric' or 'apocentric'")
        self.mode = mode
        self.device = torch.device(
            device if device is not None else
            'cuda:%d' % torch.cuda.current_device())
        self.onthefly = bool(onthefly)
        self.prev = None
        self.launches = 0          # kernels launched through the C ABI
        self.timing = None         # list of (start, stop, n) CUDA events of
        #                            the fused kernel when profiling is on
        self.copy_stream = torch.cuda.Stream(self.device)   # device -> host
        self.h2d_stream = torch.cuda.Stream(self.device)    # host -> device
        self._pool = {}
        self._hpool = {}
        self._step = 0             # submitted snapshots
        self._consumed = {}        # ring slot -> event: inputs read by the kernel
        self._last_events = 0      # event count of the last collected snapshot
        # multi-GPU: the event lists stay in HBM (StepResult.d_ids / d_ang) for
        # the NCCL exchange instead of being copied to the host per rank
        self.events_on_device = False
        self.wait_before_submit = None   # event the next submit must wait for
        self.sm_reserve = 0        # SMs the fused kernel leaves to other streams
        self._cur_main = None
        self._copy_st = C.c_void_p(self.copy_stream.cuda_stream)
        self._bulk_stream = None

    # -- buffers -------------------------------------------------------------
    RING = 3      # generations / in-flight steps a buffer name cycles through

    def _empty(self, n, dtype):
        """Uninitialised device buffer from torch's caching allocator (used
        for buffers handed to the caller)."""
        return torch.empty(max(int(n), 1), dtype=dtype, device=self.device)[:int(n)]

    def _buf(self, name, n, dtype, slot=None):
        """Persistent device buffer ``name`` of the ring slot of the current
        step.  The tracker owns its working set in HBM: a buffer is allocated
        once, grows geometrically when a later snapshot needs more (particle
        counts drift by a few per cent between snapshots) and is otherwise
        reused every ``RING`` steps.  Reuse is safe without events because every
        kernel that touches these buffers is ordered on the main stream; the
        copy stream only reads per-step outputs, which ``collect`` has
        synchronised on long before their slot comes round again."""
        slot = self._step % self.RING if slot is None else slot
        item = _ITEMSIZE[dtype]
        nbytes = max(int(n), 1) * item
        key = (name, slot)
        raw = self._pool.get(key)
        if raw is None or raw.numel() < nbytes:
            # (re)allocate the whole ring of this name at once, so that every
            # allocation happens in the first step that needs it
            cap = -(-int(nbytes * 1.25) // 512) * 512
            for k in range(self.RING):
                old = self._pool.get((name, k))
                if old is None or old.numel() < nbytes:
                    self._pool[(name, k)] = torch.empty(
                        cap, dtype=torch.uint8, device=self.device)
            raw = self._pool[key]
        return raw[:int(n) * item].view(dtype)

    def _main(self):
        # (inside submit_device the current stream is looked up once: the
        # torch.cuda.current_stream call costs ~10 us, a submit used it 8 times)
        cur = self._cur_main
        return cur if cur is not None else torch.cuda.current_stream(self.device)

    def _stream(self):
        return C.c_void_p(self._main().cuda_stream)

    def _to_device(self, arr, dtype=None, name=None):
        """Host numpy array -> device tensor.

        With ``name`` the destination is the tracker's ring buffer of that name
        and the copy runs on the host->device stream, so that the upload of
        snapshot s+1 overlaps the kernels of snapshot s (double-buffered
        ingest); the main stream waits for it in ``_inputs_ready``.  Pageable
        arrays are staged through pinned memory first."""
        arr = np.ascontiguousarray(arr, dtype=dtype)
        t = torch.from_numpy(arr.reshape(-1))
        if t.numel() == 0:
            return torch.empty(0, dtype=t.dtype, device=self.device)
        if not t.is_pinned():
            # pageable memory: stage through pinned memory (the tracker's own
            # ring when the destination is one of its buffers)
            if name is None:
                buf = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
            else:
                buf = self._hbuf('stage_' + name, t.numel(), t.dtype)
            _stage_copy(buf, t)
            t = buf
        if name is None:
            return t.to(self.device, non_blocking=True)
        dst = self._buf(name, t.numel(), t.dtype)
        with torch.cuda.stream(self.h2d_stream):
            dst.copy_(t, non_blocking=True)
        return dst

    HOST_RING = 4   # result buffers stay valid until two further submits
    SMALL_COPY = 256 << 10   # read-backs up to this size are written by a kernel

    def _hbuf(self, name, n, dtype, step=None, reserve=0):
        """Persistent PINNED host buffer ``name`` of a ring slot (cudaHostAlloc
        costs ~0.5 ms per call, so nothing is pinned in steady state).  Result
        arrays handed to the caller are views into these buffers: they stay
        valid until ``HOST_RING - 2`` further snapshots have been submitted."""
        step = self._step if step is None else step
        item = _ITEMSIZE[dtype]
        nbytes = max(int(n), 1) * item
        key = (name, step % self.HOST_RING)
        raw = self._hpool.get(key)
        if raw is None or raw.numel() < nbytes:
            cap = -(-int(max(nbytes, reserve * item) * 1.25) // 4096) * 4096
            for k in range(self.HOST_RING):     # whole ring at once (see _buf)
                old = self._hpool.get((name, k))
                if old is None or old.numel() < nbytes:
                    self._hpool[(name, k)] = torch.empty(
                        cap, dtype=torch.uint8, pin_memory=True)
            raw = self._hpool[key]
        return raw[:int(n) * item].view(dtype)

    def _to_host_async(self, t, n=None, name=None, step=None, reserve=0,
                       via=None):
        """Device tensor -> pinned host tensor on the copy stream (the caller
        synchronises before reading).  With ``name`` the destination is the
        tracker's pinned ring buffer of that name."""
        n = t.numel() if n is None else int(n)
        if name is None:
            h = torch.empty(n, dtype=t.dtype, pin_memory=True)
        else:
            h = self._hbuf(name, n, t.dtype, step, reserve)
        if n:
            # (through the C ABI on the copy stream: no stream switch of the host
            # framework per copy.  Small read-backs -- what collect() and the
            # exchange wait for -- are written by a kernel: a DMA copy would
            # queue behind the event lists in flight to the host)
            nbytes = n * _ITEMSIZE[t.dtype]
            small = nbytes <= self.SMALL_COPY and nbytes % 4 == 0 and \
                t.data_ptr() % 4 == 0 and via is None
            fn = lib.oa_copy_small if small else lib.oa_copy_async
            self.launches += int(small)       # (a kernel of this library)
            check(fn(h.data_ptr(), t.data_ptr(), nbytes,
                     self._copy_st if via is None else via))
        return h

    def to_host(self, *tensors, stream=None):
        """Device tensors produced on ``stream`` (default: the main stream) ->
        pinned host tensors (one synchronisation for all of them)."""
        done = torch.cuda.Event()
        done.record(stream if stream is not None else self._main())
        self.copy_stream.wait_event(done)
        out = [self._to_host_async(t) for t in tensors]
        self.copy_stream.synchronize()
        return out

    def to_host_async(self, *tensors, stream=None, names=None, reserve=0,
                      step=None, bulk=False):
        """Like ``to_host`` without the synchronisation: returns the pinned
        tensors and the event that marks their completion.  With ``names`` the
        destinations are the tracker's pinned ring buffers of those names, slot
        ``step`` mod HOST_RING (default: the number of submitted snapshots).
        ``bulk``: a large copy (the merged event lists of the multi-GPU
        exchange) goes on a stream of its own, so that the few-kilobyte
        read-backs every ``collect`` waits for do not queue behind it."""
        cs = self.copy_stream
        if bulk:
            if self._bulk_stream is None:
                self._bulk_stream = torch.cuda.Stream(self.device)
            cs = self._bulk_stream
        done = torch.cuda.Event()
        done.record(stream if stream is not None else self._main())
        cs.wait_event(done)
        via = C.c_void_p(cs.cuda_stream) if bulk else None
        if bulk:
            # the source was allocated on another stream (the exchange's): if its
            # owner drops it before this copy has run, the caching allocator must
            # not hand the block to that stream's next kernels
            for t in tensors:
                if t.is_cuda:
                    t.record_stream(cs)
        names = names or [None] * len(tensors)
        out = [self._to_host_async(t, name=nm, reserve=reserve, step=step, via=via)
               for t, nm in zip(tensors, names)]
        ready = torch.cuda.Event()
        ready.record(cs)
        return out + [ready]

    # -- one snapshot ----------------------------------------------------------
    def step(self, snapshot, halo_exists, region_positions, region_bulk_vels,
             H=0.0, want_angles=False, diagnostics=False, gpos=None):
        """Process one snapshot given as HOST arrays (the loader's dict).

        Mirrors the arguments the reference's ``track`` closure captures
        (``track_orbits.py:147-155``).  Returns a ``StepResult``; its event
        fields are ``None`` for a snapshot without a previous generation.
        """
        return self.collect(self.submit(
            snapshot, halo_exists, region_positions, region_bulk_vels, H,
            want_angles, diagnostics, gpos))

    def submit(self, snapshot, halo_exists, region_positions, region_bulk_vels,
               H=0.0, want_angles=False, diagnostics=False, gpos=None):
        coords = _as_f(snapshot['coordinates'], 'coordinates')
        vels = _as_f(snapshot['velocities'], 'velocities')
        if coords.dtype != vels.dtype:
            coords = coords.astype(np.float64)
            vels = vels.astype(np.float64)
        ids = np.asarray(snapshot['ids'])
        n = len(ids)
        masses = snapshot['masses']
        # the ring slot is free once the kernel that read it three steps ago
        # has finished
        slot = self._step % self.RING
        ev = self._consumed.get(slot)
        if ev is not None:
            self.h2d_stream.wait_event(ev)
        dev = {
            'pos': self._to_device(coords, name='in_pos'),
            'vel': self._to_device(vels, name='in_vel'),
            'ids': self._to_device(ids, np.int64, name='in_ids'),
            'mass': self._to_device(_as_f(masses, 'masses'), name='in_mass')
            if isinstance(masses, np.ndarray) else None,
        }
        uploaded = torch.cuda.Event()
        uploaded.record(self.h2d_stream)
        self._main().wait_event(uploaded)
        offsets = np.concatenate((
            np.asarray(snapshot['region_offsets'], dtype=np.int64), [n]))
        return self.submit_device(
            dev, n, coords.dtype, ids.dtype, offsets, halo_exists,
            region_positions, region_bulk_vels, H,
            box_size=snapshot.get('box_size'),
            redshift=snapshot.get('redshift', 0.0),
            mass_dtype=masses.dtype if isinstance(masses, np.ndarray) else None,
            want_angles=want_angles, diagnostics=diagnostics,
            gpos=self._to_device(gpos, np.int64) if gpos is not None else None)

    def step_device(self, *args, **kw):
        """``submit_device`` + ``collect``."""
        return self.collect(self.submit_device(*args, **kw))

    def submit_device(self, dev, n, data_dtype, ids_dtype, offsets,
                      halo_exists, region_positions, region_bulk_vels, H,
                      box_size=None, redshift=0.0, mass_dtype=None,
                      want_angles=False, diagnostics=False, gpos=None):
        """Same as ``submit`` with the particle arrays already in HBM
        (``dev`` = dict of flat torch tensors ``pos``, ``vel``, ``ids``,
        optional ``mass``)."""
        self._cur_main = torch.cuda.current_stream(self.device)
        try:
            return self._submit_device(
                dev, n, data_dtype, ids_dtype, offsets, halo_exists,
                region_positions, region_bulk_vels, H, box_size, redshift,
                mass_dtype, want_angles, diagnostics, gpos)
        finally:
            self._cur_main = None

    def _submit_device(self, dev, n, data_dtype, ids_dtype, offsets,
                       halo_exists, region_positions, region_bulk_vels, H,
                       box_size, redshift, mass_dtype, want_angles, diagnostics,
                       gpos):
        st = self._stream()
        if self.wait_before_submit is not None:
            # e.g. the multi-GPU exchange still reading the ring's event buffers
            self._main().wait_event(self.wait_before_submit)
            self.wait_before_submit = None
        data_dtype = np.dtype(data_dtype)
        halo_exists = np.asarray(halo_exists)
        n_h = len(halo_exists)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        assert len(offsets) == n_h + 1
        region_positions = np.asarray(region_positions)
        centre_f32 = region_positions.dtype == np.float32
        x64 = data_dtype == np.float64
        if self.onthefly:
            frame_f64 = x64
        else:
            frame_f64 = x64 or not centre_f32
        prev = self.prev
        if prev is not None and prev.frame_f64 != frame_f64:
            raise _lib.OrbitB200Error(
                "the halo-frame dtype changed between snapshots (float%d -> "
                "float%d); keep coordinates / region centres in one dtype"
                % (64 if prev.frame_f64 else 32, 64 if frame_f64 else 32))

        # ---- region table (oa_region rows) ---------------------------------
        # assembled by oa_region_rows_host straight into the pinned staging
        # buffer of the packed host->device copy [rows | offsets | seg_begin]
        # (numpy restatement: tests/test_abi_and_host.py)
        derive_bulk = region_bulk_vels is None
        cen = region_positions if region_positions.dtype in _F else \
            region_positions.astype(np.float64)
        cen = np.ascontiguousarray(cen).reshape(-1)
        if not derive_bulk:
            bulk = np.asarray(region_bulk_vels)
            bulk_dtype = bulk.dtype if bulk.dtype in _F else np.dtype(
                np.float64)
            bulk = np.ascontiguousarray(bulk, dtype=bulk_dtype).reshape(-1)
        else:
            bulk = None
            bulk_dtype = data_dtype if mass_dtype is None else np.result_type(
                data_dtype, mass_dtype)
        bulk_f32 = bulk_dtype == np.float32
        hx = np.ascontiguousarray(halo_exists, dtype=np.int64)
        nb_rows, nb_off = 128 * n_h, 8 * (n_h + 1)
        pack = self._hbuf('pack', nb_rows + nb_off + 8 * max(n_h, 1),
                          torch.uint8)
        hp = pack.numpy()
        buckets = np.empty(n_h, dtype=np.int64)
        matched = np.empty(n_h, dtype=np.bool_)
        prev_index = np.empty(n_h, dtype=np.int32)
        n_m_c = C.c_int(0)
        have_prev = prev is not None and len(prev.halo_exists) > 0
        if have_prev:
            phx = prev.halo_ids64
            p_args = (phx.ctypes.data, len(phx), prev.offsets.ctypes.data,
                      prev.buckets.ctypes.data)
        else:
            p_args = (None, 0, None, None)
        check(lib.oa_region_rows_host(
            n_h, offsets.ctypes.data, cen.ctypes.data,
            _lib.dtype_code(cen.dtype), bulk.ctypes.data if bulk is not None
            else None, _lib.dtype_code(bulk.dtype) if bulk is not None else 0,
            hx.ctypes.data, *p_args, pack.data_ptr(), buckets.ctypes.data,
            matched.ctypes.data, prev_index.ctypes.data,
            pack.data_ptr() + nb_rows + nb_off, C.byref(n_m_c)))
        n_m = n_m_c.value
        hp[nb_rows:nb_rows + nb_off] = offsets.view(np.uint8)
        used = nb_rows + nb_off + 8 * max(n_m, 1)
        d_pack = self._buf('pack', used, torch.uint8)
        d_pack.copy_(pack[:used], non_blocking=True)
        d_rows = d_pack[:nb_rows]
        d_off = d_pack[nb_rows:nb_rows + nb_off]
        d_seg = d_pack[nb_rows + nb_off:]

        d_bulk_out = None
        if derive_bulk:
            ws_bytes = lib.oa_bulk_workspace_bytes(n, n_h)
            ws = self._buf('bulk_ws', ws_bytes, torch.uint8)
            d_bulk_out = self._buf('bulk_out', max(3 * n_h, 1), torch.float64)
            check(lib.oa_bulk_velocity(
                ptr(dev['vel']), _lib.dtype_code(data_dtype), ptr(dev['mass']),
                _lib.dtype_code(mass_dtype) if mass_dtype is not None else 0,
                ptr(d_off), n_h, n, int(bulk_f32), ptr(d_rows),
                ptr(d_bulk_out), ptr(ws), ws_bytes, st))
            self.launches += 2

        lens = np.diff(offsets)
        gen = Generation()
        gen.n = n
        gen.frame_f64 = frame_f64
        gen.ids_dtype = np.dtype(ids_dtype)
        gen.offsets = offsets
        gen.halo_exists = halo_exists
        gen.halo_ids64 = hx
        gen.gpos = gpos
        gen.buckets = buckets
        gen.pjoin = self.impl in ('pjoin', 'pj2')
        gen.pj2 = self.impl == 'pj2'
        gen.ids = dev['ids']
        diag = dangle = out_angle = None
        if gen.pjoin:
            if frame_f64 or x64 or not centre_f32 or want_angles or diagnostics:
                raise _lib.OrbitB200Error(
                    "impl=%r needs float32 data and region centres and "
                    "has no per-particle outputs (checkpoint angles, "
                    "diagnostics); use impl='hash'" % self.impl)
            if prev is not None and (not prev.pjoin or prev.pj2 != gen.pj2):
                raise _lib.OrbitB200Error("generations of two implementations")
            launch = self._launch_pj2 if gen.pj2 else self._launch_pjoin
            launch(gen, prev, dev, d_rows, prev_index, matched,
                   lens, bulk_f32, box_size, H, redshift, st)
            tile_ws = a = None
        else:
            # ---- new generation buffers ------------------------------------------
            gen.index_bits = lib.oa_index_bits(int(lens.max()) if n_h else 0)
            rec_bytes = lib.oa_record_bytes(int(frame_f64))
            gen.rec = self._buf('rec', max(n, 1) * rec_bytes, torch.uint8)
            gen.tab = self._buf('tab', lib.oa_table_slots(n, n_h), torch.int32)
            gen.n_buckets = lib.oa_table_buckets(n, n_h)
            gen.mark = self._buf('mark', max(n, 1) + 8, torch.int16)

            fdt = torch.float64 if frame_f64 else torch.float32
            if want_angles:
                out_angle = self._empty(max(n, 1), torch.int16)
            if diagnostics:
                vr_dt = fdt if self.onthefly else torch.float64
                diag = {'rhat': self._empty(3 * n, fdt), 'vr': self._empty(n, vr_dt),
                        'r': self._empty(n, fdt),
                        'match': self._empty(n, torch.int64)}
            if self.onthefly and prev is not None:
                dangle = self._empty(max(prev.n, 1), fdt)

            a = _lib.TrackArgs()
            a.pos, a.vel, a.ids = ptr(dev['pos']), ptr(dev['vel']), ptr(dev['ids'])
            a.n_cur = n
            a.cur_off, a.regions = ptr(d_off), ptr(d_rows)
            a.n_regions = n_h
            a.data_dtype = int(x64)
            a.frame_dtype = int(frame_f64)
            a.centre_f32 = int(centre_f32)
            a.bulk_f32 = int(bulk_f32)
            a.periodic = int(box_size is not None)
            a.onthefly = int(self.onthefly)
            a.mode = _lib.OA_MODE[self.mode]
            if box_size is not None:
                box = np.broadcast_to(
                    np.asarray(box_size, dtype=np.float64), (3,))
                a.box[0], a.box[1], a.box[2] = float(box[0]), float(box[1]), \
                    float(box[2])
            a.hubble = float(H)
            a.one_plus_z = 1 + float(redshift)
            if prev is not None:
                a.rec_prev, a.tab_prev = ptr(prev.rec), ptr(prev.tab)
                a.tab_prev_buckets = prev.n_buckets
                a.n_prev = prev.n
                a.prev_index_bits = prev.index_bits
                a.mark_prev = ptr(prev.mark)
            else:
                a.prev_index_bits = 1
            a.cur_index_bits = gen.index_bits
            a.rec_cur, a.tab_cur, a.mark_cur = ptr(gen.rec), ptr(gen.tab), \
                ptr(gen.mark)
            a.out_angle = ptr(out_angle)
            if diag is not None:
                a.out_rhat, a.out_vr, a.out_r = ptr(diag['rhat']), \
                    ptr(diag['vr']), ptr(diag['r'])
                a.out_match = ptr(diag['match'])
            a.dangle_prev = ptr(dangle)
            a.tab_cur_buckets = gen.n_buckets
            a.sm_reserve = self.sm_reserve
            a.workspace_bytes = lib.oa_track_workspace_bytes(n)
            tile_ws = self._buf('chunk_ws', a.workspace_bytes, torch.uint8)
            a.workspace = ptr(tile_ws)
            check(lib.oa_table_clear(ptr(gen.tab), n, n_h, st))
            if self.timing is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), \
                    torch.cuda.Event(enable_timing=True)
                ev0.record(self._main())
            check(lib.oa_track_fused(C.byref(a), st))
            if self.timing is not None:
                ev1.record(self._main())
                self.timing.append((ev0, ev1, n))
            self.launches += 3
        consumed = torch.cuda.Event()
        consumed.record(self._main())
        self._consumed[self._step % self.RING] = consumed

        p = Pending()
        p.step = self._step
        p.n, p.n_h, p.n_m = n, n_h, n_m
        p.gen, p.prev, p.matched = gen, prev, matched
        p.diag, p.dangle = diag, dangle
        p.derive_bulk, p.bulk_dtype = derive_bulk, bulk_dtype
        p.region_bulk_vels = region_bulk_vels
        p.keep = (dev, d_pack, a, tile_ws)       # inputs stay alive until collected
        p.h_bulk = p.h_angle = p.h_small = None
        p.sel = p.d_ids = p.d_ang = None

        # ---- ordered event compaction, all enqueued without a host sync -------
        if prev is not None and not self.onthefly:
            cap = max(min(prev.n, n), 1)
            ws_bytes = lib.oa_select_workspace_bytes(prev.n)
            ws = self._buf('sel_ws', ws_bytes, torch.uint8)
            d_small = self._buf('small', n_m + 1, torch.int64)  # offsets..., total
            d_total = d_small[n_m:]
            check(lib.oa_select_count(
                ptr(prev.mark), prev.n, _lib.OA_SEL_NE, _lib.OA_NO_EVENT,
                ptr(ws), ws_bytes, ptr(d_total), st))
            # (valid until RING further snapshots have been submitted)
            p.sel = self._buf('sel', cap, torch.int64)
            p.d_ids = self._buf('ev_ids', cap, torch.int64)
            p.d_ang = self._buf('ev_ang', cap, torch.int16)
            if prev.pjoin:
                check(lib.oa_select_gather_events_ids(
                    ptr(prev.mark), prev.n, ptr(ws), ptr(prev.ids), ptr(p.sel),
                    ptr(p.d_ids), ptr(p.d_ang), st))
            else:
                check(lib.oa_select_gather_events(
                    ptr(prev.mark), prev.n, ptr(ws), ptr(prev.rec),
                    int(prev.frame_f64), ptr(p.sel), ptr(p.d_ids), ptr(p.d_ang),
                    st))
            check(lib.oa_segment_offsets(
                ptr(p.sel), cap, ptr(d_total), ptr(d_seg), n_m, ptr(d_small),
                st))
            self.launches += 4
            p.keep += (ws, d_small)
        else:
            d_small = None
        p.d_small = d_small

        # ---- small device->host copies on the copy stream ----------------------
        done = torch.cuda.Event()
        done.record(self._main())
        p.compacted = done
        self.copy_stream.wait_event(done)
        if prev is not None and prev.pjoin and d_small is not None:
            # the event gather above read the PREVIOUS snapshot's ID array: its
            # ring slot may only be refilled by the upload stream after this
            self._consumed[(self._step - 1) % self.RING] = done
        p.h_ids = p.h_ang = None
        p.n_spec = 0
        if d_small is not None:
            p.h_small = self._to_host_async(d_small, name='h_small')
            # speculative copy of the event lists: their exact length is only
            # known on the device, so copy as many records as the last snapshot
            # produced (+25 %) now and the remainder, if any, in collect()
            if not self.events_on_device:
                p.n_spec = min(cap, int(self._last_events * 1.25) + 1024)
                # (pinned capacity for one event per six particles up front:
                # re-pinning costs milliseconds)
                p.h_ids = self._to_host_async(p.d_ids, p.n_spec, 'h_ids',
                                              reserve=cap // 6)
                p.h_ang = self._to_host_async(p.d_ang, p.n_spec, 'h_ang',
                                              reserve=cap // 6)
        p.h_overflow = None
        if gen.pj2:
            p.h_overflow = self._to_host_async(self._pj2_overflow, 1, 'h_ovf')
        if derive_bulk:
            p.h_bulk = self._to_host_async(d_bulk_out, 3 * n_h, 'h_bulk')
            p.keep += (d_bulk_out,)
        if out_angle is not None:
            p.h_angle = self._to_host_async(out_angle, n, 'h_angle')
            p.keep += (out_angle,)
        p.small_done = torch.cuda.Event()
        p.small_done.record(self.copy_stream)
        self.prev = gen
        self._step += 1
        return p

    def _launch_pjoin(self, gen, prev, dev, d_rows, prev_index, matched, lens,
                      bulk_f32, box_size, H, redshift, st):
        """Enqueue ``oa_pjoin_step`` for the current snapshot (see pjoin.py)."""
        n, n_h = gen.n, len(lens)
        prev_bits = np.full(n_h, -1, dtype=np.int32)
        prev_pb = np.zeros(n_h, dtype=np.int64)
        prev_counts = np.zeros(n_h, dtype=np.int64)
        if prev is not None and matched.any():
            k = prev_index[matched]
            prev_bits[matched] = prev.pj_bits[k]
            prev_pb[matched] = prev.pj_pb[k]
            prev_counts[matched] = prev.offsets[k + 1] - prev.offsets[k]
        if self._planner is None:
            self._planner = pjoin.Planner(lib)
        plan = self._planner(gen.offsets, prev_bits, prev_pb, prev_counts)
        gen.pj_bits, gen.pj_pb = plan.bits, plan.pb
        # one packed host->device copy: [plan rows | group_first | range_start]
        nb_rows = plan.rows.nbytes
        nb_grp = -(-plan.group_first.nbytes // 16) * 16
        nb_rng = -(-plan.range_start.nbytes // 16) * 16
        pack = self._hbuf('pjpack', nb_rows + nb_grp + nb_rng, torch.uint8)
        hp = pack.numpy()
        hp[:nb_rows] = plan.rows.view(np.uint8)
        hp[nb_rows:nb_rows + plan.group_first.nbytes] = \
            plan.group_first.view(np.uint8)
        hp[nb_rows + nb_grp:nb_rows + nb_grp + plan.range_start.nbytes] = \
            plan.range_start.view(np.uint8)
        d_pack = self._buf('pjpack', pack.numel(), torch.uint8)
        d_pack.copy_(pack, non_blocking=True)
        gen.rec = self._buf('rec', max(n, 1) * 32, torch.uint8)
        gen.mark = self._buf('mark', max(n, 1) + 8, torch.int16)
        gen.part_off = self._buf('poff', max(plan.n_entries, 1), torch.int32)
        ws_bytes = lib.oa_pjoin_workspace_bytes(n_h, plan.n_entries, plan.total)
        ws = self._buf('pj_ws', ws_bytes, torch.uint8)

        a = pjoin.PJoinArgs()
        a.pos, a.vel, a.ids = ptr(dev['pos']), ptr(dev['vel']), ptr(dev['ids'])
        a.n_cur = n
        a.regions = ptr(d_rows)
        a.plan = ptr(d_pack)
        a.group_first = ptr(d_pack[nb_rows:])
        a.range_start = ptr(d_pack[nb_rows + nb_grp:])
        a.n_regions, a.n_groups, a.n_ranges = n_h, plan.n_groups, plan.n_ranges
        a.centre_f32, a.bulk_f32 = 1, int(bulk_f32)
        a.periodic = int(box_size is not None)
        if box_size is not None:
            box = np.broadcast_to(np.asarray(box_size, dtype=np.float64), (3,))
            a.box[0], a.box[1], a.box[2] = float(box[0]), float(box[1]), \
                float(box[2])
        a.mode = _lib.OA_MODE[self.mode]
        a.hubble_on, a.hubble = int(float(H) != 0.0), float(H)
        a.one_plus_z = 1 + float(redshift)
        if prev is not None:
            a.rec_prev, a.part_off_prev = ptr(prev.rec), ptr(prev.part_off)
            a.mark_prev, a.n_prev = ptr(prev.mark), prev.n
        a.rec_cur, a.part_off_cur = ptr(gen.rec), ptr(gen.part_off)
        a.mark_cur = ptr(gen.mark)
        a.workspace, a.workspace_bytes = ptr(ws), ws_bytes
        a.n_part_entries, a.total_tickets = plan.n_entries, plan.total
        a.sm_reserve = self.sm_reserve
        if self.timing is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), \
                torch.cuda.Event(enable_timing=True)
            ev0.record(self._main())
        check(lib.oa_pjoin_step(C.byref(a), st))
        if self.timing is not None:
            ev1.record(self._main())
            self.timing.append((ev0, ev1, n))
        self.launches += 3          # memset, item expansion, persistent kernel
        self._pj_keep = (a, d_pack, ws)

    def _launch_pj2(self, gen, prev, dev, d_rows, prev_index, matched, lens,
                    bulk_f32, box_size, H, redshift, st):
        """Enqueue ``oa_pj2_step`` for the current snapshot (see pj2.py)."""
        n, n_h = gen.n, len(lens)
        prev_plan = np.zeros((4, n_h), dtype=np.uint32)
        if prev is not None and matched.any():
            k = prev_index[matched]
            nonempty = (prev.offsets[k + 1] - prev.offsets[k]) > 0
            cols = np.flatnonzero(matched)[nonempty]
            prev_plan[:, cols] = prev.pj2_plan[:, k[nonempty]]
        if self._planner is None:
            self._planner = pj2.Planner(lib)
            self._pj2_overflow = torch.zeros(4, dtype=torch.int32,
                                             device=self.device)
        plan = self._planner(gen.offsets, *prev_plan)
        gen.pj2_plan = np.stack((plan.P, plan.cap, plan.base, plan.pb))
        # one packed host->device copy: [rows | group_first | group_off | range_start]
        parts = (plan.rows, plan.group_first, plan.group_off, plan.range_start)
        sizes = [-(-v.nbytes // 16) * 16 for v in parts]
        pack = self._hbuf('pj2pack', sum(sizes), torch.uint8)
        hp = pack.numpy()
        at, starts = 0, []
        for v, sz in zip(parts, sizes):
            hp[at:at + v.nbytes] = v.view(np.uint8).reshape(-1)
            starts.append(at)
            at += sz
        d_pack = self._buf('pj2pack', pack.numel(), torch.uint8)
        d_pack.copy_(pack, non_blocking=True)
        gen.rec = self._buf('rec', max(plan.n_slots, 1) * 32, torch.uint8)
        gen.mark = self._buf('mark', max(n, 1) + 8, torch.int16)
        gen.fill = self._buf('fill', max(plan.n_entries, 1), torch.int32)
        ws_bytes = lib.oa_pj2_workspace_bytes(plan.n_groups, plan.total)
        ws = self._buf('pj_ws', ws_bytes, torch.uint8)

        a = pj2.PJ2Args()
        a.pos, a.vel, a.ids = ptr(dev['pos']), ptr(dev['vel']), ptr(dev['ids'])
        a.n_cur = n
        a.regions = ptr(d_rows)
        a.plan = ptr(d_pack[starts[0]:])
        a.group_first = ptr(d_pack[starts[1]:])
        a.group_off = ptr(d_pack[starts[2]:])
        a.range_start = ptr(d_pack[starts[3]:])
        a.n_regions, a.n_groups, a.n_ranges = n_h, plan.n_groups, plan.n_ranges
        a.centre_f32, a.bulk_f32 = 1, int(bulk_f32)
        a.periodic = int(box_size is not None)
        if box_size is not None:
            box = np.broadcast_to(np.asarray(box_size, dtype=np.float64), (3,))
            a.box[0], a.box[1], a.box[2] = float(box[0]), float(box[1]), \
                float(box[2])
        a.mode = _lib.OA_MODE[self.mode]
        a.hubble_on, a.hubble = int(float(H) != 0.0), float(H)
        a.one_plus_z = 1 + float(redshift)
        if prev is not None:
            a.rec_prev, a.fill_prev = ptr(prev.rec), ptr(prev.fill)
            a.mark_prev, a.n_prev = ptr(prev.mark), prev.n
        a.rec_cur, a.fill_cur = ptr(gen.rec), ptr(gen.fill)
        a.mark_cur = ptr(gen.mark)
        a.workspace, a.workspace_bytes = ptr(ws), ws_bytes
        a.n_part_entries, a.n_rec_slots = plan.n_entries, plan.n_slots
        a.total_tickets = plan.total
        a.sm_reserve = self.sm_reserve
        a.overflow = ptr(self._pj2_overflow)
        if self.timing is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), \
                torch.cuda.Event(enable_timing=True)
            ev0.record(self._main())
        check(lib.oa_pj2_step(C.byref(a), st))
        if self.timing is not None:
            ev1.record(self._main())
            self.timing.append((ev0, ev1, n))
        self.launches += 4          # 2 memsets, item expansion, persistent kernel
        self._pj_keep = (a, d_pack, ws)

    def collect_keep(self, p):
        """``collect`` that leaves the snapshot's device inputs and per-particle
        outputs (``p.keep``, ``p.diag``, ``p.dangle``) alive for the caller."""
        return self.collect(p, release=False)

    def collect(self, p, release=True):
        """Wait for a submitted snapshot and return its ``StepResult``."""
        res = StepResult()
        res.step = p.step
        res.n = p.n
        res.diag = p.diag
        res.hinds = np.flatnonzero(p.matched)
        res.apsis_ids = res.apsis_angles = res.apsis_offsets = None
        res.apsis_prev_index = None
        res.n_events = 0
        res.angles = None
        res.d_ids = res.d_ang = None
        res.d_sel = res.d_ids_buf = res.d_ang_buf = res.d_small = None
        res.host_ready = res.host_slice = None
        res.compacted = p.compacted
        res.prev_gen = p.prev
        res.prepack = getattr(p, 'prepack', None)
        p.small_done.synchronize()
        if p.h_overflow is not None and int(p.h_overflow[0]) != 0:
            raise _lib.OrbitB200Error(
                "impl='pj2': %d records did not fit their ID-hash partition "
                "(head room of %d sigmas exceeded: duplicated or adversarial "
                "IDs?); use impl='hash'" % (int(p.h_overflow[0]), pj2.SIGMAS))
        if p.derive_bulk:
            res.bulk_velocities = p.h_bulk.numpy().reshape(p.n_h, 3).astype(
                p.bulk_dtype)
        else:
            res.bulk_velocities = np.asarray(p.region_bulk_vels)
        if p.h_angle is not None:
            res.angles = p.h_angle.numpy().view(np.float16)
        if p.h_small is not None:
            small = p.h_small.numpy()
            total = int(small[p.n_m])
            res.apsis_offsets = small.copy()
            self._last_events = total
            res.n_events = total
            res.apsis_prev_index = p.sel[:total]
            res.d_ids, res.d_ang = p.d_ids[:total], p.d_ang[:total]
            # whole device buffers + [offsets | total] for the multi-GPU pack
            res.d_sel, res.d_ids_buf, res.d_ang_buf = p.sel, p.d_ids, p.d_ang
            res.d_small = p.d_small
            if self.events_on_device:
                if release:
                    p.keep = None
                return res
            if total > p.n_spec:          # the speculative copy fell short
                p.h_ids = self._to_host_async(p.d_ids, total, 'h_ids', p.step)
                p.h_ang = self._to_host_async(p.d_ang, total, 'h_ang', p.step)
                self.copy_stream.synchronize()
            res.n_events = total
            ids = p.h_ids.numpy()[:total]
            res.apsis_ids = ids if p.prev.ids_dtype == np.int64 else \
                ids.astype(p.prev.ids_dtype)
            res.apsis_angles = p.h_ang.numpy()[:total].view(np.float16)
            res.apsis_prev_index = p.sel[:total]
        if release:
            p.keep = None
        return res

    # -- checkpoint of the carried device state ----------------
    def save_state(self):
        """The carried state of the last submitted snapshot as host arrays:
        everything ``track(j)`` needs from the previous snapshot (reference
        ``track_orbits.py:234-240``: ids, unit vectors, radial velocities, angle
        accumulators) in this library's record layout, plus the ID table.  A run
        restored with ``load_state`` continues at the NEXT snapshot without
        reloading this one (the reference re-processes the last saved snapshot
        and restores only the angles, ``:93-101, 229-232``).  Synchronous."""
        gen = self.prev
        if gen is None:
            raise _lib.OrbitB200Error("no snapshot has been processed yet")
        if gen.pjoin:
            raise _lib.OrbitB200Error(
                "device-state checkpoints need impl='hash'")
        torch.cuda.current_stream(self.device).synchronize()
        meta = np.array([gen.n, gen.index_bits, gen.n_buckets,
                         int(gen.frame_f64), len(gen.halo_exists)],
                        dtype=np.int64)
        state = {'meta': meta, 'rec': gen.rec.cpu().numpy(),
                 'tab': gen.tab.cpu().numpy(),
                 'offsets': np.asarray(gen.offsets, dtype=np.int64),
                 'halo_exists': np.asarray(gen.halo_exists, dtype=np.int64),
                 'buckets': np.asarray(gen.buckets, dtype=np.int64),
                 'ids_dtype': np.frombuffer(
                     gen.ids_dtype.str.encode().ljust(8), dtype=np.uint8).copy()}
        if gen.gpos is not None:
            state['gpos'] = gen.gpos.cpu().numpy()
        return state

    def load_state(self, state):
        """Inverse of ``save_state`` on a fresh tracker."""
        meta = np.asarray(state['meta'], dtype=np.int64)
        n, index_bits, n_buckets, frame_f64, n_h = (int(v) for v in meta)
        slot = (self._step - 1) % self.RING      # the slot of "the previous step"
        gen = Generation()
        gen.n, gen.index_bits, gen.n_buckets = n, index_bits, n_buckets
        gen.frame_f64 = bool(frame_f64)
        gen.offsets = np.ascontiguousarray(state['offsets'], dtype=np.int64)
        gen.halo_exists = np.asarray(state['halo_exists'], dtype=np.int64)
        gen.halo_ids64 = np.ascontiguousarray(gen.halo_exists, dtype=np.int64)
        gen.buckets = np.ascontiguousarray(state['buckets'], dtype=np.int64)
        gen.ids_dtype = np.dtype(bytes(np.asarray(
            state['ids_dtype'], dtype=np.uint8)).decode().strip())
        gen.pjoin = gen.pj2 = False
        gen.ids = None
        rec = np.ascontiguousarray(state['rec'], dtype=np.uint8)
        tab = np.ascontiguousarray(state['tab'], dtype=np.int32)
        if len(rec) != max(n, 1) * lib.oa_record_bytes(int(gen.frame_f64)) or \
                len(tab) != lib.oa_table_slots(n, n_h) or \
                len(gen.offsets) != n_h + 1:
            raise ValueError("device-state checkpoint is inconsistent")
        gen.rec = self._buf('rec', len(rec), torch.uint8, slot)
        gen.rec.copy_(torch.from_numpy(rec))
        gen.tab = self._buf('tab', len(tab), torch.int32, slot)
        gen.tab.copy_(torch.from_numpy(tab))
        gen.mark = self._buf('mark', max(n, 1) + 8, torch.int16, slot)
        check(lib.oa_fill_u16(ptr(gen.mark), max(n, 1) + 8, _lib.OA_NO_EVENT,
                              self._stream()))
        gen.gpos = None
        if 'gpos' in state:
            gen.gpos = torch.from_numpy(np.ascontiguousarray(
                state['gpos'], dtype=np.int64)).to(self.device)
        self.launches += 1
        self.prev = gen

    def load_angles(self, angles):
        """Replace the angle accumulators of the current generation with a
        host float16 array in block order (resume, reference
        ``track_orbits.py:229-232``)."""
        gen = self.prev
        if gen is not None and gen.pjoin:
            raise _lib.OrbitB200Error("impl='pjoin' cannot resume from a checkpoint")
        angles = np.ascontiguousarray(angles, dtype=np.float16)
        if gen is None or len(angles) != gen.n:
            raise ValueError("checkpoint does not match the resumed snapshot")
        d = self._to_device(angles.view(np.int16))
        check(lib.oa_set_record_angles(ptr(gen.rec), int(gen.frame_f64), ptr(d),
                                       gen.n, self._stream()))
        self.launches += 1

    # -- synchronous helpers (on-the-fly driver, tests) -------------------------
    def select(self, marks, n, op, value):
        """Ascending positions i < n with ``marks[i] (op) value`` (device
        int64 tensor) and their count."""
        st = self._stream()
        ws_bytes = lib.oa_select_workspace_bytes(n)
        ws = self._empty(ws_bytes, torch.uint8)
        d_total = self._empty(1, torch.int64)
        check(lib.oa_select_count(ptr(marks), n, op, value, ptr(ws), ws_bytes,
                                  ptr(d_total), st))
        total = int(d_total.item())
        sel = self._empty(max(total, 1), torch.int64)
        if total:
            check(lib.oa_select_gather(ptr(marks), n, op, value, ptr(ws),
                                       ptr(sel), st))
        self.launches += 3
        return sel, total

    def segment_offsets(self, sel, total, seg_begin):
        """Number of selected positions below each segment start (host)."""
        st = self._stream()
        seg_begin = np.ascontiguousarray(seg_begin, dtype=np.int64)
        if len(seg_begin) == 0:
            return np.zeros(0, dtype=np.int64)
        d_seg = self._to_device(seg_begin)
        d_out = self._empty(len(seg_begin), torch.int64)
        check(lib.oa_segment_offsets(ptr(sel), total, None, ptr(d_seg),
                                     len(seg_begin), ptr(d_out), st))
        self.launches += 1
        return d_out.cpu().numpy()
