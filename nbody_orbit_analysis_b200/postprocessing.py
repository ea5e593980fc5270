"""Drop-in ``orbitanalysis.postprocessing.Apsides`` on B200.

Same class, methods, arguments, ``ValueError``s and collated-file layout as the
reference ``postprocessing.py:8-240``.  File reading/writing stays on the host
(``storage.py``); the reductions move to the GPU:

* ``collate_apsides``: the collated state -- per halo the unique event IDs that
  passed ``angles > angle_cut`` so far and their counts, i.e. the reference's
  ``np.unique(..., return_counts=True)`` over a halo's whole history
  (``:133-141``) -- is ONE device table ascending in (halo, ID).  Per snapshot
  only the new events are sorted (two radix sorts) and run-length encoded, then
  merged into the table (``oa_merge_find`` / ``oa_merge_place``): the history is
  never re-sorted, O(E_total) instead of the reference's O(n_snap x E_total);
* ``save_final_apsis_counts``: the per-halo ``myin1d`` look-ups (``:222-232``)
  become one segmented binary-search join.

Quirks kept on purpose: ``*_counts_final`` is float64
(``np.empty``, ``:224``); ``halo_offsets`` only counts the halos present at the
snapshot while ``particle_IDs`` concatenates every pool (``:134-142``).
"""
import time

import numpy as np
import torch

from . import _lib, storage
from ._device import DeviceContext
from ._lib import lib, check, ptr
from .utils import myin1d


class Apsides:

    def __init__(self, filename, device=None):
        """Read the file header of a ``track_orbits`` result
        (reference ``postprocessing.py:10-28``)."""
        self.filename = filename
        self._device = device
        with storage.File(filename, 'r') as hf:
            keys = list(hf.keys())
            self.snapshot_numbers = np.array(
                [int(k.split('_')[1]) for k in keys])
            self.final_halo_ids = hf[keys[-1]]['halo_IDs'][:]
            self.mode = hf.attrs['mode']
            if 'box_size' in hf.attrs:
                self.box_size = hf.attrs['box_size']

    # ------------------------------------------------------------------------
    def collate_apsides(self, halo_ids=None, snapshot_number=None,
                        angle_cut=np.pi / 4, save_final_counts=False,
                        data_type=None, savefile=None, verbose=True):
        """Reference ``postprocessing.py:30-174``."""
        t_start = time.time()
        tag = self.mode[:-3] + 'er'
        if halo_ids is None:
            halo_ids = self.final_halo_ids
        elif len(np.intersect1d(self.final_halo_ids, halo_ids)) < len(
                halo_ids):
            self.missing_halo_ids = np.setdiff1d(
                halo_ids, self.final_halo_ids)
            raise ValueError(
                "Some of the halo IDs supplied were not tracked. The missing "
                "IDs have been stored in the attribute `missing_halo_ids`.")
        halo_ids = np.asarray(halo_ids)
        if snapshot_number is None:
            last = len(self.snapshot_numbers) - 1
        else:
            last = int(np.flatnonzero(
                self.snapshot_numbers == snapshot_number)[0])

        # numpy compares a float16 array with a PYTHON float in float16
        # (NEP 50 weak scalar), with a numpy scalar in that scalar's precision
        cut = float(np.float16(angle_cut)) if type(angle_cut) in (float, int) \
            else float(angle_cut)

        ctx = DeviceContext(self._device)
        st = ctx.stream()
        n_pool = len(halo_ids)
        # collated state: (pool index, ID, count) device int64 arrays, ascending
        # in (pool, ID); a snapshot's events are MERGED into it (section f-3)
        table_state = None
        idtype = None

        for s in self.snapshot_numbers[:last + 1]:
            is_final = s == self.snapshot_numbers[-1]
            with storage.File(self.filename, 'r') as hf:
                g = hf['snapshot_%03d' % s]
                region_positions = g['region_positions'][:]
                region_radii = g['region_radii'][:]
                bulk_velocities = g['bulk_velocities'][:]
                halo_ids_current = g['halo_IDs'][:]
                halo_ids_final = halo_ids_current if is_final else \
                    g['final_descendant_IDs'][:]
                common = np.intersect1d(halo_ids_final, halo_ids)
                hinds1 = myin1d(halo_ids_final, common)
                hinds2 = myin1d(halo_ids, common)
                if len(g[tag + '_IDs']) == 0:
                    continue
                if idtype is None:
                    idtype = g[tag + '_IDs'].dtype if data_type is None \
                        else np.dtype(data_type)
                offs = np.asarray(g['region_offsets'][:], dtype=np.int64)
                ev_ids = g[tag + '_IDs'][:]
                ev_ang = g['angles'][:]

            # ---- this snapshot's events -> pool --------------------------------
            E = len(ev_ids)
            n_m = len(offs) - 1
            table = np.full(max(n_m, 1), n_pool, dtype=np.int32)  # n_pool: drop
            table[hinds1] = hinds2
            d_ids = ctx.upload(ev_ids, np.int64)
            d_ang = ctx.upload(np.ascontiguousarray(
                ev_ang, dtype=np.float16).view(np.int16))
            marks = ctx.empty(E + 8, torch.int16)
            check(lib.oa_angle_cut(ptr(d_ang), E, cut, ptr(marks), st))
            sel, n_sel = ctx.select(marks, E, _lib.OA_SEL_EQ, 1)
            if n_sel:
                seg32 = ctx.empty(E, torch.int32)
                # (keep both uploads referenced until the launch: a released
                # tensor's block is handed to the very next allocation)
                d_offs, d_table = ctx.upload(offs), ctx.upload(table)
                check(lib.oa_expand_segments(ptr(d_offs), n_m, ptr(d_table), E,
                                             ptr(seg32), st))
                new_ids = ctx.gather_i64(d_ids, sel, n_sel)[:n_sel]
                new_seg = ctx.gather_i64(seg32.to(torch.int64), sel,
                                         n_sel)[:n_sel]
                keep = new_seg < n_pool           # halos that are collated
                new_ids, new_seg = new_ids[keep], new_seg[keep]
                n_new = int(new_ids.numel())
                if n_new:
                    table_state = self._merge_new(
                        ctx, table_state, new_ids, new_seg, n_new, n_pool)
            ctx.launches += 2

            # ---- unique IDs + counts of every pool ------------------------------
            if table_state is not None:
                t_seg, t_ids, t_cnt = table_state
                n_runs = int(t_ids.numel())
                # number of unique IDs in the pools before pool h
                d_keys = ctx.upload(np.arange(n_pool + 1, dtype=np.int64))
                d_poff = ctx.empty(n_pool + 1, torch.int64)
                check(lib.oa_segment_offsets(ptr(t_seg), n_runs, None,
                                             ptr(d_keys), n_pool + 1,
                                             ptr(d_poff), st))
                ctx.launches += 1
                pool_off = d_poff[:n_pool + 1].cpu().numpy()
                particle_ids = t_ids.cpu().numpy().astype(idtype, copy=False)
                particle_counts = t_cnt.cpu().numpy()
            else:
                pool_off = np.zeros(n_pool + 1, dtype=np.int64)
                particle_ids = np.zeros(0, dtype=idtype)
                particle_counts = np.zeros(0, dtype=np.int64)
            pool_len = np.diff(pool_off)
            present = np.zeros(n_pool, dtype=bool)
            present[hinds2] = True
            lens = [int(pool_len[i]) for i in range(n_pool) if present[i]]

            with storage.File(savefile, 'a') as hf:
                g = hf.create_group('snapshot_%03d' % s)
                g.create_dataset('particle_IDs', data=particle_ids)
                g.create_dataset(tag + '_counts', data=particle_counts)
                g.create_dataset('halo_offsets',
                                 data=np.cumsum([0] + lens)[:-1])
                if not is_final:
                    g.create_dataset('final_descendant_IDs',
                                     data=halo_ids_final[hinds1])
                g.create_dataset('halo_IDs', data=halo_ids_current[hinds1])
                g.create_dataset('halo_positions',
                                 data=region_positions[hinds1])
                g.create_dataset('halo_velocities',
                                 data=bulk_velocities[hinds1])
                g.create_dataset('region_radii', data=region_radii[hinds1])

        if save_final_counts:
            self.save_final_apsis_counts(savefile, verbose=verbose)
        if verbose:
            print('Collated apsides in {} s\n'.format(time.time() - t_start))

    # ------------------------------------------------------------------------
    @staticmethod
    def _merge_new(ctx, table, new_ids, new_seg, n_new, n_pool):
        """Merge one snapshot's cut-passing events ``(new_seg, new_ids)`` into
        the collated table: the reference's per-halo ``np.unique(...,
        return_counts=True)`` over the whole history (``postprocessing.py:
        133-141``) without touching the history -- only the new events are
        sorted, the table is merged (``oa_merge_find`` / ``oa_merge_place``)."""
        st = ctx.stream()
        # (pool, ID)-sorted new events, run-length encoded
        _, order, _ = ctx.argsort_values(new_ids, n_new)
        seg_sorted = ctx.gather_i64(new_seg, order, n_new)
        seg_sorted, order = ctx.sort_pairs(
            seg_sorted, order, n_new, max(n_pool - 1, 1).bit_length())
        ids_sorted = ctx.gather_i64(new_ids, order, n_new)
        head = ctx.empty(n_new + 8, torch.int16)
        check(lib.oa_run_heads(ptr(seg_sorted), ptr(ids_sorted), n_new,
                               ptr(head), st))
        starts, n_runs = ctx.select(head, n_new, _lib.OA_SEL_EQ, 1)
        counts = ctx.empty(n_runs, torch.int64)
        check(lib.oa_run_lengths(ptr(starts), n_runs, n_new, ptr(counts), st))
        u_ids = ctx.gather_i64(ids_sorted, starts, n_runs)[:n_runs]
        u_seg = ctx.gather_i64(seg_sorted, starts, n_runs)[:n_runs]
        counts = counts[:n_runs]
        ctx.launches += 2
        if table is None:
            return u_seg, u_ids, counts
        t_seg, t_ids, t_cnt = table
        n_tab = int(t_ids.numel())
        lb = ctx.empty(n_runs, torch.int64)
        miss = ctx.empty(n_runs + 8, torch.int16)
        check(lib.oa_merge_find(ptr(t_seg), ptr(t_ids), ptr(t_cnt), n_tab,
                                ptr(u_seg), ptr(u_ids), ptr(counts), n_runs,
                                ptr(lb), ptr(miss), st))
        msel, n_miss = ctx.select(miss, n_runs, _lib.OA_SEL_EQ, 1)
        ctx.launches += 1
        if n_miss == 0:
            return table
        out = [ctx.empty(n_tab + n_miss, torch.int64)[:n_tab + n_miss]
               for _ in range(3)]
        check(lib.oa_merge_place(ptr(t_seg), ptr(t_ids), ptr(t_cnt), n_tab,
                                 ptr(u_seg), ptr(u_ids), ptr(counts), ptr(lb),
                                 ptr(msel), n_miss, ptr(out[0]), ptr(out[1]),
                                 ptr(out[2]), st))
        ctx.launches += 1
        return tuple(out)

    # ------------------------------------------------------------------------
    def save_final_apsis_counts(self, collated_file, snapshot_numbers=None,
                                verbose=True):
        """Reference ``postprocessing.py:176-240``: for every earlier snapshot,
        the FINAL passage count of each of its particles."""
        tag = self.mode[:-3] + 'er'
        ctx = DeviceContext(self._device)
        st = ctx.stream()
        with storage.File(collated_file, 'r+') as hf:
            keys = np.array(list(hf.keys()))
            fin = hf[keys[-1]]
            ids_final = fin['particle_IDs'][:]
            counts_final = fin[tag + '_counts'][:]
            halo_ids = fin['halo_IDs'][:]
            fo = np.concatenate((np.asarray(fin['halo_offsets'][:],
                                            dtype=np.int64), [len(ids_final)]))
            if snapshot_numbers is None:
                todo = keys[:-1]
            else:
                nums = np.array([int(k.split('_')[-1]) for k in keys])
                todo = keys[np.isin(nums, snapshot_numbers)]
            # the final per-halo ID lists are sorted (np.unique): they are the
            # keys of a segmented binary-search join
            nf = len(ids_final)
            bias = int(ids_final.min()) if nf else 0
            d_keys = ctx.upload((ids_final.astype(np.int64) - bias))
            d_fo = ctx.upload(fo)
            d_cf = ctx.upload(counts_final, np.int64)
            for key in todo:
                g = hf[key]
                ids = g['particle_IDs'][:]
                desc = g['final_descendant_IDs'][:]
                n = len(ids)
                so = np.concatenate((np.asarray(g['halo_offsets'][:],
                                                dtype=np.int64), [n]))
                hinds = myin1d(halo_ids, desc)
                retro = np.empty(n)
                if n and len(hinds):
                    n_seg = len(so) - 1
                    table = np.full(n_seg, -1, dtype=np.int32)
                    table[:len(hinds)] = hinds
                    q_seg = ctx.empty(n, torch.int32)
                    d_so, d_table = ctx.upload(so), ctx.upload(table)
                    d_q = ctx.upload(ids, np.int64)
                    check(lib.oa_expand_segments(ptr(d_so), n_seg, ptr(d_table),
                                                 n, ptr(q_seg), st))
                    pos = ctx.empty(n, torch.int64)
                    check(lib.oa_lookup_sorted(
                        ptr(d_keys), None, nf, ptr(d_q), None, bias,
                        ptr(q_seg), ptr(d_fo), n, ptr(pos), st))
                    found = pos[:n] >= 0
                    vals = ctx.gather_i64(d_cf, torch.clamp(pos[:n], min=0), n)
                    ctx.launches += 2
                    ok = found.cpu().numpy()
                    retro[ok] = vals[:n].cpu().numpy()[ok]
                g.create_dataset(tag + '_counts_final', data=retro)
        if verbose:
            print('Saved final {} counts\n'.format(tag))
