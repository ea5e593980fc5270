/* liborbit_b200 -- C ABI of the B200-native orbit-tracking hot path.
 *
 * The reference (s-balu/nbody-orbit-analysis, package `orbitanalysis` v0.1) is
 * pure Python/numpy and has no FFI of its own: its "operator interface" for
 * this path is the set of per-region numpy functions called from
 * `track_orbits.py:147-217`.  Each entry point below states which of those it
 * replaces (file:line in the reference tree) -- a maintainer binds them with
 * ctypes exactly as INTEGRATION.md shows.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types;
 *   - every `const void* / void*` array argument is a DEVICE pointer unless the
 *     name ends in `_host`;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = default stream);
 *   - all functions return 0 on success or a negative OA_ERR_* code, never
 *     throw, and never synchronise unless documented; the message of the last
 *     failure on the calling thread is available from `oa_last_error()`;
 *   - nothing here falls back to the CPU: without a CUDA device every compute
 *     entry point returns OA_ERR_CUDA.
 */
#ifndef ORBIT_B200_H
#define ORBIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OA_OK 0
#define OA_ERR_ARG (-1)
#define OA_ERR_CUDA (-2)
#define OA_ERR_UNSUPPORTED (-3)

#define OA_F32 0
#define OA_F64 1

#define OA_MODE_PERICENTRIC 0 /* v_r: - -> + (track_orbits.py:311-312) */
#define OA_MODE_APOCENTRIC 1  /* v_r: + -> - (track_orbits.py:313-314) */

/* ABI version; bumped whenever a struct below changes. */
#define OA_ABI_VERSION 18

int oa_abi_version(void);
const char* oa_last_error(void);

/* Device facts used by the host to size grids/buffers: SM count, compute
 * capability, L2 bytes.  Fails (OA_ERR_CUDA) when no device is visible. */
int oa_device_info(int* sm_count, int* cc_major, int* cc_minor,
                   int64_t* l2_bytes, int64_t* hbm_bytes);

/* ---------------------------------------------------------------------------
 * Region table: one row per region block of the CURRENT snapshot.
 * Replaces the per-halo Python arguments of `track(j)` / `region_frame`
 * (track_orbits.py:147-155, 247): region centre, bulk velocity and which block
 * of the previous snapshot holds the same halo (track_orbits.py:162-165).
 * ------------------------------------------------------------------------- */
typedef struct oa_region {
    double centre[3];    /* regions()[0][j]                                   */
    double bulk[3];      /* regions()[2][j], or written by oa_bulk_velocity   */
    int64_t prev_begin;  /* first particle of the halo's previous block, or -1 */
    int64_t prev_count;  /* its length (0 when there is no previous block)    */
    int64_t prev_bucket; /* oa_table_bucket_begin() of that previous block    */
    int64_t cur_bucket;  /* oa_table_bucket_begin() of this block             */
    int64_t cur_begin;   /* first particle of this block (= cur_off[j])       */
    int64_t cur_count;   /* its length                                        */
    float centre_f[3];   /* (float)centre: used when the frame is float32     */
    float bulk_f[3];     /* (float)bulk  (oa_bulk_velocity keeps it in sync)  */
    int64_t reserved;
} oa_region;             /* 128 bytes; the table must be 16-byte aligned      */

/* Host twin of the device routine that restates numpy's pairwise summation
 * (np.sum of a contiguous 1-D float array; used for the mass sum of the derived
 * bulk velocity, track_orbits.py:270-274): HOST pointer, result in *out. */
int oa_pairwise_sum_host(const void* a, int dtype, int64_t n, double* out);

/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault) on `stream`: the small
 * read-backs of a snapshot (reference: the numpy arrays `track` returns,
 * track_orbits.py:186-187, live on the host) without a stream switch in the
 * host framework.  Pinned host memory for a truly asynchronous copy. */
int oa_copy_async(void* dst, const void* src, size_t bytes, void* stream);
/* The same for a few kilobytes, done by a kernel (SM stores into pinned,
 * UVA-mapped host memory) instead of a copy engine: a small read-back that the
 * host waits for must not queue behind the event lists in flight on the DMA
 * engine.  `dst` / `src` / `bytes` multiples of 4. */
int oa_copy_small(void* dst, const void* src, size_t bytes, void* stream);

/* Host-to-host copy by up to `n_threads` threads (HOST pointers, no CUDA call):
 * the staging copy of a loader's PAGEABLE arrays into the pinned ingest ring.
 * The reference hands the loader's arrays to numpy as they are
 * (track_orbits.py:130-131, 234-237); a GPU path has to stage them through
 * pinned memory before the DMA engine can take them, and one core's memcpy
 * (~10 GB/s) is slower than the host-to-device link (~52 GB/s).  Each thread
 * copies one contiguous, 4096-byte aligned part with the C library's memcpy
 * (non-temporal stores for large sizes: no read-for-ownership of the
 * destination).  At least 8 MiB per thread (fewer threads for smaller copies);
 * n_threads <= 1 or less than 16 MiB: plain memcpy on the calling thread.
 * Blocks until the copy is complete. */
int oa_host_copy(void* dst, const void* src, size_t bytes, int n_threads);

/* The region table on the HOST (host pointers, no CUDA call): what the Python
 * driver otherwise assembles with ~20 numpy calls per snapshot.  For region j:
 * centre / bulk from the catalogue arrays (float32 or float64, `bulk` may be NULL
 * when the bulk velocity is derived on the device), the block range from
 * `offsets`, and -- when `halo_ids[j]` is found in the ascending
 * `prev_halo_ids` (track_orbits.py:162-165) -- the previous block and its
 * bucket range.  Also written: buckets_out[j] (oa_table_bucket_begin of this
 * block), matched_out[j] (0/1), prev_index_out[j] (position in prev_halo_ids or
 * -1), seg_begin_out[m] = previous block start of the m-th matched region; the
 * number of matched regions is returned in *n_matched. */
int oa_region_rows_host(int n_regions, const int64_t* offsets,
                        const void* centres, int centre_dtype,
                        const void* bulk, int bulk_dtype,
                        const int64_t* halo_ids, const int64_t* prev_halo_ids,
                        int n_prev_regions, const int64_t* prev_offsets,
                        const int64_t* prev_buckets, oa_region* rows,
                        int64_t* buckets_out, uint8_t* matched_out,
                        int32_t* prev_index_out, int64_t* seg_begin_out,
                        int* n_matched);

/* Bytes of one carried-state record (32 for an OA_F32 frame, 64 for OA_F64). */
size_t oa_record_bytes(int frame_dtype);

/* ID hash table of one snapshot (one uint32 array):
 *   [ fill counters, one uint32 per bucket, padded to a multiple of 8 words |
 *     slot buckets, 8 uint32 slots (one 32-byte sector) per bucket ]
 * slot = fingerprint << index_bits | block-local particle index.  The buckets
 * of region block j (start `block_start`, length len) are
 *   [oa_table_bucket_begin(block_start, j), ... + len/OA_BUCKET_LOAD + 1).
 * Only the counters need clearing between snapshots (oa_table_clear).        */
#define OA_BUCKET_WORDS 8
#define OA_BUCKET_SLOTS 8
#define OA_BUCKET_LOAD 3   /* particles per bucket (mean fill of the 8 slots) */
int64_t oa_table_bucket_begin(int64_t block_start, int64_t region_index);
/* Number of buckets of the table for n region-particles in n_regions. */
int64_t oa_table_buckets(int64_t n, int64_t n_regions);
/* Number of uint32 words of that table (counters + slots). */
int64_t oa_table_slots(int64_t n, int64_t n_regions);
/* Bits needed to store a block-local particle index given the largest block. */
int oa_index_bits(int64_t max_block_len);

/* ---------------------------------------------------------------------------
 * (b0) Derived bulk velocity of every region block.
 * Replaces `np.mean(vel[sl], axis=0)` / the mass-weighted sum of
 * track_orbits.py:267-284 and track_orbits_onthefly.py:96-110.
 * Sums are accumulated in float64 in a fixed (deterministic) order; the mean is
 * rounded to float32 when `round_f32` is set (numpy returns the input dtype).
 * Writes regions[j].bulk and, if not NULL, bulk_out[j*3..] (float64; NaN for an
 * empty block, like numpy's mean of an empty slice).
 * workspace: oa_bulk_workspace_bytes(n, n_regions) bytes.
 * ------------------------------------------------------------------------- */
size_t oa_bulk_workspace_bytes(int64_t n, int n_regions);
int oa_bulk_velocity(const void* vel, int vel_dtype, const void* mass,
                     int mass_dtype, const int64_t* cur_off, int n_regions,
                     int64_t n, int round_f32, oa_region* regions,
                     double* bulk_out, void* workspace, size_t workspace_bytes,
                     void* stream);

/* ---------------------------------------------------------------------------
 * (a)+(b) Fused per-snapshot tracking kernel.
 *
 * For every particle of the current snapshot, in block order, ONE kernel:
 *   1. halo frame: x' = wrap(x - centre), r, rhat, v_r          region_frame
 *      (track_orbits.py:247-290, utils.py:24-33; on-the-fly variant
 *      track_orbits_onthefly.py:71-120 when `onthefly` is set);
 *   2. finds the particle's record in the same halo's previous block by probing
 *      that block's ID hash table                     compare_radial_velocities
 *      (track_orbits.py:300-309 = setdiff1d/in1d/delete + utils.myin1d :4-11);
 *   3. sign-flip test, angle change arccos(rhat_prev . rhat), float16 angle
 *      accumulator update/reset     (track_orbits.py:311-325, calc_angles :330-351);
 *   4. writes the new 32/64-byte record, the event mark of the PREVIOUS
 *      particle (so that events can be emitted in previous-block order,
 *      track_orbits.py:315-316) and inserts its own ID into the current
 *      snapshot's hash table for the next call.
 *
 * dtypes: `data_dtype` is the dtype of pos/vel; `frame_dtype` the dtype numpy
 * would give the halo-frame coordinates (result_type(pos, centre); on-the-fly:
 * the dtype of pos); `centre_f32` / `bulk_f32` say whether the catalogue rows
 * were float32 (they decide where numpy rounds to float32).  v_r is float64
 * (track_orbits.py:275-288 under numpy>=2) unless `onthefly`.
 * ------------------------------------------------------------------------- */
typedef struct oa_track_args {
    /* current snapshot (inputs) */
    const void* pos;        /* (n_cur,3) data_dtype                           */
    const void* vel;        /* (n_cur,3) data_dtype                           */
    const int64_t* ids;     /* (n_cur,)                                       */
    int64_t n_cur;
    const int64_t* cur_off; /* (n_regions+1,) block starts + n_cur            */
    const oa_region* regions; /* (n_regions,)                                 */
    int32_t n_regions;
    int32_t data_dtype;     /* OA_F32 / OA_F64                                */
    int32_t frame_dtype;    /* OA_F32 / OA_F64                                */
    int32_t centre_f32;
    int32_t bulk_f32;
    int32_t periodic;       /* 'box_size' in snapshot                         */
    int32_t onthefly;       /* on-the-fly arithmetic + per-match outputs      */
    int32_t mode;           /* OA_MODE_*                                      */
    double box[3];
    double hubble;          /* H(z), utils.py:36-39 (host scalar)             */
    double one_plus_z;
    /* previous generation (read) */
    const void* rec_prev;   /* (n_prev,) records, NULL if none                */
    const uint32_t* tab_prev; /* table filled by the previous call             */
    int64_t tab_prev_buckets; /* oa_table_buckets() of that table              */
    int64_t n_prev;
    int32_t prev_index_bits;
    int32_t cur_index_bits;
    uint16_t* mark_prev;    /* (n_prev,) event marks, updated                 */
    /* current generation (written) */
    void* rec_cur;          /* (n_cur,) records                               */
    uint32_t* tab_cur;      /* oa_table_slots() words, oa_table_clear()ed      */
    int64_t tab_cur_buckets; /* oa_table_buckets(n_cur, n_regions)             */
    uint16_t* mark_cur;     /* (n_cur,) written: "no event"                   */
    void* workspace;        /* oa_track_workspace_bytes(n_cur) bytes          */
    size_t workspace_bytes;
    /* The kernel is persistent, one CTA per SM.  sm_reserve > 0 leaves that many
     * SMs free so that kernels of other streams (the NCCL event exchange of the
     * previous snapshot) run concurrently instead of after it. */
    int32_t sm_reserve;
    int32_t reserved0;
    /* optional per-particle outputs, NULL to skip */
    void* out_rhat;         /* (n_cur,3) frame_dtype                          */
    void* out_vr;           /* (n_cur,) float64 (float32 if onthefly && F32)  */
    void* out_r;            /* (n_cur,) frame_dtype                           */
    uint16_t* out_angle;    /* (n_cur,) float16 bits: checkpoint 'angles'     */
    int64_t* out_match;     /* (n_cur,) previous index or -1                  */
    void* dangle_prev;      /* (n_prev,) frame_dtype, on-the-fly 'angles'     */
} oa_track_args;

int oa_track_fused(const oa_track_args* args, void* stream);
size_t oa_track_workspace_bytes(int64_t n_cur);
/* Zero the fill counters of a table of oa_table_slots(n, n_regions) words.
 * Must precede the oa_track_fused call that fills `tab_cur` (kept separate so
 * that the fused kernel can be timed on its own). */
int oa_table_clear(uint32_t* tab, int64_t n, int64_t n_regions, void* stream);
/* sizeof(oa_track_args) as compiled -- lets a binding verify its struct mirror. */
size_t oa_track_args_size(void);

/* ---------------------------------------------------------------------------
 * (a)+(b), second implementation: region-at-a-time PARTITIONED HASH JOIN.
 *
 * Same reference code as oa_track_fused (track(j), track_orbits.py:147-185:
 * region_frame :247-290, compare_radial_velocities :293-327, calc_angles
 * :330-351), for the float32-data / float32-frame / float64-v_r tracking mode
 * (not on-the-fly, no per-particle diagnostic outputs).  The carried records of
 * a region are stored partitioned by ID hash (2^bits partitions of at most
 * ~OA_PJOIN_TARGET particles), so that the previous partition fits a
 * shared-memory hash table and every global access is a stream; one persistent
 * kernel runs COUNT / SCAN / SCATTER / JOIN items in ticket order with
 * per-region dependency counters (csrc/oa_pjoin_core.cuh).
 *
 * Host-built plan (one row per region, plus a sentinel row with the totals):
 *   bits_cur   = max(bits of the halo's previous block, smallest b with
 *                cur_count <= OA_PJOIN_TARGET << b), at most OA_PJOIN_MAX_BITS;
 *   pb_cur     = exclusive prefix over regions of (1 << bits_cur) + 1: the
 *                region's entries in part_off_cur (partition starts + end);
 *   tile_first = exclusive prefix of SCATTER tiles
 *                (bits_cur > 0 ? ceil(cur_count / OA_PJOIN_TILE) : 0);
 *   count_first = the same for COUNT tiles of OA_PJOIN_CTILE particles;
 *   scan_first = exclusive prefix of (bits_cur > 0);
 *   join_first = exclusive prefix of JOIN items (bits_cur > 0:
 *                bits_prev >= 0 ? 1 << bits_prev : 0; bits_cur == 0: 1 for the
 *                first region of a PACK, 0 for its other members);
 *   pack_len   = consecutive small regions (bits_cur == 0, same group) handled by
 *                one item: at most OA_PJOIN_PACK_MAX regions, OA_PJOIN_TILE
 *                particles and OA_PJOIN_REC_CAP previous records together -- so
 *                that a catalogue of many tiny halos does not cost one CTA-wide
 *                item per halo.  A region that exceeds these limits alone is a
 *                pack of one.
 * Regions are grouped (group_first: first region of each group + n_regions);
 * range_start[4 * s + stage] = first ticket of stage `stage` (0 JOIN, 1
 * SCATTER, 2 SCAN, 3 COUNT) in superstep s, which holds the items of group
 * s - 3 / s - 2 / s - 1 / s respectively (n_ranges = 4 * (n_groups + 3),
 * plus one end entry).  A small pre-kernel expands the plan into the explicit
 * item list (8 bytes per ticket, in the workspace) that the CTAs index by ticket.
 * ------------------------------------------------------------------------- */
/* Compile-time shape of the kernel (tuning builds override them with -D; the
 * host reads them back with oa_pjoin_config, nothing else hard-codes them). */
#ifndef OA_PJOIN_THREADS
#define OA_PJOIN_THREADS 512   /* threads per CTA                              */
#endif
#ifndef OA_PJOIN_MIN_CTAS
#define OA_PJOIN_MIN_CTAS 2    /* resident CTAs per SM the kernel is built for */
#endif
#ifndef OA_PJOIN_TILE
#define OA_PJOIN_TILE 2048     /* particles per SCATTER item                   */
#endif
#ifndef OA_PJOIN_CTILE
#define OA_PJOIN_CTILE 8192    /* particles per COUNT item                     */
#endif
#ifndef OA_PJOIN_REC_CAP
#define OA_PJOIN_REC_CAP 2944  /* previous records per shared-memory table     */
#endif
#ifndef OA_PJOIN_TARGET
#define OA_PJOIN_TARGET 2304   /* largest mean partition size                  */
#endif
#ifndef OA_PJOIN_TMA
#define OA_PJOIN_TMA 0         /* 1: table side of a JOIN loaded by one TMA bulk copy */
#endif
#define OA_PJOIN_MAX_BITS 12
#define OA_PJOIN_PACK_MAX 32   /* regions per pack of small regions            */

typedef struct oa_pjoin_region {
    uint32_t pb_cur;
    uint32_t pb_prev;     /* the halo's entries in part_off_prev               */
    int32_t bits_cur;
    int32_t bits_prev;    /* -1: the halo has no previous block                */
    uint32_t tile_first;
    uint32_t join_first;
    uint32_t scan_first;
    uint32_t count_first;
    uint32_t pack_len;    /* bits_cur == 0: regions of the pack that starts here */
    uint32_t reserved[3]; /*               (0: member of an earlier pack)        */
} oa_pjoin_region;        /* 48 bytes */

typedef struct oa_pjoin_args {
    const float* pos;       /* (n_cur,3)                                      */
    const float* vel;       /* (n_cur,3)                                      */
    const int64_t* ids;     /* (n_cur,)                                       */
    int64_t n_cur;
    const oa_region* regions;      /* (n_regions,) centre, bulk, block ranges  */
    const oa_pjoin_region* plan;   /* (n_regions + 1,)                         */
    const uint32_t* group_first;   /* (n_groups + 1,)                          */
    const uint32_t* range_start;   /* (n_ranges + 1,)                          */
    int32_t n_regions;
    int32_t n_groups;
    int32_t n_ranges;
    int32_t centre_f32;
    int32_t bulk_f32;
    int32_t periodic;
    int32_t mode;           /* OA_MODE_*                                      */
    int32_t hubble_on;
    double box[3];
    double hubble;
    double one_plus_z;
    /* previous generation */
    const void* rec_prev;          /* (n_prev,) 32-byte records, partitioned   */
    const uint32_t* part_off_prev; /* partition offsets written by that call   */
    uint16_t* mark_prev;           /* (n_prev,) event marks in BLOCK order     */
    int64_t n_prev;
    /* current generation (written) */
    void* rec_cur;
    uint32_t* part_off_cur;        /* (n_part_entries,)                        */
    uint16_t* mark_cur;            /* (n_cur,) written: "no event"             */
    void* workspace;               /* oa_pjoin_workspace_bytes(), any content  */
    size_t workspace_bytes;
    int64_t n_part_entries;        /* plan[n_regions].pb_cur                   */
    int32_t sm_reserve;
    uint32_t total_tickets;        /* range_start[n_ranges] (host copy)        */
} oa_pjoin_args;

/* The plan on the HOST (all pointers are host pointers; no CUDA call): fills
 * rows[n_regions + 1], bits_out / pb_out [n_regions] (kept by the caller for the
 * next snapshot), group_first (capacity n_regions + 1) and range_start (capacity
 * 4 * (n_regions + 3) + 1).  prev_bits[j] / prev_pb[j] / prev_counts[j]: partition
 * bits, first part_off entry and length of the same halo's previous block
 * (bits -1 = none).  Groups are
 * runs of regions whose first particle lies in the same window of
 * `lag_particles` particles. */
typedef struct oa_pjoin_plan_info {
    int64_t n_part_entries;
    uint32_t total_tickets;
    int32_t n_groups;
    int32_t n_ranges;
    int32_t max_bits;
} oa_pjoin_plan_info;
int oa_pjoin_plan_host(const int64_t* offsets, int n_regions,
                       const int32_t* prev_bits, const int64_t* prev_pb,
                       const int64_t* prev_counts,
                       int64_t target, int64_t lag_particles,
                       oa_pjoin_region* rows, int32_t* bits_out, int64_t* pb_out,
                       uint32_t* group_first, uint32_t* range_start,
                       oa_pjoin_plan_info* info);
size_t oa_pjoin_workspace_bytes(int n_regions, int64_t n_part_entries,
                                uint32_t total_tickets);
/* out[0..6] = THREADS, MIN_CTAS, TILE, CTILE, REC_CAP, TARGET, MAX_BITS,
 * out[7] = dynamic shared memory per CTA in bytes. */
void oa_pjoin_config(int32_t* out8);
size_t oa_pjoin_args_size(void);
/* Profiling builds (-DOA_PJOIN_STATS=1) only: out16 = [cycles per stage JOIN,
 * SCATTER, SCAN, COUNT | cycles waiting for dependencies | ... | items per stage
 * at [8..11] | CTA cycles at [12]]; returns 0 when built without the counters. */
int oa_pjoin_stats(uint64_t* out16, int reset);
int oa_pjoin_step(const oa_pjoin_args* args, void* stream);

/* ---------------------------------------------------------------------------
 * Partitioned hash join, second generation (oa_pj2_step, csrc/oa_pj2.cu): the
 * same step as oa_pjoin_step -- track(j) = region_frame +
 * compare_radial_velocities + calc_angles (track_orbits.py:147-185, 247-351;
 * utils.py:4-33) for float32 data and a float32 catalogue -- built around what
 * the first generation's profile showed (profiles/r02_c1_*): it was bound by
 * CTA-wide barriers and exposed load latency, not by HBM.
 *
 *   - Partitions have a FIXED capacity, so there is no COUNT / SCAN pass: region
 *     j owns P_j partitions (a power of two) of cap_j record slots each;
 *     partition p = record slots [base_j + p cap_j, ... + fill[pb_j + p]).  A
 *     particle's partition is the top log2(P_j) bits of hash32(ID).  P_j never
 *     shrinks for a halo, so cur partition p joins prev partition p >> shift.
 *     cap_j = mean + OA_PJ2_SIGMAS sqrt(mean) + 16, mean = ceil(len_j / P_j); a
 *     partition that still overflows drops the record and counts it in
 *     `overflow` (the snapshot's result is then invalid; the host raises).
 *   - Two kinds of work items: SCATTER (tile of OA_PJ2_TILE consecutive
 *     particles of the snapshot, whatever regions they belong to: halo frame,
 *     32 B record -> its partition, slot taken with one atomic on the fill
 *     word) and JOIN (cur partition p of region j against prev partition
 *     p >> shift in a shared-memory table: sign test, arccos, float16
 *     accumulator written back into the record, event mark at the previous
 *     particle's block position).  Items are taken by ticket; the ticket order
 *     interleaves SCATTER of group g with JOIN of group g - OA_PJ2_LAG (groups =
 *     runs of regions of ~group_particles particles), so that new records are
 *     re-read from L2.  A JOIN waits for its group's tile counter.
 *   - Every CTA has a PRODUCER warp: it takes tickets, waits for the item's
 *     dependency and moves the item's inputs (tile: IDs, positions, velocities;
 *     join: both partitions) into one of two shared-memory stages with TMA bulk
 *     copies (cp.async.bulk + mbarrier complete_tx) while the consumer warps
 *     work on the other stage.  Consumers never wait for global loads and
 *     synchronise among themselves with one named barrier per JOIN.
 * ------------------------------------------------------------------------- */
#ifndef OA_PJ2_THREADS
#define OA_PJ2_THREADS 480     /* CONSUMER threads per CTA (+ one producer warp) */
#endif
#ifndef OA_PJ2_MIN_CTAS
#define OA_PJ2_MIN_CTAS 2
#endif
#ifndef OA_PJ2_TILE
#define OA_PJ2_TILE 1408       /* particles per SCATTER item (multiple of 4)   */
#endif
#ifndef OA_PJ2_CAP
#define OA_PJ2_CAP 704         /* record slots per partition at most (<= 1024) */
#endif
#ifndef OA_PJ2_TARGET
#define OA_PJ2_TARGET 416      /* mean records per partition at most           */
#endif
#ifndef OA_PJ2_SIGMAS
#define OA_PJ2_SIGMAS 8        /* head room of a partition in Poisson sigmas   */
#endif
#ifndef OA_PJ2_LAG
#define OA_PJ2_LAG 1           /* groups between a region's SCATTER and its JOIN */
#endif

typedef struct oa_pj2_region {
    uint32_t P_cur;       /* partitions of this block (power of two, >= 1)     */
    uint32_t cap_cur;     /* record slots per partition                        */
    uint32_t base_cur;    /* first record slot of partition 0                  */
    uint32_t pb_cur;      /* entry of partition 0 in the fill array            */
    uint32_t P_prev;      /* 0: the halo has no previous block                 */
    uint32_t cap_prev;
    uint32_t base_prev;
    uint32_t pb_prev;
    uint32_t shift;       /* P_cur = P_prev << shift                           */
    uint32_t group;       /* group of regions this block belongs to            */
    uint32_t join_first;  /* JOIN items before this region (P_cur each when P_prev > 0) */
    uint32_t reserved;
} oa_pj2_region;          /* 48 bytes */

typedef struct oa_pj2_args {
    const float* pos;       /* (n_cur,3), 16-byte aligned (TMA)                */
    const float* vel;
    const int64_t* ids;
    int64_t n_cur;
    const oa_region* regions;      /* (n_regions,)                             */
    const oa_pj2_region* plan;     /* (n_regions + 1,)                         */
    const uint32_t* group_first;   /* (n_groups + 1,) first region of a group  */
    const uint32_t* group_off;     /* (n_groups + 1,) first particle of a group */
    const uint32_t* range_start;   /* (n_ranges + 1,) first ticket of a range  */
    int32_t n_regions;
    int32_t n_groups;
    int32_t n_ranges;              /* 2 * (n_groups + OA_PJ2_LAG): range 2 s = SCATTER of group s, 2 s + 1 = JOIN of group s - LAG */
    int32_t centre_f32;            /* must be 1 (float32 catalogue)            */
    int32_t bulk_f32;
    int32_t periodic;
    int32_t mode;
    int32_t hubble_on;
    double box[3];
    double hubble;
    double one_plus_z;
    const void* rec_prev;          /* previous generation: record slots,       */
    const uint32_t* fill_prev;     /*   fill per partition,                    */
    uint16_t* mark_prev;           /*   event marks in BLOCK order (n_prev,)   */
    int64_t n_prev;
    void* rec_cur;                 /* (n_rec_slots,) 32-byte records           */
    uint32_t* fill_cur;            /* (n_part_entries,) zeroed by the call     */
    uint16_t* mark_cur;            /* (n_cur,) written: "no event"             */
    void* workspace;               /* oa_pj2_workspace_bytes()                 */
    size_t workspace_bytes;
    int64_t n_part_entries;
    int64_t n_rec_slots;
    int32_t sm_reserve;
    uint32_t total_tickets;
    uint32_t* overflow;            /* device word, += 1 per dropped record (never cleared here) */
} oa_pj2_args;

typedef struct oa_pj2_plan_info {
    int64_t n_part_entries;
    int64_t n_rec_slots;
    uint32_t total_tickets;
    int32_t n_groups;
    int32_t n_ranges;
    uint32_t max_P;
} oa_pj2_plan_info;
/* The plan on the HOST (host pointers, no CUDA call).  Per region of THIS
 * snapshot: prev_P / prev_cap / prev_base / prev_pb of the same halo's previous
 * block (prev_P 0 = none or empty).  Writes rows[n_regions + 1], the carried
 * P_out / cap_out / base_out / pb_out [n_regions], group_first and group_off
 * (capacity n_regions + 1 each), range_start (capacity 2 * (n_regions +
 * OA_PJ2_LAG) + 1).  `target` <= 0: OA_PJ2_TARGET. */
int oa_pj2_plan_host(const int64_t* offsets, int n_regions, const uint32_t* prev_P,
                     const uint32_t* prev_cap, const uint32_t* prev_base,
                     const uint32_t* prev_pb, int64_t group_particles, int32_t target,
                     oa_pj2_region* rows, uint32_t* P_out, uint32_t* cap_out,
                     uint32_t* base_out, uint32_t* pb_out, uint32_t* group_first,
                     uint32_t* group_off, uint32_t* range_start,
                     oa_pj2_plan_info* info);
size_t oa_pj2_workspace_bytes(int n_groups, uint32_t total_tickets);
/* out[0..7] = THREADS (consumers), MIN_CTAS, TILE, CAP, TARGET, SIGMAS, LAG,
 * dynamic shared memory per CTA in bytes. */
void oa_pj2_config(int32_t* out8);
size_t oa_pj2_args_size(void);
/* Profiling builds (-DOA_PJ2_STATS=1): out16 = [consumer cycles SCATTER, JOIN |
 * producer cycles waiting for a dependency, for a free stage | consumer cycles
 * waiting for a full stage | ... | items at [8..9] | CTA cycles at [12]];
 * returns 0 without the counters. */
int oa_pj2_stats(uint64_t* out16, int reset);
int oa_pj2_step(const oa_pj2_args* args, void* stream);

/* ---------------------------------------------------------------------------
 * Ordered selection ("np.argwhere(cond).flatten()", track_orbits.py:315 and
 * the result assembly :199-217): positions i (ascending) of marks that satisfy
 * a predicate, plus per-segment offsets.
 *   OA_SEL_NE : marks[i] != value      OA_SEL_EQ : marks[i] == value
 * Two steps so that the host can size the output exactly:
 *   oa_select_count  -> writes tile counts to workspace and the total to
 *                       *total_dev (device int64);
 *   oa_select_gather -> writes the selected positions.
 * ------------------------------------------------------------------------- */
#define OA_SEL_NE 0
#define OA_SEL_EQ 1
size_t oa_select_workspace_bytes(int64_t n);
int oa_select_count(const uint16_t* marks, int64_t n, int op, uint16_t value,
                    void* workspace, size_t workspace_bytes,
                    int64_t* total_dev, void* stream);
int oa_select_gather(const uint16_t* marks, int64_t n, int op, uint16_t value,
                     const void* workspace, int64_t* sel_out, void* stream);
/* Event lists in one pass (after oa_select_count with OA_SEL_NE, OA_NO_EVENT):
 * positions, IDs (from the previous records) and float16 angles (the marks
 * themselves) of all events, in previous-block order: apsis_inds / apsis_ids /
 * apsis_angles of track_orbits.py:315-316, 347. */
int oa_select_gather_events(const uint16_t* marks, int64_t n,
                            const void* workspace, const void* rec,
                            int frame_dtype, int64_t* sel_out, int64_t* ids_out,
                            uint16_t* angles_out, void* stream);
/* Same, with the IDs taken from the previous snapshot's ID array (block order)
 * instead of its records -- for generations whose records are not in block
 * order (oa_pjoin_step). */
int oa_select_gather_events_ids(const uint16_t* marks, int64_t n,
                                const void* workspace, const int64_t* ids,
                                int64_t* sel_out, int64_t* ids_out,
                                uint16_t* angles_out, void* stream);
/* Consumers of a selection.  `n_sel` is the number of selected positions; when
 * `n_dev` is not NULL it points to the exact count ON THE DEVICE (as written by
 * oa_select_count) and `n_sel` is only an upper bound used to size the launch,
 * so a whole snapshot can be enqueued without a host synchronisation.
 *
 * offsets_out[k] = number of selected positions < seg_begin[k]   (k < n_seg);
 * the caller appends the total.  np.cumsum([0]+lens), track_orbits.py:214. */
int oa_segment_offsets(const int64_t* sel, int64_t n_sel, const int64_t* n_dev,
                       const int64_t* seg_begin, int n_seg,
                       int64_t* offsets_out, void* stream);

/* Gathers driven by a selection. */
int oa_gather_record_ids(const void* rec, int frame_dtype, const int64_t* sel,
                         int64_t n_sel, const int64_t* n_dev, int64_t* ids_out,
                         void* stream);
int oa_gather_u16(const uint16_t* src, const int64_t* sel, int64_t n_sel,
                  const int64_t* n_dev, uint16_t* out, void* stream);
int oa_gather_i64(const int64_t* src, const int64_t* sel, int64_t n_sel,
                  const int64_t* n_dev, int64_t* out, void* stream);
int oa_gather_f(const void* src, int dtype, const int64_t* sel, int64_t n_sel,
                const int64_t* n_dev, void* out, void* stream);
/* marks[i] = (match[i] < 0) ? 1 : 0  -- "entered" predicate (onthefly :168) */
int oa_mark_unmatched(const int64_t* match, int64_t n, uint16_t* marks,
                      void* stream);
int oa_fill_u16(uint16_t* dst, int64_t n, uint16_t value, void* stream);
/* Overwrite the float16 angle accumulator of n records: resume from the
 * reference's '.checkpoint' dataset (track_orbits.py:229-232). */
int oa_set_record_angles(void* rec, int frame_dtype, const uint16_t* angles,
                         int64_t n, void* stream);

/* ---------------------------------------------------------------------------
 * LSD radix sort of (uint64 key, uint64 value) pairs on bits
 * [begin_bit, end_bit), stable.  Replaces the argsort/unique calls of
 * utils.py:10, track_orbits_onthefly.py:145,168 (sorted entered/departed
 * lists), progenitors.py:52,82,108 and postprocessing.py:135.
 * Result is left in keys_out/vals_out.  workspace: oa_sort_workspace_bytes(n).
 * ------------------------------------------------------------------------- */
size_t oa_sort_workspace_bytes(int64_t n);
int oa_sort_pairs_u64(const uint64_t* keys_in, const uint64_t* vals_in,
                      uint64_t* keys_out, uint64_t* vals_out, int64_t n,
                      int begin_bit, int end_bit, void* workspace,
                      size_t workspace_bytes, void* stream);
/* Sort keys for per-segment sorting of an ID list laid out in segments
 * [seg_off[s], seg_off[s+1]): key_hi = segment index; key_lo = id - id_minmax[0]
 * for segments whose sort_flag is non-zero (all, if sort_flag is NULL), else the
 * element's own position (keeps the original order, onthefly :176-177);
 * index[i] = i.  Two stable oa_sort_pairs_u64 passes (key_lo, then key_hi)
 * give "sorted within every segment". */
int oa_segment_sort_keys(const int64_t* ids, int64_t n, const int64_t* seg_off,
                         int n_seg, const uint8_t* sort_flag,
                         const int64_t* id_minmax, uint64_t* key_lo,
                         uint64_t* key_hi, uint64_t* index, void* stream);
/* head[i] = 1 where a new (segment, id) run starts in a sorted sequence;
 * np.unique(return_counts=True) per halo, postprocessing.py:133-141. */
int oa_run_heads(const uint64_t* seg, const int64_t* ids, int64_t n,
                 uint16_t* head, void* stream);
/* counts[k] = length of run k given the ascending run start positions. */
int oa_run_lengths(const int64_t* starts, int64_t n_runs, int64_t n,
                   int64_t* counts, void* stream);
/* Multi-GPU event ordering: merge n_lists event lists, each
 * ascending in `keys` (= position of the event particle in the UNSHARDED
 * previous snapshot, distinct across lists) and stored back to back
 * (list r = [list_off[r], list_off[r+1])), into one list in key order -- the
 * reference's event order, track_orbits.py:315-316. */
int oa_merge_event_lists(const int64_t* keys, const int64_t* ids,
                         const uint16_t* angles, int64_t n,
                         const int64_t* list_off, int n_lists, int64_t* ids_out,
                         uint16_t* angles_out, void* stream);
/* The same exchange without host round trips (every rank, every snapshot):
 *   oa_pack_events    -> this rank's send buffer: true event count, per-halo
 *                        counts, then up to `cap` (key, ID, angle) records;
 *                        `small` = [offsets[n_seg] | total] as left on the device
 *                        by oa_segment_offsets / oa_select_count;
 *   (one NCCL all-gather of oa_exchange_bytes(n_seg, cap) bytes per rank)
 *   oa_merge_gathered -> merged ID / angle lists in key order and
 *                        info = [total | global offsets[n_seg+1] | sizes[world] |
 *                        overflow flag (some rank had more than `cap` events)].
 * gpos / sel / ids / angles may be NULL when the local event list is empty
 * (the count is read from `small` on the device); same for oa_split_quantiles
 * and oa_pack_split. */
size_t oa_exchange_bytes(int n_seg, int64_t cap);
int oa_pack_events(const int64_t* gpos, const int64_t* sel, const int64_t* ids,
                   const uint16_t* angles, const int64_t* small, int n_seg,
                   int64_t cap, void* out, void* stream);
int oa_merge_gathered(const void* gathered, int world, int n_seg, int64_t cap,
                      int64_t* ids_out, uint16_t* angles_out, int64_t* info,
                      void* stream);
/* Scalable exchange (volume per rank independent of the number of GPUs): the
 * key space is cut into `world` ranges at quantiles of the order key; rank q
 * receives range q from everyone (all-to-all), merges and owns that contiguous
 * slice of the global lists.
 *   oa_split_quantiles : q_out[world-1] = this rank's quantile keys (every
 *                        rank's events are a uniform sample of all events);
 *   (all-gather of the proposals)
 *   oa_pack_split      : splitters = per-quantile median of the proposals;
 *                        send buffer = `world` blocks of oa_exchange_bytes(0, cap)
 *                        bytes; counts[n_seg] = this rank's per-halo counts;
 *                        bnd_ws: world+1 int64;
 *   (all-to-all of the blocks; the per-halo counts are summed over the ranks)
 *   oa_merge_blocks    : this rank's slice in key order, info = [size | largest
 *                        received block before truncation (> cap: repeat)]. */
int oa_split_quantiles(const int64_t* gpos, const int64_t* sel, const int64_t* small,
                       int n_seg, int world, int64_t* q_out, void* stream);
int oa_pack_split(const int64_t* gpos, const int64_t* sel, const int64_t* ids,
                  const uint16_t* angles, const int64_t* small, int n_seg,
                  const int64_t* proposals, int world, int64_t cap,
                  int64_t* bnd_ws, void* out, int64_t* counts, void* stream);
int oa_merge_blocks(const void* recv, int world, int64_t cap, int64_t* ids_out,
                    uint16_t* angles_out, int64_t* info, void* stream);
/* Batched exchange (several snapshots per exchange): append the local events of
 * one snapshot to a staging area -- keys = gpos[sel[i]] | tag (tag = snapshot
 * slot << 58), IDs, angles at [ev_base, ev_base + n_local); its n_seg segment
 * offsets, shifted by ev_base, at out_small[0..n_seg) and the running total at
 * out_small[n_seg].  The staged arrays of a batch are exchanged like ONE
 * snapshot with the concatenated segments (keys order snapshot-major). */
int oa_stage_events(const int64_t* gpos, const int64_t* sel, const int64_t* ids,
                    const uint16_t* angles, const int64_t* small, int n_seg,
                    int64_t n_local, int64_t tag, int64_t ev_base,
                    int64_t* out_keys, int64_t* out_ids, uint16_t* out_angles,
                    int64_t* out_small, void* stream);
/* min and max of an int64 array -> out_dev[0], out_dev[1] (device). */
int oa_minmax_i64(const int64_t* x, int64_t n, int64_t* out_dev, void* stream);

/* ---------------------------------------------------------------------------
 * Consumers of the tracking path: progenitors.py and postprocessing.py.
 * All joins are binary searches into lists sorted with oa_sort_pairs_u64.
 * ------------------------------------------------------------------------- */
/* r_out[i] = |wrap(pos[i] - centre of i's region)| in float64 with numpy's
 * rounding points: get_central_particle_ids, progenitors.py:41-51.
 * `centres` is (n_regions,3) float64 on the device, `box_host` 3 doubles on the
 * HOST (ignored unless periodic). */
int oa_central_radii(const void* pos, int data_dtype, const int64_t* cur_off,
                     int n_regions, const double* centres, int centre_f32,
                     int periodic, const double* box_host, int64_t n,
                     double* r_out, void* stream);
/* out[q] = src[order[seg_off[j] + q - out_off[j]]] for q in output segment j:
 * the first entries of every sorted segment (argsort(...)[:n], :52-53). */
int oa_segment_heads(const int64_t* src, const uint64_t* order,
                     const int64_t* seg_off, const int64_t* out_off, int n_seg,
                     int64_t n_out, int64_t* out, void* stream);
/* flags[order[i]] = head[i] (first occurrences back in original order:
 * np.unique(return_index=True), progenitors.py:82-84). */
int oa_scatter_flags(const uint16_t* head, const uint64_t* order, int64_t n,
                     uint16_t* flags, void* stream);
/* pos_out[i] = vals[k] (or k when vals is NULL) with keys[k] == query[i]-bias,
 * -1 if absent or flags[i] == 0; keys ascending and unique.  With q_seg/key_off
 * the search is confined to segment q_seg[i] of the keys (-1: skip).
 * np.in1d + utils.myin1d: progenitors.py:95-99, postprocessing.py:222-232. */
int oa_lookup_sorted(const uint64_t* keys, const uint64_t* vals, int64_t n_keys,
                     const int64_t* query, const uint16_t* flags, int64_t bias,
                     const int32_t* q_seg, const int64_t* key_off, int64_t m,
                     int64_t* pos_out, void* stream);
/* keys[i] = descendant(i) << 32 | halo(where[i]), ~0 when where[i] < 0
 * (progenitors.py:92-106); oa_vote_reduce on the SORTED keys gives the
 * plurality halo per descendant, ties to the smallest index, -1 if none
 * (:107-115).  best_ws: n_desc uint64. */
int oa_vote_keys(const int64_t* where, const int64_t* halo_off, int n_halos,
                 const int64_t* tracked_off, int n_desc, int64_t m,
                 uint64_t* keys, void* stream);
int oa_vote_reduce(const uint64_t* sorted_keys, int64_t m, int n_desc,
                   uint64_t* best_ws, int64_t* out, void* stream);
/* marks[i] = float16(angles[i]) > cut  (postprocessing.py:124-127). */
int oa_angle_cut(const uint16_t* angles, int64_t n, double cut, uint16_t* marks,
                 void* stream);
/* Region extraction, the loader step in front of the tracking path
 * (example_script.py:36-67: per halo, np.argwhere(|recenter(x - c)| < R) over all
 * particles).  A uniform grid (grid_lo / grid_inv_cell / grid_dim; cell index =
 * floor((x - lo) inv_cell), wrapped when periodic) holds per cell the regions
 * whose sphere touches it (CSR cell_start / cell_regions, built on the host);
 * one thread per particle tests the regions of its cell with numpy's arithmetic
 * (frame_dtype = promoted dtype of `coordinates - position`; utils.py:13-33) and
 * emits key = region << 32 | particle.  keys == NULL counts only; *counter
 * (device) receives the number of pairs.  Sorting the keys (oa_sort_pairs_u64)
 * gives the reference layout; oa_gather_by_key gathers rows of 3 (`rows3`) or
 * scalars of elem_bytes in {4, 8} by the particle index in the low 32 bits. */
int oa_region_pairs(const void* pos, int data_dtype, int64_t n, const double* centres,
                    const float* centres_f, const double* radii, int frame_dtype,
                    const int32_t* cell_start, const int32_t* cell_regions,
                    const double* grid_lo, const double* grid_inv_cell,
                    const int32_t* grid_dim, const double* box, int periodic,
                    uint64_t* keys, int64_t capacity, uint64_t* counter, void* stream);
int oa_gather_by_key(const void* src, int elem_bytes, int rows3, const uint64_t* keys,
                     int64_t n, void* out, void* stream);

/* Incremental collation (postprocessing.py:121-141: the reference re-runs
 * np.unique over a halo's WHOLE event history at every snapshot).  The collated
 * state is a table (pool, ID, count) ascending in (pool, ID); a snapshot's new
 * events, sorted and run-length encoded the same way, are merged into it:
 *   oa_merge_find : lb[k] = lower bound of new key k in the table; an equal key
 *                   adds new_cnt[k] to tab_cnt[lb[k]] (miss[k] = 0), else miss[k] = 1;
 *   oa_merge_place: with miss_sel = ascending positions of the misses, writes the
 *                   merged table of n_tab + n_miss rows. */
int oa_merge_find(const int64_t* tab_seg, const int64_t* tab_ids, int64_t* tab_cnt,
                  int64_t n_tab, const int64_t* new_seg, const int64_t* new_ids,
                  const int64_t* new_cnt, int64_t n_new, int64_t* lb, uint16_t* miss,
                  void* stream);
int oa_merge_place(const int64_t* tab_seg, const int64_t* tab_ids, const int64_t* tab_cnt,
                   int64_t n_tab, const int64_t* new_seg, const int64_t* new_ids,
                   const int64_t* new_cnt, const int64_t* lb, const int64_t* miss_sel,
                   int64_t n_miss, int64_t* out_seg, int64_t* out_ids, int64_t* out_cnt,
                   void* stream);
/* seg_out[i] = table[segment of i] (segment index itself if table is NULL). */
int oa_expand_segments(const int64_t* seg_off, int n_seg, const int32_t* table,
                       int64_t n, int32_t* seg_out, void* stream);

/* ---------------------------------------------------------------------------
 * Benchmark workload generator (not a reference function):
 * rosette orbits about drifting halo centres, generated directly in HBM.
 *   oa_synth_keys : per universe particle, key = (halo << 40 | shuffle hash)
 *                   if it lies inside its region at time t, else a sentinel
 *                   that sorts last; counts present particles per halo.
 *   oa_synth_fill : after sorting (key, particle) pairs with oa_sort_pairs_u64,
 *                   writes ids / positions / velocities in block order.
 * ------------------------------------------------------------------------- */
typedef struct oa_synth_params {
    uint64_t seed;
    int64_t n_universe;
    int64_t id_stride, id_offset;   /* id = particle * stride + offset        */
    const int64_t* halo_start;      /* (n_halos+1,) universe ranges           */
    const double* halo_radius;      /* (n_halos,)                             */
    const double* halo_c0;          /* (n_halos,3) centre at t = 0            */
    const double* halo_vh;          /* (n_halos,3) drift = bulk velocity      */
    double box;
    double t;                       /* snapshot index (time in snapshots)     */
    int32_t n_halos;
    int32_t periodic;
} oa_synth_params;

int oa_synth_keys(const oa_synth_params* params, uint64_t* keys, uint64_t* vals,
                  int64_t* halo_counts, void* stream);
int oa_synth_fill(const oa_synth_params* params, const uint64_t* order,
                  int64_t n_present, int data_dtype, void* pos, void* vel,
                  int64_t* ids, void* stream);
size_t oa_synth_params_size(void);

#ifdef __cplusplus
}
#endif
#endif /* ORBIT_B200_H */
