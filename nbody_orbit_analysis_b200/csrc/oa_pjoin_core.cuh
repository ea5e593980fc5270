// Region-at-a-time partitioned hash join -- second implementation of the fused
// tracking step (reference: track(j) = region_frame + compare_radial_velocities +
// calc_angles, track_orbits.py:147-185, 247-351).  See DESIGN.md section 4.3.
//
// Why: oa_track_kernel does four random global accesses per particle and is
// bound by the SM's L1TEX (one wavefront per lane per divergent access,
// profiles/r01_ablation.md), not by HBM.  Here the carried records of a region
// are kept PARTITIONED by ID hash into pieces that fit shared memory, so the
// match of a particle against the previous snapshot is a shared-memory probe and
// every global access is a stream:
//
//   COUNT   tile of 1024 IDs of region j  -> histogram over the 2^b partitions
//   SCAN    region j                       -> partition offsets / cursors
//   SCATTER tile of region j: halo frame of every particle, new 32 B record
//           written to its partition (one full sector per particle)
//   JOIN    previous partition q of region j -> shared-memory hash table; the
//           records of the current partitions q*f .. (q+1)*f-1 stream through
//           it: sign test, arccos, float16 accumulator, event mark.
//
// One persistent kernel runs all four stages: CTAs take tickets in order, the
// ticket order interleaves the stages of neighbouring groups of regions
// (JOIN of group g-3, SCATTER g-2, SCAN g-1, COUNT g), and per-region counters
// carry the dependencies.  A region's new records are therefore re-read by its
// JOIN while they are still in B200's 126 MB L2: HBM sees the inputs once
// (32 B), the previous records once (32 B) and the new records once (32 B).
//
// This header is compiled twice: by nvcc into liborbit_b200.so (device code,
// IEEE intrinsics), and by g++ with -DPJ_HOST_EMUL into the test-only
// tests/pjoin_emul library, where a CTA is 512 real threads and a barrier --
// the CPU tests run the very same stage code against the oracle.
#pragma once
#include <stdint.h>
#include <math.h>

#include "../../include/orbit_b200.h"

#ifdef PJ_HOST_EMUL
#define PJ_FN inline
#define PJ_UNROLL4
#define PJ_UNROLL_ALL
#else
#include <cuda_fp16.h>
#define PJ_FN __device__ __forceinline__
#define PJ_UNROLL4 _Pragma("unroll 4")
#define PJ_UNROLL_ALL _Pragma("unroll")
#endif

namespace pj {

constexpr int THREADS = OA_PJOIN_THREADS;
constexpr int TILE = OA_PJOIN_TILE;          // particles per SCATTER item
constexpr int CTILE = OA_PJOIN_CTILE;        // particles per COUNT item
constexpr int REC_CAP = OA_PJOIN_REC_CAP;    // previous records per table build
constexpr int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
constexpr int SLOTS = next_pow2(REC_CAP + REC_CAP / 3);   // shared-memory hash slots
constexpr int SCAN_PER = (1 << OA_PJOIN_MAX_BITS) / THREADS;   // scan values per thread
constexpr int MAX_BITS = OA_PJOIN_MAX_BITS;  // at most 2^12 partitions per region
constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr uint16_t NO_EVENT = 0x8000u;       // = OA_NO_EVENT of the legacy path

enum Stage { JOIN = 0, SCATTER = 1, SCAN = 2, COUNT = 3 };

// carried state of one region-particle: exactly one 32-byte sector
struct alignas(16) Rec {
    int64_t id;
    float rx, ry, rz;    // unit vector to the particle in the halo frame
    float vr;            // sign-faithful float copy of the float64 v_r
    uint32_t pos;        // position of the particle in its snapshot (block order)
    uint16_t angle;      // float16 bits of the swept-angle accumulator
    uint16_t flags;      // bit 0: matched by an earlier table batch of this join
};
static_assert(sizeof(Rec) == 32, "record must be one sector");

struct alignas(16) U4 { uint32_t x, y, z, w; };

// ---- shared-memory layout (bytes) ----------------------------------------------------
// join    : records [REC_CAP] | slots [SLOTS]
// scatter : ids | pos | vel | partition [TILE] | rank [TILE] | hist
// count   : hist            scan : values [4096] | partials [THREADS]
constexpr int SM_JOIN_REC = 0;
constexpr int SM_JOIN_SLOT = SM_JOIN_REC + REC_CAP * 32;
constexpr int SM_JOIN_END = SM_JOIN_SLOT + SLOTS * 4;
constexpr int SM_IDS = 0;
constexpr int SM_POS = SM_IDS + TILE * 8;
constexpr int SM_VEL = SM_POS + TILE * 12;
constexpr int SM_PART = SM_VEL + TILE * 12;
constexpr int SM_RANK = SM_PART + TILE * 2;
constexpr int SM_HIST = SM_RANK + TILE * 2;
constexpr int SM_SCATTER_END = SM_HIST + (1 << MAX_BITS) * 4;
constexpr int SM_SCAN_VAL = 0;
constexpr int SM_SCAN_PART = SM_SCAN_VAL + (1 << MAX_BITS) * 4;
constexpr int SM_BODY = SM_JOIN_END > SM_SCATTER_END ? SM_JOIN_END : SM_SCATTER_END;
constexpr int SM_BCAST = SM_BODY;            // 4 words of CTA-wide broadcast
constexpr int SM_MBAR = SM_BCAST + 16;       // one mbarrier (TMA build of the table)
constexpr int PACK_MAX = OA_PJOIN_PACK_MAX;  // small regions per work item
constexpr int SM_PACK = SM_MBAR + 16;        // member tables of a pack: 3 x (PACK_MAX + 1) words
constexpr int SM_BYTES = SM_PACK + 3 * (PACK_MAX + 1) * 4 + 4;
static_assert(SM_JOIN_SLOT % 16 == 0 && SM_HIST % 4 == 0, "alignment");
static_assert(SM_SCAN_PART + THREADS * 4 <= SM_BODY, "scan scratch");
static_assert(SCAN_PER * THREADS == (1 << MAX_BITS), "scan: whole values per thread");
static_assert(REC_CAP < 4095, "record index must fit 12 bits, 4095 is reserved");
static_assert(TILE % THREADS == 0 && CTILE % (8 * THREADS) == 0, "tile loops");

// launch constants derived on the host
struct Const {
    float half_box[3];        // largest float <= L/2 (float-frame wrap test)
    uint32_t total_tickets;
};

// views into the (zeroed) workspace
struct Work {
    uint32_t* ticket;
    uint32_t* done_count;     // [n_regions] COUNT tiles finished
    uint32_t* done_scan;      // [n_regions] 1 when the offsets are final
    uint32_t* done_scatter;   // [n_regions] SCATTER tiles finished
    uint32_t* cursor;         // [n_part_entries] counts, then write cursors
    uint64_t* items;          // [total_tickets] work items in ticket order
};

// ---- IEEE arithmetic without contraction (numpy's rounding points) ---------------------
#ifdef PJ_HOST_EMUL
PJ_FN float fadd(float a, float b) { return a + b; }
PJ_FN float fsub(float a, float b) { return a - b; }
PJ_FN float fmul(float a, float b) { return a * b; }
PJ_FN float fdiv(float a, float b) { return a / b; }
PJ_FN float fsqrt(float a) { return sqrtf(a); }
PJ_FN double dadd(double a, double b) { return a + b; }
PJ_FN double dsub(double a, double b) { return a - b; }
PJ_FN double dmul(double a, double b) { return a * b; }
PJ_FN double ddiv(double a, double b) { return a / b; }
// float32 <-> float16 bits (GCC's _Float16 conversions round to nearest even)
PJ_FN uint16_t half_bits(float f) {
    const _Float16 h = (_Float16)f;
    uint16_t b;
    __builtin_memcpy(&b, &h, 2);
    return b;
}
PJ_FN float half_value(uint16_t b) {
    _Float16 h;
    __builtin_memcpy(&h, &b, 2);
    return (float)h;
}
#else
PJ_FN float fadd(float a, float b) { return __fadd_rn(a, b); }
PJ_FN float fsub(float a, float b) { return __fsub_rn(a, b); }
PJ_FN float fmul(float a, float b) { return __fmul_rn(a, b); }
PJ_FN float fdiv(float a, float b) { return __fdiv_rn(a, b); }
PJ_FN float fsqrt(float a) { return __fsqrt_rn(a); }
PJ_FN double dadd(double a, double b) { return __dadd_rn(a, b); }
PJ_FN double dsub(double a, double b) { return __dsub_rn(a, b); }
PJ_FN double dmul(double a, double b) { return __dmul_rn(a, b); }
PJ_FN double ddiv(double a, double b) { return __ddiv_rn(a, b); }
PJ_FN uint16_t half_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }
PJ_FN float half_value(uint16_t h) { return __half2float(__ushort_as_half(h)); }
#endif

// numpy einsum('...i,...i') over 3 terms: float32 (p0+p1)+p2, float64 (p0+p2)+p1
PJ_FN float dot3f(const float* a, const float* b) {
    return fadd(fadd(fmul(a[0], b[0]), fmul(a[1], b[1])), fmul(a[2], b[2]));
}
PJ_FN double dot3d(const double* a, const double* b) {
    return dadd(dadd(dmul(a[0], b[0]), dmul(a[2], b[2])), dmul(a[1], b[1]));
}
// float copy of v_r whose `< 0` / `> 0` tests agree with the float64 value
PJ_FN float sign_faithful(double v) {
    float f = (float)v;
    if (f == 0.0f && v != 0.0) f = (v > 0.0) ? 1.401298464e-45f : -1.401298464e-45f;
    return f;
}

PJ_FN uint64_t mix64(uint64_t x) {             // = oa_mix64
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}
// partition of an ID among the 2^bits partitions of its region: top hash bits, so
// that doubling the partition count splits every partition in two
PJ_FN uint32_t part_of(uint64_t hash, int bits) {
    return bits > 0 ? (uint32_t)(hash >> 32) >> (32 - bits) : 0u;
}

// ---- halo frame of one particle (region_frame, track_orbits.py:247-290) ------------------
// float32 data, float32 frame, float64 v_r (numpy >= 2 promotion);
// same rounding points as stage_frame<float, float, double> in oa_track.cu
PJ_FN void frame(const oa_pjoin_args& a, const Const& k, const oa_region& R, const float* x,
                 const float* v, float* rh, double* vr) {
    float d[3];
    for (int q = 0; q < 3; ++q) {
        if (!a.centre_f32) {
            double dd = dsub((double)x[q], R.centre[q]);
            if (a.periodic) {
                const double Lq = a.box[q], h = Lq * 0.5;
                if (dd > h) dd = dsub(dd, Lq);
                if (dd < -h) dd = dadd(dd, Lq);
            }
            d[q] = (float)dd;
        } else {
            d[q] = fsub(x[q], R.centre_f[q]);
        }
    }
    if (a.centre_f32 && a.periodic) {
        for (int q = 0; q < 3; ++q) {
            float df = d[q];
            const float hf = k.half_box[q];
            if (df > hf) df = (float)dsub((double)df, a.box[q]);
            if (df < -hf) df = (float)dadd((double)df, a.box[q]);
            d[q] = df;
        }
    }
    const float r = fsqrt(dot3f(d, d));
    for (int q = 0; q < 3; ++q) rh[q] = fdiv(d[q], r);
    double w[3], rd[3];
    for (int q = 0; q < 3; ++q) {
        double wk;
        if (!a.bulk_f32) wk = dsub((double)v[q], R.bulk[q]);
        else wk = (double)fsub(v[q], R.bulk_f[q]);
        if (a.hubble_on) wk = dadd(wk, ddiv(dmul(a.hubble, (double)d[q]), a.one_plus_z));
        w[q] = wk;
        rd[q] = (double)rh[q];
    }
    *vr = dot3d(w, rd);
}

// ---- CTA-wide helpers --------------------------------------------------------------------------
// CX (execution context) provides: tid(), sync(), smem(), atomic_add / atomic_cas
// on uint32 (shared or global), load_acquire / release_add (gpu scope), backoff(),
// fail(), ld_cg / load_rec_cg (global data another CTA of this launch may have
// written), ld_stream / ld_last (read-only data touched for the last time),
// store_rec.
// (a dependency always has a smaller ticket, i.e. a CTA that is running: the wait
// is short.  A wait of seconds can only be a bug -- it ends the kernel with an
// error instead of hanging the device.)
constexpr uint32_t MAX_SPINS = 1u << 25;
template <class CX>
PJ_FN void wait_ge(CX& cx, const uint32_t* p, uint32_t need) {
    if (cx.tid() == 0) {
        uint32_t spins = 0;
        const uint64_t t0 = cx.clock();
        while (cx.load_acquire(p) < need) {
            cx.backoff();
            if (++spins > MAX_SPINS) cx.fail();
        }
        cx.stat_add(4, cx.clock() - t0);      // cycles spent waiting for a dependency
    }
    cx.sync();
}
template <class CX>
PJ_FN void signal(CX& cx, uint32_t* p) {
    cx.sync();                         // all global writes of the item are issued
    if (cx.tid() == 0) cx.release_add(p, 1u);
}

struct Item {
    int stage;
    int region;
    uint32_t idx;
};

PJ_FN uint32_t stage_prefix(const oa_pjoin_region* plan, int stage, int j) {
    return stage == JOIN ? plan[j].join_first
         : stage == SCAN ? plan[j].scan_first
         : stage == COUNT ? plan[j].count_first : plan[j].tile_first;
}

PJ_FN uint32_t stage_count_of(const oa_pjoin_region* plan, int stage, int j) {
    return stage_prefix(plan, stage, j + 1) - stage_prefix(plan, stage, j);
}

// Work items in ticket order.  Range r = 4 * superstep + stage holds the items of
// group (superstep - lag[stage]), lags: JOIN 3, SCATTER 2, SCAN 1, COUNT 0; inside
// a range, regions in order and a region's items in order.  item = stage << 62 |
// region << 32 | index.  (Every dependency of an item has a smaller ticket.)
PJ_FN uint64_t encode_item(int stage, int region, uint32_t idx) {
    return ((uint64_t)stage << 62) | ((uint64_t)(uint32_t)region << 32) | idx;
}
PJ_FN Item decode_item(uint64_t v) {
    Item it;
    it.stage = (int)(v >> 62);
    it.region = (int)((v >> 32) & 0x3FFFFFFFu);
    it.idx = (uint32_t)v;
    return it;
}
// first ticket of region j's items of `stage`
PJ_FN uint32_t item_base(const oa_pjoin_args& a, int stage, int j) {
    int lo = 0, hi = a.n_groups - 1;                    // group of region j
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int)a.group_first[mid] <= j) lo = mid; else hi = mid - 1;
    }
    const int r = 4 * (lo + (3 - stage)) + stage;
    return a.range_start[r] + stage_prefix(a.plan, stage, j) -
           stage_prefix(a.plan, stage, (int)a.group_first[lo]);
}

// items of region j, written by `nthreads` cooperating threads
PJ_FN void expand_region(const oa_pjoin_args& a, uint64_t* items, int j, int tid,
                         int nthreads) {
    for (int stage = 0; stage < 4; ++stage) {
        const uint32_t cnt = stage_count_of(a.plan, stage, j);
        if (cnt == 0) continue;
        const uint32_t base = item_base(a, stage, j);
        for (uint32_t i = (uint32_t)tid; i < cnt; i += (uint32_t)nthreads)
            items[base + i] = encode_item(stage, j, i);
    }
}

PJ_FN uint32_t tiles_of(int64_t count) { return (uint32_t)((count + TILE - 1) / TILE); }
PJ_FN uint32_t ctiles_of(int64_t count) { return (uint32_t)((count + CTILE - 1) / CTILE); }

// ---- COUNT ---------------------------------------------------------------------------------------
template <class CX>
PJ_FN void stage_count(CX& cx, const oa_pjoin_args& a, const Work& w, int j, uint32_t t) {
    const oa_region& R = a.regions[j];
    const oa_pjoin_region& P = a.plan[j];
    const int bits = P.bits_cur, nP = 1 << bits;
    const int64_t begin = R.cur_begin + (int64_t)t * CTILE;
    const int64_t left = R.cur_count - (int64_t)t * CTILE;
    const int cnt = (int)(left < CTILE ? left : CTILE);
    uint32_t* hist = reinterpret_cast<uint32_t*>(cx.smem() + SM_HIST);
    for (int p = cx.tid(); p < nP; p += THREADS) hist[p] = 0;
    cx.sync();
    for (int k0 = 0; k0 < CTILE / THREADS; k0 += 8) {
        int64_t id[8];
        PJ_UNROLL_ALL
        for (int k = 0; k < 8; ++k) {                           // 8 IDs in flight
            const int i = cx.tid() + (k0 + k) * THREADS;
            id[k] = i < cnt ? a.ids[begin + i] : 0;
        }
        PJ_UNROLL_ALL
        for (int k = 0; k < 8; ++k)
            if (cx.tid() + (k0 + k) * THREADS < cnt)
                cx.atomic_add(&hist[part_of(mix64((uint64_t)id[k]), bits)], 1u);
    }
    cx.sync();
    for (int p = cx.tid(); p < nP; p += THREADS)
        if (hist[p]) cx.atomic_add(&w.cursor[P.pb_cur + p], hist[p]);
    signal(cx, &w.done_count[j]);
}

// ---- SCAN ----------------------------------------------------------------------------------------
template <class CX>
PJ_FN void stage_scan(CX& cx, const oa_pjoin_args& a, const Work& w, int j) {
    const oa_region& R = a.regions[j];
    const oa_pjoin_region& P = a.plan[j];
    const int nP = 1 << P.bits_cur;
    wait_ge(cx, &w.done_count[j], ctiles_of(R.cur_count));
    uint32_t* val = reinterpret_cast<uint32_t*>(cx.smem() + SM_SCAN_VAL);
    uint32_t* part = reinterpret_cast<uint32_t*>(cx.smem() + SM_SCAN_PART);
    const int tid = cx.tid();
    // SCAN_PER consecutive values per thread, then a scan of the partial sums
    uint32_t own = 0;
    for (int e = 0; e < SCAN_PER; ++e) {
        const int p = tid * SCAN_PER + e;
        const uint32_t v = p < nP ? cx.ld_cg(&w.cursor[P.pb_cur + p]) : 0u;
        val[p] = own;                  // exclusive within the thread
        own += v;
    }
    part[tid] = own;
    cx.sync();
    for (int d = 1; d < THREADS; d <<= 1) {
        const uint32_t add = tid >= d ? part[tid - d] : 0u;
        cx.sync();
        part[tid] += add;
        cx.sync();
    }
    const uint32_t base = (uint32_t)R.cur_begin + part[tid] - own;
    for (int e = 0; e < SCAN_PER; ++e) {
        const int p = tid * SCAN_PER + e;
        if (p < nP) {
            const uint32_t off = base + val[p];
            a.part_off_cur[P.pb_cur + p] = off;
            w.cursor[P.pb_cur + p] = off;
        }
    }
    if (tid == 0) a.part_off_cur[P.pb_cur + nP] = (uint32_t)(R.cur_begin + R.cur_count);
    signal(cx, &w.done_scan[j]);
}

// inputs of one tile -> shared memory (coalesced element loads)
template <class CX>
PJ_FN void load_tile(CX& cx, const oa_pjoin_args& a, int64_t begin, int cnt) {
    int64_t* s_ids = reinterpret_cast<int64_t*>(cx.smem() + SM_IDS);
    float* s_pos = reinterpret_cast<float*>(cx.smem() + SM_POS);
    float* s_vel = reinterpret_cast<float*>(cx.smem() + SM_VEL);
    // last use of the inputs: streaming loads, they must not displace the records
    // that wait in L2 for their JOIN
    // (fixed trip counts + predicates: fully unrolled, every load of the tile is
    // in flight before the first shared-memory store)
    PJ_UNROLL_ALL
    for (int k = 0; k < TILE / THREADS; ++k) {
        const int i = cx.tid() + k * THREADS;
        if (i < cnt) s_ids[i] = cx.ld_last(a.ids + begin + i);
    }
    PJ_UNROLL_ALL
    for (int k = 0; k < 3 * TILE / THREADS; ++k) {
        const int i = cx.tid() + k * THREADS;
        if (i < 3 * cnt) s_pos[i] = cx.ld_last(a.pos + 3 * begin + i);
    }
    PJ_UNROLL_ALL
    for (int k = 0; k < 3 * TILE / THREADS; ++k) {
        const int i = cx.tid() + k * THREADS;
        if (i < 3 * cnt) s_vel[i] = cx.ld_last(a.vel + 3 * begin + i);
    }
}

// new record of particle i of the staged tile (angle accumulator 0: a particle
// without a match starts from zero, track_orbits.py:180-183, 344-346)
template <class CX>
PJ_FN Rec make_record(CX& cx, const oa_pjoin_args& a, const Const& k, const oa_region& R,
                      int i, int64_t c) {
    const int64_t* s_ids = reinterpret_cast<const int64_t*>(cx.smem() + SM_IDS);
    const float* s_pos = reinterpret_cast<const float*>(cx.smem() + SM_POS);
    const float* s_vel = reinterpret_cast<const float*>(cx.smem() + SM_VEL);
    float rh[3];
    double vr;
    frame(a, k, R, s_pos + 3 * i, s_vel + 3 * i, rh, &vr);
    Rec rec;
    rec.id = s_ids[i];
    rec.rx = rh[0]; rec.ry = rh[1]; rec.rz = rh[2];
    rec.vr = sign_faithful(vr);
    rec.pos = (uint32_t)c;
    rec.angle = 0;
    rec.flags = 0;
    return rec;
}

// ---- SCATTER -------------------------------------------------------------------------------------
template <class CX>
PJ_FN void stage_scatter(CX& cx, const oa_pjoin_args& a, const Const& k, const Work& w, int j,
                         uint32_t t) {
    const oa_region& R = a.regions[j];
    const oa_pjoin_region& P = a.plan[j];
    const int bits = P.bits_cur, nP = 1 << bits;
    const int64_t begin = R.cur_begin + (int64_t)t * TILE;
    const int64_t left = R.cur_count - (int64_t)t * TILE;
    const int cnt = (int)(left < TILE ? left : TILE);
    const int64_t* s_ids = reinterpret_cast<const int64_t*>(cx.smem() + SM_IDS);
    uint16_t* s_part = reinterpret_cast<uint16_t*>(cx.smem() + SM_PART);
    uint16_t* s_rank = reinterpret_cast<uint16_t*>(cx.smem() + SM_RANK);
    uint32_t* hist = reinterpret_cast<uint32_t*>(cx.smem() + SM_HIST);
    Rec* rec_cur = static_cast<Rec*>(a.rec_cur);

    load_tile(cx, a, begin, cnt);
    for (int p = cx.tid(); p < nP; p += THREADS) hist[p] = 0;
    cx.sync();
    // pass 1: partition and rank (inside the tile) of every particle
    for (int i = cx.tid(); i < cnt; i += THREADS) {
        const uint32_t p = part_of(mix64((uint64_t)s_ids[i]), bits);
        s_part[i] = (uint16_t)p;
        s_rank[i] = (uint16_t)cx.atomic_add(&hist[p], 1u);
    }
    cx.sync();
    // the offsets of this region are final before any of its tiles reserves room
    wait_ge(cx, &w.done_scan[j], 1u);
    for (int p = cx.tid(); p < nP; p += THREADS)
        if (hist[p]) hist[p] = cx.atomic_add(&w.cursor[P.pb_cur + p], hist[p]);
    cx.sync();
    // pass 2: halo frame, record -> its place in the partition (one sector)
    for (int i = cx.tid(); i < cnt; i += THREADS) {
        cx.store_rec(rec_cur + (hist[s_part[i]] + s_rank[i]),
                     make_record(cx, a, k, R, i, begin + i));
        a.mark_cur[begin + i] = NO_EVENT;
    }
    signal(cx, &w.done_scatter[j]);
}

// one current record against the shared-memory table of previous records:
// compare_radial_velocities + calc_angles (track_orbits.py:311-325, 330-351)
// (prev_lo, prev_cnt): block range of the halo's previous block -- a table that
// holds several halos (a pack) may contain the same ID once per halo
PJ_FN void probe_one(const oa_pjoin_args& a, const Rec* s_rec, const uint32_t* slots,
                     const Rec& cur, Rec* cur_global, uint32_t prev_lo = 0u,
                     uint32_t prev_cnt = 0xFFFFFFFFu) {
    const uint32_t h = (uint32_t)mix64((uint64_t)cur.id);
    uint32_t s = h & (SLOTS - 1);
    Rec prev;
    for (;;) {
        const uint32_t v = slots[s];
        if (v == EMPTY) return;             // newly entered: accumulator stays 0
        if ((v >> 12) == (h >> 12)) {       // 20-bit fingerprint: the record is read
            prev = s_rec[v & 0xFFFu];       // once, its ID settles the match
            if (prev.id == cur.id && prev.pos - prev_lo < prev_cnt) break;
        }
        s = (s + 1) & (SLOTS - 1);
    }
    const float pr[3] = {prev.rx, prev.ry, prev.rz}, cr[3] = {cur.rx, cur.ry, cur.rz};
    const float dang = acosf(dot3f(pr, cr));
    bool ev;
    if (a.mode == OA_MODE_PERICENTRIC) ev = (prev.vr < 0) && (cur.vr > 0);
    else ev = (prev.vr > 0) && (cur.vr < 0);
    float run = fadd(half_value(prev.angle), dang);
    if (ev) {
        a.mark_prev[prev.pos] = half_bits(run);
        run = 0.0f;
    }
    // angle accumulator + "matched" flag: the last word of the record
    reinterpret_cast<uint32_t*>(cur_global)[7] = (uint32_t)half_bits(run) | (1u << 16);
}

// every staged previous record -> open-addressing table (slot = 20-bit
// fingerprint | 12-bit record index)
template <class CX>
PJ_FN void table_insert_all(CX& cx, const Rec* s_rec, uint32_t* slots, int nb) {
    for (int i = cx.tid(); i < nb; i += THREADS) {
        const uint32_t h = (uint32_t)mix64((uint64_t)s_rec[i].id);
        const uint32_t val = ((h >> 12) << 12) | (uint32_t)i;
        uint32_t s = h & (SLOTS - 1);
        while (cx.atomic_cas(&slots[s], EMPTY, val) != EMPTY) s = (s + 1) & (SLOTS - 1);
    }
}

// last m in [0, n) with arr[m] <= x (arr non-decreasing, arr[0] <= x): equal
// entries belong to empty members and are skipped by taking the last
PJ_FN int member_of(const uint32_t* arr, int n, uint32_t x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (arr[mid] <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ---- JOIN ----------------------------------------------------------------------------------------
// previous records [pb, pe) against current records [cb, ce) (both of one halo,
// the current range holds every particle whose hash falls in the previous one)
template <class CX>
PJ_FN void join_ranges(CX& cx, const oa_pjoin_args& a, uint32_t pb, uint32_t pe, uint32_t cb,
                       uint32_t ce) {
    Rec* s_rec = reinterpret_cast<Rec*>(cx.smem() + SM_JOIN_REC);
    uint32_t* slots = reinterpret_cast<uint32_t*>(cx.smem() + SM_JOIN_SLOT);
    const Rec* rec_prev = static_cast<const Rec*>(a.rec_prev);
    Rec* rec_cur = static_cast<Rec*>(a.rec_cur);
    const bool multi = pe - pb > (uint32_t)REC_CAP;
    const int ncur = (int)(ce - cb);
    for (uint32_t bs = pb; bs < pe; bs += REC_CAP) {
        const int nb = (int)(pe - bs < (uint32_t)REC_CAP ? pe - bs : (uint32_t)REC_CAP);
        for (int s = cx.tid(); s < SLOTS; s += THREADS) slots[s] = EMPTY;
#if OA_PJOIN_TMA
        // one TMA bulk copy (cp.async.bulk + mbarrier) instead of register staging;
        // the slot clear above overlaps it
        cx.bulk_load(s_rec, rec_prev + bs, (uint32_t)nb * 32u);
#else
        {   // previous records -> shared memory, 16 bytes per thread and step
            const U4* src = reinterpret_cast<const U4*>(rec_prev + bs);
            U4* dst = reinterpret_cast<U4*>(s_rec);
            constexpr int STEPS = (2 * REC_CAP + THREADS - 1) / THREADS;
            for (int k0 = 0; k0 < STEPS; k0 += 6) {
                PJ_UNROLL_ALL
                for (int k = 0; k < 6; ++k) {                   // 6 x 16 B in flight
                    const int q = cx.tid() + (k0 + k) * THREADS;
                    if (q < 2 * nb) dst[q] = cx.ld_stream(src + q);
                }
            }
        }
#endif
        cx.sync();
        table_insert_all(cx, s_rec, slots, nb);
        cx.sync();
        // one record of lookahead: the load of the next record is in flight while
        // this one is probed
        int c = cx.tid();
        Rec nxt = {};
        if (c < ncur) nxt = cx.load_rec_cg(rec_cur + cb + c);
        while (c < ncur) {
            const Rec cur = nxt;
            const int cn = c + THREADS;
            if (cn < ncur) nxt = cx.load_rec_cg(rec_cur + cb + cn);
            if (!(multi && (cur.flags & 1u)))
                probe_one(a, s_rec, slots, cur, rec_cur + cb + c);
            c = cn;
        }
        cx.sync();                     // the table is rebuilt by the next batch / item
    }
}

// ---- a pack of small regions -------------------------------------------------------------------
// `L` consecutive regions j .. j+L-1, each a single partition in block order, at
// most TILE particles and REC_CAP previous records together (host plan): one tile
// of inputs, one table that holds the previous blocks of all members back to back.
template <class CX>
PJ_FN void stage_pack(CX& cx, const oa_pjoin_args& a, const Const& k, int j, int L) {
    uint32_t* m_cur = reinterpret_cast<uint32_t*>(cx.smem() + SM_PACK);   // block starts + end
    uint32_t* m_pre = m_cur + (PACK_MAX + 1);      // prefix of previous-record counts
    uint32_t* m_pbeg = m_pre + (PACK_MAX + 1);     // previous block starts
    Rec* s_rec = reinterpret_cast<Rec*>(cx.smem() + SM_JOIN_REC);
    uint32_t* slots = reinterpret_cast<uint32_t*>(cx.smem() + SM_JOIN_SLOT);
    Rec* rec_cur = static_cast<Rec*>(a.rec_cur);
    const Rec* rec_prev = static_cast<const Rec*>(a.rec_prev);
    if (cx.tid() == 0) {
        uint32_t acc = 0;
        for (int m = 0; m < L; ++m) {
            const oa_region& R = a.regions[j + m];
            const oa_pjoin_region& P = a.plan[j + m];
            const uint32_t pc = (P.bits_prev >= 0 && R.prev_count > 0) ? (uint32_t)R.prev_count : 0u;
            m_cur[m] = (uint32_t)R.cur_begin;
            m_pre[m] = acc;
            m_pbeg[m] = pc ? (uint32_t)R.prev_begin : 0u;
            acc += pc;
            a.part_off_cur[P.pb_cur] = (uint32_t)R.cur_begin;
            a.part_off_cur[P.pb_cur + 1] = (uint32_t)(R.cur_begin + R.cur_count);
        }
        const oa_region& last = a.regions[j + L - 1];
        m_cur[L] = (uint32_t)(last.cur_begin + last.cur_count);
        m_pre[L] = acc;
    }
    cx.sync();
    const uint32_t begin = m_cur[0];
    const int cnt = (int)(m_cur[L] - begin);
    const int nb = (int)m_pre[L];
    // frame + records of all members (block order), one tile
    load_tile(cx, a, (int64_t)begin, cnt);
    cx.sync();
    for (int i = cx.tid(); i < cnt; i += THREADS) {
        const uint32_t c = begin + (uint32_t)i;
        const int m = member_of(m_cur, L, c);
        cx.store_rec(rec_cur + c, make_record(cx, a, k, a.regions[j + m], i, (int64_t)c));
        a.mark_cur[c] = NO_EVENT;
    }
    cx.sync();
    if (nb == 0) return;                           // no member has a previous block
    for (int s = cx.tid(); s < SLOTS; s += THREADS) slots[s] = EMPTY;
    {
        U4* dst = reinterpret_cast<U4*>(s_rec);
        for (int q = cx.tid(); q < 2 * nb; q += THREADS) {
            const uint32_t r = (uint32_t)q >> 1;
            const int m = member_of(m_pre, L, r);
            const U4* src = reinterpret_cast<const U4*>(rec_prev + m_pbeg[m] + (r - m_pre[m]));
            dst[q] = cx.ld_stream(src + (q & 1));
        }
    }
    cx.sync();
    table_insert_all(cx, s_rec, slots, nb);
    cx.sync();
    for (int i = cx.tid(); i < cnt; i += THREADS) {
        const uint32_t c = begin + (uint32_t)i;
        const int m = member_of(m_cur, L, c);
        const uint32_t pc = m_pre[m + 1] - m_pre[m];
        if (pc == 0) continue;                     // a halo without a previous block
        const Rec cur = cx.load_rec_cg(rec_cur + c);
        probe_one(a, s_rec, slots, cur, rec_cur + c, m_pbeg[m], pc);
    }
}

template <class CX>
PJ_FN void stage_join(CX& cx, const oa_pjoin_args& a, const Const& k, const Work& w, int j,
                      uint32_t q) {
    const oa_region& R = a.regions[j];
    const oa_pjoin_region& P = a.plan[j];
    if (P.bits_cur > 0) {
        // partitioned region: previous partition q, current partitions q*f .. (q+1)*f - 1
        wait_ge(cx, &w.done_scatter[j], tiles_of(R.cur_count));
        const int db = P.bits_cur - P.bits_prev;
        const uint32_t pb = a.part_off_prev[P.pb_prev + q], pe = a.part_off_prev[P.pb_prev + q + 1];
        const uint32_t cb = cx.ld_cg(&a.part_off_cur[P.pb_cur + (q << db)]);
        const uint32_t ce = cx.ld_cg(&a.part_off_cur[P.pb_cur + ((q + 1) << db)]);
        join_ranges(cx, a, pb, pe, cb, ce);
        return;
    }
    if (P.pack_len > 1) {
        stage_pack(cx, a, k, j, (int)P.pack_len);
        return;
    }
    // small region (one partition = the block itself, in block order): frame and
    // records tile by tile, then the join against the previous block
    Rec* rec_cur = static_cast<Rec*>(a.rec_cur);
    const uint32_t tiles = tiles_of(R.cur_count);
    for (uint32_t t = 0; t < tiles; ++t) {
        const int64_t begin = R.cur_begin + (int64_t)t * TILE;
        const int cnt = (int)(R.cur_count - (int64_t)t * TILE < TILE
                                  ? R.cur_count - (int64_t)t * TILE : TILE);
        load_tile(cx, a, begin, cnt);
        cx.sync();
        for (int i = cx.tid(); i < cnt; i += THREADS) {
            cx.store_rec(rec_cur + begin + i, make_record(cx, a, k, R, i, begin + i));
            a.mark_cur[begin + i] = NO_EVENT;
        }
        cx.sync();
    }
    if (cx.tid() == 0) {
        a.part_off_cur[P.pb_cur] = (uint32_t)R.cur_begin;
        a.part_off_cur[P.pb_cur + 1] = (uint32_t)(R.cur_begin + R.cur_count);
    }
    if (P.bits_prev >= 0 && R.prev_count > 0)
        join_ranges(cx, a, (uint32_t)R.prev_begin, (uint32_t)(R.prev_begin + R.prev_count),
                    (uint32_t)R.cur_begin, (uint32_t)(R.cur_begin + R.cur_count));
}

// ---- the persistent CTA loop --------------------------------------------------------------------
template <class CX>
PJ_FN void run(CX& cx, const oa_pjoin_args& a, const Const& k, const Work& w) {
    uint32_t* bc = reinterpret_cast<uint32_t*>(cx.smem() + SM_BCAST);
    // ONE barrier per item hands out the next ticket and separates the items'
    // uses of shared memory: the ticket word alternates between two slots, so the
    // word of item n is not rewritten before every thread has passed the barrier
    // of item n + 1, i.e. has read it.
    for (uint32_t n = 0;; ++n) {
        if (cx.tid() == 0) bc[n & 1u] = cx.atomic_add(w.ticket, 1u);
        cx.sync();
        const uint32_t t = bc[n & 1u];
        if (t >= k.total_tickets) break;
        const Item it = decode_item(w.items[t]);
        const uint64_t t0 = cx.clock();
        if (it.stage == COUNT) stage_count(cx, a, w, it.region, it.idx);
        else if (it.stage == SCAN) stage_scan(cx, a, w, it.region);
        else if (it.stage == SCATTER) stage_scatter(cx, a, k, w, it.region, it.idx);
        else stage_join(cx, a, k, w, it.region, it.idx);
        // (profiling builds only, -DOA_PJOIN_STATS=1: cycles and items per stage)
        if (cx.tid() == 0) {
            cx.stat_add(it.stage, cx.clock() - t0);
            cx.stat_add(8 + it.stage, 1);
        }
    }
}

}  // namespace pj
