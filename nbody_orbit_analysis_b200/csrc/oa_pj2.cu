// Partitioned hash join, second generation (include/orbit_b200.h: oa_pj2_step).
// Replaces track(j) = region_frame + compare_radial_velocities + calc_angles of
// the reference (track_orbits.py:147-185, 247-351; utils.py:4-33) for float32
// data and a float32 catalogue.  DESIGN.md section 4.4.
//
// Structure (what the profile of the first generation asked for,
// profiles/r02_c1_*: no CTA-wide barriers on the critical path, no exposed
// global-load latency):
//
//   * one persistent kernel, OA_PJ2_MIN_CTAS CTAs per SM, each CTA = 1 PRODUCER
//     warp + OA_PJ2_THREADS consumer threads and two shared-memory stages;
//   * the producer takes tickets, resolves the item (tile of inputs / pair of
//     partitions), waits for the item's dependency, and fills a stage with
//     cp.async.bulk (TMA) copies that complete on the stage's `full` mbarrier;
//     consumers release a stage through its `empty` mbarrier, warp by warp;
//   * SCATTER (a tile of 1408 consecutive particles): halo frame of every
//     particle with numpy's rounding points, 32 B record stored into the ID-hash
//     partition of its region; the slot comes from one atomicAdd on the
//     partition's fill word.  Warps are autonomous: no barrier, each warp
//     signals its share of the tile to the group counter;
//   * JOIN (cur partition p vs prev partition p >> shift): previous records ->
//     open-addressing index in shared memory (generation-tagged entries, never
//     cleared), ONE named barrier, current records probe it: sign test, arccos,
//     float16 accumulator written back, event mark.
#include "oa_common.cuh"

#ifndef OA_PJ2_STATS
#define OA_PJ2_STATS 0
#endif
__device__ unsigned long long g_pj2_stats[16];

namespace pj2 {

constexpr int CONSUMERS = OA_PJ2_THREADS;
constexpr int CWARPS = CONSUMERS / 32;
constexpr int THREADS = CONSUMERS + 32;      // + the producer warp (the LAST warp)
constexpr int TILE = OA_PJ2_TILE;
constexpr int CAP = OA_PJ2_CAP;
constexpr int KMAX = (TILE + CONSUMERS - 1) / CONSUMERS;     // particles per thread and tile
constexpr int JMAX = (CAP + CONSUMERS - 1) / CONSUMERS;      // records per thread and partition
constexpr int STAGE_BYTES = TILE * 32;
constexpr int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
constexpr int SLOTS = next_pow2(2 * CAP);   // index entries per stage
constexpr uint16_t NO_EVENT = OA_NO_EVENT;
static_assert(CONSUMERS % 32 == 0 && TILE % 4 == 0, "shape");
static_assert(2 * CAP * 32 <= STAGE_BYTES, "a stage holds both partitions of a JOIN");
static_assert(CAP <= 1024 && SLOTS >= 2 * CAP, "index entry: 10-bit record index");

enum Kind : uint32_t { SCATTER = 0, JOIN = 1, EXIT = 2 };

// carried state of one region-particle: one 32-byte sector (same as generation 1)
struct alignas(16) Rec {
    int64_t id;
    float rx, ry, rz;    // unit vector to the particle in the halo frame
    float vr;            // sign-faithful float copy of the float64 v_r
    uint32_t pos;        // position of the particle in its snapshot (block order)
    uint16_t angle;      // float16 bits of the swept-angle accumulator
    uint16_t flags;
};
static_assert(sizeof(Rec) == 32, "record must be one sector");

// what the producer tells the consumers about a stage
struct alignas(16) Desc {
    uint32_t kind;
    uint32_t region;     // SCATTER: region of the tile's first particle; JOIN: the region
    uint32_t idx;        // SCATTER: tile; JOIN: current partition
    uint32_t gen;        // JOIN: generation tag of the index entries (1..255)
    uint32_t nA;         // JOIN: previous records in the stage
    uint32_t nB;         // JOIN: current records in the stage / SCATTER: particles
    uint32_t direct;     // SCATTER: 1 = inputs are read from global memory (last, partial tile)
    uint32_t slot0;      // JOIN: first record slot of the current partition
};

// shared-memory layout (bytes)
constexpr int SM_STAGE = 0;
constexpr int SM_INDEX = SM_STAGE + 2 * STAGE_BYTES;
constexpr int SM_DESC = SM_INDEX + 2 * SLOTS * 4;
constexpr int SM_BAR = SM_DESC + 2 * (int)sizeof(Desc);     // full[2], empty[2]
constexpr int SM_BYTES = SM_BAR + 4 * 8;

struct Const {
    float half_box[3];        // largest float <= L/2 (float-frame wrap test)
    uint32_t total_tickets;
    uint32_t need_scale;      // signals per finished tile (= consumer warps)
};

struct Work {
    uint2* items;             // [total_tickets] (kind << 31 | region, idx)
    uint32_t* ticket;
    uint32_t* done;           // [n_groups] consumer-warp signals of the group's tiles
};

// ---- IEEE arithmetic without contraction (numpy's rounding points) -----------------------
OA_D float fadd(float a, float b) { return __fadd_rn(a, b); }
OA_D float fsub(float a, float b) { return __fsub_rn(a, b); }
OA_D float fmul(float a, float b) { return __fmul_rn(a, b); }
OA_D float fdiv(float a, float b) { return __fdiv_rn(a, b); }
OA_D double dadd(double a, double b) { return __dadd_rn(a, b); }
OA_D double dsub(double a, double b) { return __dsub_rn(a, b); }
OA_D double dmul(double a, double b) { return __dmul_rn(a, b); }
OA_D double ddiv(double a, double b) { return __ddiv_rn(a, b); }
// numpy einsum('...i,...i') over 3 terms: float32 (p0+p1)+p2, float64 (p0+p2)+p1
OA_D float dot3f(const float* a, const float* b) {
    return fadd(fadd(fmul(a[0], b[0]), fmul(a[1], b[1])), fmul(a[2], b[2]));
}
OA_D double dot3d(const double* a, const double* b) {
    return dadd(dadd(dmul(a[0], b[0]), dmul(a[2], b[2])), dmul(a[1], b[1]));
}
// float copy of v_r whose `< 0` / `> 0` tests agree with the float64 value
OA_D float sign_faithful(double v) {
    float f = (float)v;
    if (f == 0.0f && v != 0.0) f = (v > 0.0) ? 1.401298464e-45f : -1.401298464e-45f;
    return f;
}

// ---- PTX helpers -----------------------------------------------------------------------
OA_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
OA_D void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
OA_D void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
OA_D void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
OA_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n"
                 " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 " selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
OA_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
OA_D void tma_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
                 "[%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
OA_D void tma_load_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                        uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
                 ".L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
OA_D uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// generic-proxy accesses before / async-proxy (TMA) accesses after
OA_D void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
OA_D void consumer_barrier() { asm volatile("bar.sync 1, %0;" :: "n"(CONSUMERS) : "memory"); }
OA_D uint32_t ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
OA_D void red_release(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
OA_D void store_rec(Rec* p, const Rec& r) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&r);
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]),
                    "r"(w[6]), "r"(w[7])
                 : "memory");
}
#if OA_PJ2_STATS
OA_D uint64_t now() { return (uint64_t)clock64(); }
OA_D void stat_add(int i, uint64_t v) { atomicAdd(&g_pj2_stats[i], (unsigned long long)v); }
#else
OA_D uint64_t now() { return 0; }
OA_D void stat_add(int, uint64_t) {}
#endif

constexpr uint32_t MAX_SPINS = 1u << 24;      // a dependency wait of seconds is a bug: trap

OA_D uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// ---- halo frame of one particle (region_frame, track_orbits.py:247-290) -------------------
// float32 data, float32 centre, float64 v_r (numpy >= 2 promotion);
// same rounding points as stage_frame<float, float, double> in oa_track.cu
struct RegionCache {
    int j;
    int64_t end;             // cur_begin + cur_count
    float c[3], bf[3];
    double bd[3];
    uint32_t P, cap, base, pb;
};
OA_D void load_region(const oa_pj2_args& a, int j, RegionCache& R) {
    const oa_region& g = a.regions[j];
    R.j = j;
    R.end = g.cur_begin + g.cur_count;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        R.c[q] = g.centre_f[q];
        R.bf[q] = g.bulk_f[q];
        R.bd[q] = g.bulk[q];
    }
    const oa_pj2_region& p = a.plan[j];
    R.P = p.P_cur; R.cap = p.cap_cur; R.base = p.base_cur; R.pb = p.pb_cur;
}
OA_D void frame(const oa_pj2_args& a, const Const& k, const RegionCache& R, const float* x,
                const float* v, float* rh, double* vr) {
    float d[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) d[q] = fsub(x[q], R.c[q]);
    if (a.periodic) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float df = d[q];
            const float hf = k.half_box[q];
            if (df > hf) df = (float)dsub((double)df, a.box[q]);
            if (df < -hf) df = (float)dadd((double)df, a.box[q]);
            d[q] = df;
        }
    }
    const float r = __fsqrt_rn(dot3f(d, d));
    double w[3], rd[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        rh[q] = fdiv(d[q], r);
        double wk;
        if (!a.bulk_f32) wk = dsub((double)v[q], R.bd[q]);
        else wk = (double)fsub(v[q], R.bf[q]);
        if (a.hubble_on) wk = dadd(wk, ddiv(dmul(a.hubble, (double)d[q]), a.one_plus_z));
        w[q] = wk;
        rd[q] = (double)rh[q];
    }
    *vr = dot3d(w, rd);
}

// ---- SCATTER: consumer side ------------------------------------------------------------------
OA_D void scatter_tile(const oa_pj2_args& a, const Const& k, const Work& w, const Desc& d,
                       const unsigned char* stage, int tid) {
    const int64_t c0 = (int64_t)d.idx * TILE;
    const int cnt = (int)d.nB;
    const int64_t* s_ids = reinterpret_cast<const int64_t*>(stage);
    const float* s_pos = reinterpret_cast<const float*>(stage + TILE * 8);
    const float* s_vel = reinterpret_cast<const float*>(stage + TILE * 20);
    Rec* rec_cur = static_cast<Rec*>(a.rec_cur);

    RegionCache R;
    load_region(a, (int)d.region, R);
    // pass 1: region, partition, slot (the atomics of all KMAX particles are in
    // flight during the frame arithmetic of pass 2)
    int64_t id[KMAX];
    uint32_t slot[KMAX], dst[KMAX], cap[KMAX];
    int jk[KMAX];
#pragma unroll
    for (int q = 0; q < KMAX; ++q) {
        const int i = tid + q * CONSUMERS;
        slot[q] = 0;
        dst[q] = 0;
        cap[q] = 0;
        jk[q] = R.j;
        id[q] = 0;
        if (i < cnt) {
            const int64_t c = c0 + i;
            while (c >= R.end) load_region(a, R.j + 1, R);
            jk[q] = R.j;
            id[q] = d.direct ? __ldcs(a.ids + c) : s_ids[i];
            const uint32_t hi = (uint32_t)(oa_mix64((uint64_t)id[q]) >> 32);
            const uint32_t p = __umulhi(hi, R.P);
            slot[q] = atomicAdd(a.fill_cur + R.pb + p, 1u);      // consumed in pass 2
            dst[q] = R.base + p * R.cap;
            cap[q] = R.cap;
        }
    }
    // pass 2: frame, record -> its slot
    if (R.j != jk[0]) load_region(a, jk[0], R);
#pragma unroll
    for (int q = 0; q < KMAX; ++q) {
        const int i = tid + q * CONSUMERS;
        if (i < cnt) {
            const int64_t c = c0 + i;
            if (R.j != jk[q]) load_region(a, jk[q], R);
            float x[3], v[3], rh[3];
            double vr;
            if (d.direct) {
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    x[e] = __ldcs(a.pos + 3 * c + e);
                    v[e] = __ldcs(a.vel + 3 * c + e);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    x[e] = s_pos[3 * i + e];
                    v[e] = s_vel[3 * i + e];
                }
            }
            frame(a, k, R, x, v, rh, &vr);
            Rec rec;
            rec.id = id[q];
            rec.rx = rh[0]; rec.ry = rh[1]; rec.rz = rh[2];
            rec.vr = sign_faithful(vr);
            rec.pos = (uint32_t)c;
            rec.angle = 0;                 // no match: the accumulator starts from zero
            rec.flags = 0;                 // (track_orbits.py:180-183, 344-346)
            a.mark_cur[c] = NO_EVENT;
            // (a partition that is full drops the record: reported, the host raises)
            if (slot[q] < cap[q]) store_rec(rec_cur + dst[q] + slot[q], rec);
            else atomicAdd(a.overflow, 1u);
        }
    }
}

// groups whose particles the tile [c0, c1) touches get one signal per consumer warp
OA_D void signal_groups(const oa_pj2_args& a, const Work& w, uint32_t c0, uint32_t c1) {
    int lo = 0, hi = a.n_groups - 1;           // last group with group_off[g] <= c0
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(a.group_off + mid) <= c0) lo = mid; else hi = mid - 1;
    }
    for (int g = lo; g < a.n_groups; ++g) {
        const uint32_t g0 = __ldg(a.group_off + g), g1 = __ldg(a.group_off + g + 1);
        if (g0 >= c1) break;
        if (g1 > c0 && g1 > g0) red_release(w.done + g, 1u);
    }
}

// ---- JOIN: consumer side ----------------------------------------------------------------------
// index entry: generation (8 bits) | fingerprint (14 bits) | record index (10 bits);
// an entry of another generation is an empty slot
OA_D void join_item(const oa_pj2_args& a, const Desc& d, const unsigned char* stage,
                    uint32_t* tab, int tid) {
    const Rec* A = reinterpret_cast<const Rec*>(stage);
    const Rec* B = reinterpret_cast<const Rec*>(stage + CAP * 32);
    const uint32_t gen = d.gen << 24;
    const int nA = (int)d.nA, nB = (int)d.nB;
    // build
#pragma unroll
    for (int q = 0; q < JMAX; ++q) {
        const int i = tid + q * CONSUMERS;
        if (i < nA) {
            const uint32_t h = (uint32_t)oa_mix64((uint64_t)A[i].id);
            const uint32_t val = gen | (((h >> 11) & 0x3FFFu) << 10) | (uint32_t)i;
            uint32_t s = h & (SLOTS - 1);
            for (;;) {
                const uint32_t v = tab[s];
                if ((v & 0xFF000000u) != gen) {
                    const uint32_t old = atomicCAS(&tab[s], v, val);
                    if (old == v) break;
                    if ((old & 0xFF000000u) != gen) continue;      // lost to a stale writer: retry
                }
                s = (s + 1) & (SLOTS - 1);
            }
        }
    }
    consumer_barrier();
    // probe
    uint32_t* out = static_cast<uint32_t*>(a.rec_cur) + 8 * (size_t)d.slot0;
#pragma unroll
    for (int q = 0; q < JMAX; ++q) {
        const int i = tid + q * CONSUMERS;
        if (i < nB) {
            const uint4 b0 = *reinterpret_cast<const uint4*>(&B[i]);
            const uint4 b1 = *(reinterpret_cast<const uint4*>(&B[i]) + 1);
            const int64_t id = (int64_t)(((uint64_t)b0.y << 32) | b0.x);
            const uint32_t h = (uint32_t)oa_mix64((uint64_t)id);
            const uint32_t fp = ((h >> 11) & 0x3FFFu) << 10;
            uint32_t s = h & (SLOTS - 1);
            int hit = -1;
            for (;;) {
                const uint32_t v = tab[s];
                if ((v & 0xFF000000u) != gen) break;        // newly entered: accumulator stays 0
                if ((v & 0x00FFFC00u) == fp) {
                    const int idx = (int)(v & 0x3FFu);
                    if (A[idx].id == id) { hit = idx; break; }
                }
                s = (s + 1) & (SLOTS - 1);
            }
            if (hit >= 0) {
                // compare_radial_velocities + calc_angles (track_orbits.py:311-325, 330-351)
                const uint4 a0 = *reinterpret_cast<const uint4*>(&A[hit]);
                const uint4 a1 = *(reinterpret_cast<const uint4*>(&A[hit]) + 1);
                const float pr[3] = {__uint_as_float(a0.z), __uint_as_float(a0.w),
                                     __uint_as_float(a1.x)};
                const float cr[3] = {__uint_as_float(b0.z), __uint_as_float(b0.w),
                                     __uint_as_float(b1.x)};
                const float pvr = __uint_as_float(a1.y), cvr = __uint_as_float(b1.y);
                const float dang = acosf(dot3f(pr, cr));
                bool ev;
                if (a.mode == OA_MODE_PERICENTRIC) ev = (pvr < 0) && (cvr > 0);
                else ev = (pvr > 0) && (cvr < 0);
                float run = fadd(__half2float(__ushort_as_half((uint16_t)(a1.w & 0xFFFFu))), dang);
                if (ev) {
                    a.mark_prev[a1.z] = __half_as_ushort(__float2half_rn(run));
                    run = 0.0f;
                }
                out[8 * (size_t)i + 7] = (uint32_t)__half_as_ushort(__float2half_rn(run));
            }
        }
    }
}

// ---- PRODUCER (the last warp) ---------------------------------------------------------------------
// PB lanes resolve PB consecutive tickets side by side (ticket -> item -> plan row
// -> dependency -> partition fills: four dependent global round trips that one
// lane alone could not hide behind a ~3000-cycle item), then hand their stages to
// the consumers in ticket order.
constexpr int PB = 4;

struct Resolved {
    uint32_t kind, region, idx, nA, nB, direct, slot0, group, need;
    const void* srcA;
    const void* srcB;
    oa_pj2_region pl;
    bool ready;              // dependency satisfied and fills read
};

// the fills of a JOIN's two partitions (valid once the group's tiles are finished)
OA_D void read_fills(const oa_pj2_args& a, Resolved& r) {
    fence_proxy_async();      // the records were written through the generic proxy
    const oa_pj2_region& pl = r.pl;
    const uint32_t pp = r.idx >> pl.shift;
    const uint32_t nB = __ldcg(a.fill_cur + pl.pb_cur + r.idx);
    const uint32_t nA = __ldg(a.fill_prev + pl.pb_prev + pp);
    r.nB = nB < pl.cap_cur ? nB : pl.cap_cur;
    r.nA = nA < pl.cap_prev ? nA : pl.cap_prev;
    r.slot0 = pl.base_cur + r.idx * pl.cap_cur;
    r.srcA = static_cast<const Rec*>(a.rec_prev) + ((size_t)pl.base_prev + (size_t)pp * pl.cap_prev);
    r.srcB = static_cast<const Rec*>(a.rec_cur) + (size_t)r.slot0;
    r.ready = true;
}

// everything that needs no waiting; a JOIN whose dependency is already satisfied
// (the usual case: its tiles are a whole group of tickets behind) is complete
OA_D void resolve(const oa_pj2_args& a, const Const& k, const Work& w, uint32_t t, Resolved& r) {
    const uint2 it = w.items[t];
    r.kind = it.x >> 31;
    r.region = it.x & 0x7FFFFFFFu;
    r.idx = it.y;
    r.slot0 = 0;
    if (r.kind == SCATTER) {
        const int64_t c0 = (int64_t)r.idx * TILE;
        const int64_t left = a.n_cur - c0;
        r.nB = (uint32_t)(left < TILE ? left : TILE);
        r.nA = 0;
        r.direct = r.nB < (uint32_t)TILE;       // copy sizes must be multiples of 16 bytes
        r.ready = true;
        return;
    }
    r.pl = a.plan[r.region];
    r.direct = 0;
    r.group = r.pl.group;
    // dependency: every tile that touches the group's particles is finished
    const uint32_t g0 = __ldg(a.group_off + r.group), g1 = __ldg(a.group_off + r.group + 1);
    r.need = g1 > g0 ? (ceil_div(g1, TILE) - g0 / TILE) * k.need_scale : 0u;
    r.ready = false;
    if (ld_acquire(w.done + r.group) >= r.need) read_fills(a, r);
}

OA_D void wait_dependency(const oa_pj2_args& a, const Work& w, Resolved& r) {
    uint32_t spins = 0;
    const uint64_t t0 = now();
    while (ld_acquire(w.done + r.group) < r.need) {
        __nanosleep(64);
        if (++spins > MAX_SPINS) __trap();
    }
    stat_add(4, now() - t0);
    read_fills(a, r);
}

__global__ void __launch_bounds__(THREADS, OA_PJ2_MIN_CTAS)
oa_pj2_kernel(const __grid_constant__ oa_pj2_args a, const __grid_constant__ Const k,
              const __grid_constant__ Work w) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    uint64_t* empty = full + 2;
    Desc* desc = reinterpret_cast<Desc*>(smem + SM_DESC);
    uint32_t* index = reinterpret_cast<uint32_t*>(smem + SM_INDEX);
    const int tid = (int)threadIdx.x;

    for (int s = tid; s < 2 * SLOTS; s += THREADS) index[s] = 0;     // generation 0 = never used
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&empty[0], CWARPS);
        mbar_init(&empty[1], CWARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t t_start = now();

    if (tid >= CONSUMERS) {
        // ===================== producer =====================
        const int lane = tid - CONSUMERS;
        const uint64_t pol_first = policy_evict_first();
        uint32_t n = 0;                      // stages handed to the consumers
        uint32_t njoin[2] = {0, 0};          // JOINs that used the index table of a stage
        for (;;) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(w.ticket, (uint32_t)PB);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            const uint32_t t = base + (uint32_t)lane;
            const bool mine = lane < PB && t < k.total_tickets;
            Resolved r;
            r.kind = EXIT;
            r.ready = true;
            r.nA = r.nB = 0;
            if (mine) resolve(a, k, w, t, r);
            // stages are handed over in ticket order, one lane after the other (an
            // unsatisfied dependency is waited for only now: every earlier ticket
            // of this warp has been issued, so the wait cannot be for one of them)
            for (int l = 0; l < PB; ++l) {
                bool use = false;
                if (lane == l && mine) {
                    if (!r.ready) wait_dependency(a, w, r);
                    use = r.kind == SCATTER || (r.nA != 0 && r.nB != 0);   // else: nothing to match
                    if (use) {
                        const uint32_t s = n & 1u;
                        {
                            const uint64_t t0 = now();
                            mbar_wait(&empty[s], ((n >> 1) & 1u) ^ 1u);
                            stat_add(5, now() - t0);
                        }
                        Desc d;
                        d.kind = r.kind; d.region = r.region; d.idx = r.idx;
                        d.nA = r.nA; d.nB = r.nB; d.direct = r.direct;
                        d.slot0 = r.slot0;
                        d.gen = 0;
                        unsigned char* st = smem + SM_STAGE + s * STAGE_BYTES;
                        if (r.kind == JOIN) {
                            d.gen = njoin[s] % 255u + 1u;
                            desc[s] = d;
                            mbar_arrive_expect_tx(&full[s], (r.nA + r.nB) * 32u);
                            tma_load_hint(st, r.srcA, r.nA * 32u, &full[s], pol_first);
                            tma_load(st + CAP * 32, r.srcB, r.nB * 32u, &full[s]);
                        } else if (r.direct) {
                            desc[s] = d;
                            mbar_arrive(&full[s]);
                        } else {
                            desc[s] = d;
                            const int64_t c0 = (int64_t)r.idx * TILE;
                            mbar_arrive_expect_tx(&full[s], (uint32_t)TILE * 32u);
                            tma_load_hint(st, a.ids + c0, TILE * 8u, &full[s], pol_first);
                            tma_load_hint(st + TILE * 8, a.pos + 3 * c0, TILE * 12u, &full[s],
                                          pol_first);
                            tma_load_hint(st + TILE * 20, a.vel + 3 * c0, TILE * 12u, &full[s],
                                          pol_first);
                        }
                    }
                }
                const unsigned used = __ballot_sync(0xFFFFFFFFu, use);
                if (used) {
                    const unsigned isjoin = __ballot_sync(0xFFFFFFFFu, use && r.kind == JOIN);
                    if (isjoin) ++njoin[n & 1u];
                    ++n;
                }
            }
            if (base + PB > k.total_tickets) {       // the ticket counter ran out: last batch
                if (lane == 0) {
                    const uint32_t s = n & 1u;
                    mbar_wait(&empty[s], ((n >> 1) & 1u) ^ 1u);
                    Desc d;
                    d.kind = EXIT;
                    d.region = d.idx = d.gen = d.nA = d.nB = d.direct = d.slot0 = 0;
                    desc[s] = d;
                    mbar_arrive(&full[s]);
                }
                break;
            }
        }
        return;
    }

    // ===================== consumers =====================
    const int lane = tid & 31;
    for (uint32_t n = 0;; ++n) {
        const uint32_t s = n & 1u;
        {
            const uint64_t t0 = now();
            mbar_wait(&full[s], (n >> 1) & 1u);
            if (tid == 0) stat_add(6, now() - t0);
        }
        const Desc d = desc[s];
        if (d.kind == EXIT) break;
        const unsigned char* st = smem + SM_STAGE + s * STAGE_BYTES;
        const uint64_t t0 = now();
        if (d.kind == SCATTER) {
            scatter_tile(a, k, w, d, st, tid);
            // this warp's records are written (generic proxy); the JOIN that waits for
            // the group counter reads them with TMA (async proxy)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&empty[s]);
                fence_proxy_async();
                const uint32_t c0 = d.idx * (uint32_t)TILE;
                signal_groups(a, w, c0, c0 + d.nB);
            }
        } else {
            join_item(a, d, st, index + s * SLOTS, tid);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (tid == 0) {
            stat_add((int)d.kind, now() - t0);
            stat_add(8 + (int)d.kind, 1);
        }
    }
    if (tid == 0) stat_add(12, now() - t_start);
}

// item list in ticket order: range 2 s = SCATTER of group s (the tiles that START
// in the group's particle range), range 2 s + 1 = JOIN of group s - LAG
__global__ void oa_pj2_expand_kernel(const __grid_constant__ oa_pj2_args a, uint2* items,
                                     uint32_t total) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int lo = 0, hi = a.n_ranges - 1;            // last range with range_start[r] <= t
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.range_start[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const uint32_t i = t - a.range_start[lo];
    if ((lo & 1) == 0) {
        const int g = lo >> 1;
        const uint32_t tile = ceil_div(a.group_off[g], TILE) + i;
        const int64_t c0 = (int64_t)tile * TILE;
        int jl = 0, jh = a.n_regions - 1;       // last region with cur_begin <= c0
        while (jl < jh) {
            const int mid = (jl + jh + 1) >> 1;
            if (a.regions[mid].cur_begin <= c0) jl = mid; else jh = mid - 1;
        }
        items[t] = make_uint2((uint32_t)jl, tile);
    } else {
        const int g = (lo >> 1) - OA_PJ2_LAG;
        int jl = (int)a.group_first[g], jh = (int)a.group_first[g + 1] - 1;
        const uint32_t target = a.plan[jl].join_first + i;
        while (jl < jh) {                        // last region with join_first <= target
            const int mid = (jl + jh + 1) >> 1;
            if (a.plan[mid].join_first <= target) jl = mid; else jh = mid - 1;
        }
        items[t] = make_uint2(0x80000000u | (uint32_t)jl, target - a.plan[jl].join_first);
    }
}

size_t work_words(int n_groups) { return 4 + (size_t)(n_groups > 0 ? n_groups : 0); }

}  // namespace pj2

extern "C" size_t oa_pj2_workspace_bytes(int n_groups, uint32_t total_tickets) {
    return 8 * (size_t)total_tickets + 4 * pj2::work_words(n_groups);
}

extern "C" int oa_pj2_stats(uint64_t* out16, int reset) {
    if (!out16) return 0;
    unsigned long long h[16];
    if (cudaMemcpyFromSymbol(h, g_pj2_stats, sizeof(h)) != cudaSuccess) return 0;
    for (int i = 0; i < 16; ++i) out16[i] = h[i];
    if (reset) {
        for (int i = 0; i < 16; ++i) h[i] = 0;
        if (cudaMemcpyToSymbol(g_pj2_stats, h, sizeof(h)) != cudaSuccess) return 0;
    }
    return OA_PJ2_STATS;
}

extern "C" size_t oa_pj2_args_size(void) { return sizeof(oa_pj2_args); }

extern "C" void oa_pj2_config(int32_t* out8) {
    out8[0] = pj2::CONSUMERS;
    out8[1] = OA_PJ2_MIN_CTAS;
    out8[2] = pj2::TILE;
    out8[3] = pj2::CAP;
    out8[4] = OA_PJ2_TARGET;
    out8[5] = OA_PJ2_SIGMAS;
    out8[6] = OA_PJ2_LAG;
    out8[7] = pj2::SM_BYTES;
}

// The plan on the host: partition counts (powers of two that never shrink for a
// halo), capacities, record-slot and fill-entry layout, groups, ticket ranges.
extern "C" int oa_pj2_plan_host(const int64_t* offsets, int n_regions, const uint32_t* prev_P,
                                const uint32_t* prev_cap, const uint32_t* prev_base,
                                const uint32_t* prev_pb, int64_t group_particles, int32_t target,
                                oa_pj2_region* rows, uint32_t* P_out, uint32_t* cap_out,
                                uint32_t* base_out, uint32_t* pb_out, uint32_t* group_first,
                                uint32_t* group_off, uint32_t* range_start,
                                oa_pj2_plan_info* info) {
    OA_REQUIRE(n_regions >= 0 && group_particles >= 1 && rows && group_first && group_off &&
               range_start && info &&
               (n_regions == 0 || (offsets && prev_P && prev_cap && prev_base && prev_pb &&
                                   P_out && cap_out && base_out && pb_out)),
               "oa_pj2_plan_host: bad arguments");
    if (target <= 0) target = OA_PJ2_TARGET;
    uint64_t pb = 0, base = 0, joins = 0;
    uint32_t max_P = 0;
    int n_groups = 0;
    int64_t gid_prev = -1;
    for (int j = 0; j < n_regions; ++j) {
        const int64_t len = offsets[j + 1] - offsets[j];
        OA_REQUIRE(len >= 0, "oa_pj2_plan_host: offsets must not decrease");
        uint64_t P = 1;
        while ((int64_t)(P * (uint64_t)target) < len) P <<= 1;
        if (prev_P[j] > P) P = prev_P[j];
        OA_REQUIRE((P & (P - 1)) == 0 && P <= (1u << 30), "oa_pj2_plan_host: bad partition count");
        const uint64_t mean = ((uint64_t)len + P - 1) / P;
        uint64_t cap = mean + (uint64_t)(OA_PJ2_SIGMAS * sqrt((double)mean)) + 16;
        if (cap > OA_PJ2_CAP) cap = OA_PJ2_CAP;
        OA_REQUIRE(mean <= (uint64_t)OA_PJ2_CAP, "oa_pj2_plan_host: target exceeds OA_PJ2_CAP");
        oa_pj2_region& r = rows[j];
        r.P_cur = (uint32_t)P;
        r.cap_cur = (uint32_t)cap;
        r.base_cur = (uint32_t)base;
        r.pb_cur = (uint32_t)pb;
        r.P_prev = prev_P[j];
        r.cap_prev = prev_cap[j];
        r.base_prev = prev_base[j];
        r.pb_prev = prev_pb[j];
        r.shift = 0;
        if (prev_P[j]) {
            OA_REQUIRE((prev_P[j] & (prev_P[j] - 1)) == 0, "oa_pj2_plan_host: bad prev_P");
            while (((uint64_t)prev_P[j] << r.shift) < P) ++r.shift;
        }
        r.join_first = (uint32_t)joins;
        r.reserved = 0;
        P_out[j] = (uint32_t)P; cap_out[j] = (uint32_t)cap;
        base_out[j] = (uint32_t)base; pb_out[j] = (uint32_t)pb;
        if (P > max_P) max_P = (uint32_t)P;
        const int64_t gid = offsets[j] / group_particles;
        if (j == 0 || gid != gid_prev) {
            group_first[n_groups] = (uint32_t)j;
            group_off[n_groups] = (uint32_t)offsets[j];
            ++n_groups;
        }
        gid_prev = gid;
        r.group = (uint32_t)(n_groups - 1);
        pb += P;
        base += P * cap;
        if (prev_P[j] && len > 0) joins += P;
    }
    const int64_t n = n_regions ? offsets[n_regions] : 0;
    OA_REQUIRE(pb < ((uint64_t)1 << 32) && base < ((uint64_t)1 << 32) && joins < ((uint64_t)1 << 31) &&
               n < ((int64_t)1 << 32),
               "oa_pj2_plan_host: plan does not fit 32-bit counters");
    oa_pj2_region& e = rows[n_regions];
    e = oa_pj2_region{};
    e.P_cur = 1;
    e.base_cur = (uint32_t)base;
    e.pb_cur = (uint32_t)pb;
    e.join_first = (uint32_t)joins;
    e.group = (uint32_t)n_groups;
    group_first[n_groups] = (uint32_t)n_regions;
    group_off[n_groups] = (uint32_t)n;

    const int n_ranges = 2 * (n_groups + OA_PJ2_LAG);
    uint64_t t = 0;
    for (int s = 0; s < n_groups + OA_PJ2_LAG; ++s) {
        range_start[2 * s] = (uint32_t)t;
        if (s < n_groups) {
            const uint64_t g0 = group_off[s], g1 = group_off[s + 1];
            t += (g1 + OA_PJ2_TILE - 1) / OA_PJ2_TILE - (g0 + OA_PJ2_TILE - 1) / OA_PJ2_TILE;
        }
        range_start[2 * s + 1] = (uint32_t)t;
        const int g = s - OA_PJ2_LAG;
        if (g >= 0) t += rows[group_first[g + 1]].join_first - rows[group_first[g]].join_first;
    }
    OA_REQUIRE(t < ((uint64_t)1 << 32), "oa_pj2_plan_host: too many work items");
    range_start[n_ranges] = (uint32_t)t;
    info->n_part_entries = (int64_t)pb;
    info->n_rec_slots = (int64_t)base;
    info->total_tickets = (uint32_t)t;
    info->n_groups = n_groups;
    info->n_ranges = n_ranges;
    info->max_P = max_P;
    return OA_OK;
}

extern "C" int oa_pj2_step(const oa_pj2_args* args, void* stream) {
    OA_REQUIRE(args != nullptr, "oa_pj2_step: args is NULL");
    const oa_pj2_args& a = *args;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OA_REQUIRE(a.n_cur >= 0 && a.n_regions >= 0 && a.n_groups >= 0, "oa_pj2_step: negative size");
    OA_REQUIRE(a.n_cur < ((int64_t)1 << 32) && a.n_prev < ((int64_t)1 << 32),
               "oa_pj2_step: more than 2^32-1 region-particles on one GPU");
    OA_REQUIRE(a.mode == OA_MODE_PERICENTRIC || a.mode == OA_MODE_APOCENTRIC,
               "oa_pj2_step: bad mode %d", a.mode);
    OA_REQUIRE(a.centre_f32 == 1, "oa_pj2_step: float32 region centres only");
    if (a.n_regions == 0 || a.total_tickets == 0) return OA_OK;
    OA_REQUIRE(a.regions && a.plan && a.group_first && a.group_off && a.range_start &&
               a.rec_cur && a.fill_cur && a.mark_cur && a.workspace && a.overflow,
               "oa_pj2_step: NULL required pointer");
    OA_REQUIRE(a.n_cur == 0 || (a.pos && a.vel && a.ids), "oa_pj2_step: NULL input array");
    OA_REQUIRE(a.n_ranges == 2 * (a.n_groups + OA_PJ2_LAG), "oa_pj2_step: bad n_ranges");
    OA_REQUIRE(a.n_prev == 0 || !a.rec_prev || (a.fill_prev && a.mark_prev),
               "oa_pj2_step: previous generation incomplete");
    OA_REQUIRE(a.workspace_bytes >= oa_pj2_workspace_bytes(a.n_groups, a.total_tickets),
               "oa_pj2_step: workspace too small (need oa_pj2_workspace_bytes)");
    OA_REQUIRE((reinterpret_cast<uintptr_t>(a.rec_cur) & 31u) == 0 &&
               (reinterpret_cast<uintptr_t>(a.rec_prev) & 31u) == 0 &&
               (reinterpret_cast<uintptr_t>(a.workspace) & 7u) == 0,
               "oa_pj2_step: records must be 32-byte aligned, the workspace 8-byte aligned");
    OA_REQUIRE(((reinterpret_cast<uintptr_t>(a.pos) | reinterpret_cast<uintptr_t>(a.vel) |
                 reinterpret_cast<uintptr_t>(a.ids)) & 15u) == 0,
               "oa_pj2_step: input arrays must be 16-byte aligned (TMA bulk copies)");

    static bool configured = false;
    if (!configured) {
        OA_CUDA_CHECK(cudaFuncSetAttribute(pj2::oa_pj2_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           pj2::SM_BYTES));
        configured = true;
    }
    int dev = 0, sms = OA_NUM_SMS, per_sm = 0;
    OA_CUDA_CHECK(cudaGetDevice(&dev));
    OA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    OA_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pj2::oa_pj2_kernel,
                                                                pj2::THREADS, pj2::SM_BYTES));
    OA_REQUIRE(per_sm >= 1, "oa_pj2_step: the kernel does not fit an SM");
    if (a.sm_reserve > 0 && a.sm_reserve < sms) sms -= a.sm_reserve;

    pj2::Const k;
    for (int q = 0; q < 3; ++q) {
        const double h = a.box[q] * 0.5;           // largest float <= L/2
        float hf = (float)h;
        if ((double)hf > h) hf = nextafterf(hf, -INFINITY);
        k.half_box[q] = hf;
    }
    k.total_tickets = a.total_tickets;
    k.need_scale = pj2::CWARPS;

    pj2::Work w;
    w.items = static_cast<uint2*>(a.workspace);
    uint32_t* ws = reinterpret_cast<uint32_t*>(w.items + a.total_tickets);
    OA_CUDA_CHECK(cudaMemsetAsync(ws, 0, 4 * pj2::work_words(a.n_groups), st));
    OA_CUDA_CHECK(cudaMemsetAsync(a.fill_cur, 0, 4 * (size_t)a.n_part_entries, st));
    w.ticket = ws;
    w.done = ws + 4;
    pj2::oa_pj2_expand_kernel<<<(a.total_tickets + 255) / 256, 256, 0, st>>>(a, w.items,
                                                                             a.total_tickets);
    OA_LAUNCH_CHECK();
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > (int64_t)a.total_tickets) grid = a.total_tickets;
    pj2::oa_pj2_kernel<<<(unsigned)grid, pj2::THREADS, pj2::SM_BYTES, st>>>(a, k, w);
    OA_LAUNCH_CHECK();
    return OA_OK;
}
