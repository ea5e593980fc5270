// Fused per-snapshot orbit-tracking kernel (sm_100a).
//
// One pass over the current snapshot's region blocks does the whole of the
// reference's per-halo `track(j)` (track_orbits.py:147-185): halo frame,
// ID match against the previous block, apsis detection, float16 angle update.
// See include/orbit_b200.h (oa_track_fused) for the contract and DESIGN.md for
// the memory layout and the roofline accounting.
//
// Structure: a persistent kernel, one 768-thread CTA per SM (80 registers per
// thread, nothing spills), in which every WARP is an autonomous software
// pipeline over chunks of 32 consecutive particles (chunk c -> warp c mod
// #warps, so the warps of the whole GPU sweep the snapshot as one compact
// front).  A warp owns a ring of D shared-memory slots with one mbarrier each;
// lane 0 fills slot t mod D with four 1-D TMA bulk copies (cp.async.bulk ...
// mbarrier::complete_tx, L2 evict-first): ids, positions, velocities and the
// oa_region rows the chunk touches.  Per iteration a warp
//   A. hashes the IDs of the NEXT chunk and issues its insert atomic and its
//      bucket load (they are in flight during everything below),
//   B. compares the 8 slots of THIS chunk's bucket (loaded an iteration ago) and
//      issues the gather of the candidate's previous record,
//   C. does the halo-frame arithmetic of this chunk while that gather is out,
//      hands the slot back to TMA (chunk t+D),
//   D. verifies the ID, tests for an apsis, updates the angle, writes the record.
// There is no block-wide barrier and no inter-warp wait.
//
// What bounds it (tools/micro/random_access.cu, profiles/r01_ablation.md): a
// warp access whose 32 lanes hit 32 different lines costs the SM's L1TEX ~2.4
// cycles per lane, whatever the level it hits.  13.5 M random 32 B gathers
// alone take 0.117 ms on B200, 13.5 M atomics 0.129 ms, 13.5 M scattered 4 B
// stores 0.09 ms, the streaming part 0.14 ms; the kernel's four random
// accesses per particle (bucket, record, counter, slot) put its floor near
// 0.45 ms and it runs in 0.535 ms.
//
// Matching: a region-segmented, bucketised hash table in global memory.  The
// table of a snapshot is two arrays: per-bucket fill counters (4 B / bucket,
// ~1 B / particle: L2-resident, so clearing them and the atomics on them cost
// next to no HBM traffic) and 32-byte slot buckets (8 slots = fingerprint |
// block-local index; never cleared -- a stale slot can only produce a candidate
// that the 64-bit ID check of the record rejects).
//   insert : k = atomicAdd(count[b], 1); slot[b][k] = value
//   probe  : one 32 B sector (ld.global.v8), XOR/compare in registers, then
//            the 32 B record of the candidate (one more v8 load), which carries
//            the ID for the exact check AND rhat / v_r / angle: match + state
//            read is one gather.  Overflow (count > 8) spills to the next
//            bucket of the same region and is resolved out of line.
// The buckets of a region are touched while the front crosses that region,
// i.e. from B200's 126 MB L2; data that is touched once is loaded/stored with
// an evict-first policy.
//
// Arithmetic mirrors numpy's evaluation order and rounding points (no FMA
// contraction: compiled with -fmad=false and *_rn intrinsics where the order
// matters) -- see SURVEY.md 2.2 / 7.4-7.6.
#include "oa_common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace {

#ifndef OA_TRACK_WARPS
#define OA_TRACK_WARPS 24      // 768 threads: 80 registers each, nothing spills
#endif
constexpr int TRACK_WARPS = OA_TRACK_WARPS;     // one CTA per SM
constexpr int TRACK_THREADS = 32 * TRACK_WARPS;
constexpr int CHUNK = 32;                       // particles per warp step
constexpr int RW = 4;                           // region rows staged per chunk

// ---- IEEE arithmetic without contraction ------------------------------------------
template <typename T> struct Ar;
template <> struct Ar<float> {
    static OA_D float add(float a, float b) { return __fadd_rn(a, b); }
    static OA_D float mul(float a, float b) { return __fmul_rn(a, b); }
    static OA_D float div(float a, float b) { return __fdiv_rn(a, b); }
    static OA_D float sqrt(float a) { return __fsqrt_rn(a); }
    static OA_D float acos(float a) { return acosf(a); }
    // numpy einsum('...i,...i') over 3 float32 terms: (p0 + p1) + p2
    static OA_D float dot3(float a0, float a1, float a2, float b0, float b1,
                           float b2) {
        return add(add(mul(a0, b0), mul(a1, b1)), mul(a2, b2));
    }
    static OA_D __half to_half(float a) { return __float2half_rn(a); }
};
template <> struct Ar<double> {
    static OA_D double add(double a, double b) { return __dadd_rn(a, b); }
    static OA_D double mul(double a, double b) { return __dmul_rn(a, b); }
    static OA_D double div(double a, double b) { return __ddiv_rn(a, b); }
    static OA_D double sqrt(double a) { return __dsqrt_rn(a); }
    static OA_D double acos(double a) { return ::acos(a); }
    // numpy einsum('...i,...i') over 3 float64 terms: (p0 + p2) + p1
    static OA_D double dot3(double a0, double a1, double a2, double b0,
                            double b1, double b2) {
        return add(add(mul(a0, b0), mul(a2, b2)), mul(a1, b1));
    }
    static OA_D __half to_half(double a) { return __double2half(a); }
};

// float copy of v_r whose `< 0` / `> 0` tests agree with the float64 value
OA_D float sign_faithful(double v) {
    float f = (float)v;
    if (f == 0.0f && v != 0.0) f = (v > 0.0) ? 1.401298464e-45f : -1.401298464e-45f;
    return f;
}
OA_D float sign_faithful(float v) { return v; }

template <typename TF, typename TVR>
OA_D void set_vr(OaRec<TF>& rec, TVR vr) {
    if constexpr (std::is_same<TF, float>::value) rec.vr = sign_faithful(vr);
    else rec.vr = (double)vr;
}

// ---- L2 cache-policy hinted memory operations ---------------------------------------
OA_D uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
OA_D uint64_t policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
OA_D uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
struct __align__(32) V8 { uint32_t w[8]; };
// one whole 32-byte sector per thread (sm_100 256-bit global access)
OA_D V8 ld32_nc(const void* a, uint64_t pol) {
    V8 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]),
                   "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7])
                 : "l"(a), "l"(pol));
    return v;
}
OA_D void st32(void* a, const V8& v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v8.u32 [%0], "
                 "{%1,%2,%3,%4,%5,%6,%7,%8}, %9;"
                 :: "l"(a), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]),
                    "r"(v.w[4]), "r"(v.w[5]), "r"(v.w[6]), "r"(v.w[7]), "l"(pol)
                 : "memory");
}
OA_D int64_t ld8_nc(const int64_t* a, uint64_t pol) {
    int64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s64 %0, [%1], %2;"
                 : "=l"(v) : "l"(a), "l"(pol));
    return v;
}
OA_D float ld_elem(const float* a, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
    return v;
}
OA_D double ld_elem(const double* a, uint64_t pol) {
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol));
    return v;
}
OA_D void st2(uint16_t* a, uint16_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u16 [%0], %1, %2;" :: "l"(a), "h"(v), "l"(pol) : "memory");
}

template <typename TF>
OA_D OaRec<TF> load_rec(const OaRec<TF>* p, uint64_t pol) {
    OaRec<TF> r;
    V8* dst = reinterpret_cast<V8*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(OaRec<TF>) / 32); ++i)
        dst[i] = ld32_nc(reinterpret_cast<const V8*>(p) + i, pol);
    return r;
}
template <typename TF>
OA_D void store_rec(OaRec<TF>* p, const OaRec<TF>& r, uint64_t pol) {
    const V8* src = reinterpret_cast<const V8*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(OaRec<TF>) / 32); ++i)
        st32(reinterpret_cast<V8*>(p) + i, src[i], pol);
}

// ---- mbarrier / TMA (1-D bulk copy) -----------------------------------------------------
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
OA_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n"
                     " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     " selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
OA_D void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                      uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
                 ".L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}

// ---- table geometry (closed forms, shared with the host through the C ABI) ------------
// Region j (block start `off`, length len) owns buckets
//   [off/3 + j, off/3 + j + len/3 + 1): mean fill <= 3 of 8 slots, so that only
// 0.4 % of the buckets overflow (Poisson) -- an overflow costs its whole warp a
// second, serialised round trip.
OA_HD uint32_t bucket_count(int64_t len) {          // len < 2^31 on one GPU
    return (uint32_t)len / (uint32_t)OA_BUCKET_LOAD + 1u;
}
OA_HD int64_t bucket_begin(int64_t block_start, int64_t region) {
    return block_start / OA_BUCKET_LOAD + region;
}
OA_HD int64_t table_buckets(int64_t n, int64_t n_regions) {
    return n / OA_BUCKET_LOAD + n_regions + 2;
}
// counters first (rounded up to whole sectors), then the slot buckets
OA_HD int64_t table_count_words(int64_t buckets) { return (buckets + 7) / 8 * 8; }

// ---- region rows ------------------------------------------------------------------------------
// The kernel reads oa_region rows (128 B, include/orbit_b200.h) as staged by
// TMA, or -- for chunks that touch more than RW regions -- a private copy.
using Row = oa_region;
static_assert(sizeof(Row) == 128, "oa_region layout");

OA_D void load_row(const oa_region* __restrict__ regions, int j, Row* out) {
    const int4* src = reinterpret_cast<const int4*>(regions + j);
    int4* dst = reinterpret_cast<int4*>(out);
#pragma unroll
    for (int q = 0; q < (int)(sizeof(Row) / 16); ++q) dst[q] = __ldg(src + q);
}

// last j in [lo, hi] with off[j] <= c   (off is non-decreasing; empty blocks
// share their start with the next block and are skipped by taking the last)
OA_D int find_region(const int64_t* __restrict__ off, int lo, int hi, int64_t c) {
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(off + mid) <= c) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// L2 eviction policies, created once per thread (warp-uniform values)
struct Policies {
    uint64_t first, normal, last;
};

struct ChunkMeta {
    int jlo, jhi;        // regions the chunk touches
    int nrows;           // rows staged in the slot (0: more than RW, look up)
    int tma;             // inputs were staged by TMA
};

// ---- derived launch constants (host -> kernel) ----------------------------------------------
struct TrackConst {
    const int2* chunk_regions;    // (n_chunks,) first / last region of every chunk
    unsigned int* ticket;         // TICKETS counters, 32 words apart
    const uint32_t* cnt_prev;     // bucket fill counters of the previous table
    const uint32_t* slot_prev;    // its slot buckets
    uint32_t* cnt_cur;
    uint32_t* slot_cur;
    int n_chunks;
    float half_box[3];            // largest float <= L/2 (float-frame wrap test)
    int tma_ok;                   // base pointers are 16-byte aligned
};

// ---- rare paths, kept out of line so they do not cost registers -----------------------
// (plain loads here: ptxas 12.9 crashes on 256-bit inline-asm loads inside a
// function that is not inlined)
template <typename TF>
__device__ __noinline__ int64_t probe_slow(const uint32_t* __restrict__ cnt,
                                           const uint32_t* __restrict__ slot,
                                           uint32_t bucket0, uint32_t nb, uint32_t home,
                                           uint32_t fpshift, uint32_t prev_count,
                                           const OaRec<TF>* __restrict__ rec_prev,
                                           uint32_t prev_begin, int64_t id,
                                           bool home_has_no_match) {
    uint32_t b = home;
    for (uint32_t t = 0; t < nb; ++t) {
        const uint32_t fill = __ldcg(cnt + bucket0 + b);
        uint32_t m = fill < OA_BUCKET_SLOTS ? fill : OA_BUCKET_SLOTS;
        if (t == 0 && home_has_no_match) m = 0;      // already compared in registers
        const uint32_t* B = slot + (size_t)(bucket0 + b) * OA_BUCKET_WORDS;
        for (uint32_t e = 0; e < m; ++e) {
            const uint32_t x = __ldcg(B + e) ^ fpshift;
            if (x < prev_count) {
                const int64_t q = (int64_t)prev_begin + (int64_t)x;
                if (__ldcg(&rec_prev[q].id) == id) return q;
            }
        }
        if (fill <= OA_BUCKET_SLOTS) return -1;      // never overflowed: a miss
        b = (b + 1 == nb) ? 0u : b + 1;
    }
    return -1;
}

__device__ __noinline__ void insert_slow(uint32_t* __restrict__ cnt,
                                         uint32_t* __restrict__ slot, uint32_t bucket0,
                                         uint32_t nb, uint32_t home, uint32_t val) {
    uint32_t b = home;
    for (uint32_t t = 1; t < nb; ++t) {
        b = (b + 1 == nb) ? 0u : b + 1;
        const uint32_t k = atomicAdd(cnt + bucket0 + b, 1u);
        if (k < OA_BUCKET_SLOTS) {
            slot[(size_t)(bucket0 + b) * OA_BUCKET_WORDS + k] = val;
            return;
        }
    }
}

// Chunks beyond a warp's static first rounds are handed out by ticket.  One
// counter would serialise in L2 (~2 ns per same-address atomic: 0.4 M tickets per
// launch doubled the kernel time); TICKETS counters on separate 128-byte lines,
// counter c serving the chunks == c (mod TICKETS), are each hit 64 times less.
constexpr int TICKETS = 64;

// first / last region of every chunk (one thread per chunk)
__global__ void track_chunks_kernel(const int64_t* __restrict__ off, int n_regions,
                                    int64_t n, int n_chunks, int2* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    // chunk tickets of the tracking kernel: TICKETS counters, one per 128-byte line
    if (t < TICKETS * 16) out[n_chunks + t] = make_int2(0, 0);
    if (t == 0 && n_chunks < TICKETS * 16)
        for (int q = n_chunks; q < TICKETS * 16; ++q) out[n_chunks + q] = make_int2(0, 0);
    if (t >= n_chunks) return;
    const int64_t first = (int64_t)t * CHUNK;
    const int64_t last = min(first + (int64_t)CHUNK, n) - 1;
    const int jlo = find_region(off, 0, n_regions - 1, first);
    const int jhi = find_region(off, jlo, n_regions - 1, last);
    out[t] = make_int2(jlo, jhi);
}

template <typename TX>
struct SlotLayout {
    static constexpr int IDS = 0;
    static constexpr int POS = IDS + 8 * CHUNK;
    static constexpr int VEL = POS + (int)sizeof(TX) * 3 * CHUNK;
    static constexpr int ROWS = VEL + (int)sizeof(TX) * 3 * CHUNK;
    static constexpr int META = ROWS + (int)sizeof(Row) * RW;
    static constexpr int BYTES = META + 16;
    // ring depth: four chunks of float data, three of double data
    static constexpr int DEPTH = sizeof(TX) == 4 ? 4 : 3;
    static_assert(POS % 16 == 0 && VEL % 16 == 0 && ROWS % 16 == 0 && BYTES % 16 == 0,
                  "TMA destinations must be 16-byte aligned");
};

template <typename T> struct Vec3 { T x, y, z; };

// Per-particle state of the two overlapped stages.  `Probe` lives from the
// issue of the table traffic of a chunk (one iteration AHEAD of its frame
// arithmetic) to the end of that chunk; `Frame` only inside one iteration.
struct Probe {
    uint32_t ins_k, ins_val, ins_idx, ins_nb;     // insert into this snapshot's table
    uint32_t fpshift, prev_count, prev_begin;     // probe of the previous table
    uint32_t prb_idx, prb_nb;
    V8 bk;
};
template <typename TF, typename TVR>
struct Frame {
    TF rh[3], r;
    TVR vr;
};

// Stage A: hash the ID, issue the insert atomic on this snapshot's table and the
// load of the particle's bucket in the previous table.  Nothing here waits.
OA_D Probe stage_hash(const oa_track_args& a, const TrackConst& k, const Policies& pol,
                      const Row* R, const int64_t id, const uint32_t c) {
    Probe H;
    const uint64_t hsh = oa_mix64((uint64_t)id);
    const uint32_t h_slot = (uint32_t)(hsh >> 32), h_fp = (uint32_t)hsh;
    const int pbits = a.prev_index_bits, cbits = a.cur_index_bits;

    // (no L2 hint on the table writes: evict-last on them measured 20 % slower)
    H.ins_nb = bucket_count(R->cur_count);
    H.ins_idx = (uint32_t)R->cur_bucket + oa_slot(h_slot, H.ins_nb);
    H.ins_val = ((h_fp >> cbits) << cbits) | (c - (uint32_t)R->cur_begin);
    H.ins_k = atomicAdd(k.cnt_cur + H.ins_idx, 1u);

    H.prev_count = (a.rec_prev != nullptr && R->prev_count > 0) ? (uint32_t)R->prev_count : 0u;
    H.fpshift = (h_fp >> pbits) << pbits;
    H.prev_begin = (uint32_t)R->prev_begin;
    H.prb_nb = bucket_count(H.prev_count);
    H.prb_idx = (uint32_t)R->prev_bucket + oa_slot(h_slot, H.prb_nb);
#pragma unroll
    for (int e = 0; e < 8; ++e) H.bk.w[e] = 0xFFFFFFFFu;     // never a candidate
    if (H.prev_count > 0)
        H.bk = ld32_nc(k.slot_prev + (size_t)H.prb_idx * OA_BUCKET_WORDS, pol.last);
    return H;
}

OA_D Probe empty_probe() {
    Probe H;
    H.ins_k = H.ins_val = H.ins_idx = H.ins_nb = 0;
    H.fpshift = H.prev_count = H.prev_begin = H.prb_idx = H.prb_nb = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) H.bk.w[e] = 0xFFFFFFFFu;
    return H;
}

// row of the region of particle `c` of a chunk whose rows are staged at `rows`
// (`own`: private copy when the chunk touches more regions than a slot stages)
OA_D const Row* find_row(const oa_track_args& a, const unsigned char* rows,
                         const ChunkMeta& meta, const int64_t c, Row* own) {
    const Row* R = reinterpret_cast<const Row*>(rows);
    if (meta.nrows > 0) {
        // first staged row whose block ends beyond this particle (empty blocks in
        // between end where they begin and are skipped)
        for (int q = 1; q < meta.nrows; ++q)
            if (c >= R->cur_begin + R->cur_count) ++R;
        return R;
    }
    load_row(a.regions, find_region(a.cur_off, meta.jlo, meta.jhi, c), own);
    return own;
}

// Stage C: halo frame of one particle (region_frame).  `R` is the row of the
// particle's region (shared memory, or a private copy on the rare path).
template <typename TX, typename TF, typename TVR, bool HUBBLE>
OA_D void stage_frame(const oa_track_args& a, const TrackConst& k, const Row* R,
                      const Vec3<TX> xin, const Vec3<TX> vin, Frame<TF, TVR>& F) {
    using AF = Ar<TF>;
    const TX x[3] = {xin.x, xin.y, xin.z}, v[3] = {vin.x, vin.y, vin.z};
    TF d[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (std::is_same<TX, double>::value || !a.centre_f32) {
            double dd = __dsub_rn((double)x[q], R->centre[q]);
            if (a.periodic) {
                const double Lq = a.box[q], h = Lq * 0.5;
                if (dd > h) dd = __dsub_rn(dd, Lq);
                if (dd < -h) dd = __dadd_rn(dd, Lq);
            }
            d[q] = (TF)dd;
        } else {
            d[q] = (TF)__fsub_rn((float)x[q], R->centre_f[q]);
        }
    }
    if (!(std::is_same<TX, double>::value || !a.centre_f32) && a.periodic) {
        // float frame: `(double)df > L/2` <=> `df > rd_float(L/2)`; only the few
        // particles of a halo that straddles a box face take the branch
        const bool out = fabsf((float)d[0]) > k.half_box[0] ||
                         fabsf((float)d[1]) > k.half_box[1] ||
                         fabsf((float)d[2]) > k.half_box[2];
        if (out) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                float df = (float)d[q];
                const float hf = k.half_box[q];
                if (df > hf) df = (float)__dsub_rn((double)df, a.box[q]);
                if (df < -hf) df = (float)__dadd_rn((double)df, a.box[q]);
                d[q] = (TF)df;
            }
        }
    }
    F.r = AF::sqrt(AF::dot3(d[0], d[1], d[2], d[0], d[1], d[2]));
#pragma unroll
    for (int q = 0; q < 3; ++q) F.rh[q] = AF::div(d[q], F.r);
    TVR w[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double wk;
        if (std::is_same<TX, double>::value || !a.bulk_f32)
            wk = __dsub_rn((double)v[q], R->bulk[q]);
        else
            wk = (double)__fsub_rn((float)v[q], R->bulk_f[q]);
        if (HUBBLE)
            wk = __dadd_rn(wk, __ddiv_rn(__dmul_rn(a.hubble, (double)d[q]), a.one_plus_z));
        w[q] = (TVR)wk;
    }
    F.vr = Ar<TVR>::dot3(w[0], w[1], w[2], (TVR)F.rh[0], (TVR)F.rh[1], (TVR)F.rh[2]);
}

// Stage B: compare the bucket's 8 slots in registers and issue the second round
// trip -- the candidate's record, or (no fingerprint matched) the fill counter
// of the home bucket, which tells a newly entered particle from an overflow.
template <typename TF>
OA_D void stage_candidate(const oa_track_args& a, const TrackConst& k, const Policies& pol,
                          const Probe& H, uint32_t& cand, OaRec<TF>& prev, uint32_t& fill) {
    const OaRec<TF>* __restrict__ rec_prev = static_cast<const OaRec<TF>*>(a.rec_prev);
    cand = 0xFFFFFFFFu;
    fill = 0;
    if (H.prev_count == 0) return;
#pragma unroll
    for (int e = OA_BUCKET_SLOTS - 1; e >= 0; --e) {
        const uint32_t xr = H.bk.w[e] ^ H.fpshift;       // == index iff fingerprint equal
        if (xr < H.prev_count) cand = xr;
    }
    // normal eviction priority: HBM delivers 64 B, and the other half is the
    // record of a neighbour that is gathered soon (-25 % HBM reads)
    if (cand != 0xFFFFFFFFu)
        prev = load_rec(rec_prev + ((size_t)H.prev_begin + cand), pol.normal);
    else
        fill = __ldcg(k.cnt_prev + H.prb_idx);
}

// Stage D: exact ID check, apsis test, angle accumulator, outputs.
template <typename TF, typename TVR, bool DIAG>
OA_D void stage_finish(const oa_track_args& a, const TrackConst& k, const Policies& pol,
                       const int64_t id, const uint32_t c, const Probe& H,
                       const Frame<TF, TVR>& F, const uint32_t cand, OaRec<TF>& prev,
                       const uint32_t fill) {
    using AF = Ar<TF>;
    const OaRec<TF>* __restrict__ rec_prev = static_cast<const OaRec<TF>*>(a.rec_prev);
    OaRec<TF>* __restrict__ rec_cur = static_cast<OaRec<TF>*>(a.rec_cur);

    int64_t p = -1;
    if (H.prev_count > 0) {
        bool slow;
        if (cand != 0xFFFFFFFFu) {
            p = (int64_t)H.prev_begin + (int64_t)cand;
            slow = prev.id != id;                     // stale slot / collision
        } else {
            slow = fill > OA_BUCKET_SLOTS;            // overflowed home bucket
        }
        if (slow) {
            const uint32_t home = oa_slot((uint32_t)(oa_mix64((uint64_t)id) >> 32), H.prb_nb);
            p = probe_slow<TF>(k.cnt_prev, k.slot_prev, H.prb_idx - home, H.prb_nb, home,
                               H.fpshift, H.prev_count, rec_prev, H.prev_begin, id,
                               cand == 0xFFFFFFFFu);
            if (p >= 0) prev = load_rec(rec_prev + p, pol.normal);
        }
    }

    __half angle_new = __ushort_as_half((unsigned short)0);
    if (p >= 0) {
        const TF dotp = AF::dot3((TF)prev.rx, (TF)prev.ry, (TF)prev.rz, F.rh[0], F.rh[1],
                                 F.rh[2]);
        const TF dang = AF::acos(dotp);
        bool ev;
        if (a.mode == OA_MODE_PERICENTRIC) ev = (prev.vr < 0) && (F.vr > 0);
        else ev = (prev.vr > 0) && (F.vr < 0);
        if (a.onthefly) {
            if (a.dangle_prev) static_cast<TF*>(a.dangle_prev)[p] = dang;
            a.mark_prev[p] = ev ? (uint16_t)1 : (uint16_t)0;
        } else {
            TF run = AF::add((TF)__half2float(prev.angle), dang);
            if (ev) {
                a.mark_prev[p] = __half_as_ushort(AF::to_half(run));
                run = (TF)0;
            }
            angle_new = AF::to_half(run);
        }
    }

    OaRec<TF> rec;
    rec.id = id;
    rec.rx = F.rh[0]; rec.ry = F.rh[1]; rec.rz = F.rh[2];
    set_vr<TF, TVR>(rec, F.vr);
    rec.r = F.r;
    rec.angle = angle_new;
    rec.flags = 0;
    store_rec(rec_cur + c, rec, pol.first);
    st2(a.mark_cur + c, OA_NO_EVENT, pol.first);

    if (H.ins_k < OA_BUCKET_SLOTS) {
        k.slot_cur[(size_t)H.ins_idx * OA_BUCKET_WORDS + H.ins_k] = H.ins_val;
    } else {           // home bucket full (rare): spill to the next ones
        const uint32_t home = oa_slot((uint32_t)(oa_mix64((uint64_t)id) >> 32), H.ins_nb);
        insert_slow(k.cnt_cur, k.slot_cur, H.ins_idx - home, H.ins_nb, home, H.ins_val);
    }

    if (DIAG) {        // optional per-particle outputs (checkpoint, diagnostics, on-the-fly)
        if (a.out_rhat) {
            TF* o = static_cast<TF*>(a.out_rhat) + 3 * (size_t)c;
            o[0] = F.rh[0]; o[1] = F.rh[1]; o[2] = F.rh[2];
        }
        if (a.out_vr) static_cast<TVR*>(a.out_vr)[c] = F.vr;
        if (a.out_r) static_cast<TF*>(a.out_r)[c] = F.r;
        if (a.out_angle) a.out_angle[c] = __half_as_ushort(angle_new);
        if (a.out_match) a.out_match[c] = p;
    }
}

template <typename TX, typename TF, typename TVR, bool HUBBLE, bool DIAG>
__global__ void __launch_bounds__(TRACK_THREADS, 1)
oa_track_kernel(const __grid_constant__ oa_track_args a, const __grid_constant__ TrackConst k) {
    using L = SlotLayout<TX>;
    constexpr int D = L::DEPTH;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[TRACK_WARPS * D];

    const int lane = threadIdx.x & 31;
    // warp index through a shuffle: the compiler then knows it is warp-uniform
    // and keeps the chunk / slot / TMA address arithmetic on the uniform datapath
    const int warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0);
    const Policies pol = {policy_evict_first(), policy_evict_normal(), policy_evict_last()};
    unsigned char* const wsm = smem + (size_t)warp * D * L::BYTES;
    uint64_t* const full = bars + warp * D;
    const int64_t n = a.n_cur;
    const int stride = gridDim.x * TRACK_WARPS;
    const int first = blockIdx.x * TRACK_WARPS + warp;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < D; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // lane 0: stage chunk `ch` (regions jr.x..jr.y) into slot `s`
    auto issue = [&](int ch, int s, int2 jr) {
        unsigned char* st = wsm + s * L::BYTES;
        const int64_t base = (int64_t)ch * CHUNK;
        const int nrows = jr.y - jr.x + 1;
        const bool tma = k.tma_ok && (base + CHUNK <= n);
        ChunkMeta* m = reinterpret_cast<ChunkMeta*>(st + L::META);
        m->jlo = jr.x;
        m->jhi = jr.y;
        m->nrows = nrows <= RW ? nrows : 0;
        m->tma = tma ? 1 : 0;
        constexpr uint32_t B_IDS = 8u * CHUNK, B_X = (uint32_t)sizeof(TX) * 3u * CHUNK;
        const uint32_t b_rows = nrows <= RW ? (uint32_t)sizeof(Row) * nrows : 0u;
        mbar_arrive_expect_tx(&full[s], (tma ? B_IDS + 2u * B_X : 0u) + b_rows);
        if (tma) {
            tma_load_1d(st + L::IDS, a.ids + base, B_IDS, &full[s], pol.first);
            tma_load_1d(st + L::POS, static_cast<const TX*>(a.pos) + 3 * base, B_X, &full[s],
                        pol.first);
            tma_load_1d(st + L::VEL, static_cast<const TX*>(a.vel) + 3 * base, B_X, &full[s],
                        pol.first);
        }
        if (b_rows)
            tma_load_1d(st + L::ROWS, a.regions + jr.x, b_rows, &full[s], pol.last);
    };

    // stage A of the chunk staged at `st`: wait for its inputs, hash, issue its
    // table traffic.  (A macro-like lambda would not be inlined twice; the
    // per-particle state must stay in registers.)
#define OA_HASH_CHUNK(CH, ST, BAR, PHASE, H_OUT, ID_OUT)                                   \
    do {                                                                                   \
        mbar_wait((BAR), (PHASE));                                                         \
        const ChunkMeta meta_ = *reinterpret_cast<const ChunkMeta*>((ST) + L::META);       \
        const int64_t c_ = (int64_t)(CH) * CHUNK + lane;                                   \
        (ID_OUT) = 0;                                                                      \
        (H_OUT) = empty_probe();                                                           \
        if (c_ < n) {                                                                      \
            (ID_OUT) = meta_.tma ? reinterpret_cast<const int64_t*>((ST) + L::IDS)[lane]   \
                                 : ld8_nc(a.ids + c_, pol.first);                          \
            Row own_;                                                                      \
            const Row* R_ = find_row(a, (ST) + L::ROWS, meta_, c_, &own_);                 \
            (H_OUT) = stage_hash(a, k, pol, R_, (ID_OUT), (uint32_t)c_);                   \
        }                                                                                  \
    } while (0)

    // Chunk schedule.  The first D + 2 chunks of a warp are static (chunk c ->
    // warp c mod stride: the whole GPU sweeps the snapshot as one compact front);
    // every later chunk is taken by TICKET, so that a warp on an SM it shares with
    // another kernel (the NCCL channels of the multi-GPU exchange) simply takes
    // fewer chunks instead of holding the whole launch back.  Position p of the
    // warp's sequence lives in wq[p & 7]; lane 0 requests the ticket of position
    // p + D + 3 during iteration p and stores it one iteration later (the atomic's
    // round trip never stalls the warp).
    __shared__ int wq_all[TRACK_WARPS * 8];
    int* const wq = wq_all + warp * 8;
    if (lane < D + 2) wq[lane] = first + lane * stride;
    __syncwarp();
    const unsigned int static_chunks = (unsigned int)(D + 2) * (unsigned int)stride;
    unsigned int t_pend = 0;
    int my_c = first % TICKETS;               // this warp's ticket counter
    bool drained = false;                     // every counter is past the last chunk
    if (lane == 0) t_pend = atomicAdd(k.ticket + 32 * my_c, 1u);

    // ---- prologue: D chunks in flight, table traffic of the first one issued ------
    if (lane == 0) {
        for (int s = 0; s < D; ++s) {
            const int ch = wq[s];
            if (ch < k.n_chunks) issue(ch, s, __ldg(k.chunk_regions + ch));
        }
    }
    int2 jr_next = make_int2(0, 0);           // regions of the chunk D steps ahead
    if (wq[D] < k.n_chunks) jr_next = __ldg(k.chunk_regions + wq[D]);

    int s = 0;
    uint32_t phase = 0;
    Probe H = empty_probe();
    int64_t id = 0;
    if (first < k.n_chunks) OA_HASH_CHUNK(first, wsm, &full[0], 0u, H, id);

    for (int p = 0;; ++p) {
        const int ch = wq[p & 7];
        if (ch >= k.n_chunks) break;
        unsigned char* st = wsm + s * L::BYTES;
        const int sn = (s + 1 == D) ? 0 : s + 1;
        const uint32_t phase_n = (s + 1 == D) ? (phase ^ 1u) : phase;
        const int64_t c = (int64_t)ch * CHUNK + lane;
        const bool active = c < n;

        // ---- stage A of the NEXT chunk: its bucket load and insert atomic are in
        //      flight during everything below ---------------------------------------
        Probe Hn = empty_probe();
        int64_t id_n = 0;
        const int ch_n = wq[(p + 1) & 7];
        if (ch_n < k.n_chunks)
            OA_HASH_CHUNK(ch_n, wsm + sn * L::BYTES, &full[sn], phase_n, Hn, id_n);

        // ---- stage B: candidate of THIS chunk (its bucket arrived an iteration ago)
        uint32_t cand = 0xFFFFFFFFu, fill = 0;
        OaRec<TF> prev;
        if (active) stage_candidate<TF>(a, k, pol, H, cand, prev, fill);

        // ---- stage C: frame arithmetic, overlapping the record round trip -----------
        Frame<TF, TVR> F;
        F.r = 0; F.vr = 0; F.rh[0] = F.rh[1] = F.rh[2] = 0;
        if (active) {
            const ChunkMeta meta = *reinterpret_cast<const ChunkMeta*>(st + L::META);
            Vec3<TX> x, v;
            if (meta.tma) {
                const TX* sp = reinterpret_cast<const TX*>(st + L::POS) + 3 * lane;
                const TX* sv = reinterpret_cast<const TX*>(st + L::VEL) + 3 * lane;
                x.x = sp[0]; x.y = sp[1]; x.z = sp[2];
                v.x = sv[0]; v.y = sv[1]; v.z = sv[2];
            } else {
                const TX* gp = static_cast<const TX*>(a.pos) + 3 * c;
                const TX* gv = static_cast<const TX*>(a.vel) + 3 * c;
                x.x = ld_elem(gp, pol.first); x.y = ld_elem(gp + 1, pol.first);
                x.z = ld_elem(gp + 2, pol.first);
                v.x = ld_elem(gv, pol.first); v.y = ld_elem(gv + 1, pol.first);
                v.z = ld_elem(gv + 2, pol.first);
            }
            Row own;
            const Row* R = find_row(a, st + L::ROWS, meta, c, &own);
            stage_frame<TX, TF, TVR, HUBBLE>(a, k, R, x, v, F);
        }
        // the slot has been consumed: refill it with the chunk D steps ahead
        __syncwarp();
        const int ahead = wq[(p + D) & 7];
        if (ahead < k.n_chunks) {
            if (lane == 0) issue(ahead, s, jr_next);
            const int ahead2 = wq[(p + D + 1) & 7];
            if (ahead2 < k.n_chunks) jr_next = __ldg(k.chunk_regions + ahead2);
        }
        // the ticket requested an iteration ago is position p + D + 2; request the next
        if (lane == 0) {
            int chunk = k.n_chunks;
            if (!drained) {
                unsigned int t = t_pend;
                for (int tries = 0;; ++tries) {
                    const unsigned long long c =
                        (unsigned long long)static_chunks + (unsigned long long)t * TICKETS + my_c;
                    if (c < (unsigned long long)k.n_chunks) { chunk = (int)c; break; }
                    // this counter is exhausted (the tail of the launch): take
                    // chunks of the next one
                    if (tries == TICKETS) { drained = true; break; }
                    my_c = (my_c + 1 == TICKETS) ? 0 : my_c + 1;
                    t = atomicAdd(k.ticket + 32 * my_c, 1u);
                }
            }
            wq[(p + D + 2) & 7] = chunk;
            if (!drained) t_pend = atomicAdd(k.ticket + 32 * my_c, 1u);
        }

        // ---- stage D ------------------------------------------------------------------
        if (active) stage_finish<TF, TVR, DIAG>(a, k, pol, id, (uint32_t)c, H, F, cand, prev, fill);

        H = Hn;
        id = id_n;
        s = sn;
        phase = phase_n;
    }
#undef OA_HASH_CHUNK
}

template <typename TX, typename TF, typename TVR, bool HUBBLE, bool DIAG>
int launch_track_impl(const oa_track_args& a, cudaStream_t st) {
    auto kern = oa_track_kernel<TX, TF, TVR, HUBBLE, DIAG>;
    using L = SlotLayout<TX>;
    constexpr int smem_bytes = TRACK_WARPS * L::DEPTH * L::BYTES;
    static bool configured = false;     // per instantiation
    if (!configured) {
        OA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           smem_bytes));
        configured = true;
    }
    int dev = 0, sms = OA_NUM_SMS;
    OA_CUDA_CHECK(cudaGetDevice(&dev));
    OA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

    TrackConst k;
    k.n_chunks = (int)((a.n_cur + CHUNK - 1) / CHUNK);
    OA_REQUIRE(a.workspace &&
                   a.workspace_bytes >= sizeof(int2) * ((size_t)k.n_chunks + 1 + TICKETS * 16),
               "oa_track_fused: workspace too small (need oa_track_workspace_bytes)");
    int2* chunks = static_cast<int2*>(a.workspace);
    k.chunk_regions = chunks;
    k.ticket = reinterpret_cast<unsigned int*>(chunks + k.n_chunks);
    k.cnt_prev = a.tab_prev;
    k.slot_prev = a.tab_prev ? a.tab_prev + table_count_words(a.tab_prev_buckets) : nullptr;
    k.cnt_cur = a.tab_cur;
    k.slot_cur = a.tab_cur + table_count_words(a.tab_cur_buckets);
    for (int q = 0; q < 3; ++q) {
        // largest float <= L/2
        const double h = a.box[q] * 0.5;
        float hf = (float)h;
        if ((double)hf > h) hf = nextafterf(hf, -INFINITY);
        k.half_box[q] = hf;
    }
    OA_REQUIRE((reinterpret_cast<uintptr_t>(a.regions) & 15u) == 0,
               "oa_track_fused: the region table must be 16-byte aligned");
    k.tma_ok = ((reinterpret_cast<uintptr_t>(a.ids) | reinterpret_cast<uintptr_t>(a.pos) |
                 reinterpret_cast<uintptr_t>(a.vel)) & 15u) == 0;

    const int tb = 256;
    track_chunks_kernel<<<(unsigned)((k.n_chunks + tb - 1) / tb), tb, 0, st>>>(
        a.cur_off, a.n_regions, a.n_cur, k.n_chunks, chunks);
    OA_LAUNCH_CHECK();
    int64_t grid = sms;                         // persistent: one CTA per SM
    if (a.sm_reserve > 0 && a.sm_reserve < sms) grid = sms - a.sm_reserve;
    const int64_t ctas_needed = (k.n_chunks + TRACK_WARPS - 1) / TRACK_WARPS;
    if (grid > ctas_needed) grid = ctas_needed;
    kern<<<(unsigned)grid, TRACK_THREADS, smem_bytes, st>>>(a, k);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

template <typename TX, typename TF, typename TVR, bool HUBBLE>
int launch_track(const oa_track_args& a, cudaStream_t st) {
    const bool diag = a.out_rhat || a.out_vr || a.out_r || a.out_angle || a.out_match;
    return diag ? launch_track_impl<TX, TF, TVR, HUBBLE, true>(a, st)
                : launch_track_impl<TX, TF, TVR, HUBBLE, false>(a, st);
}

}  // namespace

extern "C" size_t oa_record_bytes(int frame_dtype) {
    return frame_dtype == OA_F64 ? sizeof(OaRec<double>) : sizeof(OaRec<float>);
}

extern "C" int64_t oa_table_bucket_begin(int64_t block_start, int64_t region_index) {
    return bucket_begin(block_start, region_index);
}

extern "C" int64_t oa_table_buckets(int64_t n, int64_t n_regions) {
    return table_buckets(n, n_regions);
}

extern "C" int64_t oa_table_slots(int64_t n, int64_t n_regions) {
    const int64_t b = table_buckets(n, n_regions);
    return table_count_words(b) + b * OA_BUCKET_WORDS;
}

extern "C" int oa_index_bits(int64_t max_block_len) {
    int bits = 1;
    while (bits < 31 && ((int64_t)1 << bits) < max_block_len) ++bits;
    return bits;
}

extern "C" size_t oa_track_workspace_bytes(int64_t n_cur) {
    return sizeof(int2) * (size_t)((n_cur + CHUNK - 1) / CHUNK + 1 + TICKETS * 16);
}

// Only the fill counters are cleared; slot buckets are validated by the counters
// (and, on the fast path, by the ID check of the record).
extern "C" int oa_table_clear(uint32_t* tab, int64_t n, int64_t n_regions, void* stream) {
    OA_REQUIRE(tab && n >= 0, "oa_table_clear: bad arguments");
    OA_CUDA_CHECK(cudaMemsetAsync(
        tab, 0, sizeof(uint32_t) * (size_t)table_count_words(table_buckets(n, n_regions)),
        static_cast<cudaStream_t>(stream)));
    return OA_OK;
}

extern "C" int oa_track_fused(const oa_track_args* args, void* stream) {
    OA_REQUIRE(args != nullptr, "oa_track_fused: args is NULL");
    const oa_track_args& a = *args;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OA_REQUIRE(a.n_cur >= 0 && a.n_regions >= 0, "oa_track_fused: negative size");
    OA_REQUIRE(a.n_cur < ((int64_t)1 << 31) && a.n_prev < ((int64_t)1 << 31),
               "oa_track_fused: more than 2^31-1 region-particles on one GPU "
               "(shard the snapshot across GPUs)");
    OA_REQUIRE(a.mode == OA_MODE_PERICENTRIC || a.mode == OA_MODE_APOCENTRIC,
               "oa_track_fused: bad mode %d", a.mode);
    OA_REQUIRE(a.data_dtype == OA_F32 || a.data_dtype == OA_F64, "bad data_dtype");
    OA_REQUIRE(a.frame_dtype == OA_F32 || a.frame_dtype == OA_F64, "bad frame_dtype");
    OA_REQUIRE(!(a.data_dtype == OA_F64 && a.frame_dtype == OA_F32),
               "oa_track_fused: float64 data cannot have a float32 frame");
    OA_REQUIRE(a.frame_dtype == OA_F64 || a.centre_f32 || a.onthefly,
               "oa_track_fused: a float32 frame needs float32 region centres");
    if (a.n_cur == 0) return OA_OK;
    OA_REQUIRE(a.pos && a.vel && a.ids && a.cur_off && a.regions && a.rec_cur &&
               a.tab_cur && a.mark_cur, "oa_track_fused: NULL required pointer");
    OA_REQUIRE(a.n_regions >= 1, "oa_track_fused: particles without a region");
    OA_REQUIRE(a.n_prev == 0 || !a.rec_prev || (a.tab_prev && a.mark_prev),
               "oa_track_fused: previous generation incomplete");
    OA_REQUIRE(a.cur_index_bits >= 1 && a.cur_index_bits <= 31, "bad cur_index_bits");
    OA_REQUIRE(!a.rec_prev || (a.prev_index_bits >= 1 && a.prev_index_bits <= 31),
               "bad prev_index_bits");
    OA_REQUIRE(a.tab_cur_buckets > 0 && (!a.tab_prev || a.tab_prev_buckets > 0),
               "oa_track_fused: table bucket counts missing (oa_table_buckets)");

    const bool hub = (a.hubble != 0.0) && !a.onthefly;
    const bool x64 = a.data_dtype == OA_F64, f64 = a.frame_dtype == OA_F64;
    if (a.onthefly) {
        // on-the-fly: frame and v_r in the data dtype, no Hubble term
        OA_REQUIRE(x64 == f64, "oa_track_fused: on-the-fly frame dtype must equal data dtype");
        if (x64) return launch_track<double, double, double, false>(a, st);
        return launch_track<float, float, float, false>(a, st);
    }
    if (x64) {
        return hub ? launch_track<double, double, double, true>(a, st)
                   : launch_track<double, double, double, false>(a, st);
    }
    if (f64) {
        return hub ? launch_track<float, double, double, true>(a, st)
                   : launch_track<float, double, double, false>(a, st);
    }
    return hub ? launch_track<float, float, double, true>(a, st)
               : launch_track<float, float, double, false>(a, st);
}
