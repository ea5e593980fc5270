// Fused per-snapshot orbit-tracking kernel (sm_100a).
//
// One pass over the current snapshot's region blocks does the whole of the
// reference's per-halo `track(j)` (track_orbits.py:147-185): halo frame,
// ID match against the previous block, apsis detection, float16 angle update.
// See include/orbit_b200.h (oa_track_fused) for the contract and DESIGN.md for
// the memory layout and the roofline accounting.
//
// Matching: a region-segmented, bucketised hash table in global memory that
// stays in B200's 126 MB L2 while a block is being processed.  A bucket is one
// 32-byte sector = {count, 7 slots}; slot = fingerprint | block-local index.
//   insert : k = atomicAdd(bucket.count, 1); bucket.slot[k] = value   (1 atomic)
//   probe  : one sector read, fingerprint compare in registers, then the 32 B
//            record of the candidate -- which carries the 64-bit ID for the
//            exact check AND rhat / v_r / angle, so match + state read is one
//            gather.  Overflow (count > 7, <1 % of keys) spills to the next
//            bucket of the same region.
// Everything that is touched once (ids / positions / velocities / records /
// marks) is loaded and stored with an L2 evict-first policy so that only the
// tables compete for L2.
//
// Latency: no block-level barriers; each warp owns chunks of 32*ITEMS
// consecutive particles and runs the dependent chain
//   inputs -> bucket -> record -> outputs
// with all ITEMS loads of a stage in flight together.
//
// Arithmetic mirrors numpy's evaluation order and rounding points (no FMA
// contraction: compiled with -fmad=false and *_rn intrinsics where the order
// matters) -- see SURVEY.md 2.2 / 7.4-7.6.
#include "oa_common.cuh"
#include <type_traits>

namespace {

constexpr int TRACK_THREADS = 256;
constexpr int TRACK_WARPS = TRACK_THREADS / 32;

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
OA_D uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
OA_D int4 ld16_nc(const void* a, uint64_t pol) {     // read-only, no L1 allocation
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a), "l"(pol));
    return v;
}
OA_D int64_t ld8_nc(const int64_t* a, uint64_t pol) {
    int64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s64 %0, [%1], %2;"
                 : "=l"(v) : "l"(a), "l"(pol));
    return v;
}
// positions / velocities: (n,3) rows read with three strided scalar loads per
// thread -- keep L1 allocation (the three loads of a warp share their lines)
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
OA_D void st16(void* a, const int4& v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
OA_D void st2(uint16_t* a, uint16_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u16 [%0], %1, %2;" :: "l"(a), "h"(v), "l"(pol) : "memory");
}
OA_D void st4(uint32_t* a, uint32_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(a), "r"(v), "l"(pol) : "memory");
}
OA_D uint32_t atom_add(uint32_t* a, uint32_t v, uint64_t pol) {
    uint32_t o;
    asm volatile("atom.global.add.L2::cache_hint.u32 %0, [%1], %2, %3;"
                 : "=r"(o) : "l"(a), "r"(v), "l"(pol) : "memory");
    return o;
}

template <typename TF>
OA_D OaRec<TF> load_rec(const OaRec<TF>* p, uint64_t pol) {
    OaRec<TF> r;
    int4* dst = reinterpret_cast<int4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(OaRec<TF>) / 16); ++i)
        dst[i] = ld16_nc(reinterpret_cast<const int4*>(p) + i, pol);
    return r;
}
template <typename TF>
OA_D void store_rec(OaRec<TF>* p, const OaRec<TF>& r, uint64_t pol) {
    const int4* src = reinterpret_cast<const int4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(OaRec<TF>) / 16); ++i)
        st16(reinterpret_cast<int4*>(p) + i, src[i], pol);
}

struct Bucket {
    uint32_t w[OA_BUCKET_WORDS];   // w[0] = count, w[1..7] = slots
};
OA_D Bucket load_bucket(const uint32_t* tab, int64_t bucket, uint64_t pol) {
    Bucket b;
    const int4* src = reinterpret_cast<const int4*>(tab + bucket * OA_BUCKET_WORDS);
    const int4 lo = ld16_nc(src, pol), hi = ld16_nc(src + 1, pol);
    b.w[0] = lo.x; b.w[1] = lo.y; b.w[2] = lo.z; b.w[3] = lo.w;
    b.w[4] = hi.x; b.w[5] = hi.y; b.w[6] = hi.z; b.w[7] = hi.w;
    return b;
}

struct RegionRow {
    double c[3];
    double b[3];
    int64_t prev_begin, prev_count, prev_bucket, cur_bucket;
};
OA_D RegionRow load_region(const oa_region* regions, int j) {
    RegionRow R;
    const double2* src = reinterpret_cast<const double2*>(regions + j);
    const double2 a0 = __ldg(src + 0), a1 = __ldg(src + 1), a2 = __ldg(src + 2);
    const longlong2* tail = reinterpret_cast<const longlong2*>(regions + j) + 3;
    const longlong2 t0 = __ldg(tail), t1 = __ldg(tail + 1);
    R.c[0] = a0.x; R.c[1] = a0.y; R.c[2] = a1.x;
    R.b[0] = a1.y; R.b[1] = a2.x; R.b[2] = a2.y;
    R.prev_begin = t0.x; R.prev_count = t0.y;
    R.prev_bucket = t1.x; R.cur_bucket = t1.y;
    return R;
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

OA_D uint32_t bucket_count(int64_t len) { return (uint32_t)((2 * len) / 7 + 1); }
OA_HD int64_t bucket_begin(int64_t block_start, int64_t region) {
    return (2 * block_start) / 7 + region;
}

// ---- rare paths, kept out of line so they do not cost registers -----------------------
template <typename TF>
__device__ __noinline__ int64_t probe_slow(const uint32_t* __restrict__ tab,
                                           int64_t bucket_begin, uint32_t nb,
                                           uint32_t home, uint32_t fp, int pbits,
                                           uint32_t pmask,
                                           const OaRec<TF>* __restrict__ rec_prev,
                                           int64_t prev_begin, int64_t id) {
    const uint64_t pol = policy_evict_last(), pol_s = policy_evict_first();
    uint32_t b = home;
    for (uint32_t t = 0; t < nb; ++t) {
        const Bucket B = load_bucket(tab, bucket_begin + b, pol);
        const uint32_t cnt = B.w[0];
        const uint32_t m = cnt < OA_BUCKET_SLOTS ? cnt : OA_BUCKET_SLOTS;
        for (uint32_t e = 1; e <= m; ++e) {
            if ((B.w[e] >> pbits) == fp) {
                const int64_t q = prev_begin + (int64_t)(B.w[e] & pmask);
                const OaRec<TF> r = load_rec(rec_prev + q, pol_s);
                if (r.id == id) return q;
            }
        }
        if (cnt <= OA_BUCKET_SLOTS) return -1;      // never overflowed: a miss
        b = (b + 1 == nb) ? 0u : b + 1;
    }
    return -1;
}

__device__ __noinline__ void insert_slow(uint32_t* __restrict__ tab, int64_t bucket_begin,
                                         uint32_t nb, uint32_t home, uint32_t val) {
    uint32_t b = home;
    for (uint32_t t = 1; t < nb; ++t) {
        b = (b + 1 == nb) ? 0u : b + 1;
        uint32_t* bk = tab + (bucket_begin + b) * OA_BUCKET_WORDS;
        const uint32_t k = atomicAdd(bk, 1u);
        if (k < OA_BUCKET_SLOTS) { bk[1 + k] = val; return; }
    }
}

template <typename TX, typename TF, typename TVR, bool HUBBLE, int ITEMS>
__global__ void __launch_bounds__(TRACK_THREADS)
oa_track_kernel(const oa_track_args a) {
    using AF = Ar<TF>;
    constexpr int CHUNK = 32 * ITEMS;

    const TX* __restrict__ pos = static_cast<const TX*>(a.pos);
    const TX* __restrict__ vel = static_cast<const TX*>(a.vel);
    const int64_t* __restrict__ off = a.cur_off;
    const OaRec<TF>* __restrict__ rec_prev = static_cast<const OaRec<TF>*>(a.rec_prev);
    OaRec<TF>* __restrict__ rec_cur = static_cast<OaRec<TF>*>(a.rec_cur);
    const bool have_prev = (a.rec_prev != nullptr) && (a.n_prev > 0);
    const int pbits = a.prev_index_bits, cbits = a.cur_index_bits;
    const uint32_t pmask = (1u << pbits) - 1u;
    const int64_t n = a.n_cur;
    const int64_t n_chunks = (n + CHUNK - 1) / CHUNK;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * TRACK_WARPS + (threadIdx.x >> 5);
    const int64_t warps = (int64_t)gridDim.x * TRACK_WARPS;
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_keep = policy_evict_last();

    for (int64_t chunk = warp0; chunk < n_chunks; chunk += warps) {
        const int64_t base = chunk * CHUNK;
        const int64_t last = min(base + (int64_t)CHUNK, n) - 1;

        // ---- stage 1: inputs (all ITEMS in flight) ---------------------------------
        int64_t id[ITEMS];
        TX x[ITEMS][3], v[ITEMS][3];
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const int64_t c = base + it * 32 + lane;
            if (c < n) {
                id[it] = ld8_nc(a.ids + c, pol_stream);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    x[it][k] = ld_elem(pos + 3 * c + k, pol_stream);
                    v[it][k] = ld_elem(vel + 3 * c + k, pol_stream);
                }
            }
        }
        // region range of this chunk (warp-uniform loads, overlap with stage 1)
        const int jlo = find_region(off, 0, a.n_regions - 1, base);
        const int jhi = find_region(off, jlo, a.n_regions - 1, last);

        // ---- stage 2: frame, hash, table traffic issued ---------------------------
        TF rh[ITEMS][3], r[ITEMS];
        TVR vr[ITEMS];
        Bucket bk[ITEMS];
        uint32_t ins_k[ITEMS], ins_val[ITEMS], fp[ITEMS];
        uint32_t* ins_ptr[ITEMS];
        int64_t prev_begin[ITEMS];
        int jreg[ITEMS];
        bool probe[ITEMS];
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const int64_t c = base + it * 32 + lane;
            probe[it] = false;
            if (c >= n) continue;
            const int j = find_region(off, jlo, jhi, c);
            jreg[it] = j;
            const RegionRow R = load_region(a.regions, j);
            const int64_t cur_begin = __ldg(off + j);
            const int64_t cur_len = __ldg(off + j + 1) - cur_begin;

            // halo frame (region_frame)
            TF d[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (std::is_same<TX, double>::value || !a.centre_f32) {
                    double dd = __dsub_rn((double)x[it][k], R.c[k]);
                    if (a.periodic) {
                        const double L = a.box[k], h = L * 0.5;
                        if (dd > h) dd = __dsub_rn(dd, L);
                        if (dd < -h) dd = __dadd_rn(dd, L);
                    }
                    d[k] = (TF)dd;
                } else {
                    float df = __fsub_rn((float)x[it][k], (float)R.c[k]);
                    if (a.periodic) {
                        const double L = a.box[k], h = L * 0.5;
                        if ((double)df > h) df = (float)__dsub_rn((double)df, L);
                        if ((double)df < -h) df = (float)__dadd_rn((double)df, L);
                    }
                    d[k] = (TF)df;
                }
            }
            r[it] = AF::sqrt(AF::dot3(d[0], d[1], d[2], d[0], d[1], d[2]));
#pragma unroll
            for (int k = 0; k < 3; ++k) rh[it][k] = AF::div(d[k], r[it]);
            TVR w[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double wk;
                if (std::is_same<TX, double>::value || !a.bulk_f32)
                    wk = __dsub_rn((double)v[it][k], R.b[k]);
                else
                    wk = (double)__fsub_rn((float)v[it][k], (float)R.b[k]);
                if (HUBBLE)
                    wk = __dadd_rn(wk, __ddiv_rn(__dmul_rn(a.hubble, (double)d[k]),
                                                 a.one_plus_z));
                w[k] = (TVR)wk;
            }
            vr[it] = Ar<TVR>::dot3(w[0], w[1], w[2], (TVR)rh[it][0], (TVR)rh[it][1],
                                   (TVR)rh[it][2]);

            const uint64_t hsh = oa_mix64((uint64_t)id[it]);
            const uint32_t h_slot = (uint32_t)(hsh >> 32), h_fp = (uint32_t)hsh;

            // insert into the current table: the atomic is issued now, its result
            // is consumed at the very end
            ins_ptr[it] = a.tab_cur + (R.cur_bucket + oa_slot(h_slot, bucket_count(cur_len))) *
                                          OA_BUCKET_WORDS;
            ins_val[it] = ((h_fp >> cbits) << cbits) | (uint32_t)(c - cur_begin);
            ins_k[it] = atom_add(ins_ptr[it], 1u, pol_keep);

            // probe of the previous table: one sector
            if (have_prev && R.prev_count > 0) {
                probe[it] = true;
                prev_begin[it] = R.prev_begin;
                fp[it] = h_fp >> pbits;
                bk[it] = load_bucket(
                    a.tab_prev,
                    R.prev_bucket + oa_slot(h_slot, bucket_count(R.prev_count)), pol_keep);
            }
        }

        // ---- stage 3: candidate record loads ------------------------------------------
        int64_t p[ITEMS];
        OaRec<TF> prev[ITEMS];
        bool slow[ITEMS];
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            p[it] = -1;
            slow[it] = false;
            if (!probe[it]) continue;
            const uint32_t cnt = bk[it].w[0];
            const uint32_t m = cnt < OA_BUCKET_SLOTS ? cnt : OA_BUCKET_SLOTS;
            int64_t cand = -1;
#pragma unroll
            for (int e = OA_BUCKET_SLOTS; e >= 1; --e)
                if ((uint32_t)e <= m && (bk[it].w[e] >> pbits) == fp[it])
                    cand = prev_begin[it] + (int64_t)(bk[it].w[e] & pmask);
            if (cand >= 0) {
                prev[it] = load_rec(rec_prev + cand, pol_stream);
                p[it] = cand;
            } else if (cnt > OA_BUCKET_SLOTS) {
                slow[it] = true;             // bucket overflowed: look further
            }
        }

        // ---- stage 4: verify, apsis test, angle accumulator, outputs ------------------
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const int64_t c = base + it * 32 + lane;
            if (c >= n) continue;
            if (p[it] >= 0 && prev[it].id != id[it]) { p[it] = -1; slow[it] = true; }
            if (slow[it]) {        // fingerprint collision or overflowed bucket (rare)
                const RegionRow R = load_region(a.regions, jreg[it]);
                const uint32_t nb = bucket_count(R.prev_count);
                const uint32_t home =
                    oa_slot((uint32_t)(oa_mix64((uint64_t)id[it]) >> 32), nb);
                p[it] = probe_slow<TF>(a.tab_prev, R.prev_bucket, nb, home, fp[it], pbits,
                                       pmask, rec_prev, R.prev_begin, id[it]);
                if (p[it] >= 0) prev[it] = load_rec(rec_prev + p[it], pol_stream);
            }

            __half angle_new = __ushort_as_half((unsigned short)0);
            if (p[it] >= 0) {
                const TF dotp = AF::dot3((TF)prev[it].rx, (TF)prev[it].ry, (TF)prev[it].rz,
                                         rh[it][0], rh[it][1], rh[it][2]);
                const TF dang = AF::acos(dotp);
                bool ev;
                if (a.mode == OA_MODE_PERICENTRIC) ev = (prev[it].vr < 0) && (vr[it] > 0);
                else ev = (prev[it].vr > 0) && (vr[it] < 0);
                if (a.onthefly) {
                    if (a.dangle_prev) static_cast<TF*>(a.dangle_prev)[p[it]] = dang;
                    a.mark_prev[p[it]] = ev ? (uint16_t)1 : (uint16_t)0;
                } else {
                    TF run = AF::add((TF)__half2float(prev[it].angle), dang);
                    if (ev) {
                        a.mark_prev[p[it]] = __half_as_ushort(AF::to_half(run));
                        run = (TF)0;
                    }
                    angle_new = AF::to_half(run);
                }
            }

            OaRec<TF> rec;
            rec.id = id[it];
            rec.rx = rh[it][0]; rec.ry = rh[it][1]; rec.rz = rh[it][2];
            set_vr<TF, TVR>(rec, vr[it]);
            rec.r = r[it];
            rec.angle = angle_new;
            rec.flags = 0;
            store_rec(rec_cur + c, rec, pol_stream);
            st2(a.mark_cur + c, OA_NO_EVENT, pol_stream);

            if (ins_k[it] < OA_BUCKET_SLOTS) {
                st4(ins_ptr[it] + 1 + ins_k[it], ins_val[it], pol_keep);
            } else {               // home bucket full (rare): spill to the next ones
                const int64_t cb = __ldg(off + jreg[it]);
                const uint32_t nb = bucket_count(__ldg(off + jreg[it] + 1) - cb);
                const int64_t first = bucket_begin(cb, jreg[it]);
                const uint32_t home =
                    (uint32_t)((ins_ptr[it] - a.tab_cur) / OA_BUCKET_WORDS - first);
                insert_slow(a.tab_cur, first, nb, home, ins_val[it]);
            }

            if (a.out_rhat) {
                TF* o = static_cast<TF*>(a.out_rhat) + 3 * c;
                o[0] = rh[it][0]; o[1] = rh[it][1]; o[2] = rh[it][2];
            }
            if (a.out_vr) static_cast<TVR*>(a.out_vr)[c] = vr[it];
            if (a.out_r) static_cast<TF*>(a.out_r)[c] = r[it];
            if (a.out_angle) a.out_angle[c] = __half_as_ushort(angle_new);
            if (a.out_match) a.out_match[c] = p[it];
        }
    }
}

template <typename TX, typename TF, typename TVR, bool HUBBLE>
int launch_track(const oa_track_args& a, cudaStream_t st) {
    constexpr int ITEMS = 2;
    auto kern = oa_track_kernel<TX, TF, TVR, HUBBLE, ITEMS>;
    int per_sm = 0;
    OA_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &per_sm, kern, TRACK_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    int dev = 0, sms = OA_NUM_SMS;
    OA_CUDA_CHECK(cudaGetDevice(&dev));
    OA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t per_block = (int64_t)TRACK_WARPS * 32 * ITEMS;
    const int64_t blocks_needed = (a.n_cur + per_block - 1) / per_block;
    int64_t grid = (int64_t)sms * per_sm;       // persistent: a multiple of 148
    if (grid > blocks_needed) grid = blocks_needed;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, TRACK_THREADS, 0, st>>>(a);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

}  // namespace

extern "C" size_t oa_record_bytes(int frame_dtype) {
    return frame_dtype == OA_F64 ? sizeof(OaRec<double>) : sizeof(OaRec<float>);
}

// Bucket b of region j (block start `off`, length `len`) lives at bucket index
// (2*off)/7 + j + b,  b < (2*len)/7 + 1:  closed form, no prefix sum needed.
extern "C" int64_t oa_table_bucket_begin(int64_t block_start, int64_t region_index) {
    return bucket_begin(block_start, region_index);
}

extern "C" int64_t oa_table_slots(int64_t n, int64_t n_regions) {
    return ((2 * n) / 7 + n_regions + 2) * OA_BUCKET_WORDS;
}

extern "C" int oa_index_bits(int64_t max_block_len) {
    int bits = 1;
    while (bits < 31 && ((int64_t)1 << bits) < max_block_len) ++bits;
    return bits;
}

extern "C" int oa_table_clear(uint32_t* tab, int64_t n, int64_t n_regions, void* stream) {
    OA_REQUIRE(tab && n >= 0, "oa_table_clear: bad arguments");
    OA_CUDA_CHECK(cudaMemsetAsync(tab, 0,
                                  sizeof(uint32_t) * (size_t)oa_table_slots(n, n_regions),
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
    if (a.n_cur == 0) return OA_OK;
    OA_REQUIRE(a.pos && a.vel && a.ids && a.cur_off && a.regions && a.rec_cur &&
               a.tab_cur && a.mark_cur, "oa_track_fused: NULL required pointer");
    OA_REQUIRE(a.n_prev == 0 || !a.rec_prev || (a.tab_prev && a.mark_prev),
               "oa_track_fused: previous generation incomplete");
    OA_REQUIRE(a.cur_index_bits >= 1 && a.cur_index_bits <= 31, "bad cur_index_bits");
    OA_REQUIRE(!a.rec_prev || (a.prev_index_bits >= 1 && a.prev_index_bits <= 31),
               "bad prev_index_bits");

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
