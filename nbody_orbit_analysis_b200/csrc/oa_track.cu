// Fused per-snapshot orbit-tracking kernel (sm_100a).
//
// One pass over the current snapshot's region blocks does the whole of the
// reference's per-halo `track(j)` (track_orbits.py:147-185): halo frame,
// ID match against the previous block, apsis detection, float16 angle update.
// See include/orbit_b200.h (oa_track_fused) for the contract and DESIGN.md for
// the memory layout and the roofline accounting.
//
// Matching is a region-segmented open-addressing hash table in global memory
// that lives in B200's 126 MB L2 while a block is being processed: segment of
// region q = slots [2*begin_q, 2*(begin_q+len_q)), load factor 1/2, linear
// probing, slot = fingerprint | block-local index.  A probe hit is verified
// against the 64-bit ID stored in the previous record, which is the same 32 B
// sector that carries rhat / v_r / angle, so match + state read is one gather.
//
// Arithmetic mirrors numpy's evaluation order and rounding points (no FMA
// contraction: this file is compiled with -fmad=false and uses *_rn intrinsics
// where the order matters) -- see SURVEY.md 2.2 / 7.4-7.6.
#include "oa_common.cuh"
#include <type_traits>

namespace {

constexpr int TRACK_THREADS = 256;

template <typename T> struct Ar;   // IEEE ops without contraction
template <> struct Ar<float> {
    static OA_D float add(float a, float b) { return __fadd_rn(a, b); }
    static OA_D float sub(float a, float b) { return __fsub_rn(a, b); }
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
    static OA_D double sub(double a, double b) { return __dsub_rn(a, b); }
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

// 16-byte vector copies of a record
template <typename TF>
OA_D OaRec<TF> load_rec(const OaRec<TF>* p) {
    OaRec<TF> r;
    const int4* src = reinterpret_cast<const int4*>(p);
    int4* dst = reinterpret_cast<int4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(OaRec<TF>) / 16); ++i) dst[i] = __ldg(src + i);
    return r;
}
template <typename TF>
OA_D void store_rec(OaRec<TF>* p, const OaRec<TF>& r) {
    int4* dst = reinterpret_cast<int4*>(p);
    const int4* src = reinterpret_cast<const int4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(OaRec<TF>) / 16); ++i) dst[i] = src[i];
}

struct RegionRow {
    double c[3];
    double b[3];
    int64_t prev_begin;
    int64_t prev_count;
};

OA_D RegionRow load_region(const oa_region* regions, int j) {
    RegionRow R;
    const double2* src = reinterpret_cast<const double2*>(regions + j);
    double2 a0 = __ldg(src + 0), a1 = __ldg(src + 1), a2 = __ldg(src + 2);
    const longlong2* tail = reinterpret_cast<const longlong2*>(regions + j) + 3;
    longlong2 t = __ldg(tail);
    R.c[0] = a0.x; R.c[1] = a0.y; R.c[2] = a1.x;
    R.b[0] = a1.y; R.b[1] = a2.x; R.b[2] = a2.y;
    R.prev_begin = t.x; R.prev_count = t.y;
    return R;
}

// last j in [lo, hi] with off[j] <= c   (off is non-decreasing; empty blocks
// share their start with the next block and are skipped by taking the last)
OA_D int find_region(const int64_t* __restrict__ off, int lo, int hi, int64_t c) {
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(off + mid) <= c) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <typename TX, typename TF, typename TVR, bool HUBBLE, int ITEMS>
__global__ void __launch_bounds__(TRACK_THREADS)
oa_track_kernel(const oa_track_args a) {
    using AF = Ar<TF>;
    constexpr int TILE = TRACK_THREADS * ITEMS;
    __shared__ int s_j[2];

    const TX* __restrict__ pos = static_cast<const TX*>(a.pos);
    const TX* __restrict__ vel = static_cast<const TX*>(a.vel);
    const int64_t* __restrict__ ids = a.ids;
    const int64_t* __restrict__ off = a.cur_off;
    const OaRec<TF>* __restrict__ rec_prev = static_cast<const OaRec<TF>*>(a.rec_prev);
    OaRec<TF>* __restrict__ rec_cur = static_cast<OaRec<TF>*>(a.rec_cur);
    const uint32_t* __restrict__ tab_prev = a.tab_prev;
    const bool have_prev = (a.rec_prev != nullptr) && (a.n_prev > 0);
    const uint32_t pmask = (a.prev_index_bits >= 32) ? 0xFFFFFFFFu
                                                     : ((1u << a.prev_index_bits) - 1u);
    const int pbits = a.prev_index_bits, cbits = a.cur_index_bits;
    const int64_t n = a.n_cur;
    const int64_t n_tiles = (n + TILE - 1) / TILE;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * TILE;
        const int64_t last = min(base + (int64_t)TILE, n) - 1;
        if (threadIdx.x == 0) s_j[0] = find_region(off, 0, a.n_regions - 1, base);
        if (threadIdx.x == 32) s_j[1] = find_region(off, 0, a.n_regions - 1, last);
        __syncthreads();
        const int jlo = s_j[0], jhi = s_j[1];

#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const int64_t c = base + (int64_t)it * TRACK_THREADS + threadIdx.x;
            if (c >= n) continue;

            // ---- inputs ----------------------------------------------------------
            const int64_t id = __ldg(ids + c);
            TX x[3], v[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                x[k] = __ldg(pos + 3 * c + k);
                v[k] = __ldg(vel + 3 * c + k);
            }
            const int j = find_region(off, jlo, jhi, c);
            const RegionRow R = load_region(a.regions, j);
            const int64_t cur_begin = __ldg(off + j);
            const int64_t cur_len = __ldg(off + j + 1) - cur_begin;

            // ---- halo frame (region_frame) -------------------------------------
            TF d[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (std::is_same<TX, double>::value || !a.centre_f32) {
                    double dd = __dsub_rn((double)x[k], R.c[k]);
                    if (a.periodic) {
                        const double L = a.box[k], h = L * 0.5;
                        if (dd > h) dd = __dsub_rn(dd, L);
                        if (dd < -h) dd = __dadd_rn(dd, L);
                    }
                    d[k] = (TF)dd;
                } else {
                    float df = __fsub_rn((float)x[k], (float)R.c[k]);
                    if (a.periodic) {
                        const double L = a.box[k], h = L * 0.5;
                        if ((double)df > h) df = (float)__dsub_rn((double)df, L);
                        if ((double)df < -h) df = (float)__dadd_rn((double)df, L);
                    }
                    d[k] = (TF)df;
                }
            }
            const TF r = AF::sqrt(AF::dot3(d[0], d[1], d[2], d[0], d[1], d[2]));
            TF rh[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) rh[k] = AF::div(d[k], r);

            TVR w[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double wk;
                if (std::is_same<TX, double>::value || !a.bulk_f32)
                    wk = __dsub_rn((double)v[k], R.b[k]);
                else
                    wk = (double)__fsub_rn((float)v[k], (float)R.b[k]);
                if (HUBBLE)
                    wk = __dadd_rn(wk, __ddiv_rn(__dmul_rn(a.hubble, (double)d[k]),
                                                 a.one_plus_z));
                w[k] = (TVR)wk;
            }
            const TVR vr = Ar<TVR>::dot3(w[0], w[1], w[2], (TVR)rh[0], (TVR)rh[1],
                                         (TVR)rh[2]);

            // ---- match against the halo's previous block ----------------------
            const uint64_t hsh = oa_mix64((uint64_t)id);
            const uint32_t h_slot = (uint32_t)(hsh >> 32);
            const uint32_t h_fp = (uint32_t)hsh;
            int64_t p = -1;
            OaRec<TF> prev;
            if (have_prev && R.prev_count > 0) {
                const uint32_t cap = (uint32_t)(2 * R.prev_count);
                const uint32_t* seg = tab_prev + 2 * R.prev_begin;
                const uint32_t fp = (pbits >= 32) ? 0u : (h_fp >> pbits);
                uint32_t s = oa_slot(h_slot, cap);
                for (uint32_t probes = 0; probes < cap; ++probes) {
                    const uint32_t val = __ldg(seg + s);
                    if (val == OA_EMPTY) break;
                    if (pbits >= 32 || (val >> pbits) == fp) {
                        const int64_t q = R.prev_begin + (int64_t)(val & pmask);
                        prev = load_rec(rec_prev + q);
                        if (prev.id == id) { p = q; break; }
                    }
                    s = (s + 1 == cap) ? 0u : s + 1;
                }
            }

            // ---- apsis test + angle accumulator (compare_radial_velocities,
            //      calc_angles) ---------------------------------------------------
            __half angle_new = __ushort_as_half((unsigned short)0);
            if (p >= 0) {
                const TF dotp = AF::dot3((TF)prev.rx, (TF)prev.ry, (TF)prev.rz,
                                         rh[0], rh[1], rh[2]);
                const TF dang = AF::acos(dotp);
                bool ev;
                if (a.mode == OA_MODE_PERICENTRIC) ev = (prev.vr < 0) && (vr > 0);
                else ev = (prev.vr > 0) && (vr < 0);
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

            // ---- new state ----------------------------------------------------------
            OaRec<TF> rec;
            rec.id = id;
            rec.rx = rh[0]; rec.ry = rh[1]; rec.rz = rh[2];
            set_vr<TF, TVR>(rec, vr);
            rec.r = r;
            rec.angle = angle_new;
            rec.flags = 0;
            store_rec(rec_cur + c, rec);
            a.mark_cur[c] = OA_NO_EVENT;

            {   // insert into the current table (probed by the next snapshot)
                const uint32_t cap = (uint32_t)(2 * cur_len);
                uint32_t* seg = a.tab_cur + 2 * cur_begin;
                const uint32_t local = (uint32_t)(c - cur_begin);
                const uint32_t val = (cbits >= 32) ? local
                                                   : (((h_fp >> cbits) << cbits) | local);
                uint32_t s = oa_slot(h_slot, cap);
                for (uint32_t probes = 0; probes < cap; ++probes) {
                    const uint32_t old = atomicCAS(seg + s, OA_EMPTY, val);
                    if (old == OA_EMPTY) break;
                    s = (s + 1 == cap) ? 0u : s + 1;
                }
            }

            // ---- optional per-particle outputs ------------------------------------
            if (a.out_rhat) {
                TF* o = static_cast<TF*>(a.out_rhat) + 3 * c;
                o[0] = rh[0]; o[1] = rh[1]; o[2] = rh[2];
            }
            if (a.out_vr) static_cast<TVR*>(a.out_vr)[c] = vr;
            if (a.out_r) static_cast<TF*>(a.out_r)[c] = r;
            if (a.out_angle) a.out_angle[c] = __half_as_ushort(angle_new);
            if (a.out_match) a.out_match[c] = p;
        }
        __syncthreads();
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
    const int64_t tiles = (a.n_cur + TRACK_THREADS * ITEMS - 1) / (TRACK_THREADS * ITEMS);
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, TRACK_THREADS, 0, st>>>(a);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

}  // namespace

extern "C" size_t oa_record_bytes(int frame_dtype) {
    return frame_dtype == OA_F64 ? sizeof(OaRec<double>) : sizeof(OaRec<float>);
}

extern "C" int64_t oa_table_slots(int64_t n) { return 2 * n + 2; }

extern "C" int oa_index_bits(int64_t max_block_len) {
    // the all-ones pattern is reserved for OA_EMPTY: need 2^bits - 1 > max index
    int bits = 1;
    while (bits < 32 && ((int64_t)1 << bits) - 1 <= max_block_len) ++bits;
    return bits;
}

extern "C" int oa_table_clear(uint32_t* tab, int64_t n, void* stream) {
    OA_REQUIRE(tab && n >= 0, "oa_table_clear: bad arguments");
    OA_CUDA_CHECK(cudaMemsetAsync(tab, 0xFF, sizeof(uint32_t) * (size_t)oa_table_slots(n),
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
    OA_REQUIRE(a.cur_index_bits >= 1 && a.cur_index_bits <= 32, "bad cur_index_bits");

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
