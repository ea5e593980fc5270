// Sorted-list joins and segmented reductions for the two consumers of the
// tracking path (sm_100a):
//   * progenitors.py  (get_central_particle_ids :5-56, find_main_progenitors :59-117)
//   * postprocessing.py (Apsides.collate_apsides :30-174, save_final_apsis_counts :176-240)
// Both are integer joins on particle IDs over lists that are sorted with
// oa_sort_pairs_u64 first; everything here is one thread per element with a
// binary search into an L2-resident (or at worst HBM-streamed) sorted array.
#include "oa_common.cuh"

namespace {

inline unsigned blocks_for(int64_t n, int threads) {
    return (unsigned)((n + threads - 1) / threads);
}

// last j in [0, n_seg) with off[j] <= i   (off non-decreasing)
OA_D int find_segment(const int64_t* __restrict__ off, int n_seg, int64_t i) {
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(off + mid) <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ---- progenitors.py:41-51: radius of every particle about its region centre ----------
// numpy: tmp = coords[sl] - pos (dtype = result_type(coords, pos)); wrap in that
// dtype with float64 box; stored into a float64 array; r = sqrt(einsum) in
// float64 with the (p0 + p2) + p1 order.
template <typename TX>
__global__ void central_radii_kernel(const TX* __restrict__ pos,
                                     const int64_t* __restrict__ off, int n_regions,
                                     const double* __restrict__ centres, int centre_f32,
                                     int periodic, double bx, double by, double bz,
                                     int64_t n, double* __restrict__ r_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = find_segment(off, n_regions, i);
    const double box[3] = {bx, by, bz};
    double d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double c = __ldg(centres + 3 * j + k);
        if (sizeof(TX) == 8 || !centre_f32) {
            double dd = __dsub_rn((double)pos[3 * i + k], c);
            if (periodic) {
                const double h = box[k] * 0.5;
                if (dd > h) dd = __dsub_rn(dd, box[k]);
                if (dd < -h) dd = __dadd_rn(dd, box[k]);
            }
            d[k] = dd;
        } else {
            float df = __fsub_rn((float)pos[3 * i + k], (float)c);
            if (periodic) {
                const double h = box[k] * 0.5;
                if ((double)df > h) df = (float)__dsub_rn((double)df, box[k]);
                if ((double)df < -h) df = (float)__dadd_rn((double)df, box[k]);
            }
            d[k] = (double)df;
        }
    }
    const double s = __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[2], d[2])),
                               __dmul_rn(d[1], d[1]));
    r_out[i] = __dsqrt_rn(s);
}

// out[q] = src[order[seg_off[j] + (q - out_off[j])]] for q in output segment j:
// the first min(n, len_j) entries of every sorted segment (progenitors.py:52-53)
__global__ void segment_heads_kernel(const int64_t* __restrict__ src,
                                     const uint64_t* __restrict__ order,
                                     const int64_t* __restrict__ seg_off,
                                     const int64_t* __restrict__ out_off, int n_seg,
                                     int64_t n_out, int64_t* __restrict__ out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_out) return;
    const int j = find_segment(out_off, n_seg, q);
    out[q] = src[order[__ldg(seg_off + j) + (q - __ldg(out_off + j))]];
}

// flags[order[i]] = head[i]: "first occurrence" marks back in original order
// (np.unique(return_index=True), progenitors.py:82-84; the sort is stable, so
// the head of a run is the smallest original position)
__global__ void scatter_flags_kernel(const uint16_t* __restrict__ head,
                                     const uint64_t* __restrict__ order, int64_t n,
                                     uint16_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[order[i]] = head[i];
}

// pos_out[i] = vals[k] where keys[k] == query[i] - bias (keys ascending, unique),
// else -1; queries whose flag is 0 are skipped (-1).  np.in1d + myin1d,
// progenitors.py:95-99.  With segments (q_seg / key_off given) the search is
// confined to the query's own segment: the per-halo myin1d of
// postprocessing.py:222-232.
__global__ void lookup_sorted_kernel(const uint64_t* __restrict__ keys,
                                     const uint64_t* __restrict__ vals, int64_t n_keys,
                                     const int64_t* __restrict__ query,
                                     const uint16_t* __restrict__ flags, int64_t bias,
                                     const int32_t* __restrict__ q_seg,
                                     const int64_t* __restrict__ key_off, int64_t m,
                                     int64_t* __restrict__ pos_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int64_t res = -1;
    if (flags == nullptr || flags[i]) {
        int64_t lo = 0, hi = n_keys;
        bool ok = true;
        if (q_seg != nullptr) {
            const int32_t sgm = q_seg[i];
            ok = sgm >= 0;
            if (ok) { lo = __ldg(key_off + sgm); hi = __ldg(key_off + sgm + 1); }
        }
        if (ok) {
            const uint64_t q = (uint64_t)(query[i] - bias);
            const int64_t end = hi;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(keys + mid) < q) lo = mid + 1; else hi = mid;
            }
            if (lo < end && __ldg(keys + lo) == q)
                res = vals != nullptr ? (int64_t)__ldg(vals + lo) : lo;
        }
    }
    pos_out[i] = res;
}

// host halo of every tracked particle and the key of the plurality vote:
// key = descendant << 32 | halo, or ~0 when the particle is in no halo
// (progenitors.py:92-106)
__global__ void vote_keys_kernel(const int64_t* __restrict__ where,
                                 const int64_t* __restrict__ halo_off, int n_halos,
                                 const int64_t* __restrict__ tracked_off, int n_desc,
                                 int64_t m, uint64_t* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int64_t w = where[i];
    if (w < 0) { keys[i] = ~0ull; return; }
    const uint64_t halo = (uint64_t)find_segment(halo_off, n_halos, w);
    const uint64_t desc = (uint64_t)find_segment(tracked_off, n_desc, i);
    keys[i] = (desc << 32) | halo;
}

// one thread per element of the SORTED vote keys: the head of a run counts its
// length and offers (count, smallest halo wins ties) to its descendant
// (np.unique(return_counts) + first argmax, progenitors.py:107-115)
__global__ void vote_reduce_kernel(const uint64_t* __restrict__ keys, int64_t m,
                                   unsigned long long* __restrict__ best) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint64_t k = keys[i];
    if (k == ~0ull) return;
    if (i > 0 && keys[i - 1] == k) return;          // not a run head
    int64_t lo = i, hi = m;                          // first position with key > k
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) <= k) lo = mid + 1; else hi = mid;
    }
    const unsigned long long count = (unsigned long long)(lo - i);
    const unsigned long long halo = k & 0xFFFFFFFFull;
    atomicMax(best + (k >> 32), (count << 32) | (0xFFFFFFFFull - halo));
}

__global__ void vote_decode_kernel(const unsigned long long* __restrict__ best, int n_desc,
                                   int64_t* __restrict__ out) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_desc) return;
    const unsigned long long b = best[d];
    out[d] = b == 0 ? -1 : (int64_t)(0xFFFFFFFFull - (b & 0xFFFFFFFFull));
}

// marks[i] = 1 where angle (float16 bits) > cut (float16 comparison, NaN -> 0):
// `angles > angle_cut`, postprocessing.py:124-127
__global__ void angle_cut_kernel(const uint16_t* __restrict__ angles, int64_t n, double cut,
                                 uint16_t* __restrict__ marks) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = (double)__half2float(__ushort_as_half(angles[i]));
    marks[i] = (a > cut) ? 1 : 0;
}

// seg_out[i] = table[seg(i)] where seg(i) = segment of element i  (-1 allowed)
__global__ void expand_segments_kernel(const int64_t* __restrict__ seg_off, int n_seg,
                                       const int32_t* __restrict__ table, int64_t n,
                                       int32_t* __restrict__ seg_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = find_segment(seg_off, n_seg, i);
    seg_out[i] = table != nullptr ? table[s] : s;
}

}  // namespace

extern "C" int oa_central_radii(const void* pos, int data_dtype, const int64_t* cur_off,
                                int n_regions, const double* centres, int centre_f32,
                                int periodic, const double* box_host, int64_t n,
                                double* r_out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(pos && cur_off && centres && r_out && n_regions >= 1,
               "oa_central_radii: bad arguments");
    const double bx = periodic ? box_host[0] : 0, by = periodic ? box_host[1] : 0,
                 bz = periodic ? box_host[2] : 0;
    if (data_dtype == OA_F64)
        central_radii_kernel<double><<<blocks_for(n, 256), 256, 0, st>>>(
            static_cast<const double*>(pos), cur_off, n_regions, centres, centre_f32,
            periodic, bx, by, bz, n, r_out);
    else
        central_radii_kernel<float><<<blocks_for(n, 256), 256, 0, st>>>(
            static_cast<const float*>(pos), cur_off, n_regions, centres, centre_f32,
            periodic, bx, by, bz, n, r_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_segment_heads(const int64_t* src, const uint64_t* order,
                                const int64_t* seg_off, const int64_t* out_off, int n_seg,
                                int64_t n_out, int64_t* out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_out <= 0) return OA_OK;
    OA_REQUIRE(src && order && seg_off && out_off && out && n_seg >= 1,
               "oa_segment_heads: bad arguments");
    segment_heads_kernel<<<blocks_for(n_out, 256), 256, 0, st>>>(src, order, seg_off, out_off,
                                                                 n_seg, n_out, out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_scatter_flags(const uint16_t* head, const uint64_t* order, int64_t n,
                                uint16_t* flags, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(head && order && flags, "oa_scatter_flags: NULL pointer");
    scatter_flags_kernel<<<blocks_for(n, 256), 256, 0, st>>>(head, order, n, flags);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_lookup_sorted(const uint64_t* keys, const uint64_t* vals, int64_t n_keys,
                                const int64_t* query, const uint16_t* flags, int64_t bias,
                                const int32_t* q_seg, const int64_t* key_off, int64_t m,
                                int64_t* pos_out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (m <= 0) return OA_OK;
    OA_REQUIRE(query && pos_out && (keys || n_keys == 0) && (!q_seg || key_off),
               "oa_lookup_sorted: bad arguments");
    lookup_sorted_kernel<<<blocks_for(m, 256), 256, 0, st>>>(keys, vals, n_keys, query, flags,
                                                             bias, q_seg, key_off, m, pos_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_vote_keys(const int64_t* where, const int64_t* halo_off, int n_halos,
                            const int64_t* tracked_off, int n_desc, int64_t m,
                            uint64_t* keys, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (m <= 0) return OA_OK;
    OA_REQUIRE(where && halo_off && tracked_off && keys && n_halos >= 1 && n_desc >= 1,
               "oa_vote_keys: bad arguments");
    vote_keys_kernel<<<blocks_for(m, 256), 256, 0, st>>>(where, halo_off, n_halos, tracked_off,
                                                         n_desc, m, keys);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_vote_reduce(const uint64_t* sorted_keys, int64_t m, int n_desc,
                              uint64_t* best_ws, int64_t* out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OA_REQUIRE(n_desc >= 0 && (n_desc == 0 || (best_ws && out)), "oa_vote_reduce: bad arguments");
    if (n_desc == 0) return OA_OK;
    OA_CUDA_CHECK(cudaMemsetAsync(best_ws, 0, sizeof(uint64_t) * (size_t)n_desc, st));
    if (m > 0) {
        OA_REQUIRE(sorted_keys, "oa_vote_reduce: NULL keys");
        vote_reduce_kernel<<<blocks_for(m, 256), 256, 0, st>>>(
            sorted_keys, m, reinterpret_cast<unsigned long long*>(best_ws));
        OA_LAUNCH_CHECK();
    }
    vote_decode_kernel<<<blocks_for(n_desc, 256), 256, 0, st>>>(
        reinterpret_cast<const unsigned long long*>(best_ws), n_desc, out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_angle_cut(const uint16_t* angles, int64_t n, double cut, uint16_t* marks,
                            void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(angles && marks, "oa_angle_cut: NULL pointer");
    angle_cut_kernel<<<blocks_for(n, 256), 256, 0, st>>>(angles, n, cut, marks);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_expand_segments(const int64_t* seg_off, int n_seg, const int32_t* table,
                                  int64_t n, int32_t* seg_out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(seg_off && seg_out && n_seg >= 1, "oa_expand_segments: bad arguments");
    expand_segments_kernel<<<blocks_for(n, 256), 256, 0, st>>>(seg_off, n_seg, table, n, seg_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

// ---- incremental collation (postprocessing.py:121-141) --------------------------------------
// The collated state of all snapshots so far is a table of (pool, ID, count),
// ascending in (pool, ID) -- per pool exactly np.unique(..., return_counts=True)
// of that pool's event IDs.  A new snapshot's events, sorted and run-length
// encoded the same way, are MERGED into it instead of re-sorting the whole
// history: (1) every new key finds its lower bound in the table; an equal key
// adds its count, the others are "misses"; (2) the table and the misses are
// written to the merged table: a table row moves up by the number of misses in
// front of it, miss m lands at lower_bound + m.
namespace {
__device__ __forceinline__ bool key_less(int64_t sa, int64_t ia, int64_t sb, int64_t ib) {
    return sa < sb || (sa == sb && ia < ib);
}

__global__ void merge_find_kernel(const int64_t* __restrict__ t_seg,
                                  const int64_t* __restrict__ t_ids, int64_t* t_cnt, int64_t n_tab,
                                  const int64_t* __restrict__ n_seg, const int64_t* __restrict__ n_ids,
                                  const int64_t* __restrict__ n_cnt, int64_t n_new,
                                  int64_t* __restrict__ lb, uint16_t* __restrict__ miss) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_new) return;
    const int64_t s = n_seg[k], id = n_ids[k];
    int64_t lo = 0, hi = n_tab;                      // first row with key >= (s, id)
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (key_less(t_seg[mid], t_ids[mid], s, id)) lo = mid + 1; else hi = mid;
    }
    lb[k] = lo;
    const bool hit = lo < n_tab && t_seg[lo] == s && t_ids[lo] == id;
    miss[k] = hit ? 0 : 1;
    if (hit) t_cnt[lo] += n_cnt[k];                  // new keys are unique: no conflict
}

__global__ void merge_place_kernel(const int64_t* __restrict__ t_seg,
                                   const int64_t* __restrict__ t_ids,
                                   const int64_t* __restrict__ t_cnt, int64_t n_tab,
                                   const int64_t* __restrict__ n_seg, const int64_t* __restrict__ n_ids,
                                   const int64_t* __restrict__ n_cnt,
                                   const int64_t* __restrict__ lb, const int64_t* __restrict__ msel,
                                   int64_t n_miss, int64_t* __restrict__ o_seg,
                                   int64_t* __restrict__ o_ids, int64_t* __restrict__ o_cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tab) {
        int64_t lo = 0, hi = n_miss;                 // misses with lower bound <= i
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (lb[msel[mid]] <= i) lo = mid + 1; else hi = mid;
        }
        o_seg[i + lo] = t_seg[i];
        o_ids[i + lo] = t_ids[i];
        o_cnt[i + lo] = t_cnt[i];
    } else if (i < n_tab + n_miss) {
        const int64_t m = i - n_tab, k = msel[m], at = lb[k] + m;
        o_seg[at] = n_seg[k];
        o_ids[at] = n_ids[k];
        o_cnt[at] = n_cnt[k];
    }
}
}  // namespace

extern "C" int oa_merge_find(const int64_t* tab_seg, const int64_t* tab_ids, int64_t* tab_cnt,
                             int64_t n_tab, const int64_t* new_seg, const int64_t* new_ids,
                             const int64_t* new_cnt, int64_t n_new, int64_t* lb, uint16_t* miss,
                             void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_new <= 0) return OA_OK;
    OA_REQUIRE(n_tab >= 0 && new_seg && new_ids && new_cnt && lb && miss &&
               (n_tab == 0 || (tab_seg && tab_ids && tab_cnt)), "oa_merge_find: bad arguments");
    merge_find_kernel<<<blocks_for(n_new, 256), 256, 0, st>>>(tab_seg, tab_ids, tab_cnt, n_tab,
                                                              new_seg, new_ids, new_cnt, n_new, lb,
                                                              miss);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_merge_place(const int64_t* tab_seg, const int64_t* tab_ids,
                              const int64_t* tab_cnt, int64_t n_tab, const int64_t* new_seg,
                              const int64_t* new_ids, const int64_t* new_cnt, const int64_t* lb,
                              const int64_t* miss_sel, int64_t n_miss, int64_t* out_seg,
                              int64_t* out_ids, int64_t* out_cnt, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_tab + n_miss <= 0) return OA_OK;
    OA_REQUIRE(n_tab >= 0 && n_miss >= 0 && out_seg && out_ids && out_cnt &&
               (n_tab == 0 || (tab_seg && tab_ids && tab_cnt)) &&
               (n_miss == 0 || (new_seg && new_ids && new_cnt && lb && miss_sel)),
               "oa_merge_place: bad arguments");
    merge_place_kernel<<<blocks_for(n_tab + n_miss, 256), 256, 0, st>>>(
        tab_seg, tab_ids, tab_cnt, n_tab, new_seg, new_ids, new_cnt, lb, miss_sel, n_miss, out_seg,
        out_ids, out_cnt);
    OA_LAUNCH_CHECK();
    return OA_OK;
}
