// Segment helpers for the sorted per-halo ID lists of the on-the-fly path
// (np.setdiff1d returns sorted unique values, track_orbits_onthefly.py:145,168)
// and for the per-halo unique/count reductions of postprocessing.py:133-141.
#include "oa_common.cuh"

namespace {

// segment of element i: last s with seg_off[s] <= i   (seg_off has n_seg+1 entries)
OA_D int find_segment(const int64_t* __restrict__ seg_off, int n_seg, int64_t i) {
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(seg_off + mid) <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// key_lo = (id - id_min) for segments that must be sorted by ID, otherwise the
// element's own position (keeps the original order); key_hi = segment index.
__global__ void segment_sort_keys_kernel(const int64_t* __restrict__ ids, int64_t n,
                                         const int64_t* __restrict__ seg_off, int n_seg,
                                         const uint8_t* __restrict__ sort_flag,
                                         const int64_t* __restrict__ id_minmax,
                                         uint64_t* __restrict__ key_lo,
                                         uint64_t* __restrict__ key_hi,
                                         uint64_t* __restrict__ index) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = find_segment(seg_off, n_seg, i);
    const bool sorted = sort_flag == nullptr || sort_flag[s] != 0;
    key_lo[i] = sorted ? (uint64_t)(ids[i] - __ldg(id_minmax)) : (uint64_t)i;
    key_hi[i] = (uint64_t)s;
    index[i] = (uint64_t)i;
}

// run-length encoding of a sorted (segment, id) sequence: head[i] = 1 where a new
// (segment, id) run starts
__global__ void run_heads_kernel(const uint64_t* __restrict__ seg, const int64_t* __restrict__ ids,
                                 int64_t n, uint16_t* __restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || seg[i] != seg[i - 1] || ids[i] != ids[i - 1]) ? 1 : 0;
}

// counts[k] = start of run k+1 - start of run k   (starts = selected positions)
__global__ void run_lengths_kernel(const int64_t* __restrict__ starts, int64_t n_runs,
                                   int64_t n, int64_t* __restrict__ counts) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_runs) return;
    const int64_t next = (k + 1 < n_runs) ? starts[k + 1] : n;
    counts[k] = next - starts[k];
}

// Multi-way merge of n_lists ascending key lists laid out back to back
// (list r = [list_off[r], list_off[r+1])): the output position of an element is
// its own rank plus the number of smaller keys in every other list (keys are
// distinct across lists: they are positions in one unsharded snapshot).
__global__ void merge_lists_kernel(const int64_t* __restrict__ keys,
                                   const int64_t* __restrict__ ids,
                                   const uint16_t* __restrict__ angles, int64_t n,
                                   const int64_t* __restrict__ list_off, int n_lists,
                                   int64_t* __restrict__ ids_out,
                                   uint16_t* __restrict__ angles_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t key = keys[i];
    int64_t pos = 0;
    for (int r = 0; r < n_lists; ++r) {
        const int64_t b = __ldg(list_off + r), e = __ldg(list_off + r + 1);
        if (i >= b && i < e) { pos += i - b; continue; }
        int64_t lo = b, hi = e;                 // first element of list r with key' >= key
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(keys + mid) < key) lo = mid + 1; else hi = mid;
        }
        pos += lo - b;
    }
    ids_out[pos] = ids[i];
    angles_out[pos] = angles[i];
}

inline unsigned blocks_for(int64_t n, int threads) {
    return (unsigned)((n + threads - 1) / threads);
}

}  // namespace

extern "C" int oa_segment_sort_keys(const int64_t* ids, int64_t n, const int64_t* seg_off,
                                    int n_seg, const uint8_t* sort_flag,
                                    const int64_t* id_minmax, uint64_t* key_lo,
                                    uint64_t* key_hi, uint64_t* index, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(ids && seg_off && id_minmax && key_lo && key_hi && index && n_seg >= 1,
               "oa_segment_sort_keys: bad arguments");
    segment_sort_keys_kernel<<<blocks_for(n, 256), 256, 0, st>>>(
        ids, n, seg_off, n_seg, sort_flag, id_minmax, key_lo, key_hi, index);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_merge_event_lists(const int64_t* keys, const int64_t* ids,
                                    const uint16_t* angles, int64_t n,
                                    const int64_t* list_off, int n_lists, int64_t* ids_out,
                                    uint16_t* angles_out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(keys && ids && angles && list_off && ids_out && angles_out && n_lists >= 1,
               "oa_merge_event_lists: bad arguments");
    merge_lists_kernel<<<blocks_for(n, 256), 256, 0, st>>>(keys, ids, angles, n, list_off,
                                                           n_lists, ids_out, angles_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_run_heads(const uint64_t* seg, const int64_t* ids, int64_t n,
                            uint16_t* head, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(seg && ids && head, "oa_run_heads: NULL pointer");
    run_heads_kernel<<<blocks_for(n, 256), 256, 0, st>>>(seg, ids, n, head);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_run_lengths(const int64_t* starts, int64_t n_runs, int64_t n,
                              int64_t* counts, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_runs <= 0) return OA_OK;
    OA_REQUIRE(starts && counts, "oa_run_lengths: NULL pointer");
    run_lengths_kernel<<<blocks_for(n_runs, 256), 256, 0, st>>>(starts, n_runs, n, counts);
    OA_LAUNCH_CHECK();
    return OA_OK;
}
