// Ordered selection / compaction and small gathers (sm_100a).
//
// The fused tracking kernel leaves one 16-bit mark per PREVIOUS-block particle;
// the reference emits events in previous-block order (track_orbits.py:315-316),
// so the event list is "positions whose mark satisfies a predicate, ascending".
// Two-step stream compaction: per-tile counts -> exclusive scan -> ordered
// gather with a block-wide prefix sum (warp shuffles + ballot-free blocked
// layout: each thread owns 8 consecutive marks = one 16-byte load).
#include "oa_common.cuh"

namespace {

constexpr int SEL_THREADS = 256;
constexpr int SEL_ITEMS = 8;
constexpr int SEL_TILE = SEL_THREADS * SEL_ITEMS;   // 2048 marks = 4 KB

OA_D bool sel_pred(uint16_t m, int op, uint16_t value) {
    return op == OA_SEL_EQ ? (m == value) : (m != value);
}

// flags of the 8 marks owned by this thread, bit i = item i selected
OA_D uint32_t load_flags(const uint16_t* __restrict__ marks, int64_t n,
                         int64_t first, int op, uint16_t value) {
    uint32_t bits = 0;
    if (first + SEL_ITEMS <= n) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(marks + first));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (sel_pred((uint16_t)(w[i] & 0xFFFFu), op, value)) bits |= 1u << (2 * i);
            if (sel_pred((uint16_t)(w[i] >> 16), op, value)) bits |= 1u << (2 * i + 1);
        }
    } else {
        for (int i = 0; i < SEL_ITEMS; ++i)
            if (first + i < n && sel_pred(marks[first + i], op, value)) bits |= 1u << i;
    }
    return bits;
}

__global__ void __launch_bounds__(SEL_THREADS)
sel_count_kernel(const uint16_t* __restrict__ marks, int64_t n, int op,
                 uint16_t value, uint32_t* __restrict__ tile_counts) {
    const int64_t tile = blockIdx.x;
    const int64_t first = tile * SEL_TILE + (int64_t)threadIdx.x * SEL_ITEMS;
    uint32_t cnt = first < n ? __popc(load_flags(marks, n, first, op, value)) : 0u;
    uint32_t total;
    (void)oa_block_exclusive_scan<SEL_THREADS>(cnt, &total);
    if (threadIdx.x == 0) tile_counts[tile] = total;
}

// single-CTA exclusive scan of the tile counts (<= ~1M tiles for 2^31 marks)
__global__ void __launch_bounds__(1024)
sel_scan_kernel(const uint32_t* __restrict__ tile_counts, int64_t n_tiles,
                int64_t* __restrict__ tile_offsets, int64_t* __restrict__ total_dev) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_tiles; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < n_tiles ? (int64_t)tile_counts[i] : 0;
        int64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int64_t w = s_warp[lane];
            int64_t winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int64_t t = __shfl_up_sync(0xFFFFFFFFu, winc, d);
                if (lane >= d) winc += t;
            }
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        const int64_t carry = s_carry;
        const int64_t excl = carry + s_warp[warp] + inc - v;
        if (i < n_tiles) tile_offsets[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_dev = s_carry;
}

__global__ void __launch_bounds__(SEL_THREADS)
sel_gather_kernel(const uint16_t* __restrict__ marks, int64_t n, int op,
                  uint16_t value, const int64_t* __restrict__ tile_offsets,
                  int64_t* __restrict__ sel_out) {
    const int64_t tile = blockIdx.x;
    const int64_t first = tile * SEL_TILE + (int64_t)threadIdx.x * SEL_ITEMS;
    const uint32_t bits = first < n ? load_flags(marks, n, first, op, value) : 0u;
    uint32_t total;
    const uint32_t excl = oa_block_exclusive_scan<SEL_THREADS>(__popc(bits), &total);
    if (bits) {
        int64_t o = tile_offsets[tile] + excl;
#pragma unroll
        for (int i = 0; i < SEL_ITEMS; ++i)
            if (bits & (1u << i)) sel_out[o++] = first + i;
    }
}

// Event variant: the selected marks ARE the float16 event angles, and the event
// ID is in the previous record at the same position -- one kernel writes the
// three event arrays (positions, IDs, angles).
// (REC: OaRec<float>, OaRec<double>, or IdOnly = the snapshot's plain ID array)
struct IdOnly { int64_t id; };
template <typename REC>
__global__ void __launch_bounds__(SEL_THREADS)
sel_gather_events_kernel(const uint16_t* __restrict__ marks, int64_t n,
                         const int64_t* __restrict__ tile_offsets,
                         const REC* __restrict__ rec, int64_t* __restrict__ sel_out,
                         int64_t* __restrict__ ids_out, uint16_t* __restrict__ angles_out) {
    const int64_t tile = blockIdx.x;
    const int64_t first = tile * SEL_TILE + (int64_t)threadIdx.x * SEL_ITEMS;
    const uint32_t bits = first < n ? load_flags(marks, n, first, OA_SEL_NE, OA_NO_EVENT) : 0u;
    uint32_t total;
    const uint32_t excl = oa_block_exclusive_scan<SEL_THREADS>(__popc(bits), &total);
    if (bits) {
        int64_t o = tile_offsets[tile] + excl;
#pragma unroll
        for (int i = 0; i < SEL_ITEMS; ++i)
            if (bits & (1u << i)) {
                sel_out[o] = first + i;
                ids_out[o] = rec[first + i].id;
                angles_out[o] = marks[first + i];
                ++o;
            }
    }
}

// Every consumer of a selection takes `n_sel` (a host-side upper bound) and an
// optional device pointer `n_dev` to the exact count, so that a whole snapshot
// can be enqueued without a host synchronisation in the middle.
OA_D int64_t actual_count(int64_t n_sel, const int64_t* __restrict__ n_dev) {
    if (n_dev == nullptr) return n_sel;
    const int64_t v = __ldg(n_dev);
    return v < n_sel ? v : n_sel;
}

__global__ void seg_offsets_kernel(const int64_t* __restrict__ sel, int64_t n_sel,
                                   const int64_t* __restrict__ n_dev,
                                   const int64_t* __restrict__ seg_begin, int n_seg,
                                   int64_t* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_seg) return;
    const int64_t key = seg_begin[k];
    int64_t lo = 0, hi = actual_count(n_sel, n_dev);   // first i with sel[i] >= key
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sel[mid] < key) lo = mid + 1; else hi = mid;
    }
    out[k] = lo;
}

template <typename TF>
__global__ void gather_rec_ids_kernel(const OaRec<TF>* __restrict__ rec,
                                      const int64_t* __restrict__ sel, int64_t n_sel,
                                      const int64_t* __restrict__ n_dev,
                                      int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < actual_count(n_sel, n_dev)) out[i] = rec[sel[i]].id;
}

template <typename T>
__global__ void gather_kernel(const T* __restrict__ src, const int64_t* __restrict__ sel,
                              int64_t n_sel, const int64_t* __restrict__ n_dev,
                              T* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < actual_count(n_sel, n_dev)) out[i] = src[sel[i]];
}

__global__ void mark_unmatched_kernel(const int64_t* __restrict__ match, int64_t n,
                                      uint16_t* __restrict__ marks) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) marks[i] = match[i] < 0 ? (uint16_t)1 : (uint16_t)0;
}

template <typename TF>
__global__ void set_rec_angles_kernel(OaRec<TF>* __restrict__ rec,
                                      const uint16_t* __restrict__ angles, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rec[i].angle = __ushort_as_half(angles[i]);
}

__global__ void fill_u16_kernel(uint16_t* __restrict__ dst, int64_t n, uint16_t value) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = value;
}

inline unsigned blocks_for(int64_t n, int threads) {
    return (unsigned)((n + threads - 1) / threads);
}

inline int64_t sel_tiles(int64_t n) { return (n + SEL_TILE - 1) / SEL_TILE; }

// workspace layout: [tile_offsets int64 x tiles][tile_counts u32 x tiles]
inline int64_t* ws_offsets(void* ws) { return static_cast<int64_t*>(ws); }
inline uint32_t* ws_counts(void* ws, int64_t tiles) {
    return reinterpret_cast<uint32_t*>(static_cast<int64_t*>(ws) + tiles);
}

}  // namespace

extern "C" size_t oa_select_workspace_bytes(int64_t n) {
    const int64_t t = sel_tiles(n > 0 ? n : 1);
    return (size_t)t * (sizeof(int64_t) + sizeof(uint32_t)) + 16;
}

extern "C" int oa_select_count(const uint16_t* marks, int64_t n, int op,
                               uint16_t value, void* workspace,
                               size_t workspace_bytes, int64_t* total_dev,
                               void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OA_REQUIRE(n >= 0 && total_dev, "oa_select_count: bad arguments");
    OA_REQUIRE(op == OA_SEL_NE || op == OA_SEL_EQ, "oa_select_count: bad op");
    if (n == 0) {
        OA_CUDA_CHECK(cudaMemsetAsync(total_dev, 0, sizeof(int64_t), st));
        return OA_OK;
    }
    OA_REQUIRE(marks && workspace &&
               workspace_bytes >= oa_select_workspace_bytes(n),
               "oa_select_count: workspace too small");
    OA_REQUIRE((reinterpret_cast<uintptr_t>(marks) & 15) == 0,
               "oa_select_count: marks must be 16-byte aligned");
    const int64_t tiles = sel_tiles(n);
    sel_count_kernel<<<(unsigned)tiles, SEL_THREADS, 0, st>>>(
        marks, n, op, value, ws_counts(workspace, tiles));
    OA_LAUNCH_CHECK();
    sel_scan_kernel<<<1, 1024, 0, st>>>(ws_counts(workspace, tiles), tiles,
                                        ws_offsets(workspace), total_dev);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_select_gather(const uint16_t* marks, int64_t n, int op,
                                uint16_t value, const void* workspace,
                                int64_t* sel_out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) return OA_OK;
    OA_REQUIRE(marks && workspace && sel_out, "oa_select_gather: NULL pointer");
    const int64_t tiles = sel_tiles(n);
    sel_gather_kernel<<<(unsigned)tiles, SEL_THREADS, 0, st>>>(
        marks, n, op, value, ws_offsets(const_cast<void*>(workspace)), sel_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_select_gather_events(const uint16_t* marks, int64_t n,
                                       const void* workspace, const void* rec,
                                       int frame_dtype, int64_t* sel_out, int64_t* ids_out,
                                       uint16_t* angles_out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) return OA_OK;
    OA_REQUIRE(marks && workspace && rec && sel_out && ids_out && angles_out,
               "oa_select_gather_events: NULL pointer");
    const int64_t tiles = sel_tiles(n);
    const int64_t* offs = ws_offsets(const_cast<void*>(workspace));
    if (frame_dtype == OA_F64)
        sel_gather_events_kernel<OaRec<double>><<<(unsigned)tiles, SEL_THREADS, 0, st>>>(
            marks, n, offs, static_cast<const OaRec<double>*>(rec), sel_out, ids_out, angles_out);
    else
        sel_gather_events_kernel<OaRec<float>><<<(unsigned)tiles, SEL_THREADS, 0, st>>>(
            marks, n, offs, static_cast<const OaRec<float>*>(rec), sel_out, ids_out, angles_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_select_gather_events_ids(const uint16_t* marks, int64_t n,
                                           const void* workspace, const int64_t* ids,
                                           int64_t* sel_out, int64_t* ids_out,
                                           uint16_t* angles_out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) return OA_OK;
    OA_REQUIRE(marks && workspace && ids && sel_out && ids_out && angles_out,
               "oa_select_gather_events_ids: NULL pointer");
    const int64_t tiles = sel_tiles(n);
    const int64_t* offs = ws_offsets(const_cast<void*>(workspace));
    sel_gather_events_kernel<IdOnly><<<(unsigned)tiles, SEL_THREADS, 0, st>>>(
        marks, n, offs, reinterpret_cast<const IdOnly*>(ids), sel_out, ids_out, angles_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_segment_offsets(const int64_t* sel, int64_t n_sel,
                                  const int64_t* n_dev, const int64_t* seg_begin,
                                  int n_seg, int64_t* offsets_out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_seg <= 0) return OA_OK;
    OA_REQUIRE(seg_begin && offsets_out && (sel || n_sel == 0),
               "oa_segment_offsets: NULL pointer");
    seg_offsets_kernel<<<blocks_for(n_seg, 128), 128, 0, st>>>(
        sel, n_sel, n_dev, seg_begin, n_seg, offsets_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_gather_record_ids(const void* rec, int frame_dtype,
                                    const int64_t* sel, int64_t n_sel,
                                    const int64_t* n_dev, int64_t* ids_out,
                                    void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_sel <= 0) return OA_OK;
    OA_REQUIRE(rec && sel && ids_out, "oa_gather_record_ids: NULL pointer");
    if (frame_dtype == OA_F64)
        gather_rec_ids_kernel<double><<<blocks_for(n_sel, 256), 256, 0, st>>>(
            static_cast<const OaRec<double>*>(rec), sel, n_sel, n_dev, ids_out);
    else
        gather_rec_ids_kernel<float><<<blocks_for(n_sel, 256), 256, 0, st>>>(
            static_cast<const OaRec<float>*>(rec), sel, n_sel, n_dev, ids_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_gather_u16(const uint16_t* src, const int64_t* sel, int64_t n_sel,
                             const int64_t* n_dev, uint16_t* out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_sel <= 0) return OA_OK;
    OA_REQUIRE(src && sel && out, "oa_gather_u16: NULL pointer");
    gather_kernel<uint16_t><<<blocks_for(n_sel, 256), 256, 0, st>>>(src, sel, n_sel,
                                                                    n_dev, out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_gather_i64(const int64_t* src, const int64_t* sel, int64_t n_sel,
                             const int64_t* n_dev, int64_t* out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_sel <= 0) return OA_OK;
    OA_REQUIRE(src && sel && out, "oa_gather_i64: NULL pointer");
    gather_kernel<int64_t><<<blocks_for(n_sel, 256), 256, 0, st>>>(src, sel, n_sel,
                                                                   n_dev, out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_gather_f(const void* src, int dtype, const int64_t* sel,
                           int64_t n_sel, const int64_t* n_dev, void* out,
                           void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_sel <= 0) return OA_OK;
    OA_REQUIRE(src && sel && out, "oa_gather_f: NULL pointer");
    if (dtype == OA_F64)
        gather_kernel<double><<<blocks_for(n_sel, 256), 256, 0, st>>>(
            static_cast<const double*>(src), sel, n_sel, n_dev, static_cast<double*>(out));
    else
        gather_kernel<float><<<blocks_for(n_sel, 256), 256, 0, st>>>(
            static_cast<const float*>(src), sel, n_sel, n_dev, static_cast<float*>(out));
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_mark_unmatched(const int64_t* match, int64_t n, uint16_t* marks,
                                 void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(match && marks, "oa_mark_unmatched: NULL pointer");
    mark_unmatched_kernel<<<blocks_for(n, 256), 256, 0, st>>>(match, n, marks);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_fill_u16(uint16_t* dst, int64_t n, uint16_t value, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(dst, "oa_fill_u16: NULL pointer");
    fill_u16_kernel<<<blocks_for(n, 256), 256, 0, st>>>(dst, n, value);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_set_record_angles(void* rec, int frame_dtype, const uint16_t* angles,
                                    int64_t n, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(rec && angles, "oa_set_record_angles: NULL pointer");
    if (frame_dtype == OA_F64)
        set_rec_angles_kernel<double><<<blocks_for(n, 256), 256, 0, st>>>(
            static_cast<OaRec<double>*>(rec), angles, n);
    else
        set_rec_angles_kernel<float><<<blocks_for(n, 256), 256, 0, st>>>(
            static_cast<OaRec<float>*>(rec), angles, n);
    OA_LAUNCH_CHECK();
    return OA_OK;
}
