// LSD radix sort of (uint64 key, uint64 value) pairs, 8 bits per pass (sm_100a).
//
// Serves the places where the reference sorts: sorted `setdiff1d` outputs of
// the on-the-fly path (track_orbits_onthefly.py:145,168), `argsort` in
// progenitors.py:52 / utils.py:10 and `np.unique` in postprocessing.py:135.
// (Particle-to-previous-block matching itself does not sort -- see oa_track.cu.)
//
// Per pass: (1) per-tile digit histogram, (2) exclusive scan of the
// digit-major histogram matrix (two levels), (3) stable scatter using
// warp-level match-any ranking.  Stability comes from the fixed key order
// (warp, iteration, lane) inside a tile and tile-major global offsets.
#include "oa_common.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;      // 2048 keys per CTA
constexpr int RS_BINS = 256;
constexpr int SCAN_CHUNK = RS_THREADS * 8;          // histogram entries per scan CTA

inline int64_t rs_tiles(int64_t n) { return (n + RS_TILE - 1) / RS_TILE; }
inline int64_t rs_chunks(int64_t tiles) {
    return (tiles * RS_BINS + SCAN_CHUNK - 1) / SCAN_CHUNK;
}

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, uint32_t mask,
               int64_t n_tiles, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_bins[RS_BINS];
    s_bins[threadIdx.x] = 0;
    __syncthreads();
    const int64_t tile = blockIdx.x;
    const int64_t base = tile * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int64_t k = base + (int64_t)i * RS_THREADS + threadIdx.x;
        if (k < n) atomicAdd(&s_bins[(uint32_t)(__ldg(keys + k) >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * n_tiles + tile] = s_bins[threadIdx.x];
}

// level 1: exclusive scan inside chunks of SCAN_CHUNK entries (in place)
__global__ void __launch_bounds__(RS_THREADS)
rs_scan_chunks_kernel(uint32_t* __restrict__ hist, int64_t m,
                      uint32_t* __restrict__ chunk_totals) {
    const int64_t first = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * 8;
    uint32_t v[8];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[i] = (first + i < m) ? hist[first + i] : 0u;
        sum += v[i];
    }
    uint32_t total;
    uint32_t run = oa_block_exclusive_scan<RS_THREADS>(sum, &total);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (first + i < m) hist[first + i] = run;
        run += v[i];
    }
    if (threadIdx.x == 0) chunk_totals[blockIdx.x] = total;
}

// level 2: single-CTA exclusive scan of the chunk totals (in place)
__global__ void __launch_bounds__(1024)
rs_scan_totals_kernel(uint32_t* __restrict__ totals, int64_t n_chunks) {
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_chunks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < n_chunks ? totals[i] : 0u;
        uint32_t total;
        const uint32_t excl = oa_block_exclusive_scan<1024>(v, &total);
        const uint32_t carry = s_carry;
        if (i < n_chunks) totals[i] = carry + excl;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
}

template <bool HAS_VALS>
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint64_t* __restrict__ keys_in,
                  const uint64_t* __restrict__ vals_in,
                  uint64_t* __restrict__ keys_out, uint64_t* __restrict__ vals_out,
                  int64_t n, int shift, uint32_t mask, int64_t n_tiles,
                  const uint32_t* __restrict__ hist,
                  const uint32_t* __restrict__ chunk_prefix) {
    __shared__ uint32_t s_warp_hist[RS_WARPS][RS_BINS];
    __shared__ uint32_t s_gbase[RS_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS)
        (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();

    const int64_t tile = blockIdx.x;
    const int64_t wbase = tile * RS_TILE + (int64_t)warp * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int64_t k = wbase + i * 32 + lane;
        const bool valid = k < n;
        key[i] = valid ? __ldg(keys_in + k) : 0ull;
        // invalid lanes get a digit outside 0..255 so they only match each other
        const uint32_t d = valid ? ((uint32_t)(key[i] >> shift) & mask) : 0x100u;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        const uint32_t before = __popc(peers & lt_mask);
        uint32_t pre = 0;
        if (valid) pre = s_warp_hist[warp][d];
        __syncwarp();
        if (valid && before == 0) s_warp_hist[warp][d] = pre + __popc(peers);
        __syncwarp();
        rank[i] = pre + before;
    }
    __syncthreads();

    {   // per digit: exclusive scan over the warps + global base of this tile
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t t = s_warp_hist[w][d];
            s_warp_hist[w][d] = run;
            run += t;
        }
        const int64_t e = (int64_t)d * n_tiles + tile;
        s_gbase[d] = hist[e] + chunk_prefix[e / SCAN_CHUNK];
    }
    __syncthreads();

#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int64_t k = wbase + i * 32 + lane;
        if (k < n) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
            const int64_t pos = (int64_t)s_gbase[d] + s_warp_hist[warp][d] + rank[i];
            keys_out[pos] = key[i];
            if (HAS_VALS) vals_out[pos] = __ldg(vals_in + k);
        }
    }
}

__global__ void minmax_i64_kernel(const int64_t* __restrict__ x, int64_t n,
                                  int64_t* __restrict__ out) {
    int64_t lo = INT64_MAX, hi = INT64_MIN;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = __ldg(x + i);
        lo = min(lo, v);
        hi = max(hi, v);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo = min(lo, __shfl_down_sync(0xFFFFFFFFu, lo, d));
        hi = max(hi, __shfl_down_sync(0xFFFFFFFFu, hi, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(reinterpret_cast<long long*>(out), (long long)lo);
        atomicMax(reinterpret_cast<long long*>(out + 1), (long long)hi);
    }
}

__global__ void minmax_init_kernel(int64_t* out) {
    out[0] = INT64_MAX;
    out[1] = INT64_MIN;
}

struct SortWs {
    uint64_t* tmp_keys;
    uint64_t* tmp_vals;
    uint32_t* hist;
    uint32_t* chunk_totals;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline SortWs carve(void* ws, int64_t n) {
    const int64_t tiles = rs_tiles(n);
    char* p = static_cast<char*>(ws);
    SortWs w;
    w.tmp_keys = reinterpret_cast<uint64_t*>(p);
    p += align_up((size_t)n * 8, 256);
    w.tmp_vals = reinterpret_cast<uint64_t*>(p);
    p += align_up((size_t)n * 8, 256);
    w.hist = reinterpret_cast<uint32_t*>(p);
    p += align_up((size_t)tiles * RS_BINS * 4, 256);
    w.chunk_totals = reinterpret_cast<uint32_t*>(p);
    return w;
}

}  // namespace

extern "C" size_t oa_sort_workspace_bytes(int64_t n) {
    if (n < 1) n = 1;
    const int64_t tiles = rs_tiles(n);
    return 2 * align_up((size_t)n * 8, 256) + align_up((size_t)tiles * RS_BINS * 4, 256) +
           align_up((size_t)rs_chunks(tiles) * 4, 256) + 256;
}

extern "C" int oa_sort_pairs_u64(const uint64_t* keys_in, const uint64_t* vals_in,
                                 uint64_t* keys_out, uint64_t* vals_out, int64_t n,
                                 int begin_bit, int end_bit, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OA_REQUIRE(n >= 0 && begin_bit >= 0 && end_bit <= 64 && begin_bit <= end_bit,
               "oa_sort_pairs_u64: bad arguments");
    if (n == 0) return OA_OK;
    OA_REQUIRE(n < ((int64_t)1 << 32), "oa_sort_pairs_u64: n must be < 2^32");
    OA_REQUIRE(keys_in && keys_out && workspace, "oa_sort_pairs_u64: NULL pointer");
    OA_REQUIRE((vals_in == nullptr) == (vals_out == nullptr),
               "oa_sort_pairs_u64: vals_in / vals_out must both be given or both NULL");
    OA_REQUIRE(workspace_bytes >= oa_sort_workspace_bytes(n),
               "oa_sort_pairs_u64: workspace too small");
    OA_REQUIRE(keys_in != keys_out && (vals_in == nullptr || vals_in != vals_out),
               "oa_sort_pairs_u64: in-place sorting is not supported");
    const bool has_vals = vals_in != nullptr;
    const int passes = (end_bit - begin_bit + 7) / 8;
    if (passes == 0) {
        OA_CUDA_CHECK(cudaMemcpyAsync(keys_out, keys_in, (size_t)n * 8,
                                      cudaMemcpyDeviceToDevice, st));
        if (has_vals)
            OA_CUDA_CHECK(cudaMemcpyAsync(vals_out, vals_in, (size_t)n * 8,
                                          cudaMemcpyDeviceToDevice, st));
        return OA_OK;
    }
    SortWs w = carve(workspace, n);
    const int64_t tiles = rs_tiles(n);
    const int64_t m = tiles * RS_BINS;
    const int64_t chunks = rs_chunks(tiles);

    const uint64_t* src_k = keys_in;
    const uint64_t* src_v = vals_in;
    for (int pass = 0; pass < passes; ++pass) {
        // the last pass must land in *_out: passes-1-pass even -> out
        const bool to_out = ((passes - 1 - pass) % 2) == 0;
        uint64_t* dst_k = to_out ? keys_out : w.tmp_keys;
        uint64_t* dst_v = to_out ? vals_out : w.tmp_vals;
        const int shift = begin_bit + 8 * pass;
        const int width = (end_bit - shift) < 8 ? (end_bit - shift) : 8;
        const uint32_t mask = (1u << width) - 1u;
        rs_hist_kernel<<<(unsigned)tiles, RS_THREADS, 0, st>>>(src_k, n, shift, mask, tiles, w.hist);
        OA_LAUNCH_CHECK();
        rs_scan_chunks_kernel<<<(unsigned)chunks, RS_THREADS, 0, st>>>(w.hist, m,
                                                                       w.chunk_totals);
        OA_LAUNCH_CHECK();
        rs_scan_totals_kernel<<<1, 1024, 0, st>>>(w.chunk_totals, chunks);
        OA_LAUNCH_CHECK();
        if (has_vals)
            rs_scatter_kernel<true><<<(unsigned)tiles, RS_THREADS, 0, st>>>(
                src_k, src_v, dst_k, dst_v, n, shift, mask, tiles, w.hist, w.chunk_totals);
        else
            rs_scatter_kernel<false><<<(unsigned)tiles, RS_THREADS, 0, st>>>(
                src_k, nullptr, dst_k, nullptr, n, shift, mask, tiles, w.hist, w.chunk_totals);
        OA_LAUNCH_CHECK();
        src_k = dst_k;
        src_v = dst_v;
    }
    return OA_OK;
}

extern "C" int oa_minmax_i64(const int64_t* x, int64_t n, int64_t* out_dev, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OA_REQUIRE(out_dev && (x || n == 0), "oa_minmax_i64: NULL pointer");
    minmax_init_kernel<<<1, 1, 0, st>>>(out_dev);
    OA_LAUNCH_CHECK();
    if (n > 0) {
        int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
        if (blocks > OA_NUM_SMS * 8) blocks = OA_NUM_SMS * 8;
        minmax_i64_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, out_dev);
        OA_LAUNCH_CHECK();
    }
    return OA_OK;
}
