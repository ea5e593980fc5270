// Error state, version and device queries of liborbit_b200.
#include "oa_common.cuh"
#include <string.h>
#include <thread>
#include <vector>

static thread_local char g_err[512] = "";

void oa_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* oa_last_error(void) { return g_err; }

extern "C" int oa_abi_version(void) { return OA_ABI_VERSION; }

extern "C" int oa_device_info(int* sm_count, int* cc_major, int* cc_minor,
                              int64_t* l2_bytes, int64_t* hbm_bytes) {
    int dev = 0;
    OA_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    OA_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (l2_bytes) *l2_bytes = (int64_t)prop.l2CacheSize;
    if (hbm_bytes) *hbm_bytes = (int64_t)prop.totalGlobalMem;
    return OA_OK;
}

// Copy by a kernel instead of a copy engine: for the few-kilobyte read-backs the
// host waits for every snapshot (event counts, exchange sizes).  The DMA engine of
// a direction serves the copies of ALL streams in issue order, so such a copy
// would sit behind megabytes of event lists that are in flight to the host; SM
// stores into (UVA-mapped) pinned host memory are independent of that queue.
namespace {
__global__ void copy_small_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src,
                                  size_t words) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words;
         i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
    __threadfence_system();
}
}  // namespace

extern "C" int oa_copy_small(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes == 0) return OA_OK;
    OA_REQUIRE(dst && src && bytes % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 3u) == 0,
               "oa_copy_small: pointers and size must be multiples of 4 bytes");
    const size_t words = bytes / 4;
    const unsigned blocks = (unsigned)((words + 255) / 256 < 64 ? (words + 255) / 256 : 64);
    copy_small_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<uint32_t*>(dst), static_cast<const uint32_t*>(src), words);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

// plain asynchronous copy on an explicit stream (host code: a small device->host
// or host->device copy without switching the framework's current stream)
extern "C" int oa_copy_async(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes == 0) return OA_OK;
    OA_REQUIRE(dst && src, "oa_copy_async: NULL pointer");
    OA_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault,
                                  static_cast<cudaStream_t>(stream)));
    return OA_OK;
}

extern "C" size_t oa_track_args_size(void) { return sizeof(oa_track_args); }
extern "C" size_t oa_synth_params_size(void) { return sizeof(oa_synth_params); }

// ---- host-side assembly of the region table (see include/orbit_b200.h) -------------
extern "C" int oa_host_copy(void* dst, const void* src, size_t bytes, int n_threads) {
    OA_REQUIRE(bytes == 0 || (dst && src), "oa_host_copy: null pointer");
    // (a thread costs tens of microseconds to start: at least 8 MiB of copy each)
    const size_t min_part = (size_t)8 << 20;
    size_t parts = n_threads > 1 ? (size_t)n_threads : 1;
    if (parts > 64) parts = 64;
    if (parts > bytes / min_part) parts = bytes / min_part;
    if (parts <= 1) {
        if (bytes) memcpy(dst, src, bytes);
        return OA_OK;
    }
    // part boundaries on 4 KiB multiples of the byte offset (the last part takes
    // the remainder)
    const size_t step = ((bytes / parts) + 4095) & ~(size_t)4095;
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    std::vector<std::thread> pool;
    try {
        for (size_t t = 1; t < parts; ++t) {
            const size_t lo = t * step;
            if (lo >= bytes) break;
            const size_t hi = (t + 1 == parts || lo + step > bytes) ? bytes : lo + step;
            pool.emplace_back([=] { memcpy(d + lo, s + lo, hi - lo); });
        }
    } catch (...) {
        // (thread creation failed: the parts not handed out are copied below)
        for (auto& th : pool) th.join();
        memcpy(dst, src, bytes);
        return OA_OK;
    }
    memcpy(d, s, step < bytes ? step : bytes);
    for (auto& th : pool) th.join();
    return OA_OK;
}

extern "C" int oa_region_rows_host(int n_regions, const int64_t* offsets,
                                   const void* centres, int centre_dtype,
                                   const void* bulk, int bulk_dtype,
                                   const int64_t* halo_ids, const int64_t* prev_halo_ids,
                                   int n_prev_regions, const int64_t* prev_offsets,
                                   const int64_t* prev_buckets, oa_region* rows,
                                   int64_t* buckets_out, uint8_t* matched_out,
                                   int32_t* prev_index_out, int64_t* seg_begin_out,
                                   int* n_matched) {
    OA_REQUIRE(n_regions >= 0 && n_matched, "oa_region_rows_host: bad arguments");
    *n_matched = 0;
    if (n_regions == 0) return OA_OK;
    OA_REQUIRE(offsets && centres && halo_ids && rows && buckets_out && matched_out &&
               prev_index_out && seg_begin_out, "oa_region_rows_host: NULL pointer");
    OA_REQUIRE((centre_dtype == OA_F32 || centre_dtype == OA_F64) &&
               (!bulk || bulk_dtype == OA_F32 || bulk_dtype == OA_F64),
               "oa_region_rows_host: bad dtype");
    OA_REQUIRE(n_prev_regions == 0 || (prev_halo_ids && prev_offsets && prev_buckets),
               "oa_region_rows_host: previous generation incomplete");
    // pass 1 (sequential, cheap): position of every halo in the previous
    // (ascending) list.  Catalogues keep their order from snapshot to snapshot:
    // the entry after the last hit is tried first (a binary search per halo is 17
    // cache misses at 100 k halos: 6 ms per snapshot)
    int m = 0, hint = 0;
    for (int j = 0; j < n_regions; ++j) {
        const int64_t id = halo_ids[j];
        int lo = hint;
        if (!(lo < n_prev_regions && prev_halo_ids[lo] == id)) {
            int hi = n_prev_regions;
            lo = 0;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (prev_halo_ids[mid] < id) lo = mid + 1; else hi = mid;
            }
        }
        const bool hit = lo < n_prev_regions && prev_halo_ids[lo] == id;
        if (hit) hint = lo + 1;
        matched_out[j] = hit ? 1 : 0;
        prev_index_out[j] = hit ? lo : -1;
        if (hit) seg_begin_out[m++] = prev_offsets[lo];
    }
    // pass 2: the 128-byte rows (12.8 MB at 100 k halos): independent, so large
    // catalogues are filled by a few threads
    auto fill = [&](int j0, int j1) {
        for (int j = j0; j < j1; ++j) {
            oa_region& r = rows[j];
            memset(&r, 0, sizeof(r));
            for (int q = 0; q < 3; ++q) {
                if (centre_dtype == OA_F32) {
                    const float c = static_cast<const float*>(centres)[3 * j + q];
                    r.centre[q] = (double)c;
                    r.centre_f[q] = c;
                } else {
                    const double c = static_cast<const double*>(centres)[3 * j + q];
                    r.centre[q] = c;
                    r.centre_f[q] = (float)c;
                }
                if (bulk) {
                    if (bulk_dtype == OA_F32) {
                        const float b = static_cast<const float*>(bulk)[3 * j + q];
                        r.bulk[q] = (double)b;
                        r.bulk_f[q] = b;
                    } else {
                        const double b = static_cast<const double*>(bulk)[3 * j + q];
                        r.bulk[q] = b;
                        r.bulk_f[q] = (float)b;
                    }
                }
            }
            r.cur_begin = offsets[j];
            r.cur_count = offsets[j + 1] - offsets[j];
            r.cur_bucket = oa_table_bucket_begin(offsets[j], j);
            buckets_out[j] = r.cur_bucket;
            r.prev_begin = -1;
            const int lo = prev_index_out[j];
            if (lo >= 0) {
                r.prev_begin = prev_offsets[lo];
                r.prev_count = prev_offsets[lo + 1] - prev_offsets[lo];
                r.prev_bucket = prev_buckets[lo];
            }
        }
    };
    unsigned hw = std::thread::hardware_concurrency();
    const int n_thr = (n_regions >= 16384 && hw >= 2) ? (int)(hw < 4 ? hw : 4) : 1;
    if (n_thr == 1) {
        fill(0, n_regions);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_thr; ++t) {
            const int j0 = (int)((int64_t)n_regions * t / n_thr);
            const int j1 = (int)((int64_t)n_regions * (t + 1) / n_thr);
            pool.emplace_back(fill, j0, j1);
        }
        for (auto& th : pool) th.join();
    }
    *n_matched = m;
    return OA_OK;
}
