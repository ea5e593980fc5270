// Region extraction: the loader step in front of the tracking path (reference
// example_script.py:36-67 -- for every halo, `np.argwhere(|recenter(x - c)| < R)`
// over ALL particles, O(N x n_halo)).  Here: a uniform grid over the box holds,
// per cell, the regions whose sphere touches the cell (built on the host, the
// catalogue is small); one thread per particle tests only the regions of its
// cell with exactly the reference's arithmetic (same min-image code as the
// tracking kernels, utils.py:24-33) and emits a key `region << 32 | particle`.
// Sorting the keys gives the reference's layout: regions in catalogue order,
// particle indices ascending inside a region.
#include "oa_common.cuh"

namespace {

struct GridParams {
    double lo[3];        // lower corner of the grid
    double inv_cell[3];  // cells per unit length
    double box[3];
    int dim[3];          // cells per axis
    int periodic;
};

template <typename TX>
__device__ __forceinline__ int cell_of(const GridParams& g, const TX* x, int* c) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double u = ((double)x[q] - g.lo[q]) * g.inv_cell[q];
        int k = (int)floor(u);
        if (g.periodic) {
            k %= g.dim[q];
            if (k < 0) k += g.dim[q];
        } else if (k < 0 || k >= g.dim[q]) {
            return -1;
        }
        c[q] = k;
    }
    return (c[2] * g.dim[1] + c[1]) * g.dim[0] + c[0];
}

// r < R with numpy's arithmetic: d = x - c in the promoted dtype T, single +-L
// wrap with strict comparisons (computed in float64, stored in T), r = sqrt of
// the 3-term einsum in T, comparison against the float64 radius.
template <typename TX, typename T>
__device__ __forceinline__ bool inside(const GridParams& g, const TX* x, const double* c64,
                                       const float* c32, double radius) {
    T d[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (sizeof(T) == 4) d[q] = (T)__fsub_rn((float)x[q], c32[q]);
        else d[q] = (T)__dsub_rn((double)x[q], c64[q]);
        if (g.periodic) {
            const double h = g.box[q] * 0.5, dd = (double)d[q];
            if (dd > h) d[q] = (T)__dsub_rn(dd, g.box[q]);
            if ((double)d[q] < -h) d[q] = (T)__dadd_rn((double)d[q], g.box[q]);
        }
    }
    double r;
    if (sizeof(T) == 4) {
        const float a = __fmul_rn((float)d[0], (float)d[0]), b = __fmul_rn((float)d[1], (float)d[1]),
                    e = __fmul_rn((float)d[2], (float)d[2]);
        r = (double)__fsqrt_rn(__fadd_rn(__fadd_rn(a, b), e));
    } else {
        const double a = __dmul_rn((double)d[0], (double)d[0]),
                     b = __dmul_rn((double)d[1], (double)d[1]),
                     e = __dmul_rn((double)d[2], (double)d[2]);
        r = __dsqrt_rn(__dadd_rn(__dadd_rn(a, e), b));
    }
    return r < radius;
}

template <typename TX, typename T>
__global__ void region_pairs_kernel(const TX* __restrict__ pos, int64_t n,
                                    const double* __restrict__ centres,
                                    const float* __restrict__ centres_f,
                                    const double* __restrict__ radii,
                                    const int32_t* __restrict__ cell_start,
                                    const int32_t* __restrict__ cell_regions,
                                    const __grid_constant__ GridParams g,
                                    uint64_t* __restrict__ keys, int64_t capacity,
                                    unsigned long long* __restrict__ counter) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int hits = 0;
    uint64_t mine[4];                // most particles sit in 0-2 regions
    if (i < n) {
        const TX x[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
        int c[3];
        const int cell = cell_of(g, x, c);
        if (cell >= 0) {
            for (int e = cell_start[cell]; e < cell_start[cell + 1]; ++e) {
                const int j = cell_regions[e];
                if (inside<TX, T>(g, x, centres + 3 * j, centres_f + 3 * j, radii[j])) {
                    const uint64_t key = ((uint64_t)(uint32_t)j << 32) | (uint64_t)(uint32_t)i;
                    if (hits < 4) {
                        mine[hits] = key;
                    } else if (keys) {          // rare: flush directly
                        const unsigned long long at = atomicAdd(counter, 1ull);
                        if ((int64_t)at < capacity) keys[at] = key;
                    } else {
                        atomicAdd(counter, 1ull);
                    }
                    ++hits;
                }
            }
        }
    }
    const int first = hits < 4 ? hits : 4;
    // one atomic per warp for the common case
    const unsigned lane = threadIdx.x & 31u;
    unsigned total = (unsigned)first, excl = 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xFFFFFFFFu, total, d);
        if (lane >= (unsigned)d) total += t;
    }
    excl = total - (unsigned)first;
    const unsigned warp_total = __shfl_sync(0xFFFFFFFFu, total, 31);
    unsigned long long base = 0;
    if (lane == 31 && warp_total) base = atomicAdd(counter, (unsigned long long)warp_total);
    base = __shfl_sync(0xFFFFFFFFu, base, 31);
    if (keys) {
        for (int h = 0; h < first; ++h) {
            const int64_t at = (int64_t)(base + excl) + h;
            if (at < capacity) keys[at] = mine[h];
        }
    }
}

template <typename TX>
__global__ void gather_rows3_kernel(const TX* __restrict__ src, const uint64_t* __restrict__ keys,
                                    int64_t n, TX* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    const int64_t k = t / 3;
    const int q = (int)(t - 3 * k);
    out[t] = src[3 * (int64_t)(uint32_t)keys[k] + q];
}

template <typename TV>
__global__ void gather_low32_kernel(const TV* __restrict__ src, const uint64_t* __restrict__ keys,
                                    int64_t n, TV* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = src[(uint32_t)keys[k]];
}

inline unsigned blocks(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

// keys == NULL: count only.  *counter (device, zeroed by the call) receives the
// number of (region, particle) pairs.
extern "C" int oa_region_pairs(const void* pos, int data_dtype, int64_t n, const double* centres,
                               const float* centres_f, const double* radii, int frame_dtype,
                               const int32_t* cell_start, const int32_t* cell_regions,
                               const double* grid_lo, const double* grid_inv_cell,
                               const int32_t* grid_dim, const double* box, int periodic,
                               uint64_t* keys, int64_t capacity, uint64_t* counter, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OA_REQUIRE(n >= 0 && n < ((int64_t)1 << 32), "oa_region_pairs: particle index must fit 32 bits");
    OA_REQUIRE(counter && grid_lo && grid_inv_cell && grid_dim, "oa_region_pairs: NULL pointer");
    OA_REQUIRE((data_dtype == OA_F32 || data_dtype == OA_F64) &&
               (frame_dtype == OA_F32 || frame_dtype == OA_F64) &&
               !(data_dtype == OA_F64 && frame_dtype == OA_F32),
               "oa_region_pairs: bad dtype combination");
    OA_REQUIRE(!periodic || box, "oa_region_pairs: periodic without a box");
    OA_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(uint64_t), st));
    if (n == 0) return OA_OK;
    OA_REQUIRE(pos && centres && centres_f && radii && cell_start && cell_regions,
               "oa_region_pairs: NULL array");
    GridParams g;
    for (int q = 0; q < 3; ++q) {
        g.lo[q] = grid_lo[q];
        g.inv_cell[q] = grid_inv_cell[q];
        g.dim[q] = grid_dim[q];
        g.box[q] = box ? box[q] : 0.0;
    }
    g.periodic = periodic;
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(counter);
    const unsigned nb = blocks(n, 256);
    if (data_dtype == OA_F64)
        region_pairs_kernel<double, double><<<nb, 256, 0, st>>>(
            static_cast<const double*>(pos), n, centres, centres_f, radii, cell_start,
            cell_regions, g, keys, capacity, cnt);
    else if (frame_dtype == OA_F64)
        region_pairs_kernel<float, double><<<nb, 256, 0, st>>>(
            static_cast<const float*>(pos), n, centres, centres_f, radii, cell_start,
            cell_regions, g, keys, capacity, cnt);
    else
        region_pairs_kernel<float, float><<<nb, 256, 0, st>>>(
            static_cast<const float*>(pos), n, centres, centres_f, radii, cell_start,
            cell_regions, g, keys, capacity, cnt);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

// out[k] = src[low 32 bits of keys[k]]: rows of 3 (`rows3` != 0) or scalars of
// `elem_bytes` in {4, 8}
extern "C" int oa_gather_by_key(const void* src, int elem_bytes, int rows3, const uint64_t* keys,
                                int64_t n, void* out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(src && keys && out && (elem_bytes == 4 || elem_bytes == 8),
               "oa_gather_by_key: bad arguments");
    if (rows3) {
        if (elem_bytes == 4)
            gather_rows3_kernel<float><<<blocks(3 * n, 256), 256, 0, st>>>(
                static_cast<const float*>(src), keys, n, static_cast<float*>(out));
        else
            gather_rows3_kernel<double><<<blocks(3 * n, 256), 256, 0, st>>>(
                static_cast<const double*>(src), keys, n, static_cast<double*>(out));
    } else {
        if (elem_bytes == 4)
            gather_low32_kernel<uint32_t><<<blocks(n, 256), 256, 0, st>>>(
                static_cast<const uint32_t*>(src), keys, n, static_cast<uint32_t*>(out));
        else
            gather_low32_kernel<uint64_t><<<blocks(n, 256), 256, 0, st>>>(
                static_cast<const uint64_t*>(src), keys, n, static_cast<uint64_t*>(out));
    }
    OA_LAUNCH_CHECK();
    return OA_OK;
}
