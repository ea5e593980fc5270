// (b0) Derived bulk velocity per region block (sm_100a).
//
// Replaces `np.mean(vel[sl], axis=0)` / `np.sum(m[:,None]*vel)/np.sum(m)`
// (track_orbits.py:267-284, track_orbits_onthefly.py:96-110).  numpy adds the
// rows one after another in the input dtype; a parallel reduction cannot
// reproduce that rounding, so sums are accumulated in float64 in a FIXED order
// (deterministic run to run) and the documented deviation is the float32
// accumulation error of the reference itself.
//
// Two kernels, no atomics: (1) every (tile, region) intersection gets one
// partial sum, stored at index tile+region (unique because both indices are
// monotone along the particle axis); (2) one warp per region adds its partials.
#include "oa_common.cuh"

namespace {

constexpr int BULK_THREADS = 256;
constexpr int BULK_TILE = 4096;

OA_D int find_region(const int64_t* __restrict__ off, int lo, int hi, int64_t c) {
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(off + mid) <= c) lo = mid; else hi = mid - 1;
    }
    return lo;
}

struct Acc {
    double x, y, z, m;
};

OA_D Acc acc_add(Acc a, const Acc& b) {
    a.x += b.x; a.y += b.y; a.z += b.z; a.m += b.m;
    return a;
}

OA_D Acc warp_reduce(Acc a) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        Acc b;
        b.x = __shfl_down_sync(0xFFFFFFFFu, a.x, d);
        b.y = __shfl_down_sync(0xFFFFFFFFu, a.y, d);
        b.z = __shfl_down_sync(0xFFFFFFFFu, a.z, d);
        b.m = __shfl_down_sync(0xFFFFFFFFu, a.m, d);
        a = acc_add(a, b);
    }
    return a;
}

template <typename TV, typename TM>
OA_D Acc accumulate(const TV* __restrict__ vel, const TM* __restrict__ mass,
                    int64_t lo, int64_t hi, int lane, int stride) {
    Acc a = {0.0, 0.0, 0.0, 0.0};
    for (int64_t i = lo + lane; i < hi; i += stride) {
        const double m = mass ? (double)__ldg(mass + i) : 1.0;
        a.x += m * (double)__ldg(vel + 3 * i + 0);
        a.y += m * (double)__ldg(vel + 3 * i + 1);
        a.z += m * (double)__ldg(vel + 3 * i + 2);
        a.m += m;
    }
    return a;
}

template <typename TV, typename TM>
__global__ void __launch_bounds__(BULK_THREADS)
bulk_partial_kernel(const TV* __restrict__ vel, const TM* __restrict__ mass,
                    const int64_t* __restrict__ off, int n_regions, int64_t n,
                    double4* __restrict__ partial) {
    __shared__ int s_j[2];
    __shared__ Acc s_acc[BULK_THREADS / 32];
    const int64_t tile = blockIdx.x;
    const int64_t base = tile * BULK_TILE;
    const int64_t end = min(base + (int64_t)BULK_TILE, n);
    if (threadIdx.x == 0) s_j[0] = find_region(off, 0, n_regions - 1, base);
    if (threadIdx.x == 32) s_j[1] = find_region(off, 0, n_regions - 1, end - 1);
    __syncthreads();
    const int jlo = s_j[0], jhi = s_j[1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (jlo == jhi) {
        // the whole tile lies inside one block: block-wide reduction
        Acc a = accumulate(vel, mass, base, end, threadIdx.x, BULK_THREADS);
        a = warp_reduce(a);
        if (lane == 0) s_acc[warp] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            Acc t = s_acc[0];
            for (int w = 1; w < BULK_THREADS / 32; ++w) t = acc_add(t, s_acc[w]);
            partial[tile + jlo] = make_double4(t.x, t.y, t.z, t.m);
        }
    } else {
        // several blocks touch this tile: one warp per (tile, region) piece
        for (int j = jlo + warp; j <= jhi; j += BULK_THREADS / 32) {
            const int64_t lo = max(base, __ldg(off + j));
            const int64_t hi = min(end, __ldg(off + j + 1));
            if (hi <= lo) continue;
            Acc a = accumulate(vel, mass, lo, hi, lane, 32);
            a = warp_reduce(a);
            if (lane == 0) partial[tile + j] = make_double4(a.x, a.y, a.z, a.m);
        }
    }
}

__global__ void __launch_bounds__(128)
bulk_finalize_kernel(const double4* __restrict__ partial,
                     const int64_t* __restrict__ off, int n_regions, int round_f32,
                     oa_region* __restrict__ regions, double* __restrict__ bulk_out) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= n_regions) return;
    const int64_t lo = off[j], hi = off[j + 1];
    Acc a = {0.0, 0.0, 0.0, 0.0};
    if (hi > lo) {
        const int64_t t0 = lo / BULK_TILE, t1 = (hi - 1) / BULK_TILE;
        for (int64_t t = t0 + lane; t <= t1; t += 32) {
            const double4 p = partial[t + j];
            a.x += p.x; a.y += p.y; a.z += p.z; a.m += p.w;
        }
    }
    a = warp_reduce(a);
    if (lane == 0) {
        double b[3] = {a.x / a.m, a.y / a.m, a.z / a.m};   // 0/0 = NaN when empty
        for (int k = 0; k < 3; ++k) {
            if (round_f32) b[k] = (double)(float)b[k];
            regions[j].bulk[k] = b[k];
            regions[j].bulk_f[k] = (float)b[k];
            if (bulk_out) bulk_out[3 * j + k] = b[k];
        }
    }
}

inline int64_t bulk_tiles(int64_t n) { return (n + BULK_TILE - 1) / BULK_TILE; }

template <typename TV, typename TM>
int launch_bulk(const void* vel, const void* mass, const int64_t* off, int n_regions,
                int64_t n, int round_f32, oa_region* regions, double* bulk_out,
                void* workspace, cudaStream_t st) {
    double4* partial = static_cast<double4*>(workspace);
    if (n > 0) {
        bulk_partial_kernel<TV, TM><<<(unsigned)bulk_tiles(n), BULK_THREADS, 0, st>>>(
            static_cast<const TV*>(vel), static_cast<const TM*>(mass), off, n_regions,
            n, partial);
        OA_LAUNCH_CHECK();
    }
    const int warps_per_block = 4;
    bulk_finalize_kernel<<<(n_regions + warps_per_block - 1) / warps_per_block,
                           warps_per_block * 32, 0, st>>>(
        partial, off, n_regions, round_f32, regions, bulk_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

}  // namespace

extern "C" size_t oa_bulk_workspace_bytes(int64_t n, int n_regions) {
    return (size_t)(bulk_tiles(n > 0 ? n : 0) + n_regions + 1) * sizeof(double4);
}

extern "C" int oa_bulk_velocity(const void* vel, int vel_dtype, const void* mass,
                                int mass_dtype, const int64_t* cur_off,
                                int n_regions, int64_t n, int round_f32,
                                oa_region* regions, double* bulk_out,
                                void* workspace, size_t workspace_bytes,
                                void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_regions <= 0) return OA_OK;
    OA_REQUIRE(cur_off && regions && workspace && (vel || n == 0),
               "oa_bulk_velocity: NULL pointer");
    OA_REQUIRE(workspace_bytes >= oa_bulk_workspace_bytes(n, n_regions),
               "oa_bulk_velocity: workspace too small");
    OA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 31) == 0,
               "oa_bulk_velocity: workspace must be 32-byte aligned");
    const bool v64 = vel_dtype == OA_F64, m64 = mass_dtype == OA_F64;
    if (v64) {
        if (m64 || !mass)
            return launch_bulk<double, double>(vel, mass, cur_off, n_regions, n, round_f32,
                                               regions, bulk_out, workspace, st);
        return launch_bulk<double, float>(vel, mass, cur_off, n_regions, n, round_f32,
                                          regions, bulk_out, workspace, st);
    }
    if (m64 && mass)
        return launch_bulk<float, double>(vel, mass, cur_off, n_regions, n, round_f32,
                                          regions, bulk_out, workspace, st);
    return launch_bulk<float, float>(vel, mass, cur_off, n_regions, n, round_f32,
                                     regions, bulk_out, workspace, st);
}
