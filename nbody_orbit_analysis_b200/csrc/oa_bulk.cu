// (b0) Derived bulk velocity per region block (sm_100a), BIT-EXACT.
//
// Replaces `np.mean(vel[sl], axis=0)` / `np.sum(m[:,None]*vel, axis=0)/np.sum(m)`
// (track_orbits.py:267-284, track_orbits_onthefly.py:96-110).  The rounding of
// those sums is part of the result: a bulk velocity that differs in the last
// bits flips the sign of v_r for particles near their apsis and changes the
// event lists.  numpy's order is therefore reproduced exactly (checked against
// numpy 2.3 on the host, tests/test_abi_and_host.py restates both orders):
//
//   * a reduction over axis 0 of an (n, 3) array adds the ROWS ONE AFTER
//     ANOTHER in the array's dtype (the reduced axis is the outer loop: no
//     pairwise blocking) -- three dependent chains per region;
//   * `np.sum` of the contiguous 1-D mass slice is numpy's PAIRWISE sum: below 8
//     elements a plain loop from zero, up to 128 elements eight interleaved
//     accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a tail,
//     above that a split at n/2 rounded down to a multiple of 8;
//   * `np.mean` divides the float sum by float(n); the weighted form divides
//     by the mass sum cast to the product dtype.
//
// One warp per region: the lanes stage tiles of rows in shared memory
// (coalesced), lanes 0..2 run the three chains, a chain of n float adds costs
// ~5 cycles per row (2.5 ms for a block of 800 k particles; the path is only
// taken when the catalogue gives no bulk velocities).
#include "oa_common.cuh"

namespace {

constexpr int ROWS = 256;          // rows per staged tile
constexpr int WARPS = 4;           // regions per CTA

// numpy's pairwise sum (numpy/_core/src/umath/loops_utils.h.src), T arithmetic.
// Leaf: at most 128 elements.
template <typename T>
__host__ __device__ __forceinline__ T pairwise_leaf(const T* __restrict__ a, int64_t n) {
    if (n < 8) {
        T res = (T)0;
        for (int64_t i = 0; i < n; ++i) res = res + a[i];
        return res;
    }
    T r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = r[j] + a[i + j];
    }
    T res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res = res + a[i];
    return res;
}
// The recursion `sum(a, n) = sum(a, n2) + sum(a + n2, n - n2)`, n2 = n/2 rounded
// down to a multiple of 8, without a call stack (depth <= 25 for n < 2^31).
template <typename T>
__host__ __device__ T pairwise_sum(const T* __restrict__ a, int64_t n) {
    int64_t f_off[32], f_n[32];
    T f_left[32];
    unsigned char f_phase[32];       // 0: nothing done, 1: left half summed
    int top = 0;
    f_off[0] = 0; f_n[0] = n; f_phase[0] = 0;
    T ret = (T)0;
    while (top >= 0) {
        const int64_t o = f_off[top], m = f_n[top];
        if (m <= 128) {
            ret = pairwise_leaf(a + o, m);
            --top;
            continue;
        }
        int64_t n2 = m / 2;
        n2 -= n2 % 8;
        if (f_phase[top] == 0) {             // descend into the left half
            f_phase[top] = 1;
            ++top;
            f_off[top] = o; f_n[top] = n2; f_phase[top] = 0;
        } else if (f_phase[top] == 1) {      // left done (in ret): descend right
            f_left[top] = ret;
            f_phase[top] = 2;
            ++top;
            f_off[top] = o + n2; f_n[top] = m - n2; f_phase[top] = 0;
        } else {                             // both done
            ret = f_left[top] + ret;
            --top;
        }
    }
    return ret;
}

template <typename TV, typename TM, typename TP>
__global__ void __launch_bounds__(32 * WARPS)
bulk_exact_kernel(const TV* __restrict__ vel, const TM* __restrict__ mass,
                  const int64_t* __restrict__ off, int n_regions, int round_f32,
                  oa_region* __restrict__ regions, double* __restrict__ bulk_out) {
    __shared__ TV s_v[WARPS][3 * ROWS];
    __shared__ TM s_m[WARPS][ROWS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int j = blockIdx.x * WARPS + w;
    if (j >= n_regions) return;
    const int64_t lo = off[j], hi = off[j + 1], n = hi - lo;
    TP acc = (TP)0;
    for (int64_t base = lo; base < hi; base += ROWS) {
        const int rows = (int)((hi - base) < ROWS ? (hi - base) : ROWS);
        for (int e = lane; e < 3 * rows; e += 32) s_v[w][e] = __ldg(vel + 3 * base + e);
        if (mass)
            for (int e = lane; e < rows; e += 32) s_m[w][e] = __ldg(mass + base + e);
        __syncwarp();
        if (lane < 3) {
            for (int r = 0; r < rows; ++r) {
                // (no contraction: the build uses -fmad=false)
                const TP p = mass ? (TP)s_m[w][r] * (TP)s_v[w][3 * r + lane]
                                  : (TP)s_v[w][3 * r + lane];
                acc = (base == lo && r == 0) ? p : acc + p;
            }
        }
        __syncwarp();
    }
    TP denom = (TP)(double)n;              // np.mean: sum / float(n)
    if (mass) {
        TM d = (TM)0;
        if (lane == 3) d = pairwise_sum(mass + lo, n);
        d = __shfl_sync(0xFFFFFFFFu, d, 3);
        denom = (TP)d;
    }
    if (lane < 3) {
        double b = (double)(acc / denom);  // 0/0 = NaN for an empty block
        if (round_f32) b = (double)(float)b;
        regions[j].bulk[lane] = b;
        regions[j].bulk_f[lane] = (float)b;
        if (bulk_out) bulk_out[3 * j + lane] = b;
    }
}

template <typename TV, typename TM, typename TP>
int launch_bulk(const void* vel, const void* mass, const int64_t* off, int n_regions,
                int round_f32, oa_region* regions, double* bulk_out, cudaStream_t st) {
    bulk_exact_kernel<TV, TM, TP><<<(n_regions + WARPS - 1) / WARPS, 32 * WARPS, 0, st>>>(
        static_cast<const TV*>(vel), static_cast<const TM*>(mass), off, n_regions, round_f32,
        regions, bulk_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

}  // namespace

// host twin of the device routine (same source): lets the CPU tests hold the
// emulation of numpy's pairwise sum against numpy itself
extern "C" int oa_pairwise_sum_host(const void* a, int dtype, int64_t n, double* out) {
    OA_REQUIRE(out && (a || n == 0) && n >= 0 && (dtype == OA_F32 || dtype == OA_F64),
               "oa_pairwise_sum_host: bad arguments");
    if (dtype == OA_F32) *out = (double)pairwise_sum(static_cast<const float*>(a), n);
    else *out = pairwise_sum(static_cast<const double*>(a), n);
    return OA_OK;
}

// (kept for the ABI: the exact kernel needs no workspace)
extern "C" size_t oa_bulk_workspace_bytes(int64_t n, int n_regions) {
    (void)n;
    (void)n_regions;
    return 256;
}

extern "C" int oa_bulk_velocity(const void* vel, int vel_dtype, const void* mass,
                                int mass_dtype, const int64_t* cur_off,
                                int n_regions, int64_t n, int round_f32,
                                oa_region* regions, double* bulk_out,
                                void* workspace, size_t workspace_bytes,
                                void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    (void)workspace;
    (void)workspace_bytes;
    if (n_regions <= 0) return OA_OK;
    OA_REQUIRE(cur_off && regions && (vel || n == 0), "oa_bulk_velocity: NULL pointer");
    const bool v64 = vel_dtype == OA_F64, m64 = mass_dtype == OA_F64;
    // product dtype = numpy's result_type(mass, velocity)
    if (!mass) {
        if (v64) return launch_bulk<double, double, double>(vel, nullptr, cur_off, n_regions,
                                                            round_f32, regions, bulk_out, st);
        return launch_bulk<float, float, float>(vel, nullptr, cur_off, n_regions, round_f32,
                                                regions, bulk_out, st);
    }
    if (v64) {
        if (m64) return launch_bulk<double, double, double>(vel, mass, cur_off, n_regions,
                                                            round_f32, regions, bulk_out, st);
        return launch_bulk<double, float, double>(vel, mass, cur_off, n_regions, round_f32,
                                                  regions, bulk_out, st);
    }
    if (m64) return launch_bulk<float, double, double>(vel, mass, cur_off, n_regions, round_f32,
                                                       regions, bulk_out, st);
    return launch_bulk<float, float, float>(vel, mass, cur_off, n_regions, round_f32, regions,
                                            bulk_out, st);
}
