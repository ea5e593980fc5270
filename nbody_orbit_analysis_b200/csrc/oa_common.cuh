// Shared declarations for liborbit_b200 (sm_100a only).
//
// Everything in csrc/ is written for one target, NVIDIA B200 (sm_100a); there
// is no multi-arch dispatch and no CPU fallback.  The C-ABI surface is declared
// in include/orbit_b200.h; this header holds what the kernels share.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/orbit_b200.h"

#define OA_HD __host__ __device__ __forceinline__
#define OA_D __device__ __forceinline__

// ---- error reporting --------------------------------------------------------
void oa_set_error(const char* fmt, ...);

#define OA_CUDA_CHECK(expr)                                                    \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) {                                               \
            oa_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,         \
                         cudaGetErrorString(_e));                              \
            return OA_ERR_CUDA;                                                \
        }                                                                      \
    } while (0)

#define OA_REQUIRE(cond, ...)                                                  \
    do {                                                                       \
        if (!(cond)) {                                                         \
            oa_set_error(__VA_ARGS__);                                         \
            return OA_ERR_ARG;                                                 \
        }                                                                      \
    } while (0)

#define OA_LAUNCH_CHECK() OA_CUDA_CHECK(cudaGetLastError())

// B200: 148 SMs.  Grids of persistent-style kernels are sized in multiples of it.
constexpr int OA_NUM_SMS = 148;

// ---- hashing of particle IDs --------------------------------------------------
// One 64-bit mix; the high word picks the slot, the low word supplies the
// fingerprint bits stored next to the index in the slot.
OA_HD uint64_t oa_mix64(uint64_t x) {
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}

// slot of a hash in a table segment of `cap` slots (Lemire range reduction)
OA_HD uint32_t oa_slot(uint32_t h, uint32_t cap) {
    return (uint32_t)(((uint64_t)h * (uint64_t)cap) >> 32);
}

// ---- record layouts (carried state, one per region-particle) -------------------
// 32 B (float frame) / 64 B (double frame): a probe hit costs ONE 32 B sector
// for the float record -- id (verification), unit vector, v_r, radius, angle.
template <typename T> struct OaRec;

template <> struct __align__(16) OaRec<float> {
    int64_t id;
    float rx, ry, rz;
    float vr;        // sign-faithful float copy of v_r (see oa_sign_faithful)
    float r;
    __half angle;    // swept-angle accumulator, re-rounded every snapshot
    uint16_t flags;
};
static_assert(sizeof(OaRec<float>) == 32, "float record must be one sector");

template <> struct __align__(16) OaRec<double> {
    int64_t id;
    double rx, ry, rz;
    double vr;
    double r;
    __half angle;
    uint16_t flags;
    uint32_t pad;
};
static_assert(sizeof(OaRec<double>) == 64, "double record must be two sectors");

// event mark meaning "no event" in the per-previous-particle mark array.  A
// float16 angle is never -0.0 (angles are >= +0 or NaN), so the pattern is free.
constexpr uint16_t OA_NO_EVENT = 0x8000u;

// ---- block-level scans (device helpers shared by select / sort) ---------------
OA_D uint32_t oa_warp_inclusive_scan(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive prefix of `v` over a block of THREADS threads; *total = block sum.
// Contains __syncthreads(): every thread of the block must call it.
template <int THREADS>
OA_D uint32_t oa_block_exclusive_scan(uint32_t v, uint32_t* total) {
    constexpr int WARPS = THREADS / 32;
    static_assert(WARPS >= 1 && WARPS <= 32, "block size");
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = oa_warp_inclusive_scan(v);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < WARPS ? s_warp[lane] : 0u;
        const uint32_t winc = oa_warp_inclusive_scan(w);
        s_warp[lane] = winc - w;
        if (lane == 31) s_total = winc;
    }
    __syncthreads();
    const uint32_t out = inc - v + s_warp[warp];
    *total = s_total;
    __syncthreads();
    return out;
}
