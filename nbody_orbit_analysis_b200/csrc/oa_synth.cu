// On-device generator of the synthetic benchmark workload.
//
// Not part of the reference's path: it produces the INPUTS of the throughput
// runs directly in HBM (rosette orbits about drifting halo centres in a
// periodic box, region membership churn, per-snapshot block shuffle) so that
// bench.py does not spend minutes generating 256^3 particles with numpy.
// The block shuffle is a radix sort (oa_sort.cu) on (halo, hash) keys.
#include "oa_common.cuh"
#include <math.h>

namespace {

OA_D uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

OA_D double uniform(uint64_t seed, uint64_t stream, uint64_t idx) {
    const uint64_t k = splitmix64(seed * 0x100000001B3ull + stream);
    const uint64_t z = splitmix64(idx ^ k);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

OA_D int find_halo(const int64_t* __restrict__ start, int n_halos, int64_t u) {
    int lo = 0, hi = n_halos - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(start + mid) <= u) lo = mid; else hi = mid - 1;
    }
    return lo;
}

struct Radial {
    double a, e, w, phi;
};

OA_D Radial radial_elements(uint64_t seed, uint64_t pid, double R) {
    Radial q;
    q.a = (0.05 + 1.15 * uniform(seed, 1, pid)) * R;
    q.e = 0.1 + 0.7 * uniform(seed, 2, pid);
    q.w = 0.1 + 0.9 * uniform(seed, 3, pid);
    q.phi = 6.283185307179586 * uniform(seed, 4, pid);
    return q;
}

__global__ void synth_keys_kernel(oa_synth_params p, uint64_t* __restrict__ keys,
                                  uint64_t* __restrict__ vals,
                                  unsigned long long* __restrict__ halo_counts) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= p.n_universe) return;
    const int h = find_halo(p.halo_start, p.n_halos, u);
    const uint64_t pid = (uint64_t)(u * p.id_stride + p.id_offset);
    const double R = __ldg(p.halo_radius + h);
    const Radial q = radial_elements(p.seed, pid, R);
    const double r = q.a * (1.0 - q.e * cos(q.w * p.t + q.phi));
    uint64_t key;
    if (r <= R) {
        const uint64_t salt = splitmix64(p.seed * 7919ull +
                                         104729ull * (uint64_t)(p.t + 1.0));
        key = ((uint64_t)h << 40) | (splitmix64(pid ^ salt) >> 24);
        atomicAdd(halo_counts + h, 1ull);
    } else {
        key = (uint64_t)p.n_halos << 40;     // sorts behind every real block
    }
    keys[u] = key;
    vals[u] = (uint64_t)u;
}

template <typename T>
__global__ void synth_fill_kernel(oa_synth_params p, const uint64_t* __restrict__ order,
                                  int64_t n_present, T* __restrict__ pos,
                                  T* __restrict__ vel, int64_t* __restrict__ ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_present) return;
    const int64_t u = (int64_t)order[i];
    const int h = find_halo(p.halo_start, p.n_halos, u);
    const uint64_t pid = (uint64_t)(u * p.id_stride + p.id_offset);
    const double R = __ldg(p.halo_radius + h);
    const Radial q = radial_elements(p.seed, pid, R);
    const double psi0 = 6.283185307179586 * uniform(p.seed, 5, pid);
    const double kappa = 0.55 + 0.4 * uniform(p.seed, 6, pid);
    const double cz = 2.0 * uniform(p.seed, 7, pid) - 1.0;
    const double az = 6.283185307179586 * uniform(p.seed, 8, pid);
    const double sz = sqrt(fmax(1.0 - cz * cz, 0.0));
    const double n[3] = {sz * cos(az), sz * sin(az), cz};
    double ref[3] = {0.0, 0.0, 1.0};
    if (!(fabs(n[2]) < 0.9)) { ref[0] = 1.0; ref[2] = 0.0; }
    double e1[3] = {n[1] * ref[2] - n[2] * ref[1], n[2] * ref[0] - n[0] * ref[2],
                    n[0] * ref[1] - n[1] * ref[0]};
    const double inv = 1.0 / sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    e1[0] *= inv; e1[1] *= inv; e1[2] *= inv;
    const double e2[3] = {n[1] * e1[2] - n[2] * e1[1], n[2] * e1[0] - n[0] * e1[2],
                          n[0] * e1[1] - n[1] * e1[0]};
    const double ph = q.w * p.t + q.phi;
    const double r = q.a * (1.0 - q.e * cos(ph));
    const double rdot = q.a * q.e * q.w * sin(ph);
    const double psi = kappa * q.w * p.t + psi0;
    const double cp = cos(psi), sp = sin(psi);
    const double rt = r * kappa * q.w;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double er = cp * e1[k] + sp * e2[k];
        const double et = -sp * e1[k] + cp * e2[k];
        const double vh = __ldg(p.halo_vh + 3 * h + k);
        double c = __ldg(p.halo_c0 + 3 * h + k) + vh * p.t;
        double x = c + r * er;
        if (p.periodic) {
            x = fmod(x, p.box);
            if (x < 0.0) x += p.box;
        }
        T xs = (T)x;
        if (p.periodic && (double)xs >= p.box) xs = (T)0;   // float rounding up to L
        pos[3 * i + k] = xs;
        vel[3 * i + k] = (T)(vh + rdot * er + rt * et);
    }
    ids[i] = (int64_t)pid;
}

}  // namespace

extern "C" int oa_synth_keys(const oa_synth_params* params, uint64_t* keys,
                             uint64_t* vals, int64_t* halo_counts, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OA_REQUIRE(params && keys && vals && halo_counts, "oa_synth_keys: NULL pointer");
    const oa_synth_params& p = *params;
    OA_REQUIRE(p.n_halos >= 1 && p.n_halos < (1 << 23), "oa_synth_keys: bad n_halos");
    OA_CUDA_CHECK(cudaMemsetAsync(halo_counts, 0, sizeof(int64_t) * p.n_halos, st));
    if (p.n_universe <= 0) return OA_OK;
    synth_keys_kernel<<<(unsigned)((p.n_universe + 255) / 256), 256, 0, st>>>(
        p, keys, vals, reinterpret_cast<unsigned long long*>(halo_counts));
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_synth_fill(const oa_synth_params* params, const uint64_t* order,
                             int64_t n_present, int data_dtype, void* pos, void* vel,
                             int64_t* ids, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OA_REQUIRE(params, "oa_synth_fill: NULL params");
    if (n_present <= 0) return OA_OK;
    OA_REQUIRE(order && pos && vel && ids, "oa_synth_fill: NULL pointer");
    const unsigned blocks = (unsigned)((n_present + 255) / 256);
    if (data_dtype == OA_F64)
        synth_fill_kernel<double><<<blocks, 256, 0, st>>>(
            *params, order, n_present, static_cast<double*>(pos),
            static_cast<double*>(vel), ids);
    else
        synth_fill_kernel<float><<<blocks, 256, 0, st>>>(
            *params, order, n_present, static_cast<float*>(pos),
            static_cast<float*>(vel), ids);
    OA_LAUNCH_CHECK();
    return OA_OK;
}
