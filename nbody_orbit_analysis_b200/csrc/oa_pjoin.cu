// Device build of the partitioned hash join (csrc/oa_pjoin_core.cuh): execution
// context on a CUDA thread block, the persistent kernel and its C-ABI launcher
// (include/orbit_b200.h: oa_pjoin_step).
#include "oa_common.cuh"
#include "oa_pjoin_core.cuh"

#ifndef OA_PJOIN_STATS
#define OA_PJOIN_STATS 0
#endif
// profiling builds: [0..3] cycles per stage (thread 0 of every CTA), [4] cycles
// waiting for dependencies, [8..11] items per stage, [12] CTA cycles in the kernel
__device__ unsigned long long g_pj_stats[16];

namespace {

struct DevCtx {
#if OA_PJOIN_STATS
    OA_D uint64_t clock() const { return (uint64_t)clock64(); }
    OA_D void stat_add(int i, uint64_t v) const {
        atomicAdd(&g_pj_stats[i], (unsigned long long)v);
    }
#else
    OA_D uint64_t clock() const { return 0; }
    OA_D void stat_add(int, uint64_t) const {}
#endif
    unsigned char* sm;
    uint32_t parity;        // phase of the TMA mbarrier (same in every thread)
    OA_D int tid() const { return (int)threadIdx.x; }
    OA_D void sync() const { __syncthreads(); }
    OA_D unsigned char* smem() const { return sm; }
    OA_D uint32_t atomic_add(uint32_t* p, uint32_t v) const { return atomicAdd(p, v); }
    OA_D uint32_t atomic_cas(uint32_t* p, uint32_t c, uint32_t v) const {
        return atomicCAS(p, c, v);
    }
    OA_D uint32_t load_acquire(const uint32_t* p) const {
        uint32_t v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
    }
    OA_D void release_add(uint32_t* p, uint32_t v) const {
        __threadfence();
        atomicAdd(p, v);
    }
    OA_D void backoff() const { __nanosleep(100); }
    OA_D void fail() const { __trap(); }
    // a value another CTA of this launch may have written: L2, never L1
    OA_D uint32_t ld_cg(const uint32_t* p) const { return __ldcg(p); }
    // whole records as one 256-bit access (sm_100): a warp moves 1 KB per instruction
    OA_D pj::Rec load_rec_cg(const pj::Rec* p) const {
        pj::Rec r;
        uint32_t* w = reinterpret_cast<uint32_t*>(&r);
        asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]),
                       "=r"(w[6]), "=r"(w[7])
                     : "l"(p) : "memory");
        return r;
    }
    OA_D void store_rec(pj::Rec* p, const pj::Rec& r) const {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&r);
        asm volatile("st.global.cg.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]),
                        "r"(w[6]), "r"(w[7])
                     : "memory");
    }
    // global -> shared bulk copy, complete for every thread on return.  Called by
    // all threads of the CTA after a barrier (the destination is not in use).
    OA_D void bulk_load(void* dst, const void* src, uint32_t bytes) {
        uint64_t* bar = reinterpret_cast<uint64_t*>(sm + pj::SM_MBAR);
        const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
        if (bytes == 0) return;
        if (threadIdx.x == 0) {
            // earlier generic-proxy accesses to `dst` are ordered before the
            // async-proxy writes of the copy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                         :: "r"(bar_a), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
                         "[%0], [%1], %2, [%3];"
                         :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes),
                            "r"(bar_a) : "memory");
        }
        uint32_t ok;
        do {
            asm volatile("{\n .reg .pred p;\n"
                         " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                         " selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(bar_a), "r"(parity) : "memory");
        } while (!ok);
        parity ^= 1u;
    }
    OA_D int64_t ld_last(const int64_t* p) const { return __ldcs(p); }
    OA_D float ld_last(const float* p) const { return __ldcs(p); }
    // read-only for this launch and touched once: streaming, no L1 allocation
    OA_D pj::U4 ld_stream(const pj::U4* p) const {
        const uint4 v = __ldcs(reinterpret_cast<const uint4*>(p));
        pj::U4 r;
        r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
        return r;
    }
};

__global__ void __launch_bounds__(pj::THREADS, OA_PJOIN_MIN_CTAS)
oa_pjoin_kernel(const __grid_constant__ oa_pjoin_args a, const __grid_constant__ pj::Const k,
                const __grid_constant__ pj::Work w) {
    extern __shared__ __align__(128) unsigned char smem[];
    DevCtx cx;
    cx.sm = smem;
    cx.parity = 0;
#if OA_PJOIN_TMA
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(smem + pj::SM_MBAR)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
#endif
    const uint64_t t0 = cx.clock();
    pj::run(cx, a, k, w);
    if (threadIdx.x == 0) cx.stat_add(12, cx.clock() - t0);
}

// one block per region writes the region's work items (see pj::expand_region)
__global__ void oa_pjoin_expand_kernel(const __grid_constant__ oa_pjoin_args a,
                                       uint64_t* __restrict__ items) {
    pj::expand_region(a, items, (int)blockIdx.x, (int)threadIdx.x, (int)blockDim.x);
}

// workspace: [ items (8 B each) | ticket (4 words) | done x 3 per region | cursors ]
size_t work_words(int n_regions, int64_t n_part_entries) {
    return 4 + 3 * (size_t)n_regions + (size_t)n_part_entries;
}

}  // namespace

extern "C" size_t oa_pjoin_workspace_bytes(int n_regions, int64_t n_part_entries,
                                           uint32_t total_tickets) {
    return 8 * (size_t)total_tickets + 4 * work_words(n_regions, n_part_entries);
}

// profiling builds (-DOA_PJOIN_STATS=1): read (and optionally clear) the counters;
// returns 0 when the library was built without them
extern "C" int oa_pjoin_stats(uint64_t* out16, int reset) {
    if (!out16) return 0;
    unsigned long long h[16];
    if (cudaMemcpyFromSymbol(h, g_pj_stats, sizeof(h)) != cudaSuccess) return 0;
    for (int i = 0; i < 16; ++i) out16[i] = h[i];
    if (reset) {
        for (int i = 0; i < 16; ++i) h[i] = 0;
        if (cudaMemcpyToSymbol(g_pj_stats, h, sizeof(h)) != cudaSuccess) return 0;
    }
    return OA_PJOIN_STATS;
}

extern "C" size_t oa_pjoin_args_size(void) { return sizeof(oa_pjoin_args); }

extern "C" void oa_pjoin_config(int32_t* out8) {
    out8[0] = pj::THREADS;
    out8[1] = OA_PJOIN_MIN_CTAS;
    out8[2] = pj::TILE;
    out8[3] = pj::CTILE;
    out8[4] = pj::REC_CAP;
    out8[5] = OA_PJOIN_TARGET;
    out8[6] = pj::MAX_BITS;
    out8[7] = pj::SM_BYTES;
}

// host-side plan (restated in numpy in pjoin.py:make_plan, which the tests
// compare with this function)
extern "C" int oa_pjoin_plan_host(const int64_t* offsets, int n_regions,
                                  const int32_t* prev_bits, const int64_t* prev_pb,
                                  const int64_t* prev_counts,
                                  int64_t target, int64_t lag_particles,
                                  oa_pjoin_region* rows, int32_t* bits_out, int64_t* pb_out,
                                  uint32_t* group_first, uint32_t* range_start,
                                  oa_pjoin_plan_info* info) {
    OA_REQUIRE(n_regions >= 0 && target >= 1 && lag_particles >= 1 && rows && group_first &&
               range_start && info && (n_regions == 0 || (offsets && prev_bits && prev_pb &&
                                                          prev_counts && bits_out && pb_out)),
               "oa_pjoin_plan_host: bad arguments");
    uint64_t pb = 0, tiles = 0, ctiles = 0, joins = 0, scans = 0;
    int n_groups = 0, max_bits = 0;
    int64_t gid_prev = -1;
    int leader = -1;                       // first region of the open pack
    int64_t pack_cur = 0, pack_prev = 0;
    for (int j = 0; j < n_regions; ++j) {
        const int64_t len = offsets[j + 1] - offsets[j];
        OA_REQUIRE(len >= 0, "oa_pjoin_plan_host: offsets must not decrease");
        int bits = 0;
        while (bits < OA_PJOIN_MAX_BITS && len > (target << bits)) ++bits;
        if (prev_bits[j] > bits) bits = prev_bits[j];
        OA_REQUIRE(bits <= OA_PJOIN_MAX_BITS, "oa_pjoin_plan_host: bad prev_bits");
        if (bits > max_bits) max_bits = bits;
        oa_pjoin_region& r = rows[j];
        r.pb_cur = (uint32_t)pb;
        r.pb_prev = prev_bits[j] >= 0 ? (uint32_t)prev_pb[j] : 0u;
        r.bits_cur = bits;
        r.bits_prev = prev_bits[j];
        r.tile_first = (uint32_t)tiles;
        r.count_first = (uint32_t)ctiles;
        r.join_first = (uint32_t)joins;
        r.scan_first = (uint32_t)scans;
        r.pack_len = 0;
        r.reserved[0] = r.reserved[1] = r.reserved[2] = 0;
        bits_out[j] = bits;
        pb_out[j] = (int64_t)pb;
        pb += ((uint64_t)1 << bits) + 1;
        const int64_t gid = offsets[j] / lag_particles;
        const bool new_group = j == 0 || gid != gid_prev;
        if (new_group) group_first[n_groups++] = (uint32_t)j;
        gid_prev = gid;
        if (bits > 0) {
            tiles += (uint64_t)((len + OA_PJOIN_TILE - 1) / OA_PJOIN_TILE);
            ctiles += (uint64_t)((len + OA_PJOIN_CTILE - 1) / OA_PJOIN_CTILE);
            scans += 1;
            if (prev_bits[j] >= 0) joins += (uint64_t)1 << prev_bits[j];
            leader = -1;
        } else {
            // packs of consecutive small regions of one group
            const int64_t plen = prev_bits[j] >= 0 ? prev_counts[j] : 0;
            const bool fits = leader >= 0 && !new_group &&
                              rows[leader].pack_len < OA_PJOIN_PACK_MAX &&
                              pack_cur + len <= OA_PJOIN_TILE &&
                              pack_prev + plen <= OA_PJOIN_REC_CAP;
            if (fits) {
                rows[leader].pack_len += 1;
                pack_cur += len;
                pack_prev += plen;
            } else {
                leader = j;
                r.pack_len = 1;
                pack_cur = len;
                pack_prev = plen;
                joins += 1;
            }
        }
    }
    OA_REQUIRE(pb < ((uint64_t)1 << 32) && tiles + ctiles + joins + scans < ((uint64_t)1 << 32),
               "oa_pjoin_plan_host: plan does not fit 32-bit counters");
    oa_pjoin_region& e = rows[n_regions];
    e.pb_cur = (uint32_t)pb;
    e.pb_prev = 0;
    e.bits_cur = e.bits_prev = 0;
    e.tile_first = (uint32_t)tiles;
    e.count_first = (uint32_t)ctiles;
    e.join_first = (uint32_t)joins;
    e.scan_first = (uint32_t)scans;
    e.pack_len = 0;
    e.reserved[0] = e.reserved[1] = e.reserved[2] = 0;
    group_first[n_groups] = (uint32_t)n_regions;

    // ticket ranges: superstep s holds JOIN of group s-3, SCATTER s-2, SCAN s-1, COUNT s
    const int n_ranges = 4 * (n_groups + 3);
    uint32_t t = 0;
    for (int s = 0; s < n_groups + 3; ++s)
        for (int st = 0; st < 4; ++st) {
            range_start[4 * s + st] = t;
            const int g = s - (3 - st);
            if (g < 0 || g >= n_groups) continue;
            const oa_pjoin_region &lo = rows[group_first[g]], &hi = rows[group_first[g + 1]];
            t += st == pj::JOIN ? hi.join_first - lo.join_first
               : st == pj::SCAN ? hi.scan_first - lo.scan_first
               : st == pj::COUNT ? hi.count_first - lo.count_first
                                 : hi.tile_first - lo.tile_first;
        }
    range_start[n_ranges] = t;
    info->n_part_entries = (int64_t)pb;
    info->total_tickets = t;
    info->n_groups = n_groups;
    info->n_ranges = n_ranges;
    info->max_bits = max_bits;
    return OA_OK;
}

extern "C" int oa_pjoin_step(const oa_pjoin_args* args, void* stream) {
    OA_REQUIRE(args != nullptr, "oa_pjoin_step: args is NULL");
    const oa_pjoin_args& a = *args;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OA_REQUIRE(a.n_cur >= 0 && a.n_regions >= 0 && a.n_groups >= 0, "oa_pjoin_step: negative size");
    OA_REQUIRE(a.n_cur < ((int64_t)1 << 32) && a.n_prev < ((int64_t)1 << 32),
               "oa_pjoin_step: more than 2^32-1 region-particles on one GPU");
    OA_REQUIRE(a.mode == OA_MODE_PERICENTRIC || a.mode == OA_MODE_APOCENTRIC,
               "oa_pjoin_step: bad mode %d", a.mode);
    if (a.n_regions == 0) return OA_OK;
    OA_REQUIRE(a.regions && a.plan && a.group_first && a.range_start && a.rec_cur &&
               a.part_off_cur && a.mark_cur && a.workspace,
               "oa_pjoin_step: NULL required pointer");
    OA_REQUIRE(a.n_cur == 0 || (a.pos && a.vel && a.ids), "oa_pjoin_step: NULL input array");
    OA_REQUIRE(a.n_ranges == 4 * (a.n_groups + 3), "oa_pjoin_step: n_ranges != 4 (n_groups + 3)");
    OA_REQUIRE(a.n_prev == 0 || !a.rec_prev || (a.part_off_prev && a.mark_prev),
               "oa_pjoin_step: previous generation incomplete");
    OA_REQUIRE(a.workspace_bytes >=
                   oa_pjoin_workspace_bytes(a.n_regions, a.n_part_entries, a.total_tickets),
               "oa_pjoin_step: workspace too small (need oa_pjoin_workspace_bytes)");
    OA_REQUIRE((reinterpret_cast<uintptr_t>(a.rec_cur) & 31u) == 0 &&
               (reinterpret_cast<uintptr_t>(a.rec_prev) & 31u) == 0,
               "oa_pjoin_step: records must be 32-byte aligned");

    static bool configured = false;
    if (!configured) {
        OA_CUDA_CHECK(cudaFuncSetAttribute(oa_pjoin_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           pj::SM_BYTES));
        configured = true;
    }
    int dev = 0, sms = OA_NUM_SMS, per_sm = 0;
    OA_CUDA_CHECK(cudaGetDevice(&dev));
    OA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // every CTA of the grid must be resident: items wait for items with smaller
    // tickets, which only resident CTAs can hold
    OA_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, oa_pjoin_kernel,
                                                                pj::THREADS, pj::SM_BYTES));
    OA_REQUIRE(per_sm >= 1, "oa_pjoin_step: the kernel does not fit an SM");
    if (a.sm_reserve > 0 && a.sm_reserve < sms) sms -= a.sm_reserve;

    // (the ticket ranges live on the device; the host that built them passes
    // their total, a read-back would cost a synchronisation)
    const uint32_t total = a.total_tickets;
    if (total == 0) return OA_OK;

    pj::Const k;
    for (int q = 0; q < 3; ++q) {
        const double h = a.box[q] * 0.5;           // largest float <= L/2
        float hf = (float)h;
        if ((double)hf > h) hf = nextafterf(hf, -INFINITY);
        k.half_box[q] = hf;
    }
    k.total_tickets = total;

    OA_REQUIRE((reinterpret_cast<uintptr_t>(a.workspace) & 7u) == 0,
               "oa_pjoin_step: workspace must be 8-byte aligned");
    pj::Work w;
    w.items = static_cast<uint64_t*>(a.workspace);
    uint32_t* ws = reinterpret_cast<uint32_t*>(w.items + total);
    OA_CUDA_CHECK(cudaMemsetAsync(ws, 0, 4 * work_words(a.n_regions, a.n_part_entries), st));
    w.ticket = ws;
    w.done_count = ws + 4;
    w.done_scan = w.done_count + a.n_regions;
    w.done_scatter = w.done_scan + a.n_regions;
    w.cursor = w.done_scatter + a.n_regions;
    oa_pjoin_expand_kernel<<<(unsigned)a.n_regions, 128, 0, st>>>(a, w.items);
    OA_LAUNCH_CHECK();

    int64_t grid = (int64_t)sms * per_sm;
    if (grid > (int64_t)total) grid = total;
    oa_pjoin_kernel<<<(unsigned)grid, pj::THREADS, pj::SM_BYTES, st>>>(a, k, w);
    OA_LAUNCH_CHECK();
    return OA_OK;
}
