// Test infrastructure: the partitioned-join stage code (csrc/oa_pjoin_core.cuh)
// compiled for the HOST.  A CTA is THREADS real threads and a pthread barrier,
// atomics are the compiler's, and several CTAs run concurrently so that the
// ticket order, the dependency counters and the shared-memory phases are
// exercised as on the device.  Built by tests/test_pjoin_emul.py with
//   g++ -O1 -ffp-contract=off -pthread -shared -fPIC -DPJ_HOST_EMUL
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../nbody_orbit_analysis_b200/csrc/oa_pjoin_core.cuh"

namespace {

struct Cta {
    pthread_barrier_t barrier;
    unsigned char* smem;
};

struct HostCtx {
    Cta* cta;
    int tid_;
    int tid() const { return tid_; }
    void sync() const { pthread_barrier_wait(&cta->barrier); }
    unsigned char* smem() const { return cta->smem; }
    uint32_t atomic_add(uint32_t* p, uint32_t v) const {
        return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
    }
    uint32_t atomic_cas(uint32_t* p, uint32_t c, uint32_t v) const {
        __atomic_compare_exchange_n(p, &c, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
        return c;
    }
    uint32_t load_acquire(const uint32_t* p) const { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
    void release_add(uint32_t* p, uint32_t v) const { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
    void backoff() const { sched_yield(); }
    void fail() const { abort(); }
    uint32_t ld_cg(const uint32_t* p) const { return __atomic_load_n(p, __ATOMIC_RELAXED); }
    pj::U4 ld_stream(const pj::U4* p) const { return *p; }
    void bulk_load(void* dst, const void* src, uint32_t bytes) const {
        if (tid_ == 0) memcpy(dst, src, bytes);
        sync();
    }
    int64_t ld_last(const int64_t* p) const { return *p; }
    float ld_last(const float* p) const { return *p; }
    pj::Rec load_rec_cg(const pj::Rec* p) const { return *p; }
    void store_rec(pj::Rec* p, const pj::Rec& r) const { *p = r; }
    uint64_t clock() const { return 0; }
    void stat_add(int, uint64_t) const {}
};

struct ThreadArg {
    HostCtx cx;
    const oa_pjoin_args* a;
    const pj::Const* k;
    const pj::Work* w;
};

void* thread_main(void* p) {
    ThreadArg* t = static_cast<ThreadArg*>(p);
    pj::run(t->cx, *t->a, *t->k, *t->w);
    return nullptr;
}

}  // namespace

extern "C" int pj_emul_config(int32_t* out8) {
    out8[0] = pj::THREADS;
    out8[1] = OA_PJOIN_MIN_CTAS;
    out8[2] = pj::TILE;
    out8[3] = pj::CTILE;
    out8[4] = pj::REC_CAP;
    out8[5] = OA_PJOIN_TARGET;
    out8[6] = pj::MAX_BITS;
    out8[7] = pj::SM_BYTES;
    return (int)sizeof(oa_pjoin_args);
}

// float16 helpers exposed for the test of the conversions themselves
extern "C" unsigned pj_emul_half_bits(float f) { return pj::half_bits(f); }
extern "C" float pj_emul_half_value(unsigned b) { return pj::half_value((uint16_t)b); }

// Same contract as oa_pjoin_step with HOST pointers; `n_ctas` CTAs run concurrently.
extern "C" int pj_emul_step(const oa_pjoin_args* args, int n_ctas) {
    const oa_pjoin_args& a = *args;
    if (a.n_regions == 0 || a.total_tickets == 0) return 0;
    pj::Const k;
    for (int q = 0; q < 3; ++q) {
        const double h = a.box[q] * 0.5;
        float hf = (float)h;
        if ((double)hf > h) hf = nextafterf(hf, -INFINITY);
        k.half_box[q] = hf;
    }
    k.total_tickets = a.total_tickets;
    pj::Work w;
    w.items = static_cast<uint64_t*>(a.workspace);
    uint32_t* ws = reinterpret_cast<uint32_t*>(w.items + a.total_tickets);
    memset(ws, 0, 4 * (4 + 3 * (size_t)a.n_regions + (size_t)a.n_part_entries));
    w.ticket = ws;
    w.done_count = ws + 4;
    w.done_scan = w.done_count + a.n_regions;
    w.done_scatter = w.done_scan + a.n_regions;
    w.cursor = w.done_scatter + a.n_regions;
    for (int j = 0; j < a.n_regions; ++j)        // the expand pre-kernel
        for (int t = 0; t < 128; ++t) pj::expand_region(a, w.items, j, t, 128);


    std::vector<Cta> ctas(n_ctas);
    std::vector<ThreadArg> targs((size_t)n_ctas * pj::THREADS);
    std::vector<pthread_t> th(targs.size());
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    for (int c = 0; c < n_ctas; ++c) {
        pthread_barrier_init(&ctas[c].barrier, nullptr, pj::THREADS);
        // poisoned shared memory: nothing may rely on its initial content
        ctas[c].smem = static_cast<unsigned char*>(aligned_alloc(128, (pj::SM_BYTES + 127) / 128 * 128));
        memset(ctas[c].smem, 0xA5, pj::SM_BYTES);
        for (int t = 0; t < pj::THREADS; ++t) {
            ThreadArg& ta = targs[(size_t)c * pj::THREADS + t];
            ta.cx.cta = &ctas[c];
            ta.cx.tid_ = t;
            ta.a = &a;
            ta.k = &k;
            ta.w = &w;
        }
    }
    int rc = 0;
    for (size_t i = 0; i < targs.size(); ++i)
        if (pthread_create(&th[i], &attr, thread_main, &targs[i]) != 0) { rc = -1; th.resize(i); break; }
    for (size_t i = 0; i < th.size(); ++i) pthread_join(th[i], nullptr);
    for (int c = 0; c < n_ctas; ++c) {
        pthread_barrier_destroy(&ctas[c].barrier);
        free(ctas[c].smem);
    }
    pthread_attr_destroy(&attr);
    return rc;
}
