// Segment helpers for the sorted per-halo ID lists of the on-the-fly path
// (np.setdiff1d returns sorted unique values, track_orbits_onthefly.py:145,168)
// and for the per-halo unique/count reductions of postprocessing.py:133-141.
#include "oa_common.cuh"

namespace {

// segment of element i: last s with seg_off[s] <= i   (seg_off has n_seg+1 entries)
OA_D int find_segment(const int64_t* __restrict__ seg_off, int n_seg, int64_t i) {
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(seg_off + mid) <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// key_lo = (id - id_min) for segments that must be sorted by ID, otherwise the
// element's own position (keeps the original order); key_hi = segment index.
__global__ void segment_sort_keys_kernel(const int64_t* __restrict__ ids, int64_t n,
                                         const int64_t* __restrict__ seg_off, int n_seg,
                                         const uint8_t* __restrict__ sort_flag,
                                         const int64_t* __restrict__ id_minmax,
                                         uint64_t* __restrict__ key_lo,
                                         uint64_t* __restrict__ key_hi,
                                         uint64_t* __restrict__ index) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = find_segment(seg_off, n_seg, i);
    const bool sorted = sort_flag == nullptr || sort_flag[s] != 0;
    key_lo[i] = sorted ? (uint64_t)(ids[i] - __ldg(id_minmax)) : (uint64_t)i;
    key_hi[i] = (uint64_t)s;
    index[i] = (uint64_t)i;
}

// run-length encoding of a sorted (segment, id) sequence: head[i] = 1 where a new
// (segment, id) run starts
__global__ void run_heads_kernel(const uint64_t* __restrict__ seg, const int64_t* __restrict__ ids,
                                 int64_t n, uint16_t* __restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || seg[i] != seg[i - 1] || ids[i] != ids[i - 1]) ? 1 : 0;
}

// counts[k] = start of run k+1 - start of run k   (starts = selected positions)
__global__ void run_lengths_kernel(const int64_t* __restrict__ starts, int64_t n_runs,
                                   int64_t n, int64_t* __restrict__ counts) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_runs) return;
    const int64_t next = (k + 1 < n_runs) ? starts[k + 1] : n;
    counts[k] = next - starts[k];
}

// Multi-way merge of n_lists ascending key lists laid out back to back
// (list r = [list_off[r], list_off[r+1])): the output position of an element is
// its own rank plus the number of smaller keys in every other list (keys are
// distinct across lists: they are positions in one unsharded snapshot).
__global__ void merge_lists_kernel(const int64_t* __restrict__ keys,
                                   const int64_t* __restrict__ ids,
                                   const uint16_t* __restrict__ angles, int64_t n,
                                   const int64_t* __restrict__ list_off, int n_lists,
                                   int64_t* __restrict__ ids_out,
                                   uint16_t* __restrict__ angles_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t key = keys[i];
    int64_t pos = 0;
    for (int r = 0; r < n_lists; ++r) {
        const int64_t b = __ldg(list_off + r), e = __ldg(list_off + r + 1);
        if (i >= b && i < e) { pos += i - b; continue; }
        int64_t lo = b, hi = e;                 // first element of list r with key' >= key
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(keys + mid) < key) lo = mid + 1; else hi = mid;
        }
        pos += lo - b;
    }
    ids_out[pos] = ids[i];
    angles_out[pos] = angles[i];
}

// ---- multi-GPU event exchange without host round trips ---------------------------------
// Send buffer of one rank (bytes, 16-aligned sections):
//   [ int64 size | int64 counts[n_seg] | pad | int64 keys[cap] | int64 ids[cap] |
//     uint16 angles[cap] | pad ]
// all ranks use the same n_seg and cap, so the all-gathered buffer is `world`
// such chunks back to back.
struct ExchangeLayout {
    int64_t keys, ids, angles, bytes;
};
__host__ __device__ inline ExchangeLayout exchange_layout(int n_seg, int64_t cap) {
    ExchangeLayout L;
    int64_t o = 8 * (1 + (int64_t)n_seg);
    o = (o + 15) & ~15ll;
    L.keys = o;
    o += 8 * cap;
    L.ids = o;
    o += 8 * cap;
    L.angles = o;
    o += 2 * cap;
    L.bytes = (o + 15) & ~15ll;
    return L;
}

// small[0..n_seg) = per-segment offsets of the local event list, small[n_seg] =
// its length (as written by oa_segment_offsets / oa_select_count)
__global__ void pack_events_kernel(const int64_t* __restrict__ gpos,
                                   const int64_t* __restrict__ sel,
                                   const int64_t* __restrict__ ids,
                                   const uint16_t* __restrict__ angles,
                                   const int64_t* __restrict__ small, int n_seg, int64_t cap,
                                   unsigned char* __restrict__ out) {
    const ExchangeLayout L = exchange_layout(n_seg, cap);
    const int64_t total = small[n_seg];
    const int64_t m = total < cap ? total : cap;
    int64_t* hdr = reinterpret_cast<int64_t*>(out);
    int64_t* okeys = reinterpret_cast<int64_t*>(out + L.keys);
    int64_t* oids = reinterpret_cast<int64_t*>(out + L.ids);
    uint16_t* oang = reinterpret_cast<uint16_t*>(out + L.angles);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t0 == 0) hdr[0] = total;                       // true size, even if truncated
    for (int64_t s = t0; s < n_seg; s += stride)
        hdr[1 + s] = (s + 1 < n_seg ? small[s + 1] : total) - small[s];
    for (int64_t i = t0; i < m; i += stride) {
        okeys[i] = gpos[sel[i]];
        oids[i] = ids[i];
        oang[i] = angles[i];
    }
}

// global per-segment offsets = exclusive scan of the summed per-rank counts
// (one block; n_seg is the number of halos), total and per-rank sizes
__global__ void merge_offsets_kernel(const unsigned char* __restrict__ gathered, int world,
                                     int n_seg, int64_t cap, int64_t* __restrict__ info) {
    // info: [total | offsets[n_seg + 1] | sizes[world] | overflow]
    const ExchangeLayout L = exchange_layout(n_seg, cap);
    __shared__ int64_t s_carry;
    __shared__ int64_t s_part[1024];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_seg; base += blockDim.x) {
        const int s = base + threadIdx.x;
        int64_t v = 0;
        if (s < n_seg)
            for (int r = 0; r < world; ++r)
                v += reinterpret_cast<const int64_t*>(gathered + (size_t)r * L.bytes)[1 + s];
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < (int)blockDim.x; d <<= 1) {        // inclusive scan
            const int64_t t = threadIdx.x >= (unsigned)d ? s_part[threadIdx.x - d] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        if (s < n_seg) info[1 + s] = s_carry + s_part[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry += s_part[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int64_t total = 0, overflow = 0;
        for (int r = 0; r < world; ++r) {
            const int64_t sz = reinterpret_cast<const int64_t*>(gathered + (size_t)r * L.bytes)[0];
            info[2 + n_seg + r] = sz;
            if (sz > cap) overflow = 1;
            total += sz < cap ? sz : cap;
        }
        info[0] = total;
        info[1 + n_seg] = s_carry;
        info[2 + n_seg + world] = overflow;
    }
}

// every gathered record finds its place: own rank + number of smaller keys in
// the other ranks' (ascending) lists
__global__ void merge_gathered_kernel(const unsigned char* __restrict__ gathered, int world,
                                      int n_seg, int64_t cap, int64_t* __restrict__ ids_out,
                                      uint16_t* __restrict__ angles_out) {
    const ExchangeLayout L = exchange_layout(n_seg, cap);
    const int r = blockIdx.y;
    const unsigned char* mine = gathered + (size_t)r * L.bytes;
    int64_t size = reinterpret_cast<const int64_t*>(mine)[0];
    if (size > cap) size = cap;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    const int64_t key = reinterpret_cast<const int64_t*>(mine + L.keys)[i];
    int64_t pos = i;
    for (int q = 0; q < world; ++q) {
        if (q == r) continue;
        const unsigned char* other = gathered + (size_t)q * L.bytes;
        int64_t hi = reinterpret_cast<const int64_t*>(other)[0];
        if (hi > cap) hi = cap;
        const int64_t* keys = reinterpret_cast<const int64_t*>(other + L.keys);
        int64_t lo = 0;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(keys + mid) < key) lo = mid + 1; else hi = mid;
        }
        pos += lo;
    }
    ids_out[pos] = reinterpret_cast<const int64_t*>(mine + L.ids)[i];
    angles_out[pos] = reinterpret_cast<const uint16_t*>(mine + L.angles)[i];
}

// ---- scalable variant: key-range partitioned merge (all-to-all) ---------------------------
// Every rank's events are a uniform sample of all events (particles are sharded
// by ID), so local quantiles of the order key are global quantiles: they split
// the key space into `world` ranges of (nearly) equal event counts.  Rank q
// receives from every rank the events of range q, merges the `world` ascending
// lists and owns that contiguous part of the global lists.  Volume per rank:
// its own share, whatever the number of GPUs.
//
// local quantile keys: q_out[j] = key of local event floor(total * (j+1) / world)
__global__ void quantile_keys_kernel(const int64_t* __restrict__ gpos,
                                     const int64_t* __restrict__ sel,
                                     const int64_t* __restrict__ small, int n_seg, int world,
                                     int64_t* __restrict__ q_out) {
    const int j = threadIdx.x;
    if (j >= world - 1) return;
    const int64_t total = small[n_seg];
    // a rank without events proposes +inf (it is out-voted by the median below)
    q_out[j] = total > 0 ? gpos[sel[total * (j + 1) / world]] : INT64_MAX;
}

// splitters = per-quantile median over the ranks' proposals; bounds of the local
// (ascending) event list: bnd[q] = first local event with key >= splitter[q-1]
__global__ void split_bounds_kernel(const int64_t* __restrict__ proposals, int world,
                                    const int64_t* __restrict__ gpos,
                                    const int64_t* __restrict__ sel,
                                    const int64_t* __restrict__ small, int n_seg,
                                    int64_t* __restrict__ bnd) {
    const int q = threadIdx.x;                      // 0 .. world
    if (q > world) return;
    const int64_t total = small[n_seg];
    if (q == 0) { bnd[0] = 0; return; }
    if (q == world) { bnd[world] = total; return; }
    // median of proposals[r][q-1] over ranks r (world <= 64: insertion into registers)
    int64_t v[64];
    for (int r = 0; r < world; ++r) {
        int64_t x = proposals[r * (world - 1) + (q - 1)];
        int k = r;
        while (k > 0 && v[k - 1] > x) { v[k] = v[k - 1]; --k; }
        v[k] = x;
    }
    const int64_t split = v[(world - 1) / 2];
    int64_t lo = 0, hi = total;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (gpos[sel[mid]] < split) lo = mid + 1; else hi = mid;
    }
    bnd[q] = lo;
}

// send buffer: `world` blocks of exchange_layout(0, cap) bytes, block q = events
// bnd[q] .. bnd[q+1]; counts[] = per-segment event counts of this rank
__global__ void pack_split_kernel(const int64_t* __restrict__ gpos,
                                  const int64_t* __restrict__ sel,
                                  const int64_t* __restrict__ ids,
                                  const uint16_t* __restrict__ angles,
                                  const int64_t* __restrict__ small, int n_seg,
                                  const int64_t* __restrict__ bnd, int world, int64_t cap,
                                  unsigned char* __restrict__ out,
                                  int64_t* __restrict__ counts) {
    const ExchangeLayout L = exchange_layout(0, cap);
    const int64_t total = small[n_seg];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t s = t0; s < n_seg; s += stride)
        counts[s] = (s + 1 < n_seg ? small[s + 1] : total) - small[s];
    if (t0 < world)
        reinterpret_cast<int64_t*>(out + (size_t)t0 * L.bytes)[0] = bnd[t0 + 1] - bnd[t0];
    for (int64_t i = t0; i < total; i += stride) {
        int q = 0;
        while (q + 1 < world && i >= bnd[q + 1]) ++q;
        const int64_t k = i - bnd[q];
        if (k >= cap) continue;                       // overflow: flagged by the receiver
        unsigned char* blk = out + (size_t)q * L.bytes;
        reinterpret_cast<int64_t*>(blk + L.keys)[k] = gpos[sel[i]];
        reinterpret_cast<int64_t*>(blk + L.ids)[k] = ids[i];
        reinterpret_cast<uint16_t*>(blk + L.angles)[k] = angles[i];
    }
}

// Batched exchange: the local events of one snapshot are appended to a staging
// area that collects several snapshots (keys tagged with the snapshot's slot in
// their high bits, segment offsets shifted by the events staged before), so that
// ONE exchange orders the events of all of them.  stage_small[n_seg] = running
// total (the next snapshot overwrites it with its first offset, the same value).
__global__ void stage_events_kernel(const int64_t* __restrict__ gpos,
                                    const int64_t* __restrict__ sel,
                                    const int64_t* __restrict__ ids,
                                    const uint16_t* __restrict__ angles,
                                    const int64_t* __restrict__ small, int n_seg,
                                    int64_t n_local, int64_t tag, int64_t ev_base,
                                    int64_t* __restrict__ out_keys,
                                    int64_t* __restrict__ out_ids,
                                    uint16_t* __restrict__ out_angles,
                                    int64_t* __restrict__ out_small) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t s = t0; s < n_seg; s += stride) out_small[s] = small[s] + ev_base;
    if (t0 == 0) out_small[n_seg] = ev_base + n_local;
    for (int64_t i = t0; i < n_local; i += stride) {
        out_keys[ev_base + i] = gpos[sel[i]] | tag;
        out_ids[ev_base + i] = ids[i];
        out_angles[ev_base + i] = angles[i];
    }
}

// merge of the `world` received blocks (same layout, block r from rank r) into
// this rank's slice; info = [slice size | largest received block before
// truncation to cap (> cap: overflow, the exchange must be repeated)]
__global__ void merge_blocks_kernel(const unsigned char* __restrict__ recv, int world,
                                    int64_t cap, int64_t* __restrict__ ids_out,
                                    uint16_t* __restrict__ angles_out,
                                    int64_t* __restrict__ info) {
    const ExchangeLayout L = exchange_layout(0, cap);
    const int r = blockIdx.y;
    const unsigned char* mine = recv + (size_t)r * L.bytes;
    const int64_t true_size = reinterpret_cast<const int64_t*>(mine)[0];
    const int64_t size = true_size < cap ? true_size : cap;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && r == 0) {
        int64_t tot = 0, largest = 0;
        for (int q = 0; q < world; ++q) {
            const int64_t sz = reinterpret_cast<const int64_t*>(recv + (size_t)q * L.bytes)[0];
            if (sz > largest) largest = sz;
            tot += sz < cap ? sz : cap;
        }
        info[0] = tot;
        info[1] = largest;
    }
    // The keys of this CTA's elements are ascending, so their lower bounds in
    // another list lie between the lower bounds of the CTA's first and last key:
    // one full binary search per (CTA, list) and end, then every element searches
    // a window of about blockDim.x entries (a few cache lines) instead of the
    // whole list -- (world - 1) x 17 dependent, scattered loads per event were
    // slowing the tracking kernel the exchange overlaps with.
    __shared__ int64_t s_lo[64], s_hi[64];
    const int64_t first = (int64_t)blockIdx.x * blockDim.x;
    if (first >= size) return;                       // (uniform for the CTA)
    const int64_t last = min(first + (int64_t)blockDim.x, size) - 1;
    const int64_t* my_keys = reinterpret_cast<const int64_t*>(mine + L.keys);
    for (int q = threadIdx.x; q < 2 * world; q += blockDim.x) {
        const int list = q >> 1;
        const int64_t key = (q & 1) ? my_keys[last] : my_keys[first];
        const unsigned char* other = recv + (size_t)list * L.bytes;
        int64_t hi = reinterpret_cast<const int64_t*>(other)[0];
        if (hi > cap) hi = cap;
        const int64_t* keys = reinterpret_cast<const int64_t*>(other + L.keys);
        int64_t lo = 0;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(keys + mid) < key) lo = mid + 1; else hi = mid;
        }
        if (q & 1) s_hi[list] = lo; else s_lo[list] = lo;
    }
    __syncthreads();
    if (i >= size) return;
    const int64_t key = my_keys[i];
    int64_t pos = i;
    for (int q = 0; q < world; ++q) {
        if (q == r) continue;
        const int64_t* keys =
            reinterpret_cast<const int64_t*>(recv + (size_t)q * L.bytes + L.keys);
        int64_t lo = s_lo[q], hi = s_hi[q];
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(keys + mid) < key) lo = mid + 1; else hi = mid;
        }
        pos += lo;
    }
    ids_out[pos] = reinterpret_cast<const int64_t*>(mine + L.ids)[i];
    angles_out[pos] = reinterpret_cast<const uint16_t*>(mine + L.angles)[i];
}

inline unsigned blocks_for(int64_t n, int threads) {
    return (unsigned)((n + threads - 1) / threads);
}

}  // namespace

extern "C" int oa_segment_sort_keys(const int64_t* ids, int64_t n, const int64_t* seg_off,
                                    int n_seg, const uint8_t* sort_flag,
                                    const int64_t* id_minmax, uint64_t* key_lo,
                                    uint64_t* key_hi, uint64_t* index, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(ids && seg_off && id_minmax && key_lo && key_hi && index && n_seg >= 1,
               "oa_segment_sort_keys: bad arguments");
    segment_sort_keys_kernel<<<blocks_for(n, 256), 256, 0, st>>>(
        ids, n, seg_off, n_seg, sort_flag, id_minmax, key_lo, key_hi, index);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_merge_event_lists(const int64_t* keys, const int64_t* ids,
                                    const uint16_t* angles, int64_t n,
                                    const int64_t* list_off, int n_lists, int64_t* ids_out,
                                    uint16_t* angles_out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(keys && ids && angles && list_off && ids_out && angles_out && n_lists >= 1,
               "oa_merge_event_lists: bad arguments");
    merge_lists_kernel<<<blocks_for(n, 256), 256, 0, st>>>(keys, ids, angles, n, list_off,
                                                           n_lists, ids_out, angles_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" size_t oa_exchange_bytes(int n_seg, int64_t cap) {
    return (size_t)exchange_layout(n_seg, cap).bytes;
}

extern "C" int oa_pack_events(const int64_t* gpos, const int64_t* sel, const int64_t* ids,
                              const uint16_t* angles, const int64_t* small, int n_seg,
                              int64_t cap, void* out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // (the event count is read from small[n_seg] on the device: gpos / sel / ids /
    // angles are never dereferenced for an empty list and may then be NULL -- a
    // zero-element torch tensor has data_ptr() == 0)
    OA_REQUIRE(small && out && n_seg >= 0 && cap >= 1, "oa_pack_events: bad arguments");
    int64_t blocks = (cap + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    pack_events_kernel<<<(unsigned)blocks, 256, 0, st>>>(
        gpos, sel, ids, angles, small, n_seg, cap, static_cast<unsigned char*>(out));
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_merge_gathered(const void* gathered, int world, int n_seg, int64_t cap,
                                 int64_t* ids_out, uint16_t* angles_out, int64_t* info,
                                 void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OA_REQUIRE(gathered && ids_out && angles_out && info && world >= 1 && n_seg >= 0 && cap >= 1,
               "oa_merge_gathered: bad arguments");
    merge_offsets_kernel<<<1, 1024, 0, st>>>(static_cast<const unsigned char*>(gathered), world,
                                             n_seg, cap, info);
    OA_LAUNCH_CHECK();
    dim3 grid(blocks_for(cap, 256), (unsigned)world);
    merge_gathered_kernel<<<grid, 256, 0, st>>>(static_cast<const unsigned char*>(gathered),
                                                world, n_seg, cap, ids_out, angles_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_split_quantiles(const int64_t* gpos, const int64_t* sel,
                                  const int64_t* small, int n_seg, int world,
                                  int64_t* q_out, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // (gpos / sel may be NULL for an empty list, see oa_pack_events)
    OA_REQUIRE(small && q_out && world >= 1 && world <= 64,
               "oa_split_quantiles: bad arguments");
    if (world == 1) return OA_OK;
    quantile_keys_kernel<<<1, 64, 0, st>>>(gpos, sel, small, n_seg, world, q_out);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_pack_split(const int64_t* gpos, const int64_t* sel, const int64_t* ids,
                             const uint16_t* angles, const int64_t* small, int n_seg,
                             const int64_t* proposals, int world, int64_t cap,
                             int64_t* bnd_ws, void* out, int64_t* counts, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // (gpos / sel / ids / angles may be NULL for an empty list, see oa_pack_events)
    OA_REQUIRE(small && bnd_ws && out && counts &&
               world >= 1 && world <= 64 && cap >= 1 && (world == 1 || proposals),
               "oa_pack_split: bad arguments");
    split_bounds_kernel<<<1, 96, 0, st>>>(proposals, world, gpos, sel, small, n_seg, bnd_ws);
    OA_LAUNCH_CHECK();
    int64_t blocks = (cap * world + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    pack_split_kernel<<<(unsigned)blocks, 256, 0, st>>>(
        gpos, sel, ids, angles, small, n_seg, bnd_ws, world, cap,
        static_cast<unsigned char*>(out), counts);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_stage_events(const int64_t* gpos, const int64_t* sel, const int64_t* ids,
                               const uint16_t* angles, const int64_t* small, int n_seg,
                               int64_t n_local, int64_t tag, int64_t ev_base,
                               int64_t* out_keys, int64_t* out_ids, uint16_t* out_angles,
                               int64_t* out_small, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OA_REQUIRE(small && out_small && n_seg >= 0 && n_local >= 0 && ev_base >= 0 &&
               (n_local == 0 || (gpos && sel && ids && angles && out_keys && out_ids &&
                                 out_angles)),
               "oa_stage_events: bad arguments");
    int64_t work = n_local > n_seg ? n_local : n_seg;
    int64_t blocks = (work + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 4096) blocks = 4096;
    stage_events_kernel<<<(unsigned)blocks, 256, 0, st>>>(gpos, sel, ids, angles, small, n_seg,
                                                          n_local, tag, ev_base, out_keys,
                                                          out_ids, out_angles, out_small);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_merge_blocks(const void* recv, int world, int64_t cap, int64_t* ids_out,
                               uint16_t* angles_out, int64_t* info, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OA_REQUIRE(recv && ids_out && angles_out && info && world >= 1 && world <= 64 && cap >= 1,
               "oa_merge_blocks: bad arguments (at most 64 ranks)");
    dim3 grid(blocks_for(cap, 256), (unsigned)world);
    merge_blocks_kernel<<<grid, 256, 0, st>>>(static_cast<const unsigned char*>(recv), world,
                                              cap, ids_out, angles_out, info);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_run_heads(const uint64_t* seg, const int64_t* ids, int64_t n,
                            uint16_t* head, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return OA_OK;
    OA_REQUIRE(seg && ids && head, "oa_run_heads: NULL pointer");
    run_heads_kernel<<<blocks_for(n, 256), 256, 0, st>>>(seg, ids, n, head);
    OA_LAUNCH_CHECK();
    return OA_OK;
}

extern "C" int oa_run_lengths(const int64_t* starts, int64_t n_runs, int64_t n,
                              int64_t* counts, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_runs <= 0) return OA_OK;
    OA_REQUIRE(starts && counts, "oa_run_lengths: NULL pointer");
    run_lengths_kernel<<<blocks_for(n_runs, 256), 256, 0, st>>>(starts, n_runs, n, counts);
    OA_LAUNCH_CHECK();
    return OA_OK;
}
