// Native episode ingest: PrioritizedReplayBuffer.add for host-resident episodes as ONE call.
//
// Replaces the per-key NumPy assignments of DataStorage.add (algorithm/replay_buffer.py:30-56)
// and the max-priority insert of PrioritizedReplayBuffer.add (:293-315) on the actor -> learner
// boundary: the columns of the episode are packed into a pinned staging buffer (ring of 4, each
// guarded by an event), cross PCIe in one cudaMemcpyAsync and are written into the rings by
// k_storage_write_table; k_leaf_max + k_per_add give the new rows the current maximum priority.
// The handle owns only its staging buffers and events; rings, tree and ids stay the caller's.
#include <stdlib.h>
#include <string.h>

#include <new>

#include "common.cuh"
#include "tree_apply.cuh"

namespace asac {
// ---- small episodes (the actor -> learner traffic of a running learner: a handful of rows, <= 1024, a few KB):
// leaf max, ring writes and the max-priority insert in ONE launch instead of memset + three kernels.  Every CTA
// scans its share of the leaves (atomicMax on the float's bit pattern: priorities are >= 0) and takes a ticket; the
// CTA that draws the last one has seen every partial maximum and does the rest alone — the row copies (one warp per
// row) and k_per_add's leaf write + ancestor recompute (block_tree_apply), then resets the two scratch words.
// Same reads, same values and same order of writes as asac_tree_leaf_max -> asac_storage_write_table ->
// asac_per_add: the tree and the rings come out bit-identical.
struct IngestArgs {
    AsacWriteTable table;
    float *nodes;
    int64_t capacity, first_id, T;
    int levels, ignore_size, buffer_empty;
    int64_t *store_ids;
    unsigned int *scratch;   // [0] max bits, [1] tickets
    const float *td_max;
};

__device__ __forceinline__ void ingest_lane_copy(void *dst, const void *src, int64_t bytes, int lane) {
    const uintptr_t bits = (uintptr_t)dst | (uintptr_t)src | (uintptr_t)bytes;
    if ((bits & 15) == 0) {
        const int4 *s4 = reinterpret_cast<const int4 *>(src);
        int4 *d4 = reinterpret_cast<int4 *>(dst);
        for (int64_t i = lane; i < (bytes >> 4); i += 32) d4[i] = s4[i];
    } else if ((bits & 3) == 0) {
        const int32_t *s1 = reinterpret_cast<const int32_t *>(src);
        int32_t *d1 = reinterpret_cast<int32_t *>(dst);
        for (int64_t i = lane; i < (bytes >> 2); i += 32) d1[i] = s1[i];
    } else {
        const uint8_t *s0 = reinterpret_cast<const uint8_t *>(src);
        uint8_t *d0 = reinterpret_cast<uint8_t *>(dst);
        for (int64_t i = lane; i < bytes; i += 32) d0[i] = s0[i];
    }
}

__global__ void __launch_bounds__(1024) k_ingest_small(const IngestArgs a) {
    __shared__ TreeApplySmem s_apply;
    __shared__ float s_m[32];
    __shared__ int s_last;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (!a.buffer_empty) {  // the reference's self._sum_tree.max (replay_buffer.py:296-299)
        const float4 *v4 = reinterpret_cast<const float4 *>(a.nodes + a.capacity);
        const int64_t n4 = a.capacity >> 2;
        float m = 0.f;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + t; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
            const float4 v = __ldg(v4 + i);
            m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        }
        for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + t; i < a.capacity;
             i += (int64_t)gridDim.x * blockDim.x)
            m = fmaxf(m, a.nodes[a.capacity + i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) s_m[warp] = m;
        __syncthreads();
        if (t == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_m[w]);
            atomicMax(a.scratch, __float_as_uint(m));
        }
    }
    if (t == 0) {
        __threadfence();
        s_last = atomicAdd(a.scratch + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float max_p = a.buffer_empty ? a.td_max[0] : __uint_as_float(*reinterpret_cast<volatile unsigned int *>(a.scratch));
    // DataStorage.add (replay_buffer.py:30-56): one warp per row, all columns
    for (int64_t w = warp; w < a.T; w += blockDim.x >> 5) {
        const int64_t id = (a.first_id + w) % (10 * a.capacity);
        const int64_t slot = id & (a.capacity - 1);
        for (int c = 0; c < a.table.n_columns; ++c) {
            const int64_t rb = a.table.col[c].row_bytes;
            ingest_lane_copy(reinterpret_cast<uint8_t *>(a.table.col[c].ring) + slot * rb,
                             reinterpret_cast<const uint8_t *>(a.table.col[c].rows) + w * rb, rb, lane);
        }
    }
    // PrioritizedReplayBuffer.add (:293-307): the body of k_per_add for ONE chunk (T <= 1024)
    bool active = t < a.T;
    int slot = 0;
    float value = 0.f;
    if (active) {
        const int64_t i = t;
        active = (i + a.capacity >= a.T);
        const int64_t id = (a.first_id + i) % (10 * a.capacity);
        slot = (int)(id & (a.capacity - 1));
        if (active) a.store_ids[slot] = id;
        value = max_p;
        if (a.ignore_size > 0 && (slot >= a.capacity - a.ignore_size || i >= a.T - a.ignore_size)) value = 0.f;
    }
    block_tree_apply(a.nodes, a.capacity, a.levels, slot, value, active, s_apply);
    if (t == 0) {
        a.scratch[0] = 0u;
        a.scratch[1] = 0u;
    }
}
}  // namespace asac

namespace {
constexpr int kSlots = 4;
struct StageSlot {
    uint8_t *host = nullptr, *dev = nullptr;
    size_t bytes = 0;
    cudaEvent_t done = nullptr;
    bool used = false;
};
inline size_t align16(size_t n) { return (n + 15) & ~(size_t)15; }
}  // namespace

struct AsacIngest {
    int64_t capacity;
    int n_columns;
    void *rings[ASAC_MAX_COLUMNS];
    int64_t row_bytes[ASAC_MAX_COLUMNS];
    float *nodes;
    int64_t *store_ids;
    float *max_p;            // device scratch scalar
    const float *td_max;     // device scalar: priority of the very first rows (td_error_max)
    int device;
    int next;
    unsigned int *scratch;   // device: [0] max bits, [1] tickets of k_ingest_small (both zero between launches)
    StageSlot slot[kSlots];
};

extern "C" int asac_ingest_create(AsacIngest **out, int64_t capacity, int n_columns, void *const *rings,
                                  const int64_t *row_bytes, float *nodes, int64_t *store_ids, float *max_p_scratch,
                                  const float *td_max_dev) {
    ASAC_REQUIRE(out && rings && row_bytes && nodes && store_ids && max_p_scratch && td_max_dev,
                 "asac_ingest_create: null argument");
    ASAC_REQUIRE(n_columns > 0 && n_columns <= ASAC_MAX_COLUMNS, "asac_ingest_create: %d columns", n_columns);
    ASAC_REQUIRE(capacity > 0 && (capacity & (capacity - 1)) == 0, "asac_ingest_create: capacity is not a power of two");
    AsacIngest *h = new (std::nothrow) AsacIngest();
    ASAC_REQUIRE(h != nullptr, "asac_ingest_create: out of memory");
    h->capacity = capacity;
    h->n_columns = n_columns;
    for (int c = 0; c < n_columns; ++c) {
        h->rings[c] = rings[c];
        h->row_bytes[c] = row_bytes[c];
    }
    h->nodes = nodes; h->store_ids = store_ids; h->max_p = max_p_scratch; h->td_max = td_max_dev;
    h->next = 0;
    cudaError_t e = cudaGetDevice(&h->device);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->scratch, 2 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(h->scratch, 0, 2 * sizeof(unsigned int));
    if (e != cudaSuccess) {
        delete h;
        asac::set_error("asac_ingest_create: %s", cudaGetErrorString(e));
        return (int)e;
    }
    *out = h;
    return ASAC_OK;
}

extern "C" void asac_ingest_destroy(AsacIngest *h) {
    if (!h) return;
    for (StageSlot &s : h->slot) {
        if (s.done) { cudaEventSynchronize(s.done); cudaEventDestroy(s.done); }
        if (s.host) cudaFreeHost(s.host);
        if (s.dev) cudaFree(s.dev);
    }
    if (h->scratch) cudaFree(h->scratch);
    delete h;
}

extern "C" int64_t asac_ingest_row_bytes(const AsacIngest *h) {
    int64_t n = 0;
    for (int c = 0; c < h->n_columns; ++c) n += h->row_bytes[c];
    return n;
}

// host_columns[c]: [T, row_bytes[c]] contiguous host memory (pageable or pinned), T <= capacity.
// first_id: DataStorage._id before the call; buffer_empty: the buffer held no rows (size == 0).
extern "C" int asac_ingest_add(AsacIngest *h, const void *const *host_columns, int64_t T, int64_t first_id,
                               int ignore_size, int buffer_empty, void *stream) {
    ASAC_REQUIRE(h && host_columns, "asac_ingest_add: null argument");
    ASAC_REQUIRE(T > 0 && T <= h->capacity, "asac_ingest_add: T %lld outside (0, capacity]", (long long)T);
    cudaStream_t st = (cudaStream_t)stream;
    size_t total = 0;
    for (int c = 0; c < h->n_columns; ++c) total += align16((size_t)(T * h->row_bytes[c]));
    StageSlot &s = h->slot[h->next];
    h->next = (h->next + 1) % kSlots;
    if (s.used) ASAC_CUDA(cudaEventSynchronize(s.done));  // the copy that last read this pinned buffer is done
    if (s.bytes < total) {
        const size_t want = total < ((size_t)1 << 16) ? ((size_t)1 << 16) : total * 2;
        if (s.host) ASAC_CUDA(cudaFreeHost(s.host));
        if (s.dev) ASAC_CUDA(cudaFree(s.dev));
        s.host = s.dev = nullptr; s.bytes = 0;
        ASAC_CUDA(cudaHostAlloc((void **)&s.host, want, cudaHostAllocDefault));
        ASAC_CUDA(cudaMalloc((void **)&s.dev, want));
        s.bytes = want;
    }
    if (!s.done) ASAC_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    AsacWriteTable table;
    memset(&table, 0, sizeof(table));
    table.n_columns = h->n_columns;
    size_t off = 0;
    for (int c = 0; c < h->n_columns; ++c) {
        const size_t n = (size_t)(T * h->row_bytes[c]);
        if (n) memcpy(s.host + off, host_columns[c], n);
        table.col[c].ring = h->rings[c];
        table.col[c].rows = s.dev + off;
        table.col[c].row_bytes = h->row_bytes[c];
        off += align16(n);
    }
    if (off) ASAC_CUDA(cudaMemcpyAsync(s.dev, s.host, off, cudaMemcpyHostToDevice, st));
    ASAC_CUDA(cudaEventRecord(s.done, st));
    s.used = true;
    static const bool fused_small = [] {
        const char *e = getenv("ASAC_INGEST_FUSED");
        return !(e && e[0] == '0');
    }();
    if (fused_small && T <= 1024 && off <= ((size_t)256 << 10)) {
        asac::IngestArgs a;
        a.table = table;
        a.nodes = h->nodes; a.capacity = h->capacity; a.first_id = first_id; a.T = T;
        a.levels = asac::tree_levels(h->capacity); a.ignore_size = ignore_size; a.buffer_empty = buffer_empty;
        a.store_ids = h->store_ids; a.scratch = h->scratch; a.td_max = h->td_max;
        // the ring ids of DataStorage wrap at 10 x capacity (replay_buffer.py:49); write_table gets them reduced
        a.first_id = first_id;
        int64_t want = (h->capacity / 4 + 1023) / 1024;
        const int blocks = buffer_empty ? 1 : (int)(want < 1 ? 1 : (want > 64 ? 64 : want));
        asac::k_ingest_small<<<blocks, 1024, 0, st>>>(a);
        ASAC_LAUNCHED("k_ingest_small");
        return ASAC_OK;
    }
    int rc;
    const float *max_p = h->td_max;  // replay_buffer.py:296-299
    if (!buffer_empty) {
        if ((rc = asac_tree_leaf_max(h->nodes, h->capacity, h->max_p, stream)) != ASAC_OK) return rc;
        max_p = h->max_p;
    }
    if ((rc = asac_storage_write_table(&table, h->capacity, first_id % (10 * h->capacity), T, stream)) != ASAC_OK)
        return rc;
    return asac_per_add(h->nodes, h->capacity, h->store_ids, first_id, T, max_p, ignore_size, stream);
}
