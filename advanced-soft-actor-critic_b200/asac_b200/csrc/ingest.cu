// Native episode ingest: PrioritizedReplayBuffer.add for host-resident episodes as ONE call.
//
// Replaces the per-key NumPy assignments of DataStorage.add (algorithm/replay_buffer.py:30-56)
// and the max-priority insert of PrioritizedReplayBuffer.add (:293-315) on the actor -> learner
// boundary: the columns of the episode are packed into a pinned staging buffer (ring of 4, each
// guarded by an event), cross PCIe in one cudaMemcpyAsync and are written into the rings by
// k_storage_write_table; k_leaf_max + k_per_add give the new rows the current maximum priority.
// The handle owns only its staging buffers and events; rings, tree and ids stay the caller's.
#include <string.h>

#include <new>

#include "common.cuh"

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
