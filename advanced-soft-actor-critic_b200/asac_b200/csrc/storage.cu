// HBM-resident transition storage: ring writes, window gather fused with the learner-side
// padding rule, and guarded write-backs (sm_100a).
//
// Replaces DataStorage (algorithm/replay_buffer.py:21-142), the window gather of
// _prefetch_loop (:356-362), update_transitions (:429-434) and the padding block of
// SAC_Base._sample_from_replay_buffer (algorithm/sac_base.py:2435-2453).
//
// Layout: one contiguous [capacity, row_bytes] array per stored key (struct of arrays) plus
// the int64 `_id` column.  All kernels move whole rows with the widest unit the row size
// and pointer alignment allow (16 / 4 / 1 bytes), consecutive lanes on consecutive units.
#include "common.cuh"

namespace asac {

__device__ __forceinline__ int copy_unit(const void *a, const void *b, int64_t bytes) {
    const uintptr_t bits = (uintptr_t)a | (uintptr_t)b | (uintptr_t)bytes;
    return (bits & 15) == 0 ? 16 : ((bits & 3) == 0 ? 4 : 1);
}

__device__ __forceinline__ void lane_copy(void *dst, const void *src, int64_t bytes, int lane, int nlanes) {
    const int unit = copy_unit(dst, src, bytes);
    if (unit == 16) {
        const int4 *s = reinterpret_cast<const int4 *>(src);
        int4 *d = reinterpret_cast<int4 *>(dst);
        for (int64_t i = lane; i < (bytes >> 4); i += nlanes) d[i] = s[i];
    } else if (unit == 4) {
        const int32_t *s = reinterpret_cast<const int32_t *>(src);
        int32_t *d = reinterpret_cast<int32_t *>(dst);
        for (int64_t i = lane; i < (bytes >> 2); i += nlanes) d[i] = s[i];
    } else {
        const uint8_t *s = reinterpret_cast<const uint8_t *>(src);
        uint8_t *d = reinterpret_cast<uint8_t *>(dst);
        for (int64_t i = lane; i < bytes; i += nlanes) d[i] = s[i];
    }
}

// one warp per new row
__global__ void k_storage_write_rows(uint8_t *ring, int64_t capacity, int64_t first_id, const uint8_t *rows,
                                     int64_t T, int64_t row_bytes) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= T) return;
    const int64_t id = (first_id + w) % (10 * capacity);
    const int64_t slot = id & (capacity - 1);
    lane_copy(ring + slot * row_bytes, rows + w * row_bytes, row_bytes, lane, 32);
}

// one warp per new row, all columns of the episode in one launch (rows staged column after column
// in one device buffer by a single H2D copy)
__global__ void k_storage_write_table(AsacWriteTable table, int64_t capacity, int64_t first_id, int64_t T) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= T) return;
    const int64_t id = (first_id + w) % (10 * capacity);
    const int64_t slot = id & (capacity - 1);
    for (int c = 0; c < table.n_columns; ++c) {
        const int64_t rb = table.col[c].row_bytes;
        lane_copy(reinterpret_cast<uint8_t *>(table.col[c].ring) + slot * rb,
                  reinterpret_cast<const uint8_t *>(table.col[c].rows) + w * rb, rb, lane, 32);
    }
}

// one warp per (batch element, time step) window row
__global__ void k_storage_gather(AsacColumnTable table, int64_t capacity, const int64_t *data_ids, int batch,
                                 int prev_n, int L, const float *padding_action, uint8_t *out_padding_mask) {
    pdl_wait();
    pdl_trigger();
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (int64_t)batch * L) return;
    const int b = (int)(w / L), t = (int)(w % L);
    const int64_t id = data_ids[b];
    const int64_t slot = (id + t - prev_n) & (capacity - 1);
    const int64_t anchor = id & (capacity - 1);
    const int32_t *index_ring = reinterpret_cast<const int32_t *>(table.col[table.index_column].ring);
    // sac_base.py:2441-2443
    const bool invalid = (t != prev_n) && ((index_ring[slot] - index_ring[anchor]) != (t - prev_n));
    if (lane == 0) out_padding_mask[w] = invalid ? 1 : 0;
    for (int c = 0; c < table.n_columns; ++c) {
        const AsacColumn col = table.col[c];
        uint8_t *dst = reinterpret_cast<uint8_t *>(col.out) + w * (int64_t)col.out_stride + col.out_offset;
        const uint8_t *src = reinterpret_cast<const uint8_t *>(col.ring) + slot * (int64_t)col.row_bytes;
        if (!invalid || col.role == ASAC_ROLE_COPY) {
            lane_copy(dst, src, col.row_bytes, lane, 32);
            continue;
        }
        switch (col.role) {
            case ASAC_ROLE_INDEX:
                if (lane == 0) *reinterpret_cast<int32_t *>(dst) = -1;
                break;
            case ASAC_ROLE_ACTION:
                for (int i = lane; i < col.row_bytes / 4; i += 32) reinterpret_cast<float *>(dst)[i] = padding_action[i];
                break;
            case ASAC_ROLE_REWARD:
                if (lane == 0) *reinterpret_cast<float *>(dst) = 0.f;
                break;
            case ASAC_ROLE_DONE:
                if (lane == 0) *dst = 1;
                break;
            case ASAC_ROLE_MU_PROB:
                for (int i = lane; i < col.row_bytes / 4; i += 32) reinterpret_cast<float *>(dst)[i] = 1.f;
                break;
            default:  // ASAC_ROLE_HIDDEN
                for (int i = lane; i < col.row_bytes; i += 32) dst[i] = 0;
                break;
        }
    }
}

// one warp per (batch element, write-back row)
__global__ void k_storage_scatter(uint8_t *ring, int64_t capacity, const int64_t *store_ids,
                                  const int64_t *data_ids, int batch, int first_offset, int n_rows,
                                  const uint8_t *rows, int64_t row_bytes, int64_t rows_b_stride,
                                  const uint8_t *padding_mask, int64_t mask_b_stride) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (int64_t)batch * n_rows) return;
    const int b = (int)(w / n_rows), t = (int)(w % n_rows);
    if (padding_mask && padding_mask[b * mask_b_stride + t]) return;  // sac_base.py:2596,2605
    const int64_t id = data_ids[b] + first_offset + t;
    const int64_t slot = id & (capacity - 1);
    if (store_ids[slot] != id) return;  // replay_buffer.py:431-434
    // Duplicate targets: NumPy's fancy assignment keeps the LAST write in flat (b, t) order
    // (pointers are stacked on axis 1 and flattened, sac_base.py:2590-2591, 2600-2601).  A row
    // yields when a later, unpadded (b', t') aims at the same id: data_ids[b'] + t' == data_ids[b] + t.
    bool loses = false;
    for (int b2 = b + lane; b2 < batch; b2 += 32) {
        const int64_t t2 = data_ids[b] + t - data_ids[b2];
        if (t2 < 0 || t2 >= n_rows) continue;
        if (b2 == b && t2 <= t) continue;
        if (padding_mask && padding_mask[b2 * mask_b_stride + t2]) continue;
        loses = true;
    }
    if (__any_sync(0xffffffffu, loses)) return;
    lane_copy(ring + slot * row_bytes, rows + (b * rows_b_stride + t) * row_bytes, row_bytes, lane, 32);
}

}  // namespace asac

using namespace asac;

static inline bool pow2(int64_t x) { return x > 0 && (x & (x - 1)) == 0; }

extern "C" int asac_storage_write_rows(void *ring, int64_t capacity, int64_t first_id, const void *rows, int64_t T,
                                       int64_t row_bytes, void *stream) {
    ASAC_REQUIRE(pow2(capacity), "asac_storage_write_rows: capacity is not a power of two");
    if (T <= 0 || row_bytes <= 0) return ASAC_OK;
    const int threads = 256;
    const int64_t blocks = (T * 32 + threads - 1) / threads;
    k_storage_write_rows<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<uint8_t *>(ring), capacity, first_id, reinterpret_cast<const uint8_t *>(rows), T, row_bytes);
    ASAC_LAUNCHED("k_storage_write_rows");
    return ASAC_OK;
}

extern "C" int asac_storage_write_table(const AsacWriteTable *table_host, int64_t capacity, int64_t first_id,
                                        int64_t T, void *stream) {
    ASAC_REQUIRE(pow2(capacity), "asac_storage_write_table: capacity is not a power of two");
    ASAC_REQUIRE(table_host && table_host->n_columns > 0 && table_host->n_columns <= ASAC_MAX_COLUMNS,
                 "asac_storage_write_table: bad column table");
    if (T <= 0) return ASAC_OK;
    const int threads = 256;
    const int64_t blocks = (T * 32 + threads - 1) / threads;
    k_storage_write_table<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(*table_host, capacity, first_id, T);
    ASAC_LAUNCHED("k_storage_write_table");
    return ASAC_OK;
}

extern "C" int asac_storage_gather(const AsacColumnTable *table_host, int64_t capacity, const int64_t *data_ids,
                                   int batch, int prev_n, int post_n, const float *padding_action,
                                   uint8_t *out_padding_mask, void *stream) {
    ASAC_REQUIRE(pow2(capacity), "asac_storage_gather: capacity is not a power of two");
    ASAC_REQUIRE(table_host && table_host->n_columns > 0 && table_host->n_columns <= ASAC_MAX_COLUMNS,
                 "asac_storage_gather: bad column table");
    ASAC_REQUIRE(table_host->index_column >= 0 && table_host->index_column < table_host->n_columns &&
                     table_host->col[table_host->index_column].row_bytes == 4,
                 "asac_storage_gather: index column must be int32");
    ASAC_REQUIRE(batch > 0 && prev_n >= 0 && post_n >= 0, "asac_storage_gather: bad window");
    const int L = prev_n + 1 + post_n;
    const int threads = 256;
    const int64_t blocks = ((int64_t)batch * L * 32 + threads - 1) / threads;
    ASAC_CUDA(launch_ex(k_storage_gather, dim3((unsigned)blocks), dim3(threads), 0, (cudaStream_t)stream, 0, true,
                        *table_host, capacity, data_ids, batch, prev_n, L, padding_action, out_padding_mask));
    ASAC_LAUNCHED("k_storage_gather");
    return ASAC_OK;
}

extern "C" int asac_storage_scatter(void *ring, int64_t capacity, const int64_t *store_ids, const int64_t *data_ids,
                                    int batch, int first_offset, int n_rows, const void *rows, int64_t row_bytes,
                                    int64_t rows_b_stride, const uint8_t *padding_mask, int64_t mask_b_stride,
                                    void *stream) {
    ASAC_REQUIRE(pow2(capacity), "asac_storage_scatter: capacity is not a power of two");
    if (batch <= 0 || n_rows <= 0 || row_bytes <= 0) return ASAC_OK;
    const int threads = 256;
    const int64_t blocks = ((int64_t)batch * n_rows * 32 + threads - 1) / threads;
    k_storage_scatter<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<uint8_t *>(ring), capacity, store_ids, data_ids, batch, first_offset, n_rows,
        reinterpret_cast<const uint8_t *>(rows), row_bytes, rows_b_stride, padding_mask, mask_b_stride);
    ASAC_LAUNCHED("k_storage_scatter");
    return ASAC_OK;
}
