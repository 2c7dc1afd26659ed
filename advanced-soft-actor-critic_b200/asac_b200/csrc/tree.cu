// GPU segment tree + prioritized-replay bookkeeping kernels (sm_100a).
//
// Replaces SumTree / PrioritizedReplayBuffer index arithmetic of the reference
// (algorithm/replay_buffer.py:145-242, 293-307, 347-354, 412-427).  Layout and numeric
// contract are documented in include/asac_b200.h and DESIGN.md §3.
#include <math.h>

#include "common.cuh"

namespace asac {

static inline int tree_levels(int64_t capacity) {
    int l = 0;
    while (((int64_t)1 << l) < capacity) ++l;
    return l;
}

static inline bool is_pow2(int64_t x) { return x > 0 && (x & (x - 1)) == 0; }

// ------------------------------------------------------------------------------------
// Block-cooperative leaf write + ancestor recompute.  One thread per updated leaf, one CTA.
// Duplicate slots: the highest thread index wins (NumPy fancy-assignment order).
// Every parent is recomputed as fp32 (left + right) from L2 (.cg loads/stores), level by
// level behind a CTA barrier, so the result is independent of thread scheduling and
// bit-identical to replay_buffer.py:176-183.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void block_tree_apply(float *nodes, int64_t capacity, int levels, int slot,
                                                 float value, bool active, bool check_dups, int *s_slot) {
    const int t = threadIdx.x;
    bool winner = active;
    if (check_dups) {
        s_slot[t] = active ? slot : -1;
        __syncthreads();
        if (active) {
            for (int u = t + 1; u < (int)blockDim.x; ++u) {
                if (s_slot[u] == slot) {
                    winner = false;
                    break;
                }
            }
        }
    }
    int64_t node = capacity + slot;
    if (winner) __stcg(&nodes[node], value);
    const float2 *pairs = reinterpret_cast<const float2 *>(nodes);
    for (int l = 0; l < levels; ++l) {
        __syncthreads();
        node >>= 1;
        if (winner) {
            float2 ch = __ldcg(pairs + node);  // children 2*node, 2*node+1
            __stcg(&nodes[node], __fadd_rn(ch.x, ch.y));
        }
    }
}

__global__ void __launch_bounds__(1024) k_tree_update(float *nodes, int64_t capacity, int levels,
                                                      const int64_t *slots, const float *p, int k) {
    __shared__ int s_slot[1024];
    const int t = threadIdx.x;
    const bool active = t < k;
    int slot = 0;
    float value = 0.f;
    if (active) {
        slot = (int)slots[t];
        value = p[t];
    }
    block_tree_apply(nodes, capacity, levels, slot, value, active, true, s_slot);
}

__global__ void k_tree_level(float *nodes, int64_t first, int64_t count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        int64_t node = first + i;
        float2 ch = reinterpret_cast<const float2 *>(nodes)[node];
        nodes[node] = __fadd_rn(ch.x, ch.y);
    }
}

// max over the leaves: 128-bit coalesced loads, warp shuffle + one atomic per CTA.
// priorities are >= 0 so the IEEE bit pattern orders like an unsigned integer.
__global__ void __launch_bounds__(256) k_leaf_max(const float *leaves, int64_t n, unsigned int *out) {
    float m = 0.f;
    const int64_t n4 = n >> 2;
    const float4 *v4 = reinterpret_cast<const float4 *>(leaves);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = __ldg(v4 + i);
        m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, leaves[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float s_m[8];
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_m[w]);
        atomicMax(out, __float_as_uint(m));
    }
}

// replay_buffer.py:195-205 for one sample; returns the data slot, *p_out = leaf priority
__device__ __forceinline__ int tree_descend(const float *nodes, int64_t capacity, int levels, double v,
                                            float *p_out) {
    const float2 *pairs = reinterpret_cast<const float2 *>(nodes);
    int64_t node = 1;
    float leaf = nodes[1];
    for (int l = 0; l < levels; ++l) {
        float2 ch = __ldg(pairs + node);
        const bool left = (v <= (double)ch.x) || (ch.y == 0.f);
        if (!left) v -= (double)ch.x;
        node = 2 * node + (left ? 0 : 1);
        leaf = left ? ch.x : ch.y;
    }
    *p_out = leaf;
    return (int)(node - capacity);
}

__device__ __forceinline__ double stratum_draw(float total, int batch, int i, double u) {
    const float seg = __fdiv_rn(total, (float)batch);           // np.float32 / int -> fp32
    const double lo = (double)i * (double)seg;                    // int64 * f32 scalar -> f64
    const double hi = (double)(i + 1) * (double)seg;
    return __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), u));       // np.random.uniform(lo, hi)
}

__global__ void k_tree_sample(const float *nodes, int64_t capacity, int levels, int batch,
                              const double *unit_uniform, uint64_t seed, const int64_t *draw_counter,
                              int32_t *out_slot, float *out_p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    double u;
    if (unit_uniform) {
        u = unit_uniform[i];
    } else {
        uint32_t r[4];
        philox4(seed, (uint64_t)draw_counter[0], (uint64_t)i, r);
        u = u01_double(r[0], r[1]);
    }
    const float total = nodes[1];
    float p;
    out_slot[i] = tree_descend(nodes, capacity, levels, stratum_draw(total, batch, i, u), &p);
    out_p[i] = p;
}

// replay_buffer.py:347-354 — one CTA, batch <= 1024
__global__ void __launch_bounds__(1024) k_per_sample(const float *nodes, int64_t capacity, int levels,
                                                     const int64_t *store_ids, int batch,
                                                     const double *unit_uniform, uint64_t seed,
                                                     int64_t *draw_counter, double *per_state,
                                                     int32_t *out_slot, int64_t *out_data_id, float *out_p,
                                                     float *out_w) {
    __shared__ float s_min[32];
    const int t = threadIdx.x;
    const bool active = t < batch;
    const float total = nodes[1];
    float p = 0.f, w = INFINITY;
    if (active) {
        double u;
        if (unit_uniform) {
            u = unit_uniform[t];
        } else {
            uint32_t r[4];
            philox4(seed, (uint64_t)draw_counter[0], (uint64_t)t, r);
            u = u01_double(r[0], r[1]);
        }
        const int slot = tree_descend(nodes, capacity, levels, stratum_draw(total, batch, t, u), &p);
        out_slot[t] = slot;
        out_data_id[t] = store_ids[slot];
        out_p[t] = p;
        w = __fdiv_rn(p, total);
    }
    float m = w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((t & 31) == 0) s_min[t >> 5] = m;
    const double beta = fmin(1.0, per_state[0] + per_state[1]);
    __syncthreads();
    if (t < 32) {
        m = (t < (int)((blockDim.x + 31) >> 5)) ? s_min[t] : INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (t == 0) s_min[0] = m;
    }
    __syncthreads();
    if (active) out_w[t] = (float)pow((double)__fdiv_rn(w, s_min[0]), -beta);
    if (t == 0) {
        per_state[0] = beta;
        if (!unit_uniform) draw_counter[0] += 1;
    }
}

// replay_buffer.py:412-427 — one CTA, k <= 1024
__global__ void __launch_bounds__(1024) k_per_update(float *nodes, int64_t capacity, int levels,
                                                     const int64_t *store_ids, const int64_t *data_ids,
                                                     const float *td, int k, float td_min, float td_max,
                                                     float alpha, int precomputed_p, double *per_state) {
    __shared__ int s_slot[1024];
    const int t = threadIdx.x;
    bool active = t < k;
    int slot = 0;
    float value = 0.f;
    int bad = 0;
    if (active) {
        const int64_t id = data_ids[t];
        slot = (int)(id & (capacity - 1));
        float e = td[t];
        if (precomputed_p) {
            value = e;
        } else {
            // np.clip keeps NaN; np.power(float32, python float) is an fp32 power with alpha
            // rounded to fp32: evaluated in fp64 and rounded once.
            float c = fminf(fmaxf(e, td_min), td_max);
            if (isnan(e)) { c = e; bad = 1; }
            value = (float)pow((double)c, (double)alpha);
        }
        active = (store_ids[slot] == id);
    }
    if (__syncthreads_or(bad)) {
        if (t == 0) per_state[3] = 1.0;  // the reference raises 'td_error has nan'
        return;
    }
    block_tree_apply(nodes, capacity, levels, slot, value, active, true, s_slot);
}

// replay_buffer.py:293-307 + :43-54 — one CTA per chunk of <= 1024 new rows
__global__ void __launch_bounds__(1024) k_per_add(float *nodes, int64_t capacity, int levels,
                                                  int64_t *store_ids, int64_t first_id, int64_t T,
                                                  int64_t chunk_begin, int chunk_len, const float *max_p,
                                                  int ignore_size) {
    __shared__ int s_slot[1];
    const int t = threadIdx.x;
    bool active = t < chunk_len;
    int slot = 0;
    float value = 0.f;
    if (active) {
        const int64_t i = chunk_begin + t;
        // consecutive ids: row i is overwritten by row i + capacity of the same call (NumPy keeps the
        // last write), so only the last `capacity` rows of a longer episode survive
        active = (i + capacity >= T);
        const int64_t id = (first_id + i) % (10 * capacity);
        slot = (int)(id & (capacity - 1));
        if (active) store_ids[slot] = id;
        value = max_p[0];
        if (ignore_size > 0 && (slot >= capacity - ignore_size || i >= T - ignore_size)) value = 0.f;
    }
    block_tree_apply(nodes, capacity, levels, slot, value, active, false, s_slot);
}

}  // namespace asac

using namespace asac;

extern "C" int asac_tree_update(float *nodes, int64_t capacity, const int64_t *slots, const float *p,
                                int64_t k, void *stream) {
    ASAC_REQUIRE(is_pow2(capacity), "asac_tree_update: capacity %lld is not a power of two", (long long)capacity);
    ASAC_REQUIRE(k >= 0, "asac_tree_update: k < 0");
    const int levels = tree_levels(capacity);
    for (int64_t off = 0; off < k; off += 1024) {
        const int n = (int)((k - off) < 1024 ? (k - off) : 1024);
        const int threads = ((n + 31) / 32) * 32;
        k_tree_update<<<1, threads, 0, (cudaStream_t)stream>>>(nodes, capacity, levels, slots + off, p + off, n);
        ASAC_LAUNCHED("k_tree_update");
    }
    return ASAC_OK;
}

extern "C" int asac_tree_rebuild(float *nodes, int64_t capacity, void *stream) {
    ASAC_REQUIRE(is_pow2(capacity), "asac_tree_rebuild: capacity is not a power of two");
    for (int64_t first = capacity >> 1; first >= 1; first >>= 1) {
        const int64_t count = first;  // nodes [first, 2*first)
        const int threads = 256;
        const int blocks = (int)((count + threads - 1) / threads);
        k_tree_level<<<blocks, threads, 0, (cudaStream_t)stream>>>(nodes, first, count);
        ASAC_LAUNCHED("k_tree_level");
    }
    return ASAC_OK;
}

extern "C" int asac_tree_leaf_max(const float *nodes, int64_t capacity, float *out, void *stream) {
    ASAC_REQUIRE(is_pow2(capacity), "asac_tree_leaf_max: capacity is not a power of two");
    ASAC_CUDA(cudaMemsetAsync(out, 0, sizeof(float), (cudaStream_t)stream));
    int64_t want = (capacity / 4 + 255) / 256;
    int blocks = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
    k_leaf_max<<<blocks, 256, 0, (cudaStream_t)stream>>>(nodes + capacity, capacity,
                                                         reinterpret_cast<unsigned int *>(out));
    ASAC_LAUNCHED("k_leaf_max");
    return ASAC_OK;
}

extern "C" int asac_tree_sample(const float *nodes, int64_t capacity, int batch, const double *unit_uniform,
                                uint64_t seed, const int64_t *draw_counter, int32_t *out_slot, float *out_p,
                                void *stream) {
    ASAC_REQUIRE(is_pow2(capacity), "asac_tree_sample: capacity is not a power of two");
    ASAC_REQUIRE(batch > 0, "asac_tree_sample: batch <= 0");
    ASAC_REQUIRE(unit_uniform || draw_counter, "asac_tree_sample: need unit_uniform or draw_counter");
    const int threads = 128;
    k_tree_sample<<<(batch + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(
        nodes, capacity, tree_levels(capacity), batch, unit_uniform, seed, draw_counter, out_slot, out_p);
    ASAC_LAUNCHED("k_tree_sample");
    return ASAC_OK;
}

extern "C" int asac_per_sample(const float *nodes, int64_t capacity, const int64_t *store_ids, int batch,
                               const double *unit_uniform, uint64_t seed, int64_t *draw_counter,
                               double *per_state, int32_t *out_slot, int64_t *out_data_id, float *out_p,
                               float *out_is_weight, void *stream) {
    ASAC_REQUIRE(is_pow2(capacity), "asac_per_sample: capacity is not a power of two");
    ASAC_REQUIRE(batch > 0 && batch <= 1024, "asac_per_sample: batch %d outside (0, 1024]", batch);
    ASAC_REQUIRE(unit_uniform || draw_counter, "asac_per_sample: need unit_uniform or draw_counter");
    const int threads = ((batch + 31) / 32) * 32;
    k_per_sample<<<1, threads, 0, (cudaStream_t)stream>>>(nodes, capacity, tree_levels(capacity), store_ids, batch,
                                                          unit_uniform, seed, draw_counter, per_state, out_slot,
                                                          out_data_id, out_p, out_is_weight);
    ASAC_LAUNCHED("k_per_sample");
    return ASAC_OK;
}

extern "C" int asac_per_update(float *nodes, int64_t capacity, const int64_t *store_ids, const int64_t *data_ids,
                               const float *td, int k, float td_min, float td_max, float alpha, int precomputed_p,
                               double *per_state, void *stream) {
    ASAC_REQUIRE(is_pow2(capacity), "asac_per_update: capacity is not a power of two");
    ASAC_REQUIRE(k > 0 && k <= 1024, "asac_per_update: k %d outside (0, 1024]", k);
    const int threads = ((k + 31) / 32) * 32;
    k_per_update<<<1, threads, 0, (cudaStream_t)stream>>>(nodes, capacity, tree_levels(capacity), store_ids, data_ids,
                                                          td, k, td_min, td_max, alpha, precomputed_p, per_state);
    ASAC_LAUNCHED("k_per_update");
    return ASAC_OK;
}

extern "C" int asac_per_add(float *nodes, int64_t capacity, int64_t *store_ids, int64_t first_id, int64_t T,
                            const float *max_p, int ignore_size, void *stream) {
    ASAC_REQUIRE(is_pow2(capacity), "asac_per_add: capacity is not a power of two");
    ASAC_REQUIRE(T > 0, "asac_per_add: T <= 0");
    const int levels = tree_levels(capacity);
    for (int64_t off = 0; off < T; off += 1024) {
        const int n = (int)((T - off) < 1024 ? (T - off) : 1024);
        const int threads = ((n + 31) / 32) * 32;
        k_per_add<<<1, threads, 0, (cudaStream_t)stream>>>(nodes, capacity, levels, store_ids, first_id, T, off, n,
                                                           max_p, ignore_size);
        ASAC_LAUNCHED("k_per_add");
    }
    return ASAC_OK;
}
