// GPU segment tree + prioritized-replay bookkeeping kernels (sm_100a).
//
// Replaces SumTree / PrioritizedReplayBuffer index arithmetic of the reference
// (algorithm/replay_buffer.py:145-242, 293-307, 347-354, 412-427).  Layout and numeric
// contract are documented in include/asac_b200.h and DESIGN.md §3.
#include <math.h>

#include "common.cuh"
#include "tree_apply.cuh"

namespace asac {

__global__ void __launch_bounds__(1024) k_tree_update(float *nodes, int64_t capacity, int levels,
                                                      const int64_t *slots, const float *p, int k) {
    __shared__ TreeApplySmem s_apply;
    const int t = threadIdx.x;
    const bool active = t < k;
    int slot = 0;
    float value = 0.f;
    if (active) {
        slot = (int)slots[t];
        value = p[t];
    }
    block_tree_apply(nodes, capacity, levels, slot, value, active, s_apply);
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

// One decision of replay_buffer.py:195-205: go left iff v <= left || right == 0, else v -= left
// (float64 compare / subtract against fp32 nodes, NumPy promotion).
__device__ __forceinline__ int descend_step(float left, float right, double &v, float &leaf) {
    const bool go_left = (v <= (double)left) || (right == 0.f);
    if (!go_left) v -= (double)left;
    leaf = go_left ? left : right;
    return go_left ? 0 : 1;
}

// K levels below `node` with ONE memory round trip: the 2^j descendants of `node` on sub-level j
// are the contiguous, aligned run nodes[node << j .. (node << j) + 2^j), so all K sub-levels are
// requested before the first comparison (speculative loads: 2^(K+1) - 2 floats, one is used per
// level).  The decisions themselves are replay_buffer.py:195-205 unchanged.
template <int K>
__device__ __forceinline__ int64_t descend_block(const float *__restrict__ nodes, int64_t node, double &v,
                                                 float &leaf) {
    float sub[(2 << K) - 2];  // sub-level j (1..K) at offset 2^j - 2
    {
        const float2 c = __ldg(reinterpret_cast<const float2 *>(nodes + (node << 1)));
        sub[0] = c.x; sub[1] = c.y;
    }
#pragma unroll
    for (int j = 2; j <= K; ++j) {
#pragma unroll
        for (int q = 0; q < (1 << j) / 4; ++q) {
            const int o = (1 << j) - 2 + 4 * q;
            const float4 c = __ldg(reinterpret_cast<const float4 *>(nodes + (node << j)) + q);
            sub[o] = c.x; sub[o + 1] = c.y; sub[o + 2] = c.z; sub[o + 3] = c.w;
        }
    }
    int path = 0;  // decisions so far, oldest in the most significant bit
#pragma unroll
    for (int j = 1; j <= K; ++j) {
        // the pair under `path` on sub-level j: halve the candidate run once per earlier decision
        float cand[1 << K];
#pragma unroll
        for (int i = 0; i < (1 << j); ++i) cand[i] = sub[(1 << j) - 2 + i];
#pragma unroll
        for (int b = 0; b < j - 1; ++b) {
            const bool bit = (path >> (j - 2 - b)) & 1;
            const int half = (1 << j) >> (b + 1);
#pragma unroll
            for (int i = 0; i < half; ++i) cand[i] = bit ? cand[i + half] : cand[i];
        }
        path = (path << 1) | descend_step(cand[0], cand[1], v, leaf);
    }
    return (node << K) + path;
}

// replay_buffer.py:195-205 for one sample; returns the data slot, *p_out = leaf priority.
// `top` (shared memory, may be null) mirrors nodes[0 .. 2^(top_levels+1)), i.e. levels 0..top_levels.
__device__ __forceinline__ int tree_descend(const float *__restrict__ nodes, int64_t capacity, int levels, double v,
                                            float *p_out, const float *top = nullptr, int top_levels = 0) {
    int64_t node = 1;
    float leaf = top ? top[1] : nodes[1];
    int l = 0;
    for (; l < top_levels; ++l) node = 2 * node + descend_step(top[2 * node], top[2 * node + 1], v, leaf);
    // K = 3 (one 8-byte and three 16-byte requests): ptxas keeps exactly this group ahead of the
    // first decision; with K = 4 it sinks the 64-byte sub-level below it, i.e. a second round trip
    for (; l + 3 <= levels; l += 3) node = descend_block<3>(nodes, node, v, leaf);
    const int rest = levels - l;
    if (rest == 2) node = descend_block<2>(nodes, node, v, leaf);
    else if (rest == 1) node = descend_block<1>(nodes, node, v, leaf);
    *p_out = leaf;
    return (int)(node - capacity);
}

// levels of the tree mirrored in shared memory by the sampling kernels (nodes[0 .. 2^(n+1)))
constexpr int TREE_TOP_LEVELS = 10;
__device__ __forceinline__ int stage_tree_top(const float *__restrict__ nodes, int levels, float *top) {
    const int n = levels < TREE_TOP_LEVELS ? levels : TREE_TOP_LEVELS;
    const float4 *src = reinterpret_cast<const float4 *>(nodes);
    float4 *dst = reinterpret_cast<float4 *>(top);
    if (n >= 1) {
        for (int i = threadIdx.x; i < ((2 << n) >> 2); i += blockDim.x) dst[i] = __ldg(src + i);
    } else {
        if (threadIdx.x < 2) top[threadIdx.x] = __ldg(nodes + threadIdx.x);
    }
    __syncthreads();
    return n;
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
    __shared__ __align__(16) float s_top[2 << TREE_TOP_LEVELS];
    const int top_levels = stage_tree_top(nodes, levels, s_top);
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
    const float total = s_top[1];
    float p;
    out_slot[i] = tree_descend(nodes, capacity, levels, stratum_draw(total, batch, i, u), &p, s_top, top_levels);
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
    __shared__ __align__(16) float s_top[2 << TREE_TOP_LEVELS];
    pdl_wait();
    pdl_trigger();
    const int top_levels = stage_tree_top(nodes, levels, s_top);
    const int t = threadIdx.x;
    const bool active = t < batch;
    const float total = s_top[1];
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
        const int slot = tree_descend(nodes, capacity, levels, stratum_draw(total, batch, t, u), &p, s_top,
                                      top_levels);
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
        per_state[2] = (double)s_min[0];  // smallest sampling probability of this batch (sharded replay: asac_per_shard_weights)
        if (!unit_uniform) draw_counter[0] += 1;
    }
}

// Sharded replay (one tree per GPU, B draws each): a transition of shard r is drawn with probability
// p / total_r, so its importance weight is ((p / total_r) / min)^-beta with the minimum taken over the batches
// of ALL shards — the single-buffer rule of replay_buffer.py:352-354 applied to the union of the draws.
// global_min[0] is that minimum (MIN all-reduce of per_state[2]); total_r is the shard's own root.
__global__ void __launch_bounds__(1024) k_per_shard_weights(const float *nodes, int batch, const float *p,
                                                            const double *per_state, const double *global_min,
                                                            float *out_w) {
    const int t = threadIdx.x;
    if (t >= batch) return;
    const float total = nodes[1];
    const float w = __fdiv_rn(p[t], total);
    out_w[t] = (float)pow((double)__fdiv_rn(w, (float)global_min[0]), -per_state[0]);
}

// replay_buffer.py:412-427 — one CTA, k <= 1024
__global__ void __launch_bounds__(1024) k_per_update(float *nodes, int64_t capacity, int levels,
                                                     const int64_t *store_ids, const int64_t *data_ids,
                                                     const float *td, int k, float td_min, float td_max,
                                                     float alpha, int precomputed_p, double *per_state) {
    __shared__ TreeApplySmem s_apply;
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
            value = td_to_priority(e, td_min, td_max, alpha, &bad);
        }
        active = (store_ids[slot] == id);
    }
    if (__syncthreads_or(bad)) {
        if (t == 0) per_state[3] = 1.0;  // the reference raises 'td_error has nan'
        return;
    }
    block_tree_apply(nodes, capacity, levels, slot, value, active, s_apply);
}

// replay_buffer.py:293-307 + :43-54 — one CTA per chunk of <= 1024 new rows
__global__ void __launch_bounds__(1024) k_per_add(float *nodes, int64_t capacity, int levels,
                                                  int64_t *store_ids, int64_t first_id, int64_t T,
                                                  int64_t chunk_begin, int chunk_len, const float *max_p,
                                                  int ignore_size) {
    __shared__ TreeApplySmem s_apply;
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
    block_tree_apply(nodes, capacity, levels, slot, value, active, s_apply);
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
    const int threads = batch >= 8192 ? 1024 : 128;  // bulk: amortise the 8 KB tree-top stage over more samples
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
    ASAC_CUDA(launch_ex(k_per_sample, dim3(1), dim3(threads), 0, (cudaStream_t)stream, 0, true, nodes, capacity,
                        tree_levels(capacity), store_ids, batch, unit_uniform, seed, draw_counter, per_state, out_slot,
                        out_data_id, out_p, out_is_weight));
    ASAC_LAUNCHED("k_per_sample");
    return ASAC_OK;
}

extern "C" int asac_per_shard_weights(const float *nodes, int batch, const float *p, const double *per_state,
                                      const double *global_min, float *out_is_weight, void *stream) {
    ASAC_REQUIRE(nodes && p && per_state && global_min && out_is_weight, "asac_per_shard_weights: null pointer");
    ASAC_REQUIRE(batch > 0 && batch <= 1024, "asac_per_shard_weights: batch %d outside (0, 1024]", batch);
    k_per_shard_weights<<<1, ((batch + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(nodes, batch, p, per_state, global_min,
                                                                                 out_is_weight);
    ASAC_LAUNCHED("k_per_shard_weights");
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
