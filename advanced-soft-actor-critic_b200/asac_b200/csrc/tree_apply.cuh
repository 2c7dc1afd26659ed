// Block-cooperative sum-tree update shared by the replay kernels (tree.cu) and the fused step
// epilogue (sac.cu).  See replay_buffer.py:172-183 of the reference.
#pragma once
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
//
// Every touched parent becomes fp32 (left + right) exactly as replay_buffer.py:176-183, but the
// D dependent global round trips of a level-by-level walk are replaced by ONE: the siblings of
// every node on a leaf's path are prefetched at once (they are only used where the sibling
// subtree holds no updated leaf), and the level-by-level recomputation runs on the union of
// the paths in shared memory.  The updated leaves are kept as a sorted, doubly linked list of
// "alive" entries; on each level two alive siblings merge (the left one survives and adds the
// right one's value), a lone child adds its prefetched sibling.  Results go back with
// fire-and-forget stores.  Stratified samples arrive sorted by leaf, so the bitonic sort below
// only runs for caller-supplied id lists.
// ------------------------------------------------------------------------------------
constexpr int TREE_CHUNK = 8;  // levels per register set of prefetched siblings

struct TreeApplySmem {
    unsigned long long key[1024];
    int node[1024];
    float val[1024];
    short next[1024], prev[1024];
    int warp_sum[32];
};

__device__ __forceinline__ void block_tree_apply(float *nodes, int64_t capacity, int levels, int slot, float value,
                                                 bool active, TreeApplySmem &sm) {
    const int t = threadIdx.x, nthr = blockDim.x, lane = t & 31, warp = t >> 5;
    // ---- sort key (slot, thread): inactive entries sort to the end
    const unsigned long long mykey =
        active ? (((unsigned long long)(unsigned)slot << 32) | (unsigned)t) : 0xFFFFFFFFFFFFFFFFull;
    sm.key[t] = mykey;
    sm.val[t] = value;
    __syncthreads();
    const bool unsorted = t > 0 && sm.key[t - 1] > mykey;
    if (__syncthreads_or(unsorted)) {
        int n2 = 32;
        while (n2 < nthr) n2 <<= 1;
        for (int i = nthr + t; i < n2; i += nthr) sm.key[i] = 0xFFFFFFFFFFFFFFFFull;
        __syncthreads();
        for (int k = 2; k <= n2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = t; i < n2; i += nthr) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = sm.key[i], b = sm.key[ixj];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) { sm.key[i] = b; sm.key[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        }
    }
    // ---- winners: the last entry of each run of equal slots (= the highest thread index)
    const unsigned long long k0 = sm.key[t];
    const bool valid = k0 != 0xFFFFFFFFFFFFFFFFull;
    const unsigned long long k1 = (t + 1 < nthr) ? sm.key[t + 1] : 0xFFFFFFFFFFFFFFFFull;
    const bool winner = valid && (k1 == 0xFFFFFFFFFFFFFFFFull || (k1 >> 32) != (k0 >> 32));
    const float wval = valid ? sm.val[(int)(k0 & 0xFFFFFFFFu)] : 0.f;
    // ---- dense rank of the winners (block exclusive scan)
    const unsigned ballot = __ballot_sync(0xffffffffu, winner);
    if (lane == 0) sm.warp_sum[warp] = __popc(ballot);
    __syncthreads();  // also: every thread has read its key / value before the arrays are reused
    int base = 0, total = 0;
    for (int w = 0; w < (nthr >> 5); ++w) {
        const int c = sm.warp_sum[w];
        if (w < warp) base += c;
        total += c;
    }
    const int d = base + __popc(ballot & ((1u << lane) - 1u));
    __syncthreads();
    if (winner) {
        sm.node[d] = (int)(capacity + (int64_t)(k0 >> 32));
        sm.val[d] = wval;
        sm.next[d] = (short)(d + 1 < total ? d + 1 : -1);
        sm.prev[d] = (short)(d - 1);
    }
    __syncthreads();
    // ---- entry t (< total) walks its leaf's path, TREE_CHUNK levels per register set; the next
    // chunk's siblings are requested before the current chunk is consumed (one exposed round trip)
    bool alive = t < total;
    const int leaf_node = alive ? sm.node[t] : 1;
    int node = leaf_node;
    if (alive) __stcg(nodes + node, sm.val[t]);
    float cur[TREE_CHUNK], nxt[TREE_CHUNK];
#pragma unroll
    for (int j = 0; j < TREE_CHUNK; ++j) cur[j] = (alive && j < levels) ? __ldcg(nodes + ((leaf_node >> j) ^ 1)) : 0.f;
#pragma unroll 1
    for (int l0 = 0; l0 < levels; l0 += TREE_CHUNK) {
#pragma unroll
        for (int j = 0; j < TREE_CHUNK; ++j) {
            const int l = l0 + TREE_CHUNK + j;
            nxt[j] = (alive && l < levels) ? __ldcg(nodes + ((leaf_node >> l) ^ 1)) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < TREE_CHUNK; ++j) {
            if (l0 + j < levels) {  // block-uniform
                float nv = 0.f;
                int merged = -1;
                bool dies = false;
                if (alive) {
                    const float mine = sm.val[t];
                    if ((node & 1) == 0) {
                        const int u = sm.next[t];
                        if (u >= 0 && sm.node[u] == node + 1) {
                            merged = u;
                            nv = __fadd_rn(mine, sm.val[u]);
                        } else {
                            nv = __fadd_rn(mine, cur[j]);
                        }
                    } else {
                        const int q = sm.prev[t];
                        if (q >= 0 && sm.node[q] == node - 1) dies = true;
                        else nv = __fadd_rn(cur[j], mine);
                    }
                }
                __syncthreads();
                if (alive && !dies) {
                    node >>= 1;
                    sm.node[t] = node;
                    sm.val[t] = nv;
                    __stcg(nodes + node, nv);
                    if (merged >= 0) {
                        const int w = sm.next[merged];
                        sm.next[t] = (short)w;
                        if (w >= 0) sm.prev[w] = (short)t;
                    }
                }
                if (dies) alive = false;
                __syncthreads();
            }
        }
#pragma unroll
        for (int j = 0; j < TREE_CHUNK; ++j) cur[j] = nxt[j];
    }
}

// PrioritizedReplayBuffer.update's priority (replay_buffer.py:415-422): clip(td, min, max) ** alpha.
// np.clip keeps NaN (*bad is set); np.power(float32, python float) is an fp32 power with alpha
// rounded to fp32: evaluated in fp64 and rounded once.
__device__ __forceinline__ float td_to_priority(float td, float td_min, float td_max, float alpha, int *bad) {
    float c = fminf(fmaxf(td, td_min), td_max);
    if (isnan(td)) { c = td; *bad = 1; }
    return (float)pow((double)c, (double)alpha);
}

}  // namespace asac
