// Stock-net forward on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32 with
// the accumulator in tensor memory, 3xTF32 split operands for fp32-level accuracy.
//
// Replaces, for large row counts, the chain of nn.Linear + exact-erf GELU + residual launches
// of LinearLayers (algorithm/nn_models/layers/linear_layers.py:24-119) behind ModelQ / ModelPolicy
// forwards (nn_models/q.py:74-91, nn_models/policy.py:150-170) — the evaluation the actor side
// and the large-batch value passes spend their time in.
//
// One persistent CTA per SM walks 128-row tiles, TWO tile groups of 8 warps each when shared
// memory allows: each group owns its A operands, its TMEM accumulator columns, its mbarrier and
// a named barrier, and runs the layers of its tile independently of the other group, so the
// tensor pipe works on one group's MMAs while the CUDA cores run the other group's epilogue
// (v1 ran MMA and epilogue of a single tile back to back: tensor pipe 7.8 % active).
//   * operands live in shared memory in the UMMA K-major, no-swizzle "core matrix" layout
//     (8 rows x 16 bytes contiguous; consecutive K chunks 128 B apart, 8-row groups SBO apart),
//     each as a (hi, lo) pair: hi = x with the 13 low mantissa bits cleared (exact in tf32),
//     lo = x - hi.  D = A_hi.B_lo + A_lo.B_hi + A_hi.B_hi accumulates in fp32 in TMEM, which
//     restores ~21 mantissa bits per product (plain kind::tf32 keeps 10 and cannot meet the
//     1e-5 parity bound of the north star);
//   * the weights of every layer are split and laid out once per CTA;
//   * one elected thread issues the 3 x K/8 MMAs of a layer and commits them to an mbarrier;
//   * the epilogue maps TMEM lane = row to one thread (two warps per 32-lane quarter, 32
//     accumulator columns each): tcgen05.ld -> bias + exact-erf GELU + residual -> split ->
//     16-byte stores straight into the next layer's A operand; the head is one more MMA
//     (N = 16, zero padded) whose epilogue writes the rows of the result.
#include <math.h>

#include "common.cuh"
#include "mlp_tile.cuh"
#include "tc_engine.cuh"

namespace asac {

constexpr int TC_ROWS = 128;     // UMMA_M
constexpr int TC_THREADS = 256;  // per tile group: 8 warps, 2 per TMEM lane quarter
constexpr int TC_HEAD_N = 16;    // padded head width (UMMA_N % 16 == 0 for M = 128)
constexpr int TC_PREFETCH = 2;   // float4 registers per thread holding the next tile's input rows

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------- shared-memory plan
struct TcPlan {
    int K0p;                  // first-layer K padded to 8
    int off_a_hi[2], off_a_lo[2];   // per tile group: [128, max(K0p, H)] activations
    int off_w[ASAC_MAX_DEPTH + 1][2];  // per layer (head last): hi, lo
    int off_bias;             // depth * H + TC_HEAD_N
    int off_misc;             // mbarrier, tmem address
    int total;                // floats
};
__host__ __device__ __forceinline__ TcPlan tc_plan(const NetShape &s, int groups) {
    TcPlan p;
    const int H = s.hidden;
    p.K0p = round_up(s.in_dim, 8);
    const int ka = p.K0p > H ? p.K0p : H;
    int o = 0;
    for (int g = 0; g < 2; ++g) {
        p.off_a_hi[g] = o; o += g < groups ? TC_ROWS * ka : 0;
        p.off_a_lo[g] = o; o += g < groups ? TC_ROWS * ka : 0;
    }
    for (int l = 0; l < s.depth; ++l) {
        const int K = l == 0 ? p.K0p : H;
        p.off_w[l][0] = o; o += H * K;
        p.off_w[l][1] = o; o += H * K;
    }
    p.off_w[s.depth][0] = o; o += TC_HEAD_N * H;
    p.off_w[s.depth][1] = o; o += TC_HEAD_N * H;
    p.off_bias = o; o += s.depth * H + TC_HEAD_N;
    p.off_misc = o; o += 8;   // 2 mbarriers (16 B), tmem address
    p.total = o;
    return p;
}

struct TcArgs {
    const float *params, *x;
    float *out;
    NetShape s;
    int64_t rows;
};

__device__ __forceinline__ void group_sync(int group) {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(TC_THREADS) : "memory");
}

template <int GROUPS>
__global__ void __launch_bounds__(TC_THREADS *GROUPS, 1) k_mlp_forward_tc(const TcArgs a) {
    extern __shared__ __align__(128) float smem_tc[];
    float *sm = smem_tc;
    const NetShape s = a.s;
    const int H = s.hidden, d = s.depth, O = s.out_dim;
    const TcPlan pl = tc_plan(s, GROUPS);
    const int group = threadIdx.x / TC_THREADS;            // tile group of this thread
    const int tid = threadIdx.x % TC_THREADS, warp = tid >> 5, lane = tid & 31;  // within the group
    float *a_hi = sm + pl.off_a_hi[group], *a_lo = sm + pl.off_a_lo[group], *bias = sm + pl.off_bias;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + pl.off_misc) + group;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + pl.off_misc + 4);
    // per group 128 columns: trunk accumulator at +0, head accumulator at +64
    constexpr uint32_t TMEM_COLS = 128 * GROUPS;

    // ---- one-time setup: TMEM, mbarrier, split weights in the UMMA layout
    if (threadIdx.x < 32) tmem_alloc(tmem_slot, TMEM_COLS);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    for (int l = 0; l <= d; ++l) {
        const int K = l == 0 ? s.in_dim : H, Kp = l == 0 ? pl.K0p : H;
        const int N = l == d ? O : H, Np = l == d ? TC_HEAD_N : H;
        const float *W = a.params + net_w_off(s, l);
        float *w_hi = sm + pl.off_w[l][0], *w_lo = sm + pl.off_w[l][1];
        for (int i = threadIdx.x; i < Np * Kp; i += TC_THREADS * GROUPS) {
            const int n = i / Kp, k = i - n * Kp;
            const float w = (n < N && k < K) ? __ldg(W + (int64_t)n * K + k) : 0.f;
            float hi, lo;
            split_tf32(w, hi, lo);
            const int off = umma_off(n, k, Kp);
            w_hi[off] = hi;
            w_lo[off] = lo;
        }
        const float *b = a.params + net_b_off(s, l);
        for (int i = threadIdx.x; i < Np; i += TC_THREADS * GROUPS) bias[l * H + i] = i < N ? __ldg(b + i) : 0.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot + (uint32_t)group * 128u;
    const uint32_t idesc_trunk = umma_idesc_tf32(TC_ROWS, H), idesc_head = umma_idesc_tf32(TC_ROWS, TC_HEAD_N);
    const uint32_t a_hi_addr = smem_u32(a_hi), a_lo_addr = smem_u32(a_lo);

    // this thread's accumulator slice: TMEM lane = row, 32 columns
    const int row = (warp & 3) * 32 + lane;
    const int col0 = (warp >> 2) * 32;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t phase = 0;

    const int64_t n_tiles = (a.rows + TC_ROWS - 1) / TC_ROWS;
    // input prefetch registers: TC_PREFETCH float4 per thread cover a tile of up to 16 input columns
    const bool prefetch = (TC_ROWS * s.in_dim <= TC_PREFETCH * 4 * TC_THREADS) && ((((uintptr_t)a.x) & 15) == 0);
    float4 nx[TC_PREFETCH];
    auto load_tile = [&](int64_t t) {
#pragma unroll
        for (int q = 0; q < TC_PREFETCH; ++q) {
            nx[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int e = 4 * (tid + TC_THREADS * q);
            if (t < n_tiles && e < TC_ROWS * s.in_dim) {
                const int64_t g = t * TC_ROWS * s.in_dim + e, total = a.rows * s.in_dim;
                if (g + 3 < total) {
                    nx[q] = __ldg(reinterpret_cast<const float4 *>(a.x + g));
                } else {
                    if (g < total) nx[q].x = __ldg(a.x + g);
                    if (g + 1 < total) nx[q].y = __ldg(a.x + g + 1);
                    if (g + 2 < total) nx[q].z = __ldg(a.x + g + 2);
                }
            }
        }
    };
    if (prefetch) load_tile((int64_t)blockIdx.x * GROUPS + group);
    for (int64_t tile = (int64_t)blockIdx.x * GROUPS + group; tile < n_tiles; tile += (int64_t)gridDim.x * GROUPS) {
        const int64_t r0 = tile * TC_ROWS;
        // ---- stage the tile's input rows (split) as the first A operand.  Narrow inputs are
        // prefetched: the NEXT tile's rows are requested into registers before this tile's layers
        // run, so their global-memory latency hides behind a whole tile of work.
        {
            const int K = s.in_dim, Kp = pl.K0p;
            if (prefetch) {
#pragma unroll
                for (int q = 0; q < TC_PREFETCH; ++q) {
                    const float xv[4] = {nx[q].x, nx[q].y, nx[q].z, nx[q].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int e = 4 * (tid + TC_THREADS * q) + j;
                        if (e < TC_ROWS * K) {
                            const int r = e / K, k = e - r * K;
                            float hi, lo;
                            split_tf32(xv[j], hi, lo);
                            const int off = umma_off(r, k, Kp);
                            a_hi[off] = hi;
                            a_lo[off] = lo;
                        }
                    }
                }
                for (int i = tid; i < TC_ROWS * (Kp - K); i += TC_THREADS) {
                    const int r = i / (Kp - K), k = K + i - r * (Kp - K);
                    const int off = umma_off(r, k, Kp);
                    a_hi[off] = 0.f;
                    a_lo[off] = 0.f;
                }
                load_tile(tile + (int64_t)gridDim.x * GROUPS);
            } else {
                for (int i = tid; i < TC_ROWS * Kp; i += TC_THREADS) {
                    const int r = i / Kp, k = i - r * Kp;
                    const float x = (r0 + r < a.rows && k < K) ? __ldg(a.x + (r0 + r) * K + k) : 0.f;
                    float hi, lo;
                    split_tf32(x, hi, lo);
                    const int off = umma_off(r, k, Kp);
                    a_hi[off] = hi;
                    a_lo[off] = lo;
                }
            }
        }
        fence_proxy_async();
        tc_fence_before();
        group_sync(group);

        for (int l = 0; l <= d; ++l) {
            const int Kp = l == 0 ? pl.K0p : H;
            const bool head = l == d;
            // ---- MMA issue: cross terms first, the dominant hi.hi last
            if (tid == 0) {
                tc_fence_after();
                const uint32_t w_hi_addr = smem_u32(sm + pl.off_w[l][0]), w_lo_addr = smem_u32(sm + pl.off_w[l][1]);
                const uint32_t sbo = (uint32_t)Kp * 32;  // bytes between 8-row groups: Kp/4 chunks x 128 B
                const uint32_t d_tmem = tmem_base + (head ? 64u : 0u);
                const uint32_t idesc = head ? idesc_head : idesc_trunk;
                uint32_t acc = 0;
                for (int term = 0; term < 3; ++term) {
                    const uint32_t aa = term == 1 ? a_lo_addr : a_hi_addr;
                    const uint32_t bb = term == 0 ? w_lo_addr : w_hi_addr;
                    for (int k8 = 0; k8 < Kp / 8; ++k8) {
                        tc_mma_tf32(d_tmem, umma_desc(aa + k8 * 256, 128, sbo), umma_desc(bb + k8 * 256, 128, sbo), idesc,
                                    acc);
                        acc = 1;
                    }
                }
                tc_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            tc_fence_after();

            if (!head) {
                // ---- epilogue: bias + exact-erf GELU (+ residual) -> split -> next A operand
                float v[32];
                tmem_ld32(lane_addr + (uint32_t)col0, v);
                const bool residual = (l == 0 ? s.in_dim : H) == H;
                const float *bl = bias + l * H + col0;
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const int off = umma_off(row, col0 + 4 * c4, H);
                    float4 xh = make_float4(0.f, 0.f, 0.f, 0.f), xl = xh;
                    if (residual) {
                        xh = *reinterpret_cast<const float4 *>(a_hi + off);
                        xl = *reinterpret_cast<const float4 *>(a_lo + off);
                    }
                    const float xr[4] = {xh.x + xl.x, xh.y + xl.y, xh.z + xl.z, xh.w + xl.w};
                    float hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float z = v[4 * c4 + j] + bl[4 * c4 + j];
                        float y = gelu_erf(z);
                        if (residual) y = y + xr[j];
                        split_tf32(y, hi[j], lo[j]);
                    }
                    // a non-residual first layer may have K0p != H: its output still uses the H-wide layout,
                    // which overlaps the input operand -> all reads of this layer's A are complete (MMA committed)
                    *reinterpret_cast<float4 *>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4 *>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
                fence_proxy_async();
            } else if (warp < 4) {
                float v[16];
                tmem_ld16(lane_addr + 64u, v);
                if (r0 + row < a.rows) {
                    const float *bl = bias + d * H;
#pragma unroll
                    for (int o = 0; o < TC_HEAD_N; ++o)
                        if (o < O) a.out[(r0 + row) * O + o] = v[o] + bl[o];
                }
            }
            tc_fence_before();
            group_sync(group);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x < 32) tmem_dealloc(*tmem_slot, TMEM_COLS);
}


// ---------------------------------------------------------------- features-on-M engine: standalone forward
// The layer engine of the update kernels (tc_engine.cuh) on its own: one CTA per `rows_per_cta` rows
// (a multiple of 8, <= 256), D[64, R] = W . X^T per layer, activations kept IN PLACE as (hi, lo) operand
// planes, next layer's weights fetched into registers while the tensor pipe and the epilogue work.
// `variant` 0: epilogue on tcgen05.ld.16x256b fragments (all 32 lanes of a warp carry data);
// `variant` 1: tcgen05.ld.32x32b (lane = feature; half of each warp idles with M = 64).  Exists so that the
// fragment mapping is pinned by a test before the fused kernels rely on it.
struct TcfArgs {
    const float *params, *x;
    float *out;
    NetShape s;
    int64_t rows;
    int rows_per_cta, variant;
    long long *probe;  // optional: per layer {issue start, MMAs done, epilogue done, layer done} clocks of CTA 0, thread 0
};

__global__ void __launch_bounds__(NT, 1) k_mlp_forward_tcf(const TcfArgs a) {
    extern __shared__ __align__(128) float smem_tc[];
    const NetShape s = a.s;
    const int H = s.hidden, d = s.depth, O = s.out_dim;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int RC = a.rows_per_cta;
    const int64_t r0 = (int64_t)blockIdx.x * RC;
    const int rows = (int)min((int64_t)RC, a.rows - r0);
    const int R = round_up(rows, 8);
    const int K0p = round_up(s.in_dim, 8);
    const int ka = K0p > H ? K0p : H;
    // plan: activations (hi, lo) [RC, ka]; two weight slots (hi, lo) [64, ka]; biases; mbarrier + tmem slot
    float *x_hi = smem_tc, *x_lo = x_hi + RC * ka;
    float *w_slot[2][2];
    float *p = x_lo + RC * ka;
    for (int q = 0; q < 2; ++q)
        for (int h = 0; h < 2; ++h) { w_slot[q][h] = p; p += TCF_M * ka; }
    float *bias = p; p += 2 * TCF_M;
    uint64_t *bar = reinterpret_cast<uint64_t *>(p);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(p + 2);

    const uint32_t tmem_cols = tcf_tmem_cols(RC, a.variant == 2);
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < R * K0p; i += NT) {
        const int r = i / K0p, k = i - r * K0p;
        const float x = (r < rows && k < s.in_dim) ? __ldg(a.x + (r0 + r) * s.in_dim + k) : 0.f;
        float hi, lo;
        split_tf32(x, hi, lo);
        const int off = umma_off(r, k, K0p);
        x_hi[off] = hi;
        x_lo[off] = lo;
    }
    tcf_stage_weights(w_slot[0][0], w_slot[0][1], a.params + net_w_off(s, 0), H, s.in_dim, K0p);
    if (tid < TCF_M) bias[tid] = __ldg(a.params + net_b_off(s, 0) + tid);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    uint32_t phase = 0;

    if (a.variant != 1) {  // the fused kernels' own layer routine (tc_engine.cuh); 2: cross terms in their own accumulator
        TcfCtx cx;
        cx.x_hi = x_hi; cx.x_lo = x_lo;
        for (int q = 0; q < 2; ++q)
            for (int h = 0; h < 2; ++h) cx.w[q][h] = w_slot[q][h];
        cx.bias = bias; cx.bar = bar; cx.tmem = tmem; cx.phase = 0; cx.slot = 0;
        cx.cross_cols = a.variant == 2 ? (uint32_t)round_up(RC, 8) : 0u;
        TcfJob job = tcf_trunk_job(s, a.params, 0);
        for (int l = 0; l <= d; ++l) {
            const bool probe = a.probe && blockIdx.x == 0 && tid == 0;
            if (probe) a.probe[l * 5 + 0] = clock64();
            if (l < d) {
                const TcfJob nxt = l + 1 < d ? tcf_trunk_job(s, a.params, l + 1) : tcf_head_job(s, a.params);
                tcf_layer<false>(cx, job, &nxt, R, job.K == TCF_M, nullptr, 0, 0);
                job = nxt;
            } else {
                tcf_layer<true>(cx, job, nullptr, R, false, a.out + r0 * O, O, rows);
            }
            if (probe) a.probe[l * 5 + 4] = clock64();
        }
        tc_fence_after();
        if (warp == 0) tmem_dealloc(tmem, tmem_cols);
        return;
    }
    for (int l = 0; l <= d; ++l) {
        const int Kp = l == 0 ? K0p : H;
        const bool head = l == d;
        const int cur = l & 1, nxt = cur ^ 1;
        const bool probe = a.probe && blockIdx.x == 0 && tid == 0;
        if (probe) a.probe[l * 5 + 0] = clock64();
        if (warp == 0)
            tcf_issue(tmem, tmem, smem_u32(w_slot[cur][0]), smem_u32(w_slot[cur][1]), smem_u32(x_hi), smem_u32(x_lo), Kp, R,
                      bar);
        if (probe) a.probe[l * 5 + 1] = clock64();
        // next layer's weights (and bias): global -> registers while the MMAs run
        float4 wn[2];
        float bn = 0.f;
        const bool more = l < d;
        const int Nn = (l + 1 == d) ? O : H;  // rows of the next weight matrix (the head is zero padded to 64)
        if (more) {
            const float *Wn = a.params + net_w_off(s, l + 1);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int e = 4 * (tid + NT * q);  // element index in [64, H]
                const int n = e / H;
                wn[q] = n < Nn ? __ldg(reinterpret_cast<const float4 *>(Wn + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (tid < TCF_M) bn = tid < Nn ? __ldg(a.params + net_b_off(s, l + 1) + tid) : 0.f;
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        if (probe) a.probe[l * 5 + 2] = clock64();

        const bool residual = !head && Kp == H;
        const float *bl = bias + cur * TCF_M;
        const int sp = warp & 3, cgrp = warp >> 2;
        const uint32_t lane_base = tmem + ((uint32_t)(32 * sp) << 16);
        for (int c8 = cgrp; c8 < (R >> 3); c8 += 4) {
            if (a.variant == 0) {
                float v[4];
                tmem_ld_16x256b(lane_base + (uint32_t)(c8 * 8), v);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = 16 * sp + (lane >> 2) + 8 * h;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int r = c8 * 8 + 2 * (lane & 3) + cc;
                        const float z = v[2 * h + cc] + bl[j];
                        if (head) {
                            if (j < O && r < rows) a.out[(r0 + r) * O + j] = z;
                        } else {
                            const int off = umma_off(r, j, H);
                            float y = gelu_erf(z);
                            if (residual) y = y + (x_hi[off] + x_lo[off]);
                            float hi, lo;
                            split_tf32(y, hi, lo);
                            x_hi[off] = hi;
                            x_lo[off] = lo;
                        }
                    }
                }
            } else {
                float v[8];
                tmem_ld_32x32b_x8(lane_base + (uint32_t)(c8 * 8), v);
                if (lane < 16) {
                    const int j = 16 * sp + lane;
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        const int r = c8 * 8 + cc;
                        const float z = v[cc] + bl[j];
                        if (head) {
                            if (j < O && r < rows) a.out[(r0 + r) * O + j] = z;
                        } else {
                            const int off = umma_off(r, j, H);
                            float y = gelu_erf(z);
                            if (residual) y = y + (x_hi[off] + x_lo[off]);
                            float hi, lo;
                            split_tf32(y, hi, lo);
                            x_hi[off] = hi;
                            x_lo[off] = lo;
                        }
                    }
                }
            }
        }
        if (probe) a.probe[l * 5 + 3] = clock64();
        if (more) {  // registers -> the other weight slot (its last reader, layer l - 1, has committed)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int e = 4 * (tid + NT * q);
                const int n = e / H, k = e - n * H;
                const float wv[4] = {wn[q].x, wn[q].y, wn[q].z, wn[q].w};
                float hi[4], lo[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) split_tf32(wv[t], hi[t], lo[t]);
                const int off = umma_off(n, k, H);
                *reinterpret_cast<float4 *>(w_slot[nxt][0] + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4 *>(w_slot[nxt][1] + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            if (tid < TCF_M) bias[nxt * TCF_M + tid] = bn;
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (probe) a.probe[l * 5 + 4] = clock64();
    }
    tc_fence_after();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace asac

using namespace asac;

extern "C" int asac_mlp_forward_tc(const float *params, int in_dim, int hidden, int depth, int out_dim, const float *x,
                                   int64_t rows, float *out, void *stream) {
    ASAC_UNSUPPORTED(hidden != 64, "asac_mlp_forward_tc: hidden width %d (the tensor-core tile is built for 64)", hidden);
    ASAC_UNSUPPORTED(depth < 1 || depth > ASAC_MAX_DEPTH, "asac_mlp_forward_tc: depth %d", depth);
    ASAC_UNSUPPORTED(out_dim < 1 || out_dim > TC_HEAD_N, "asac_mlp_forward_tc: out_dim %d > %d", out_dim, TC_HEAD_N);
    ASAC_REQUIRE(in_dim > 0 && rows > 0, "asac_mlp_forward_tc: bad sizes");
    TcArgs a;
    a.params = params; a.x = x; a.out = out;
    a.s = NetShape{in_dim, hidden, depth, out_dim};
    a.rows = rows;
    int dev = 0, sms = 148;
    ASAC_CUDA(cudaGetDevice(&dev));
    ASAC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t tiles = (rows + TC_ROWS - 1) / TC_ROWS;
    // two tile groups per CTA when their operands fit and there is more than one tile per SM
    const int bytes2 = tc_plan(a.s, 2).total * 4 + 128, bytes1 = tc_plan(a.s, 1).total * 4 + 128;
    const int groups = (bytes2 <= 227 * 1024 && tiles > sms) ? 2 : 1;
    const int bytes = groups == 2 ? bytes2 : bytes1;
    ASAC_UNSUPPORTED(bytes > 227 * 1024, "asac_mlp_forward_tc: %d bytes of shared memory", bytes);
    static thread_local int granted[2][16];
    if (dev >= 16 || granted[groups - 1][dev] < bytes) {
        if (groups == 2)
            ASAC_CUDA(cudaFuncSetAttribute(k_mlp_forward_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        else
            ASAC_CUDA(cudaFuncSetAttribute(k_mlp_forward_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        if (dev < 16) granted[groups - 1][dev] = bytes;
    }
    const int64_t ctas = (tiles + groups - 1) / groups;
    const unsigned grid = (unsigned)(ctas < sms ? ctas : sms);
    if (groups == 2) k_mlp_forward_tc<2><<<grid, TC_THREADS * 2, bytes, (cudaStream_t)stream>>>(a);
    else k_mlp_forward_tc<1><<<grid, TC_THREADS, bytes, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_mlp_forward_tc");
    return ASAC_OK;
}

static int mlp_forward_tcf(const float *params, int in_dim, int hidden, int depth, int out_dim, const float *x,
                           int64_t rows, float *out, int rows_per_cta, int variant, long long *probe, void *stream) {
    ASAC_UNSUPPORTED(hidden != 64, "asac_mlp_forward_tcf: hidden width %d (UMMA_M is the hidden width: 64)", hidden);
    ASAC_UNSUPPORTED(depth < 1 || depth > ASAC_MAX_DEPTH, "asac_mlp_forward_tcf: depth %d", depth);
    ASAC_UNSUPPORTED(out_dim < 1 || out_dim > TCF_M, "asac_mlp_forward_tcf: out_dim %d > 64", out_dim);
    ASAC_REQUIRE(in_dim > 0 && rows > 0, "asac_mlp_forward_tcf: bad sizes");
    ASAC_REQUIRE(rows_per_cta >= 8 && rows_per_cta <= TCF_MAX_ROWS && rows_per_cta % 8 == 0,
                 "asac_mlp_forward_tcf: rows_per_cta %d must be a multiple of 8 in [8, 256]", rows_per_cta);
    TcfArgs a;
    a.params = params; a.x = x; a.out = out;
    a.s = NetShape{in_dim, hidden, depth, out_dim};
    a.rows = rows; a.rows_per_cta = rows_per_cta; a.variant = variant; a.probe = probe;
    const int K0p = round_up(in_dim, 8), ka = K0p > hidden ? K0p : hidden;
    const int bytes = (2 * rows_per_cta * ka + 4 * TCF_M * ka + 2 * TCF_M + 8) * 4 + 128;
    ASAC_UNSUPPORTED(bytes > 227 * 1024, "asac_mlp_forward_tcf: %d bytes of shared memory", bytes);
    static thread_local int granted[16];
    int dev = 0;
    ASAC_CUDA(cudaGetDevice(&dev));
    if (dev >= 16 || granted[dev] < bytes) {
        ASAC_CUDA(cudaFuncSetAttribute(k_mlp_forward_tcf, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        if (dev < 16) granted[dev] = bytes;
    }
    k_mlp_forward_tcf<<<(unsigned)((rows + rows_per_cta - 1) / rows_per_cta), NT, bytes, (cudaStream_t)stream>>>(a);
    ASAC_LAUNCHED("k_mlp_forward_tcf");
    return ASAC_OK;
}

extern "C" int asac_mlp_forward_tcf(const float *params, int in_dim, int hidden, int depth, int out_dim, const float *x,
                                    int64_t rows, float *out, int rows_per_cta, int variant, void *stream) {
    return mlp_forward_tcf(params, in_dim, hidden, depth, out_dim, x, rows, out, rows_per_cta, variant, nullptr, stream);
}

// debug: the same launch with per-layer phase clocks of CTA 0 written to probe[(depth + 1) * 5] (device memory)
extern "C" int asac_mlp_forward_tcf_probe(const float *params, int in_dim, int hidden, int depth, int out_dim,
                                          const float *x, int64_t rows, float *out, int rows_per_cta, int variant,
                                          long long *probe, void *stream) {
    return mlp_forward_tcf(params, in_dim, hidden, depth, out_dim, x, rows, out, rows_per_cta, variant, probe, stream);
}
