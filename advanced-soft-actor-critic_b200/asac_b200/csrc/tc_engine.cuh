// tcgen05 building blocks shared by the tensor-core kernels (sm_100a).
//
// Two orientations of one 64-wide ResBlock layer  Z = X . W^T  exist in this library:
//   * rows on M (mlp_tc.cu):      D[128 rows, 64] = X[128, K] . W[64, K]^T      — bulk forward, 128-row tiles;
//   * features on M (this file):  D[64, R]        = W[64, K] . X[R, K]^T        — the UPDATE kernels.
// The second form is what makes tensor cores usable on the latency-bound update path: UMMA_M is the
// hidden width (64: the smallest M tcgen05 has), UMMA_N is the number of batch rows a CTA holds — any
// multiple of 8 from 8 to 256 — so a 16-row tile costs 24 small MMAs (192 cycles) instead of padding
// the rows to 128, and a 208-row tile (get_l_probs over a burn-in window) is ONE MMA batch.
// The accumulator is transposed: TMEM lane = output feature, TMEM column = batch row; with M = 64 the
// data sit in lanes 0-15 of each of the four 32-lane sub-partitions (feature j -> lane 32 (j / 16) + j % 16).
//
// Operands live in shared memory in the UMMA K-major, no-swizzle core-matrix layout (8 rows x 16 bytes
// contiguous; K chunks 128 B apart; 8-row groups Kp * 32 B apart), each as a (hi, lo) pair for 3xTF32:
// hi = x with the 13 low mantissa bits cleared (exact in tf32), lo = x - hi;
// D = A_hi.B_lo + A_lo.B_hi + A_hi.B_hi accumulated in fp32 keeps ~21 mantissa bits per product, which is
// what the 1e-5 parity bound needs (plain tf32 keeps 10).
#pragma once
#include "common.cuh"
#include "mlp_tile.cuh"

namespace asac {

// ---------------------------------------------------------------- tcgen05 primitives
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32, cta_group::1
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor): start address,
// leading (K-chunk) byte offset and stride (8-row group) byte offset, all in 16-byte units; version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// cute::UMMA::InstrDescriptor: D fp32, A/B tf32, both K-major, dense
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// float offset of element (row, k) of an operand with Kp (multiple of 8) columns in the core-matrix layout
__host__ __device__ __forceinline__ int umma_off(int row, int k, int Kp) {
    return (row >> 3) * (Kp * 8) + (k >> 2) * 32 + (row & 7) * 4 + (k & 3);
}
// Both halves are ROUNDED to tf32 (cvt.rna: nearest, 11 significant bits) here, on the CUDA cores: the tensor core
// itself truncates its operands, and a truncated hi leaves a 13-bit lo of which the hardware then drops the two
// lowest bits — a biased error of up to 2^-20 |x| per operand (measured: rms error of a 3-layer net 2.6e-7 of scale
// against 3.6e-8 for FFMA).  With hi rounded, |lo| <= 2^-12 |x| and its own rounding costs <= 2^-24 |x|: fp32 level.
// (integer form of cvt.rna.tf32.f32 — add half an ulp of the 11-bit significand, clear the 13 low bits; a carry
//  ripples into the exponent as it should.  Two full-rate ALU instructions: the cvt itself goes through the
//  quarter-rate conversion pipe and cost 2 us per value pass, measured.)
__device__ __forceinline__ float rna_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = rna_tf32(x);
    lo = rna_tf32(x - hi);
}

// ---------------------------------------------------------------- features-on-M layer engine
constexpr int TCF_M = 64;            // UMMA_M = hidden width
constexpr int TCF_MAX_ROWS = 256;    // UMMA_N limit
constexpr uint32_t TCF_TMEM_COLS = 256;
// TMEM columns of a CTA that runs row batches of up to R rows: the allocation is a power of two >= 32; with the
// cross terms in their own accumulator (at column offset `cross`) twice the rows
__host__ __device__ __forceinline__ uint32_t tcf_tmem_cols(int R, bool split) {
    uint32_t need = (uint32_t)(split ? 2 * R : R), c = 32;
    while (c < need) c <<= 1;
    return c;
}

// One WARP (all 32 lanes enter; one elected lane issues): the 3 x Kp/8 MMAs of  D[64, R] = W[64, Kp] . X[R, Kp]^T
// (cross terms first, hi.hi last) and the commit.  w_* / x_*: shared-memory byte addresses of the (hi, lo) operand
// planes; R a multiple of 8, 8..256.
// Issue cost matters on the latency-bound update path: from ONE divergent thread (`if (tid == 0)`) every operand
// of every tcgen05.mma went through R2UR and an issue took ~100 cycles (2400 per 64-wide layer, measured).  With a
// warp-uniform branch, an elected lane and fully unrolled descriptor arithmetic the operands stay in uniform registers.
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void tc_mma_tf32_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate, uint32_t elected) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(elected)
        : "memory");
}
template <int K8>
__device__ __forceinline__ void tcf_issue_steps(uint32_t d_tmem, uint32_t d_cross, uint32_t w_hi, uint32_t w_lo,
                                                uint32_t x_hi, uint32_t x_lo, uint32_t sbo, uint32_t idesc,
                                                uint32_t elected) {
    const uint64_t hi_mask = ((uint64_t)((128u >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
                             ((uint64_t)1 << 46);
#pragma unroll
    for (int term = 0; term < 3; ++term) {
        const uint32_t aa = term == 1 ? w_lo : w_hi;
        const uint32_t bb = term == 0 ? x_lo : x_hi;
#pragma unroll
        for (int k8 = 0; k8 < K8; ++k8) {
            const uint64_t ad = hi_mask | (uint64_t)(((aa + k8 * 256) >> 4) & 0x3FFF);
            const uint64_t bd = hi_mask | (uint64_t)(((bb + k8 * 256) >> 4) & 0x3FFF);
            // the first MMA into an accumulator overwrites it: term 0 opens the cross-term accumulator, term 2 the
            // hi.hi accumulator when that is a separate one
            const bool first = k8 == 0 && (term == 0 || (term == 2 && d_cross != d_tmem));
            tc_mma_tf32_elect(term == 2 ? d_tmem : d_cross, ad, bd, idesc, first ? 0u : 1u, elected);
        }
    }
}
// d_cross: accumulator of the two cross terms — d_tmem itself (one accumulation chain of 3 Kp / 8 MMAs) or a second
// column range (the epilogue adds the two): the tensor core truncates the running sum at every MMA, so the error of
// a chain grows with its length and with the magnitude of what it holds; the cross terms are 2^-11 of the result and
// their chain's truncation is invisible, which leaves Kp / 8 truncations on the hi.hi chain instead of 3 Kp / 8.
__device__ __forceinline__ void tcf_issue(uint32_t d_tmem, uint32_t d_cross, uint32_t w_hi, uint32_t w_lo, uint32_t x_hi,
                                          uint32_t x_lo, int Kp, int R, uint64_t *bar) {
    tc_fence_after();
    const uint32_t sbo = (uint32_t)Kp * 32;
    const uint32_t idesc = umma_idesc_tf32(TCF_M, R);
    const uint32_t elected = elect_one();
    if (Kp == 64) {
        tcf_issue_steps<8>(d_tmem, d_cross, w_hi, w_lo, x_hi, x_lo, sbo, idesc, elected);
    } else if (Kp == 8) {
        tcf_issue_steps<1>(d_tmem, d_cross, w_hi, w_lo, x_hi, x_lo, sbo, idesc, elected);
    } else if (Kp == 16) {
        tcf_issue_steps<2>(d_tmem, d_cross, w_hi, w_lo, x_hi, x_lo, sbo, idesc, elected);
    } else {
#pragma unroll 1
        for (int term = 0; term < 3; ++term) {
            const uint32_t aa = term == 1 ? w_lo : w_hi;
            const uint32_t bb = term == 0 ? x_lo : x_hi;
#pragma unroll 2
            for (int k8 = 0; k8 < (Kp >> 3); ++k8) {
                const bool first = k8 == 0 && (term == 0 || (term == 2 && d_cross != d_tmem));
                tc_mma_tf32_elect(term == 2 ? d_tmem : d_cross, umma_desc(aa + k8 * 256, 128, sbo),
                                  umma_desc(bb + k8 * 256, 128, sbo), idesc, first ? 0u : 1u, elected);
            }
        }
    }
    if (elected) tc_commit(bar);
    __syncwarp();
}

// Accumulator fragment of a warp: 16 features (lanes 0-15 of the warp's TMEM sub-partition) x 8 batch rows.
// tcgen05.ld.16x256b.x1: thread t holds features {t / 4, t / 4 + 8} (relative to 16 * (warp % 4)) for the
// columns {2 (t % 4), 2 (t % 4) + 1} of the 8-column chunk:  v[0], v[1] = (f0, c0), (f0, c0 + 1);
// v[2], v[3] = (f0 + 8, c0), (f0 + 8, c0 + 1).  All 32 threads carry data.
__device__ __forceinline__ void tmem_ld_16x256b_nowait(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b(uint32_t taddr, float (&v)[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(taddr)
                 : "memory");
    tmem_wait_ld();
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
// tcgen05.ld.32x32b.x8: thread t reads 8 consecutive columns of TMEM lane 32 (warp % 4) + t (with M = 64 only
// threads 0-15 of a warp see accumulator rows)
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// Converts `n_rows` x K weights W[n, k] (row-major, global) into the (hi, lo) operand planes of a [64, Kp]
// A operand (rows >= n_rows and columns >= K zero).  All NT threads.
__device__ __forceinline__ void tcf_stage_weights(float *w_hi, float *w_lo, const float *W, int n_rows, int K, int Kp) {
    for (int i = threadIdx.x; i < TCF_M * Kp; i += NT) {
        const int n = i / Kp, k = i - n * Kp;
        const float w = (n < n_rows && k < K) ? __ldg(W + (int64_t)n * K + k) : 0.f;
        float hi, lo;
        split_tf32(w, hi, lo);
        const int off = umma_off(n, k, Kp);
        w_hi[off] = hi;
        w_lo[off] = lo;
    }
}

// ---------------------------------------------------------------- engine context of a CTA
struct TcfCtx {
    float *x_hi, *x_lo;   // activation operand planes [rows, <= 64] (in place from layer to layer)
    float *w[2][2];       // two weight slots x (hi, lo): [64, <= 64]
    float *bias;          // [2][64]
    uint64_t *bar;
    uint32_t tmem, phase;
    uint32_t cross_cols;  // column offset of the cross-term accumulator (0: one accumulator for all three terms)
    int slot;             // weight slot holding the CURRENT job
};

// One GEMM stage of a net: rows of W actually present (N <= 64; the rest of the A operand is zero), K, K padded to 8
struct TcfJob {
    const float *W, *b;
    int N, K, Kp;
};
__device__ __forceinline__ TcfJob tcf_trunk_job(const NetShape &s, const float *params, int l) {
    const int K = net_k(s, l);
    return TcfJob{params + net_w_off(s, l), params + net_b_off(s, l), s.hidden, K, round_up(K, 8)};
}
__device__ __forceinline__ TcfJob tcf_head_job(const NetShape &s, const float *params) {
    return TcfJob{params + net_w_off(s, s.depth), params + net_b_off(s, s.depth), s.out_dim, s.hidden, s.hidden};
}

// A job's weights on their way global -> registers -> operand planes.  Two mappings:
//  * K == 64: 16-byte loads; a quarter warp covers the 8 rows of a core-matrix group at one K chunk, so its
//    16-byte stores fill 128 contiguous bytes (conflict free) and a warp reads two full 32-byte sectors per row;
//  * otherwise (first layers, K = S or S + A): scalar loads, <= 8 per thread.
struct TcfWeights {
    float v[8];
    float b;
};
__device__ __forceinline__ void tcf_prefetch(TcfWeights &p, const TcfJob &j) {
    const int tid = threadIdx.x;
    if (j.K == 64) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = tid + NT * q, w = i >> 5, lane = i & 31;
            const int n = 8 * (w >> 2) + (lane & 7), c = 4 * (w & 3) + (lane >> 3);
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < j.N) t = __ldg(reinterpret_cast<const float4 *>(j.W + (int64_t)n * 64 + 4 * c));
            p.v[4 * q] = t.x; p.v[4 * q + 1] = t.y; p.v[4 * q + 2] = t.z; p.v[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = tid + NT * q;
            const int n = i / j.Kp, k = i - n * j.Kp;
            p.v[q] = (i < TCF_M * j.Kp && n < j.N && k < j.K) ? __ldg(j.W + (int64_t)n * j.K + k) : 0.f;
        }
    }
    p.b = (tid < TCF_M && tid < j.N && j.b) ? __ldg(j.b + tid) : 0.f;
}
__device__ __forceinline__ void tcf_store(const TcfWeights &p, const TcfJob &j, float *w_hi, float *w_lo, float *bias) {
    const int tid = threadIdx.x;
    if (j.K == 64) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = tid + NT * q, w = i >> 5, lane = i & 31;
            const int off = (w >> 2) * 512 + (4 * (w & 3) + (lane >> 3)) * 32 + (lane & 7) * 4;
            float hi[4], lo[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) split_tf32(p.v[4 * q + t], hi[t], lo[t]);
            *reinterpret_cast<float4 *>(w_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4 *>(w_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = tid + NT * q;
            if (i < TCF_M * j.Kp) {
                const int n = i / j.Kp, k = i - n * j.Kp;
                float hi, lo;
                split_tf32(p.v[q], hi, lo);
                const int off = umma_off(n, k, j.Kp);
                w_hi[off] = hi;
                w_lo[off] = lo;
            }
        }
    }
    if (tid < TCF_M) bias[tid] = p.b;
}

// Runs the CURRENT job on the R rows (multiple of 8) in the activation planes and leaves the NEXT job's weights
// in the other slot.  Trunk layer: y = gelu(z + b) (+ x when `residual`), written back IN PLACE as (hi, lo) with
// the 64-wide layout; z is also stored to `z_save[r * 64 + j]` when given.  Head: out[r * ldo + j] = z + b for
// j < cur.N, r < rows_valid; the planes are left untouched.  Ends with a CTA barrier.
template <bool HEAD>
__device__ __forceinline__ void tcf_layer(TcfCtx &cx, const TcfJob &cur, const TcfJob *next, int R, bool residual,
                                          float *out, int ldo, int rows_valid, float *z_save = nullptr) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s = cx.slot;
    if (warp == 0)
        tcf_issue(cx.tmem, cx.tmem + cx.cross_cols, smem_u32(cx.w[s][0]), smem_u32(cx.w[s][1]), smem_u32(cx.x_hi),
                  smem_u32(cx.x_lo), cur.Kp, R, cx.bar);
    TcfWeights pre;
    if (next) tcf_prefetch(pre, *next);
    const float *bl = cx.bias + s * TCF_M;
    const int sp = warp & 3;
    const uint32_t lane_base = cx.tmem + ((uint32_t)(32 * sp) << 16);
    const int f0 = 16 * sp + (lane >> 2);
    const float bj[2] = {bl[f0], bl[f0 + 8]};
    // the next job's weights go to the other slot while the tensor pipe works (their loads were issued above and
    // the slot's last reader committed two jobs ago): off the critical path of the epilogue
    if (next) tcf_store(pre, *next, cx.w[s ^ 1][0], cx.w[s ^ 1][1], cx.bias + (s ^ 1) * TCF_M);
    mbar_wait(cx.bar, cx.phase);
    cx.phase ^= 1;
    tc_fence_after();
#pragma unroll 1
    for (int c8 = warp >> 2; c8 < (R >> 3); c8 += 4) {
        float v[4];
        if (cx.cross_cols) {
            uint32_t ra[4], rb[4];
            tmem_ld_16x256b_nowait(lane_base + (uint32_t)(c8 * 8), ra);
            tmem_ld_16x256b_nowait(lane_base + cx.cross_cols + (uint32_t)(c8 * 8), rb);
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = __uint_as_float(ra[q]) + __uint_as_float(rb[q]);
        } else {
            tmem_ld_16x256b(lane_base + (uint32_t)(c8 * 8), v);
        }
        const int r0 = c8 * 8 + 2 * (lane & 3);
        if constexpr (HEAD) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int j = f0 + 8 * h, r = r0 + cc;
                    if (j < cur.N && r < rows_valid) out[r * ldo + j] = v[2 * h + cc] + bj[h];
                }
        } else {
            // loads first, four independent GELU chains, stores last (a store between two loads of the same planes
            // serialised the four elements: 1000 cycles per chunk, measured)
            int off[4];
            float res[4], y[4];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) off[2 * h + cc] = umma_off(r0 + cc, f0 + 8 * h, TCF_M);
            if (residual) {
                float xh[4], xl[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { xh[q] = cx.x_hi[off[q]]; xl[q] = cx.x_lo[off[q]]; }
#pragma unroll
                for (int q = 0; q < 4; ++q) res[q] = xh[q] + xl[q];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float z = v[q] + bj[q >> 1];
                if (z_save) z_save[(r0 + (q & 1)) * TCF_M + f0 + 8 * (q >> 1)] = z;
                y[q] = gelu_erf(z);
                if (residual) y[q] = y[q] + res[q];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float hi, lo;
                split_tf32(y[q], hi, lo);
                cx.x_hi[off[q]] = hi;
                cx.x_lo[off[q]] = lo;
            }
        }
    }
    cx.slot = s ^ 1;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
}

// element (r, k) of the activation planes (Kp-wide layout)
__device__ __forceinline__ void tcf_put(TcfCtx &cx, int r, int k, int Kp, float x) {
    float hi, lo;
    split_tf32(x, hi, lo);
    const int off = umma_off(r, k, Kp);
    cx.x_hi[off] = hi;
    cx.x_lo[off] = lo;
}

}  // namespace asac
