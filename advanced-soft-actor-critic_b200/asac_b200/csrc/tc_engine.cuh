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
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

// ---------------------------------------------------------------- features-on-M layer engine
constexpr int TCF_M = 64;            // UMMA_M = hidden width
constexpr int TCF_MAX_ROWS = 256;    // UMMA_N limit
constexpr uint32_t TCF_TMEM_COLS = 256;

// One thread: the 3 x Kp/8 MMAs of  D[64, R] = W[64, Kp] . X[R, Kp]^T  (cross terms first, hi.hi last) and the commit.
// w_* / x_*: shared-memory byte addresses of the (hi, lo) operand planes; R a multiple of 8, 8..256.
__device__ __forceinline__ void tcf_issue(uint32_t d_tmem, uint32_t w_hi, uint32_t w_lo, uint32_t x_hi, uint32_t x_lo,
                                          int Kp, int R, uint64_t *bar) {
    tc_fence_after();
    const uint32_t sbo = (uint32_t)Kp * 32;
    const uint32_t idesc = umma_idesc_tf32(TCF_M, R);
    uint32_t acc = 0;
#pragma unroll 1
    for (int term = 0; term < 3; ++term) {
        const uint32_t aa = term == 1 ? w_lo : w_hi;
        const uint32_t bb = term == 0 ? x_lo : x_hi;
#pragma unroll 1
        for (int k8 = 0; k8 < (Kp >> 3); ++k8) {
            tc_mma_tf32(d_tmem, umma_desc(aa + k8 * 256, 128, sbo), umma_desc(bb + k8 * 256, 128, sbo), idesc, acc);
            acc = 1;
        }
    }
    tc_commit(bar);
}

// Accumulator fragment of a warp: 16 features (lanes 0-15 of the warp's TMEM sub-partition) x 8 batch rows.
// tcgen05.ld.16x256b.x1: thread t holds features {t / 4, t / 4 + 8} (relative to 16 * (warp % 4)) for the
// columns {2 (t % 4), 2 (t % 4) + 1} of the 8-column chunk:  v[0], v[1] = (f0, c0), (f0, c0 + 1);
// v[2], v[3] = (f0 + 8, c0), (f0 + 8, c0 + 1).  All 32 threads carry data.
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

}  // namespace asac
