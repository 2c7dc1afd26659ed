// 16-row layer passes on the warp-level tensor-core path (mma.sync.m16n8k8 tf32, 3xTF32), sm_100a.
//
// NOT part of libasac_b200.so: this is the variant the micro-benchmark tools/ubench/layer_chain.cu measures against the
// FFMA routine (profiles/r02n_ubench_layer_chain.txt).  On B200 the legacy HMMA tf32 instruction issues about once per
// 28 cycles per scheduler — 36 MAC/clk against FFMA's 32 — so a 16-row pass takes 1 537 cycles here against 2 116 for
// the adopted FFMA tile, with four times the error: not worth the precision.  Kept as the record of that measurement.
//
// The update kernels of a replay-sized step hold ONE 16-row tile per CTA, so a layer is a 16 x H x K GEMM:
// exactly one m16 tile tall.  The FFMA routine of mlp_tile.cuh needs 512 FFMA issue cycles per scheduler,
// a K-split exchange through shared memory and two CTA barriers for it (~2 us per layer measured); tcgen05
// has a fixed cost per layer (operand planes, commit, TMEM read-back) that only pays from two passes on
// (tc_engine.cuh).  Here warp w owns the 8 output columns of n-tile w: A (the rows) and B (the staged weights)
// come straight from shared memory into registers, are split into tf32 hi + lo on the CUDA cores (round to
// nearest, so the tensor core's truncation of its operands never sees a discarded bit), three MMAs per k-step
// (hi.hi into one accumulator, hi.lo + lo.hi into another: the small terms are summed apart), and the epilogue
// (bias, exact-erf GELU, residual) runs on the accumulator registers: no partial exchange, one barrier.
//
// Fragment layout of mma.m16n8k8.row.col (PTX ISA, "Matrix Fragments for mma.m16n8k8"), g = lane >> 2, t = lane & 3:
//   A: a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4)       B: b0 (k = t, n = g)  b1 (k = t + 4, n = g)
//   C: c0 (g, 2t)  c1 (g, 2t + 1)  c2 (g + 8, 2t)  c3 (g + 8, 2t + 1)
// Shared-memory strides of 4 mod 32 floats (lda = ldw = H + 4) make every fragment load bank-conflict free; the
// TMA-staged weights ([H][32] segments, 16-byte chunk index XOR (row & 7)) are conflict free as well.
#pragma once
#include "mlp_tile.cuh"

namespace asac {

// x = hi + lo with hi = tf32(x) rounded to nearest (ties away) and lo = tf32(x - hi): integer form of
// cvt.rna.tf32.f32 (the cvt itself issues on the quarter-rate pipe)
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    const float r = x - __uint_as_float(hi);  // exact
    lo = (__float_as_uint(r) + 0x1000u) & 0xffffe000u;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// one k-step of C += A . B in 3xTF32; `a`, `b` hold fp32 values
__device__ __forceinline__ void mma_3x(float (&big)[4], float (&small)[4], const float (&a)[4], const float (&b)[2]) {
    uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
    for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
    mma_tf32(small, ah, bl);
    mma_tf32(small, al, bh);
    mma_tf32(big, ah, bh);
}

// ResBlock forward over ONE 16-row pass:  Z = X . W^T + b ;  Y = gelu(Z) (+ X when residual).
//   WARPS = 8:  warps 0..7 run all of K for one n-tile each (H = 64), warps 8..15 wait at the barrier
//   WARPS = 16: warp w -> n-tile w & 7, K half w >> 3; the upper halves' accumulators meet in `part`
//   TWO (with WARPS = 8): warps 8..15 run a second, independent 16-row tile (rows 16..31 of X / Z / Y)
// Ws: staged weights [H][ldw] (SWZ: TMA segments).  K == H == 64.
template <int WARPS, bool SWZ, bool TWO>
__device__ __noinline__ void mma_layer_forward(const float *X, int ldx, const float *Ws, int ldw, const float *bs, float *Zs,
                                               float *Ys, int ldy, bool residual, float *part) {
    ASAC_SMEM(X); ASAC_SMEM(Ws); ASAC_SMEM(bs); ASAC_SMEM(Ys); ASAC_SMEM(part);
    constexpr int H = 64, KSTEPS = H / 8, KSPL = WARPS / 8, PER = KSTEPS / KSPL;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int tile = warp & 7, ks = TWO ? 0 : warp >> 3;
    if (TWO) {
        X += (warp >> 3) * 16 * ldx;
        Ys += (warp >> 3) * 16 * ldy;
        if (Zs) Zs += (warp >> 3) * 16 * ldy;
    }
    float big[4] = {0.f, 0.f, 0.f, 0.f}, small[4] = {0.f, 0.f, 0.f, 0.f};
    if (ks < KSPL) {
        const float *xr0 = X + g * ldx + t, *xr1 = xr0 + 8 * ldx;
        const int n = tile * 8 + g;
        const float *wr = Ws + n * ldw + t;
#pragma unroll
        for (int kk = ks * PER; kk < ks * PER + PER; ++kk) {
            const int k0 = kk * 8;
            const float a[4] = {xr0[k0], xr1[k0], xr0[k0 + 4], xr1[k0 + 4]};
            float b[2];
            if (SWZ) {
                b[0] = Ws[swz_off(H, n, k0 + t)];
                b[1] = Ws[swz_off(H, n, k0 + t + 4)];
            } else {
                b[0] = wr[k0];
                b[1] = wr[k0 + 4];
            }
            mma_3x(big, small, a, b);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) big[i] += small[i];
    }
    if (KSPL > 1) {
        float4 *p4 = reinterpret_cast<float4 *>(part);
        if (ks == 1) p4[tile * 32 + lane] = make_float4(big[0], big[1], big[2], big[3]);
        __syncthreads();
        if (ks == 0) {
            const float4 o = p4[tile * 32 + lane];
            big[0] += o.x; big[1] += o.y; big[2] += o.z; big[3] += o.w;
        }
    }
    if (ks == 0) {
        const int col = tile * 8 + 2 * t;
        const float2 bb = *reinterpret_cast<const float2 *>(bs + col);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = g + 8 * h;
            const float z0 = big[2 * h] + bb.x, z1 = big[2 * h + 1] + bb.y;
            if (Zs) *reinterpret_cast<float2 *>(Zs + row * ldy + col) = make_float2(z0, z1);
            float y0 = gelu_erf(z0), y1 = gelu_erf(z1);
            if (residual) {
                const float2 xv = *reinterpret_cast<const float2 *>(X + row * ldx + col);
                y0 = y0 + xv.x; y1 = y1 + xv.y;
            }
            *reinterpret_cast<float2 *>(Ys + row * ldy + col) = make_float2(y0, y1);
        }
    }
    __syncthreads();
}

// Variant with operands that are ALREADY split: X planes (hi at X, lo at X + xplane; fp32 values for the residual at
// Xf) and, with WSPLIT, weight planes (hi at Ws, lo at Ws + wplane).  The epilogue writes the next layer's planes.
template <bool WSPLIT>
__device__ __noinline__ void mma_layer_forward_ps(const float *Xf, const float *X, int xplane, int ldx, const float *Ws,
                                                  int wplane, int ldw, const float *bs, float *Zs, float *Yf, float *Y,
                                                  int ldy, bool residual) {
    ASAC_SMEM(Xf); ASAC_SMEM(X); ASAC_SMEM(Ws); ASAC_SMEM(bs); ASAC_SMEM(Yf); ASAC_SMEM(Y);
    constexpr int H = 64, KSTEPS = H / 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    if (warp < 8) {
        const int tile = warp;
        float big[4] = {0.f, 0.f, 0.f, 0.f}, small[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t *xh0 = reinterpret_cast<const uint32_t *>(X) + g * ldx + t, *xh1 = xh0 + 8 * ldx;
        const uint32_t *wr = reinterpret_cast<const uint32_t *>(Ws) + (tile * 8 + g) * ldw + t;
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) {
            const int k0 = kk * 8;
            const uint32_t ah[4] = {xh0[k0], xh1[k0], xh0[k0 + 4], xh1[k0 + 4]};
            const uint32_t al[4] = {xh0[xplane + k0], xh1[xplane + k0], xh0[xplane + k0 + 4], xh1[xplane + k0 + 4]};
            uint32_t bh[2], bl[2];
            if (WSPLIT) {
                bh[0] = wr[k0]; bh[1] = wr[k0 + 4];
                bl[0] = wr[wplane + k0]; bl[1] = wr[wplane + k0 + 4];
            } else {
                split_tf32(__uint_as_float(wr[k0]), bh[0], bl[0]);
                split_tf32(__uint_as_float(wr[k0 + 4]), bh[1], bl[1]);
            }
            mma_tf32(small, ah, bl);
            mma_tf32(small, al, bh);
            mma_tf32(big, ah, bh);
        }
        const int col = tile * 8 + 2 * t;
        const float2 bb = *reinterpret_cast<const float2 *>(bs + col);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = g + 8 * h;
            const float z0 = (big[2 * h] + small[2 * h]) + bb.x, z1 = (big[2 * h + 1] + small[2 * h + 1]) + bb.y;
            if (Zs) *reinterpret_cast<float2 *>(Zs + row * ldy + col) = make_float2(z0, z1);
            float y0 = gelu_erf(z0), y1 = gelu_erf(z1);
            if (residual) {
                const float2 xv = *reinterpret_cast<const float2 *>(Xf + row * ldx + col);
                y0 = y0 + xv.x; y1 = y1 + xv.y;
            }
            *reinterpret_cast<float2 *>(Yf + row * ldy + col) = make_float2(y0, y1);
            uint32_t h0, l0, h1, l1;
            split_tf32(y0, h0, l0);
            split_tf32(y1, h1, l1);
            *reinterpret_cast<uint2 *>(Y + row * ldy + col) = make_uint2(h0, h1);
            *reinterpret_cast<uint2 *>(Y + xplane + row * ldy + col) = make_uint2(l0, l1);
        }
    }
    __syncthreads();
}

}  // namespace asac
