// Row-tile MLP building blocks for the SAC update kernels (sm_100a, fp32 FFMA path).
//
// A CTA of NT = 128 threads owns a tile of rows whose activations live in shared memory;
// layer weights are streamed L2 -> shared memory with cp.async (double buffered) and the
// CTA walks the net layer by layer.  Thread mapping of the GEMM core: lane group
// cg = tid & 15 owns columns {cg, cg+16, ...} (CM of them), row group rg = tid >> 4 owns
// rows {2rg, 2rg+1} of each 16-row pass; operands are read as float4 along K.
//
// Numerics follow the reference's torch fp32 ops: nn.Linear + exact-erf GELU + residual
// (algorithm/nn_models/layers/linear_layers.py:46-56).  The library is compiled with
// -fmad=false, so only the explicit fmaf() of the GEMM cores contract.
#pragma once
#include "common.cuh"

namespace asac {

constexpr int NT = 128;        // threads per CTA in every tiled kernel
constexpr int PASS_ROWS = 16;  // rows per GEMM pass

__host__ __device__ __forceinline__ int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------- flat parameter layout
struct NetShape {
    int in_dim, hidden, depth, out_dim;
};
__host__ __device__ __forceinline__ int64_t net_w_off(const NetShape &s, int l) {  // l == depth -> head
    // closed form (a loop here was unrolled at every call site: 31 % of k_value_pass's instructions)
    if (l == 0) return 0;
    return (int64_t)s.hidden * (s.in_dim + 1) + (int64_t)(l - 1) * s.hidden * (s.hidden + 1);
}
__host__ __device__ __forceinline__ int net_k(const NetShape &s, int l) { return l == 0 ? s.in_dim : s.hidden; }
__host__ __device__ __forceinline__ int net_n(const NetShape &s, int l) { return l == s.depth ? s.out_dim : s.hidden; }
__host__ __device__ __forceinline__ int64_t net_b_off(const NetShape &s, int l) {
    return net_w_off(s, l) + (int64_t)net_n(s, l) * net_k(s, l);
}
__host__ __device__ __forceinline__ int64_t net_count(const NetShape &s) {
    return net_b_off(s, s.depth) + s.out_dim;
}
__host__ __device__ __forceinline__ int64_t net_stride(const NetShape &s) { return (net_count(s) + 3) & ~(int64_t)3; }

// The layer routines below are deliberately NOT inlined: a tile kernel walks ~10 layers through
// three or four call sites, and with everything inlined (x4 width instantiations) k_value_pass
// grew to 18 K SASS instructions (290 KB) — the ncu source page showed > 50 % of the warp stall
// samples in `no_instruction` (instruction-cache misses) with only 4 warps per SM to hide them.
// One copy per width keeps the hot loop resident.  Pointer arguments that live in shared memory
// are declared to the compiler with ASAC_SMEM so the out-of-line code still uses LDS/STS.
#define ASAC_SMEM(p) __builtin_assume(__isShared(p))

// ---------------------------------------------------------------- cp.async helpers
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Stage W[N][K] (row-major, global) into Ws and b[N] into bs.
//   transpose == false : Ws[n * ldw + k], ldw = round_up(K,4) + 4, pad columns zeroed
//   transpose == true  : Ws[k * ldw + n], ldw = N + 4   (for dX = dZ . W)
__device__ __noinline__ void stage_weights(float *Ws, float *bs, const float *W, const float *b, int N, int K,
                                           bool transpose) {
    ASAC_SMEM(Ws);
    const int tid = threadIdx.x;
    if (!transpose) {
        const int K4 = round_up(K, 4), ldw = K4 + 4;
        if ((K & 3) == 0 && (((uintptr_t)W) & 15) == 0) {
            const int per_row = K >> 2;
            for (int i = tid; i < N * per_row; i += NT) {
                const int n = i / per_row, c = i - n * per_row;
                cp_async16(Ws + n * ldw + 4 * c, W + (int64_t)n * K + 4 * c);
            }
        } else {
            for (int i = tid; i < N * K; i += NT) {
                const int n = i / K, k = i - n * K;
                cp_async4(Ws + n * ldw + k, W + i);
            }
            if (K4 != K) {
                const int padc = K4 - K;
                for (int i = tid; i < N * padc; i += NT) Ws[(i / padc) * ldw + K + (i % padc)] = 0.f;
            }
        }
    } else {
        const int ldw = N + 4;
        for (int i = tid; i < N * K; i += NT) {
            const int n = i / K, k = i - n * K;
            cp_async4(Ws + k * ldw + n, W + i);
        }
    }
    if (bs && b)
        for (int i = tid; i < N; i += NT) cp_async4(bs + i, b + i);
}

// ---------------------------------------------------------------- activation
__device__ __forceinline__ float gelu_erf(float z) {
    return (z * 0.5f) * (1.f + erff(z * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float z) {
    const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * z * z);
    return cdf + z * pdf;
}

// ---------------------------------------------------------------- GEMM core
// acc[i][c] += sum_k A[(r + i) * lda + k] * Ws[(cg + 16 c) * ldw + k],  k < K4 (multiple of 4)
template <int CM>
__device__ __forceinline__ void gemm_core(const float *__restrict__ A, int lda, int K4,
                                          const float *__restrict__ Ws, int ldw, int r, int cg,
                                          float (&acc)[2][CM]) {
    const float *a0p = A + r * lda;
    const float *a1p = a0p + lda;
    const float *wp = Ws + cg * ldw;
#pragma unroll 2
    for (int k = 0; k < K4; k += 4) {
        const float4 a0 = *reinterpret_cast<const float4 *>(a0p + k);
        const float4 a1 = *reinterpret_cast<const float4 *>(a1p + k);
#pragma unroll
        for (int c = 0; c < CM; ++c) {
            const float4 w = *reinterpret_cast<const float4 *>(wp + 16 * c * ldw + k);
            acc[0][c] = fmaf(a0.x, w.x, acc[0][c]);
            acc[1][c] = fmaf(a1.x, w.x, acc[1][c]);
            acc[0][c] = fmaf(a0.y, w.y, acc[0][c]);
            acc[1][c] = fmaf(a1.y, w.y, acc[1][c]);
            acc[0][c] = fmaf(a0.z, w.z, acc[0][c]);
            acc[1][c] = fmaf(a1.z, w.z, acc[1][c]);
            acc[0][c] = fmaf(a0.w, w.w, acc[0][c]);
            acc[1][c] = fmaf(a1.w, w.w, acc[1][c]);
        }
    }
}

// ResBlock forward over `nrows` (multiple of 16) rows:
//   Z = X . W^T + b ;  Y = gelu(Z) (+ X when residual)        Zs may be null (no backward)
template <int CM>
__device__ __noinline__ void layer_forward_t(const float *X, int ldx, int K4, const float *Ws, const float *bs,
                                             float *Zs, float *Ys, int ldy, int nrows, bool residual) {
    ASAC_SMEM(X); ASAC_SMEM(Ws); ASAC_SMEM(bs); ASAC_SMEM(Ys);
    if (Zs) ASAC_SMEM(Zs);
    const int tid = threadIdx.x, cg = tid & 15, rg = tid >> 4;
    const int ldw = K4 + 4;
    for (int r0 = 0; r0 < nrows; r0 += PASS_ROWS) {
        const int r = r0 + 2 * rg;
        float acc[2][CM];
#pragma unroll
        for (int c = 0; c < CM; ++c) acc[0][c] = acc[1][c] = bs[cg + 16 * c];
        gemm_core<CM>(X, ldx, K4, Ws, ldw, r, cg, acc);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
            for (int c = 0; c < CM; ++c) {
                const int j = cg + 16 * c;
                const float z = acc[i][c];
                if (Zs) Zs[(r + i) * ldy + j] = z;
                float y = gelu_erf(z);
                if (residual) y = y + X[(r + i) * ldx + j];
                Ys[(r + i) * ldy + j] = y;
            }
        }
    }
}

__device__ __forceinline__ void layer_forward(int hidden, const float *X, int ldx, int K4, const float *Ws,
                                              const float *bs, float *Zs, float *Ys, int ldy, int nrows,
                                              bool residual) {
    switch (hidden >> 4) {
        case 1: layer_forward_t<1>(X, ldx, K4, Ws, bs, Zs, Ys, ldy, nrows, residual); break;
        case 2: layer_forward_t<2>(X, ldx, K4, Ws, bs, Zs, Ys, ldy, nrows, residual); break;
        case 4: layer_forward_t<4>(X, ldx, K4, Ws, bs, Zs, Ys, ldy, nrows, residual); break;
        default: layer_forward_t<8>(X, ldx, K4, Ws, bs, Zs, Ys, ldy, nrows, residual); break;
    }
}

// dX = dZ . W (+ dY when the block was residual), W staged transposed: Wt[k * ldw + j], ldw = H + 4.
// Output columns k < H (hidden -> hidden layers only).
template <int CM>
__device__ __noinline__ void layer_input_grad_t(const float *dZ, int ld, int H, const float *Wt, const float *dY,
                                                float *dX, int nrows, bool residual) {
    ASAC_SMEM(dZ); ASAC_SMEM(Wt); ASAC_SMEM(dY); ASAC_SMEM(dX);
    const int tid = threadIdx.x, cg = tid & 15, rg = tid >> 4;
    const int ldw = H + 4;
    for (int r0 = 0; r0 < nrows; r0 += PASS_ROWS) {
        const int r = r0 + 2 * rg;
        float acc[2][CM];
#pragma unroll
        for (int c = 0; c < CM; ++c) acc[0][c] = acc[1][c] = 0.f;
        gemm_core<CM>(dZ, ld, H, Wt, ldw, r, cg, acc);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
            for (int c = 0; c < CM; ++c) {
                const int k = cg + 16 * c;
                float v = acc[i][c];
                if (residual) v = v + dY[(r + i) * ld + k];
                dX[(r + i) * ld + k] = v;
            }
        }
    }
}

__device__ __forceinline__ void layer_input_grad(int hidden, const float *dZ, int ld, const float *Wt,
                                                 const float *dY, float *dX, int nrows, bool residual) {
    switch (hidden >> 4) {
        case 1: layer_input_grad_t<1>(dZ, ld, hidden, Wt, dY, dX, nrows, residual); break;
        case 2: layer_input_grad_t<2>(dZ, ld, hidden, Wt, dY, dX, nrows, residual); break;
        case 4: layer_input_grad_t<4>(dZ, ld, hidden, Wt, dY, dX, nrows, residual); break;
        default: layer_input_grad_t<8>(dZ, ld, hidden, Wt, dY, dX, nrows, residual); break;
    }
}

// dZ = dY * gelu'(Z)   (in place over dY), rows beyond `valid_rows` are zeroed
__device__ __forceinline__ void gelu_backward(float *dY, const float *Z, int ld, int H, int nrows, int valid_rows) {
    for (int i = threadIdx.x; i < nrows * H; i += NT) {
        const int r = i / H, j = i - r * H;
        const float g = dY[r * ld + j];
        dY[r * ld + j] = r < valid_rows ? g * gelu_erf_grad(Z[r * ld + j]) : 0.f;
    }
}

// Partial weight / bias gradient of one layer over the tile's rows:
//   gW[j * K + k] = sum_r dZ[r][j] * X[r][k],  gb[j] = sum_r dZ[r][j]
// 4x4 output blocks, one float4 of dZ and one of X per row.
__device__ __noinline__ void layer_weight_grad(const float *dZ, int ldz, const float *X, int ldx, int H, int K,
                                               int nrows, float *gW, float *gb) {
    ASAC_SMEM(dZ); ASAC_SMEM(X);
    const int tid = threadIdx.x;
    const int K4 = round_up(K, 4);
    const int nJB = H >> 2, nKB = K4 >> 2;
    const bool vec = ((K & 3) == 0) && ((((uintptr_t)gW) & 15) == 0);
    for (int blk = tid; blk < nJB * nKB; blk += NT) {
        const int jb = blk % nJB, kb = blk / nJB;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        for (int r = 0; r < nrows; ++r) {
            const float4 dz = *reinterpret_cast<const float4 *>(dZ + r * ldz + 4 * jb);
            const float4 x = *reinterpret_cast<const float4 *>(X + r * ldx + 4 * kb);
            const float dzv[4] = {dz.x, dz.y, dz.z, dz.w};
            const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(dzv[a], xv[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float *dst = gW + (int64_t)(4 * jb + a) * K + 4 * kb;
            if (vec) {
                *reinterpret_cast<float4 *>(dst) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (4 * kb + b < K) dst[b] = acc[a][b];
            }
        }
    }
    for (int j = tid; j < H; j += NT) {
        float s = 0.f;
        for (int r = 0; r < nrows; ++r) s += dZ[r * ldz + j];
        gb[j] = s;
    }
}

// Linear head: out[r * O + o] = X[r] . Wh[o] + bh[o]   (Wh, bh in global memory, 8 lanes per dot)
__device__ __noinline__ void head_forward(const float *X, int ldx, int H, const float *Wh, const float *bh,
                                          int O, int nrows, float *out) {
    ASAC_SMEM(X); ASAC_SMEM(out);
    const int tid = threadIdx.x, grp = tid >> 3, sub = tid & 7;
    const int total = nrows * O;
    for (int d0 = 0; d0 < total; d0 += NT / 8) {
        const int d = d0 + grp;
        float s = 0.f;
        if (d < total) {
            const int r = d / O, o = d - r * O;
            const float *x = X + r * ldx;
            const float *w = Wh + (int64_t)o * H;
            for (int k = sub; k < H; k += 8) s = fmaf(x[k], __ldg(w + k), s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (d < total && sub == 0) out[d] = s + __ldg(bh + (d % O));
    }
}

// Head backward.  dO[r * O + o] (rows >= valid rows must be zero):
//   gWh[o * H + j] = sum_r dO[r][o] X[r][j];  gbh[o] = sum_r dO[r][o];  dH[r][j] = sum_o dO[r][o] Wh[o][j]
__device__ __noinline__ void head_backward(const float *dO, int O, const float *X, int ldx, int H,
                                           const float *Wh, int nrows, float *gWh, float *gbh, float *dH,
                                           int ldh) {
    ASAC_SMEM(dO); ASAC_SMEM(dH);
    if (X) ASAC_SMEM(X);
    const int tid = threadIdx.x;
    if (gWh) {
        for (int i = tid; i < O * H; i += NT) {
            const int o = i / H, j = i - o * H;
            float s = 0.f;
            for (int r = 0; r < nrows; ++r) s = fmaf(dO[r * O + o], X[r * ldx + j], s);
            gWh[i] = s;
        }
        for (int o = tid; o < O; o += NT) {
            float s = 0.f;
            for (int r = 0; r < nrows; ++r) s += dO[r * O + o];
            gbh[o] = s;
        }
    }
    for (int i = tid; i < nrows * H; i += NT) {
        const int r = i / H, j = i - r * H;
        float s = 0.f;
        for (int o = 0; o < O; ++o) s = fmaf(dO[r * O + o], __ldg(Wh + (int64_t)o * H + j), s);
        dH[r * ldh + j] = s;
    }
}

// ---------------------------------------------------------------- whole-net forward
// Shared-memory plan of a tile kernel (all sizes in floats):
//   act buffers : nrows * lda each, lda = max(H, round_up(in,4)) + 4
//   weight ring : 2 x wsz, wsz = H * (max(H, round_up(in,4)) + 4) + H
struct TileSmem {
    float *w[2];   // weight ring
    float *bias[2];
};

__host__ __device__ __forceinline__ int tile_lda(int hidden, int in_dim) {
    const int k4 = round_up(in_dim, 4);
    return (hidden > k4 ? hidden : k4) + 4;
}
__host__ __device__ __forceinline__ int tile_wsz(int hidden, int in_dim) {
    return hidden * tile_lda(hidden, in_dim) + hidden;
}

// Runs the `depth` ResBlocks of a stock net over rows held in `x0` (nrows x lda, input columns
// [0, in_dim) valid, pad columns up to round_up(in,4) zero).
//   save == nullptr : ping-pongs between bufA and bufB, returns the buffer holding the output
//   save != nullptr : layer l reads save_x[l] and writes z to save_z[l], y to save_x[l+1]
// The caller must have issued no outstanding cp.async groups.
__device__ __forceinline__ float *net_trunk_forward(const NetShape &s, const float *params, const TileSmem &sm,
                                                    float *x0, float *bufA, float *bufB, float **save_x,
                                                    float **save_z, int lda, int nrows) {
    const int H = s.hidden;
    stage_weights(sm.w[0], sm.bias[0], params + net_w_off(s, 0), params + net_b_off(s, 0), H, s.in_dim, false);
    cp_async_commit();
    float *x = x0;
    for (int l = 0; l < s.depth; ++l) {
        if (l + 1 < s.depth) {
            stage_weights(sm.w[(l + 1) & 1], sm.bias[(l + 1) & 1], params + net_w_off(s, l + 1),
                          params + net_b_off(s, l + 1), H, H, false);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int K = net_k(s, l);
        float *y = save_x ? save_x[l + 1] : (x == bufA ? bufB : bufA);
        layer_forward(H, x, lda, round_up(K, 4), sm.w[l & 1], sm.bias[l & 1], save_z ? save_z[l] : nullptr, y, lda,
                      nrows, K == H);
        __syncthreads();
        x = y;
    }
    return x;
}

}  // namespace asac
