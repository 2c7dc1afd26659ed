// Micro-benchmark: a chain of D ResBlock layers (H = 64) over ONE 16-row tile per CTA, the unit every
// update kernel of the step is made of.  Variants of the layer routine are timed with clock64 (cycles per
// layer, CTA 0) and checked against an fp64 host evaluation of the same chain.
//   0: the FFMA K-split routine of mlp_tile.cuh (layer_forward_t<4, false>)
//   1: mma.sync m16n8k8 3xTF32, 8 warps x full K, epilogue in registers, one barrier per layer
//   2: mma.sync m16n8k8 3xTF32, 16 warps x K halves, smem reduction
//   4: 0 with the two epilogue outputs of a thread interleaved;  5 / 6: register tile 4 x 2 / 8 x 1 (see mlp_tile.cuh)
//   3: as 1, warps 8..15 run a second independent 16-row tile (the same rows again)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I../../include
//        -I../../advanced-soft-actor-critic_b200/asac_b200/csrc layer_chain.cu -o build/layer_chain
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "mlp_tile.cuh"
#include "mma_tile.cuh"

namespace asac {
void set_error(const char *, ...) {}
void count_launch(int) {}
}  // namespace asac
using namespace asac;

constexpr int H = 64, LDW = 68, LDA = 68, SLOT = 64 * LDW + 64;

template <int V>
__global__ void __launch_bounds__(NT) k_chain(const float *W, const float *bias, const float *X, float *Y, long long *cycles,
                                              int D) {
    extern __shared__ float4 smem4[];
    float *sm = reinterpret_cast<float *>(smem4);
    float *Ws = sm;                       // D slots
    float *xa = Ws + D * SLOT;            // [16][LDA] x 2 planes
    float *xb = xa + 2 * 16 * LDA;
    float *part = xb + 2 * 16 * LDA;      // 4 x 16 x 64
    const int tid = threadIdx.x;
    for (int i = tid; i < D * 64 * 64; i += NT) {
        const int l = i / 4096, r = (i / 64) % 64, k = i % 64;
        Ws[l * SLOT + r * LDW + k] = W[i];
    }
    for (int i = tid; i < D * 64; i += NT) Ws[(i / 64) * SLOT + 64 * LDW + (i % 64)] = bias[i];
    for (int i = tid; i < 16 * 64; i += NT) {
        const float v = X[(int64_t)blockIdx.x * 1024 + i];
        xa[(i / 64) * LDA + (i % 64)] = v;
        if (V == 3) xa[(16 + i / 64) * LDA + (i % 64)] = v;
    }
    __syncthreads();
    float *x = xa, *y = xb;
    // variants 7 / 8: split planes.  X: [fp32 | hi | lo] x 16 rows in `part` (x) and after it (y); W lo planes in place of
    // the upper half of the slots is not possible, so they live behind everything (extra shared memory)
    float *px = part, *py = part + 3 * 16 * LDA, *wlo = part + 6 * 16 * LDA;
    if (V >= 7) {
        for (int i = tid; i < 16 * 64; i += NT) {
            const float v = xa[(i / 64) * LDA + (i % 64)];
            uint32_t h, l; split_tf32(v, h, l);
            px[(i / 64) * LDA + (i % 64)] = v;
            px[16 * LDA + (i / 64) * LDA + (i % 64)] = __uint_as_float(h);
            px[32 * LDA + (i / 64) * LDA + (i % 64)] = __uint_as_float(l);
        }
        if (V == 8)
            for (int i = tid; i < D * 64 * 64; i += NT) {
                const int l = i / 4096, r = (i / 64) % 64, k = i % 64;
                uint32_t h, lo; split_tf32(Ws[l * SLOT + r * LDW + k], h, lo);
                Ws[l * SLOT + r * LDW + k] = __uint_as_float(h);
                wlo[l * 64 * LDW + r * LDW + k] = __uint_as_float(lo);
            }
        __syncthreads();
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int l = 0; l < D; ++l) {
        const float *w = Ws + l * SLOT, *b = w + 64 * LDW;
        if (V == 0) layer_forward_t<4, false>(x, LDA, 64, w, b, nullptr, y, LDA, 16, true, part);
        if (V == 1) mma_layer_forward<8, false, false>(x, LDA, w, LDW, b, nullptr, y, LDA, true, part);
        if (V == 2) mma_layer_forward<16, false, false>(x, LDA, w, LDW, b, nullptr, y, LDA, true, part);
        if (V == 3) mma_layer_forward<8, false, true>(x, LDA, w, LDW, b, nullptr, y, LDA, true, part);
        if (V == 4) layer_forward_hh<4, false, 2>(x, LDA, w, b, nullptr, y, LDA, 16, true, part);
        if (V == 5) layer_forward_hh<4, false, 4>(x, LDA, w, b, nullptr, y, LDA, 16, true, part);
        if (V == 6) layer_forward_hh<4, false, 8>(x, LDA, w, b, nullptr, y, LDA, 16, true, part);
        if (V == 7) mma_layer_forward_ps<false>(px, px + 16 * LDA, 16 * LDA, LDA, w, 0, LDW, b, nullptr, py, py + 16 * LDA, LDA, true);
        if (V == 8) mma_layer_forward_ps<true>(px, px + 16 * LDA, 16 * LDA, LDA, w, (int)(wlo + l * 64 * LDW - w), LDW, b, nullptr, py, py + 16 * LDA, LDA, true);
        if (V >= 7) { float *t2 = px; px = py; py = t2; }
        float *t = x; x = y; y = t;
    }
    const long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (V >= 7) x = px;
    for (int i = tid; i < 16 * 64; i += NT) Y[(int64_t)blockIdx.x * 1024 + i] = x[(i / 64) * LDA + (i % 64)];
}

static double gelu64(double z) { return 0.5 * z * (1.0 + erf(z * 0.70710678118654752440)); }

template <int V>
static void run(const char *name, int grid, int D, const float *dW, const float *db, const float *dX, float *dY,
                long long *dC, const std::vector<double> &ref, double scale) {
    const size_t smem = (size_t)(D * SLOT + 4 * 16 * LDA + (V >= 7 ? 6 * 16 * LDA + (V == 8 ? D * 64 * LDW : 0) : 4 * 16 * 64)) * 4;
    cudaFuncSetAttribute(k_chain<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) k_chain<V><<<grid, NT, smem>>>(dW, db, dX, dY, dC, D);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) k_chain<V><<<grid, NT, smem>>>(dW, db, dX, dY, dC, D);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<float> y((size_t)grid * 1024);
    std::vector<long long> c(grid);
    cudaMemcpy(y.data(), dY, y.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(c.data(), dC, grid * 8, cudaMemcpyDeviceToHost);
    double maxe = 0, rms = 0;
    for (size_t i = 0; i < y.size(); ++i) { const double e = fabs(y[i] - ref[i]); maxe = fmax(maxe, e); rms += e * e; }
    printf("%-44s grid %3d D %2d: %7.1f cycles/layer (CTA 0), kernel %6.2f us, max err %.3g rms %.3g of scale %.3g\n", name, grid,
           D, (double)c[0] / D, ms * 1000 / 20, maxe / scale, sqrt(rms / y.size()) / scale, scale);
#ifdef ASAC_PROBES
    long long seg[8], zero[8] = {0};
    cudaMemcpyFromSymbol(seg, g_layer_seg, sizeof(seg));
    cudaMemcpyToSymbol(g_layer_seg, zero, sizeof(zero));
    if (seg[4]) printf("    segments per pass: GEMM+partials %.0f, barrier %.0f, epilogue %.0f, barrier %.0f cycles\n",
                       (double)seg[0] / seg[4], (double)seg[1] / seg[4], (double)seg[2] / seg[4], (double)seg[3] / seg[4]);
#endif
}

int main(int argc, char **argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;  // one variant (for ncu)
    const int grid = 128, D = 5;  // (5 layers: variant 8 keeps a second plane of every weight in shared memory)
    std::vector<float> W((size_t)D * 4096), b(D * 64), X((size_t)grid * 1024);
    srand(1);
    auto u = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto &v : W) v = u() * 0.125f;  // nn.Linear default: U(-1/sqrt(K), 1/sqrt(K))
    for (auto &v : b) v = u() * 0.125f;
    for (auto &v : X) v = u();
    std::vector<double> ref(X.begin(), X.end());
    for (int g = 0; g < grid; ++g)
        for (int l = 0; l < D; ++l) {
            double y[1024];
            for (int r = 0; r < 16; ++r)
                for (int j = 0; j < 64; ++j) {
                    double z = b[l * 64 + j];
                    for (int k = 0; k < 64; ++k) z += ref[g * 1024 + r * 64 + k] * (double)W[l * 4096 + j * 64 + k];
                    y[r * 64 + j] = gelu64(z) + ref[g * 1024 + r * 64 + j];
                }
            for (int i = 0; i < 1024; ++i) ref[g * 1024 + i] = y[i];
        }
    double scale = 0;
    for (double v : ref) scale = fmax(scale, fabs(v));
    float *dW, *db, *dX, *dY; long long *dC;
    cudaMalloc(&dW, W.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dY, X.size() * 4);
    cudaMalloc(&dC, grid * 8);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    for (int g : {1, 128}) {
        if (only < 0 || only == 0) run<0>("0 FFMA K-split (mlp_tile.cuh)", g, D, dW, db, dX, dY, dC, ref, scale);
        if (only < 0 || only == 1) run<1>("1 mma.sync 3xTF32, 8 warps full K", g, D, dW, db, dX, dY, dC, ref, scale);
        if (only < 0 || only == 2) run<2>("2 mma.sync 3xTF32, 16 warps K halves", g, D, dW, db, dX, dY, dC, ref, scale);
        if (only < 0 || only == 3) run<3>("3 mma.sync 3xTF32, two 16-row tiles", g, D, dW, db, dX, dY, dC, ref, scale);
        if (only < 0 || only == 4) run<4>("4 FFMA K-split 2 rows x 4 cols, epilogue x2", g, D, dW, db, dX, dY, dC, ref, scale);
        if (only < 0 || only == 5) run<5>("5 FFMA K-split 4 rows x 2 cols (cg = lane)", g, D, dW, db, dX, dY, dC, ref, scale);
        if (only < 0 || only == 7) run<7>("7 mma.sync 3xTF32, A planes pre-split", g, D, dW, db, dX, dY, dC, ref, scale);
        if (only < 0 || only == 8) run<8>("8 mma.sync 3xTF32, A and W planes pre-split", g, D, dW, db, dX, dY, dC, ref, scale);
        if (only < 0 || only == 6) run<6>("6 FFMA K-split 8 rows x 1 col", g, D, dW, db, dX, dY, dC, ref, scale);
    }
    return 0;
}
