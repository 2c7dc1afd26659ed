// Row-tile MLP building blocks for the SAC update kernels (sm_100a, fp32 FFMA path).
//
// A CTA of NT = 512 threads (16 warps, 4 per SM sub-partition) owns a tile of <= 16 rows whose
// activations live in shared memory and walks the net layer by layer.
//
//   * Weights: every layer the kernel will need is a "job" of a WeightPipe, staged ahead of its
//     use into a slot guarded by an mbarrier; consumers wait on the barrier's phase, so the staging
//     of later layers overlaps the math of the current one.  Hidden->hidden layers (K a multiple of
//     32) arrive by TMA: ONE thread issues one `cp.async.bulk.tensor` per 32-column segment from a
//     4-D tensor map over {k, row, layer, ensemble member} of the flat parameter buffer, with the
//     128-byte swizzle, so the [rows][32] tiles land bank-conflict-free without padding and the
//     readers only XOR the 16-byte chunk index with (row & 7).  First layers (K = 6, 8, ...) are
//     copied by all threads with cp.async into padded rows.  Slots are recycled round robin when
//     shared memory cannot hold all jobs.
//   * GEMM core: thread (cg, rg, ks) = (tid & 15, (tid >> 4) & 7, tid >> 7) accumulates rows
//     {2rg, 2rg+1} x CM columns over the ks-th quarter of K with float4 operand loads; the four
//     K-partials meet in shared memory and the epilogue (bias, exact-erf GELU, residual) runs on
//     all 512 threads, two outputs each.  At replay batch sizes the step is latency-bound (one
//     16-row tile per SM): 4 warps per scheduler hide the LDS->FFMA latency that a 128-thread
//     CTA exposed (ncu: 'short_scoreboard' on the first FFMA after the loads).
//
// Numerics follow the reference's torch fp32 ops: nn.Linear + exact-erf GELU + residual
// (algorithm/nn_models/layers/linear_layers.py:46-56).  The library is compiled with
// -fmad=false, so only the explicit fmaf() of the GEMM cores contract.
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the encoder is resolved at run time, see sac.cu)

#include "common.cuh"

namespace asac {

constexpr int NT = 512;        // threads per CTA in every tiled kernel
constexpr int KSPLIT = 4;      // K is split over tid >> 7
constexpr int PASS_ROWS = 16;  // rows per GEMM pass

__host__ __device__ __forceinline__ int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Code placement: a tile kernel walks ~10 layers through three or four call sites.  With everything
// inlined (x4 width instantiations per site) k_value_pass grew to 18 K SASS instructions (290 KB)
// and > 50 % of the warp stall samples were `no_instruction` (instruction-cache misses), so the
// per-layer routines are out of line: one copy per width keeps the hot loop resident.  (Measured
// alternative, same box: out-of-line TRUNK walks with the layers inlined into them — one call per
// net instead of one per layer — were 2.2 % slower.)  Pointer PARAMETERS that live in shared memory
// are declared with ASAC_SMEM so the out-of-line code still uses LDS/STS.  Hints on pointers that
// are derived inside inlined code are not safe: the compiler folded one to false and deleted a
// whole trunk walk.
#define ASAC_SMEM(p) __builtin_assume(__isShared(p))

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

// ---------------------------------------------------------------- mbarrier / TMA primitives
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared (16-byte aligned addresses, size a multiple of 16)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
// the mbarrier receives one arrival when all prior cp.async of this thread have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- weight pipe
struct WeightJob {
    const float *W;  // [N, K] row-major (global)
    const float *b;  // [N] or null
    int N, K;
    const CUtensorMap *map;  // non-null: stage by TMA (K % 32 == 0); coordinates {k, row, layer, net}
    int layer, net;
};
constexpr int MAX_WEIGHT_JOBS = 20;
constexpr int MAX_WEIGHT_SLOTS = 16;
constexpr int PIPE_LOOKAHEAD = MAX_WEIGHT_SLOTS;

// shared-memory row stride of a staged [N, K] matrix: K rounded to 4 plus 4 floats, so that the
// 16-byte operand loads of 8 consecutive rows fall into distinct banks (stride = 4 mod 32 for K = 64)
__host__ __device__ __forceinline__ int weight_ld(int K) { return round_up(K, 4) + 4; }
// slots start on 1024-byte boundaries (the 128-byte swizzle pattern is a function of address bits 7-9)
__host__ __device__ __forceinline__ int weight_slot_floats(int N, int K) { return round_up(N * weight_ld(K) + N, 256); }
// float offset of element (n, k) of a TMA-staged [N][K] matrix: K / 32 segments of [N][32] floats, the
// 16-byte chunk index XORed with (n & 7)  (CU_TENSOR_MAP_SWIZZLE_128B)
__device__ __forceinline__ int swz_off(int N, int n, int k) {
    return (k >> 5) * (N * 32) + n * 32 + ((((k >> 2) & 7) ^ (n & 7)) << 2) + (k & 3);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, int c3,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
            "r"(smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

struct WeightPipe {
    float *slots;      // n_slots x slot_floats
    uint64_t *bars;    // n_slots mbarriers (NT deferred arrivals each, one per thread, plus the TMA bytes)
    WeightJob *jobs;   // shared-memory job table
    int n_slots, slot_floats, n_jobs, issued, consumed;
    int n_open;  // jobs [0, n_open) may be issued (a kernel that starts ahead of its predecessor under programmatic
                 // dependent launch opens the jobs that read the predecessor's output after griddepcontrol.wait)
};

// Issues every job whose slot is free; returns the new `issued`.  Executed by ALL threads at
// CTA-uniform points, after a __syncthreads() that follows the last read of any slot being
// recycled.  Hidden -> hidden layers: thread 0 registers the bytes (expect_tx) and issues one TMA
// tensor copy per 32-column segment.  First layers and biases: every thread copies its share with
// cp.async.  Each thread then posts one deferred arrival on the slot's mbarrier
// (`cp.async.mbarrier.arrive.noinc`: it fires when that thread's copies have landed), so a slot is
// complete after NT arrivals plus the TMA bytes.
// (History, each variant measured on the GPU: one cp.async.bulk per weight ROW from one warp made
// the consumers' wait the largest stall of the value pass; staging from warp 0 alone lengthened
// the kernels' setup phase from 6.8 to 8.0 us; limiting the lookahead only moved time between
// phases.  The weights are not the bottleneck: CTA (0,0) waits 3.4 us per STEP on them.)
static __device__ __noinline__ int pipe_fill_impl(float *slots, uint64_t *bars, const WeightJob *jobs, int n_slots,
                                           int slot_floats, int n_jobs, int issued, int consumed) {
    ASAC_SMEM(slots); ASAC_SMEM(jobs);
    const int tid = threadIdx.x;
    const int ahead = n_slots < PIPE_LOOKAHEAD ? n_slots : PIPE_LOOKAHEAD;
#pragma unroll 1
    while (issued < n_jobs && issued - consumed < ahead) {
        const WeightJob j = jobs[issued];
        const int slot = issued % n_slots;
        float *Ws = slots + (int64_t)slot * slot_floats;
        float *bs;
        if (j.map) {
            // TMA: K / 32 tensor copies of [N][32] floats, 128-byte swizzle
            bs = Ws + j.N * j.K;
            // (issued by lane 0 of warp `issued mod 16`: the initial fill of ~10 jobs from thread 0 alone was 2.2 us of
            // serial expect_tx + TMA issue in every kernel's setup)
            if (tid == ((issued & (NT / 32 - 1)) << 5)) {
                // generic-proxy reads of a RECYCLED slot precede the async-proxy writes (a slot's first use has none)
                if (issued >= n_slots) fence_proxy_async();
                mbar_expect_tx(bars + slot, (unsigned)(j.N * j.K * 4));
                for (int seg = 0; seg < (j.K >> 5); ++seg)
                    tma_load_4d(Ws + seg * j.N * 32, j.map, seg * 32, 0, j.layer, j.net, bars + slot);
            }
        } else {
            const int ldw = weight_ld(j.K);
            bs = Ws + j.N * ldw;
            if (((j.K & 3) == 0) && ((((uintptr_t)j.W) & 15) == 0)) {
                const int per_row = j.K >> 2, total = j.N * per_row;
#pragma unroll 1
                for (int i = tid; i < total; i += NT) {
                    const int n = i / per_row, c = i - n * per_row;
                    cp_async16(Ws + n * ldw + 4 * c, j.W + (int64_t)n * j.K + 4 * c);
                }
            } else {
                // rows that are not 16-byte multiples (e.g. K = 6): 4-byte copies, zeroed pad columns
                const int K4 = round_up(j.K, 4), pad = K4 - j.K;
#pragma unroll 1
                for (int i = tid; i < j.N * j.K; i += NT) {
                    const int n = i / j.K, k = i - n * j.K;
                    cp_async4(Ws + n * ldw + k, j.W + i);
                }
#pragma unroll 1
                for (int i = tid; i < j.N * pad; i += NT) {
                    const int n = i / pad;
                    Ws[n * ldw + j.K + (i - n * pad)] = 0.f;
                }
            }
        }
        if (j.b) {
#pragma unroll 1
            for (int n = tid; n < j.N; n += NT) cp_async4(bs + n, j.b + n);
        }
        cp_async_mbar_arrive(bars + slot);
        ++issued;
    }
    return issued;
}

// all threads; `jobs` must already be written (any thread) and published by a CTA barrier
__device__ __forceinline__ void pipe_init(WeightPipe &p, float *slots, uint64_t *bars, WeightJob *jobs, int n_slots,
                                          int slot_floats, int n_jobs, int n_open = -1) {
    p.slots = slots; p.bars = bars; p.jobs = jobs;
    p.n_slots = n_slots < MAX_WEIGHT_SLOTS ? n_slots : MAX_WEIGHT_SLOTS;
    p.slot_floats = slot_floats; p.n_jobs = n_jobs; p.issued = 0; p.consumed = 0;
    p.n_open = n_open < 0 ? n_jobs : n_open;
    if ((int)threadIdx.x < p.n_slots) {
        mbar_init(bars + threadIdx.x, NT);
        fence_mbar_init();
    }
    __syncthreads();
    p.issued = pipe_fill_impl(p.slots, p.bars, p.jobs, p.n_slots, p.slot_floats, p.n_open, p.issued, p.consumed);
}
// all threads, CTA-uniform: every job may be issued from here on
__device__ __forceinline__ void pipe_open_all(WeightPipe &p) {
    p.n_open = p.n_jobs;
    p.issued = pipe_fill_impl(p.slots, p.bars, p.jobs, p.n_slots, p.slot_floats, p.n_open, p.issued, p.consumed);
}

// waits for the oldest unconsumed job; returns its staged weights / bias
// debug counters of CTA (0,0), thread 0: cycles spent waiting for staged weights / number of waits
#ifdef ASAC_PROBES
#define ASAC_PROBE_ON 1
#else
#define ASAC_PROBE_ON 0
#endif
// (compiled in with -DASAC_PROBES only: the read-modify-write of a global counter by thread 0 after every
// layer stalls the whole CTA at the next barrier — 2.7 % of the step when it was always on)
static __device__ long long g_pipe_wait[2];
static __device__ long long g_layer_seg[8];  // layer_forward segments of CTA (0,0), thread 0 (debug)
__device__ __forceinline__ void pipe_acquire(const WeightPipe &p, const float *&Ws, const float *&bs) {
    const int slot = p.consumed % p.n_slots;
    const unsigned parity = (unsigned)((p.consumed / p.n_slots) & 1);
    const bool probe = ASAC_PROBE_ON && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0;
    const long long c0 = probe ? clock64() : 0;
    mbar_wait(p.bars + slot, parity);
    if (probe) {
        g_pipe_wait[0] += clock64() - c0;
        g_pipe_wait[1] += 1;
    }
    const WeightJob &j = p.jobs[p.consumed];
    Ws = p.slots + (int64_t)slot * p.slot_floats;
    bs = Ws + (j.map ? j.N * j.K : j.N * weight_ld(j.K));
}
__device__ __forceinline__ bool pipe_front_swizzled(const WeightPipe &p) { return p.jobs[p.consumed].map != nullptr; }
// the CTA has finished reading the oldest job (a __syncthreads() must separate the reads from this
// call); refills the freed slot
__device__ __forceinline__ void pipe_release(WeightPipe &p) {
    ++p.consumed;
    if (p.issued < p.n_open)
        p.issued = pipe_fill_impl(p.slots, p.bars, p.jobs, p.n_slots, p.slot_floats, p.n_open, p.issued, p.consumed);
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

// ---------------------------------------------------------------- forward layer
// ResBlock forward over `nrows` (multiple of 16) rows:
//   Z = X . W^T + b ;  Y = gelu(Z) (+ X when residual)        Zs may be null (no backward)
// Ws: [H rows][ldw = K4 + 4] staged weights, `part`: KSPLIT x 16 x H floats of scratch.
template <int CM, bool SWZ>
__device__ __noinline__ void layer_forward_t(const float *X, int ldx, int K4, const float *Ws, const float *bs,
                                             float *Zs, float *Ys, int ldy, int nrows, bool residual, float *part) {
    ASAC_SMEM(X); ASAC_SMEM(Ws); ASAC_SMEM(bs); ASAC_SMEM(Ys); ASAC_SMEM(part);
    constexpr int H = 16 * CM;
    const int tid = threadIdx.x, cg = tid & 15, rg = (tid >> 4) & 7, ks = tid >> 7;
    const int ldw = K4 + 4;
    const int kchunk = round_up((K4 + KSPLIT - 1) / KSPLIT, 4);
    const int kb = ks * kchunk, ke = min(K4, kb + kchunk);
    const bool probe = ASAC_PROBE_ON && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0;
#pragma unroll 1
    for (int r0 = 0; r0 < nrows; r0 += PASS_ROWS) {
        const long long c0 = probe ? clock64() : 0;
        const int r = r0 + 2 * rg;
        float acc[2][CM];
#pragma unroll
        for (int c = 0; c < CM; ++c) acc[0][c] = acc[1][c] = 0.f;
        const float *a0p = X + r * ldx, *a1p = a0p + ldx, *wp = Ws + cg * ldw;
        // TMA-staged weights: K4 == H, segments of [H][32] floats, chunk index XOR (row & 7); the rows of
        // this thread are cg + 16 c, so (row & 7) == (cg & 7) for all of them
        const float *wsz = Ws + cg * 32;
        const int x7 = cg & 7;
#pragma unroll 4
        for (int k = kb; k < ke; k += 4) {
            const float4 a0 = *reinterpret_cast<const float4 *>(a0p + k);
            const float4 a1 = *reinterpret_cast<const float4 *>(a1p + k);
            const int koff = SWZ ? (k >> 5) * (H * 32) + ((((k >> 2) & 7) ^ x7) << 2) : k;
#pragma unroll
            for (int c = 0; c < CM; ++c) {
                const float4 w = SWZ ? *reinterpret_cast<const float4 *>(wsz + 16 * c * 32 + koff)
                                     : *reinterpret_cast<const float4 *>(wp + 16 * c * ldw + koff);
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
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int c = 0; c < CM; ++c) part[(ks * PASS_ROWS + 2 * rg + i) * H + cg + 16 * c] = acc[i][c];
        const long long c1 = probe ? clock64() : 0;
        __syncthreads();
        const long long c2 = probe ? clock64() : 0;
#pragma unroll 1
        for (int o = tid; o < PASS_ROWS * H; o += NT) {
            const int row = o / H, j = o - row * H;
            float z = bs[j];
#pragma unroll
            for (int s = 0; s < KSPLIT; ++s) z += part[s * PASS_ROWS * H + o];
            if (Zs) Zs[(r0 + row) * ldy + j] = z;
            float y = gelu_erf(z);
            if (residual) y = y + X[(r0 + row) * ldx + j];
            Ys[(r0 + row) * ldy + j] = y;
        }
        const long long c3 = probe ? clock64() : 0;
        __syncthreads();
        if (probe) {
            g_layer_seg[0] += c1 - c0;            // K-split GEMM + partial stores
            g_layer_seg[1] += c2 - c1;            // barrier
            g_layer_seg[2] += c3 - c2;            // epilogue
            g_layer_seg[3] += clock64() - c3;     // barrier
            g_layer_seg[4] += 1;
        }
    }
}

// Hidden -> hidden pass (K4 == H) with a register tile of RT rows x CT columns per thread (RT * CT * KSPLIT * NT ==
// 16 * H * KSPLIT ... i.e. RT * CT == 2 * H / 16).  A 128-bit shared load is served one quarter-warp (8 lanes, 128 bytes) per
// wavefront: the mapping of layer_forward_t (2 x CM, 16 column groups per warp) spends 4 wavefronts on a weight
// load that fetches 256 distinct bytes and the GEMM phase is bound by those wavefronts (1300 cycles for 512 FFMA
// issue slots, measured).  With the column group = lane (CT = 2 at H = 64) every weight load fetches 512 distinct
// bytes and the RT row loads are warp-uniform broadcasts.  Same arithmetic per output in the same order as
// layer_forward_t (k ascending inside a K quarter, quarters added in order): results are bit-identical.
template <int CM, bool SWZ, int RT>
__device__ __noinline__ void layer_forward_hh(const float *X, int ldx, const float *Ws, const float *bs, float *Zs,
                                              float *Ys, int ldy, int nrows, bool residual, float *part) {
    ASAC_SMEM(X); ASAC_SMEM(Ws); ASAC_SMEM(bs); ASAC_SMEM(Ys); ASAC_SMEM(part);
    constexpr int H = 16 * CM, K4 = H, ldw = K4 + 4, KQ = K4 / KSPLIT / 4;  // float4 steps per thread
    constexpr int CT = 2 * CM / RT, NCG = H / CT, NRG = PASS_ROWS / RT;
    static_assert(CT >= 1 && CT * RT == 2 * CM && NCG * NRG * KSPLIT == NT, "register tile");
    const int tid = threadIdx.x, cg = tid % NCG, rg = (tid / NCG) % NRG, ks = tid / (NCG * NRG);
    const int kb = ks * (K4 / KSPLIT);
#pragma unroll 1
    for (int r0 = 0; r0 < nrows; r0 += PASS_ROWS) {
        const float *ap = X + (r0 + RT * rg) * ldx + kb, *wp = Ws + cg * ldw + kb;
        const int x7 = cg & 7;
        float acc[RT][CT];
#pragma unroll
        for (int i = 0; i < RT; ++i)
#pragma unroll
            for (int c = 0; c < CT; ++c) acc[i][c] = 0.f;
#pragma unroll
        for (int q = 0; q < KQ; ++q) {
            const int k = kb + 4 * q;
            float4 a[RT], w[CT];
#pragma unroll
            for (int i = 0; i < RT; ++i) a[i] = *reinterpret_cast<const float4 *>(ap + i * ldx + 4 * q);
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                // rows cg + NCG c: (row & 7) == (cg & 7) since NCG is a multiple of 8
                const int row = cg + NCG * c;
                w[c] = SWZ ? *reinterpret_cast<const float4 *>(Ws + (k >> 5) * (H * 32) + row * 32 + ((((k >> 2) & 7) ^ x7) << 2))
                           : *reinterpret_cast<const float4 *>(wp + NCG * c * ldw + 4 * q);
            }
#pragma unroll
            for (int c = 0; c < CT; ++c)
#pragma unroll
                for (int i = 0; i < RT; ++i) {
                    acc[i][c] = fmaf(a[i].x, w[c].x, acc[i][c]);
                    acc[i][c] = fmaf(a[i].y, w[c].y, acc[i][c]);
                    acc[i][c] = fmaf(a[i].z, w[c].z, acc[i][c]);
                    acc[i][c] = fmaf(a[i].w, w[c].w, acc[i][c]);
                }
        }
#pragma unroll
        for (int i = 0; i < RT; ++i)
#pragma unroll
            for (int c = 0; c < CT; ++c) part[(ks * PASS_ROWS + RT * rg + i) * H + cg + NCG * c] = acc[i][c];
        __syncthreads();
        // epilogue: outputs tid, tid + NT, ... of the pass, their dependent chains (partials, erf) interleaved
        constexpr int PER = PASS_ROWS * H / NT;  // 2 at H = 64
        float z[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int o = tid + u * NT, j = o & (H - 1);
            z[u] = bs[j];
#pragma unroll
            for (int s = 0; s < KSPLIT; ++s) z[u] += part[s * PASS_ROWS * H + o];
        }
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int o = tid + u * NT, row = o / H, j = o & (H - 1);
            if (Zs) Zs[(r0 + row) * ldy + j] = z[u];
            float y = gelu_erf(z[u]);
            if (residual) y = y + X[(r0 + row) * ldx + j];
            Ys[(r0 + row) * ldy + j] = y;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void layer_forward(int hidden, const float *X, int ldx, int K4, const float *Ws,
                                              const float *bs, float *Zs, float *Ys, int ldy, int nrows,
                                              bool residual, float *part, bool swizzled) {
    if (hidden == 64 && K4 == 64) {  // the stock width: 4 x 2 register tile (bit-identical, 11 % fewer cycles per pass)
        if (swizzled) layer_forward_hh<4, true, 4>(X, ldx, Ws, bs, Zs, Ys, ldy, nrows, residual, part);
        else layer_forward_hh<4, false, 4>(X, ldx, Ws, bs, Zs, Ys, ldy, nrows, residual, part);
        return;
    }
    if (swizzled) {  // hidden -> hidden layers staged by TMA (hidden >= 32)
        switch (hidden >> 4) {
            case 2: layer_forward_t<2, true>(X, ldx, K4, Ws, bs, Zs, Ys, ldy, nrows, residual, part); break;
            case 4: layer_forward_t<4, true>(X, ldx, K4, Ws, bs, Zs, Ys, ldy, nrows, residual, part); break;
            default: layer_forward_t<8, true>(X, ldx, K4, Ws, bs, Zs, Ys, ldy, nrows, residual, part); break;
        }
        return;
    }
    switch (hidden >> 4) {
        case 1: layer_forward_t<1, false>(X, ldx, K4, Ws, bs, Zs, Ys, ldy, nrows, residual, part); break;
        case 2: layer_forward_t<2, false>(X, ldx, K4, Ws, bs, Zs, Ys, ldy, nrows, residual, part); break;
        case 4: layer_forward_t<4, false>(X, ldx, K4, Ws, bs, Zs, Ys, ldy, nrows, residual, part); break;
        default: layer_forward_t<8, false>(X, ldx, K4, Ws, bs, Zs, Ys, ldy, nrows, residual, part); break;
    }
}

// ---------------------------------------------------------------- input gradient
template <int CM>
__device__ __forceinline__ void load_cm(const float *p, float (&w)[CM]) {
    if constexpr (CM == 1) {
        w[0] = p[0];
    } else if constexpr (CM == 2) {
        const float2 v = *reinterpret_cast<const float2 *>(p);
        w[0] = v.x; w[1] = v.y;
    } else {
#pragma unroll
        for (int q = 0; q < CM / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4 *>(p + 4 * q);
            w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
        }
    }
}

// dX = dZ . W (+ dY when the block was residual) for a hidden -> hidden layer, 16 rows.
// W is the forward staging (Ws[j * ldw + k], ldw = H + 4): thread (cg, rg, ks) owns the CM
// consecutive columns k = CM*cg.. of rows {2rg, 2rg+1} over the ks-th quarter of j.
// Zprev != nullptr: the epilogue also writes the NEXT backward step's dZ = dX * gelu'(Zprev) (rows >= valid_rows zero)
// into dZprev — which may be the buffer dY lives in: every thread reads its dY element before it writes there.  That
// is gelu_backward() of the layer below without its own pass over the tile and its two barriers; same expression.
template <int CM, bool SWZ>
__device__ __noinline__ void layer_input_grad_t(const float *dZ, int ld, const float *Ws, const float *dY, float *dX,
                                                bool residual, float *part, const float *Zprev, float *dZprev,
                                                int valid_rows) {
    ASAC_SMEM(dZ); ASAC_SMEM(Ws); ASAC_SMEM(dY); ASAC_SMEM(dX); ASAC_SMEM(part);
    constexpr int H = 16 * CM;
    const int tid = threadIdx.x, cg = tid & 15, rg = (tid >> 4) & 7, ks = tid >> 7;
    const int ldw = H + 4;
    const int jb = ks * (H / KSPLIT), je = jb + H / KSPLIT;
    const int r = 2 * rg;
    float acc[2][CM];
#pragma unroll
    for (int c = 0; c < CM; ++c) acc[0][c] = acc[1][c] = 0.f;
#pragma unroll
    for (int j = jb; j < je; j += 4) {
        const float4 d0 = *reinterpret_cast<const float4 *>(dZ + r * ld + j);
        const float4 d1 = *reinterpret_cast<const float4 *>(dZ + (r + 1) * ld + j);
        const float d0v[4] = {d0.x, d0.y, d0.z, d0.w}, d1v[4] = {d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            float w[CM];
            if constexpr (SWZ) {  // columns CM*cg .. of row j+jj: whole 16-byte chunks (CM >= 4) or half a chunk (CM == 2)
                const int row = j + jj;
                if constexpr (CM >= 4) {
#pragma unroll
                    for (int q = 0; q < CM / 4; ++q) {
                        const float4 v = *reinterpret_cast<const float4 *>(Ws + swz_off(H, row, CM * cg + 4 * q));
                        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
                    }
                } else {
                    load_cm<CM>(Ws + swz_off(H, row, CM * cg), w);
                }
            } else {
                load_cm<CM>(Ws + (j + jj) * ldw + CM * cg, w);
            }
#pragma unroll
            for (int c = 0; c < CM; ++c) {
                acc[0][c] = fmaf(d0v[jj], w[c], acc[0][c]);
                acc[1][c] = fmaf(d1v[jj], w[c], acc[1][c]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < CM; ++c) part[(ks * PASS_ROWS + r + i) * H + CM * cg + c] = acc[i][c];
    __syncthreads();
#pragma unroll 1
    for (int o = tid; o < PASS_ROWS * H; o += NT) {
        const int row = o / H, k = o - row * H;
        float v = part[o];
#pragma unroll
        for (int s = 1; s < KSPLIT; ++s) v += part[s * PASS_ROWS * H + o];
        if (residual) v = v + dY[row * ld + k];
        dX[row * ld + k] = v;
        if (Zprev) dZprev[row * ld + k] = row < valid_rows ? v * gelu_erf_grad(Zprev[row * ld + k]) : 0.f;
    }
    __syncthreads();
}

__device__ __forceinline__ void layer_input_grad(int hidden, const float *dZ, int ld, const float *Ws,
                                                 const float *dY, float *dX, bool residual, float *part,
                                                 bool swizzled, const float *Zprev = nullptr, float *dZprev = nullptr,
                                                 int valid_rows = 0) {
    if (swizzled) {
        switch (hidden >> 4) {
            case 2: layer_input_grad_t<2, true>(dZ, ld, Ws, dY, dX, residual, part, Zprev, dZprev, valid_rows); break;
            case 4: layer_input_grad_t<4, true>(dZ, ld, Ws, dY, dX, residual, part, Zprev, dZprev, valid_rows); break;
            default: layer_input_grad_t<8, true>(dZ, ld, Ws, dY, dX, residual, part, Zprev, dZprev, valid_rows); break;
        }
        return;
    }
    switch (hidden >> 4) {
        case 1: layer_input_grad_t<1, false>(dZ, ld, Ws, dY, dX, residual, part, Zprev, dZprev, valid_rows); break;
        case 2: layer_input_grad_t<2, false>(dZ, ld, Ws, dY, dX, residual, part, Zprev, dZprev, valid_rows); break;
        case 4: layer_input_grad_t<4, false>(dZ, ld, Ws, dY, dX, residual, part, Zprev, dZprev, valid_rows); break;
        default: layer_input_grad_t<8, false>(dZ, ld, Ws, dY, dX, residual, part, Zprev, dZprev, valid_rows); break;
    }
}

// dZ = dY * gelu'(Z) over 16 rows (H a power of two); rows >= valid_rows are zeroed
static __device__ __noinline__ void gelu_backward(const float *dY, const float *Z, float *dZ, int ld, int H, int valid_rows) {
    ASAC_SMEM(dY); ASAC_SMEM(Z); ASAC_SMEM(dZ);
    const int sh = 31 - __clz(H);
#pragma unroll 1
    for (int i = threadIdx.x; i < PASS_ROWS * H; i += NT) {
        const int r = i >> sh, j = i & (H - 1);
        dZ[r * ld + j] = r < valid_rows ? dY[r * ld + j] * gelu_erf_grad(Z[r * ld + j]) : 0.f;
    }
}

// Partial weight / bias gradient of one layer over the tile's rows:
//   gW[j * K + k] = sum_r dZ[r][j] * X[r][k],  gb[j] = sum_r dZ[r][j]
// 2 x 4 output blocks, consecutive threads along k (coalesced 16-byte stores).
static __device__ __noinline__ void layer_weight_grad(const float *dZ, int ldz, const float *X, int ldx, int H, int K,
                                               int nrows, float *gW, float *gb) {
    ASAC_SMEM(dZ); ASAC_SMEM(X);
    const int tid = threadIdx.x;
    const int K4 = round_up(K, 4);
    const int nJB = H >> 1, nKB = K4 >> 2;
    const bool vec = ((K & 3) == 0) && ((((uintptr_t)gW) & 15) == 0);
#pragma unroll 1
    for (int blk = tid; blk < nJB * nKB; blk += NT) {
        const int jb = blk / nKB, kb = blk - jb * nKB;
        float acc[2][4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
        for (int r = 0; r < nrows; ++r) {
            const float2 dz = *reinterpret_cast<const float2 *>(dZ + r * ldz + 2 * jb);
            const float4 x = *reinterpret_cast<const float4 *>(X + r * ldx + 4 * kb);
            const float dzv[2] = {dz.x, dz.y};
            const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(dzv[a], xv[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            float *dst = gW + (int64_t)(2 * jb + a) * K + 4 * kb;
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

// Linear head: out[r * O + o] = X[r] . Wh[o] + bh[o]   (Wh, bh in shared or global memory, 8 lanes per dot)
static __device__ __noinline__ void head_forward(const float *X, int ldx, int H, const float *Wh, const float *bh, int O,
                                          int nrows, float *out) {
    ASAC_SMEM(X); ASAC_SMEM(out);
    const int tid = threadIdx.x, grp = tid >> 3, sub = tid & 7;
    const int total = nrows * O;
#pragma unroll 1
    for (int d0 = 0; d0 < total; d0 += NT / 8) {
        const int d = d0 + grp;
        float s = 0.f;
        if (d < total) {
            const int r = d / O, o = d - r * O;
            const float *x = X + r * ldx;
            const float *w = Wh + (int64_t)o * H;
            for (int k = sub; k < H; k += 8) s = fmaf(x[k], w[k], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (d < total && sub == 0) out[d] = s + bh[d % O];
    }
}

// Head backward.  dO[r * O + o] (rows >= valid rows must be zero):
//   gWh[o * H + j] = sum_r dO[r][o] X[r][j];  gbh[o] = sum_r dO[r][o];  dH[r][j] = sum_o dO[r][o] Wh[o][j]
// Zlast != nullptr: also dZlast = dH * gelu'(Zlast) for the last ResBlock (rows >= valid_rows zero) — gelu_backward()
// of that layer without its own pass and barriers; same expression.
static __device__ __noinline__ void head_backward(const float *dO, int O, const float *X, int ldx, int H, const float *Wh,
                                           int nrows, float *gWh, float *gbh, float *dH, int ldh,
                                           const float *Zlast = nullptr, float *dZlast = nullptr, int valid_rows = 0) {
    ASAC_SMEM(dO); ASAC_SMEM(dH);
    if (X) ASAC_SMEM(X);
    const int tid = threadIdx.x;
    if (gWh) {
#pragma unroll 1
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
#pragma unroll 1
    for (int i = tid; i < nrows * H; i += NT) {
        const int r = i / H, j = i - r * H;
        float s = 0.f;
        for (int o = 0; o < O; ++o) s = fmaf(dO[r * O + o], Wh[(int64_t)o * H + j], s);
        dH[r * ldh + j] = s;
        if (Zlast) dZlast[r * ldh + j] = r < valid_rows ? s * gelu_erf_grad(Zlast[r * ldh + j]) : 0.f;
    }
}

// ---------------------------------------------------------------- whole-net forward
__host__ __device__ __forceinline__ int tile_lda(int hidden, int in_dim) {
    const int k4 = round_up(in_dim, 4);
    return (hidden > k4 ? hidden : k4) + 4;
}
__host__ __device__ __forceinline__ int tile_wsz(int hidden, int in_dim) {
    const int a = weight_slot_floats(hidden, in_dim), b = weight_slot_floats(hidden, hidden);
    return a > b ? a : b;
}
__host__ __device__ __forceinline__ int tile_part_floats(int hidden) { return KSPLIT * PASS_ROWS * hidden; }

// copies a net's head (Whead[O, H] followed by bhead[O], contiguous in the flat layout) into shared
// memory at kernel start, so that the head passes do not pay a global round trip on the critical path
__device__ __forceinline__ void stage_head(float *dst, const NetShape &s, const float *params) {
    const float *src = params + net_w_off(s, s.depth);
    const int n = s.out_dim * s.hidden + s.out_dim;
    for (int i = threadIdx.x; i < n; i += NT) dst[i] = __ldg(src + i);
}
__host__ __device__ __forceinline__ int head_floats(int hidden, int out_dim) { return round_up(out_dim * hidden + out_dim, 4); }

// appends the `depth` trunk layers of a stock net to a job table (forward order)
// `map` (may be null): 4-D tensor map of the net family's hidden -> hidden layers, `net` the member index
__device__ __forceinline__ bool tma_layer(const NetShape &s, int l, const CUtensorMap *map) {
    return map != nullptr && l >= 1 && s.hidden >= 32;
}
// `lane` >= 0: called by a whole warp, lane n + l writes job n + l (thread 0 alone spent 1.9 us of every kernel's
// setup building the table); lane < 0: the calling thread writes them all
__device__ __forceinline__ int push_trunk_jobs(WeightJob *jobs, int n, const NetShape &s, const float *params,
                                               const CUtensorMap *map = nullptr, int net = 0, int lane = -1) {
    for (int l = 0; l < s.depth; ++l, ++n)
        if (lane < 0 || lane == n)
            jobs[n] = WeightJob{params + net_w_off(s, l), params + net_b_off(s, l), s.hidden, net_k(s, l),
                                tma_layer(s, l, map) ? map : nullptr, l - 1, net};
    return n;
}
// appends layers depth-1 .. 1 (the input-gradient passes of a backward walk; no bias needed)
__device__ __forceinline__ int push_trunk_jobs_reverse(WeightJob *jobs, int n, const NetShape &s, const float *params,
                                                       const CUtensorMap *map = nullptr, int net = 0, int lane = -1) {
    for (int l = s.depth - 1; l >= 1; --l, ++n)
        if (lane < 0 || lane == n)
            jobs[n] = WeightJob{params + net_w_off(s, l), nullptr, s.hidden, s.hidden,
                                tma_layer(s, l, map) ? map : nullptr, l - 1, net};
    return n;
}
// The same tables built by ONE WARP, lane j writing job j in closed form: a kernel's table is a sequence of up to
// four segments (a net's trunk in forward order, or its layers depth-1 .. 1 for the input-gradient walk).
struct JobSegment {
    const float *params;
    const CUtensorMap *map;
    NetShape s;
    int net;
    int reverse;  // 0: layers 0 .. depth-1 with bias; 1: layers depth-1 .. 1 without
};
__device__ __forceinline__ int segment_jobs(const JobSegment &g) { return g.reverse ? g.s.depth - 1 : g.s.depth; }
__device__ __forceinline__ void write_segment_job(WeightJob *jobs, int lane, int first, const JobSegment &g) {
    const int i = lane - first;
    if (i < 0 || i >= segment_jobs(g)) return;
    const int l = g.reverse ? g.s.depth - 1 - i : i;
    jobs[lane] = WeightJob{g.params + net_w_off(g.s, l), g.reverse ? nullptr : g.params + net_b_off(g.s, l), g.s.hidden,
                           g.reverse ? g.s.hidden : net_k(g.s, l), tma_layer(g.s, l, g.map) ? g.map : nullptr, l - 1, g.net};
}
// lane < 32 of one warp; segments with params == nullptr are skipped
__device__ __forceinline__ void write_job_table(WeightJob *jobs, int lane, const JobSegment &g0, const JobSegment &g1,
                                                const JobSegment &g2, const JobSegment &g3) {
    // pick the lane's segment first (a handful of compares), then build ONE job
    const int n0 = g0.params ? segment_jobs(g0) : 0, n1 = g1.params ? segment_jobs(g1) : 0,
              n2 = g2.params ? segment_jobs(g2) : 0, n3 = g3.params ? segment_jobs(g3) : 0;
    const int f1 = n0, f2 = n0 + n1, f3 = n0 + n1 + n2;
    if (lane < f1) write_segment_job(jobs, lane, 0, g0);
    else if (lane < f2) write_segment_job(jobs, lane, f1, g1);
    else if (lane < f3) write_segment_job(jobs, lane, f2, g2);
    else if (lane < f3 + n3) write_segment_job(jobs, lane, f3, g3);
}
// descriptor fetch of a TMA tensor map ahead of its first copy (kernel entry, one thread)
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// Per-layer activation buffers of a saved forward pass: layer l at base + l * stride.  (Pointer arrays indexed by the
// layer live in local memory: filling and reading them cost ~300 instructions of every backward kernel's setup.)
struct LayerBufs {
    float *base;
    int stride;
    __device__ LayerBufs(float *b, int s) : base(b), stride(s) {}
    __device__ LayerBufs(decltype(nullptr)) : base(nullptr), stride(0) {}
    __device__ __forceinline__ float *operator[](int l) const { return base + l * stride; }
};

// Runs the `depth` ResBlocks of a stock net over rows held in `x0` (nrows x lda, input columns
// [0, in_dim) valid, pad columns up to round_up(in,4) zero); the pipe's next `depth` jobs must be
// this net's trunk layers in forward order.
//   save == nullptr : ping-pongs between bufA and bufB, returns the buffer holding the output
//   save != nullptr : layer l reads save_x[l] and writes z to save_z[l], y to save_x[l+1]
__device__ __forceinline__ float *net_trunk_forward(const NetShape &s, WeightPipe &pipe, float *x0, float *bufA,
                                                    float *bufB, LayerBufs save_x, LayerBufs save_z, int lda, int nrows,
                                                    float *part) {
    const int H = s.hidden;
    float *x = x0;
#pragma unroll 1
    for (int l = 0; l < s.depth; ++l) {
        const float *Ws, *bs;
        const bool swz = pipe_front_swizzled(pipe);
        pipe_acquire(pipe, Ws, bs);
        const int K = net_k(s, l);
        float *y = save_x.base ? save_x[l + 1] : (x == bufA ? bufB : bufA);
        layer_forward(H, x, lda, round_up(K, 4), Ws, bs, save_z.base ? save_z[l] : nullptr, y, lda, nrows, K == H, part,
                      swz);
        pipe_release(pipe);  // layer_forward ends with a CTA barrier
        x = y;
    }
    return x;
}

}  // namespace asac
